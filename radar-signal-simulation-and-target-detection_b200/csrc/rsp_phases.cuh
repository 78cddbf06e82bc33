// rsp_phases.cuh -- the per-thread bodies of the pulse-compression, MTD and CFAR kernels,
// written as host/device "phases".  A kernel is `phase; __syncthreads(); phase; ...`; the
// host-emulation test (csrc/host_emul.cpp, built with g++, no GPU) runs the very same phases with
// `for (tid = 0; tid < nthreads; ++tid)` loops in place of the barriers and checks them against
// NumPy.  Inside one phase every butterfly touches a disjoint set of elements, so the serial
// emulation is exact.
#pragma once
#include "rsp_math.cuh"

#define RSP_PC_THREADS 256
#define RSP_MTD_THREADS 256
#define RSP_CFAR_THREADS 256

struct PadAddr {
    RSP_HD int operator()(int a) const { return rsp_pad16(a); }
};

// =============================================================================================
// Pulse compression: one overlap-save block of length L = R1*256 (radices R1,16,16).
//   x[i] = y[s0 + i], s0 = seg_start0 + g0 - (taps-1);  c = IFFT(FFT(x) . H);
//   out[g0 + i - (taps-1)] = c[i] for i >= taps-1           (fun_process_single_frame.m:115-125;
//   identical to the reference's full-length FFT convolution because that one never wraps).
// =============================================================================================
struct PcBlockArgs {
    const cf* line;      // beam line y[0..N)
    cf* out_line;        // pc line [0..G)
    const cf* tw1;       // [(R1-1)][256]  e^{-2 pi i jk/L}
    const cf* tw2;       // [15][16]       e^{-2 pi i jk/256}
    const cf* H;         // [L] filter spectrum / L, digit-reversed order of radices (R1,16,16)
    int N;               // samples per line
    int seg_start0;      // 0-based first sample of the segment (samples before it count as zero)
    int taps;            // matched-filter length
    int g0;              // first output gate of this block
    int g_end;           // one past the last gate this segment owns
};

template <int R1> RSP_HD void pc_phase_load_pass1(const PcBlockArgs& a, cf* s, int tid) {
    const int s0 = a.seg_start0 + a.g0 - (a.taps - 1);
    cf v[R1];
#pragma unroll
    for (int m = 0; m < R1; ++m) {
        const int idx = s0 + tid + m * 256;
        v[m] = (idx >= a.seg_start0 && idx < a.N) ? a.line[idx] : make_float2(0.f, 0.f);
    }
    SmallDft<R1, -1>::run(v);
    s[rsp_pad16(tid)] = v[0];
#pragma unroll
    for (int k = 1; k < R1; ++k) {
        const cf w = a.tw1[(k - 1) * 256 + tid];
        s[rsp_pad16(tid + k * 256)] = mul_tw<-1>(v[k], w.x, w.y);
    }
}

template <int R1> RSP_HD void pc_phase_pass2(const PcBlockArgs& a, cf* s, int tid) {
    for (int q = tid; q < R1 * 16; q += RSP_PC_THREADS) dif_butterfly<16, -1>(s, 256, q, a.tw2, PadAddr());
}

// last forward pass (Ls = 16, no twiddles) . H . first inverse pass, all in registers
template <int R1> RSP_HD void pc_phase_mid(const PcBlockArgs& a, cf* s, int tid) {
    for (int q = tid; q < R1 * 16; q += RSP_PC_THREADS) {
        cf v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = s[rsp_pad16(16 * q + m)];
        SmallDft<16, -1>::run(v);
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = cmul(v[k], a.H[16 * q + k]);
        SmallDft<16, +1>::run(v);
#pragma unroll
        for (int m = 0; m < 16; ++m) s[rsp_pad16(16 * q + m)] = v[m];
    }
}

template <int R1> RSP_HD void pc_phase_ipass2(const PcBlockArgs& a, cf* s, int tid) {
    for (int q = tid; q < R1 * 16; q += RSP_PC_THREADS) dit_butterfly<16, +1>(s, 256, q, a.tw2, PadAddr());
}

template <int R1> RSP_HD void pc_phase_ipass1_store(const PcBlockArgs& a, const cf* s, int tid) {
    cf v[R1];
    v[0] = s[rsp_pad16(tid)];
#pragma unroll
    for (int k = 1; k < R1; ++k) {
        const cf w = a.tw1[(k - 1) * 256 + tid];
        const cf x = s[rsp_pad16(tid + k * 256)];
        v[k] = mul_tw<+1>(x, w.x, w.y);
    }
    SmallDft<R1, +1>::run(v);
#pragma unroll
    for (int m = 0; m < R1; ++m) {
        const int i = tid + m * 256;
        const int g = a.g0 + i - (a.taps - 1);
        if (i >= a.taps - 1 && g < a.g_end) a.out_line[g] = v[m];
    }
}

// Narrow-pulse FIR + circshift (fun_process_single_frame.m:111-112,123):
//   u = filter(fir, 1, y(seg_start:end));  piece1(g) = u((g + fir_delay) mod Lseg)
RSP_HD cf pc_narrow_gate(const cf* line, int N, int seg_start0, const float* fir, int nfir, int fir_delay, int g) {
    const int Lseg = N - seg_start0;
    int ui = g + fir_delay;
    ui = ui % Lseg;
    cf acc = make_float2(0.f, 0.f);
    for (int k = 0; k < nfir; ++k) {
        const int i = ui - k;
        if (i < 0) break;
        const cf x = line[seg_start0 + i];
        acc.x += fir[k] * x.x;
        acc.y += fir[k] * x.y;
    }
    return acc;
}

// =============================================================================================
// MTD: windowed length-P Doppler FFT for a tile of TG range gates, in place in shared memory.
//   tile element (position a, gate gl) lives at s[a*(TG+1) + gl].
//   Input pulse p is stored at position perm[p] (digit reversal) so that the DIT passes leave
//   X[f] at position f; fftshift is folded into the window as (-1)^p (even P).
// =============================================================================================
struct MtdPlan {
    int P;
    int nrad;
    int radices[4];      // DIF order r0..r(k-1); DIT runs them k-1 .. 0
    int tw_off[4];       // offset of pass s in the twiddle table
};

struct StrideAddr {
    int stride, gl;
    RSP_HD int operator()(int a) const { return a * stride + gl; }
};

template <int SIGN>
RSP_HD void mtd_dit_pass(cf* s, const MtdPlan& plan, int pass, const cf* tw_all, int TG, int tid, int nthreads) {
    const int r = plan.radices[pass];
    int Ls = plan.P;
    for (int i = 0; i < pass; ++i) Ls /= plan.radices[i];
    const int nbf = plan.P / r;
    const cf* tw = tw_all + plan.tw_off[pass];
    for (int e = tid; e < nbf * TG; e += nthreads) {
        StrideAddr addr;
        addr.stride = TG + 1;
        addr.gl = e % TG;
        const int q = e / TG;
        switch (r) {
            case 16: dit_butterfly<16, SIGN>(s, Ls, q, tw, addr); break;
            case 8: dit_butterfly<8, SIGN>(s, Ls, q, tw, addr); break;
            case 4: dit_butterfly<4, SIGN>(s, Ls, q, tw, addr); break;
            default: dit_butterfly<2, SIGN>(s, Ls, q, tw, addr); break;
        }
    }
}

// =============================================================================================
// CFAR on one (pair, gate tile): S tile rows = gates [g_first - mR, g_first + TG + mR), P columns.
//   fun_process_single_frame.m:192-213.  Returns 1 when the CUT (gl, v) is a detection.
// =============================================================================================
struct CfarParams {
    int P, G;
    int guard_r, guard_v, ref_r, ref_v;
    float t_cfar;
};

RSP_HD int cfar_cut(const float* S, int ld, const CfarParams& c, int gl, int v, float* cut_out) {
    const int mR = c.guard_r + c.ref_r;
    const int mV = c.guard_v + c.ref_v;
    const float* row = S + (gl + mR) * ld;
    const float cut = row[v];
    float lead_r = 0.f, trail_r = 0.f, lead_v = 0.f, trail_v = 0.f;
    for (int i = 0; i < c.ref_r; ++i) {
        lead_r += S[(gl + i) * ld + v];
        trail_r += S[(gl + mR + c.guard_r + 1 + i) * ld + v];
    }
    for (int i = 0; i < c.ref_v; ++i) {
        lead_v += row[v - mV + i];
        trail_v += row[v + c.guard_v + 1 + i];
    }
    const float noise_r = fmaxf(lead_r / (float)c.ref_r, trail_r / (float)c.ref_r);
    const float noise_v = fmaxf(lead_v / (float)c.ref_v, trail_v / (float)c.ref_v);
    const float thr = c.t_cfar * fmaxf(noise_r, noise_v);
    *cut_out = cut;
    return cut > thr;
}
