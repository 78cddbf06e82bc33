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
// Pulse compression: one overlap-save block of length L = R1*R2*R3 handled by a GROUP of T = L/16
// threads (every thread owns 16 points in every pass), 256/T groups per CTA.
//   x[i] = y[s0 + i], s0 = seg_start0 + g0 - (taps-1);  c = IFFT(FFT(x) . H);
//   out[g0 + i - (taps-1)] = c[i] for i >= taps-1           (fun_process_single_frame.m:115-125;
//   identical to the reference's full-length FFT convolution because that one never wraps).
// Passes: DIF(R1) fused with the global load -> DIF(R2) -> [DIF(R3) . H . DIT(R3)] in registers ->
// DIT(R2) -> DIT(R1) fused with the global store: 4 shared-memory writes + 4 reads per point.
// =============================================================================================
template <int L_, int R1_, int R2_, int R3_> struct PcCfg {
    static constexpr int L = L_, R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int T = L / 16;             // threads per FFT
    static constexpr int NG = RSP_PC_THREADS / T;   // FFTs per CTA
    static constexpr int SPAN1 = L / R1;         // pass-1 butterfly stride
    static constexpr int LS2 = L / R1;           // pass-2 sub-transform length (= R2*R3)
    static constexpr int SPAN2 = LS2 / R2;       // = R3
    static constexpr int NB1 = 16 / R1, NB2 = 16 / R2, NB3 = 16 / R3;   // butterflies per thread
    static constexpr int SMEM_ELEMS = L + L / 16 + 16;                   // padded line
    static_assert(R1 * R2 * R3 == L, "radices must multiply to L");
    static_assert(T >= 32 && RSP_PC_THREADS % T == 0, "group must be whole warps");
};

struct PcBlockArgs {
    const cf* line;      // beam line y[0..N)
    cf* out_line;        // pc line [0..G)
    const cf* tw1;       // [(R1-1)][L/R1]      e^{-2 pi i jk/L}          (global)
    const cf* tw2;       // [(R2-1)][R3]        e^{-2 pi i jk/(R2 R3)}    (shared copy in the kernel)
    const cf* Hmid;      // [(NB3*R3)][T]: Hmid[(i*R3+k)*T + t] = H_dr[R3*(t + T*i) + k] / L
    int N;               // samples per line
    int seg_start0;      // 0-based first sample of the segment (samples before it count as zero)
    int taps;            // matched-filter length
    int g0;              // first output gate of this block
    int g_end;           // one past the last gate this segment owns
};

template <class Cfg> RSP_HD void pc_phase_load_pass1(const PcBlockArgs& a, cf* s, int t) {
    const int s0 = a.seg_start0 + a.g0 - (a.taps - 1);
#pragma unroll
    for (int i = 0; i < Cfg::NB1; ++i) {
        const int q = t + i * Cfg::T;
        cf v[Cfg::R1];
#pragma unroll
        for (int m = 0; m < Cfg::R1; ++m) {
            const int idx = s0 + q + m * Cfg::SPAN1;
            v[m] = (idx >= a.seg_start0 && idx < a.N) ? a.line[idx] : make_float2(0.f, 0.f);
        }
        SmallDft<Cfg::R1, -1>::run(v);
        s[rsp_pad16(q)] = v[0];
#pragma unroll
        for (int k = 1; k < Cfg::R1; ++k) {
            const cf w = a.tw1[(k - 1) * Cfg::SPAN1 + q];
            s[rsp_pad16(q + k * Cfg::SPAN1)] = mul_tw<-1>(v[k], w.x, w.y);
        }
    }
}

template <class Cfg> RSP_HD void pc_phase_pass2(const PcBlockArgs& a, cf* s, int t) {
#pragma unroll
    for (int i = 0; i < Cfg::NB2; ++i) dif_butterfly<Cfg::R2, -1>(s, Cfg::LS2, t + i * Cfg::T, a.tw2, PadAddr());
}

// last forward pass (Ls = R3, no twiddles) . H . first inverse pass, all in registers
template <class Cfg> RSP_HD void pc_phase_mid(const PcBlockArgs& a, cf* s, int t) {
#pragma unroll
    for (int i = 0; i < Cfg::NB3; ++i) {
        const int q = t + i * Cfg::T;
        cf v[Cfg::R3];
#pragma unroll
        for (int m = 0; m < Cfg::R3; ++m) v[m] = s[rsp_pad16(Cfg::R3 * q + m)];
        SmallDft<Cfg::R3, -1>::run(v);
#pragma unroll
        for (int k = 0; k < Cfg::R3; ++k) v[k] = cmul(v[k], a.Hmid[(i * Cfg::R3 + k) * Cfg::T + t]);
        SmallDft<Cfg::R3, +1>::run(v);
#pragma unroll
        for (int m = 0; m < Cfg::R3; ++m) s[rsp_pad16(Cfg::R3 * q + m)] = v[m];
    }
}

template <class Cfg> RSP_HD void pc_phase_ipass2(const PcBlockArgs& a, cf* s, int t) {
#pragma unroll
    for (int i = 0; i < Cfg::NB2; ++i) dit_butterfly<Cfg::R2, +1>(s, Cfg::LS2, t + i * Cfg::T, a.tw2, PadAddr());
}

template <class Cfg> RSP_HD void pc_phase_ipass1_store(const PcBlockArgs& a, const cf* s, int t) {
#pragma unroll
    for (int i = 0; i < Cfg::NB1; ++i) {
        const int q = t + i * Cfg::T;
        cf v[Cfg::R1];
        v[0] = s[rsp_pad16(q)];
#pragma unroll
        for (int k = 1; k < Cfg::R1; ++k) {
            const cf w = a.tw1[(k - 1) * Cfg::SPAN1 + q];
            const cf x = s[rsp_pad16(q + k * Cfg::SPAN1)];
            v[k] = mul_tw<+1>(x, w.x, w.y);
        }
        SmallDft<Cfg::R1, +1>::run(v);
#pragma unroll
        for (int m = 0; m < Cfg::R1; ++m) {
            const int io = q + m * Cfg::SPAN1;
            const int g = a.g0 + io - (a.taps - 1);
            if (io >= a.taps - 1 && g < a.g_end) a.out_line[g] = v[m];
        }
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
// MTD: windowed length-P Doppler FFT for a tile of 32 range gates, in place in shared memory.
//   tile element (position a, gate gl) lives at s[a*33 + gl]; lanes run along the gates, so every
//   pass is bank-conflict free and all twiddles are warp-uniform.
//   Input pulse p is stored at position perm[p] (digit reversal) so that the DIT passes leave
//   X[f] at position f; fftshift is folded into the window as (-1)^p (even P).
//   Compile-time plan: P = R0*R1*R2 (DIF order); DIT runs R2, R1, R0.
// =============================================================================================
#define RSP_MTD_TG 32
struct StrideAddr {
    int gl;
    RSP_HD int operator()(int a) const { return a * (RSP_MTD_TG + 1) + gl; }
};

template <int P_, int R0_, int R1_, int R2_> struct MtdCfg {
    static constexpr int P = P_, R0 = R0_, R1 = R1_, R2 = R2_;
    static_assert(R0 * R1 * R2 == P, "radices must multiply to P");
    // twiddle table offsets (pass s has (r_s - 1) * (Ls_s / r_s) entries)
    static constexpr int TW0 = 0;
    static constexpr int TW1 = TW0 + (R0 - 1) * (P / R0);
    static constexpr int TW2 = TW1 + (R1 - 1) * (P / R0 / R1);
    static constexpr int TW_COUNT = TW2 + (R2 - 1) * 1;
};

template <int R, int LS, int P> RSP_HD void mtd_dit_pass_t(cf* s, const cf* tw, int tid) {
    if (R == 1) return;
    constexpr int NBF = P / R;                        // butterflies per gate
    const int gl = tid & (RSP_MTD_TG - 1);
    StrideAddr addr;
    addr.gl = gl;
#pragma unroll
    for (int q = tid / RSP_MTD_TG; q < NBF; q += RSP_MTD_THREADS / RSP_MTD_TG)
        dit_butterfly<(R > 1 ? R : 2), -1>(s, LS, q, tw, addr);
}

template <class Cfg> RSP_HD void mtd_passes_phase(cf* s, const cf* tw, int tid, int pass) {
    // pass 0 = innermost (radix R2, Ls = R2), 1 = middle (R1, Ls = R1*R2), 2 = outermost (R0, Ls = P)
    if (pass == 0) mtd_dit_pass_t<Cfg::R2, Cfg::R2, Cfg::P>(s, tw + Cfg::TW2, tid);
    else if (pass == 1) mtd_dit_pass_t<Cfg::R1, Cfg::R1 * Cfg::R2, Cfg::P>(s, tw + Cfg::TW1, tid);
    else mtd_dit_pass_t<Cfg::R0, Cfg::P, Cfg::P>(s, tw + Cfg::TW0, tid);
}

// =============================================================================================
// CFAR on one (pair, gate tile): S tile rows = gates [g_first - mR, g_first + TG + mR), P columns.
//   fun_process_single_frame.m:192-213.  Two phases: window sums, then the decision.
//     R5[row][v] = sum_{i<ref_r} S[row+i][v]      rows [0, TG + mR + guard_r + 1)
//     D5[gl][v]  = sum_{i<ref_v} S[gl+mR][v+i]    CUT rows, v in [0, P - ref_v]
//   CUT (gl, v): lead_r = R5[gl][v], trail_r = R5[gl + mR + guard_r + 1][v],
//                lead_v = D5[gl][v - mV], trail_v = D5[gl][v + guard_v + 1].
//   The sums add left to right exactly like the reference's mean(); no running (subtractive) sums.
// =============================================================================================
struct CfarParams {
    int P, G;
    int guard_r, guard_v, ref_r, ref_v;
    float t_cfar;
};

RSP_HD int cfar_r5_rows(const CfarParams& c, int TG) { return TG + c.guard_r + c.ref_r + c.guard_r + 1; }

RSP_HD void cfar_sums_phase(const float* S, float* R5, float* D5, const CfarParams& c, int TG, int tid, int nthreads) {
    const int P = c.P, mR = c.guard_r + c.ref_r;
    const int nr5 = cfar_r5_rows(c, TG) * P;
    for (int e = tid; e < nr5; e += nthreads) {
        float acc = S[e];
        for (int i = 1; i < c.ref_r; ++i) acc += S[e + i * P];
        R5[e] = acc;
    }
    const int nd5 = TG * P;
    for (int e = tid; e < nd5; e += nthreads) {
        const int gl = e / P, v = e - gl * P;
        float acc = 0.f;
        if (v + c.ref_v <= P) {
            const float* row = S + (gl + mR) * P + v;
            acc = row[0];
            for (int i = 1; i < c.ref_v; ++i) acc += row[i];
        }
        D5[e] = acc;
    }
}

RSP_HD int cfar_decide(const float* S, const float* R5, const float* D5, const CfarParams& c, int gl, int v,
                       float* cut_out) {
    const int P = c.P, mR = c.guard_r + c.ref_r, mV = c.guard_v + c.ref_v;
    const float cut = S[(gl + mR) * P + v];
    const float lead_r = R5[gl * P + v], trail_r = R5[(gl + mR + c.guard_r + 1) * P + v];
    const float lead_v = D5[gl * P + v - mV], trail_v = D5[gl * P + v + c.guard_v + 1];
    const float noise_r = fmaxf(lead_r / (float)c.ref_r, trail_r / (float)c.ref_r);
    const float noise_v = fmaxf(lead_v / (float)c.ref_v, trail_v / (float)c.ref_v);
    *cut_out = cut;
    return cut > c.t_cfar * fmaxf(noise_r, noise_v);
}
