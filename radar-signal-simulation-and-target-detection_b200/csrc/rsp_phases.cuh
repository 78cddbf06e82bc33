// rsp_phases.cuh -- the per-thread bodies of the pulse-compression, MTD and CFAR kernels,
// written as host/device "phases".  A kernel is `phase; __syncthreads(); phase; ...`; the
// host-emulation test (csrc/host_emul.cpp, built with g++, no GPU) runs the very same phases with
// `for (tid = 0; tid < nthreads; ++tid)` loops in place of the barriers and checks them against
// NumPy.  Inside one phase every butterfly touches a disjoint set of elements, so the serial
// emulation is exact.
#pragma once
#include "rsp_math.cuh"
#include "rsp_dft_big.cuh"

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
    int in_lo, in_hi;    // input samples outside [in_lo, in_hi) count as zero (segment start .. line end)
    int seg_start0;      // output gate g reads y[seg_start0 + g - k], k < taps
    int taps;            // matched-filter length
    int g0;              // first output gate of this block
    int g_end;           // one past the last gate this segment owns
};

// Padded shared-memory addressing: pad16(base + off) == pad16(base) + off + (off >> 4) whenever
// (base & 15) + (off & 15) < 16, which holds for every access below (strides are multiples of 16, or
// base & 15 < 4 with strides of 4).  So each thread computes one padded base per butterfly and all
// other addresses are compile-time immediates.
#define RSP_POFF(off) ((off) + ((off) >> 4))

// Twiddles w[k] = W^(k j), k = 1 .. R-1, of one butterfly from the table row entries tw[(k-1)*stride].
// RSP_PC_DERIVE_TW: only the powers k = 1, 2, 4, 8 are loaded, the others are products of those (at most three
// factors, so the error stays at a few ulp).  Loads cost shared-memory / L1 wavefronts, which is what bounds the
// chain; the extra complex multiplies go to the FP32 pipe, which has room.
#ifndef RSP_PC_DERIVE_TW
#define RSP_PC_DERIVE_TW 1
#endif
template <int R> RSP_HD void pc_twiddles(const cf* tw, int stride, cf* w) {
#if RSP_PC_DERIVE_TW
#pragma unroll
    for (int k = 1; k < R; k <<= 1) w[k] = tw[(k - 1) * stride];
#pragma unroll
    for (int k = 3; k < R; ++k) {
        const int low = k & (-k);                      // lowest set bit: k = (k - low) + low, both already known
        if (low != k) w[k] = cmul(w[k - low], w[low]);
    }
#else
#pragma unroll
    for (int k = 1; k < R; ++k) w[k] = tw[(k - 1) * stride];
#endif
}

template <class Cfg> RSP_HD void pc_phase_load_pass1(const PcBlockArgs& a, cf* s, int t) {
    const int s0 = a.seg_start0 + a.g0 - (a.taps - 1);
    const bool interior = s0 >= a.in_lo && s0 + Cfg::L <= a.in_hi;       // uniform over the group
#pragma unroll
    for (int i = 0; i < Cfg::NB1; ++i) {
        const int q = t + i * Cfg::T;
        cf v[Cfg::R1];
        if (interior) {
            const cf* src = a.line + s0 + q;
#pragma unroll
            for (int m = 0; m < Cfg::R1; ++m) v[m] = src[m * Cfg::SPAN1];
        } else {
#pragma unroll
            for (int m = 0; m < Cfg::R1; ++m) {
                const int idx = s0 + q + m * Cfg::SPAN1;
                v[m] = (idx >= a.in_lo && idx < a.in_hi) ? a.line[idx] : make_float2(0.f, 0.f);
            }
        }
        cf w[Cfg::R1];
        pc_twiddles<Cfg::R1>(a.tw1 + q, Cfg::SPAN1, w);
        SmallDft<Cfg::R1, -1>::run(v);
        cf* sb = s + rsp_pad16(q);
        sb[0] = v[0];
#pragma unroll
        for (int k = 1; k < Cfg::R1; ++k) sb[RSP_POFF(k * Cfg::SPAN1)] = mul_tw<-1>(v[k], w[k].x, w[k].y);
    }
}

// one radix-R2 DIF / DIT butterfly of pass 2 (sub-transform length LS2, stride SPAN2)
template <class Cfg, int SIGN, bool DIF> RSP_HD void pc_pass2_butterfly(cf* s, const cf* tw2, int q) {
    constexpr int R = Cfg::R2, SPAN = Cfg::SPAN2;
    const int blk = q / SPAN, j = q - blk * SPAN;
    cf* sb = s + rsp_pad16(blk * Cfg::LS2 + j);
    cf v[R], w[R];
    pc_twiddles<R>(tw2 + j, SPAN, w);
    if (DIF) {
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = sb[RSP_POFF(m * SPAN)];
        SmallDft<R, SIGN>::run(v);
        sb[0] = v[0];
#pragma unroll
        for (int k = 1; k < R; ++k) sb[RSP_POFF(k * SPAN)] = mul_tw<SIGN>(v[k], w[k].x, w[k].y);
    } else {
        v[0] = sb[0];
        if (R == 16) {                                   // input twiddles folded into the first butterfly stage (fma_pm)
#pragma unroll
            for (int k = 1; k < R; ++k) v[k] = sb[RSP_POFF(k * SPAN)];
            SmallDft<16, SIGN>::run_twiddled(v, w);
        } else {
#pragma unroll
            for (int k = 1; k < R; ++k) {
                const cf x = sb[RSP_POFF(k * SPAN)];
                v[k] = mul_tw<SIGN>(x, w[k].x, w[k].y);
            }
            SmallDft<R, SIGN>::run(v);
        }
#pragma unroll
        for (int m = 0; m < R; ++m) sb[RSP_POFF(m * SPAN)] = v[m];
    }
}

template <class Cfg> RSP_HD void pc_phase_pass2(const PcBlockArgs& a, cf* s, int t) {
#pragma unroll
    for (int i = 0; i < Cfg::NB2; ++i) pc_pass2_butterfly<Cfg, -1, true>(s, a.tw2, t + i * Cfg::T);
}

// last forward pass (Ls = R3, no twiddles) . H . first inverse pass, all in registers
template <class Cfg> RSP_HD void pc_phase_mid(const PcBlockArgs& a, cf* s, int t) {
#pragma unroll
    for (int i = 0; i < Cfg::NB3; ++i) {
        const int q = t + i * Cfg::T;
        cf* sb = s + rsp_pad16(Cfg::R3 * q);
        const cf* h = a.Hmid + (i * Cfg::R3) * Cfg::T + t;
        cf v[Cfg::R3];
#pragma unroll
        for (int m = 0; m < Cfg::R3; ++m) v[m] = sb[m];
        SmallDft<Cfg::R3, -1>::run(v);
#pragma unroll
        for (int k = 0; k < Cfg::R3; ++k) v[k] = cmul(v[k], h[k * Cfg::T]);
        SmallDft<Cfg::R3, +1>::run(v);
#pragma unroll
        for (int m = 0; m < Cfg::R3; ++m) sb[m] = v[m];
    }
}

template <class Cfg> RSP_HD void pc_phase_ipass2(const PcBlockArgs& a, cf* s, int t) {
#pragma unroll
    for (int i = 0; i < Cfg::NB2; ++i) pc_pass2_butterfly<Cfg, +1, false>(s, a.tw2, t + i * Cfg::T);
}

template <class Cfg> RSP_HD void pc_phase_ipass1_store(const PcBlockArgs& a, const cf* s, int t) {
    // every output of the block is wanted when the block lies inside the segment's gate range
    const bool full = a.g0 + Cfg::L - (a.taps - 1) <= a.g_end;
#pragma unroll
    for (int i = 0; i < Cfg::NB1; ++i) {
        const int q = t + i * Cfg::T;
        const cf* sb = s + rsp_pad16(q);
        cf v[Cfg::R1], w[Cfg::R1];
        pc_twiddles<Cfg::R1>(a.tw1 + q, Cfg::SPAN1, w);
        v[0] = sb[0];
        if (Cfg::R1 == 16) {
#pragma unroll
            for (int k = 1; k < Cfg::R1; ++k) v[k] = sb[RSP_POFF(k * Cfg::SPAN1)];
            SmallDft<16, +1>::run_twiddled(v, w);
        } else {
#pragma unroll
            for (int k = 1; k < Cfg::R1; ++k) {
                const cf x = sb[RSP_POFF(k * Cfg::SPAN1)];
                v[k] = mul_tw<+1>(x, w[k].x, w[k].y);
            }
            SmallDft<Cfg::R1, +1>::run(v);
        }
        cf* dst = a.out_line + a.g0 + q - (a.taps - 1);
#pragma unroll
        for (int m = 0; m < Cfg::R1; ++m) {
            const int io = q + m * Cfg::SPAN1;
            if (io >= a.taps - 1 && (full || a.g0 + io - (a.taps - 1) < a.g_end)) dst[m * Cfg::SPAN1] = v[m];
        }
    }
}

// Narrow-pulse FIR from a shared-memory copy of the line head: ys[i] = y[seg_start0 + i].
// Valid when ngates + fir_delay <= Lseg (no circshift wrap), which the kernel checks.
RSP_HD cf pc_narrow_gate_smem(const cf* ys, const float* fir, int nfir, int fir_delay, int g) {
    const int ui = g + fir_delay;
    const int kmax = ui + 1 < nfir ? ui + 1 : nfir;
    cf acc = make_float2(0.f, 0.f);
    for (int k = 0; k < kmax; ++k) {
        const cf x = ys[ui - k];
        acc.x += fir[k] * x.x;
        acc.y += fir[k] * x.y;
    }
    return acc;
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

template <class Cfg> struct MtdInner {     // radix of the innermost non-trivial pass and its index
    static constexpr int PASS = Cfg::R2 > 1 ? 0 : (Cfg::R1 > 1 ? 1 : 2);
    static constexpr int R = Cfg::R2 > 1 ? Cfg::R2 : (Cfg::R1 > 1 ? Cfg::R1 : Cfg::R0);
};

// Innermost DIT pass straight from global memory.  Butterfly q (Ls = R, stride 1, no twiddles) owns
// positions q*R + k; with the digit reversal of radices (R0,R1,R2) the pulse stored at position
// q*R + k is  base(q) + k * (P/R)  (affine in k), so the digit-reversed placement needs no table:
//   (R0,R1,1): base = q              (R0,R1,R2): base = q / R1 + R0 * (q % R1)        (R0,1,1): base = 0
//   src = pc + b*ldg + g0 (pulse stride = pstride elements); gate_ok masks gates beyond G.
template <class Cfg>
RSP_HD void mtd_first_pass_t(cf* s, const cf* src, size_t pstride, const float* win, bool gate_ok, int tid) {
    constexpr int R = MtdInner<Cfg>::R, P = Cfg::P, NBF = P / R, KSTR = P / R;
    const int gl = tid & (RSP_MTD_TG - 1);
#pragma unroll
    for (int q = tid / RSP_MTD_TG; q < NBF; q += RSP_MTD_THREADS / RSP_MTD_TG) {
        const int base = Cfg::R2 > 1 ? (q / Cfg::R1 + Cfg::R0 * (q % Cfg::R1)) : q;
        const cf* col = src + (size_t)base * pstride + gl;
        const float* wq = win + base;
        const unsigned ps = (unsigned)pstride;                 // P * pstride elements fit 32 bits (the cube is < 16 GB)
        cf v[R];
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = gate_ok ? col[(unsigned)(k * KSTR) * ps] : make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < R; ++k) v[k] = cscale(v[k], wq[k * KSTR]);
        SmallDft<R, -1>::run(v);
        cf* dst = s + (q * R) * (RSP_MTD_TG + 1) + gl;
#pragma unroll
        for (int m = 0; m < R; ++m) dst[m * (RSP_MTD_TG + 1)] = v[m];
    }
}

template <class Cfg> RSP_HD void mtd_passes_phase(cf* s, const cf* tw, int tid, int pass) {
    // pass 0 = innermost (radix R2, Ls = R2), 1 = middle (R1, Ls = R1*R2), 2 = outermost (R0, Ls = P)
    if (pass == 0) mtd_dit_pass_t<Cfg::R2, Cfg::R2, Cfg::P>(s, tw + Cfg::TW2, tid);
    else if (pass == 1) mtd_dit_pass_t<Cfg::R1, Cfg::R1 * Cfg::R2, Cfg::P>(s, tw + Cfg::TW1, tid);
    else mtd_dit_pass_t<Cfg::R0, Cfg::P, Cfg::P>(s, tw + Cfg::TW0, tid);
}

// =============================================================================================
// MTD for any P (the reference's native P = 332 = 4 * 83): P = R * Q with R the largest power of two <= 8
// dividing P.  Work item (k < Q, gate): the R length-Q DFTs of the decimated sequences x[a + R n] share
// their twiddle W_Q^{nk} (direct sums, Q^2 each), then one radix-R butterfly with W_P^{ak} gives the R
// outputs X[k + Q d].  R * Q^2 = P * Q multiply-adds per line instead of P^2.
//   xin[p*(TG+1) + gl] windowed input, xout[row*(TG+1) + gl] fftshifted output, stw[m] = e^{-2 pi i m/P}.
// =============================================================================================
template <int R> RSP_HD void mtd_dft_item(const cf* xin, cf* xout, const cf* stw, int P, int TG, int k, int gl) {
    const int Q = P / R, ld = TG + 1;
    cf acc[R];
#pragma unroll
    for (int a = 0; a < R; ++a) acc[a] = make_float2(0.f, 0.f);
    int idx = 0;                                        // (n k) mod Q
    for (int n = 0; n < Q; ++n) {
        const cf w = stw[R * idx];                      // W_Q^{nk} = W_P^{R (nk mod Q)}
#pragma unroll
        for (int a = 0; a < R; ++a) {
            const cf x = xin[(a + R * n) * ld + gl];
            acc[a].x = fmaf(x.x, w.x, fmaf(-x.y, w.y, acc[a].x));
            acc[a].y = fmaf(x.x, w.y, fmaf(x.y, w.x, acc[a].y));
        }
        idx += k;
        if (idx >= Q) idx -= Q;
    }
#pragma unroll
    for (int a = 1; a < R; ++a) acc[a] = cmul(acc[a], stw[a * k]);     // W_P^{ak}, a k < P
    if (R > 1) SmallDft<(R > 1 ? R : 2), -1>::run(acc);                // over a -> d
    const int half = P / 2;
#pragma unroll
    for (int d = 0; d < R; ++d) {
        int row = k + Q * d + half;                    // fftshift(.,1): bin f moves to (f + floor(P/2)) mod P
        if (row >= P) row -= P;
        xout[row * ld + gl] = acc[d];
    }
}

// The same item for KT consecutive output bins k0 .. k0 + KT - 1 of one gate: every decimated input sample is
// loaded once and used for KT bins (each with its own warp-uniform twiddle), i.e. R + KT shared-memory loads
// per 4 R KT multiply-adds instead of R + 1 per 4 R.  The single-bin item is limited by the shared-memory
// pipe (9 wavefronts per 16 FMA instructions at R = 4); this one is limited by the FMA pipe.  Sums run over n
// in the same order with the same fused operations, so the results are bit-identical to mtd_dft_item.
template <int R, int KT> RSP_HD void mtd_dft_item_kt(const cf* xin, cf* xout, const cf* stw, int P, int TG, int k0, int gl) {
    const int Q = P / R, ld = TG + 1;
    cf acc[KT][R];
    int idx[KT], kk[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) {
        kk[j] = k0 + j < Q ? k0 + j : Q - 1;                // bins past the end repeat the last one and are not stored
        idx[j] = 0;
#pragma unroll
        for (int a = 0; a < R; ++a) acc[j][a] = make_float2(0.f, 0.f);
    }
    for (int n = 0; n < Q; ++n) {
        cf x[R];
#pragma unroll
        for (int a = 0; a < R; ++a) x[a] = xin[(a + R * n) * ld + gl];
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            const cf w = stw[R * idx[j]];                   // W_Q^{n k_j}
#pragma unroll
            for (int a = 0; a < R; ++a) {
                acc[j][a].x = fmaf(x[a].x, w.x, fmaf(-x[a].y, w.y, acc[j][a].x));
                acc[j][a].y = fmaf(x[a].x, w.y, fmaf(x[a].y, w.x, acc[j][a].y));
            }
            idx[j] += kk[j];
            if (idx[j] >= Q) idx[j] -= Q;
        }
    }
    const int half = P / 2;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
        if (k0 + j >= Q) break;
#pragma unroll
        for (int a = 1; a < R; ++a) acc[j][a] = cmul(acc[j][a], stw[a * kk[j]]);
        if (R > 1) SmallDft<(R > 1 ? R : 2), -1>::run(acc[j]);
#pragma unroll
        for (int d = 0; d < R; ++d) {
            int row = kk[j] + Q * d + half;
            if (row >= P) row -= P;
            xout[row * ld + gl] = acc[j][d];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Odd sub-length Q (the native P = 332 = 4 x 83): the length-Q DFTs by their even / odd parts.  With
//   S_n = x[n] + x[Q-n],  D_n = x[n] - x[Q-n]   (n = 1 .. H = (Q-1)/2, formed once per line by mtd_dft_fold_phase),
//   X[k]   = x[0] + sum_n S_n cos(2 pi n k / Q)  -  i sum_n D_n sin(2 pi n k / Q)
//   X[Q-k] = x[0] + sum_n S_n cos(...)           +  i sum_n D_n sin(...)
// a pair of bins costs 4 FMAs per (n, sub-sequence) instead of the 16 of two direct sums: a quarter of the multiply-adds
// (110 k -> 28 k per line at P = 332), which is what the direct kernel was bound by.  KP bin pairs per work item share the
// S / D loads; pair 0 is bin 0 alone.  Same radix-R recombination as mtd_dft_item.
// ---------------------------------------------------------------------------------------------
RSP_HD void mtd_dft_fold_phase(cf* xin, int P, int R, int TG, int tid, int nthreads) {
    const int Q = P / R, H = (Q - 1) / 2, ld = TG + 1;
    const int n_items = R * H * TG;
    for (int e = tid; e < n_items; e += nthreads) {
        const int an = e / TG, gl = e - an * TG;
        const int n = an / R, a = an - n * R;
        cf* lo = xin + (a + R * (n + 1)) * ld + gl;
        cf* hi = xin + (a + R * (Q - 1 - n)) * ld + gl;
        const cf x = *lo, y = *hi;
        *lo = cadd(x, y);
        *hi = csub(x, y);
    }
}

// LD = TG + 1 (row pitch of the tile) is a template parameter so that every shared-memory offset of the inner loop is an
// immediate; the twiddle index is kept pre-scaled (stw[R (n k mod Q)] = stw[(n R k) mod P]).
template <int R, int KP, int LD>
RSP_HD void mtd_dft_sym_compute(const cf* xin, const cf* stw, int P, int kp0, int gl, cf (&A)[KP][R], cf (&B)[KP][R], int (&kk)[KP]) {
    const int Q = P / R, H = (Q - 1) / 2;
    int idx[KP], step[KP];                                  // A = x0 + sum S cos,  B = - sum D sin
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        kk[j] = kp0 + j <= H ? kp0 + j : H;                 // pairs past the end repeat the last one and are not stored
        idx[j] = 0;
        step[j] = R * kk[j];
#pragma unroll
        for (int a = 0; a < R; ++a) {
            A[j][a] = xin[a * LD + gl];
            B[j][a] = make_float2(0.f, 0.f);
        }
    }
    const cf* pS = xin + gl + R * LD;                       // row n = 1 of sub-sequence 0
    const cf* pD = xin + gl + R * (Q - 1) * LD;             // row Q - 1
    for (int n = 1; n <= H; ++n, pS += R * LD, pD -= R * LD) {
        cf S[R], D[R];
#pragma unroll
        for (int a = 0; a < R; ++a) {
            S[a] = pS[a * LD];
            D[a] = pD[a * LD];
        }
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            idx[j] += step[j];
            if (idx[j] >= P) idx[j] -= P;                   // R (n k_j mod Q)
            const cf w = stw[idx[j]];                       // (cos, -sin) of 2 pi n k_j / Q
#pragma unroll
            for (int a = 0; a < R; ++a) {
                A[j][a].x = fmaf(S[a].x, w.x, A[j][a].x);
                A[j][a].y = fmaf(S[a].y, w.x, A[j][a].y);
                B[j][a].x = fmaf(D[a].x, w.y, B[j][a].x);
                B[j][a].y = fmaf(D[a].y, w.y, B[j][a].y);
            }
        }
    }
}

template <int R, int KP, int LD>
RSP_HD void mtd_dft_sym_store(cf* xout, const cf* stw, int P, int kp0, int gl, const cf (&A)[KP][R], const cf (&B)[KP][R], const int (&kk)[KP]) {
    const int Q = P / R, H = (Q - 1) / 2, half = P / 2;
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        if (kp0 + j > H) break;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (side == 1 && kk[j] == 0) break;             // bin 0 has no mirror
            const int k = side == 0 ? kk[j] : Q - kk[j];
            cf acc[R];
#pragma unroll
            for (int a = 0; a < R; ++a)                     // X[k] = A + i B,  X[Q-k] = A - i B   (B carries the minus of the sine sum)
                acc[a] = side == 0 ? make_float2(A[j][a].x - B[j][a].y, A[j][a].y + B[j][a].x)
                                   : make_float2(A[j][a].x + B[j][a].y, A[j][a].y - B[j][a].x);
#pragma unroll
            for (int a = 1; a < R; ++a) acc[a] = cmul(acc[a], stw[a * k]);
            if (R > 1) SmallDft<(R > 1 ? R : 2), -1>::run(acc);
#pragma unroll
            for (int d = 0; d < R; ++d) {
                int row = k + Q * d + half;
                if (row >= P) row -= P;
                xout[row * LD + gl] = acc[d];
            }
        }
    }
}

// work items of one tile: groups of KP bin pairs x TG gates
RSP_HD int mtd_dft_sym_groups(int Q, int KP) { return ((Q - 1) / 2 + 1 + KP - 1) / KP; }
// When one round of the CTA covers all items, the results wait in registers until every item has read the tile and then
// overwrite it: one tile instead of two in shared memory (twice the CTAs per SM at P = 332).
RSP_HD bool mtd_dft_sym_inplace(int Q, int KP, int TG, int nthreads) { return mtd_dft_sym_groups(Q, KP) * TG <= nthreads; }

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
    int pad_pitch = 1;     // vectorised kernel: 1 = conflict-free row pitches (cfar4_pitch), 0 = dense rows
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

// =============================================================================================
// CFAR, vectorised variant (P % 4 == 0): every thread works on quads of 4 consecutive Doppler bins.
//   S tile: rows = gates [g_first - mR, g_first + TG + mR), row pitch PP = P + 16 floats with the
//   data at column offset 8 (zero halo of 8 on both sides, so the aligned Doppler-window loads of
//   partially valid quads never leave the row).  R5 (range window sums) has pitch P.
//   RR / RV / GV are compile-time reference / guard lengths (RR = 0: everything at run time).
// =============================================================================================
#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif
#define RSP_CFAR_HALO 8

struct Cfar4Geom {
    int P, P4, sh;        // P4 = P/4; sh = log2(P4) when P4 is a power of two, else -1
    int PP;               // padded row pitch of S (floats): >= P + 2*RSP_CFAR_HALO and == 8 (mod 32), see cfar4_pitch
    int RP;               // row pitch of R5 (floats): >= P and == 8 (mod 32)
    int TG, rows, r5_rows;
};

// Row pitch (floats) >= need with pitch == 8 (mod 32), i.e. 32 bytes (mod 128): the decision phase walks the CUT quads
// of a row and then jumps to the next row; with such a pitch the 16-byte accesses of a quarter-warp that straddles two
// rows still fall into eight different 16-byte bank groups (the old pitches 80 and 64 floats gave 2-way conflicts on
// every straddling quarter: 37 % extra wavefronts on the seven LDS.128 of the decision, profiles/r1b_*).
RSP_HD int cfar4_pitch(int need) {
    int p = (need + 3) & ~3;
    while ((p & 31) != 8) p += 4;
    return p;
}

RSP_HD Cfar4Geom cfar4_geom(const CfarParams& c, int TG) {
    Cfar4Geom g;
    g.P = c.P;
    g.P4 = c.P / 4;
    g.sh = -1;
    for (int s = 0; s < 16; ++s)
        if ((1 << s) == g.P4) g.sh = s;
    g.PP = c.pad_pitch ? cfar4_pitch(c.P + 2 * RSP_CFAR_HALO) : c.P + 2 * RSP_CFAR_HALO;
    g.RP = c.pad_pitch ? cfar4_pitch(c.P) : c.P;
    g.TG = TG;
    g.rows = TG + 2 * (c.guard_r + c.ref_r);
    g.r5_rows = cfar_r5_rows(c, TG);
    return g;
}

RSP_HD void cfar4_split(const Cfar4Geom& g, int idx, int& row, int& c4) {
    if (g.sh >= 0) { row = idx >> g.sh; c4 = idx & (g.P4 - 1); }
    else { row = idx / g.P4; c4 = idx - row * g.P4; }
}

RSP_HD float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

template <int RR>
RSP_HD void cfar4_r5_phase(const float* S, float* R5, const CfarParams& c, const Cfar4Geom& g, int tid, int nthreads) {
    const int rr = RR > 0 ? RR : c.ref_r;
    const int n = g.r5_rows * g.P4;
    const int pp4 = g.PP / 4;
    const float4* S4 = reinterpret_cast<const float4*>(S) + RSP_CFAR_HALO / 4;
    float4* R4 = reinterpret_cast<float4*>(R5);
    for (int idx = tid; idx < n; idx += nthreads) {
        int row, c4;
        cfar4_split(g, idx, row, c4);
        float4 acc = S4[row * pp4 + c4];
#pragma unroll
        for (int i = 1; i < rr; ++i) acc = f4add(acc, S4[(row + i) * pp4 + c4]);
        R4[row * (g.RP / 4) + c4] = acc;
    }
}

// Window sums of 4 consecutive CUTs from aligned float4 loads: out[j] = sum_{i<RV} row[start + j + i],
// start = v0 + OFF with v0 % 4 == 0 and OFF a compile-time constant (may be negative).
template <int RV, int OFF> RSP_HD void cfar4_window4(const float* row_v0, float out[4]) {
    constexpr int LO = ((OFF % 4) + 4) % 4;              // offset of the first element inside its float4
    constexpr int NV = (LO + RV + 3 + 3) / 4;            // aligned float4s that cover the span
    float x[4 * NV];
    const float4* p = reinterpret_cast<const float4*>(row_v0 + (OFF - LO));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 q = p[i];
        x[4 * i] = q.x; x[4 * i + 1] = q.y; x[4 * i + 2] = q.z; x[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = x[LO + j];
#pragma unroll
        for (int i = 1; i < RV; ++i) a += x[LO + j + i];
        out[j] = a;
    }
}

// Decision for the quad (gl, c4): returns a 4-bit detection mask, cut[] receives the four S values.
template <int RR, int RV, int GV>
RSP_HD unsigned cfar4_decide_quad(const float* S, const float* R5, const CfarParams& c, const Cfar4Geom& g, int gl, int c4,
                                  float cut[4]) {
    const int rr = RR > 0 ? RR : c.ref_r, rv = RR > 0 ? RV : c.ref_v, gv = RR > 0 ? GV : c.guard_v;
    const int mR = c.guard_r + c.ref_r, mV = gv + rv;
    const int v0 = 4 * c4;
    const float* row = S + (gl + mR) * g.PP + RSP_CFAR_HALO;             // row[v] = S(gl, v)
    const float4 cq = *reinterpret_cast<const float4*>(row + v0);
    const float4 lr = *reinterpret_cast<const float4*>(R5 + gl * g.RP + v0);
    const float4 tr = *reinterpret_cast<const float4*>(R5 + (gl + mR + c.guard_r + 1) * g.RP + v0);
    float lead[4], trail[4];
    if (RR > 0) {
        cfar4_window4<(RR > 0 ? RV : 1), -(GV + RV)>(row + v0, lead);
        cfar4_window4<(RR > 0 ? RV : 1), GV + 1>(row + v0, trail);
    } else {
        for (int j = 0; j < 4; ++j) {
            float a = row[v0 + j - mV], b = row[v0 + j + gv + 1];
            for (int i = 1; i < rv; ++i) { a += row[v0 + j - mV + i]; b += row[v0 + j + gv + 1 + i]; }
            lead[j] = a;
            trail[j] = b;
        }
    }
    const float kr = c.t_cfar / (float)rr, kv = c.t_cfar / (float)rv;     // T * mean == (T/ref) * sum
    const float cu[4] = {cq.x, cq.y, cq.z, cq.w};
    const float l4[4] = {lr.x, lr.y, lr.z, lr.w}, t4[4] = {tr.x, tr.y, tr.z, tr.w};
    unsigned mask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int v = v0 + j;
        const float thr = fmaxf(kr * fmaxf(l4[j], t4[j]), kv * fmaxf(lead[j], trail[j]));
        cut[j] = cu[j];
        if (v >= mV && v < c.P - mV && cu[j] > thr) mask |= 1u << j;
    }
    return mask;
}

// =============================================================================================
// CFAR, marching variant (cfar5_kernel; P / 4 a power of two, compile-time windows).  Same cells as cfar4_*, about a
// third of the instructions:
//   * no range-sum array: a work item owns one CUT quad (4 Doppler bins) of CR consecutive gates and walks down them
//     with the rows of its leading and trailing range windows in registers (one new row of each per gate; the sums are
//     re-added left to right for every gate, exactly like the reference's mean(), no subtractive updates);
//   * range test first: the Doppler windows are only summed for quads with a cell above the range threshold
//     (cut > T max(noise_R, noise_V)  <=>  cut > T noise_R  and  cut > T noise_V, fun_process_single_frame.m:196-209);
//   * the tile keeps 4 (not 8) pad floats on each side of a row and they are never initialised: every window that
//     touches them belongs to a Doppler bin outside [mV, P - mV), which the decision masks.
// Row pitch: items are numbered chunk-major (it = chunk * nq + quad), so consecutive lanes read consecutive 16-byte
// groups only if a chunk's offset CR * pitch / 4 is congruent to nq (mod 8): cfar5_pitch picks the smallest such pitch.
// =============================================================================================
#define RSP_CFAR5_HALO 4
#define RSP_CFAR5_CR 5

RSP_HD int cfar5_nq(int P, int mV) { return (P - mV - 1) / 4 - mV / 4 + 1; }          // quads holding at least one CUT
RSP_HD int cfar5_pitch(int P, int mV) {
    const int nq = cfar5_nq(P, mV);
    int p4 = (P + 2 * RSP_CFAR5_HALO) / 4;
    while (((RSP_CFAR5_CR * p4 - nq) & 7) != 0) ++p4;                                 // CR is odd: a solution every 8 steps
    return 4 * p4;
}

RSP_HD float4 cfar5_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Doppler test of a quad that passed the range test: clears the bits of the cells at or below T * noise_V.
// row = S(CUT row, 4 c4).
template <int RV, int GV>
RSP_HD unsigned cfar5_doppler(const float* row, float kv, unsigned m, float4* cq_out) {
    float lead[4], trail[4];
    cfar4_window4<RV, -(GV + RV)>(row, lead);
    cfar4_window4<RV, GV + 1>(row, trail);
    const float4 cq = cfar5_ld4(row);
    const float cu[4] = {cq.x, cq.y, cq.z, cq.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (!(cu[j] > kv * fmaxf(lead[j], trail[j]))) m &= ~(1u << j);
    *cq_out = cq;
    return m;
}

// Rows [gl0, gl0 + n_rows) of CUT quad c4 (n_rows <= CR).  S = tile base (row 0 = gate g_first - mR), pitch in floats.
// hit(s, mask) is called for every row gl0 + s with cells above the RANGE threshold (bit j of mask = Doppler bin
// 4 c4 + j); the caller queues them for cfar5_doppler (rare, so the marching loop stays lean in registers).
template <int RR, int GV_PLUS_RV, int CR, class Hit>
RSP_HD void cfar5_march(const float* S, int pitch, const CfarParams& c, int gl0, int n_rows, int c4, Hit&& hit) {
    const int mR = c.guard_r + RR, mV = GV_PLUS_RV;
    const int v0 = 4 * c4;
    unsigned vmask = 0;                                                   // Doppler bins of the quad that are cells under test
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (v0 + j >= mV && v0 + j < c.P - mV) vmask |= 1u << j;
    const float* pl = S + RSP_CFAR5_HALO + v0 + (size_t)gl0 * pitch;      // next leading-window row: S(gl0 + i, v0)
    const float* pt = pl + (mR + c.guard_r + 1) * pitch;                  // next trailing-window row
    const float* pc = pl + mR * pitch;                                    // CUT row
    const float kr = c.t_cfar / (float)RR;                                // T * mean == (T / ref) * sum
    float4 L[RR], T[RR];                                                  // rings: window row i of step s lives in slot (s + i) % RR
#pragma unroll
    for (int i = 0; i < RR - 1; ++i) {
        L[i] = cfar5_ld4(pl); pl += pitch;
        T[i] = cfar5_ld4(pt); pt += pitch;
    }
#pragma unroll
    for (int s = 0; s < CR; ++s) {
        if (s < n_rows) {
            L[(s + RR - 1) % RR] = cfar5_ld4(pl); pl += pitch;
            T[(s + RR - 1) % RR] = cfar5_ld4(pt); pt += pitch;
            const float4 cq = cfar5_ld4(pc); pc += pitch;
            float4 ls = L[s % RR], ts = T[s % RR];
#pragma unroll
            for (int i = 1; i < RR; ++i) { ls = f4add(ls, L[(s + i) % RR]); ts = f4add(ts, T[(s + i) % RR]); }
            unsigned m = (cq.x > kr * fmaxf(ls.x, ts.x) ? 1u : 0u) | (cq.y > kr * fmaxf(ls.y, ts.y) ? 2u : 0u) |
                         (cq.z > kr * fmaxf(ls.z, ts.z) ? 4u : 0u) | (cq.w > kr * fmaxf(ls.w, ts.w) ? 8u : 0u);
            m &= vmask;
            if (m) hit(s, m);
        }
    }
}
