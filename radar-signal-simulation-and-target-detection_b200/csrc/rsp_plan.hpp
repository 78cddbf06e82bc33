// rsp_plan.hpp -- host-side (fp64) construction of FFT plans, twiddle tables and block filter
// spectra.  Host only; shared by the library (rsp_api.cu) and the host-emulation test.
#pragma once
#include <complex>
#include <vector>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "rsp_phases.cuh"

namespace rsp {

typedef std::complex<double> zc;
static const double kPi = 3.14159265358979323846264338327950288;

// Iterative radix-2 fp64 FFT (one-time setup work only: filter spectra).
inline void host_fft(std::vector<zc>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                const double ang = -2.0 * kPi * (double)k / (double)len;
                const zc w(std::cos(ang), std::sin(ang));
                const zc u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
        }
    }
}

inline std::vector<cf> make_twiddles(int Ls, int r) {   // [(k-1)*span + j] = e^{-2 pi i jk/Ls}
    const int span = Ls / r;
    std::vector<cf> t((size_t)(r - 1) * span);
    for (int k = 1; k < r; ++k)
        for (int j = 0; j < span; ++j) {
            const double ang = -2.0 * kPi * (double)((long long)j * k % Ls) / (double)Ls;
            t[(size_t)(k - 1) * span + j] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
    return t;
}

// Pulse-compression block plan: L = R1*R2*R3 (see PcCfg in rsp_phases.cuh).
struct PcPlan {
    int L = 0, R1 = 0, R2 = 0, R3 = 0, T = 0, taps = 0, valid = 0, nblk = 0;
    int seg_start0 = 0, gate0 = 0, ngates = 0;
    std::vector<cf> tw1, tw2, Hmid;
};

inline bool pc_radices(int L, int* r) {
    switch (L) {
        case 1024: r[0] = 16; r[1] = 16; r[2] = 4; return true;
        case 2048: r[0] = 8; r[1] = 16; r[2] = 16; return true;
        case 4096: r[0] = 16; r[1] = 16; r[2] = 16; return true;
    }
    return false;
}

inline bool make_pc_plan(PcPlan& pl, int L, const zc* taps, int ntaps, int seg_start0, int gate0, int ngates) {
    int r[3];
    if (!pc_radices(L, r)) return false;
    if (ntaps < 1 || ntaps > L) return false;
    pl.L = L;
    pl.R1 = r[0]; pl.R2 = r[1]; pl.R3 = r[2];
    pl.T = L / 16;
    pl.taps = ntaps;
    pl.valid = L - (ntaps - 1);
    pl.seg_start0 = seg_start0;
    pl.gate0 = gate0;
    pl.ngates = ngates;
    pl.nblk = (ngates + pl.valid - 1) / pl.valid;
    pl.tw1 = make_twiddles(L, pl.R1);
    pl.tw2 = make_twiddles(L / pl.R1, pl.R2);
    std::vector<zc> h((size_t)L, zc(0, 0));
    for (int i = 0; i < ntaps; ++i) h[i] = taps[i];
    host_fft(h);
    std::vector<cf> Hdr((size_t)L);
    for (int f = 0; f < L; ++f) {
        const int pos = rsp_digit_reverse(f, L, r, 3);
        Hdr[pos] = make_float2((float)(h[f].real() / L), (float)(h[f].imag() / L));
    }
    // Hmid[(i*R3 + k)*T + t] = Hdr[R3*(t + T*i) + k]: coalesced over the threads of a group
    const int nb3 = 16 / pl.R3;
    pl.Hmid.assign((size_t)L, make_float2(0, 0));
    for (int i = 0; i < nb3; ++i)
        for (int k = 0; k < pl.R3; ++k)
            for (int t = 0; t < pl.T; ++t) pl.Hmid[(size_t)(i * pl.R3 + k) * pl.T + t] = Hdr[(size_t)pl.R3 * (t + pl.T * i) + k];
    return true;
}

// Cost model used to pick the block length: every pass keeps all threads busy, so the work is
// proportional to the points transformed; longer blocks waste less on the (taps-1) overlap.
inline double pc_plan_cost(int L, int ntaps, int ngates) {
    const int valid = L - (ntaps - 1);
    if (valid < 1) return 1e300;
    const int nblk = (ngates + valid - 1) / valid;
    return (double)nblk * L;
}

inline int choose_pc_len(int ntaps, int ngates) {
    const int cands[3] = {1024, 2048, 4096};
    int best = 0;
    double bc = 1e300;
    for (int i = 0; i < 3; ++i) {
        const double c = pc_plan_cost(cands[i], ntaps, ngates);
        if (c < bc) { bc = c; best = cands[i]; }    // ties: the shorter block (its first pass is a cheaper radix; measured +2 %)
    }
    return best;
}

// Mixed block plan: counts[i] blocks of length {4096, 2048, 1024}[i] laid end to end over the segment's gates, chosen to
// minimise the transformed points.  At config 2 the long segment has 4826 gates and 700 taps: four 2048-point blocks
// (1349 valid gates each) transform 8192 points, one 4096 + one 2048 + one 1024 block (3397 + 1349 + 325 >= 4826) only
// 7168.  Returns the total points, or 0 when no block length fits the filter.
inline int choose_pc_mix(int ntaps, int ngates, int counts[3]) {
    const int Ls[3] = {4096, 2048, 1024};
    int best = 0, best_blocks = 0;
    counts[0] = counts[1] = counts[2] = 0;
    for (int a = 0; a <= 16; ++a)
        for (int b = 0; b <= 16; ++b)
            for (int c = 0; c <= 16; ++c) {
                const int n[3] = {a, b, c};
                long cover = 0;
                int pts = 0, blocks = 0;
                bool ok = true;
                for (int i = 0; i < 3; ++i) {
                    const int valid = Ls[i] - (ntaps - 1);
                    if (n[i] > 0 && valid < 1) ok = false;
                    cover += (long)n[i] * (valid > 0 ? valid : 0);
                    pts += n[i] * Ls[i];
                    blocks += n[i];
                }
                if (!ok || blocks == 0 || cover < ngates) continue;
                if (best == 0 || pts < best || (pts == best && blocks < best_blocks)) {
                    best = pts; best_blocks = blocks;
                    counts[0] = a; counts[1] = b; counts[2] = c;
                }
            }
    return best;
}

// Doppler plan for power-of-two P: radices (R0,R1,R2) in DIF order, fixed per P so that the kernel
// template (MtdCfg) and the host tables agree.
inline bool mtd_radices(int P, int* r) {
    switch (P) {
        case 2: r[0] = 2; r[1] = 1; r[2] = 1; return true;
        case 4: r[0] = 4; r[1] = 1; r[2] = 1; return true;
        case 8: r[0] = 8; r[1] = 1; r[2] = 1; return true;
        case 16: r[0] = 16; r[1] = 1; r[2] = 1; return true;
        case 32: r[0] = 8; r[1] = 4; r[2] = 1; return true;
        case 64: r[0] = 8; r[1] = 8; r[2] = 1; return true;
        case 128: r[0] = 16; r[1] = 8; r[2] = 1; return true;
        case 256: r[0] = 16; r[1] = 16; r[2] = 1; return true;
        case 512: r[0] = 8; r[1] = 8; r[2] = 8; return true;
    }
    return false;
}

struct DopplerPlan {
    int P = 0, r[3] = {1, 1, 1};
    std::vector<cf> tw;        // passes concatenated at the MtdCfg offsets
    std::vector<int> perm;     // perm[p] = position of input pulse p
    std::vector<int> iperm;    // iperm[pos] = pulse stored at pos
};

inline bool make_doppler_plan(DopplerPlan& dp, int P) {
    if (!mtd_radices(P, dp.r)) return false;
    dp.P = P;
    dp.tw.clear();
    int Ls = P;
    for (int s = 0; s < 3; ++s) {
        if (dp.r[s] > 1) {
            std::vector<cf> t = make_twiddles(Ls, dp.r[s]);
            dp.tw.insert(dp.tw.end(), t.begin(), t.end());
        }
        Ls /= dp.r[s];
    }
    int nrad = 0, rad[3];
    for (int s = 0; s < 3; ++s) if (dp.r[s] > 1) rad[nrad++] = dp.r[s];
    dp.perm.resize(P);
    dp.iperm.resize(P);
    for (int p = 0; p < P; ++p) {
        dp.perm[p] = rsp_digit_reverse(p, P, rad, nrad);
        dp.iperm[dp.perm[p]] = p;
    }
    return true;
}

// Round-to-nearest (ties away) fp32 -> tf32, like cvt.rna.tf32.f32; result still an fp32 bit pattern.
inline float host_tf32(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

// Weight fragments of dbf_mma_kernel: frag[(s*NT + nt)*32 + lane] = {b0h, b1h, b0l, b1l} where, with
// g = lane >> 2, t = lane & 3, c = 4s + t, b = 4nt + (g >> 1), (wr, wi) = conj(W[b][c]):
//   output column re (g even): b0 = wr, b1 = -wi;   output column im (g odd): b0 = wi, b1 = wr.
inline std::vector<float4> make_dbf_fragments(const double* W_ri /* [B][C][2] */, int B, int C, int NT, int KS) {
    std::vector<float4> f((size_t)KS * NT * 32);
    for (int s = 0; s < KS; ++s)
        for (int nt = 0; nt < NT; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3, c = 4 * s + t, b = 4 * nt + (g >> 1);
                float b0 = 0.f, b1 = 0.f;
                if (c < C && b < B) {
                    const float wr = (float)W_ri[((size_t)b * C + c) * 2], wi = (float)-W_ri[((size_t)b * C + c) * 2 + 1];
                    if ((g & 1) == 0) { b0 = wr; b1 = -wi; } else { b0 = wi; b1 = wr; }
                }
                const float b0h = host_tf32(b0), b1h = host_tf32(b1);
                f[(size_t)(s * NT + nt) * 32 + lane] = make_float4(b0h, b1h, host_tf32(b0 - b0h), host_tf32(b1 - b1h));
            }
    return f;
}

// Weight fragments for dbf_mma2_kernel (weights as the A operand of mma m16n8k8 tf32):
//   f[((s*MT + mt)*2 + {hi, lo})*32 + lane] = {a0, a1, a2, a3},  lane = 4g + t, beam b = 8 mt + g, channel c = 4 s + t,
//   a0 = A[Re b][(c, re)] = wr, a1 = A[Im b][(c, re)] = -wi, a2 = A[Re b][(c, im)] = wi, a3 = A[Im b][(c, im)] = wr
//   for out = sum_c x_c conj(W[b][c]), W = wr + i wi  (fun_process_single_frame.m:95, x * W').
inline std::vector<float4> make_dbf_fragments_wa(const double* W_ri /* [B][C][2] */, int B, int C, int MT, int KS) {
    std::vector<float4> f((size_t)KS * MT * 2 * 32);
    for (int s = 0; s < KS; ++s)
        for (int mt = 0; mt < MT; ++mt)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t = lane & 3, c = 4 * s + t, b = 8 * mt + g;
                float a[4] = {0.f, 0.f, 0.f, 0.f};
                if (c < C && b < B) {
                    const float wr = (float)W_ri[((size_t)b * C + c) * 2], wi = (float)W_ri[((size_t)b * C + c) * 2 + 1];
                    a[0] = wr; a[1] = -wi; a[2] = wi; a[3] = wr;
                }
                float h[4], l[4];
                for (int i = 0; i < 4; ++i) { h[i] = host_tf32(a[i]); l[i] = host_tf32(a[i] - h[i]); }
                f[((size_t)(s * MT + mt) * 2 + 0) * 32 + lane] = make_float4(h[0], h[1], h[2], h[3]);
                f[((size_t)(s * MT + mt) * 2 + 1) * 32 + lane] = make_float4(l[0], l[1], l[2], l[3]);
            }
    return f;
}

}  // namespace rsp
