// rsp_plan.hpp -- host-side (fp64) construction of FFT plans, twiddle tables and block filter
// spectra.  Host only; shared by the library (rsp_api.cu) and the host-emulation test.
#pragma once
#include <complex>
#include <vector>
#include <cmath>
#include <cstdint>
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

// Pulse-compression block plan: L = R1*256, radices (R1,16,16).
struct PcPlan {
    int L = 0, R1 = 0, taps = 0, valid = 0, nblk = 0;
    int seg_start0 = 0, gate0 = 0, ngates = 0;
    std::vector<cf> tw1, tw2, H;
};

inline bool make_pc_plan(PcPlan& pl, int L, const zc* taps, int ntaps, int seg_start0, int gate0, int ngates) {
    if (L != 1024 && L != 2048 && L != 4096) return false;
    if (ntaps < 1 || ntaps > L) return false;
    pl.L = L;
    pl.R1 = L / 256;
    pl.taps = ntaps;
    pl.valid = L - (ntaps - 1);
    pl.seg_start0 = seg_start0;
    pl.gate0 = gate0;
    pl.ngates = ngates;
    pl.nblk = (ngates + pl.valid - 1) / pl.valid;
    pl.tw1 = make_twiddles(L, pl.R1);
    pl.tw2 = make_twiddles(256, 16);
    std::vector<zc> h((size_t)L, zc(0, 0));
    for (int i = 0; i < ntaps; ++i) h[i] = taps[i];
    host_fft(h);
    const int radices[3] = {pl.R1, 16, 16};
    pl.H.assign((size_t)L, make_float2(0, 0));
    for (int f = 0; f < L; ++f) {
        const int pos = rsp_digit_reverse(f, L, radices, 3);
        pl.H[pos] = make_float2((float)(h[f].real() / L), (float)(h[f].imag() / L));
    }
    return true;
}

// Cost model used to pick the block length: pass 1 costs ~L*log2(R1), passes 2/3
// run on max(L/16, 256) thread slots of 16 points each.
inline double pc_plan_cost(int L, int ntaps, int ngates) {
    const int valid = L - (ntaps - 1);
    if (valid < 1) return 1e300;
    const int nblk = (ngates + valid - 1) / valid;
    const double l2r1 = std::log2((double)L / 256.0);
    const double slots = (double)(L > 4096 ? L : 4096);
    return nblk * (L * l2r1 + 2.0 * slots * 4.0);
}

inline int choose_pc_len(int ntaps, int ngates) {
    const int cands[3] = {1024, 2048, 4096};
    int best = 0;
    double bc = 1e300;
    for (int i = 0; i < 3; ++i) {
        const double c = pc_plan_cost(cands[i], ntaps, ngates);
        if (c < bc) { bc = c; best = cands[i]; }
    }
    return best;
}

// Doppler plan for power-of-two P (2 <= P <= 4096).
struct DopplerPlan {
    MtdPlan plan;
    std::vector<cf> tw;        // all passes concatenated
    std::vector<int> perm;     // perm[p] = position of input pulse p
};

inline bool make_doppler_plan(DopplerPlan& dp, int P) {
    if (P < 2 || (P & (P - 1))) return false;
    int lg = 0;
    while ((1 << lg) < P) ++lg;
    MtdPlan& pl = dp.plan;
    pl.P = P;
    pl.nrad = 0;
    int rem = lg;
    while (rem > 0) {
        int step = rem >= 4 ? 4 : rem;
        if (rem > 4 && rem < 8 && rem - step < 2 && rem - step > 0) step = rem - 2;   // avoid a trailing radix-2
        if (pl.nrad >= 4) return false;
        pl.radices[pl.nrad++] = 1 << step;
        rem -= step;
    }
    dp.tw.clear();
    int Ls = P;
    for (int s = 0; s < pl.nrad; ++s) {
        pl.tw_off[s] = (int)dp.tw.size();
        std::vector<cf> t = make_twiddles(Ls, pl.radices[s]);
        dp.tw.insert(dp.tw.end(), t.begin(), t.end());
        Ls /= pl.radices[s];
    }
    dp.perm.resize(P);
    for (int p = 0; p < P; ++p) dp.perm[p] = rsp_digit_reverse(p, P, pl.radices, pl.nrad);
    return true;
}

}  // namespace rsp
