// host_emul.cpp -- g++-built (no GPU) emulation of the kernel phases in rsp_phases.cuh.
// TEST INFRASTRUCTURE: lets tests/test_host_emulation.py check the exact per-thread code of the
// CUDA kernels against NumPy on a machine without a GPU.  Barriers become loops over tid.
#include <cstring>
#include <algorithm>
#include <vector>
#include <type_traits>
#include "rsp_plan.hpp"
#include "rsp_dft_big.cuh"

using namespace rsp;

template <class Cfg> static void run_pc_block(const PcBlockArgs& a, cf* s) {
    for (int t = 0; t < Cfg::T; ++t) pc_phase_load_pass1<Cfg>(a, s, t);
    for (int t = 0; t < Cfg::T; ++t) pc_phase_pass2<Cfg>(a, s, t);
    for (int t = 0; t < Cfg::T; ++t) pc_phase_mid<Cfg>(a, s, t);
    for (int t = 0; t < Cfg::T; ++t) pc_phase_ipass2<Cfg>(a, s, t);
    for (int t = 0; t < Cfg::T; ++t) pc_phase_ipass1_store<Cfg>(a, s, t);
}

template <class Cfg> static void run_mtd_tile(cf* s, const cf* tw, const cf* src, size_t pstride, const int* iperm,
                                              const float* win) {
    for (int t = 0; t < RSP_MTD_THREADS; ++t)
        mtd_first_pass_t<Cfg>(s, src, pstride, win, true, t);
    for (int pass = MtdInner<Cfg>::PASS + 1; pass < 3; ++pass)
        for (int t = 0; t < RSP_MTD_THREADS; ++t) mtd_passes_phase<Cfg>(s, tw, t, pass);
}

extern "C" {

int emul_small_dft(int R, int sign, float* v) {
    cf* c = reinterpret_cast<cf*>(v);
    if (sign < 0) {
        switch (R) {
            case 2: SmallDft<2, -1>::run(c); return 0;
            case 4: SmallDft<4, -1>::run(c); return 0;
            case 8: SmallDft<8, -1>::run(c); return 0;
            case 16: SmallDft<16, -1>::run(c); return 0;
            case 32: SmallDft<32, -1>::run(c); return 0;
            case 64: SmallDft<64, -1>::run(c); return 0;
        }
    } else {
        switch (R) {
            case 2: SmallDft<2, +1>::run(c); return 0;
            case 4: SmallDft<4, +1>::run(c); return 0;
            case 8: SmallDft<8, +1>::run(c); return 0;
            case 16: SmallDft<16, +1>::run(c); return 0;
            case 32: SmallDft<32, +1>::run(c); return 0;
            case 64: SmallDft<64, +1>::run(c); return 0;
        }
    }
    return -1;
}

// One segment of one line through the overlap-save blocks.  L == 0 -> library's own choice.
int emul_pc_segment(const float* line, int N, int seg_start0, int gate0, int ngates, const double* taps_ri,
                    int ntaps, int L, float* out_line, int* L_used, int* nblk_used) {
    if (L == 0) L = choose_pc_len(ntaps, ngates);
    std::vector<zc> taps(ntaps);
    for (int i = 0; i < ntaps; ++i) taps[i] = zc(taps_ri[2 * i], taps_ri[2 * i + 1]);
    PcPlan pl;
    if (!make_pc_plan(pl, L, taps.data(), ntaps, seg_start0, gate0, ngates)) return -1;
    std::vector<cf> smem((size_t)L + L / 16 + 16);
    for (int blk = 0; blk < pl.nblk; ++blk) {
        PcBlockArgs a;
        a.line = reinterpret_cast<const cf*>(line);
        a.out_line = reinterpret_cast<cf*>(out_line);
        a.tw1 = pl.tw1.data();
        a.tw2 = pl.tw2.data();
        a.Hmid = pl.Hmid.data();
        a.in_lo = seg_start0;
        a.in_hi = N;
        a.seg_start0 = seg_start0;
        a.taps = ntaps;
        a.g0 = gate0 + blk * pl.valid;
        a.g_end = gate0 + ngates;
        switch (pl.L) {
            case 1024: run_pc_block<PcCfg<1024, 16, 16, 4>>(a, smem.data()); break;
            case 2048: run_pc_block<PcCfg<2048, 8, 16, 16>>(a, smem.data()); break;
            case 4096: run_pc_block<PcCfg<4096, 16, 16, 16>>(a, smem.data()); break;
            default: return -2;
        }
    }
    if (L_used) *L_used = L;
    if (nblk_used) *nblk_used = pl.nblk;
    return 0;
}

// The mixed block plan of a segment (choose_pc_mix): parts laid end to end exactly as rsp_upload_constants builds them.
// counts_out = blocks of {4096, 2048, 1024}; returns the transformed points per line, or < 0 on error.
int emul_pc_segment_mixed(const float* line, int N, int seg_start0, int gate0, int ngates, const double* taps_ri,
                          int ntaps, float* out_line, int* counts_out) {
    int counts[3];
    const int pts = choose_pc_mix(ntaps, ngates, counts);
    if (pts <= 0) return -1;
    const int Ls[3] = {4096, 2048, 1024};
    int gate = gate0, left = ngates;
    for (int i = 0; i < 3; ++i) {
        counts_out[i] = counts[i];
        if (!counts[i]) continue;
        const int take = std::min(left, counts[i] * (Ls[i] - (ntaps - 1)));
        if (take <= 0) break;
        int lu = 0, nb = 0;
        const int rc = emul_pc_segment(line, N, seg_start0, gate, take, taps_ri, ntaps, Ls[i], out_line, &lu, &nb);
        if (rc) return rc - 10;
        if (nb > counts[i]) return -3;
        gate += take; left -= take;
    }
    return left == 0 ? pts : -4;
}

// The DBF of dbf_mma2_kernel on the host: same weight fragments (make_dbf_fragments_wa), same lane <-> (row, column, k)
// mapping of mma.m16n8k8 (A row-major 16x8: a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4); B 8x8: b0=(t,g) b1=(t+4,g);
// D 16x8: c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)), same sample order sg and the same store addresses.
// The MMA itself is done in double on hi + lo, so this checks the index algebra, not the TF32 rounding.
int emul_dbf_wa(const float* raw /* [C][N] complex */, int C, int N, const double* W_ri /* [B][C][2] */, int B,
                float* beam /* [B][N] complex, zero-initialised by the caller */) {
    if (N % 2) return -1;
    const int MT = (B + 7) / 8, KS = C <= 16 ? 4 : 8;
    if (C > 32 || B > 16) return -2;
    const std::vector<float4> Wa = make_dbf_fragments_wa(W_ri, B, C, MT, KS);
    for (int n_base = 0; n_base < N; n_base += 32) {
        // acc[lane][mt][tile j][4]
        std::vector<double> acc((size_t)32 * MT * 4 * 4, 0.0);
        for (int s = 0; s < KS; ++s)
            for (int mt = 0; mt < MT; ++mt)
                for (int j = 0; j < 4; ++j) {                    // n-tile j = 2q + parity
                    const int q = j >> 1, par = j & 1;
                    double A[16][8], Bm[8][8];
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, t = lane & 3;
                        const float4 h = Wa[((size_t)(s * MT + mt) * 2 + 0) * 32 + lane], l = Wa[((size_t)(s * MT + mt) * 2 + 1) * 32 + lane];
                        A[g][t] = (double)h.x + l.x; A[g + 8][t] = (double)h.y + l.y;
                        A[g][t + 4] = (double)h.z + l.z; A[g + 8][t + 4] = (double)h.w + l.w;
                        const int sg = (g & 1) ? g + 7 : g, c = 4 * s + t, n = n_base + 16 * q + sg + par;
                        double re = 0.0, im = 0.0;
                        if (c < C && n < N) { re = raw[((size_t)c * N + n) * 2]; im = raw[((size_t)c * N + n) * 2 + 1]; }
                        Bm[t][g] = re; Bm[t + 4][g] = im;
                    }
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, t = lane & 3;
                        double* d = &acc[(((size_t)lane * MT + mt) * 4 + j) * 4];
                        for (int k = 0; k < 8; ++k) {
                            d[0] += A[g][k] * Bm[k][2 * t]; d[1] += A[g][k] * Bm[k][2 * t + 1];
                            d[2] += A[g + 8][k] * Bm[k][2 * t]; d[3] += A[g + 8][k] * Bm[k][2 * t + 1];
                        }
                    }
                }
        for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            for (int mt = 0; mt < MT; ++mt) {
                const int b = 8 * mt + g;
                if (b >= B) continue;
                for (int q = 0; q < 2; ++q) {
                    const double* E = &acc[(((size_t)lane * MT + mt) * 4 + 2 * q) * 4];
                    const double* O = &acc[(((size_t)lane * MT + mt) * 4 + 2 * q + 1) * 4];
                    const int n = n_base + 16 * q + 2 * t;
                    float* row = beam + (size_t)b * N * 2;
                    if (n < N) { row[2 * n] = (float)E[0]; row[2 * n + 1] = (float)E[2]; row[2 * n + 2] = (float)O[0]; row[2 * n + 3] = (float)O[2]; }
                    if (n + 8 < N) { row[2 * (n + 8)] = (float)E[1]; row[2 * (n + 8) + 1] = (float)E[3]; row[2 * (n + 8) + 2] = (float)O[1]; row[2 * (n + 8) + 3] = (float)O[3]; }
                }
            }
        }
    }
    return 0;
}

int emul_cfar4_pitch(int need) { return cfar4_pitch(need); }

int emul_pc_narrow(const float* line, int N, int seg_start0, const float* fir, int nfir, int fir_delay, int ngates,
                   float* out_line) {
    cf* o = reinterpret_cast<cf*>(out_line);
    const cf* y = reinterpret_cast<const cf*>(line);
    for (int g = 0; g < ngates; ++g) {
        o[g] = pc_narrow_gate(y, N, seg_start0, fir, nfir, fir_delay, g);
        if (ngates + fir_delay <= N - seg_start0) {           // the shared-memory variant must agree exactly
            const cf alt = pc_narrow_gate_smem(y + seg_start0, fir, nfir, fir_delay, g);
            if (alt.x != o[g].x || alt.y != o[g].y) return -1;
        }
    }
    return 0;
}

// Doppler FFT of a [P][TG] tile (pulse-major, as it is read from the pc cube), window applied here.
int emul_mtd_tile(const float* x, int P, int TG, const float* win, float* out /* [TG][P] */, int* radices_out) {
    DopplerPlan dp;
    if (TG != RSP_MTD_TG || !make_doppler_plan(dp, P)) return -1;
    const cf* xi = reinterpret_cast<const cf*>(x);
    std::vector<cf> s((size_t)P * (TG + 1));
    std::vector<float> w(P);
    for (int p = 0; p < P; ++p) w[p] = win[p] * ((p & 1) ? -1.f : 1.f);
    const int* ip = dp.iperm.data();
    switch (P) {
#define X(p, a, b, c) case p: run_mtd_tile<MtdCfg<p, a, b, c>>(s.data(), dp.tw.data(), xi, (size_t)TG, ip, w.data()); break;
        X(8, 8, 1, 1) X(16, 16, 1, 1) X(32, 8, 4, 1) X(64, 8, 8, 1) X(128, 16, 8, 1) X(256, 16, 16, 1) X(512, 8, 8, 8)
#undef X
        default: return -3;
    }
    cf* o = reinterpret_cast<cf*>(out);
    for (int gl = 0; gl < TG; ++gl)
        for (int row = 0; row < P; ++row) o[(size_t)gl * P + row] = s[(size_t)row * (TG + 1) + gl];
    if (radices_out)
        for (int i = 0; i < 4; ++i) radices_out[i] = i < 3 ? dp.r[i] : 0;
    return 0;
}

// generic Doppler DFT (any P = R * Q): the work items of mtd_dft_kernel<TG, R>
int emul_mtd_dft_tile(const float* x /* [P][TG] */, int P, int TG, const float* win, float* out /* [TG][P] */) {
    const int R = (P % 8 == 0) ? 8 : (P % 4 == 0) ? 4 : (P % 2 == 0) ? 2 : 1, Q = P / R;
    const cf* xi = reinterpret_cast<const cf*>(x);
    std::vector<cf> xin((size_t)P * (TG + 1)), xout((size_t)P * (TG + 1)), stw(P);
    for (int m = 0; m < P; ++m) {
        const double ang = -2.0 * kPi * m / P;
        stw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    for (int p = 0; p < P; ++p)
        for (int gl = 0; gl < TG; ++gl) xin[(size_t)p * (TG + 1) + gl] = cscale(xi[(size_t)p * TG + gl], win[p]);
    for (int k = 0; k < Q; ++k)
        for (int gl = 0; gl < TG; ++gl) {
            if (R == 8) mtd_dft_item<8>(xin.data(), xout.data(), stw.data(), P, TG, k, gl);
            else if (R == 4) mtd_dft_item<4>(xin.data(), xout.data(), stw.data(), P, TG, k, gl);
            else if (R == 2) mtd_dft_item<2>(xin.data(), xout.data(), stw.data(), P, TG, k, gl);
            else mtd_dft_item<1>(xin.data(), xout.data(), stw.data(), P, TG, k, gl);
        }
    cf* o = reinterpret_cast<cf*>(out);
    for (int gl = 0; gl < TG; ++gl)
        for (int row = 0; row < P; ++row) o[(size_t)gl * P + row] = xout[(size_t)row * (TG + 1) + gl];
    return 0;
}

// the k-tiled work items of mtd_dft_kernel<TG, R, KT> (KT = 4 or 11 bins per item)
int emul_mtd_dft_tile_kt(const float* x /* [P][TG] */, int P, int TG, int KT, const float* win, float* out /* [TG][P] */) {
    const int R = (P % 8 == 0) ? 8 : (P % 4 == 0) ? 4 : (P % 2 == 0) ? 2 : 1, Q = P / R;
    if (!(KT == 4 || ((KT == 6 || KT == 11) && R <= 4))) return -1;
    const cf* xi = reinterpret_cast<const cf*>(x);
    std::vector<cf> xin((size_t)P * (TG + 1)), xout((size_t)P * (TG + 1), make_float2(-7.f, -7.f)), stw(P);
    for (int m = 0; m < P; ++m) {
        const double ang = -2.0 * kPi * m / P;
        stw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    for (int p = 0; p < P; ++p)
        for (int gl = 0; gl < TG; ++gl) xin[(size_t)p * (TG + 1) + gl] = cscale(xi[(size_t)p * TG + gl], win[p]);
    const int groups = (Q + KT - 1) / KT;
    for (int kg = 0; kg < groups; ++kg)
        for (int gl = 0; gl < TG; ++gl) {
#define RSP_EMUL_KT(r, kt) if (R == r && KT == kt) mtd_dft_item_kt<r, kt>(xin.data(), xout.data(), stw.data(), P, TG, kg * KT, gl);
            RSP_EMUL_KT(8, 4) RSP_EMUL_KT(4, 4) RSP_EMUL_KT(2, 4) RSP_EMUL_KT(1, 4)
            RSP_EMUL_KT(4, 11) RSP_EMUL_KT(2, 11) RSP_EMUL_KT(1, 11) RSP_EMUL_KT(4, 6) RSP_EMUL_KT(2, 6) RSP_EMUL_KT(1, 6)
#undef RSP_EMUL_KT
        }
    cf* o = reinterpret_cast<cf*>(out);
    for (int gl = 0; gl < TG; ++gl)
        for (int row = 0; row < P; ++row) o[(size_t)gl * P + row] = xout[(size_t)row * (TG + 1) + gl];
    return 0;
}

// the folded (even / odd) work items of mtd_dft_kernel<TG, R, 100 + KP> for odd Q = P / R
int emul_mtd_dft_tile_sym(const float* x /* [P][TG] */, int P, int TG, int KP, const float* win, float* out /* [TG][P] */) {
    const int R = (P % 8 == 0) ? 8 : (P % 4 == 0) ? 4 : (P % 2 == 0) ? 2 : 1, Q = P / R;
    if (R > 4 || !(Q & 1) || !(KP == 3 || KP == 4 || KP == 6)) return -1;
    const cf* xi = reinterpret_cast<const cf*>(x);
    std::vector<cf> xin((size_t)P * (TG + 1)), xout((size_t)P * (TG + 1), make_float2(-7.f, -7.f)), stw(P);
    for (int m = 0; m < P; ++m) {
        const double ang = -2.0 * kPi * m / P;
        stw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    for (int p = 0; p < P; ++p)
        for (int gl = 0; gl < TG; ++gl) xin[(size_t)p * (TG + 1) + gl] = cscale(xi[(size_t)p * TG + gl], win[p]);
    for (int t = 0; t < RSP_MTD_THREADS; ++t) mtd_dft_fold_phase(xin.data(), P, R, TG, t, RSP_MTD_THREADS);
    if (TG != 8) return -3;                                   // the row pitch is a template parameter: the test uses tiles of 8 gates
    // as the in-place kernel runs it: every item computes into its registers, barrier, every item stores over the input tile
    const int groups = mtd_dft_sym_groups(Q, KP);
    const bool inplace = mtd_dft_sym_inplace(Q, KP, TG, RSP_MTD_THREADS);
    cf* dst = inplace ? xin.data() : xout.data();
    auto run = [&](auto r_tag, auto kp_tag) {
        constexpr int RR = decltype(r_tag)::value, KK = decltype(kp_tag)::value;
        struct Regs { cf A[KK][RR], B[KK][RR]; int kk[KK]; };
        std::vector<Regs> regs((size_t)groups * TG);
        for (int e = 0; e < groups * TG; ++e)
            mtd_dft_sym_compute<RR, KK, 9>(xin.data(), stw.data(), P, (e / TG) * KK, e % TG, regs[e].A, regs[e].B, regs[e].kk);
        for (int e = 0; e < groups * TG; ++e)
            mtd_dft_sym_store<RR, KK, 9>(dst, stw.data(), P, (e / TG) * KK, e % TG, regs[e].A, regs[e].B, regs[e].kk);
    };
#define RSP_EMUL_SYM(r, kp) if (R == r && KP == kp) run(std::integral_constant<int, r>{}, std::integral_constant<int, kp>{});
    RSP_EMUL_SYM(4, 3) RSP_EMUL_SYM(2, 3) RSP_EMUL_SYM(1, 3) RSP_EMUL_SYM(4, 4) RSP_EMUL_SYM(2, 4) RSP_EMUL_SYM(1, 4)
    RSP_EMUL_SYM(4, 6) RSP_EMUL_SYM(2, 6) RSP_EMUL_SYM(1, 6)
#undef RSP_EMUL_SYM
    if (inplace) xout = xin;
    cf* o = reinterpret_cast<cf*>(out);
    for (int gl = 0; gl < TG; ++gl)
        for (int row = 0; row < P; ++row) o[(size_t)gl * P + row] = xout[(size_t)row * (TG + 1) + gl];
    return 0;
}

// CFAR over a full sum map S[G][P] (one pair) using the two tile phases on tiles of TG gates.
int emul_cfar_map(const float* S, int G, int P, int guard_r, int guard_v, int ref_r, int ref_v, float t_cfar, int TG,
                  unsigned char* det /* [G][P] */) {
    CfarParams c;
    c.P = P; c.G = G; c.guard_r = guard_r; c.guard_v = guard_v; c.ref_r = ref_r; c.ref_v = ref_v; c.t_cfar = t_cfar;
    const int mR = guard_r + ref_r, mV = guard_v + ref_v;
    const int rows = TG + 2 * mR;
    std::memset(det, 0, (size_t)G * P);
    std::vector<float> tile((size_t)rows * P), R5((size_t)cfar_r5_rows(c, TG) * P), D5((size_t)TG * P);
    for (int g_first = mR; g_first < G - mR; g_first += TG) {
        for (int row = 0; row < rows; ++row)
            for (int v = 0; v < P; ++v) {
                const int g = g_first - mR + row;
                tile[(size_t)row * P + v] = g < G ? S[(size_t)g * P + v] : 0.f;
            }
        for (int t = 0; t < RSP_CFAR_THREADS; ++t)
            cfar_sums_phase(tile.data(), R5.data(), D5.data(), c, TG, t, RSP_CFAR_THREADS);
        for (int gl = 0; gl < TG && g_first + gl < G - mR; ++gl)
            for (int v = mV; v < P - mV; ++v) {
                float cut;
                if (cfar_decide(tile.data(), R5.data(), D5.data(), c, gl, v, &cut)) det[(size_t)(g_first + gl) * P + v] = 1;
            }
    }
    return 0;
}

// Vectorised CFAR (P % 4 == 0): padded tile + quad phases, as cfar4_kernel runs them.
int emul_cfar4_map(const float* S, int G, int P, int guard_r, int guard_v, int ref_r, int ref_v, float t_cfar, int TG,
                   int use_template, unsigned char* det /* [G][P] */) {
    if (P % 4) return -1;
    CfarParams c;
    c.P = P; c.G = G; c.guard_r = guard_r; c.guard_v = guard_v; c.ref_r = ref_r; c.ref_v = ref_v; c.t_cfar = t_cfar;
    const Cfar4Geom g = cfar4_geom(c, TG);
    const int mR = guard_r + ref_r, mV = guard_v + ref_v;
    std::memset(det, 0, (size_t)G * P);
    std::vector<float> tile((size_t)g.rows * g.PP, 0.f), R5((size_t)g.r5_rows * g.RP);
    for (int g_first = mR; g_first < G - mR; g_first += TG) {
        std::fill(tile.begin(), tile.end(), 0.f);
        for (int row = 0; row < g.rows; ++row) {
            const int gg = g_first - mR + row;
            if (gg < G)
                for (int v = 0; v < P; ++v) tile[(size_t)row * g.PP + RSP_CFAR_HALO + v] = S[(size_t)gg * P + v];
        }
        for (int t = 0; t < RSP_CFAR_THREADS; ++t) {
            if (use_template && ref_r == 5) cfar4_r5_phase<5>(tile.data(), R5.data(), c, g, t, RSP_CFAR_THREADS);
            else cfar4_r5_phase<0>(tile.data(), R5.data(), c, g, t, RSP_CFAR_THREADS);
        }
        const int c_lo = mV / 4, c_hi = (P - mV - 1) / 4;
        for (int gl = 0; gl < TG && g_first + gl < G - mR; ++gl)
            for (int c4 = c_lo; c4 <= c_hi; ++c4) {
                float cut[4];
                unsigned m;
                if (use_template && ref_r == 5 && ref_v == 5 && guard_v == 10)
                    m = cfar4_decide_quad<5, 5, 10>(tile.data(), R5.data(), c, g, gl, c4, cut);
                else if (use_template && ref_r == 5 && ref_v == 4 && guard_v == 2)
                    m = cfar4_decide_quad<5, 4, 2>(tile.data(), R5.data(), c, g, gl, c4, cut);
                else m = cfar4_decide_quad<0, 0, 0>(tile.data(), R5.data(), c, g, gl, c4, cut);
                for (int j = 0; j < 4; ++j)
                    if (m & (1u << j)) det[(size_t)(g_first + gl) * P + 4 * c4 + j] = 1;
            }
    }
    return 0;
}

// Marching CFAR (cfar5_kernel): uninitialised pad columns (NaN here), chunk-major items, range test first, queued Doppler test.
int emul_cfar5_map(const float* S, int G, int P, int guard_r, int guard_v, int ref_r, int ref_v, float t_cfar, int TG,
                   unsigned char* det /* [G][P] */) {
    if (P % 4 || ref_r != 5 || !((ref_v == 5 && guard_v == 10) || (ref_v == 4 && guard_v == 2))) return -1;
    CfarParams c;
    c.P = P; c.G = G; c.guard_r = guard_r; c.guard_v = guard_v; c.ref_r = ref_r; c.ref_v = ref_v; c.t_cfar = t_cfar;
    constexpr int CR = RSP_CFAR5_CR;
    const int mR = guard_r + ref_r, mV = guard_v + ref_v, rows = TG + 2 * mR;
    const int pitch = cfar5_pitch(P, mV), nq = cfar5_nq(P, mV), c_lo = mV / 4;
    if (((CR * (pitch / 4) - nq) & 7) || pitch < P + 2 * RSP_CFAR5_HALO) return -2;
    std::memset(det, 0, (size_t)G * P);
    std::vector<float> tile((size_t)rows * pitch);
    for (int g_first = mR; g_first < G - mR; g_first += TG) {
        std::fill(tile.begin(), tile.end(), std::nanf(""));
        for (int row = 0; row < rows; ++row) {
            const int gg = g_first - mR + row;
            for (int v = 0; v < P; ++v) tile[(size_t)row * pitch + RSP_CFAR5_HALO + v] = gg < G ? S[(size_t)gg * P + v] : 0.f;
        }
        const int gl_end = std::min(TG, G - mR - g_first);
        const int n_items = ((gl_end + CR - 1) / CR) * nq;
        std::vector<unsigned> queue;
        for (int it = 0; it < n_items; ++it) {
            const int ch = it / nq, c4 = c_lo + (it - ch * nq), gl0 = ch * CR;
            auto hit = [&](int s, unsigned m) { queue.push_back(((unsigned)(gl0 + s) << 16) | ((unsigned)c4 << 4) | m); };
            if (ref_v == 5) cfar5_march<5, 15, CR>(tile.data(), pitch, c, gl0, std::min(CR, gl_end - gl0), c4, hit);
            else cfar5_march<5, 6, CR>(tile.data(), pitch, c, gl0, std::min(CR, gl_end - gl0), c4, hit);
        }
        const float kv = t_cfar / (float)ref_v;
        for (unsigned w : queue) {
            const int gl = (int)(w >> 16), c4 = (int)((w >> 4) & 0xFFFu);
            const float* row0 = tile.data() + (size_t)(gl + mR) * pitch + RSP_CFAR5_HALO;
            float4 cq;
            const unsigned m = ref_v == 5 ? cfar5_doppler<5, 10>(row0 + 4 * c4, kv, w & 15u, &cq) : cfar5_doppler<4, 2>(row0 + 4 * c4, kv, w & 15u, &cq);
            for (int j = 0; j < 4; ++j)
                if (m & (1u << j)) det[(size_t)(g_first + gl) * P + 4 * c4 + j] = 1;
        }
    }
    return 0;
}

double emul_spline5_peak(const double* y, int os) { return rsp_spline5_peak(y, os); }

int emul_digit_reverse(int f, int L, const int* radices, int nrad) { return rsp_digit_reverse(f, L, radices, nrad); }

}  // extern "C"
