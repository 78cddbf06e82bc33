// rsp_math.cuh -- host/device math shared by the kernels and by the host-emulation test.
//
// Everything here is plain C++ marked RSP_HD so that the *same* butterfly, FFT-pass, spline and
// CFAR code that runs inside the sm_100a kernels can also be compiled by g++ (no GPU) and checked
// against NumPy in tests/test_host_emulation.py.  No CUDA intrinsics in this file.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RSP_HD __host__ __device__ __forceinline__
#include <cuda_runtime.h>
#else
#define RSP_HD inline
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

typedef float2 cf;

RSP_HD cf cadd(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
RSP_HD cf csub(cf a, cf b) { return make_float2(a.x - b.x, a.y - b.y); }
RSP_HD cf cmul(cf a, cf b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
RSP_HD cf cmulc(cf a, cf b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
RSP_HD cf cscale(cf a, float s) { return make_float2(a.x * s, a.y * s); }

// Multiply by the unit twiddle of a FORWARD transform rotated a quarter turn:
// SIGN = -1 (forward, e^{-i..}): multiply by -i ; SIGN = +1 (inverse): multiply by +i.
template <int SIGN> RSP_HD cf mul_quarter(cf a) {
    return SIGN < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// Multiply by (wr + i*SIGN_adjusted wi) where (wr, wi) is the FORWARD twiddle e^{-i theta}.
template <int SIGN> RSP_HD cf mul_tw(cf a, float wr, float wi) {
    // forward: a * (wr + i wi); inverse: a * (wr - i wi)
    return SIGN < 0 ? make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr)
                    : make_float2(a.x * wr + a.y * wi, a.y * wr - a.x * wi);
}

#define RSP_SQRT1_2 0.70710678118654752440f
#define RSP_COS_PI_8 0.92387953251128675613f
#define RSP_SIN_PI_8 0.38268343236508977173f

// ---------------------------------------------------------------------------------------------
// In-register DFTs of length 2/4/8/16, natural order in and out.
// SIGN = -1: X[k] = sum_n x[n] e^{-2 pi i nk/R};  SIGN = +1: conjugate kernel (unnormalised inverse).
// ---------------------------------------------------------------------------------------------
template <int SIGN> RSP_HD void dft2(cf& a, cf& b) {
    cf t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

template <int SIGN> RSP_HD void dft4(cf& x0, cf& x1, cf& x2, cf& x3) {
    cf t0 = cadd(x0, x2), t1 = csub(x0, x2);
    cf t2 = cadd(x1, x3), t3 = mul_quarter<SIGN>(csub(x1, x3));
    x0 = cadd(t0, t2);
    x1 = cadd(t1, t3);
    x2 = csub(t0, t2);
    x3 = csub(t1, t3);
}

// y0 = a + w b, y1 = a - w b with w = (wr, wi) the FORWARD twiddle (SIGN > 0: its conjugate), as 4 + 2 fused multiply-adds:
// the product goes straight into the sum and the difference is 2 a - y0.  Two instructions fewer than multiply, add, subtract.
#ifndef RSP_FFT_FMA
#define RSP_FFT_FMA 1
#endif
template <int SIGN> RSP_HD void fma_pm(cf a, cf b, float wr, float wi, cf& y0, cf& y1) {
    if (SIGN < 0) {
        y0.x = fmaf(b.x, wr, fmaf(-b.y, wi, a.x));
        y0.y = fmaf(b.x, wi, fmaf(b.y, wr, a.y));
    } else {
        y0.x = fmaf(b.x, wr, fmaf(b.y, wi, a.x));
        y0.y = fmaf(b.y, wr, fmaf(-b.x, wi, a.y));
    }
    y1.x = fmaf(2.f, a.x, -y0.x);
    y1.y = fmaf(2.f, a.y, -y0.y);
}
// DFT-4 of (x0, w1 x1, w2 x2, w3 x3); T0 = false: x0 is multiplied by w0 as well
template <int SIGN, bool T0> RSP_HD void dft4_tw(cf& x0, cf& x1, cf& x2, cf& x3, cf w0, cf w1, cf w2, cf w3) {
    const cf a0 = T0 ? x0 : mul_tw<SIGN>(x0, w0.x, w0.y);
    cf t0, t1, t2, t3;
    fma_pm<SIGN>(a0, x2, w2.x, w2.y, t0, t1);
    const cf u = mul_tw<SIGN>(x1, w1.x, w1.y);
    fma_pm<SIGN>(u, x3, w3.x, w3.y, t2, t3);
    t3 = mul_quarter<SIGN>(t3);
    x0 = cadd(t0, t2);
    x1 = cadd(t1, t3);
    x2 = csub(t0, t2);
    x3 = csub(t1, t3);
}

template <int R, int SIGN> struct SmallDft;

template <int SIGN> struct SmallDft<2, SIGN> {
    static RSP_HD void run(cf* v) { dft2<SIGN>(v[0], v[1]); }
};
template <int SIGN> struct SmallDft<4, SIGN> {
    static RSP_HD void run(cf* v) { dft4<SIGN>(v[0], v[1], v[2], v[3]); }
};
template <int SIGN> struct SmallDft<8, SIGN> {
    // n = 2a + b (a<4, b<2), k = c + 4d (c<4, d<2)
    static RSP_HD void run(cf* v) {
        dft4<SIGN>(v[0], v[2], v[4], v[6]);   // b = 0 -> T0[c] in v[0],v[2],v[4],v[6]
        dft4<SIGN>(v[1], v[3], v[5], v[7]);   // b = 1 -> T1[c] in v[1],v[3],v[5],v[7]
        // T1[c] *= W8^c
        v[3] = mul_tw<SIGN>(v[3], RSP_SQRT1_2, -RSP_SQRT1_2);
        v[5] = mul_quarter<SIGN>(v[5]);
        v[7] = mul_tw<SIGN>(v[7], -RSP_SQRT1_2, -RSP_SQRT1_2);
        cf o[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            o[c] = cadd(v[2 * c], v[2 * c + 1]);
            o[c + 4] = csub(v[2 * c], v[2 * c + 1]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o[i];
    }
};
template <int SIGN> struct SmallDft<16, SIGN> {
    // n = 4a + b, k = c + 4d (a,b,c,d < 4)
    static RSP_HD void run(cf* v) {
#pragma unroll
        for (int b = 0; b < 4; ++b) dft4<SIGN>(v[b], v[4 + b], v[8 + b], v[12 + b]);   // t[b][c] at v[4c+b]
        run_second_stage(v);
    }
    // t[b][c] *= W16^{bc}, then for every c a DFT-4 over b
    static RSP_HD void run_second_stage(cf* v) {
#if RSP_FFT_FMA
        cf o[16];
        {   // c = 0: no twiddles
            cf a0 = v[0], a1 = v[1], a2 = v[2], a3 = v[3];
            dft4<SIGN>(a0, a1, a2, a3);
            o[0] = a0; o[4] = a1; o[8] = a2; o[12] = a3;
        }
        {   // c = 1: W16^1, W16^2, W16^3 folded into the butterfly
            cf a0 = v[4], a1 = v[5], a2 = v[6], a3 = v[7];
            dft4_tw<SIGN, true>(a0, a1, a2, a3, make_float2(1.f, 0.f), make_float2(RSP_COS_PI_8, -RSP_SIN_PI_8),
                                make_float2(RSP_SQRT1_2, -RSP_SQRT1_2), make_float2(RSP_SIN_PI_8, -RSP_COS_PI_8));
            o[1] = a0; o[5] = a1; o[9] = a2; o[13] = a3;
        }
        {   // c = 2: W16^2, W16^4 = -i (free), W16^6
            cf a0 = v[8], a2 = mul_quarter<SIGN>(v[10]);
            const cf u = mul_tw<SIGN>(v[9], RSP_SQRT1_2, -RSP_SQRT1_2);
            cf t0 = cadd(a0, a2), t1 = csub(a0, a2), t2, t3;
            fma_pm<SIGN>(u, v[11], -RSP_SQRT1_2, -RSP_SQRT1_2, t2, t3);
            t3 = mul_quarter<SIGN>(t3);
            o[2] = cadd(t0, t2); o[6] = cadd(t1, t3); o[10] = csub(t0, t2); o[14] = csub(t1, t3);
        }
        {   // c = 3: W16^3, W16^6, W16^9
            cf a0 = v[12], a1 = v[13], a2 = v[14], a3 = v[15];
            dft4_tw<SIGN, true>(a0, a1, a2, a3, make_float2(1.f, 0.f), make_float2(RSP_SIN_PI_8, -RSP_COS_PI_8),
                                make_float2(-RSP_SQRT1_2, -RSP_SQRT1_2), make_float2(-RSP_COS_PI_8, RSP_SIN_PI_8));
            o[3] = a0; o[7] = a1; o[11] = a2; o[15] = a3;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = o[i];
#else
        v[4 * 1 + 1] = mul_tw<SIGN>(v[4 * 1 + 1], RSP_COS_PI_8, -RSP_SIN_PI_8);      // e=1
        v[4 * 1 + 2] = mul_tw<SIGN>(v[4 * 1 + 2], RSP_SQRT1_2, -RSP_SQRT1_2);        // e=2
        v[4 * 1 + 3] = mul_tw<SIGN>(v[4 * 1 + 3], RSP_SIN_PI_8, -RSP_COS_PI_8);      // e=3
        v[4 * 2 + 1] = mul_tw<SIGN>(v[4 * 2 + 1], RSP_SQRT1_2, -RSP_SQRT1_2);        // e=2
        v[4 * 2 + 2] = mul_quarter<SIGN>(v[4 * 2 + 2]);                              // e=4
        v[4 * 2 + 3] = mul_tw<SIGN>(v[4 * 2 + 3], -RSP_SQRT1_2, -RSP_SQRT1_2);       // e=6
        v[4 * 3 + 1] = mul_tw<SIGN>(v[4 * 3 + 1], RSP_SIN_PI_8, -RSP_COS_PI_8);      // e=3
        v[4 * 3 + 2] = mul_tw<SIGN>(v[4 * 3 + 2], -RSP_SQRT1_2, -RSP_SQRT1_2);       // e=6
        v[4 * 3 + 3] = mul_tw<SIGN>(v[4 * 3 + 3], -RSP_COS_PI_8, RSP_SIN_PI_8);      // e=9
        cf o[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            cf a0 = v[4 * c], a1 = v[4 * c + 1], a2 = v[4 * c + 2], a3 = v[4 * c + 3];
            dft4<SIGN>(a0, a1, a2, a3);                                             // over b -> d
            o[c] = a0; o[c + 4] = a1; o[c + 8] = a2; o[c + 12] = a3;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = o[i];
#endif
    }
    // the same transform of (v[0], w[1] v[1], ..., w[15] v[15]): the input twiddles of a DIT pass folded into the first stage
    static RSP_HD void run_twiddled(cf* v, const cf* w) {
#if RSP_FFT_FMA
        dft4_tw<SIGN, true>(v[0], v[4], v[8], v[12], make_float2(1.f, 0.f), w[4], w[8], w[12]);
#pragma unroll
        for (int b = 1; b < 4; ++b) dft4_tw<SIGN, false>(v[b], v[4 + b], v[8 + b], v[12 + b], w[b], w[4 + b], w[8 + b], w[12 + b]);
        run_second_stage(v);
#else
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = mul_tw<SIGN>(v[k], w[k].x, w[k].y);
        run(v);
#endif
    }
};

// ---------------------------------------------------------------------------------------------
// Index helpers for the in-place mixed-radix transforms.
//
// DIF pass (sub-length Ls, radix R), butterfly q in [0, L/R):
//   blk = q / (Ls/R), j = q % (Ls/R), elements at blk*Ls + j + m*(Ls/R), m < R.
//   y_k = (sum_m x_m w_R^{mk}) * W_Ls^{jk}, stored at the same slots indexed by k.
// After all DIF passes (radices r1..rk) frequency f = k1 + r1*(k2 + r2*(...)) sits at position
//   pos = k1*(L/r1) + k2*(L/(r1 r2)) + ... + kk                              (digit reversal).
// DIT pass = exact inverse structure: y_k * W_Ls^{-+jk} then butterfly, run with Ls small -> large.
// ---------------------------------------------------------------------------------------------
RSP_HD int rsp_digit_reverse(int f, int L, const int* radices, int nrad) {
    int pos = 0, stride = L;
    for (int s = 0; s < nrad; ++s) {
        int r = radices[s];
        stride /= r;
        pos += (f % r) * stride;
        f /= r;
    }
    return pos;
}

// Shared-memory padding: one complex of padding every 16 keeps every pass of the in-place FFT
// bank-conflict free ("PC kernel").
RSP_HD int rsp_pad16(int a) { return a + (a >> 4); }

// Generic DIF butterfly on an array `s` (element address map AddrFn), twiddle table
// tw[(k-1)*(Ls/R) + j] = e^{-2 pi i jk/Ls} (forward sign; SIGN selects conjugation).
template <int R, int SIGN, typename Addr>
RSP_HD void dif_butterfly(cf* s, int Ls, int q, const cf* tw, Addr addr) {
    const int span = Ls / R;
    const int blk = q / span, j = q - blk * span;
    const int base = blk * Ls + j;
    cf v[R];
#pragma unroll
    for (int m = 0; m < R; ++m) v[m] = s[addr(base + m * span)];
    SmallDft<R, SIGN>::run(v);
    s[addr(base)] = v[0];
#pragma unroll
    for (int k = 1; k < R; ++k) {
        if (span == 1) {
            s[addr(base + k)] = v[k];
        } else {
            cf w = tw[(k - 1) * span + j];
            s[addr(base + k * span)] = mul_tw<SIGN>(v[k], w.x, w.y);
        }
    }
}

template <int R, int SIGN, typename Addr>
RSP_HD void dit_butterfly(cf* s, int Ls, int q, const cf* tw, Addr addr) {
    const int span = Ls / R;
    const int blk = q / span, j = q - blk * span;
    const int base = blk * Ls + j;
    cf v[R];
    v[0] = s[addr(base)];
#pragma unroll
    for (int k = 1; k < R; ++k) {
        cf x = s[addr(base + k * span)];
        if (span != 1) {
            cf w = tw[(k - 1) * span + j];
            x = mul_tw<SIGN>(x, w.x, w.y);
        }
        v[k] = x;
    }
    SmallDft<R, SIGN>::run(v);
#pragma unroll
    for (int m = 0; m < R; ++m) s[addr(base + m * span)] = v[m];
}

// ---------------------------------------------------------------------------------------------
// S9 helper: first maximum of the not-a-knot cubic spline through 5 unit-spaced points sampled
// every 1/OS cell (fun_process_single_frame.m:250-260, 265-275; interp1(...,'spline')).
// Not-a-knot with 5 knots: M1 = d1, M3 = d3, M2 = (6 d2 - d1 - d3)/4, M0 = 2M1 - M2, M4 = 2M3 - M2
// where d_i are second differences and M_i the second derivatives at the knots.
// Returns the offset of the maximum from the FIRST point, in cells.
// ---------------------------------------------------------------------------------------------
RSP_HD double rsp_spline5_peak(const double y[5], int os) {
    const double d1 = y[0] - 2.0 * y[1] + y[2];
    const double d2 = y[1] - 2.0 * y[2] + y[3];
    const double d3 = y[2] - 2.0 * y[3] + y[4];
    double M[5];
    M[1] = d1;
    M[3] = d3;
    M[2] = (6.0 * d2 - d1 - d3) * 0.25;
    M[0] = 2.0 * M[1] - M[2];
    M[4] = 2.0 * M[3] - M[2];
    double best = y[0];
    int best_q = 0;
    const int nq = 4 * os;
    for (int q = 1; q <= nq; ++q) {
        int i = q / os;
        if (i > 3) i = 3;
        const double t = (double)(q - i * os) / (double)os;
        const double b = (y[i + 1] - y[i]) - (2.0 * M[i] + M[i + 1]) / 6.0;
        const double val = y[i] + t * (b + t * (0.5 * M[i] + t * ((M[i + 1] - M[i]) / 6.0)));
        if (val > best) { best = val; best_q = q; }
    }
    return (double)best_q / (double)os;
}
