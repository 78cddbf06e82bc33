// rsp_kernels.cuh -- the sm_100a kernels of the per-frame chain.  Each kernel body is a sequence
// of phases from rsp_phases.cuh separated by __syncthreads(); see DESIGN.md for the data layout,
// the roofline that bounds each kernel and its algorithmic bytes.
//
//   matlab_to_pcn_kernel   MATLAB [P,N,C] column-major cube -> device-native raw[p][c][n]
//   dbf_kernel             S5  fun_process_single_frame.m:92-97   raw[p][c][n] -> beam[p][b][n]
//   pc_narrow_kernel       S6  fun_process_single_frame.m:111-112,123
//   pc_fft_kernel          S6  fun_process_single_frame.m:115-125 (overlap-save blocks)
//   mtd_kernel             S7  fun_process_single_frame.m:131-136 (power-of-two P)
//   mtd_dft_kernel         S7  same, any P (the reference's native P = 332)
//   cfar_kernel            S8 + S9  fun_process_single_frame.m:172-223, 241-298
#pragma once
#include <cuda_runtime.h>
#include "rsp.h"
#include "rsp_phases.cuh"

namespace rsp {

// ------------------------------------------------------------------------------------------
// layout / precision conversion (drop-in path only; the throughput path is PCN complex64)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 to_c64(float2 v) { return v; }
__device__ __forceinline__ float2 to_c64(double2 v) { return make_float2((float)v.x, (float)v.y); }

template <typename TIN>
__global__ void __launch_bounds__(256) matlab_to_pcn_kernel(const TIN* __restrict__ in, float2* __restrict__ out,
                                                            int P, int N, int C) {
    __shared__ float2 tile[32][33];
    const int c = blockIdx.z, p0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int n = n0 + i, p = p0 + threadIdx.x;
        if (n < N && p < P) tile[i][threadIdx.x] = to_c64(in[((size_t)c * N + n) * P + p]);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int p = p0 + i, n = n0 + threadIdx.x;
        if (n < N && p < P) out[((size_t)p * C + c) * N + n] = tile[threadIdx.x][i];
    }
}

__global__ void __launch_bounds__(256) c128_to_c64_kernel(const double2* __restrict__ in, float2* __restrict__ out,
                                                          size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = to_c64(in[i]);
}

// MATLAB [P,G,B] column-major (element (p,g,b) at (b*G+g)*P+p) <-> pc[p][b][g] with pitch ldg
template <typename TIN>
__global__ void __launch_bounds__(256) pgb_to_pbg_kernel(const TIN* __restrict__ in, float2* __restrict__ out, int P,
                                                         int G, int B, int ldg) {
    __shared__ float2 tile[32][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, g0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int g = g0 + i, p = p0 + threadIdx.x;
        if (g < G && p < P) tile[i][threadIdx.x] = to_c64(in[((size_t)b * G + g) * P + p]);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int p = p0 + i, g = g0 + threadIdx.x;
        if (g < G && p < P) out[((size_t)p * B + b) * ldg + g] = tile[threadIdx.x][i];
    }
}

// pc[p][b][g] (pitch ldg) -> MATLAB-ordered [b][g][p] complex64
__global__ void __launch_bounds__(256) pbg_to_bgp_kernel(const float2* __restrict__ in, float2* __restrict__ out, int P,
                                                         int G, int B, int ldg) {
    __shared__ float2 tile[32][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, g0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int p = p0 + i, g = g0 + threadIdx.x;
        if (g < G && p < P) tile[i][threadIdx.x] = in[((size_t)p * B + b) * ldg + g];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int g = g0 + i, p = p0 + threadIdx.x;
        if (g < G && p < P) out[((size_t)b * G + g) * P + p] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------
// S5: digital beamforming.  beam[p][b][n] = sum_c raw[p][c][n] * conj(W[b][c]).
// One thread owns SPT range samples (strided by 256 so every load is a coalesced 8-byte lane
// access with no alignment requirement -- the native N = 5819 is odd) and all NB beams in
// registers; the conjugated weights sit in shared memory and are read as warp broadcasts.
// HBM-bound: reads 8*C bytes, writes 8*B bytes per sample, 8*C*B flops.
// ------------------------------------------------------------------------------------------
template <int NB, int SPT>
__global__ void __launch_bounds__(256) dbf_kernel(const float2* __restrict__ raw, float2* __restrict__ beam,
                                                  const float2* __restrict__ Wc /* [C][NB] conj(W) */, int C, int N,
                                                  int ldb) {
    __shared__ float2 sW[RSP_MAX_CHANNELS * NB];
    const int tid = threadIdx.x;
    for (int i = tid; i < C * NB; i += 256) sW[i] = Wc[i];
    __syncthreads();
    const int p = blockIdx.y;
    const int n0 = blockIdx.x * (256 * SPT) + tid;
    const float2* rp = raw + (size_t)p * C * N;
    float2 acc[SPT][NB];
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[k][b] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
        float2 x[SPT];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int n = n0 + k * 256;
            x[k] = (n < N) ? __ldcs(rp + (size_t)c * N + n) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float2 w = sW[c * NB + b];
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
                acc[k][b].x = fmaf(x[k].x, w.x, fmaf(-x[k].y, w.y, acc[k][b].x));
                acc[k][b].y = fmaf(x[k].x, w.y, fmaf(x[k].y, w.x, acc[k][b].y));
            }
        }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int n = n0 + k * 256;
            if (n < N) beam[((size_t)p * NB + b) * ldb + n] = acc[k][b];
        }
}

// ------------------------------------------------------------------------------------------
// S6: pulse compression
// ------------------------------------------------------------------------------------------
struct PcKernelArgs {
    const float2* beam;
    float2* pc;
    const float2* tw1;
    const float2* tw2;
    const float2* H;
    int N, ldb, ldg, B;
    int seg_start0, taps, gate0, g_end, valid;
};

template <int R1>
__global__ void __launch_bounds__(RSP_PC_THREADS) pc_fft_kernel(const PcKernelArgs k) {
    extern __shared__ float2 pc_smem[];
    const int tid = threadIdx.x;
    const size_t line = (size_t)blockIdx.z * k.B + blockIdx.y;
    PcBlockArgs a;
    a.line = k.beam + line * k.ldb;
    a.out_line = k.pc + line * k.ldg;
    a.tw1 = k.tw1;
    a.tw2 = k.tw2;
    a.H = k.H;
    a.N = k.N;
    a.seg_start0 = k.seg_start0;
    a.taps = k.taps;
    a.g0 = k.gate0 + blockIdx.x * k.valid;
    a.g_end = k.g_end;
    pc_phase_load_pass1<R1>(a, pc_smem, tid);
    __syncthreads();
    pc_phase_pass2<R1>(a, pc_smem, tid);
    __syncthreads();
    pc_phase_mid<R1>(a, pc_smem, tid);
    __syncthreads();
    pc_phase_ipass2<R1>(a, pc_smem, tid);
    __syncthreads();
    pc_phase_ipass1_store<R1>(a, pc_smem, tid);
}

__global__ void __launch_bounds__(256) pc_narrow_kernel(const float2* __restrict__ beam, float2* __restrict__ pc,
                                                        const float* __restrict__ fir, int nfir, int fir_delay, int N,
                                                        int ldb, int ldg, int B, int seg_start0, int ngates) {
    __shared__ float sfir[256];
    for (int i = threadIdx.x; i < nfir; i += 256) sfir[i] = fir[i];
    __syncthreads();
    const size_t line = (size_t)blockIdx.y * B + blockIdx.x;
    const float2* y = beam + line * ldb;
    for (int g = threadIdx.x; g < ngates; g += 256)
        pc[line * ldg + g] = pc_narrow_gate(y, N, seg_start0, sfir, nfir, fir_delay, g);
}

// ------------------------------------------------------------------------------------------
// S7: MTD.  Tile = TG gates x P pulses of one beam.  The corner turn happens here: rows of the
// pc cube (range-contiguous) are read coalesced, the FFT runs along the strided pulse dimension
// with lanes spread over gates (bank-conflict free for any stride), and the Doppler lines leave
// transposed, contiguous in Doppler, exactly MATLAB's rdm_13beam(v,g,b) byte order.
// ------------------------------------------------------------------------------------------
struct MtdArgs {
    const float2* pc;
    float2* rdm;
    float* amp;
    const float* win;      // [P]; (-1)^p folded in for the power-of-two kernel
    const float2* tw;      // pow2: per-pass twiddles; dft: e^{-2 pi i m/P}, m < P
    const int* perm;       // pow2 only
    MtdPlan plan;
    int tw_count;
    int B, G, ldg;
};

template <int TG>
__global__ void __launch_bounds__(RSP_MTD_THREADS) mtd_kernel(const MtdArgs k) {
    extern __shared__ float2 mtd_smem[];
    const int P = k.plan.P;
    float2* tile = mtd_smem;
    float2* stw = tile + (size_t)P * (TG + 1);
    float* swin = reinterpret_cast<float*>(stw + k.tw_count);
    int* sperm = reinterpret_cast<int*>(swin + P);
    const int tid = threadIdx.x;
    for (int i = tid; i < k.tw_count; i += RSP_MTD_THREADS) stw[i] = k.tw[i];
    for (int i = tid; i < P; i += RSP_MTD_THREADS) {
        swin[i] = k.win[i];
        sperm[i] = k.perm[i];
    }
    __syncthreads();
    const int g0 = blockIdx.x * TG, b = blockIdx.y;
    for (int e = tid; e < P * TG; e += RSP_MTD_THREADS) {
        const int p = e / TG, gl = e - p * TG, g = g0 + gl;
        float2 x = make_float2(0.f, 0.f);
        if (g < k.G) x = k.pc[((size_t)p * k.B + b) * k.ldg + g];
        tile[sperm[p] * (TG + 1) + gl] = cscale(x, swin[p]);
    }
    __syncthreads();
    for (int pass = k.plan.nrad - 1; pass >= 0; --pass) {
        mtd_dit_pass<-1>(tile, k.plan, pass, stw, TG, tid, RSP_MTD_THREADS);
        __syncthreads();
    }
    for (int e = tid; e < TG * P; e += RSP_MTD_THREADS) {
        const int gl = e / P, row = e - gl * P, g = g0 + gl;
        if (g < k.G) {
            const float2 v = tile[row * (TG + 1) + gl];
            const size_t o = ((size_t)b * k.G + g) * P + row;
            __stcs(k.rdm + o, v);
            k.amp[o] = sqrtf(fmaf(v.x, v.x, v.y * v.y));
        }
    }
}

template <int TG>
__global__ void __launch_bounds__(RSP_MTD_THREADS) mtd_dft_kernel(const MtdArgs k) {
    extern __shared__ float2 mtd_smem[];
    const int P = k.plan.P;
    float2* xin = mtd_smem;
    float2* xout = xin + (size_t)P * (TG + 1);
    float2* stw = xout + (size_t)P * (TG + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < P; i += RSP_MTD_THREADS) stw[i] = k.tw[i];
    const int g0 = blockIdx.x * TG, b = blockIdx.y;
    for (int e = tid; e < P * TG; e += RSP_MTD_THREADS) {
        const int p = e / TG, gl = e - p * TG, g = g0 + gl;
        float2 x = make_float2(0.f, 0.f);
        if (g < k.G) x = k.pc[((size_t)p * k.B + b) * k.ldg + g];
        xin[p * (TG + 1) + gl] = cscale(x, k.win[p]);
    }
    __syncthreads();
    const int half = P / 2;
    for (int e = tid; e < P * TG; e += RSP_MTD_THREADS) {
        const int f = e / TG, gl = e - f * TG;
        float2 acc = make_float2(0.f, 0.f);
        int idx = 0;
        for (int p = 0; p < P; ++p) {
            const float2 w = stw[idx];
            const float2 x = xin[p * (TG + 1) + gl];
            acc.x = fmaf(x.x, w.x, fmaf(-x.y, w.y, acc.x));
            acc.y = fmaf(x.x, w.y, fmaf(x.y, w.x, acc.y));
            idx += f;
            if (idx >= P) idx -= P;
        }
        int row = f + half;                 // fftshift(.,1): bin f moves to (f + floor(P/2)) mod P
        if (row >= P) row -= P;
        xout[row * (TG + 1) + gl] = acc;
    }
    __syncthreads();
    for (int e = tid; e < TG * P; e += RSP_MTD_THREADS) {
        const int gl = e / P, row = e - gl * P, g = g0 + gl;
        if (g < k.G) {
            const float2 v = xout[row * (TG + 1) + gl];
            const size_t o = ((size_t)b * k.G + g) * P + row;
            __stcs(k.rdm + o, v);
            k.amp[o] = sqrtf(fmaf(v.x, v.x, v.y * v.y));
        }
    }
}

// ------------------------------------------------------------------------------------------
// S8 + S9: GOCA-CFAR on S = |rdm_b| + |rdm_{b+1}|, then per detection the spline refinement and
// the monopulse angle, fused into the compaction (one atomic per detection; detections are rare).
// ------------------------------------------------------------------------------------------
struct CfarArgs {
    const float* amp;            // [B][G][P]
    const float2* rdm;           // [B][G][P]
    CfarParams c;
    int* count;                  // detection counter of this CPI slot
    rsp_detection* recs;         // records of this CPI slot
    int cap;
    const double* range_axis;
    const double* vel_axis;
    const double* beam_angles;
    const double* k_slopes;
    double delta_r, delta_v;
    int complex_mode;
};

template <int TG>
__global__ void __launch_bounds__(RSP_CFAR_THREADS) cfar_kernel(const CfarArgs k) {
    extern __shared__ float cfar_smem[];
    const int P = k.c.P, G = k.c.G;
    const int mR = k.c.guard_r + k.c.ref_r, mV = k.c.guard_v + k.c.ref_v;
    const int rows = TG + 2 * mR;
    const int pair = blockIdx.y;
    const int g_first = mR + blockIdx.x * TG;
    const int tid = threadIdx.x;
    const float* A = k.amp + (size_t)pair * G * P;
    const float* Bm = A + (size_t)G * P;
    const size_t base = (size_t)(g_first - mR) * P;
    const int n_el = rows * P;
    for (int e = tid; e < n_el; e += RSP_CFAR_THREADS) {
        const int row = e / P;
        float s = 0.f;
        if (g_first - mR + row < G) s = A[base + e] + Bm[base + e];
        cfar_smem[e] = s;
    }
    __syncthreads();
    for (int e = tid; e < TG * P; e += RSP_CFAR_THREADS) {
        const int gl = e / P, v = e - gl * P, g = g_first + gl;
        if (g >= G - mR || v < mV || v >= P - mV) continue;
        float cut;
        if (!cfar_cut(cfar_smem, P, k.c, gl, v, &cut)) continue;
        const int slot = atomicAdd(k.count, 1);
        if (slot >= k.cap) continue;        // overflow is reported by the host from the count
        double yr[5], yv[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            yr[i] = (double)cfar_smem[(gl + mR - 2 + i) * P + v];
            yv[i] = (double)cfar_smem[(gl + mR) * P + v - 2 + i];
        }
        const double r_off = rsp_spline5_peak(yr, 8) - 2.0;          // fsf:237 rInterpTimes = 8
        const double v_off = rsp_spline5_peak(yv, 4) - 2.0;          // vInterpTimes = 4
        rsp_detection d;
        d.v_idx = v + 1;
        d.r_idx = g + 1;
        d.pair_idx = pair + 1;
        d.power = cut;
        d.range = k.range_axis[g] + r_off * k.delta_r;               // fsf:262
        d.velocity = k.vel_axis[v] + v_off * k.delta_v;              // fsf:278
        const size_t o = ((size_t)pair * G + g) * P + v;
        double ratio;
        const double eps = 2.220446049250313e-16;
        if (k.complex_mode) {                                        // mc:454-461
            const float2 a = k.rdm[o], bb = k.rdm[o + (size_t)G * P];
            const double nr = (double)a.x - (double)bb.x, ni = (double)a.y - (double)bb.y;
            const double dr = (double)a.x + (double)bb.x + eps, di = (double)a.y + (double)bb.y;
            ratio = (nr * dr + ni * di) / (dr * dr + di * di);
        } else {                                                     // fsf:282-285
            const double sa = (double)A[(size_t)g * P + v], sb = (double)Bm[(size_t)g * P + v];
            ratio = (sa - sb) / (sa + sb + eps);
        }
        d.angle = 0.5 * (k.beam_angles[pair] + k.beam_angles[pair + 1]) + k.k_slopes[pair] * ratio;   // fsf:286-290
        k.recs[slot] = d;
    }
}

}  // namespace rsp
