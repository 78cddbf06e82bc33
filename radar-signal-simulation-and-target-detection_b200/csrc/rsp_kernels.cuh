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
//   cfar_kernel/cfar4_kernel  S8  fun_process_single_frame.m:172-223 (records cell + the 12 values S9 needs)
//   refine_kernel             S9  fun_process_single_frame.m:241-298, one launch per batch of CPIs
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

// Drop dead intermediates from L2 without writing them back: once the consumer kernel of a buffer has
// finished, its dirty lines would otherwise be evicted to HBM later (measured: 69 MB of the 166 MB of
// DRAM traffic per CPI at config 2).  discard.global.L2 makes the contents undefined, which is exactly
// what a dead buffer is.  Used on the stream path only (the single-CPI path keeps its intermediates
// readable for rsp_get_beam / rsp_get_pc / rsp_get_amp).
struct DiscardArgs {
    void* ptr;            // 128-byte aligned, or nullptr
    size_t bytes;
};
__device__ __forceinline__ void l2_discard(const DiscardArgs& d) {
    if (!d.ptr) return;
    const size_t lines = d.bytes >> 7;
    const size_t gthreads = (size_t)gridDim.x * gridDim.y * gridDim.z * blockDim.x * blockDim.y;
    const size_t gtid = ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (blockDim.x * blockDim.y) +
                        threadIdx.y * blockDim.x + threadIdx.x;
    char* base = static_cast<char*>(d.ptr);
    for (size_t i = gtid; i < lines; i += gthreads) asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + (i << 7)) : "memory");
}

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
// S4 + S4.1: echo synthesis and noise (fun_process_single_frame.m:47-88).  One CTA owns one
// (pulse, channel) line: it first fills the line with Philox noise, then adds every target's delayed
// pulse (only the three non-zero stretches of tx_pulse are visited), so no atomics are needed.
// ------------------------------------------------------------------------------------------
struct SynthTarget {          // per target, prepared on the host in fp64
    int delay;                // delay_samples (fsf:56); targets with delay outside (0, N) are dropped (fsf:66)
    float amp;                // sqrt(SNR / P_signal_unscaled) (fsf:61-63)
    // Phase steps as 0.64 fixed-point fractions of a cycle, so that the phase of pulse p, channel c is the
    // wrapping integer dop_fix * p + steer_fix * c: exact argument reduction without fp64 on the device.
    unsigned long long dop_fix;     // frac(doppler_freq * prt) * 2^64        (cycles per pulse, fsf:57-58)
    unsigned long long steer_fix;   // frac(element_spacing * sin(El) / wavelength) * 2^64   (cycles per channel, fsf:163-169)
};
// exp(j 2 pi * phase / 2^64): the top 24 bits go through sincospif, the next 24 correct it to first order
__device__ __forceinline__ float2 unit_phasor(unsigned long long phase) {
    const float hi = (float)(unsigned)(phase >> 40) * (1.0f / 16777216.0f);                          // [0, 1), exact
    const float lo = (float)(unsigned)((phase >> 16) & 0xFFFFFFu) * (1.0f / 281474976710656.0f);     // < 2^-24
    float sn, cs;
    sincospif(2.0f * hi, &sn, &cs);
    const float d = 6.283185307179586f * lo;
    return make_float2(cs - d * sn, sn + d * cs);
}
struct SynthArgs {
    float2* raw;              // [P][C][N]
    const float2* tx;         // [N]
    const SynthTarget* tg;
    int n_targets, P, C, N;
    int seg_lo[3], seg_hi[3]; // non-zero stretches of tx (0-based, half open)
    float noise_sigma;        // sqrt(noise_power / 2) per component; 0 = no noise
    unsigned long long seed;
    uint2 round_key[10];      // Philox key schedule of `seed` (host-computed so the kernel reads it as constants)
    int tx_lo, tx_hi;         // hull of the non-zero stretches
};

// z += x * p (complex), the one statement every echo-synthesis path accumulates a target with.  Written with rounding-mode
// intrinsics so that the compiler cannot contract it differently from one kernel to the next: the staged, the gathering and
// the fused (S4 inside S5) kernels must produce the same bits for the same frame.
__device__ __forceinline__ void synth_cmac(float& zr, float& zi, const float2 x, const float2 p) {
    zr = __fadd_rn(zr, __fmaf_rn(x.x, p.x, -__fmul_rn(x.y, p.y)));
    zi = __fadd_rn(zi, __fmaf_rn(x.x, p.y, __fmul_rn(x.y, p.x)));
}

// Philox4x32-10 (Salmon et al., SC'11); round_key[r] = key + r * (0x9E3779B9, 0xBB67AE85)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, const uint2 (&round_key)[10]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ round_key[r].x, lo1, hi0 ^ ctr.w ^ round_key[r].y, lo0);
    }
    return ctr;
}
// two uniform words -> one complex standard-normal sample (Box-Muller): z = sqrt(-2 ln u1) * exp(j 2 pi u2),
// u1 = ((a >> 9) + 1/2) 2^-23, u2 = (b >> 9) 2^-23.
// Logarithm, square root, sine and cosine are the MUFU approximations (lg2: absolute 2^-22.6, sqrt: relative
// 2^-23, sin / cos: absolute 2^-21 on [-pi, pi], evaluated at 2 pi (u2 - 1/2) and negated): the radius is
// off by about 1e-7 / r, which only shows for the 0.1 % of samples with r < 0.05.  An accurate logf costs a
// third of the whole generator and buys nothing for white Gaussian noise.
__device__ __forceinline__ float2 box_muller(unsigned int a, unsigned int b) {
    // 23 random bits under the exponent of 1.0 give 1 + k 2^-23; no integer-to-float conversion (XU pipe, the
    // kernel's bottleneck together with the four MUFUs below)
    const float u1 = __uint_as_float(0x3F800000u | (a >> 9)) - (1.0f - 0.5f / 8388608.0f);   // (k + 1/2) 2^-23 in (0, 1)
    const float h2 = __uint_as_float(0x3F800000u | (b >> 9)) - 1.5f;                          // k 2^-23 - 1/2 in [-1/2, 1/2)
    float l2, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));   // -2 ln 2 * log2(u1)
    const float x = 6.283185307179586f * h2;                               // [-pi, pi)
    return make_float2(-r * __cosf(x), -r * __sinf(x));
}

// One CTA per (channel, pulse, chunk of RSP_SYNTH_CHUNK range samples).  The chunk is built in shared
// memory -- Philox noise first, then every target's delayed pulse added over the three non-zero stretches of
// tx_pulse that fall inside the chunk (targets are serialised by a barrier because their echoes may
// overlap) -- and leaves with one coalesced store, so the cube is written exactly once.
#define RSP_SYNTH_CHUNK 4096
#define RSP_SYNTH_TB 64
__global__ void __launch_bounds__(256) synth_kernel(const __grid_constant__ SynthArgs k) {
    __shared__ __align__(16) float2 s[RSP_SYNTH_CHUNK];
    __shared__ float2 s_ph[RSP_SYNTH_TB];
    __shared__ int s_delay[RSP_SYNTH_TB];
    const int c = blockIdx.x, p = blockIdx.y, tid = threadIdx.x;
    const int c0 = blockIdx.z * RSP_SYNTH_CHUNK;                            // even
    const int cn = min(RSP_SYNTH_CHUNK, k.N - c0);
    const size_t line_id = (size_t)p * k.C + c;
    // noise: counter = (pair index within the line, line id), 2 complex samples per Philox call
    for (int i = tid; 2 * i < cn; i += 256) {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k.noise_sigma > 0.f) {
            const uint4 r = philox4x32_10(make_uint4((unsigned)(c0 / 2 + i), (unsigned)line_id, (unsigned)(line_id >> 32), 0u), k.round_key);
            const float2 z0 = box_muller(r.x, r.y), z1 = box_muller(r.z, r.w);
            z = make_float4(z0.x * k.noise_sigma, z0.y * k.noise_sigma, z1.x * k.noise_sigma, z1.y * k.noise_sigma);
        }
        *reinterpret_cast<float4*>(&s[2 * i]) = z;                          // the odd tail sample of an odd chunk is never stored
    }
    for (int t0 = 0; t0 < k.n_targets; t0 += RSP_SYNTH_TB) {
        const int nb = min(RSP_SYNTH_TB, k.n_targets - t0);
        __syncthreads();                                                    // noise written / previous batch consumed
        if (tid < nb) {
            const SynthTarget tg = k.tg[t0 + tid];
            // phasor = amp * exp(j 2 pi (dop_cycles * p + steer_cycles * c)), argument reduced in integers
            const float2 u = unit_phasor(tg.dop_fix * (unsigned long long)p + tg.steer_fix * (unsigned long long)c);
            s_ph[tid] = make_float2(tg.amp * u.x, tg.amp * u.y);
            s_delay[tid] = tg.delay;
        }
        __syncthreads();
        bool dirty = false;                                                 // CTA-uniform: an earlier target wrote this chunk
        for (int t = 0; t < nb; ++t) {
            const int delay = s_delay[t];
            const int len = min(k.N, k.N - delay);                          // fsf:67
            int lo[3], hi[3];
            bool any = false;
#pragma unroll
            for (int sgi = 0; sgi < 3; ++sgi) {
                lo[sgi] = max(k.seg_lo[sgi], c0 - delay);
                hi[sgi] = min(min(k.seg_hi[sgi], len), c0 + cn - delay);
                any |= lo[sgi] < hi[sgi];
            }
            if (!any) continue;                                             // this echo misses the chunk: no work, no barrier
            if (dirty) __syncthreads();                                     // echoes may overlap: one target at a time
            dirty = true;
            const float2 ph = s_ph[t];
#pragma unroll
            for (int sgi = 0; sgi < 3; ++sgi)
                for (int i = lo[sgi] + tid; i < hi[sgi]; i += 256) {
                    const float2 x = k.tx[i];
                    float2 v = s[delay + i - c0];
                    synth_cmac(v.x, v.y, x, ph);
                    s[delay + i - c0] = v;
                }
        }
    }
    __syncthreads();
    float2* out = k.raw + line_id * k.N + c0;
    if ((reinterpret_cast<size_t>(out) & 15) == 0) {
        for (int i = tid; 2 * i + 1 < cn; i += 256) reinterpret_cast<float4*>(out)[i] = *reinterpret_cast<const float4*>(&s[2 * i]);
        if ((cn & 1) && tid == 0) out[cn - 1] = s[cn - 1];
    } else {
        for (int i = tid; i < cn; i += 256) out[i] = s[i];
    }
}

// The same cube for frames with at most RSP_SYNTH_GATHER_T targets (the reference's scenes have 1 to 5): no
// shared-memory staging and no barriers per target.  A thread makes two noise samples, then for every
// target tests whether the pair falls inside the hull of the delayed pulse train (one compare; about a
// third of the line does) and only then gathers tx_pulse -- which is zero between the pulses, so no
// per-stretch test is needed.  Targets are added in index order, like the staged kernel.
#define RSP_SYNTH_GATHER_T 8
#define RSP_SYNTH_FUSED_T 64             // targets per frame of the fused S4 + S5 kernel = RSP_MAX_FRAME_TARGETS (BASELINE config 4)
#define RSP_SYNTH_GATHER_CHUNK 8192
// Samples n0 (even) and n0 + 1 of line `line_id` = pulse * C + channel: Philox noise, then the targets in index order.
// ph[t * ph_stride] is target t's phasor for this (pulse, channel).  Statement for statement the per-pair body of
// synth_gather_kernel (kept in its measured form above), used by the fused dbf_synth_kernel.
// act0 / act1: bit t (t + 32) set = target t can reach the caller's sample range (dbf_synth_kernel culls per warp: with 64
// targets only the ~11 % whose non-zero pulse stretches overlap the warp's 16-32 samples are visited; a culled target would
// have added exact zeros, so the result is the same).
__device__ __forceinline__ float4 synth_pair(const SynthArgs& k, const float2* ph, int ph_stride, const int* delay, int n0, size_t line_id,
                                             unsigned act0 = 0xFFFFFFFFu, unsigned act1 = 0xFFFFFFFFu) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k.noise_sigma > 0.f) {
        const uint4 r = philox4x32_10(make_uint4((unsigned)(n0 / 2), (unsigned)line_id, (unsigned)(line_id >> 32), 0u), k.round_key);
        const float2 z0 = box_muller(r.x, r.y), z1 = box_muller(r.z, r.w);
        z = make_float4(z0.x * k.noise_sigma, z0.y * k.noise_sigma, z1.x * k.noise_sigma, z1.y * k.noise_sigma);
    }
    const unsigned span = (unsigned)(k.tx_hi - k.tx_lo);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        unsigned m = h == 0 ? act0 : act1;
        if (h == 0 ? k.n_targets < 32 : k.n_targets < 64) m &= (1u << (k.n_targets - 32 * h > 0 ? k.n_targets - 32 * h : 0)) - 1u;
        while (m) {                                                         // ascending target index, like the plain loop
            const int t = 32 * h + __ffs(m) - 1;
            m &= m - 1;
            const int j = n0 - delay[t] - k.tx_lo;                          // tx index of sample n0, relative to the hull
            if ((unsigned)(j + 1) <= span) {                                // sample n0 or n0 + 1 is inside the delayed hull
                const float2 p = ph[t * ph_stride];
                if ((unsigned)j < span) {
                    const float2 x = k.tx[j + k.tx_lo];
                    synth_cmac(z.x, z.y, x, p);
                }
                if ((unsigned)(j + 1) < span) {
                    const float2 x = k.tx[j + 1 + k.tx_lo];
                    synth_cmac(z.z, z.w, x, p);
                }
            }
        }
    }
    return z;
}
// The same samples for the KS channels c = c0 + 4 s of one thread of dbf_synth_kernel at once: the hull test, the delay and the
// two tx_pulse samples of a target do not depend on the channel, so they are fetched once per target and applied to every
// channel with that channel's phasor (per channel: the statements of synth_pair in the same order, hence the same bits).
// With 64 targets per frame the target loop is most of the fused kernel; this form runs a third fewer instructions in it.
template <int KS>
__device__ __forceinline__ void synth_pair_channels(const SynthArgs& k, const float2* ph /* phasor of channel c0 */, int ph_stride, const int* delay,
                                                    int n0, size_t line0 /* pulse * C + c0 */, int c0, int C, unsigned act0, unsigned act1,
                                                    float4 (&z)[KS]) {
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        z[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k.noise_sigma > 0.f && c0 + 4 * s < C) {
            const size_t line_id = line0 + 4 * s;
            const uint4 r = philox4x32_10(make_uint4((unsigned)(n0 / 2), (unsigned)line_id, (unsigned)(line_id >> 32), 0u), k.round_key);
            const float2 z0 = box_muller(r.x, r.y), z1 = box_muller(r.z, r.w);
            z[s] = make_float4(z0.x * k.noise_sigma, z0.y * k.noise_sigma, z1.x * k.noise_sigma, z1.y * k.noise_sigma);
        }
    }
    const unsigned span = (unsigned)(k.tx_hi - k.tx_lo);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        unsigned m = h == 0 ? act0 : act1;
        if (h == 0 ? k.n_targets < 32 : k.n_targets < 64) m &= (1u << (k.n_targets - 32 * h > 0 ? k.n_targets - 32 * h : 0)) - 1u;
        while (m) {                                                         // ascending target index, like the plain loop
            const int t = 32 * h + __ffs(m) - 1;
            m &= m - 1;
            const int j = n0 - delay[t] - k.tx_lo;                          // tx index of sample n0, relative to the hull
            if ((unsigned)(j + 1) <= span) {                                // sample n0 or n0 + 1 is inside the delayed hull
                const bool in0 = (unsigned)j < span, in1 = (unsigned)(j + 1) < span;
                float2 x0 = make_float2(0.f, 0.f), x1 = x0;
                if (in0) x0 = k.tx[j + k.tx_lo];
                if (in1) x1 = k.tx[j + 1 + k.tx_lo];
                const float2* pt = ph + t * ph_stride;
#pragma unroll
                for (int s = 0; s < KS; ++s) {
                    if (c0 + 4 * s < C) {
                        const float2 p = pt[4 * s];
                        if (in0) {
                            synth_cmac(z[s].x, z[s].y, x0, p);
                        }
                        if (in1) {
                            synth_cmac(z[s].z, z[s].w, x1, p);
                        }
                    }
                }
            }
        }
    }
}
__global__ void __launch_bounds__(256) synth_gather_kernel(const __grid_constant__ SynthArgs k) {
    __shared__ float2 s_ph[RSP_SYNTH_GATHER_T];
    __shared__ int s_delay[RSP_SYNTH_GATHER_T];
    const int c = blockIdx.x, p = blockIdx.y, tid = threadIdx.x;
    const int c0 = blockIdx.z * RSP_SYNTH_GATHER_CHUNK;                     // even
    const int cn = min(RSP_SYNTH_GATHER_CHUNK, k.N - c0);
    const size_t line_id = (size_t)p * k.C + c;
    if (tid < k.n_targets) {
        const SynthTarget tg = k.tg[tid];
        const float2 u = unit_phasor(tg.dop_fix * (unsigned long long)p + tg.steer_fix * (unsigned long long)c);
        s_ph[tid] = make_float2(tg.amp * u.x, tg.amp * u.y);
        s_delay[tid] = tg.delay;
    }
    __syncthreads();
    float2* out = k.raw + line_id * k.N + c0;
    const bool aligned = (reinterpret_cast<size_t>(out) & 15) == 0;
    const unsigned span = (unsigned)(k.tx_hi - k.tx_lo);
    for (int i = tid; 2 * i < cn; i += 256) {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k.noise_sigma > 0.f) {
            const uint4 r = philox4x32_10(make_uint4((unsigned)(c0 / 2 + i), (unsigned)line_id, (unsigned)(line_id >> 32), 0u), k.round_key);
            const float2 z0 = box_muller(r.x, r.y), z1 = box_muller(r.z, r.w);
            z = make_float4(z0.x * k.noise_sigma, z0.y * k.noise_sigma, z1.x * k.noise_sigma, z1.y * k.noise_sigma);
        }
        const int n0 = c0 + 2 * i;
        for (int t = 0; t < k.n_targets; ++t) {
            const int j = n0 - s_delay[t] - k.tx_lo;                        // tx index of sample n0, relative to the hull
            if ((unsigned)(j + 1) <= span) {                                // sample n0 or n0 + 1 is inside the delayed hull
                const float2 ph = s_ph[t];
                if ((unsigned)j < span) {
                    const float2 x = k.tx[j + k.tx_lo];
                    synth_cmac(z.x, z.y, x, ph);
                }
                if ((unsigned)(j + 1) < span) {
                    const float2 x = k.tx[j + 1 + k.tx_lo];
                    synth_cmac(z.z, z.w, x, ph);
                }
            }
        }
        if (aligned && 2 * i + 1 < cn) {
            reinterpret_cast<float4*>(out)[i] = z;
        } else {
            out[2 * i] = make_float2(z.x, z.y);
            if (2 * i + 1 < cn) out[2 * i + 1] = make_float2(z.z, z.w);
        }
    }
}

// ------------------------------------------------------------------------------------------
// S5: digital beamforming.  beam[p][b][n] = sum_c raw[p][c][n] * conj(W[b][c]).
// One thread owns SPT range samples (strided by the CTA width so every load is a coalesced 8-byte
// lane access with no alignment requirement -- the native N = 5819 is odd) and all NB beams in
// registers; the conjugated weights sit in shared memory and are read as warp broadcasts.  The
// channel loop is software-pipelined CU channels deep (loads of group i+1 are issued before the
// FMAs of group i) and CTAs are small (128 threads) so that enough bytes are in flight per SM.
// HBM-bound: reads 8*C bytes, writes 8*B bytes per sample, 8*C*B flops.
// ------------------------------------------------------------------------------------------
#define RSP_DBF_THREADS 128
template <int NB, int SPT, int CU>
__global__ void __launch_bounds__(RSP_DBF_THREADS) dbf_kernel(const float2* __restrict__ raw, float2* __restrict__ beam,
                                                              const float2* __restrict__ Wc /* [C][NB] conj(W) */,
                                                              int C, int N, int ldb, int* __restrict__ det_count,
                                                              const DiscardArgs dead) {
    __shared__ float2 sW[RSP_MAX_CHANNELS * NB];
    const int tid = threadIdx.x;
    l2_discard(dead);
    if (det_count && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *det_count = 0;   // first kernel of the CPI
    for (int i = tid; i < C * NB; i += RSP_DBF_THREADS) sW[i] = Wc[i];
    __syncthreads();
    const int p = blockIdx.y;
    const int n0 = blockIdx.x * (RSP_DBF_THREADS * SPT) + tid;
    const float2* rp = raw + (size_t)p * C * N;
    float2 acc[SPT][NB];
#pragma unroll
    for (int k = 0; k < SPT; ++k)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[k][b] = make_float2(0.f, 0.f);
    float2 cur[CU][SPT], nxt[CU][SPT];
    auto load_group = [&](float2 (&dst)[CU][SPT], int c0) {
#pragma unroll
        for (int u = 0; u < CU; ++u)
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
                const int n = n0 + k * RSP_DBF_THREADS, c = c0 + u;
                dst[u][k] = (n < N && c < C) ? __ldcs(rp + (size_t)c * N + n) : make_float2(0.f, 0.f);
            }
    };
    load_group(cur, 0);
    for (int c0 = 0; c0 < C; c0 += CU) {
        if (c0 + CU < C) load_group(nxt, c0 + CU);
#pragma unroll
        for (int u = 0; u < CU; ++u) {
            if (c0 + u < C) {
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float2 w = sW[(c0 + u) * NB + b];
#pragma unroll
                    for (int k = 0; k < SPT; ++k) {
                        acc[k][b].x = fmaf(cur[u][k].x, w.x, fmaf(-cur[u][k].y, w.y, acc[k][b].x));
                        acc[k][b].y = fmaf(cur[u][k].x, w.y, fmaf(cur[u][k].y, w.x, acc[k][b].y));
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < CU; ++u)
#pragma unroll
            for (int k = 0; k < SPT; ++k) cur[u][k] = nxt[u][k];
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
            const int n = n0 + k * RSP_DBF_THREADS;
            if (n < N) beam[((size_t)p * NB + b) * ldb + n] = acc[k][b];
        }
}

// ------------------------------------------------------------------------------------------
// S5 on the tensor cores (C <= 32, B <= 16): the complex contraction as a real GEMM
//   out[n][2b + {re,im}] = sum_{c,{re,im}} X[n][(c,.)] * Bm[(c,.)][2b + .]      (K = 2C, N = 2B)
// with mma.sync m16n8k8 TF32 and 3xTF32 error compensation (x = xh + xl, w = wh + wl;
// xh*wh + xh*wl + xl*wh, fp32 accumulate), which keeps fp32-level accuracy (the <= 1e-4 RDM tolerance
// rules out plain TF32).  K order inside a k-step of 8 is chosen so that a thread's two A registers of
// a row are exactly the (re, im) of one float2 load: k = t -> (channel 4s + t, re), k = t + 4 -> (.., im).
// One warp owns 32 consecutive range samples (two m16 tiles) of one pulse; all its loads are issued
// up front.  Weight fragments {b0h, b1h, b0l, b1l} are prebuilt on the host, one float4 per
// (k-step, n-tile, lane), and read from shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

#define RSP_DBF_MMA_THREADS 128
// VEC = true (N even): fragment row g <-> sample n0 + 2g, row g+8 <-> sample n0 + 2g + 1 (the row order
// of an MMA tile is free), so a thread's four A registers of a tile are ONE float4 load and its four D
// registers ONE float4 store: 8 lanes x 16 B = 128 contiguous bytes per channel / beam row.
// VEC = false (odd N, e.g. the native 5819): rows g, g+8 <-> samples n0 + g, n0 + g + 8, float2 accesses.
template <int NT, int KS, bool VEC>
__global__ void __launch_bounds__(RSP_DBF_MMA_THREADS) dbf_mma_kernel(const float2* __restrict__ raw,
                                                                      float2* __restrict__ beam,
                                                                      const float4* __restrict__ Wfrag /* [KS][NT][32] */,
                                                                      int C, int NB, int N, int ldb,
                                                                      int* __restrict__ det_count, const DiscardArgs dead) {
    __shared__ float4 sW[KS * NT * 32];
    const int tid = threadIdx.x;
    l2_discard(dead);
    if (det_count && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *det_count = 0;   // first kernel of the CPI
    for (int i = tid; i < KS * NT * 32; i += RSP_DBF_MMA_THREADS) sW[i] = Wfrag[i];
    const int lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int p = blockIdx.y;
    const int n_base = (blockIdx.x * (RSP_DBF_MMA_THREADS / 32) + w) * 32;      // 32 samples per warp = 2 tiles
    const float2* rp = raw + (size_t)p * C * N;
    // x[s][m] = {a0, a2, a1, a3} of tile m, k-step s:  (re, im) of the thread's two rows
    float4 x[KS][2];
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        const int c = 4 * s + t;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C) {
                if (VEC) {
                    const int n = n_base + 16 * m + 2 * g;
                    if (n < N) v = __ldcs(reinterpret_cast<const float4*>(rp + (size_t)c * N + n));
                } else {
                    const int n = n_base + 16 * m + g;
                    if (n < N) { const float2 a = __ldcs(rp + (size_t)c * N + n); v.x = a.x; v.y = a.y; }
                    if (n + 8 < N) { const float2 a = __ldcs(rp + (size_t)c * N + n + 8); v.z = a.x; v.w = a.y; }
                }
            }
            x[s][m] = v;
        }
    }
    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const float v[4] = {x[s][m].x, x[s][m].z, x[s][m].y, x[s][m].w};   // a0 (re,row0) a1 (re,row1) a2 (im,row0) a3 (im,row1)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                // hi = x with the 13 low mantissa bits cleared (a valid tf32), lo = x - hi (exact in
                // fp32, |lo| < 2^-10 |x|), lo truncated the same way: 3 instructions per value
                // (cvt.rna.tf32.f32 expands to ~6 on sm_100a).  Dropped terms are O(2^-20 |x w|).
                ah[m][i] = __float_as_uint(v[i]) & 0xFFFFE000u;
                al[m][i] = __float_as_uint(v[i] - __uint_as_float(ah[m][i])) & 0xFFFFE000u;
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float4 wf = sW[(s * NT + nt) * 32 + lane];
            const uint32_t b0h = __float_as_uint(wf.x), b1h = __float_as_uint(wf.y);
            const uint32_t b0l = __float_as_uint(wf.z), b1l = __float_as_uint(wf.w);
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                mma_tf32(acc[m][nt], al[m], b0h, b1h);
                mma_tf32(acc[m][nt], ah[m], b0l, b1l);
                mma_tf32(acc[m][nt], ah[m], b0h, b1h);
            }
        }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int b = 4 * nt + t;
        if (b < NB) {
            float2* row = beam + ((size_t)p * NB + b) * ldb;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (VEC) {
                    const int n = n_base + 16 * m + 2 * g;
                    if (n < N)
                        *reinterpret_cast<float4*>(row + n) = make_float4(acc[m][nt][0], acc[m][nt][1], acc[m][nt][2], acc[m][nt][3]);
                } else {
                    const int n = n_base + 16 * m + g;
                    if (n < N) row[n] = make_float2(acc[m][nt][0], acc[m][nt][1]);
                    if (n + 8 < N) row[n + 8] = make_float2(acc[m][nt][2], acc[m][nt][3]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// S5 on the tensor cores, weights as the A operand (N even): D[16 x 8] = A[16 x 8] * B[8 x 8] with
//   A rows  r < 8 -> Re of beam 8 mt + r,  r + 8 -> Im of the same beam;   k = t -> (channel 4s + t, Re x), t + 4 -> (.., Im x)
//   B cols  = 8 range samples.
// A thread's B fragment {b0, b1} of a k-step is then exactly the (re, im) pair of ONE complex sample of ONE
// channel, in load order: a float4 load (samples n, n + 1) feeds the "even" and the "odd" n-tile of a 16-sample
// group without a single register move (the data-as-A kernel above needs the order x, z, y, w and pays ~150
// moves per warp for it, a third of its instructions; on sm_100a a move costs an issue slot like an FFMA).
// The D fragments hold Re (row g) and Im (row g + 8) of beam g for samples 4t .. 4t + 3 of the group, i.e. 32
// contiguous bytes of the beam row per thread, 128 per quad.  Weight fragments (hi, lo) come straight from
// global memory through L1 (4 KB, read by every warp), so the kernel has no shared memory and no barrier.
// 3xTF32 as above: Wl*xh + Wh*xl + Wh*xh with xh = x (the MMA ignores the 13 low mantissa bits), xl = x - trunc(x).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_wa(float (&d)[4], const float4& a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
                   "r"(b0), "r"(b1));
}

// NQ = 16-sample groups per warp tile (2: 32 samples; 1 for the big shapes, C > 16 or B > 8, whose 32-sample tile would
// need 64 registers of loads + 32 accumulators: 178 registers and 8 warps per SM at config 3).  One tile per warp: software
// pipelining, several tiles per warp, split launches and L1::no_allocate loads were all measured neutral or slower
// (profiles/README.md) and are gone.
template <int MT, int KS, int NQ>
__global__ void __launch_bounds__(RSP_DBF_MMA_THREADS, ((MT == 1 && KS <= 4) || NQ == 1 ? (MT * KS * NQ <= 8 ? 8 : 5) : 1)) dbf_mma2_kernel(const float2* __restrict__ raw, float2* __restrict__ beam,
                                                                       const float4* __restrict__ Wa /* [KS][MT][2][32] */,
                                                                       int C, int NB, int N, int ldb,
                                                                       int* __restrict__ det_count, const DiscardArgs dead) {
    const int tid = threadIdx.x;
    l2_discard(dead);
    if (det_count && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *det_count = 0;   // first kernel of the CPI
    const int lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int p = blockIdx.y;
    const int n_base = (blockIdx.x * (RSP_DBF_MMA_THREADS / 32) + w) * (16 * NQ);
    if (n_base >= N) return;
    // column g of the even n-tile of a 16-sample group <-> sample sg = g (g even) or g + 7 (g odd), odd n-tile: sg + 1.
    // With this order the D fragments of lane t are samples 2t, 2t+1 and 2t+8, 2t+9, so each of the two 16-byte stores
    // of a quad is 64 contiguous bytes (full 32-byte sectors) instead of four 16-byte pieces 32 bytes apart.
    const int sg = (g & 1) ? g + 7 : g;
    const float2* rp = raw + (size_t)p * C * N + sg + (unsigned)(t * N);                 // channel t of k-step 0
    const unsigned cstep = 4u * (unsigned)N;                                           // k-step s: channel 4s + t
    float4 x[KS][NQ];
#pragma unroll
    for (int s = 0; s < KS; ++s) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            x[s][q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * s + t < C && n_base + 16 * q + sg < N) x[s][q] = __ldcs(reinterpret_cast<const float4*>(rp + s * cstep + n_base + 16 * q));
        }
    }
    float acc[MT][2 * NQ][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2 * NQ; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][j][i] = 0.f;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        // n-tile j = 2q + parity: {b0, b1} = (re, im) of sample 16q + sg + parity
        uint32_t bh[2 * NQ][2], bl[2 * NQ][2];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float v[4] = {x[s][q].x, x[s][q].y, x[s][q].z, x[s][q].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t hi = __float_as_uint(v[i]) & 0xFFFFE000u;
                bh[2 * q + (i >> 1)][i & 1] = hi;
                bl[2 * q + (i >> 1)][i & 1] = __float_as_uint(v[i] - __uint_as_float(hi)) & 0xFFFFE000u;
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const float4 ah = __ldg(Wa + ((s * MT + mt) * 2 + 0) * 32 + lane);
            const float4 al = __ldg(Wa + ((s * MT + mt) * 2 + 1) * 32 + lane);
#pragma unroll
            for (int j = 0; j < 2 * NQ; ++j) {
                mma_tf32_wa(acc[mt][j], al, bh[j][0], bh[j][1]);
                mma_tf32_wa(acc[mt][j], ah, bl[j][0], bl[j][1]);
                mma_tf32_wa(acc[mt][j], ah, bh[j][0], bh[j][1]);
            }
        }
    }
    float2* const brow = beam + (size_t)p * NB * ldb + 2 * t;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int b = 8 * mt + g;
        if (b < NB) {
            float2* row = brow + (size_t)b * ldb + n_base;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float(&E)[4] = acc[mt][2 * q];
                const float(&O)[4] = acc[mt][2 * q + 1];
                const int n = n_base + 16 * q + 2 * t;
                if (n < N) *reinterpret_cast<float4*>(row + 16 * q) = make_float4(E[0], E[2], O[0], O[2]);
                if (n + 8 < N) *reinterpret_cast<float4*>(row + 16 * q + 8) = make_float4(E[1], E[3], O[1], O[3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// S4 + S5 fused (BASELINE config 4: echo synthesis ahead of the chain, frames with <= RSP_SYNTH_GATHER_T targets): the
// samples a thread would load for its MMA B fragments are generated in registers instead -- one Philox call + Box-Muller
// per (re, im) x 2 pair and the delayed pulses of the targets gathered from tx_pulse, exactly synth_gather_kernel's
// arithmetic (synth_pair) with the same counters, so the beams equal those of the two-kernel path bit for bit -- and
// the raw cube is never written or read: 2 x 67 MB of HBM traffic per frame disappear and the DBF's load latency with
// them.  The target phasors of the CTA's pulse (n_targets x C) are computed once per CTA into shared memory.
// MMA part, fragment layout and stores as in dbf_mma2_kernel.
// ------------------------------------------------------------------------------------------
template <int MT, int KS, int NQ>
__global__ void __launch_bounds__(RSP_DBF_MMA_THREADS) dbf_synth_kernel(const __grid_constant__ SynthArgs k, float2* __restrict__ beam,
                                                                        const float4* __restrict__ Wa /* [KS][MT][2][32] */,
                                                                        int NB, int ldb, int* __restrict__ det_count, const DiscardArgs dead) {
    __shared__ float2 s_ph[RSP_SYNTH_FUSED_T * 4 * KS];       // [target][channel], channel stride 4 KS >= C
    __shared__ int s_delay[RSP_SYNTH_FUSED_T];
    const int tid = threadIdx.x;
    const int C = k.C, N = k.N;
    l2_discard(dead);
    if (det_count && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *det_count = 0;   // first kernel of the frame
    const int p = blockIdx.y;
    for (int i = tid; i < k.n_targets * C; i += RSP_DBF_MMA_THREADS) {
        const int tt = i / C, c = i - tt * C;
        const SynthTarget tg = k.tg[tt];
        const float2 u = unit_phasor(tg.dop_fix * (unsigned long long)p + tg.steer_fix * (unsigned long long)c);
        s_ph[tt * (4 * KS) + c] = make_float2(tg.amp * u.x, tg.amp * u.y);
        if (c == 0) s_delay[tt] = tg.delay;
    }
    __syncthreads();
    const int lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int n_base = (blockIdx.x * (RSP_DBF_MMA_THREADS / 32) + w) * (16 * NQ);
    if (n_base >= N) return;
    const int sg = (g & 1) ? g + 7 : g;                                  // even: a float4 is one (n0, n0 + 1) pair of a line
    unsigned act[2];                                                     // targets with a non-zero stretch inside the warp's samples
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int tt = lane + 32 * h;
        bool a = false;
        if (tt < k.n_targets) {
            const int d = s_delay[tt];
#pragma unroll
            for (int sgm = 0; sgm < 3; ++sgm)
                a = a || (k.seg_hi[sgm] > k.seg_lo[sgm] && d + k.seg_lo[sgm] < n_base + 16 * NQ && d + k.seg_hi[sgm] > n_base);
        }
        act[h] = __ballot_sync(0xFFFFFFFFu, a);
    }
    float4 x[KS][NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int n0 = n_base + 16 * q + sg;
        float4 z[KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) z[s] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 < N) synth_pair_channels<KS>(k, s_ph + t, 4 * KS, s_delay, n0, (size_t)p * C + t, t, C, act[0], act[1], z);
#pragma unroll
        for (int s = 0; s < KS; ++s) x[s][q] = z[s];
    }
    float acc[MT][2 * NQ][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 2 * NQ; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][j][i] = 0.f;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        uint32_t bh[2 * NQ][2], bl[2 * NQ][2];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float v[4] = {x[s][q].x, x[s][q].y, x[s][q].z, x[s][q].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t hi = __float_as_uint(v[i]) & 0xFFFFE000u;
                bh[2 * q + (i >> 1)][i & 1] = hi;
                bl[2 * q + (i >> 1)][i & 1] = __float_as_uint(v[i] - __uint_as_float(hi)) & 0xFFFFE000u;
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const float4 ah = __ldg(Wa + ((s * MT + mt) * 2 + 0) * 32 + lane);
            const float4 al = __ldg(Wa + ((s * MT + mt) * 2 + 1) * 32 + lane);
#pragma unroll
            for (int j = 0; j < 2 * NQ; ++j) {
                mma_tf32_wa(acc[mt][j], al, bh[j][0], bh[j][1]);
                mma_tf32_wa(acc[mt][j], ah, bl[j][0], bl[j][1]);
                mma_tf32_wa(acc[mt][j], ah, bh[j][0], bh[j][1]);
            }
        }
    }
    float2* const brow = beam + (size_t)p * NB * ldb + 2 * t;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int b = 8 * mt + g;
        if (b < NB) {
            float2* row = brow + (size_t)b * ldb + n_base;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float(&E)[4] = acc[mt][2 * q];
                const float(&O)[4] = acc[mt][2 * q + 1];
                const int n = n_base + 16 * q + 2 * t;
                if (n < N) *reinterpret_cast<float4*>(row + 16 * q) = make_float4(E[0], E[2], O[0], O[2]);
                if (n + 8 < N) *reinterpret_cast<float4*>(row + 16 * q + 8) = make_float4(E[1], E[3], O[1], O[3]);
            }
        }
    }
}

// mbarrier / bulk-copy helpers of the TMA-fed kernels (rsp_fused.cuh, rsp_dbf_tc.cuh)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------------
// S6: pulse compression.  One CTA = Cfg::NG overlap-save blocks (one per group of Cfg::T threads);
// work item = (line, block).  The medium-segment launch also computes the narrow-pulse FIR gates of
// its lines (fun_process_single_frame.m:111-112,123).
// ------------------------------------------------------------------------------------------
struct PcSegArgs {
    const float2* tw1;
    const float2* tw2;
    const float2* Hmid;
    int seg_start0, in_lo, in_hi, taps, gate0, g_end, valid;
    int nblk, n_items, n_ctas;
};

struct PcKernelArgs {
    const float2* beam;
    float2* pc;
    int N, ldb, ldg;
    PcSegArgs seg[2];            // role 0 = CTAs [0, seg[0].n_ctas), role 1 = the rest
    // narrow FIR, computed by the role-1 (medium) groups for block 0 of their line when do_narrow
    int do_narrow;
    const float* fir;
    int nfir, fir_delay, narrow_start0, narrow_gates;
    int group_bar;               // 1: the groups of a CTA synchronise separately (named barriers)
};

// Barrier over one group of T threads (whole warps) that share an overlap-save block: named barrier 1 + grp.
// The groups of a CTA are independent after the shared tables are loaded, so they need not wait for each other.
__device__ __forceinline__ void pc_group_sync(bool per_group, int grp, int nthreads) {
    if (per_group) asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthreads) : "memory");
    else __syncthreads();
}

template <class Cfg>
__device__ __forceinline__ void pc_role(const PcKernelArgs& k, const PcSegArgs& sg, int cta, float2* smem, bool narrow) {
    constexpr int NTW2 = (Cfg::R2 - 1) * Cfg::SPAN2;
    float2* stw2 = smem + Cfg::NG * Cfg::SMEM_ELEMS;
    float* sfir = reinterpret_cast<float*>(stw2 + NTW2);
    for (int i = threadIdx.x; i < NTW2; i += RSP_PC_THREADS) stw2[i] = sg.tw2[i];
    if (narrow)
        for (int i = threadIdx.x; i < k.nfir; i += RSP_PC_THREADS) sfir[i] = k.fir[i];
    const int grp = threadIdx.x / Cfg::T, t = threadIdx.x - grp * Cfg::T;
    const int item = cta * Cfg::NG + grp;
    const bool active = item < sg.n_items;
    const int line = active ? item / sg.nblk : 0, blk = active ? item - line * sg.nblk : 0;
    float2* s = smem + grp * Cfg::SMEM_ELEMS;
    PcBlockArgs a;
    a.line = k.beam + (size_t)line * k.ldb;
    a.out_line = k.pc + (size_t)line * k.ldg;
    a.tw1 = sg.tw1;
    a.tw2 = stw2;
    a.Hmid = sg.Hmid;
    a.in_lo = sg.in_lo;
    a.in_hi = sg.in_hi;
    a.seg_start0 = sg.seg_start0;
    a.taps = sg.taps;
    a.g0 = sg.gate0 + blk * sg.valid;
    a.g_end = sg.g_end;
    const bool pg = k.group_bar != 0 && Cfg::NG > 1;
    if (active) pc_phase_load_pass1<Cfg>(a, s, t);
    __syncthreads();                                 // also publishes stw2 / sfir, loaded by the whole CTA
    if (active) pc_phase_pass2<Cfg>(a, s, t);
    pc_group_sync(pg, grp, Cfg::T);
    if (active) pc_phase_mid<Cfg>(a, s, t);
    pc_group_sync(pg, grp, Cfg::T);
    if (active) pc_phase_ipass2<Cfg>(a, s, t);
    pc_group_sync(pg, grp, Cfg::T);
    if (active) pc_phase_ipass1_store<Cfg>(a, s, t);
    if (narrow) {                                   // uniform over the CTA
        const int need = k.narrow_gates + k.fir_delay;
        const bool fast = need <= k.N - k.narrow_start0 && need <= Cfg::SMEM_ELEMS;
        pc_group_sync(pg, grp, Cfg::T);
        if (active && blk == 0) {
            if (fast) {
                for (int i = t; i < need; i += Cfg::T) s[i] = a.line[k.narrow_start0 + i];
            }
        }
        pc_group_sync(pg, grp, Cfg::T);
        if (active && blk == 0) {
            for (int g = t; g < k.narrow_gates; g += Cfg::T)
                a.out_line[g] = fast ? pc_narrow_gate_smem(s, sfir, k.nfir, k.fir_delay, g)
                                     : pc_narrow_gate(a.line, k.N, k.narrow_start0, sfir, k.nfir, k.fir_delay, g);
        }
    }
}

// CfgA = block plan of role 0 (the long segment), CfgB = role 1 (the medium segment + narrow FIR).
// The short role-1 CTAs have the highest block indices, so they fill the tail of the long ones.
// Register budget of the pulse-compression kernel: RSP_PC_NREG caps the registers per thread (the kernel compiles without
// spills down to 56), which sets how many 256-thread CTAs share an SM: 64 registers -> 4 CTAs = 32 warps per SM, measured
// 20.1 -> 18.9 us per CPI against the 80 registers / 3 CTAs that __launch_bounds__(256, 3) gave (profiles/r2_pc_regs.txt).
#ifndef RSP_PC_NREG
#define RSP_PC_NREG 64
#endif
#if RSP_PC_NREG > 0
#define RSP_PC_BOUNDS __maxnreg__(RSP_PC_NREG)
#else
#define RSP_PC_BOUNDS __launch_bounds__(RSP_PC_THREADS, 3)
#endif
template <class CfgA, class CfgB>
__global__ void RSP_PC_BOUNDS pc_fft_kernel(const PcKernelArgs k) {
    extern __shared__ float2 pc_smem[];
    if ((int)blockIdx.x < k.seg[0].n_ctas) pc_role<CfgA>(k, k.seg[0], blockIdx.x, pc_smem, false);
    else pc_role<CfgB>(k, k.seg[1], blockIdx.x - k.seg[0].n_ctas, pc_smem, k.do_narrow != 0);
}

__global__ void __launch_bounds__(256) pc_narrow_kernel(const float2* __restrict__ beam, float2* __restrict__ pc,
                                                        const float* __restrict__ fir, int nfir, int fir_delay, int N,
                                                        int ldb, int ldg, int seg_start0, int ngates) {
    __shared__ float sfir[256];
    for (int i = threadIdx.x; i < nfir; i += 256) sfir[i] = fir[i];
    __syncthreads();
    const size_t line = blockIdx.x;
    const float2* y = beam + line * ldb;
    for (int g = threadIdx.x; g < ngates; g += 256)
        pc[line * ldg + g] = pc_narrow_gate(y, N, seg_start0, sfir, nfir, fir_delay, g);
}

// ------------------------------------------------------------------------------------------
// S7: MTD.  Tile = 32 gates x P pulses of one beam.  The corner turn happens here: rows of the
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
    const int* perm;       // pow2 only: iperm[pos] = pulse stored at position pos
    int P;
    unsigned p_magic;      // 2^32 / P + 1: e / P == __umulhi(e, p_magic) for e < 64 P (the read-out of the generic-P kernel)
    int B, G, ldg;
    int g_lo, g_hi;        // gates [g_lo, g_hi) of this launch (the whole map: 0, G; range-blocked path: one chunk)
    DiscardArgs dead;      // buffer whose last reader has finished (the beam cube), or {nullptr, 0}
};

template <class Cfg>
__global__ void __launch_bounds__(RSP_MTD_THREADS) mtd_kernel(const MtdArgs k) {
    constexpr int P = Cfg::P, TG = RSP_MTD_TG;
    extern __shared__ float2 mtd_smem[];
    l2_discard(k.dead);
    float2* tile = mtd_smem;                       // [P][TG + 1]
    float2* stw = mtd_smem + P * (TG + 1);         // [Cfg::TW_COUNT]
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < Cfg::TW_COUNT; i += RSP_MTD_THREADS) stw[i] = k.tw[i];
    const int g0 = k.g_lo + blockIdx.x * TG, b = blockIdx.y;
    // innermost pass straight from global memory
    mtd_first_pass_t<Cfg>(tile, k.pc + (size_t)b * k.ldg + g0, (size_t)k.B * k.ldg, k.win, g0 + lane < k.g_hi, tid);
    __syncthreads();
    if (MtdInner<Cfg>::PASS < 1 && Cfg::R1 > 1) { mtd_passes_phase<Cfg>(tile, stw, tid, 1); __syncthreads(); }
    if (MtdInner<Cfg>::PASS < 2) { mtd_passes_phase<Cfg>(tile, stw, tid, 2); __syncthreads(); }
    // transposed read-out: consecutive threads take consecutive Doppler rows of one gate; element e = gl * P + row of the
    // tile is element e of the tile's contiguous block of rdm[b][g0 ..][.] (one base pointer, immediate offsets)
    const size_t o0 = ((size_t)b * k.G + g0) * P;
    float2* const rdm = k.rdm + o0;
    float* const amp = k.amp + o0;
    const int e_end = min(TG, k.g_hi - g0) * P;
#pragma unroll 4
    for (int e = tid; e < e_end; e += RSP_MTD_THREADS) {
        const int gl = e / P, row = e - gl * P;          // P is a compile-time power of two
        const float2 v = tile[row * (TG + 1) + gl];
        __stcs(rdm + e, v);
        amp[e] = sqrtf(fmaf(v.x, v.x, v.y * v.y));
    }
}

// Arguments of the P = 64 specialisation: the window sits in the kernel arguments (constant bank).
struct MtdRegArgs {
    MtdArgs m;
    float win[64];          // kaiser(P) * (-1)^p  (fftshift folded in), fun_process_single_frame.m:134-135
};

// S7, P = 64 specialisation of the tiled kernel (8 x 8 Cooley-Tukey, every index a compile-time constant).
//   phase 1, thread (q = warp, gl = lane): pulses p = q + 8m of gate gl -> window -> DFT-8 over m -> . W64^(q k1) -> S1[k1][q][gl]
//   phase 2, thread (k1 = warp, gl):       S1[k1][.][gl] -> DFT-8 over q -> X[k1 + 8 k2] -> S2[gl][k1 + 8 k2]   (corner turn)
//   phase 3: the tile is one contiguous 16 KB block of rdm[b][g][v]; consecutive threads store consecutive bins.
// Same data flow as mtd_kernel<MtdCfg<64, 8, 8, 1>> but without run-time butterfly addressing: phase 1 is instantiated
// once per warp index (switch), so the window comes from the constant bank (kernel argument) and the 49 twiddles are
// immediates.  5.3 M instead of 7.8 M warp instructions per CPI, 32 registers.  S1 and S2 share one 16.6 KB tile (an
// extra barrier between the phase-2 loads and stores): with separate tiles (33 KB) the kernel was slower than the
// generic one -- it lives on residency (8 CTAs per SM), not on its instruction count.  Lanes run along gates in phases
// 1-2 and along Doppler bins in phase 3; the S2 pitch of 65 keeps the transposed 64-bit stores conflict free (lane
// stride 130 words = 2 banks).
__host__ __device__ constexpr float tw64_re(int i) {   // Re exp(-2 pi i q k1 / 64) at i = 8 q + k1
    constexpr float t[64] = {   // exp(-2 pi i q k1 / 64) at [8 q + k1]
    1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f,
    1.0f, 0.99518472f, 0.980785251f, 0.956940353f, 0.923879504f, 0.881921291f, 0.831469595f, 0.773010433f,
    1.0f, 0.980785251f, 0.923879504f, 0.831469595f, 0.707106769f, 0.555570245f, 0.382683426f, 0.195090324f,
    1.0f, 0.956940353f, 0.831469595f, 0.634393275f, 0.382683426f, 0.0980171412f, -0.195090324f, -0.471396744f,
    1.0f, 0.923879504f, 0.707106769f, 0.382683426f, 0.0f, -0.382683426f, -0.707106769f, -0.923879504f,
    1.0f, 0.881921291f, 0.555570245f, 0.0980171412f, -0.382683426f, -0.773010433f, -0.980785251f, -0.956940353f,
    1.0f, 0.831469595f, 0.382683426f, -0.195090324f, -0.707106769f, -0.980785251f, -0.923879504f, -0.555570245f,
    1.0f, 0.773010433f, 0.195090324f, -0.471396744f, -0.923879504f, -0.956940353f, -0.555570245f, 0.0980171412f};
    return t[i];
}
__host__ __device__ constexpr float tw64_im(int i) {
    constexpr float t[64] = {
    0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f,
    0.0f, -0.0980171412f, -0.195090324f, -0.290284663f, -0.382683426f, -0.471396744f, -0.555570245f, -0.634393275f,
    0.0f, -0.195090324f, -0.382683426f, -0.555570245f, -0.707106769f, -0.831469595f, -0.923879504f, -0.980785251f,
    0.0f, -0.290284663f, -0.555570245f, -0.773010433f, -0.923879504f, -0.99518472f, -0.980785251f, -0.881921291f,
    0.0f, -0.382683426f, -0.707106769f, -0.923879504f, -1.0f, -0.923879504f, -0.707106769f, -0.382683426f,
    0.0f, -0.471396744f, -0.831469595f, -0.99518472f, -0.923879504f, -0.634393275f, -0.195090324f, 0.290284663f,
    0.0f, -0.555570245f, -0.923879504f, -0.980785251f, -0.707106769f, -0.195090324f, 0.382683426f, 0.831469595f,
    0.0f, -0.634393275f, -0.980785251f, -0.881921291f, -0.382683426f, 0.290284663f, 0.831469595f, 0.99518472f};
    return t[i];
}

template <int Q>
__device__ __forceinline__ void mtd64_phase1(const MtdRegArgs& k, const float2* src, unsigned pstride, float2* S1, int lane) {
    cf v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = src[(unsigned)(8 * m) * pstride];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = cscale(v[m], k.win[Q + 8 * m]);
    SmallDft<8, -1>::run(v);
    S1[(0 * 8 + Q) * 32 + lane] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < 8; ++k1)
        S1[(k1 * 8 + Q) * 32 + lane] = Q == 0 ? v[k1] : mul_tw<-1>(v[k1], tw64_re(8 * Q + k1), tw64_im(8 * Q + k1));
}

#define RSP_MTD64_PITCH 65
template <bool APPROX_SQRT>
__global__ void __launch_bounds__(256, 8) mtd64_kernel(const __grid_constant__ MtdRegArgs k) {
    __shared__ float2 S2[32 * RSP_MTD64_PITCH];     // S1 (phase 1 -> 2) and S2 (phase 2 -> 3) share the storage
    float2* const S1 = S2;
    l2_discard(k.m.dead);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y, G = k.m.G, g0 = k.m.g_lo + blockIdx.x * 32, g_hi = k.m.g_hi;
    {   // phase 1: the warp index picks the compile-time instance, so window and twiddles are immediates
        const int g = g0 + lane;
        const unsigned pstride = (unsigned)k.m.B * (unsigned)k.m.ldg;
        const float2* src = k.m.pc + (size_t)b * k.m.ldg + (g < g_hi ? g : g_hi - 1) + (size_t)w * pstride;
        switch (w) {
            case 0: mtd64_phase1<0>(k, src, pstride, S1, lane); break;
            case 1: mtd64_phase1<1>(k, src, pstride, S1, lane); break;
            case 2: mtd64_phase1<2>(k, src, pstride, S1, lane); break;
            case 3: mtd64_phase1<3>(k, src, pstride, S1, lane); break;
            case 4: mtd64_phase1<4>(k, src, pstride, S1, lane); break;
            case 5: mtd64_phase1<5>(k, src, pstride, S1, lane); break;
            case 6: mtd64_phase1<6>(k, src, pstride, S1, lane); break;
            default: mtd64_phase1<7>(k, src, pstride, S1, lane); break;
        }
    }
    __syncthreads();
    {   // phase 2
        cf v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = S1[(w * 8 + q) * 32 + lane];
        SmallDft<8, -1>::run(v);
        __syncthreads();                                 // every warp has read its S1 rows
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) S2[lane * RSP_MTD64_PITCH + w + 8 * k2] = v[k2];
    }
    __syncthreads();
    {   // phase 3: element e = gl * 64 + v of the tile, e = tid + 256 i
        const int v = threadIdx.x & 63, glb = threadIdx.x >> 6;
        const size_t o = ((size_t)b * G + g0) * 64 + threadIdx.x;
        float2* rdm = k.m.rdm + o;
        float* amp = k.m.amp + o;
        const int rows = g_hi - g0;              // gates of this tile that exist (>= 1)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int gl = glb + 4 * i;
            if (gl < rows) {
                const float2 x = S2[gl * RSP_MTD64_PITCH + v];
                __stcs(rdm + 256 * i, x);
                const float sq = fmaf(x.x, x.x, x.y * x.y);
                float a;
                if (APPROX_SQRT) asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(sq));
                else a = sqrtf(sq);
                amp[256 * i] = a;
            }
        }
    }
}

// KT = 1: one output bin per work item (mtd_dft_item); KT > 1: KT bins per item share the input loads
// (mtd_dft_item_kt), which moves the kernel from the shared-memory pipe to the FMA pipe.
template <int TG, int R, int KT>
__global__ void __launch_bounds__(RSP_MTD_THREADS) mtd_dft_kernel(const __grid_constant__ MtdArgs k) {
    extern __shared__ float2 mtd_smem[];
    l2_discard(k.dead);
    const int P = k.P, Q = P / R;
    constexpr int KP = KT >= 100 ? KT - 100 : 1;
    const bool inplace = KT >= 100 && mtd_dft_sym_inplace(Q, KP, TG, RSP_MTD_THREADS);
    float2* xin = mtd_smem;
    float2* xout = inplace ? xin : xin + (size_t)P * (TG + 1);
    float2* stw = xout + (size_t)P * (TG + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < P; i += RSP_MTD_THREADS) stw[i] = k.tw[i];
    const int g0 = k.g_lo + blockIdx.x * TG, b = blockIdx.y;
    const float2* const src = k.pc + (size_t)b * k.ldg + g0;                 // element (p, gl) of the tile: src[p * B * ldg + gl]
    const unsigned pstride = (unsigned)k.B * (unsigned)k.ldg;
    const int gl_end = k.g_hi - g0;
#pragma unroll 4
    for (int e = tid; e < P * TG; e += RSP_MTD_THREADS) {
        const int p = e / TG, gl = e - p * TG;
        float2 x = make_float2(0.f, 0.f);
        if (gl < gl_end) x = src[(size_t)((unsigned)p * pstride + (unsigned)gl)];
        xin[p * (TG + 1) + gl] = cscale(x, k.win[p]);
    }
    __syncthreads();
    if (KT >= 100) {                  // odd Q: even / odd folding, KT - 100 bin pairs per work item (mtd_dft_sym_*)
        mtd_dft_fold_phase(xin, P, R, TG, tid, RSP_MTD_THREADS);
        __syncthreads();
        const int n_items = mtd_dft_sym_groups(Q, KP) * TG;
        cf A[KP][R], Bm[KP][R];
        int kk[KP];
        if (inplace) {
            const int kg = tid / TG, gl = tid - kg * TG;
            if (tid < n_items) mtd_dft_sym_compute<R, KP, TG + 1>(xin, stw, P, kg * KP, gl, A, Bm, kk);
            __syncthreads();          // every item has read the tile
            if (tid < n_items) mtd_dft_sym_store<R, KP, TG + 1>(xout, stw, P, kg * KP, gl, A, Bm, kk);
        } else {
            for (int e = tid; e < n_items; e += RSP_MTD_THREADS) {
                const int kg = e / TG, gl = e - kg * TG;
                mtd_dft_sym_compute<R, KP, TG + 1>(xin, stw, P, kg * KP, gl, A, Bm, kk);
                mtd_dft_sym_store<R, KP, TG + 1>(xout, stw, P, kg * KP, gl, A, Bm, kk);
            }
        }
    } else if (KT == 1) {
        for (int e = tid; e < Q * TG; e += RSP_MTD_THREADS) {
            const int kk = e / TG, gl = e - kk * TG;
            mtd_dft_item<R>(xin, xout, stw, P, TG, kk, gl);
        }
    } else {
        const int groups = (Q + KT - 1) / KT;
        for (int e = tid; e < groups * TG; e += RSP_MTD_THREADS) {
            const int kg = e / TG, gl = e - kg * TG;
            mtd_dft_item_kt<R, (KT > 1 && KT < 100 ? KT : 2)>(xin, xout, stw, P, TG, kg * KT, gl);
        }
    }
    __syncthreads();
    // read-out, Doppler-contiguous: element e = gl * P + row of the tile is element e of the tile's block of rdm[b][g0..][.]
    // (one base pointer, no 64-bit index arithmetic per element; the division by the run-time P is a multiply-high)
    static_assert(TG <= 64, "p_magic is exact for e < 64 P");
    const size_t o0 = ((size_t)b * k.G + g0) * P;
    float2* const rdm = k.rdm + o0;
    float* const amp = k.amp + o0;
    const int e_end = min(TG, k.g_hi - g0) * P;
#pragma unroll 4
    for (int e = tid; e < e_end; e += RSP_MTD_THREADS) {
        const int gl = (int)__umulhi((unsigned)e, k.p_magic), row = e - gl * P;
        const float2 v = xout[row * (TG + 1) + gl];
        __stcs(rdm + e, v);
        amp[e] = sqrtf(fmaf(v.x, v.x, v.y * v.y));
    }
}

// zero-Doppler notch of the stage-2 path: rows [lo, hi] of every (beam, gate) line are cleared
__global__ void __launch_bounds__(256) doppler_notch_kernel(float2* __restrict__ rdm, size_t n_lines, int P, int lo, int hi) {
    const int w = hi - lo + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines * w; i += (size_t)gridDim.x * blockDim.x) {
        const size_t line = i / w;
        rdm[line * P + lo + (int)(i - line * w)] = make_float2(0.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------
// S8: GOCA-CFAR on S = |rdm_b| + |rdm_{b+1}| (lean kernel: window sums in shared memory, one
// compare per cell, compaction with one atomic per detection -- detections are rare).
// S9: refine_kernel, one thread per detection: spline peak search + monopulse angle.
// ------------------------------------------------------------------------------------------
// What cfar_kernel records per detection: everything S9 needs, so that the refinement does not depend on
// the lane's amplitude map (which the lane's next CPI overwrites) and can run once per batch.
struct RawDet {
    int v, r, pair;              // 0-based cell
    float power;                 // S(v, r)
    float yr[5];                 // S(v, r-2 .. r+2)   (range neighbours, fsf:250-256)
    float yv[5];                 // S(v-2 .. v+2, r)   (Doppler neighbours, fsf:265-271)
    float a_re, a_im, b_re, b_im;// amplitude mode: a_re = |rdm_A|, b_re = |rdm_B|; complex mode: rdm_A, rdm_B
};

struct CfarArgs {
    const float* amp;            // [B][G][P]
    const float2* rdm;           // [B][G][P]
    CfarParams c;
    int* count;                  // detection counter of this CPI slot (zeroed by dbf_kernel)
    RawDet* raw;                 // raw records of this CPI slot
    int cap;
    int complex_mode;
    int cut_lo, cut_hi;          // cells under test: gates [cut_lo, cut_hi) (the whole map: mR, G - mR; range-blocked path: a chunk)
    DiscardArgs dead;            // the pc cube (its last reader, mtd_kernel, has finished)
};

// S is the shared-memory sum-map tile; row0 points at S(gl = CUT row, v = 0), ld = its row pitch
__device__ __forceinline__ void cfar_emit(const CfarArgs& k, const float* row0, int ld, int v, int g, int pair, float power) {
    const int slot = atomicAdd(k.count, 1);
    if (slot >= k.cap) return;              // overflow is reported by the host from the count
    RawDet d;
    d.v = v; d.r = g; d.pair = pair; d.power = power;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        d.yr[i] = row0[(i - 2) * ld + v];
        d.yv[i] = row0[v - 2 + i];
    }
    const size_t o = ((size_t)pair * k.c.G + g) * k.c.P + v, nb = (size_t)k.c.G * k.c.P;
    if (k.complex_mode) {
        const float2 a = k.rdm[o], b = k.rdm[o + nb];
        d.a_re = a.x; d.a_im = a.y; d.b_re = b.x; d.b_im = b.y;
    } else {
        d.a_re = k.amp[o]; d.a_im = 0.f; d.b_re = k.amp[o + nb]; d.b_im = 0.f;
    }
    k.raw[slot] = d;
}

// generic scalar variant (any P, any window)
template <int TG>
__global__ void __launch_bounds__(RSP_CFAR_THREADS) cfar_kernel(const CfarArgs k) {
    extern __shared__ float cfar_smem[];
    l2_discard(k.dead);
    const int P = k.c.P, G = k.c.G;
    const int mR = k.c.guard_r + k.c.ref_r, mV = k.c.guard_v + k.c.ref_v;
    const int rows = TG + 2 * mR;
    float* S = cfar_smem;
    float* R5 = S + rows * P;
    float* D5 = R5 + cfar_r5_rows(k.c, TG) * P;
    const int pair = blockIdx.y;
    const int g_first = k.cut_lo + blockIdx.x * TG;
    const int tid = threadIdx.x;
    const float* A = k.amp + (size_t)pair * G * P;
    const float* Bm = A + (size_t)G * P;
    const size_t base = (size_t)(g_first - mR) * P;
    const int n_valid = min(rows, G - (g_first - mR)) * P;        // rows that exist in the map
    for (int e = tid; e < rows * P; e += RSP_CFAR_THREADS) S[e] = e < n_valid ? A[base + e] + Bm[base + e] : 0.f;
    __syncthreads();
    cfar_sums_phase(S, R5, D5, k.c, TG, tid, RSP_CFAR_THREADS);
    __syncthreads();
    const int nv = P - 2 * mV;                       // CUT columns [mV, P - mV)
    for (int e = tid; e < TG * nv; e += RSP_CFAR_THREADS) {
        const int gl = e / nv, v = mV + (e - gl * nv), g = g_first + gl;
        if (g >= k.cut_hi) break;
        float cut;
        if (cfar_decide(S, R5, D5, k.c, gl, v, &cut)) cfar_emit(k, S + (gl + mR) * P, P, v, g, pair, cut);
    }
}

// Vectorised variant for P % 4 == 0 (see cfar4_* in rsp_phases.cuh); RR/RV/GV = compile-time range
// reference, Doppler reference and Doppler guard lengths (RR = 0: run time).
template <int TG, int RR, int RV, int GV>
#ifndef RSP_CFAR_MINB
#define RSP_CFAR_MINB 3
#endif
__global__ void __launch_bounds__(RSP_CFAR_THREADS, RSP_CFAR_MINB) cfar4_kernel(const CfarArgs k) {
    extern __shared__ float cfar_smem[];
    l2_discard(k.dead);
    const Cfar4Geom g = cfar4_geom(k.c, TG);
    const int P = k.c.P, G = k.c.G;
    const int mR = k.c.guard_r + k.c.ref_r, mV = k.c.guard_v + k.c.ref_v;
    float* S = cfar_smem;                       // [rows][PP]
    float* R5 = S + g.rows * g.PP;              // [r5_rows][RP]
    const int pair = blockIdx.y;
    const int g_first = k.cut_lo + blockIdx.x * TG;
    const int tid = threadIdx.x;
    const float4* A4 = reinterpret_cast<const float4*>(k.amp + ((size_t)pair * G + (g_first - mR)) * P);
    const float4* B4 = A4 + (size_t)G * g.P4;
    float4* S4 = reinterpret_cast<float4*>(S);
    const int pp4 = g.PP / 4, h4 = RSP_CFAR_HALO / 4;
    const int rows_valid = min(g.rows, G - (g_first - mR));
    for (int idx = tid; idx < g.rows * 2 * h4; idx += RSP_CFAR_THREADS) {      // zero the halo columns
        const int row = idx / (2 * h4), c = idx - row * (2 * h4);
        S4[row * pp4 + (c < h4 ? c : g.P4 + c)] = make_float4(0.f, 0.f, 0.f, 0.f);       // columns [0, h4) and [h4 + P4, 2 h4 + P4)
    }
    {   // S = A + B: all loads of a batch are issued before the first add (the tile is one contiguous
        // block of each map, so the loads are full 128-byte lines)
        constexpr int U = 4;
        const int n_valid4 = rows_valid * g.P4, n_all4 = g.rows * g.P4;
        for (int base = tid; base < n_all4; base += RSP_CFAR_THREADS * U) {
            float4 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * RSP_CFAR_THREADS;
                a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                b[u] = a[u];
                if (idx < n_valid4) { a[u] = A4[idx]; b[u] = B4[idx]; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * RSP_CFAR_THREADS;
                if (idx < n_all4) {
                    int row, c4;
                    cfar4_split(g, idx, row, c4);
                    S4[row * pp4 + h4 + c4] = make_float4(a[u].x + b[u].x, a[u].y + b[u].y, a[u].z + b[u].z, a[u].w + b[u].w);
                }
            }
        }
    }
    __syncthreads();
    cfar4_r5_phase<RR>(S, R5, k.c, g, tid, RSP_CFAR_THREADS);
    __syncthreads();
    const int c_lo = mV / 4, nq = (P - mV - 1) / 4 - c_lo + 1;            // quads holding at least one CUT
    const int gl_end = min(TG, k.cut_hi - g_first);
    for (int idx = tid; idx < gl_end * nq; idx += RSP_CFAR_THREADS) {
        const int gl = idx / nq, c4 = c_lo + (idx - gl * nq);
        float cut[4];
        unsigned m = cfar4_decide_quad<RR, RV, GV>(S, R5, k.c, g, gl, c4, cut);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            cfar_emit(k, S + (gl + mR) * g.PP + RSP_CFAR_HALO, g.PP, 4 * c4 + j, g_first + gl, pair, cut[j]);
        }
    }
}

// Marching variant (cfar5_* in rsp_phases.cuh): the default for compile-time windows and P % 4 == 0, P <= 1024.
// Tile = gates [g_first - mR, g_first + TG + mR) of the pair's sum map; thread (quad q, row group) fills it with two
// 16-byte loads per row (the tile is one contiguous block of each amplitude map); then work item (chunk of CR gates,
// CUT quad) walks down its gates.
template <int TG, int RR, int RV, int GV>
__global__ void __launch_bounds__(RSP_CFAR_THREADS, 3) cfar5_kernel(const CfarArgs k) {
    extern __shared__ float cfar_smem[];
    l2_discard(k.dead);
    constexpr int CR = RSP_CFAR5_CR;
    const int P = k.c.P, G = k.c.G, P4 = P >> 2;
    const int mR = k.c.guard_r + RR, mV = GV + RV;
    const int pitch = cfar5_pitch(P, mV), pitch4 = pitch >> 2, rows = TG + 2 * mR;
    const int pair = blockIdx.y, g_first = k.cut_lo + blockIdx.x * TG, tid = threadIdx.x;
    float* S = cfar_smem;
    {
        const int rstep = RSP_CFAR_THREADS / P4, r0 = tid / P4, q = tid - r0 * P4;   // threads beyond rstep * P4 idle (P4 not a power of two)
        const float4* A4 = reinterpret_cast<const float4*>(k.amp + ((size_t)pair * G + (g_first - mR)) * P) + q;
        const float4* B4 = A4 + (size_t)G * P4;
        float4* Sq = reinterpret_cast<float4*>(S + RSP_CFAR5_HALO) + q;
        const int rows_valid = min(rows, G - (g_first - mR));
#ifndef RSP_CFAR5_U
#define RSP_CFAR5_U 5
#endif
        constexpr int U = RSP_CFAR5_U;     // rows in flight per thread (2 x 16-byte loads each)
        for (int r = r0 < rstep ? r0 : rows; r < rows; r += U * rstep) {
            float4 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = r + u * rstep;
                a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                b[u] = a[u];
                if (rr < rows_valid) { a[u] = A4[(size_t)rr * P4]; b[u] = B4[(size_t)rr * P4]; }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = r + u * rstep;
                if (rr < rows) Sq[rr * pitch4] = f4add(a[u], b[u]);
            }
        }
    }
    unsigned* queue = reinterpret_cast<unsigned*>(S + rows * pitch);      // [1 + TG * nq]: count, then (gl << 16 | c4 << 4 | mask)
    if (tid == 0) queue[0] = 0;
    __syncthreads();
    const int nq = cfar5_nq(P, mV), c_lo = mV / 4;
    const int gl_end = min(TG, k.cut_hi - g_first);
    const int n_items = ((gl_end + CR - 1) / CR) * nq;
    for (int it = tid; it < n_items; it += RSP_CFAR_THREADS) {
        const int ch = it / nq, c4 = c_lo + (it - ch * nq), gl0 = ch * CR;
        cfar5_march<RR, GV + RV, CR>(S, pitch, k.c, gl0, min(CR, gl_end - gl0), c4, [&](int s, unsigned m) {
            queue[1 + atomicAdd(queue, 1u)] = ((unsigned)(gl0 + s) << 16) | ((unsigned)c4 << 4) | m;
        });
    }
    __syncthreads();
    const int n_hit = (int)queue[0];                                      // quads above the range threshold: Doppler test, then emit
    const float kv = k.c.t_cfar / (float)RV;
    for (int e = tid; e < n_hit; e += RSP_CFAR_THREADS) {
        const unsigned w = queue[1 + e];
        const int gl = (int)(w >> 16), c4 = (int)((w >> 4) & 0xFFFu);
        const float* row0 = S + (gl + mR) * pitch + RSP_CFAR5_HALO;
        float4 cq;
        unsigned m = cfar5_doppler<RV, GV>(row0 + 4 * c4, kv, w & 15u, &cq);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const float cut = j == 0 ? cq.x : j == 1 ? cq.y : j == 2 ? cq.z : cq.w;
            cfar_emit(k, row0, pitch, 4 * c4 + j, g_first + gl, pair, cut);
        }
    }
}

// S9 (fun_process_single_frame.m:241-298) for every raw record of slots [first_slot, first_slot + n_slots):
// spline peak search on the fp32 sum-map neighbours the detector saw, monopulse ratio from the two
// beams at the integer cell.  One launch per batch; blockIdx.y = slot.
struct RefineArgs {
    const int* counts;           // [slots]
    const RawDet* raw;           // [slots][cap]
    rsp_detection* recs;         // [slots][cap]
    int cap, first_slot;
    const double* range_axis;
    const double* vel_axis;
    const double* beam_angles;
    const double* k_slopes;
    double delta_r, delta_v;
    int complex_mode;
};

__global__ void __launch_bounds__(128) refine_kernel(const RefineArgs k) {
    const int slot = k.first_slot + blockIdx.y;
    const int n = min(k.counts[slot], k.cap);
    const RawDet* raw = k.raw + (size_t)slot * k.cap;
    rsp_detection* out = k.recs + (size_t)slot * k.cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const RawDet r = raw[i];
        double yr[5], yv[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) { yr[j] = (double)r.yr[j]; yv[j] = (double)r.yv[j]; }
        const double r_off = rsp_spline5_peak(yr, 8) - 2.0;          // fsf:237 rInterpTimes = 8
        const double v_off = rsp_spline5_peak(yv, 4) - 2.0;          // vInterpTimes = 4
        rsp_detection d;
        d.v_idx = r.v + 1; d.r_idx = r.r + 1; d.pair_idx = r.pair + 1;
        d.power = r.power;
        d.range = k.range_axis[r.r] + r_off * k.delta_r;             // fsf:262
        d.velocity = k.vel_axis[r.v] + v_off * k.delta_v;            // fsf:278
        double ratio;
        const double eps = 2.220446049250313e-16;
        if (k.complex_mode) {                                        // mc:454-461
            const double nr = (double)r.a_re - (double)r.b_re, ni = (double)r.a_im - (double)r.b_im;
            const double dr = (double)r.a_re + (double)r.b_re + eps, di = (double)r.a_im + (double)r.b_im;
            ratio = (nr * dr + ni * di) / (dr * dr + di * di);
        } else {                                                     // fsf:282-285
            const double sa = (double)r.a_re, sb = (double)r.b_re;
            ratio = (sa - sb) / (sa + sb + eps);
        }
        d.angle = 0.5 * (k.beam_angles[r.pair] + k.beam_angles[r.pair + 1]) + k.k_slopes[r.pair] * ratio;   // fsf:286-290
        out[i] = d;
    }
}

}  // namespace rsp
