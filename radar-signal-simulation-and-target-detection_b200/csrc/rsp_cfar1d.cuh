// rsp_cfar1d.cuh -- the per-segment 1-D range CFAR of the real-data path: local_execute_cfar / executeCFAR_2D /
// Function_CFAR1D_sub, debug_simulated_data_processing_v2.m:419-511.
//
// Map layout: A[(b R + g) V + v], Doppler fastest == MATLAB mtd_amplitude_map(v, g) of beam b.  For every gate column y of a
// segment (1-based inside the segment, the three pulse segments are detected separately, :420-428) and every Doppler row
// outside the zero-velocity notch (:446-451):
//     left  window  y - (save + ref) .. y - save - 1,   right window  y + save + 1 .. y + save + ref          (:476-479)
//     a window that leaves the segment is replaced by the other one                                          (:481-491)
//     level = max (GOCA, method 0) or min (SOCA) of the two means,  threshold = level * T,  flag = A >= threshold   (:493-504)
// Rows inside the notch get flag 0 and threshold 0 (:456-460).  The means add left to right and divide by `ref`, like mean().
// One CTA = 32 Doppler bins x 64 gates; the amplitudes of the tile and its +-(save + ref) halo are staged once in shared
// memory (|z| is computed there when the input is the complex range-Doppler map of the stage-2 context).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rsp {

#define RSP_CFAR1D_TG 64
#define RSP_CFAR1D_MAXW 64            // save + ref

struct Cfar1dArgs {
    const float* amp;        // [B][R][V] amplitudes, or nullptr
    const float2* rdm;       // [B][R][V] complex map (amplitude = |z|), used when amp == nullptr
    unsigned char* flags;    // [B][R][V]
    float* thr;              // [B][R][V] or nullptr
    int V, R, B;
    int ref, save, method;   // method 0: greatest-of, 1: smallest-of
    float t_cfar;
    int seg_lo[3], seg_hi[3];   // 0-based gate ranges [lo, hi) of the three segments (empty: lo == hi)
    int notch_lo, notch_hi;     // 0-based Doppler rows [lo, hi] excluded from the detection
};

__global__ void __launch_bounds__(256) cfar1d_kernel(const __grid_constant__ Cfar1dArgs k) {
    __shared__ float tile[(RSP_CFAR1D_TG + 2 * RSP_CFAR1D_MAXW) * 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int v = blockIdx.x * 32 + lane, g0 = blockIdx.y * RSP_CFAR1D_TG, b = blockIdx.z;
    const int W = k.ref + k.save;
    const int rows = RSP_CFAR1D_TG + 2 * W;
    const size_t base = (size_t)b * k.R * k.V;
    for (int r = w; r < rows; r += 8) {
        const int g = g0 - W + r;
        float a = 0.f;
        if (g >= 0 && g < k.R && v < k.V) {
            const size_t o = base + (size_t)g * k.V + v;
            if (k.amp) a = k.amp[o];
            else { const float2 z = k.rdm[o]; a = sqrtf(fmaf(z.x, z.x, z.y * z.y)); }
        }
        tile[r * 32 + lane] = a;
    }
    __syncthreads();
    if (v >= k.V) return;
    const bool masked = v >= k.notch_lo && v <= k.notch_hi;
    for (int gl = w; gl < RSP_CFAR1D_TG; gl += 8) {
        const int g = g0 + gl;
        if (g >= k.R) break;
        int lo = 0, hi = 0;
#pragma unroll
        for (int s = 0; s < 3; ++s)
            if (g >= k.seg_lo[s] && g < k.seg_hi[s]) { lo = k.seg_lo[s]; hi = k.seg_hi[s]; }
        float thr = 0.f;
        unsigned char flag = 0;
        if (!masked && hi > lo) {
            const bool left_ok = g - W >= lo, right_ok = g + W < hi;      // refL1 >= 1, refR2 <= ColNum
            const float* c = tile + (gl + W) * 32 + lane;                  // c[d * 32] = A(v, g + d)
            float sl = 0.f, sr = 0.f;
            for (int i = 0; i < k.ref; ++i) {
                sl += c[(-W + i) * 32];                                   // gates g - W .. g - save - 1, left to right
                sr += c[(k.save + 1 + i) * 32];                           // gates g + save + 1 .. g + W
            }
            const float ml = (left_ok ? sl : sr) / (float)k.ref, mr = (right_ok ? sr : sl) / (float)k.ref;
            const float level = k.method == 0 ? fmaxf(ml, mr) : fminf(ml, mr);
            thr = level * k.t_cfar;
            flag = c[0] >= thr ? 1 : 0;
        }
        const size_t o = base + (size_t)g * k.V + v;
        k.flags[o] = flag;
        if (k.thr) k.thr[o] = thr;
    }
}

}  // namespace rsp
