// Micro-benchmark: issue cost of the sm_100a packed fp32 instructions (FFMA2 / FADD2) against scalar
// FFMA / FADD, alone and mixed with integer and shared-memory instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_f32x2 tools/ubench/f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0,%1,%2,%3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0,%1,%2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) { unsigned r; asm volatile("add.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// MODE 0: 16 scalar FFMA / iter; 1: 8 FFMA2 / iter (same flops); 2: 16 FADD; 3: 8 FADD2;
// 4: 16 FFMA + 8 IADD; 5: 8 FFMA2 + 8 IADD; 6: 16 FFMA + 4 LDS.64; 7: 8 FFMA2 + 4 LDS.64 (+2 LDS.128 variant = 8)
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = s * i;
    __syncthreads();
    float a[16]; u64 p[8]; unsigned n[8];
    for (int i = 0; i < 16; ++i) a[i] = s + i + threadIdx.x;
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(s + i, s - i + threadIdx.x); p[i] = *(u64*)&v; n[i] = i + threadIdx.x; }
    float2 cc = make_float2(s, s * 0.5f); u64 c2 = *(u64*)&cc;
    const float2* sm2 = (const float2*)sm; const float4* sm4 = (const float4*)sm;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 4 || MODE == 6) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma1(a[i], s, a[(i + 1) & 15]);
        }
        if (MODE == 1 || MODE == 5 || MODE == 7 || MODE == 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], c2, p[(i + 1) & 7]);
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = add1(a[i], a[(i + 1) & 15]);
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = add2(p[i], p[(i + 1) & 7]);
        }
        if (MODE == 4 || MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) n[i] = iadd(n[i], n[(i + 1) & 7]);
        }
        if (MODE == 6 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { float2 v = sm2[(n[i] + it) & 1023]; a[i] += v.x; p[i] ^= __float_as_uint(v.y); }
        }
        if (MODE == 8) {
#pragma unroll
            for (int i = 0; i < 2; ++i) { float4 v = sm4[(n[i] + it) & 511]; a[i] += v.x + v.z; p[i] ^= __float_as_uint(v.y + v.w); }
        }
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
    for (int i = 0; i < 8; ++i) { float2 v = *(float2*)&p[i]; r += v.x + v.y + n[i]; }
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* out, double flops_per_iter) {
    const int iters = 4096, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 64, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double thr = (double)grid * 256 * iters;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  %6.2f ns/iter/warp-slot\n", name, ms, thr * flops_per_iter / ms * 1e-9, ms * 1e6 / iters);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("16 FFMA", out, 32); run<1>("8 FFMA2", out, 32); run<2>("16 FADD", out, 16); run<3>("8 FADD2", out, 16);
    run<4>("16 FFMA + 8 IADD", out, 32); run<5>("8 FFMA2 + 8 IADD", out, 32);
    run<6>("16 FFMA + 4 LDS.64", out, 32); run<7>("8 FFMA2 + 4 LDS.64", out, 32); run<8>("8 FFMA2 + 2 LDS.128", out, 32);
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
