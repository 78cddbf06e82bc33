// Micro-benchmark: per-SM-sub-partition cost of common instruction classes on sm_100a, alone and mixed
// with FFMA, to build the cost model used in DESIGN.md (which instructions share the FP32 datapath).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_opcost tools/ubench/opcost.cu
#include <cstdio>
#include <cuda_runtime.h>
#define OPS(X) X(0, IADD) X(1, IMAD) X(2, LOP) X(3, SHF) X(4, ISETPSEL) X(5, LEA) X(6, FMNMX) X(7, MUFU) X(8, LDS64) X(9, LDS128) X(10, STS64) X(11, FADD) X(12, PRMT) X(13, LDS32) X(14, STS128)
template <int OP> __device__ __forceinline__ void op(unsigned& a, unsigned b, float* sm, int i) {
    if (OP == 0) asm volatile("add.u32 %0,%0,%1;" : "+r"(a) : "r"(b));
    if (OP == 1) asm volatile("mad.lo.u32 %0,%0,%1,%1;" : "+r"(a) : "r"(b));
    if (OP == 2) asm volatile("xor.b32 %0,%0,%1;" : "+r"(a) : "r"(b));
    if (OP == 3) asm volatile("shf.l.wrap.b32 %0,%0,%1,7;" : "+r"(a) : "r"(b));
    if (OP == 4) asm volatile("{.reg .pred p; setp.lt.u32 p,%0,%1; selp.u32 %0,%1,%0,p;}" : "+r"(a) : "r"(b));
    if (OP == 5) asm volatile("{.reg .u32 t; shl.b32 t,%0,3; add.u32 %0,t,%1;}" : "+r"(a) : "r"(b));
    if (OP == 6) asm volatile("max.f32 %0,%0,%1;" : "+f"(*(float*)&a) : "f"(__uint_as_float(b)));
    if (OP == 7) asm volatile("sqrt.approx.f32 %0,%0;" : "+f"(*(float*)&a));
    if (OP == 8) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1},[%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((a & 1023u) << 3))); a ^= __float_as_uint(v.x) ^ __float_as_uint(v.y); }
    if (OP == 9) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3},[%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((a & 511u) << 4))); a ^= __float_as_uint(v.x) ^ __float_as_uint(v.w); }
    if (OP == 10) asm volatile("st.shared.v2.f32 [%0],{%1,%2};" ::"r"((unsigned)__cvta_generic_to_shared(sm) + ((threadIdx.x + i * 32u) & 1023u) * 8u), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)) : "memory");
    if (OP == 11) asm volatile("add.f32 %0,%0,%1;" : "+f"(*(float*)&a) : "f"(__uint_as_float(b)));
    if (OP == 12) asm volatile("prmt.b32 %0,%0,%1,0x1230;" : "+r"(a) : "r"(b));
    if (OP == 13) { float v; asm volatile("ld.shared.f32 %0,[%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm) + ((a & 2047u) << 2))); a ^= __float_as_uint(v); }
    if (OP == 14) asm volatile("st.shared.v4.f32 [%0],{%1,%2,%1,%2};" ::"r"((unsigned)__cvta_generic_to_shared(sm) + ((threadIdx.x + i * 32u) & 511u) * 16u), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)) : "memory");
}
// NF FFMA + NX ops per iteration
template <int OP, int NF, int NX> __global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = s * i;
    __syncthreads();
    float a[16]; unsigned n[8];
    for (int i = 0; i < 16; ++i) a[i] = s + i + threadIdx.x;
    for (int i = 0; i < 8; ++i) n[i] = i * 77 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f32 %0,%0,%1,%2;" : "+f"(a[i & 15]) : "f"(s), "f"(a[(i + 1) & 15]));
#pragma unroll
        for (int i = 0; i < NX; ++i) op<OP>(n[i & 7], n[(i + 1) & 7], sm, i);
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += a[i];
    for (int i = 0; i < 8; ++i) r += n[i];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
template <int OP, int NF, int NX> float run(float* out) {
    const int iters = 2048, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP, NF, NX><<<grid, 256>>>(out, 64, 1.0f);
    cudaEventRecord(e0);
    k<OP, NF, NX><<<grid, 256>>>(out, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e-3f * 1.965e9f / iters / 16.f;   // cycles per iteration per warp (16 warps per sub-partition)
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    printf("cycles per warp-iteration at 1965 MHz, 16 warps per sub-partition\n");
    printf("%-10s %10s %10s %14s %14s\n", "op", "16 X", "8 X", "16 FFMA + 8 X", "16 FFMA + 16 X");
    printf("%-10s %10.2f\n", "FFMA(16)", run<0, 16, 0>(out));
#define X(id, name) printf("%-10s %10.2f %10.2f %14.2f %14.2f\n", #name, run<id, 0, 16>(out), run<id, 0, 8>(out), run<id, 16, 8>(out), run<id, 16, 16>(out));
    OPS(X)
    cudaError_t e = cudaDeviceSynchronize(); printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
