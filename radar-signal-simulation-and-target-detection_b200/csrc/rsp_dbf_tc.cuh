// rsp_dbf_tc.cuh -- S5 (fun_process_single_frame.m:92-97) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with
// the accumulators AND the data operand in tensor memory, the weights in shared memory, the raw cube streamed in by TMA.
//
//   beam[p][b][n] = sum_c raw[p][c][n] * conj(W[b][c])      as the real GEMM   D[128 samples x 2B] = A[128 x 2C] . Bw[2C x 2B]
//       A[m][2c + {0,1}] = {Re, Im} x(c, n0 + m)                 (TMEM: lane = sample row m, column = k)
//       Bw[2b + 0][2c + {0,1}] = { wr, wi },  Bw[2b + 1][2c + {0,1}] = { -wi, wr },  W[b][c] = wr + i wi   (x * conj(W))
//   3xTF32 error compensation like the mma.sync kernels: D = Ah.Bh + Al.Bh + Ah.Bl with Ah = x truncated to tf32,
//   Al = x - Ah (exact), Bh / Bl split on the host; fp32 accumulation in TMEM.  Dropped terms are O(2^-21 |x w|).
//
// Persistent CTA (one per SM), six warps:
//   warp 0      TMA producer: one cp.async.bulk.tensor.2d per tile ([C channels] x [128 samples], 8 C x 128 bytes) into a
//               ring of NS stages, completion on mbarriers; also allocates / frees the TMEM columns
//   warp 1      MMA issuer: one lane issues the 3 x (2C / 8) tcgen05.mma of a tile and commits them to an mbarrier
//   warps 2-5   converters + epilogue: thread m owns sample row m.  Convert: raw stage -> Ah, Al written straight into
//               tensor memory with tcgen05.st (lane m, columns 2c, 2c + 1), so the operand never crosses shared memory a
//               second time (measured: as shared-memory operands in the no-swizzle K-major layout the twelve MMAs of a
//               tile cost 1700 cycles of operand fetch).  Epilogue: tcgen05.ld of the row's 2B accumulators, one 8-byte
//               store per beam (256 contiguous bytes per warp).  Operand and accumulator buffers are double: the
//               conversion of tile i + 1 and the epilogue of tile i - 1 overlap the MMAs of tile i.
// Registers: 192 threads x <= 64, everything else lives in shared / tensor memory, so the kernel leaves most of the register
// file to co-resident kernels of other lanes.  HBM-bound: 8 C bytes in + 8 B bytes out per sample.
#pragma once
#include <cuda.h>
#include "rsp_fused.cuh"

namespace rsp {

#define RSP_TC_THREADS 192
#define RSP_TC_TILE 128
#define RSP_TC_MAX_STAGES 6

struct DbfTcArgs {
    float2* beam;            // [P][B][ldb]
    const float* Bw;         // [2][Npad * Kpad] hi, lo weight operands in the canonical layout (make_dbf_tc_weights)
    int* det_count;
    int C, B, P, N, ldb;
    int Cpad;                // C rounded up to a multiple of 4 (K = 2 Cpad is a multiple of 8)
    int Npad;                // 2 B rounded up to 16 or 32 (MMA N)
    int stages;              // ring depth
    int tiles_per_pulse;     // ceil(N / 128)
    int chunk;               // > 0: CTA j owns the contiguous tiles [j chunk, (j + 1) chunk); 0: persistent, tile = j + i gridDim
    int tile_lo, tile_hi;    // tiles [tile_lo, tile_hi) of every pulse (the whole line: 0, tiles_per_pulse; range-blocked path: a chunk)
    int p_lo, p_hi;          // pulses [p_lo, p_hi)
    int p_wrap;              // RSP_PROBES builds: > 0 writes pulse p to beam row p % p_wrap (multi-CPI experiment)
    int dbg;                 // RSP_TC_DEBUG measurement aid (results wrong): 1 no MMAs, 2 no beam stores, 4 no conversion, 8 no proxy fence, 16 one MMA term
    DiscardArgs dead;
};

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // SM100 shared-memory matrix descriptor, no swizzle: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// mbarrier wait of the role warps.  RSP_TC_SLEEP_NS > 0 backs off between polls (__nanosleep): the polls are 40 % of this
// kernel's executed instructions, which cost nothing while it runs alone but take issue slots from co-resident kernels.
#ifndef RSP_TC_SLEEP_NS
#define RSP_TC_SLEEP_NS 0
#endif
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) {
#if RSP_TC_SLEEP_NS > 0
    uint32_t ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(RSP_TC_SLEEP_NS);
    }
#else
    mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                   "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int NCOL>    // 16 or 32 accumulator columns of this thread's row
__device__ __forceinline__ void tc_ld_row(uint32_t taddr, float (&v)[NCOL]) {
    uint32_t r[NCOL];
    if constexpr (NCOL == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
                     "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                       "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                       "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < NCOL; ++i) v[i] = __uint_as_float(r[i]);
}

// RSP_TC_NREG caps the registers per thread so that the kernel (192 threads) fits beside the pulse-compression CTAs of
// another lane: 3 x 256 x 72 + 192 x 48 registers fill the file exactly.
#ifndef RSP_TC_NREG
#define RSP_TC_NREG 0
#endif
#if RSP_TC_NREG > 0
#define RSP_TC_BOUNDS __maxnreg__(RSP_TC_NREG)
#else
#define RSP_TC_BOUNDS __launch_bounds__(RSP_TC_THREADS, 1)
#endif
#ifndef RSP_TC_CONV_UNROLL
#define RSP_TC_CONV_UNROLL 4          // 4-channel groups of the conversion kept in flight (1 = leanest in registers)
#endif
template <int NPAD, int CPAD>    // MMA N: 16 (B <= 8) or 32 (B <= 16); channels padded to 8 / 16 / 32 (K = 2 CPAD)
__global__ void RSP_TC_BOUNDS dbf_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                   const __grid_constant__ DbfTcArgs k) {
    extern __shared__ __align__(1024) unsigned char tsm[];
    const int C = k.C, NS = k.stages;
    constexpr int Cpad = CPAD;
    const uint32_t raw_stage = (uint32_t)C * 1024u;                  // [C][128 samples] complex64
    constexpr int KCOLS = 2 * CPAD;                                  // TMEM columns of one operand (hi or lo) buffer
    constexpr int TCOLS = 64 + 4 * KCOLS <= 128 ? 128 : 64 + 4 * KCOLS <= 256 ? 256 : 512;   // D0, D1 at columns 0 / NPAD, operands from 64
    const uint32_t b_bytes = (uint32_t)NPAD * (uint32_t)(2 * Cpad) * 4u;
    unsigned char* const s_raw = tsm;
    unsigned char* const s_bh = s_raw + (size_t)NS * raw_stage;
    unsigned char* const s_bl = s_bh + b_bytes;
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(s_bl + b_bytes);
    // barriers: [0, NS) raw full, [NS, 2 NS) raw empty, 2 NS + {0, 1}: A buffer ready, 2 NS + 2 + {0, 1}: MMAs of the buffer done
    // (operands consumed, accumulator complete), 2 NS + 4 + {0, 1}: accumulator buffer read out
    uint32_t* const s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * RSP_TC_MAX_STAGES + 6);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int B_A_READY = 2 * NS, B_MMA_DONE = 2 * NS + 2, B_D_FREE = 2 * NS + 4;

    if (k.det_count && blockIdx.x == 0 && tid == 0) *k.det_count = 0;          // first kernel of the CPI
    l2_discard(k.dead);                                                        // every thread of the grid takes its share of the lines
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(BAR(s), 1); mbar_init(BAR(NS + s), 4); }
        for (int b = 0; b < 2; ++b) { mbar_init(BAR(B_A_READY + b), 4); mbar_init(BAR(B_MMA_DONE + b), 1); mbar_init(BAR(B_D_FREE + b), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // weight operands (already in the canonical layout) -> shared memory
        const float4* src = reinterpret_cast<const float4*>(k.Bw);
        float4* dst = reinterpret_cast<float4*>(s_bh);
        for (int i = tid; i < (int)(2 * b_bytes / 16); i += RSP_TC_THREADS) dst[i] = src[i];
        fence_async_smem();
    }
    if (w == 0) {                                                              // 2 x NPAD accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int tpp = k.tile_hi - k.tile_lo;                                   // tiles per pulse of this launch
    const int n_tiles = (k.p_hi - k.p_lo) * tpp;
    const int first = k.chunk > 0 ? blockIdx.x * k.chunk : blockIdx.x, step = k.chunk > 0 ? 1 : gridDim.x;
    const int n_my = first >= n_tiles ? 0 : k.chunk > 0 ? min(k.chunk, n_tiles - first) : (n_tiles - first + step - 1) / step;

    if (w == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
#ifdef RSP_TC_EVICT_FIRST
            const uint64_t pol = l2_policy_evict_first();
#endif
            for (int i = 0; i < n_my; ++i) {
                const int s = i % NS, round = i / NS;
                if (round > 0) tc_wait(BAR(NS + s), (uint32_t)(round - 1) & 1u);
                const int tile = first + i * step;
                const int pr = tile / tpp, p = k.p_lo + pr, n0 = (k.tile_lo + tile - pr * tpp) * RSP_TC_TILE;
                mbar_expect_tx(BAR(s), raw_stage);
#ifdef RSP_TC_EVICT_FIRST
                tma_load_2d_hint(smem_u32(s_raw) + (uint32_t)s * raw_stage, &tmap, 2 * n0, p * C, BAR(s), pol);
#else
                tma_load_2d(smem_u32(s_raw) + (uint32_t)s * raw_stage, &tmap, 2 * n0, p * C, BAR(s));
#endif
            }
        }
    } else if (w == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            constexpr uint32_t lbo_b = (uint32_t)NPAD * 16u, sbo_b = 128u;   // core column / row-group strides of the weight operand
            constexpr int ksteps = KCOLS / 8;
            // descriptor of k-step ks = descriptor of step 0 + ks * (2 LBO >> 4) in the start-address field: the issue loop is
            // straight-line code with immediate offsets (a loop that rebuilt the descriptors cost ~55 cycles per MMA in the
            // one issuing thread, three times the tensor-core time of an M128 N32 K8 instruction)
            const uint64_t bdesc_h = tc_smem_desc(smem_u32(s_bh), lbo_b, sbo_b), bdesc_l = tc_smem_desc(smem_u32(s_bl), lbo_b, sbo_b);
            constexpr uint64_t kstep_b = (uint64_t)((2u * lbo_b) >> 4);
            const int first_term = (k.dbg & 1) ? 3 : (k.dbg & 16) ? 2 : 0;
            for (int i = 0; i < n_my; ++i) {
                const int buf = i & 1;
                tc_wait(BAR(B_A_READY + buf), (uint32_t)(i >> 1) & 1u);
                if (i >= 2) tc_wait(BAR(B_D_FREE + buf), (uint32_t)((i >> 1) - 1) & 1u);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(buf * NPAD);
                const uint32_t a_hi = tmem_base + 64u + (uint32_t)(2 * buf * KCOLS), a_lo = a_hi + (uint32_t)KCOLS;
                if (first_term == 0) {                                   // Al.Bh + Ah.Bl + Ah.Bh (small terms first)
#pragma unroll
                    for (int ks = 0; ks < ksteps; ++ks) tc_mma_tf32_ts(d, a_lo + 8u * ks, bdesc_h + kstep_b * ks, idesc, ks ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < ksteps; ++ks) tc_mma_tf32_ts(d, a_hi + 8u * ks, bdesc_l + kstep_b * ks, idesc, 1u);
#pragma unroll
                    for (int ks = 0; ks < ksteps; ++ks) tc_mma_tf32_ts(d, a_hi + 8u * ks, bdesc_h + kstep_b * ks, idesc, 1u);
                } else if (first_term == 2) {
#pragma unroll
                    for (int ks = 0; ks < ksteps; ++ks) tc_mma_tf32_ts(d, a_hi + 8u * ks, bdesc_h + kstep_b * ks, idesc, ks ? 1u : 0u);
                }
                tc_commit(BAR(B_MMA_DONE + buf));                        // implies tcgen05.fence::before_thread_sync
            }
        }
    } else {
        // ------------------------------------------------------------------ converters + epilogue (thread = sample row)
        const int q = w & 3, m = 32 * q + lane;                          // TMEM lanes 32 q .. 32 q + 31 belong to warp q (mod 4)
        auto epilogue = [&](int i) {
            const int tile = first + i * step;
            const int pr = tile / tpp, p = k.p_lo + pr, n = (k.tile_lo + tile - pr * tpp) * RSP_TC_TILE + m;
#ifdef RSP_PROBES
            float2* dst = k.beam + (size_t)(k.p_wrap > 0 ? p % k.p_wrap : p) * k.B * k.ldb + n;
#else
            float2* dst = k.beam + (size_t)p * k.B * k.ldb + n;
#endif
#pragma unroll
            for (int h = 0; h < NPAD / 16; ++h) {                        // 16 accumulator columns = 8 beams at a time
                float v[16];
                tc_ld_row<16>(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((i & 1) * NPAD + 16 * h), v);
                if (h == NPAD / 16 - 1) {                                // the buffer has been read out: the tensor core may reuse it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(B_D_FREE + (i & 1)));
                }
                if (n < k.N && !(k.dbg & 2)) {
#pragma unroll
                    for (int b = 0; b < 8; ++b)
                        if (8 * h + b < k.B) dst[(size_t)(8 * h + b) * k.ldb] = make_float2(v[2 * b], v[2 * b + 1]);
                }
            }
        };
        for (int i = 0; i < n_my; ++i) {
            const int s = i % NS, buf = i & 1;
            tc_wait(BAR(s), (uint32_t)(i / NS) & 1u);
            if (i >= 2) { tc_wait(BAR(B_MMA_DONE + buf), (uint32_t)((i >> 1) - 1) & 1u); tc_fence_after(); }   // operand buffer `buf` has been consumed
            const unsigned char* rs = s_raw + (size_t)s * raw_stage + (size_t)m * 8;
            const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + 64u + (uint32_t)(2 * buf * KCOLS);    // hi buffer; lo follows at + KCOLS
            if (!(k.dbg & 4)) {
                constexpr int kConvUnroll = RSP_TC_CONV_UNROLL;
#pragma unroll kConvUnroll
                for (int j = 0; j < CPAD / 4; ++j) {                      // 4 channels = 8 operand columns per store
                    float2 x[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) x[c] = 4 * j + c < C ? *reinterpret_cast<const float2*>(rs + (size_t)(4 * j + c) * 1024) : make_float2(0.f, 0.f);
                    uint32_t h[8], l[8];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        h[2 * c] = __float_as_uint(x[c].x) & 0xFFFFE000u;
                        h[2 * c + 1] = __float_as_uint(x[c].y) & 0xFFFFE000u;
                        l[2 * c] = __float_as_uint(x[c].x - __uint_as_float(h[2 * c]));
                        l[2 * c + 1] = __float_as_uint(x[c].y - __uint_as_float(h[2 * c + 1]));
                    }
                    tc_st8(ta + 8u * (uint32_t)j, h);
                    tc_st8(ta + (uint32_t)KCOLS + 8u * (uint32_t)j, l);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(BAR(B_A_READY + buf)); mbar_arrive(BAR(NS + s)); }
            if (i > 0) {                                                  // epilogue of the previous tile while the tensor core works on this one
                tc_wait(BAR(B_MMA_DONE + (buf ^ 1)), (uint32_t)((i - 1) >> 1) & 1u);
                tc_fence_after();
                epilogue(i - 1);
            }
        }
        if (n_my > 0) {
            tc_wait(BAR(B_MMA_DONE + ((n_my - 1) & 1)), (uint32_t)((n_my - 1) >> 1) & 1u);
            tc_fence_after();
            epilogue(n_my - 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (w == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
    }
}

}  // namespace rsp
