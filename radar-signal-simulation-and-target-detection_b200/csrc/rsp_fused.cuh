// rsp_fused.cuh -- S5 + S6 in one launch: digital beamforming fused into the pulse-compression input
// (fun_process_single_frame.m:92-97 and :101-127), the `dbf_pc` kernel of SURVEY.md section 7.
//
// One thread-block cluster per pulse, one CTA per beam (cluster size = B <= 8).
//   DBF phase   CTA r of the cluster streams ITS slice of the pulse's range axis -- [C channels x 128 samples]
//               tiles -- from HBM with TMA tensor copies (cp.async.bulk.tensor, 2-stage ring, completion on an
//               mbarrier), forms ALL B beams of the slice on the tensor cores (the arithmetic of dbf_mma2_kernel:
//               3xTF32 mma.sync with the weights as the A operand, bit-identical results) and scatters beam g's
//               samples into the range line that CTA g keeps in ITS shared memory (st.shared::cluster over DSMEM).
//               The raw cube is read exactly once and the beam cube never exists in global memory or L2.
//   barrier     barrier.cluster arrive.release / wait.acquire: every CTA now holds the complete line of its beam.
//   PC phase    the overlap-save blocks of the line (rsp_phases.cuh, unchanged arithmetic): groups of warps run
//               DIF -> DIF -> [DIF . H . DIT] -> DIT -> DIT on one block each with the first pass reading the
//               line from shared memory instead of L2; blocks are packed into rounds of 8 warps (config 2: one
//               4096-point block, then 2048 + 1024 + 1024 together); the medium group adds the narrow-pulse FIR.
//               pc[p][b][g] leaves in the layout the Doppler kernel expects.
// Shared memory per CTA: line 8 N bytes + workspace (the TMA ring aliases it: the two phases never overlap inside
// a CTA) + tables: 106 KB at config 2, two CTAs per SM, so the memory-bound DBF phase of one cluster overlaps the
// issue-bound FFT phase of another on the same SM.
#pragma once
#include <cuda.h>
#include "rsp_kernels.cuh"

namespace rsp {

#define RSP_FUSED_THREADS 256
#define RSP_FUSED_WARPS (RSP_FUSED_THREADS / 32)
#define RSP_FUSED_TILE 128                 // range samples per TMA tile: one 16-sample MMA group per warp
#define RSP_FUSED_STAGES 2
#define RSP_FUSED_MAX_ROUNDS 4
#define RSP_FUSED_MAX_GROUPS 8

struct DbfPcGroup {          // one overlap-save block, run by a group of whole warps in one round
    int cfg;                 // 1 = PcCfg<1024>, 2 = PcCfg<2048>, 3 = PcCfg<4096>
    int seg;                 // index into DbfPcArgs::seg
    int blk;                 // block index inside the segment
    int warp0;               // first warp of the group
    int work_off;            // complex elements from the start of the workspace
    int narrow;              // the group also computes the narrow-pulse FIR gates of the line
};

struct DbfPcArgs {
    float2* pc;              // [P][B][ldg]
    float2* beam_out;        // optional [P][B][ldb] copy of the beam lines (rsp_get_beam on the single-CPI path), or nullptr
    const float4* Wa;        // weight fragments of dbf_mma2_kernel [KS][1][2][32]
    int* det_count;          // zeroed here: first kernel of the CPI
    int C, B, P, N, ldb, ldg;
    int tiles;               // ceil(N / RSP_FUSED_TILE)
    int tma_rank4;           // 1: one 4-D tensor copy per tile; 0: eight 2-D copies (one per 16-sample group)
    int line_bytes;          // shared-memory offset of the workspace (8 N rounded up to 128)
    int work_bytes;          // workspace / TMA ring
    PcSegArgs seg[4];
    int tw2_off[4];          // complex-element offsets of the segments' tw2 copies inside the table area
    int tw2_len[4];
    int n_rounds;
    signed char warp_group[RSP_FUSED_MAX_ROUNDS][RSP_FUSED_WARPS];     // group of each warp in each round, -1 = idle
    DbfPcGroup grp[RSP_FUSED_MAX_ROUNDS][RSP_FUSED_MAX_GROUPS];
    const float* fir;
    int nfir, fir_delay, narrow_start0, narrow_gates;
    DiscardArgs dead;
    long long* dbg;          // optional [grid][8] phase timestamps (globaltimer ns) + SM id, tools/fused_diag.py; nullptr in production
    int dbg_flags;           // measurement aid: 1 = skip the DSMEM stores (results are wrong)
    int prefetch_ahead;      // > 0: during its PC phase a CTA prefetches its slice of pulse p + prefetch_ahead into L2 (the
                             // cluster that will take that pulse when this wave retires then streams from L2, not HBM)
};

__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_relaxed() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// the same copy with an L2 eviction policy (createpolicy): streaming input that nobody reads twice should not push the
// chain's L2-resident intermediates out
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

// One overlap-save block of the CTA's line (shared memory) by the group of Cfg::T threads that starts at warp G.warp0.
template <class Cfg>
__device__ __forceinline__ void dbf_pc_block(const DbfPcArgs& k, const DbfPcGroup& G, const float2* line, float2* out_line,
                                             float2* work, const float2* tables, const float* sfir, int tid) {
    const PcSegArgs& sg = k.seg[G.seg];
    const int t = tid - G.warp0 * 32;
    float2* s = work + G.work_off;
    PcBlockArgs a;
    a.line = line;
    a.out_line = out_line;
    a.tw1 = sg.tw1;
    a.tw2 = tables + k.tw2_off[G.seg];
    a.Hmid = sg.Hmid;
    a.in_lo = sg.in_lo;
    a.in_hi = sg.in_hi;
    a.seg_start0 = sg.seg_start0;
    a.taps = sg.taps;
    a.g0 = sg.gate0 + G.blk * sg.valid;
    a.g_end = sg.g_end;
    const int bar = G.warp0 + 1;                       // named barrier of the group (0 is __syncthreads)
    pc_phase_load_pass1<Cfg>(a, s, t);
    pc_group_sync(true, bar - 1, Cfg::T);
    pc_phase_pass2<Cfg>(a, s, t);
    pc_group_sync(true, bar - 1, Cfg::T);
    pc_phase_mid<Cfg>(a, s, t);
    pc_group_sync(true, bar - 1, Cfg::T);
    pc_phase_ipass2<Cfg>(a, s, t);
    pc_group_sync(true, bar - 1, Cfg::T);
    pc_phase_ipass1_store<Cfg>(a, s, t);
    if (G.narrow) {                                    // fun_process_single_frame.m:111-112,123 straight from the resident line
        const bool fast = k.narrow_gates + k.fir_delay <= k.N - k.narrow_start0;
        for (int g = t; g < k.narrow_gates; g += Cfg::T)
            out_line[g] = fast ? pc_narrow_gate_smem(line + k.narrow_start0, sfir, k.nfir, k.fir_delay, g)
                               : pc_narrow_gate(line, k.N, k.narrow_start0, sfir, k.nfir, k.fir_delay, g);
    }
}

template <int KS>     // k-steps of 4 channels (4 for C <= 16); B <= 8 beams = one m16 tile of (Re, Im) rows
__global__ void __launch_bounds__(RSP_FUSED_THREADS, 2) dbf_pc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                      const __grid_constant__ DbfPcArgs k) {
    extern __shared__ __align__(1024) unsigned char fsm[];
    float2* const line = reinterpret_cast<float2*>(fsm);
    unsigned char* const ring = fsm + k.line_bytes;                       // DBF phase
    float2* const work = reinterpret_cast<float2*>(fsm + k.line_bytes);   // PC phase (same bytes)
    float2* const tables = reinterpret_cast<float2*>(fsm + k.line_bytes + k.work_bytes);
    const int n_tab = k.tw2_off[3] + k.tw2_len[3];
    float* const sfir = reinterpret_cast<float*>(tables + n_tab);
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(sfir + 256);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int C = k.C, B = k.B, N = k.N;
    const uint32_t rank = cluster_ctarank();
    const int p = blockIdx.x / B;
    long long* const dbg = k.dbg ? k.dbg + (size_t)blockIdx.x * 8 : nullptr;
    if (dbg && tid == 0) { uint32_t sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); dbg[0] = global_ns(); dbg[7] = sm; }
    if (tid == 0) {
        for (int s = 0; s < RSP_FUSED_STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#pragma unroll 1
    for (int sgi = 0; sgi < 4; ++sgi)
        for (int i = tid; i < k.tw2_len[sgi]; i += RSP_FUSED_THREADS) tables[k.tw2_off[sgi] + i] = k.seg[sgi].tw2[i];
    for (int i = tid; i < k.nfir; i += RSP_FUSED_THREADS) sfir[i] = k.fir[i];
    __syncthreads();

    // ------------------------------------------------------------------------------------ DBF phase
    const int t_lo = (int)((long)rank * k.tiles / B), t_hi = (int)((long)(rank + 1) * k.tiles / B);
    const int n_my = t_hi - t_lo;
    const uint32_t stage_bytes = (uint32_t)(RSP_FUSED_TILE / 16) * C * 128u;       // [8 groups][C][16 samples]
    auto issue = [&](int i) {                                                      // thread 0 only
        const int s = i % RSP_FUSED_STAGES;
        const uint32_t full = smem_u32(&bars[s]);
        const uint32_t dst = smem_u32(ring) + (uint32_t)s * stage_bytes;
        const int grp0 = (t_lo + i) * (RSP_FUSED_TILE / 16);
        mbar_expect_tx(full, stage_bytes);
        if (k.tma_rank4) {
            tma_load_4d(dst, &tmap, 0, 0, grp0, p, full);
        } else {
#pragma unroll 1
            for (int j = 0; j < RSP_FUSED_TILE / 16; ++j) tma_load_2d(dst + (uint32_t)j * C * 128u, &tmap, 32 * (grp0 + j), p * C, full);
        }
    };
    if (tid == 0)
        for (int i = 0; i < RSP_FUSED_STAGES && i < n_my; ++i) issue(i);
    // every CTA of the cluster is running before anyone writes into it; nothing to publish yet, so no release / acquire
    // (a releasing arrive would wait for every global access issued so far)
    cluster_arrive_relaxed();
    cluster_wait_relaxed();
    if (dbg && tid == 0) dbg[1] = global_ns();

    const int sg_ = (g & 1) ? g + 7 : g;                // sample of the group this lane loads (dbf_mma2_kernel's column order)
    const uint32_t remote_line = g < B ? map_to_cta(smem_u32(line), (uint32_t)g) : 0u;   // beam g's line lives in CTA g
    float4 ah[KS], al[KS];                              // weight fragments stay in registers for the whole slice
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        ah[s] = __ldg(k.Wa + (s * 2 + 0) * 32 + lane);
        al[s] = __ldg(k.Wa + (s * 2 + 1) * 32 + lane);
    }
#pragma unroll 1
    for (int i = 0; i < n_my; ++i) {
        const int s = i % RSP_FUSED_STAGES;
        mbar_wait(smem_u32(&bars[s]), (uint32_t)(i / RSP_FUSED_STAGES) & 1u);
        const unsigned char* st = ring + (size_t)s * stage_bytes + (size_t)w * C * 128;       // 16-sample group w of the tile
        float4 x[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int c = 4 * ks + t;
            x[ks] = c < C ? *reinterpret_cast<const float4*>(st + c * 128 + sg_ * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // The refill below is an async-proxy write: it must not start before the loads above have RETURNED (a barrier
        // only orders them against generic-proxy writes, and an LDS queued behind the DSMEM stores of the previous tile
        // can take longer than a TMA copy that hits L2).  Making the barrier's predicate depend on every loaded
        // register forces the wait: 8 LOP3 per tile.
        unsigned dep = 0u;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
            dep |= __float_as_uint(x[ks].x) | __float_as_uint(x[ks].y) | __float_as_uint(x[ks].z) | __float_as_uint(x[ks].w);
        __syncthreads_or(dep == 0x7FB1C0DEu);           // a NaN payload no sample carries; the result is not used
        if (tid == 0 && i + RSP_FUSED_STAGES < n_my) issue(i + RSP_FUSED_STAGES);
        float E[4] = {0.f, 0.f, 0.f, 0.f}, O[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const float v[4] = {x[ks].x, x[ks].y, x[ks].z, x[ks].w};
            uint32_t bh[4], bl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bh[j] = __float_as_uint(v[j]) & 0xFFFFE000u;
                bl[j] = __float_as_uint(v[j] - __uint_as_float(bh[j])) & 0xFFFFE000u;
            }
            // even n-tile: (re, im) of sample sg, odd n-tile: sample sg + 1; same order of the three products as dbf_mma2_kernel
            mma_tf32_wa(E, al[ks], bh[0], bh[1]);
            mma_tf32_wa(E, ah[ks], bl[0], bl[1]);
            mma_tf32_wa(E, ah[ks], bh[0], bh[1]);
            mma_tf32_wa(O, al[ks], bh[2], bh[3]);
            mma_tf32_wa(O, ah[ks], bl[2], bl[3]);
            mma_tf32_wa(O, ah[ks], bh[2], bh[3]);
        }
        const int n = (t_lo + i) * RSP_FUSED_TILE + w * 16 + 2 * t;
        if (g < B && !(k.dbg_flags & 1)) {
            if (n < N) st_cluster_v4(remote_line + (uint32_t)n * 8u, make_float4(E[0], E[2], O[0], O[2]));
            if (n + 8 < N) st_cluster_v4(remote_line + (uint32_t)(n + 8) * 8u, make_float4(E[1], E[3], O[1], O[3]));
        }
    }
    if (dbg && tid == 0) dbg[2] = global_ns();
    cluster_arrive();                                   // release: my stores into the other CTAs' lines
    cluster_wait();                                     // acquire: everybody's stores into mine
    if (dbg && tid == 0) dbg[3] = global_ns();

    // ------------------------------------------------------------------------------------ PC phase
    // housekeeping that must not sit in front of a releasing barrier (the arrive waits for these stores): the dead
    // amplitude map of the lane's previous CPI is dropped from L2, the detection counter of this CPI is cleared
    l2_discard(k.dead);
    if (k.det_count && blockIdx.x == 0 && tid == 0) *k.det_count = 0;
    const size_t line_id = (size_t)p * B + rank;
    if (k.beam_out) {                                   // debug / rsp_get_beam copy (single-CPI path only)
        float4* dst = reinterpret_cast<float4*>(k.beam_out + line_id * k.ldb);
        const float4* src = reinterpret_cast<const float4*>(line);
        for (int i = tid; i < N / 2; i += RSP_FUSED_THREADS) dst[i] = src[i];
    }
    if (tid == 0 && k.prefetch_ahead > 0 && p + k.prefetch_ahead < k.P) {
        const int pn = p + k.prefetch_ahead;
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int grp0 = (t_lo + i) * (RSP_FUSED_TILE / 16);
            if (k.tma_rank4) tma_prefetch_4d(&tmap, 0, 0, grp0, pn);
            else
                for (int j = 0; j < RSP_FUSED_TILE / 16; ++j) tma_prefetch_2d(&tmap, 32 * (grp0 + j), pn * C);
        }
    }
    float2* out_line = k.pc + line_id * k.ldg;
#pragma unroll 1
    for (int r = 0; r < k.n_rounds; ++r) {
        if (r) __syncthreads();                         // the workspace is reused by the next round
        const int gi = k.warp_group[r][w];
        if (gi >= 0) {
            const DbfPcGroup& G = k.grp[r][gi];
            if (G.cfg == 3) dbf_pc_block<PcCfg<4096, 16, 16, 16>>(k, G, line, out_line, work, tables, sfir, tid);
            else if (G.cfg == 2) dbf_pc_block<PcCfg<2048, 8, 16, 16>>(k, G, line, out_line, work, tables, sfir, tid);
            else dbf_pc_block<PcCfg<1024, 16, 16, 4>>(k, G, line, out_line, work, tables, sfir, tid);
        }
        if (dbg && tid == 0 && r < 3) dbg[4 + r] = global_ns();       // thread 0's group only
    }
}


// ------------------------------------------------------------------------------------------
// S5 alone, fed by TMA tensor copies (the DBF phase of dbf_pc_kernel as its own kernel; shapes the fused kernel does not
// cover, and the two-kernel A/B path): CTA = 8 warps, one slice of `tiles_per_cta` consecutive 128-sample tiles of one
// pulse, NS-stage ring, one [8 groups][C][16 samples] tile per copy, fragment loads conflict free (4 wavefronts per
// 512 B instead of the 16 a fragment-order ld.global costs), beam rows written as 64-byte pieces like dbf_mma2_kernel.
// ------------------------------------------------------------------------------------------
#define RSP_DBFT_STAGES 3
struct DbfTmaArgs {
    float2* beam;
    const float4* Wa;
    int* det_count;
    int C, B, N, ldb, tiles, tiles_per_cta, tma_rank4;
    DiscardArgs dead;
};
template <int KS>
__global__ void __launch_bounds__(RSP_FUSED_THREADS, 4) dbf_tma2_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                        const __grid_constant__ DbfTmaArgs k) {
    extern __shared__ __align__(1024) unsigned char tsm[];
    __shared__ __align__(8) unsigned long long bars[RSP_DBFT_STAGES];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int C = k.C, B = k.B, N = k.N, p = blockIdx.y;
    if (k.det_count && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *k.det_count = 0;
    if (tid == 0) {
        for (int s = 0; s < RSP_DBFT_STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int t_lo = blockIdx.x * k.tiles_per_cta, t_hi = min(k.tiles, t_lo + k.tiles_per_cta), n_my = t_hi - t_lo;
    const uint32_t stage_bytes = (uint32_t)(RSP_FUSED_TILE / 16) * C * 128u;
    auto issue = [&](int i) {
        const int s = i % RSP_DBFT_STAGES;
        const uint32_t full = smem_u32(&bars[s]);
        const uint32_t dst = smem_u32(tsm) + (uint32_t)s * stage_bytes;
        const int grp0 = (t_lo + i) * (RSP_FUSED_TILE / 16);
        mbar_expect_tx(full, stage_bytes);
        if (k.tma_rank4) tma_load_4d(dst, &tmap, 0, 0, grp0, p, full);
        else
            for (int j = 0; j < RSP_FUSED_TILE / 16; ++j) tma_load_2d(dst + (uint32_t)j * C * 128u, &tmap, 32 * (grp0 + j), p * C, full);
    };
    if (tid == 0)
        for (int i = 0; i < RSP_DBFT_STAGES && i < n_my; ++i) issue(i);
    l2_discard(k.dead);
    const int sg_ = (g & 1) ? g + 7 : g;
    float4 ah[KS], al[KS];
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        ah[s] = __ldg(k.Wa + (s * 2 + 0) * 32 + lane);
        al[s] = __ldg(k.Wa + (s * 2 + 1) * 32 + lane);
    }
    float2* const brow = k.beam + ((size_t)p * B + g) * k.ldb;
#pragma unroll 1
    for (int i = 0; i < n_my; ++i) {
        const int s = i % RSP_DBFT_STAGES;
        mbar_wait(smem_u32(&bars[s]), (uint32_t)(i / RSP_DBFT_STAGES) & 1u);
        const unsigned char* st = tsm + (size_t)s * stage_bytes + (size_t)w * C * 128;
        float4 x[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int c = 4 * ks + t;
            x[ks] = c < C ? *reinterpret_cast<const float4*>(st + c * 128 + sg_ * 8) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        unsigned dep = 0u;                              // see dbf_pc_kernel: the loads must have returned before the refill
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
            dep |= __float_as_uint(x[ks].x) | __float_as_uint(x[ks].y) | __float_as_uint(x[ks].z) | __float_as_uint(x[ks].w);
        __syncthreads_or(dep == 0x7FB1C0DEu);
        if (tid == 0 && i + RSP_DBFT_STAGES < n_my) issue(i + RSP_DBFT_STAGES);
        float E[4] = {0.f, 0.f, 0.f, 0.f}, O[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const float v[4] = {x[ks].x, x[ks].y, x[ks].z, x[ks].w};
            uint32_t bh[4], bl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bh[j] = __float_as_uint(v[j]) & 0xFFFFE000u;
                bl[j] = __float_as_uint(v[j] - __uint_as_float(bh[j])) & 0xFFFFE000u;
            }
            mma_tf32_wa(E, al[ks], bh[0], bh[1]);
            mma_tf32_wa(E, ah[ks], bl[0], bl[1]);
            mma_tf32_wa(E, ah[ks], bh[0], bh[1]);
            mma_tf32_wa(O, al[ks], bh[2], bh[3]);
            mma_tf32_wa(O, ah[ks], bl[2], bl[3]);
            mma_tf32_wa(O, ah[ks], bh[2], bh[3]);
        }
        const int n = (t_lo + i) * RSP_FUSED_TILE + w * 16 + 2 * t;
        if (g < B) {
            if (n < N) *reinterpret_cast<float4*>(brow + n) = make_float4(E[0], E[2], O[0], O[2]);
            if (n + 8 < N) *reinterpret_cast<float4*>(brow + n + 8) = make_float4(E[1], E[3], O[1], O[3]);
        }
    }
}


#ifdef RSP_PROBES
// ------------------------------------------------------------------------------------------
// Experiment (RSP_EXP_MERGE=1, tools/stage_probe.py): do a memory-bound and an issue-bound stage overlap when their CTAs
// are guaranteed to share SMs?  One launch holds the CTAs of pc_fft_kernel AND CTAs that run the DBF of dbf_mma2_kernel
// (256 samples of one pulse each) on a scratch beam buffer, interleaved evenly in block-index order.
// ------------------------------------------------------------------------------------------
struct MergeDbfArgs {
    const float2* raw;
    float2* beam;
    const float4* Wa;
    int C, NB, N, ldb, P;
    int n_dbf, n_total;      // DBF CTAs / all CTAs of the launch
};
template <int KS>
__device__ __forceinline__ void dbf_role_256(const MergeDbfArgs& k, int idx) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
    const int per_pulse = (k.N + 255) / 256;
    const int p = idx / per_pulse, n_base = (idx - p * per_pulse) * 256 + w * 32;
    if (p >= k.P || n_base >= k.N) return;
    const int sg = (g & 1) ? g + 7 : g;
    const float2* rp = k.raw + (size_t)p * k.C * k.N + sg + (unsigned)(t * k.N);
    float4 x[KS][2];
#pragma unroll
    for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            x[s][q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * s + t < k.C && n_base + 16 * q + sg < k.N) x[s][q] = __ldcs(reinterpret_cast<const float4*>(rp + s * 4u * (unsigned)k.N + n_base + 16 * q));
        }
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float v[4] = {x[s][q].x, x[s][q].y, x[s][q].z, x[s][q].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t hi = __float_as_uint(v[i]) & 0xFFFFE000u;
                bh[2 * q + (i >> 1)][i & 1] = hi;
                bl[2 * q + (i >> 1)][i & 1] = __float_as_uint(v[i] - __uint_as_float(hi)) & 0xFFFFE000u;
            }
        }
        const float4 ah = __ldg(k.Wa + (s * 2 + 0) * 32 + lane), al = __ldg(k.Wa + (s * 2 + 1) * 32 + lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mma_tf32_wa(acc[j], al, bh[j][0], bh[j][1]);
            mma_tf32_wa(acc[j], ah, bl[j][0], bl[j][1]);
            mma_tf32_wa(acc[j], ah, bh[j][0], bh[j][1]);
        }
    }
    if (g < k.NB) {
        float2* row = k.beam + ((size_t)p * k.NB + g) * k.ldb + 2 * t + n_base;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int n = n_base + 16 * q + 2 * t;
            if (n < k.N) *reinterpret_cast<float4*>(row + 16 * q) = make_float4(acc[2 * q][0], acc[2 * q][2], acc[2 * q + 1][0], acc[2 * q + 1][2]);
            if (n + 8 < k.N) *reinterpret_cast<float4*>(row + 16 * q + 8) = make_float4(acc[2 * q][1], acc[2 * q][3], acc[2 * q + 1][1], acc[2 * q + 1][3]);
        }
    }
}
template <class CfgA, class CfgB>
__global__ void __launch_bounds__(RSP_PC_THREADS, 3) pc_dbf_merge_kernel(const PcKernelArgs k, const MergeDbfArgs d) {
    extern __shared__ float2 pc_smem[];
    const long i = blockIdx.x;
    const int before = (int)(i * d.n_dbf / d.n_total), after = (int)((i + 1) * d.n_dbf / d.n_total);
    if (after > before) { dbf_role_256<4>(d, before); return; }
    const int cta = (int)i - before;
    if (cta < k.seg[0].n_ctas) pc_role<CfgA>(k, k.seg[0], cta, pc_smem, false);
    else pc_role<CfgB>(k, k.seg[1], cta - k.seg[0].n_ctas, pc_smem, k.do_narrow != 0);
}

#endif  // RSP_PROBES

}  // namespace rsp
