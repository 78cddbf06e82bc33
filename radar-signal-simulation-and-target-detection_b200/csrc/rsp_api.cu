// rsp_api.cu -- the C ABI of include/rsp.h: context, constant upload, per-CPI launch sequence.
// There is no CPU fallback: without a CUDA device rsp_create fails with RSP_ERR_NO_DEVICE.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rsp.h"
#include "rsp_kernels.cuh"
#include "rsp_fused.cuh"
#include "rsp_dbf_tc.cuh"
#include "rsp_cfar1d.cuh"
#include "rsp_plan.hpp"

using namespace rsp;

static thread_local std::string g_create_error;

// Measurement aids (stage masks, stream priorities, uniform carve-outs, kernels with parts switched off) exist only in a
// library built with -DRSP_PROBES (build.py --variant probes -DRSP_PROBES; tools/stage_probe.py picks it up): the shipped
// library cannot be told by its environment to skip stages or to produce wrong results.
#ifdef RSP_PROBES
static int probe_env(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
#else
static int probe_env(const char*, int dflt) { return dflt; }
#endif

struct rsp_ctx {
    rsp_params prm{};
    int C = 0, B = 0, P = 0, N = 0, G = 0;
    int ldb = 0, ldg = 0;                 // row pitches (complex elements) of beam / pc
    bool have_constants = false;
    bool ran = false;
    bool rdm_in_ctx = false;
    bool own_stream = true;
    bool pow2_doppler = false;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;

    // device buffers
    float2* d_raw = nullptr;              // PCN complex64 staging for host / converted input
    void* d_stage = nullptr;              // raw staging for dtype/layout conversion
    size_t stage_bytes = 0;
    // A lane = one CUDA stream + its own intermediates.  The single-CPI calls use lane 0; the stream
    // path deals consecutive CPIs round-robin over n_lanes so that one CPI's memory-bound kernels
    // and tails overlap another CPI's FP-bound kernels.
    struct Lane {
        cudaStream_t s = nullptr;
        float2* beam = nullptr;
        float2* pc = nullptr;
        float* amp = nullptr;
        float2* raw = nullptr;            // staging cube of the pipelined host-input / frame paths (lazy)
        float2* rdm = nullptr;            // own RDM of those paths when the caller passes none (lazy; lane 0 uses d_rdm)
        unsigned long long seen_epoch = 0;  // last caller_epoch this lane was ordered behind (pipelined paths)
        cudaEvent_t done = nullptr;
    };
    std::vector<int> slot_lane;           // lane that last produced each ring slot (-1: joined to the caller's stream)
    Lane lanes[8];
    int n_lanes = 1;
    // Bumped whenever tables the kernels read are (re)written on the caller's stream or that stream changes; a
    // pipelined submission orders its lane behind the caller's stream only when the lane has not seen the
    // current epoch -- doing it per submission would serialise lane l behind everything queued on lane 0.
    unsigned long long caller_epoch = 1;
    struct GraphEntry {                   // one captured rsp_stream_enqueue batch (stream_enqueue_graphed)
        const void* raw; void* rdm; int raw_pool, rdm_pool, n_cpi, first_slot, flags;
        cudaGraphExec_t exec; bool failed; long launches;
    };
    std::vector<GraphEntry> graphs;
    unsigned long long graph_epoch = 0;
    long graph_launches = 0;
    Lane* cur = nullptr;                  // lane the launch helpers enqueue on
    bool discard = false;                 // stream path: drop dead intermediates from L2 (see l2_discard)
    cudaEvent_t fork = nullptr;
    float2* d_rdm = nullptr;
    float2* d_aux = nullptr;              // scratch for stage2 transposes
    // constants
    float2* d_W = nullptr;
    float4* d_Wfrag = nullptr;            // tensor-core DBF weight fragments
    float4* d_Wfrag_wa = nullptr;         // same for dbf_mma2_kernel (weights as the A operand)
    int pc_group_bar = 1;                 // RSP_PC_GROUP_BAR at rsp_create: per-group named barriers in pc_fft_kernel
    // S5 + S6 in one launch (dbf_pc_kernel, rsp_fused.cuh): cluster per pulse, TMA-fed DBF, beam lines in shared memory
    bool dbf_pc_ok = false;               // the shape and block plan fit the fused kernel (else: dbf + pc_fft launches)
    bool dbf_pc_enabled = false;          // RSP_FUSE_DBF_PC=1 selects the fused cluster kernel
    bool keep_beam = false;               // single-CPI path: the fused kernel also writes the beam cube for rsp_get_beam
    DbfPcArgs dbf_pc{};                   // everything but the per-launch pointers
    size_t dbf_pc_smem = 0;
    int dbf_pc_tma_rank4 = 1;
    // tcgen05 DBF (dbf_tc_kernel, rsp_dbf_tc.cuh)
    bool dbf_tc = false;                  // selected for this shape (RSP_DBF=tc forces it on, any other RSP_DBF value off)
    float* d_Bw_tc = nullptr;             // [2][Npad * Kpad] weight operands, canonical K-major core-matrix layout
    DbfTcArgs dbf_tc_args{};
    size_t dbf_tc_smem = 0;
    int dbf_tc_grid = 0;
    // Range-blocked chain (shapes whose beam + pulse-compressed cubes do not fit L2, e.g. config 3): the CPI is processed in
    // chunks of one overlap-save block's gates, DBF -> PC -> MTD -> CFAR per chunk, so every intermediate is consumed while
    // it is still in L2 instead of making a round trip through HBM
    struct Chunk { int tile_lo, tile_hi; int part, blk; bool medium; int g_lo, g_hi; int cut_lo, cut_hi; };
    std::vector<Chunk> chunks;
    bool blocked = false;
    int block_lanes = 1;
    bool dbf_tma2 = false;                // RSP_DBF=tma2: stand-alone TMA-tensor-fed DBF (dbf_tma2_kernel)
    int dbf_tma2_tiles = 8;               // tiles per CTA (RSP_DBF_TMA2_TILES)
    const float2* exp_raw = nullptr;      // RSP_EXP_MERGE experiment: the cube whose DBF rides inside the PC launch
    float2* exp_beam = nullptr;
    long exp_multi_count = 0;             // RSP_EXP_DBF_MULTI experiment
    int dbf_pc_clusters = 0;              // cudaOccupancyMaxActiveClusters of the fused kernel
    struct TmapEntry { const void* ptr; int mode; CUtensorMap map; };
    std::vector<TmapEntry> tmaps;         // tensor maps of the raw cubes seen so far (keyed by device pointer)
    long long* d_fused_dbg = nullptr;     // RSP_FUSED_DEBUG: phase timestamps of the last dbf_pc launch [grid][8]
    int fused_dbg_flags = -1;
    const SynthArgs* fused = nullptr;     // set while rsp_submit_targets enqueues a frame whose S4 is fused into the DBF
    bool fuse_synth = true;               // dbf_synth_kernel on the pipelined frame path (<= 8 targets); RSP_FUSE_SYNTH=0: two kernels
    int stages = 15;                      // RSP_STAGES at rsp_create (measurement aid, see enqueue_chain)
    bool dbf_wa = true;                   // RSP_DBF=mma selects the older data-as-A kernel
    int dbf_nt = 0, dbf_ks = 0;           // 0 = FFMA kernel
    float* d_fir = nullptr;
    int n_fir = 0;
    PcPlan med, lng;
    // stage-2 (process_stage2_mtd) plans: one per gated segment, own tables
    struct S2Seg { PcPlan pl; float2 *tw1 = nullptr, *tw2 = nullptr, *H = nullptr; int lo = 0, hi = 0; };
    S2Seg s2[3];
    bool s2_ready = false;
    int s2_notch = 0;
    float* d_s2_win = nullptr;
    float2 *d_med_tw1 = nullptr, *d_med_tw2 = nullptr, *d_med_H = nullptr;
    float2 *d_lng_tw1 = nullptr, *d_lng_tw2 = nullptr, *d_lng_H = nullptr;
    // mixed block plan of the long segment (choose_pc_mix): `lng` covers the first gates with the longest blocks, the
    // parts below continue with shorter blocks and go out in a second launch
    PcPlan lngx[2];
    float2 *d_lngx_tw1[2] = {nullptr, nullptr}, *d_lngx_tw2[2] = {nullptr, nullptr}, *d_lngx_H[2] = {nullptr, nullptr};
    DopplerPlan dop;
    float2* d_dop_tw = nullptr;
    int dop_tw_count = 0;
    int* d_dop_perm = nullptr;
    float* d_win = nullptr;
    std::vector<float> h_win, h_s2_win;   // host copies: the register MTD kernel takes the window as an argument
    bool cfar_pad = true;                 // conflict-free CFAR tile pitches (P <= 64)
    int mtd_mode = 2;                     // RSP_MTD: 0 = tile (generic), 2 = p64 (specialised tile; default when P = 64)
    bool mtd_approx_sqrt = false;         // RSP_MTD_SQRT=approx
    double *d_range_axis = nullptr, *d_vel_axis = nullptr, *d_beam_angles = nullptr, *d_k_slopes = nullptr;
    double delta_r = 0, delta_v = 0;
    // S4 synthesis (rsp_set_waveform)
    float2* d_tx = nullptr;
    SynthTarget* d_tg = nullptr;
    int tg_cap = 0;
    SynthTarget* d_tg_ring = nullptr;     // [slots][RSP_MAX_FRAME_TARGETS] descriptors of rsp_submit_targets
    SynthTarget* h_tg_ring = nullptr;     // pinned mirror
    rsp_waveform wf{};
    int tx_seg_lo[3] = {0, 0, 0}, tx_seg_hi[3] = {0, 0, 0};
    bool have_waveform = false;
    // detection ring
    int slots = 0;
    int* d_counts = nullptr;
    RawDet* d_rawdet = nullptr;           // [slots][max_detections] what cfar_kernel records
    rsp_detection* d_recs = nullptr;
    int* h_count = nullptr;               // pinned
    rsp_detection* h_recs = nullptr;      // pinned [max_detections]
    // pipelined paths (rsp_submit_cpi / rsp_submit_targets): the count and the first prefetch_k records of a slot
    // are copied to pinned memory behind the slot's kernels, so the fetch is one event wait and no round trip
    int prefetch_k = 0;
    std::vector<std::pair<uint64_t, int>> sort_keys;      // scratch of sort_detections_inplace
    std::vector<rsp_detection> sort_tmp;
    cudaStream_t fetch_stream = nullptr;  // late records of dense frames (fetch_slot)
    int* h_slot_count = nullptr;          // pinned [slots]
    rsp_detection* h_slot_recs = nullptr; // pinned [slots][prefetch_k]
    std::vector<cudaEvent_t> slot_done;   // recorded behind the prefetch copies
    std::vector<char> slot_prefetched;
    int mtd_tg = 32, mtd_r = 1, mtd_kt = 1, cfar_tg = 32, cfar_variant = 0;
    bool cfar_vec = false;
    int pulse_block = 0;                  // > 0: pulses per group of the pulse-blocked S5 -> S6 path (enqueue_chain)
    bool cfar5 = false;                   // cfar5_kernel (marching) instead of cfar4_kernel
    size_t mtd_smem = 0, cfar_smem = 0;
    // per-kernel event timing (rsp_set_profiling)
    bool profiling = false;
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
};

enum KernelClass { K_CONVERT = 0, K_DBF, K_PC_NARROW, K_PC, K_MTD, K_CFAR, K_SYNTH, K_REFINE, K_DBF_PC, K_MTD_CFAR, K_NCLASS };
static const char* kKernelNames[K_NCLASS] = {"convert", "dbf", "pc_narrow", "pc_fft", "mtd", "cfar", "synth", "refine", "dbf_pc", "mtd_cfar"};
static_assert(K_NCLASS <= RSP_MAX_KERNEL_CLASSES, "rsp_kernel_times is too small");

static cudaEvent_t take_event(rsp_ctx* c) {
    if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
// RAII bracket: records an event before and after one launch when profiling is on
struct Timed {
    rsp_ctx* c; int cls; cudaEvent_t a = nullptr;
    Timed(rsp_ctx* c_, int cls_) : c(c_), cls(cls_) {
        if (c->profiling) { a = take_event(c); cudaEventRecord(a, c->cur->s); }
    }
    ~Timed() {
        c->launches++;
        if (c->profiling) { cudaEvent_t b = take_event(c); cudaEventRecord(b, c->cur->s); c->spans.push_back({cls, a, b}); }
    }
};

static int fail(rsp_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(ctx, call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return fail(ctx, RSP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <typename T> static cudaError_t dev_alloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}
template <typename T> static cudaError_t upload(T** dptr, const std::vector<T>& h) {
    if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
    cudaError_t e = dev_alloc(dptr, h.size() ? h.size() : 1);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// RSP_CARVEOUT=1 (experiment, off by default): every chain kernel asks for the same L1 / shared-memory split (maximum
// shared memory), so that an SM never has to drain to change its carve-out between kernels of different lanes.  Measured:
// it does not make the kernels of different streams overlap any better (DBF + PC together still cost the sum of their
// times) and the smaller L1 slows both (DBF 15.3 -> 21.0 us, PC 21.8 -> 24.2 us), profiles/r1_overlap_probe.txt.
static int carveout_pct() { static const int v = probe_env("RSP_CARVEOUT", 0); return v == 1 ? 100 : v; }
template <typename KernelT> static void prefer_max_smem(KernelT kern) {
    if (carveout_pct() > 0) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct());
}
template <typename KernelT> static cudaError_t opt_in_smem(KernelT kern, size_t bytes) {
    prefer_max_smem(kern);
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Cap a kernel's CTAs per SM by asking for more dynamic shared memory than it uses: co-resident kernels
// of other lanes then keep their share of the SM (measured: CFAR at 4 CTAs/SM costs the chain 10 %).
static size_t smem_for_occupancy(size_t needed, int max_ctas_per_sm) {
    if (max_ctas_per_sm <= 0) return needed;
    // RSP_SMEM_RESERVE_KB: shared memory left free beside the capped CTAs (room for a co-resident kernel of another stream)
    static const size_t reserve = (size_t)probe_env("RSP_SMEM_RESERVE_KB", 0) * 1024;
    const size_t per_sm = 227 * 1024 - reserve;
    size_t pad_to = per_sm / (size_t)(max_ctas_per_sm + 1) + 1024;               // max_ctas + 1 no longer fit
    pad_to = std::min(pad_to, per_sm / (size_t)max_ctas_per_sm - 2048);          // ... but max_ctas still do
    return std::max(needed, pad_to);                                             // never below what the kernel uses
}

typedef PcCfg<1024, 16, 16, 4> Pc1024;
typedef PcCfg<2048, 8, 16, 16> Pc2048;
typedef PcCfg<4096, 16, 16, 16> Pc4096;
template <class Cfg> static size_t pc_smem_bytes() {
    return ((size_t)Cfg::NG * Cfg::SMEM_ELEMS + (Cfg::R2 - 1) * Cfg::SPAN2) * sizeof(float2) + 256 * sizeof(float);
}
// RSP_OCC_PC caps the pulse-compression CTAs per SM (shared-memory padding), leaving registers for the kernels of other lanes
static int pc_occ_cap() { static const int v = probe_env("RSP_OCC_PC", 0); return v; }
template <class A, class B> static size_t pc_smem_pair() { return smem_for_occupancy(std::max(pc_smem_bytes<A>(), pc_smem_bytes<B>()), pc_occ_cap()); }
// X(long plan, medium plan)
#define RSP_FOR_EACH_PC_PAIR(X) X(Pc1024, Pc1024) X(Pc2048, Pc1024) X(Pc4096, Pc1024) X(Pc1024, Pc2048) X(Pc2048, Pc2048) \
    X(Pc4096, Pc2048) X(Pc1024, Pc4096) X(Pc2048, Pc4096) X(Pc4096, Pc4096)

// generic Doppler DFT kernel: (tile gates, power-of-two factor R of P)
// (tile gates, power-of-two factor R of P, output bins per work item KT)
#define RSP_FOR_EACH_DFT_TR(X, kt) X(32, 1, kt) X(32, 2, kt) X(32, 4, kt) X(16, 1, kt) X(16, 2, kt) X(16, 4, kt) X(8, 1, kt) X(8, 2, kt) X(8, 4, kt)
#define RSP_FOR_EACH_DFT(X) RSP_FOR_EACH_DFT_TR(X, 1) RSP_FOR_EACH_DFT_TR(X, 4) RSP_FOR_EACH_DFT_TR(X, 6) RSP_FOR_EACH_DFT_TR(X, 11) \
    X(32, 8, 1) X(16, 8, 1) X(8, 8, 1) X(32, 8, 4) X(16, 8, 4) X(8, 8, 4)                                                          \
    RSP_FOR_EACH_DFT_TR(X, 103) RSP_FOR_EACH_DFT_TR(X, 104) RSP_FOR_EACH_DFT_TR(X, 106)      /* odd Q: 3 / 4 / 6 bin pairs per item */
#define RSP_FOR_EACH_POW2_P(X) X(8, 8, 1, 1) X(16, 16, 1, 1) X(32, 8, 4, 1) X(64, 8, 8, 1) X(128, 16, 8, 1) X(256, 16, 16, 1) X(512, 8, 8, 8)
// Tile width (gates) and output bins per work item of the generic Doppler DFT kernel.  The kernel's cost per
// gate is (rounds of the CTA over its work items) x (bins per item) / (tile gates), divided by how many CTAs
// fit an SM (the load and read-out phases of one tile hide behind the sums of another; measured on the
// native P = 332 = 4 x 83: (32, 11) 455 us, (16, 6) 376 us, (16, 4) 452 us, (8, 4) 370 us).  Near-ties go to
// the wider tile (longer contiguous stretches in the pulse-compressed cube).
static size_t dft_smem_bytes(int P, int tg, int R = 1, int kt = 1) {    // input tile + output tile (one tile for the in-place folded form) + twiddles
    const bool one = kt >= 100 && mtd_dft_sym_inplace(P / R, kt - 100, tg, RSP_MTD_THREADS);
    return ((size_t)(one ? 1 : 2) * P * (tg + 1) + P) * sizeof(float2);
}
static bool choose_dft_plan(int P, int R, int* tg_out, int* kt_out) {
    const int Q = P / R;
    const int tg_env = getenv("RSP_DFT_TG") ? atoi(getenv("RSP_DFT_TG")) : 0;
    const int kt_env = getenv("RSP_DFT_KT") ? atoi(getenv("RSP_DFT_KT")) : 0;
    double best = -1.0;
    for (int tg : {32, 16, 8}) {
        if (tg_env && tg != tg_env) continue;
        // odd Q >= 7: the folded form (kt = 100 + bin pairs per item) does a quarter of the multiply-adds; RSP_DFT_SYM=0 keeps the direct sums
        static const bool sym_ok = [] { const char* e = getenv("RSP_DFT_SYM"); return !(e && atoi(e) == 0); }();
        const bool sym = sym_ok && (Q & 1) && Q >= 7 && R <= 4;
        for (int kt : {1, 4, 6, 11, 103, 104, 106}) {
            if ((kt >= 100) != sym && !(kt_env && kt == kt_env && (kt < 100 || ((Q & 1) && R <= 4)))) continue;
            if ((kt > 4 && R > 4) || (kt_env ? kt != kt_env : kt == 1)) continue;     // kt = 1 only on request
            const int kp = kt >= 100 ? kt - 100 : 0;
            const size_t sm = dft_smem_bytes(P, tg, R, kt);
            if (sm > 200 * 1024) continue;
            const int regs = kt >= 100 ? (kp == 6 ? 168 : kp == 4 ? 128 : 96) : kt == 11 ? 160 : kt == 6 ? 96 : 64;
            const int ctas = (int)std::min<size_t>(std::min<size_t>(227 * 1024 / (sm + 1024), 65536 / (regs * RSP_MTD_THREADS)), 3);
            const long items = (long)(kt >= 100 ? ((Q - 1) / 2 + 1 + kp - 1) / kp : (Q + kt - 1) / kt) * tg;
            const long rounds = (items + RSP_MTD_THREADS - 1) / RSP_MTD_THREADS;
            const double cost = (double)rounds * (kt >= 100 ? kp : kt) / tg / std::max(ctas, 1);
            if (best < 0 || cost < best * 0.85) { best = cost; *tg_out = tg; *kt_out = kt; }
        }
    }
    return best >= 0;
}

static cudaError_t mtd_opt_in(int P, size_t bytes) {
    switch (P) {
#define X(p, a, b, c) case p: return opt_in_smem(mtd_kernel<MtdCfg<p, a, b, c>>, bytes);
        RSP_FOR_EACH_POW2_P(X)
#undef X
    }
    return cudaErrorInvalidValue;
}

static void plan_dbf_pc(rsp_ctx* c);
static int plan_dbf_tc(rsp_ctx* c, const rsp_constants* k);
static void plan_blocking(rsp_ctx* c);

extern "C" {

int rsp_abi_version(void) { return RSP_ABI_VERSION; }

int rsp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* rsp_last_error(const rsp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

static void drop_graphs(rsp_ctx* c);
void rsp_destroy(rsp_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->prm.device);
    drop_graphs(c);
    cudaFree(c->d_raw); cudaFree(c->d_stage); cudaFree(c->d_rdm);
    for (auto& ln : c->lanes) {
        cudaFree(ln.beam); cudaFree(ln.pc); cudaFree(ln.amp); cudaFree(ln.raw); cudaFree(ln.rdm);
        if (ln.done) cudaEventDestroy(ln.done);
        if (ln.s && &ln != &c->lanes[0]) cudaStreamDestroy(ln.s);
    }
    if (c->fork) cudaEventDestroy(c->fork);
    if (c->fetch_stream) cudaStreamDestroy(c->fetch_stream);
    cudaFree(c->d_rawdet); cudaFree(c->d_fused_dbg); cudaFree(c->d_Bw_tc);
    cudaFree(c->d_tx); cudaFree(c->d_tg); cudaFree(c->d_tg_ring);
    if (c->h_tg_ring) cudaFreeHost(c->h_tg_ring);
    for (auto& sg : c->s2) { cudaFree(sg.tw1); cudaFree(sg.tw2); cudaFree(sg.H); }
    cudaFree(c->d_s2_win);
    cudaFree(c->d_aux); cudaFree(c->d_W); cudaFree(c->d_Wfrag); cudaFree(c->d_Wfrag_wa); cudaFree(c->d_fir);
    cudaFree(c->d_med_tw1); cudaFree(c->d_med_tw2); cudaFree(c->d_med_H);
    cudaFree(c->d_lng_tw1); cudaFree(c->d_lng_tw2); cudaFree(c->d_lng_H);
    for (int i = 0; i < 2; ++i) { cudaFree(c->d_lngx_tw1[i]); cudaFree(c->d_lngx_tw2[i]); cudaFree(c->d_lngx_H[i]); }
    cudaFree(c->d_dop_tw); cudaFree(c->d_dop_perm); cudaFree(c->d_win);
    cudaFree(c->d_range_axis); cudaFree(c->d_vel_axis); cudaFree(c->d_beam_angles); cudaFree(c->d_k_slopes);
    cudaFree(c->d_counts); cudaFree(c->d_recs);
    if (c->h_count) cudaFreeHost(c->h_count);
    if (c->h_recs) cudaFreeHost(c->h_recs);
    if (c->h_slot_count) cudaFreeHost(c->h_slot_count);
    if (c->h_slot_recs) cudaFreeHost(c->h_slot_recs);
    for (auto e : c->slot_done) cudaEventDestroy(e);
    for (auto& sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rsp_create(const rsp_params* p, rsp_ctx** out) {
    if (!p || !out) return fail(nullptr, RSP_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (p->abi_version != RSP_ABI_VERSION)
        return fail(nullptr, RSP_ERR_INVALID_ARG, "abi_version %d != %d", p->abi_version, RSP_ABI_VERSION);
    if (p->n_channels < 1 || p->n_channels > RSP_MAX_CHANNELS || p->n_beams < 2 || p->n_beams > RSP_MAX_BEAMS)
        return fail(nullptr, RSP_ERR_UNSUPPORTED, "need 1..%d channels and 2..%d beams", RSP_MAX_CHANNELS, RSP_MAX_BEAMS);
    if (p->n_pulses < 2 || p->n_pulses > 4096 || p->n_samples < 64)
        return fail(nullptr, RSP_ERR_UNSUPPORTED, "need 2..4096 pulses and >= 64 samples");
    const int G = p->n_gates[0] + p->n_gates[1] + p->n_gates[2];
    if (p->n_gates[0] < 0 || p->n_gates[1] < 0 || p->n_gates[2] < 0 || G < 1)
        return fail(nullptr, RSP_ERR_INVALID_ARG, "bad gate counts");
    for (int i = 0; i < 3; ++i)
        if (p->seg_start[i] < 1 || p->seg_start[i] > p->n_samples)
            return fail(nullptr, RSP_ERR_INVALID_ARG, "seg_start[%d] out of range", i);
    if (p->guard_r < 0 || p->guard_v < 0 || p->ref_r < 1 || p->ref_v < 1 || p->guard_r + p->ref_r < 2 ||
        p->guard_v + p->ref_v < 2 || p->guard_r + p->ref_r > 64)
        return fail(nullptr, RSP_ERR_UNSUPPORTED, "CFAR windows: need ref >= 1, 2 <= guard+ref (<= 64 in range)");
    if (p->max_detections < 1) return fail(nullptr, RSP_ERR_INVALID_ARG, "max_detections < 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(nullptr, RSP_ERR_NO_DEVICE, "no CUDA device visible; librsp has no CPU fallback");
    }
    if (p->device < 0 || p->device >= ndev) return fail(nullptr, RSP_ERR_INVALID_ARG, "device %d of %d", p->device, ndev);

    rsp_ctx* c = new rsp_ctx();
    c->prm = *p;
    c->C = p->n_channels; c->B = p->n_beams; c->P = p->n_pulses; c->N = p->n_samples; c->G = G;
    c->ldb = (c->N + 3) & ~3;
    c->ldg = (c->G + 3) & ~3;
#define CUC(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            fail(nullptr, RSP_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));           \
            rsp_destroy(c);                                                                         \
            return RSP_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)
    CUC(cudaSetDevice(p->device));
    // RSP_STREAM_PRIO=high: the context's own streams get the greatest priority (measurement aid, tools/overlap_probe.py)
    int prio_lo = 0, prio_hi = 0;
    CUC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (!probe_env("RSP_STREAM_PRIO_HIGH", 0)) prio_hi = prio_lo > 0 ? 0 : prio_lo;
    CUC(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
    const size_t PBG = (size_t)c->P * c->B * c->G;
    CUC(dev_alloc(&c->d_raw, (size_t)c->P * c->C * c->N));
    CUC(dev_alloc(&c->d_rdm, PBG));
    c->stages = probe_env("RSP_STAGES", 15);
    { const char* e = getenv("RSP_FUSE_SYNTH"); c->fuse_synth = !(e && atoi(e) == 0); }
    { const char* e = getenv("RSP_PC_GROUP_BAR"); c->pc_group_bar = e ? atoi(e) : 1; }
    const char* el = getenv("RSP_LANES");
    // Default: six lanes while a lane's intermediates (beam + pulse-compressed + amplitude cubes) stay under 128 MB, else three.
    // Measured (profiles/r2e_graph_lanes_ab.txt): config 1 52 k -> 68 k CPI/s and config 2 50.3 -> 49.0 us per CPI from 3 to 6
    // lanes (more kernels to fill each other's ramps and tails); config 3 and the native shape do not care.
    const size_t lane_bytes = ((size_t)c->P * c->B * (c->ldb + c->ldg)) * sizeof(float2) + PBG * sizeof(float);
    c->n_lanes = el ? std::min(8, std::max(1, atoi(el))) : (lane_bytes < ((size_t)128 << 20) ? 6 : 3);
    CUC(cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming));
    for (int i = 0; i < c->n_lanes; ++i) {
        rsp_ctx::Lane& ln = c->lanes[i];
        if (i == 0) ln.s = c->stream; else CUC(cudaStreamCreateWithPriority(&ln.s, cudaStreamNonBlocking, prio_hi));
        CUC(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        CUC(dev_alloc(&ln.beam, (size_t)c->P * c->B * c->ldb));
        CUC(dev_alloc(&ln.pc, (size_t)c->P * c->B * c->ldg));
        CUC(dev_alloc(&ln.amp, PBG));
        CUC(cudaMemset(ln.beam, 0, (size_t)c->P * c->B * c->ldb * sizeof(float2)));
        CUC(cudaMemset(ln.pc, 0, (size_t)c->P * c->B * c->ldg * sizeof(float2)));
    }
    c->cur = &c->lanes[0];
    const char* es = getenv("RSP_STREAM_SLOTS");
    c->slots = es ? std::max(1, atoi(es)) : 128;
    CUC(dev_alloc(&c->d_counts, (size_t)c->slots));
    CUC(dev_alloc(&c->d_recs, (size_t)c->slots * p->max_detections));
    CUC(cudaMemset(c->d_counts, 0, (size_t)c->slots * sizeof(int)));
    c->slot_lane.assign((size_t)c->slots, -1);
    CUC(dev_alloc(&c->d_rawdet, (size_t)c->slots * p->max_detections));
    CUC(cudaMallocHost(reinterpret_cast<void**>(&c->h_count), sizeof(int)));
    CUC(cudaMallocHost(reinterpret_cast<void**>(&c->h_recs), (size_t)p->max_detections * sizeof(rsp_detection)));
#undef CUC
    *out = c;
    return RSP_OK;
}

int rsp_set_stream(rsp_ctx* c, void* s) {
    if (!c) return RSP_ERR_INVALID_ARG;
    c->caller_epoch++;
    CU(c, cudaSetDevice(c->prm.device));
    CU(c, cudaStreamSynchronize(c->stream));
    if (s == RSP_STREAM_OWN) {
        if (!c->own_stream) {
            CU(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
            c->own_stream = true;
        }
    } else {
        if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
        c->stream = reinterpret_cast<cudaStream_t>(s);
        c->own_stream = false;
    }
    c->lanes[0].s = c->stream;
    return RSP_OK;
}

int rsp_synchronize(rsp_ctx* c) {
    if (!c) return RSP_ERR_INVALID_ARG;
    CU(c, cudaStreamSynchronize(c->stream));
    for (int l = 1; l < c->n_lanes; ++l)          // pipelined submissions run on the lanes' own streams
        if (c->lanes[l].s) CU(c, cudaStreamSynchronize(c->lanes[l].s));
    return RSP_OK;
}

int rsp_upload_constants(rsp_ctx* c, const rsp_constants* k) {
    if (!c || !k) return RSP_ERR_INVALID_ARG;
    if (!k->dbf_weights || !k->fir || !k->mf_medium || !k->mf_long || !k->mtd_win || !k->range_axis ||
        !k->velocity_axis || !k->beam_angles_deg || !k->k_slopes)
        return fail(c, RSP_ERR_INVALID_ARG, "null table in rsp_constants");
    if (k->n_fir < 1 || k->n_fir > 256) return fail(c, RSP_ERR_UNSUPPORTED, "n_fir must be 1..256");
    c->caller_epoch++;
    CU(c, cudaSetDevice(c->prm.device));
    const int C = c->C, B = c->B, P = c->P, G = c->G;
    // conj(W) laid out [c][b]
    std::vector<float2> W((size_t)C * B);
    for (int b = 0; b < B; ++b)
        for (int ch = 0; ch < C; ++ch) {
            const rsp_c128 w = k->dbf_weights[(size_t)b * C + ch];
            W[(size_t)ch * B + b] = make_float2((float)w.re, (float)-w.im);
        }
    CU(c, upload(&c->d_W, W));
    {   // tensor-core DBF (mma.sync TF32 x3) unless RSP_DBF=ffma
        const char* e = getenv("RSP_DBF");
        c->dbf_nt = c->dbf_ks = 0;
        if (!(e && std::string(e) == "ffma")) {
            c->dbf_nt = (B + 3) / 4;
            c->dbf_ks = C <= 16 ? 4 : 8;
            CU(c, upload(&c->d_Wfrag, make_dbf_fragments(reinterpret_cast<const double*>(k->dbf_weights), B, C, c->dbf_nt, c->dbf_ks)));
            CU(c, upload(&c->d_Wfrag_wa, make_dbf_fragments_wa(reinterpret_cast<const double*>(k->dbf_weights), B, C, (B + 7) / 8, c->dbf_ks)));
            c->dbf_wa = !(e && std::string(e) == "mma");
            c->dbf_tma2 = e && std::string(e) == "tma2";
            c->dbf_tma2_tiles = std::max(1, probe_env("RSP_DBF_TMA2_TILES", 8));
        }
    }
    std::vector<float> fir(k->n_fir);
    for (int i = 0; i < k->n_fir; ++i) fir[i] = (float)k->fir[i];
    CU(c, upload(&c->d_fir, fir));
    c->n_fir = k->n_fir;

    // pulse-compression block plans
    const int g1 = c->prm.n_gates[0], g2 = c->prm.n_gates[1], g3 = c->prm.n_gates[2];
    auto plan_seg = [&](PcPlan& pl, const rsp_c128* taps, int nt, int seg_start1, int gate0, int ng,
                        const char* env) -> int {
        if (ng == 0) { pl = PcPlan(); return RSP_OK; }
        if (nt < 1 || nt > 4096) return fail(c, RSP_ERR_UNSUPPORTED, "matched filter length %d not in 1..4096", nt);
        std::vector<zc> t(nt);
        for (int i = 0; i < nt; ++i) t[i] = zc(taps[i].re, taps[i].im);
        int L = choose_pc_len(nt, ng);
        const char* e = getenv(env);
        if (e && atoi(e) > 0) L = atoi(e);
        if (!make_pc_plan(pl, L, t.data(), nt, seg_start1 - 1, gate0, ng))
            return fail(c, RSP_ERR_UNSUPPORTED, "no block plan for %d taps (L=%d)", nt, L);
        return RSP_OK;
    };
    int rc = plan_seg(c->med, k->mf_medium, k->n_mf_medium, c->prm.seg_start[1], g1, g2, "RSP_PC_LEN_MEDIUM");
    if (rc) return rc;
    rc = plan_seg(c->lng, k->mf_long, k->n_mf_long, c->prm.seg_start[2], g1 + g2, g3, "RSP_PC_LEN_LONG");
    if (rc) return rc;
    c->lngx[0] = PcPlan(); c->lngx[1] = PcPlan();
    {   // mixed block lengths when they transform fewer points than the best single length (RSP_PC_MIX=0: never)
        const char* em = getenv("RSP_PC_MIX");
        const char* el = getenv("RSP_PC_LEN_LONG");
        int counts[3];
        const int nt = k->n_mf_long;
        const int mix_pts = (g3 > 0 && nt >= 1 && nt <= 1024) ? choose_pc_mix(nt, g3, counts) : 0;
        if (!(em && atoi(em) == 0) && !(el && atoi(el) > 0) && c->lng.L && mix_pts > 0 &&
            mix_pts < c->lng.nblk * c->lng.L) {
            std::vector<zc> t(nt);
            for (int i = 0; i < nt; ++i) t[i] = zc(k->mf_long[i].re, k->mf_long[i].im);
            const int Ls[3] = {4096, 2048, 1024};
            PcPlan* parts[3] = {&c->lng, &c->lngx[0], &c->lngx[1]};
            int gate = g1 + g2, left = g3, pi = 0;
            for (int i = 0; i < 3; ++i) {
                if (!counts[i]) continue;
                const int take = std::min(left, counts[i] * (Ls[i] - (nt - 1)));
                if (take <= 0) break;
                *parts[pi] = PcPlan();
                if (!make_pc_plan(*parts[pi], Ls[i], t.data(), nt, c->prm.seg_start[2] - 1, gate, take))
                    return fail(c, RSP_ERR_UNSUPPORTED, "no block plan for %d taps (L=%d)", nt, Ls[i]);
                gate += take; left -= take; ++pi;
            }
        }
    }
    for (int i = 0; i < 2; ++i)
        if (c->lngx[i].L) { CU(c, upload(&c->d_lngx_tw1[i], c->lngx[i].tw1)); CU(c, upload(&c->d_lngx_tw2[i], c->lngx[i].tw2)); CU(c, upload(&c->d_lngx_H[i], c->lngx[i].Hmid)); }
    if (c->med.L) { CU(c, upload(&c->d_med_tw1, c->med.tw1)); CU(c, upload(&c->d_med_tw2, c->med.tw2)); CU(c, upload(&c->d_med_H, c->med.Hmid)); }
    if (c->lng.L) { CU(c, upload(&c->d_lng_tw1, c->lng.tw1)); CU(c, upload(&c->d_lng_tw2, c->lng.tw2)); CU(c, upload(&c->d_lng_H, c->lng.Hmid)); }
#define X(A, B) CU(c, opt_in_smem(pc_fft_kernel<A, B>, pc_smem_pair<A, B>()));
    RSP_FOR_EACH_PC_PAIR(X)
#undef X

    // Doppler plan
    std::vector<float> win(P);
    c->pow2_doppler = make_doppler_plan(c->dop, P);
    if (c->pow2_doppler) {
        for (int p = 0; p < P; ++p) win[p] = (float)(k->mtd_win[p] * ((p & 1) ? -1.0 : 1.0));
        CU(c, upload(&c->d_dop_tw, c->dop.tw));
        CU(c, upload(&c->d_dop_perm, c->dop.iperm));
        c->mtd_tg = RSP_MTD_TG;
        c->mtd_smem = ((size_t)P * (RSP_MTD_TG + 1) + c->dop.tw.size() + 1) * sizeof(float2);
        c->mtd_smem = smem_for_occupancy(c->mtd_smem, probe_env("RSP_OCC_MTD", 0));
        CU(c, mtd_opt_in(P, c->mtd_smem));
    } else {
        std::vector<float2> tw(P);
        for (int m = 0; m < P; ++m) {
            const double ang = -2.0 * kPi * (double)m / (double)P;
            tw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        for (int p = 0; p < P; ++p) win[p] = (float)k->mtd_win[p];
        CU(c, upload(&c->d_dop_tw, tw));
        c->mtd_r = (P % 8 == 0) ? 8 : (P % 4 == 0) ? 4 : (P % 2 == 0) ? 2 : 1;
        c->mtd_tg = 0;
        if (!choose_dft_plan(P, c->mtd_r, &c->mtd_tg, &c->mtd_kt))
            return fail(c, RSP_ERR_UNSUPPORTED, "P=%d too large for the generic Doppler DFT kernel", P);
        c->mtd_smem = dft_smem_bytes(P, c->mtd_tg, c->mtd_r, c->mtd_kt);
#define X(tg, r, kt) if (c->mtd_tg == tg && c->mtd_r == r && c->mtd_kt == kt) CU(c, opt_in_smem(mtd_dft_kernel<tg, r, kt>, c->mtd_smem));
        RSP_FOR_EACH_DFT(X)
#undef X
    }
    CU(c, upload(&c->d_win, win));
    c->h_win = win;
    { const char* e = getenv("RSP_MTD"); c->mtd_mode = !e ? 2 : !strcmp(e, "tile") ? 0 : 2; }
    { const char* e = getenv("RSP_MTD_SQRT"); c->mtd_approx_sqrt = e && !strcmp(e, "approx"); }
    {   // CFAR tile height: the largest of {64,32,16} whose shared arrays let three CTAs share an SM
        const int mR = c->prm.guard_r + c->prm.ref_r;
        c->cfar_vec = (P % 4) == 0;
        // padded (conflict-free) pitches only where the tile keeps its size: at P = 128 they push the 32-gate tile from
        // 65 KB to 73 KB and the kernel gets slower (0.101 -> 0.115 ms at config 3), at P <= 64 they are a net gain
        c->cfar_pad = P <= 64;
        { const char* e = getenv("RSP_CFAR_PAD"); if (e) c->cfar_pad = atoi(e) != 0; }
        auto smem_for = [&](int tg) {
            const bool pad = c->cfar_vec && c->cfar_pad;
            const size_t r5 = (size_t)(tg + mR + c->prm.guard_r + 1) * (pad ? cfar4_pitch(P) : P);
            return c->cfar_vec ? ((size_t)(tg + 2 * mR) * (pad ? cfar4_pitch(P + 2 * RSP_CFAR_HALO) : P + 2 * RSP_CFAR_HALO) + r5) * sizeof(float)
                               : ((size_t)(tg + 2 * mR) * P + r5 + (size_t)tg * P) * sizeof(float);
        };
        c->cfar_tg = 16;
        for (int tg : {64, 32, 16})
            if (smem_for(tg) <= 72 * 1024) { c->cfar_tg = tg; break; }
        if (const char* e = getenv("RSP_CFAR_TG")) { const int v = atoi(e); if (v == 64 || v == 32 || v == 16) c->cfar_tg = v; }
        c->cfar_smem = smem_for(c->cfar_tg);
        if (c->cfar_smem > 200 * 1024) return fail(c, RSP_ERR_UNSUPPORTED, "CFAR tile does not fit shared memory");
        c->cfar_smem = smem_for_occupancy(c->cfar_smem, probe_env("RSP_OCC_CFAR", 3));
        const int rr = c->prm.ref_r, rv = c->prm.ref_v, gv = c->prm.guard_v;
        c->cfar_variant = (rr == 5 && rv == 5 && gv == 10) ? 1 : (rr == 5 && rv == 4 && gv == 2) ? 2 : 0;
#define RSP_CFAR_DISPATCH(TGV, ACTION)                                                      \
        if (!c->cfar_vec) { ACTION(cfar_kernel<TGV>) }                                          \
        else if (c->cfar_variant == 1) { ACTION((cfar4_kernel<TGV, 5, 5, 10>)) }                \
        else if (c->cfar_variant == 2) { ACTION((cfar4_kernel<TGV, 5, 4, 2>)) }                 \
        else { ACTION((cfar4_kernel<TGV, 0, 0, 0>)) }
#define RSP_CFAR5_CASE(TGV, ACTION)                                                             \
        if (c->cfar_tg == TGV) {                                                                \
            if (c->cfar_variant == 1) { ACTION((cfar5_kernel<TGV, 5, 5, 10>)) }                     \
            else { ACTION((cfar5_kernel<TGV, 5, 4, 2>)) }                                           \
        }
#define RSP_CFAR5_DISPATCH(ACTION) RSP_CFAR5_CASE(20, ACTION) RSP_CFAR5_CASE(40, ACTION) RSP_CFAR5_CASE(80, ACTION) RSP_CFAR5_CASE(120, ACTION)
#define OPTIN(K) CU(c, opt_in_smem(K, c->cfar_smem));
        // marching kernel (cfar5_kernel): compile-time windows, P / 4 a power of two; RSP_CFAR=quad keeps cfar4_kernel
        const int p4 = P / 4;
        { const char* e = getenv("RSP_CFAR"); c->cfar5 = c->cfar_vec && c->cfar_variant != 0 && p4 >= 1 && p4 <= RSP_CFAR_THREADS && !(e && !strcmp(e, "quad")); }
        if (c->cfar5) {
            const int mV = c->prm.guard_v + c->prm.ref_v;
            auto smem5 = [&](int tg) { return ((size_t)(tg + 2 * mR) * cfar5_pitch(P, mV) + 1 + (size_t)tg * cfar5_nq(P, mV)) * sizeof(float); };
            c->cfar_tg = 20;
            for (int tg : {120, 80, 40, 20})      // measured at config 2: 120 > 80 > 40 (profiles/r2d_cfar5_ab.txt)
                if (smem5(tg) <= 72 * 1024) { c->cfar_tg = tg; break; }
            if (const char* e = getenv("RSP_CFAR5_TG")) { const int v = atoi(e); if (v == 20 || v == 40 || v == 80 || v == 120) c->cfar_tg = v; }
            c->cfar_smem = smem5(c->cfar_tg);
            if (c->cfar_smem > 200 * 1024) return fail(c, RSP_ERR_UNSUPPORTED, "CFAR tile does not fit shared memory");
            RSP_CFAR5_DISPATCH(OPTIN)
        }
        else if (c->cfar_tg == 64) { RSP_CFAR_DISPATCH(64, OPTIN) }
        else if (c->cfar_tg == 32) { RSP_CFAR_DISPATCH(32, OPTIN) }
        else { RSP_CFAR_DISPATCH(16, OPTIN) }
#undef OPTIN
    }

    CU(c, upload(&c->d_range_axis, std::vector<double>(k->range_axis, k->range_axis + G)));
    CU(c, upload(&c->d_vel_axis, std::vector<double>(k->velocity_axis, k->velocity_axis + P)));
    CU(c, upload(&c->d_beam_angles, std::vector<double>(k->beam_angles_deg, k->beam_angles_deg + B)));
    CU(c, upload(&c->d_k_slopes, std::vector<double>(k->k_slopes, k->k_slopes + (B - 1))));
    c->delta_r = k->delta_r;
    c->delta_v = k->delta_v;
    plan_dbf_pc(c);
    if (int rc_tc = plan_dbf_tc(c, k)) return rc_tc;
    plan_blocking(c);
    c->have_constants = true;
    return RSP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// launch sequence for one CPI (device-resident PCN complex64 input)
// ------------------------------------------------------------------------------------------
static DiscardArgs dead_buf(const rsp_ctx* c, void* p, size_t bytes) {
    DiscardArgs d;
    d.ptr = (c->discard && (reinterpret_cast<uintptr_t>(p) & 127) == 0) ? p : nullptr;
    d.bytes = bytes;
    return d;
}
static DiscardArgs dead_amp(const rsp_ctx* c) { return dead_buf(c, c->cur->amp, (size_t)c->P * c->B * c->G * sizeof(float)); }
static DiscardArgs dead_beam(const rsp_ctx* c) { return dead_buf(c, c->cur->beam, (size_t)c->P * c->B * c->ldb * sizeof(float2)); }
static DiscardArgs dead_pc(const rsp_ctx* c) { return dead_buf(c, c->cur->pc, (size_t)c->P * c->B * c->ldg * sizeof(float2)); }


// ------------------------------------------------------------------------------------------
// dbf_pc_kernel (rsp_fused.cuh): round schedule, shared-memory layout, tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// Tensor map of one raw cube raw[p][c][n] (complex64 = 2 floats).  rank4: {32 floats of a 16-sample group, C channels,
// N/16 groups, P pulses} with box {32, C, 8, 1}: ONE copy lands a [8 groups][C][16 samples] tile, the layout in which the
// MMA fragment loads (4 channels x 8 sample pairs per warp instruction) are 512 contiguous bytes.  rank2 (fallback should a
// driver reject the non-monotonic strides): {2 N floats, C P rows} with box {32, C}, eight copies per tile.
enum { TMAP_GROUPS_2D = 0, TMAP_GROUPS_4D = 1, TMAP_ROWS_2D = 2 };   // TMAP_ROWS_2D: box {256 floats, C rows} = [C][128 samples] (dbf_tc_kernel)
static bool encode_raw_tmap(const rsp_ctx* c, const void* raw, int mode, CUtensorMap* out) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t N = (cuuint64_t)c->N, C = (cuuint64_t)c->C, P = (cuuint64_t)c->P;
    if (mode == TMAP_GROUPS_4D) {
        const cuuint64_t dims[4] = {32, C, N / 16, P};
        const cuuint64_t strides[3] = {N * 8, 128, C * N * 8};
        const cuuint32_t box[4] = {32, (cuuint32_t)C, RSP_FUSED_TILE / 16, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(raw), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    const cuuint64_t dims[2] = {2 * N, C * P};
    const cuuint64_t strides[1] = {N * 8};
    const cuuint32_t box[2] = {mode == TMAP_ROWS_2D ? 256u : 32u, (cuuint32_t)C};
    const cuuint32_t es[2] = {1, 1};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(raw), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int pc_cfg_id(int L) { return L == 4096 ? 3 : L == 2048 ? 2 : L == 1024 ? 1 : 0; }
static int pc_smem_elems(int L) { return L + L / 16 + 16; }

// Decide whether the fused kernel covers this context and lay out its rounds: every overlap-save block of one line
// (long segment, its mixed-length continuation, medium segment) is a group of L/512 warps; first-fit decreasing into
// rounds of 8 warps.  Called at the end of rsp_upload_constants.
static void plan_dbf_pc(rsp_ctx* c) {
    c->dbf_pc_ok = false;
    c->tmaps.clear();
    if (encode_tiled_fn() && c->N % 16 == 0) {   // tensor-map form: probe the single-copy rank-4 map once (any 16-byte aligned address will do)
        const char* e = getenv("RSP_FUSED_TMA");
        CUtensorMap probe;
        c->dbf_pc_tma_rank4 = !(e && !strcmp(e, "2d")) && encode_raw_tmap(c, c->d_raw, TMAP_GROUPS_4D, &probe);
    }
    { const char* e = getenv("RSP_FUSE_DBF_PC"); c->dbf_pc_enabled = e && atoi(e) != 0; }   // opt-in: measured slower than dbf_tc + pc_fft
    if (!c->dbf_pc_enabled || !c->dbf_nt || !c->dbf_wa || c->dbf_tma2) return;
    if (c->B > 8 || c->C > 16 || c->N % 16 != 0 || c->N > 8192 || (c->ldb & 1) || !encode_tiled_fn()) return;
    const bool narrow = c->prm.n_gates[0] > 0;
    if (narrow && !c->med.L) return;                   // the narrow FIR rides on the medium group
    if (!c->med.L && !c->lng.L) return;
    DbfPcArgs& a = c->dbf_pc;
    a = DbfPcArgs{};
    const PcPlan* pl[4] = {&c->lng, &c->lngx[0], &c->lngx[1], &c->med};
    const float2* tw1[4] = {c->d_lng_tw1, c->d_lngx_tw1[0], c->d_lngx_tw1[1], c->d_med_tw1};
    const float2* tw2[4] = {c->d_lng_tw2, c->d_lngx_tw2[0], c->d_lngx_tw2[1], c->d_med_tw2};
    const float2* H[4] = {c->d_lng_H, c->d_lngx_H[0], c->d_lngx_H[1], c->d_med_H};
    struct Item { int L, seg, blk; };
    std::vector<Item> items;
    int tab = 0;
    for (int i = 0; i < 4; ++i) {
        a.tw2_off[i] = tab;
        a.tw2_len[i] = 0;
        std::memset(&a.seg[i], 0, sizeof a.seg[i]);
        if (!pl[i]->L) continue;
        if (!pc_cfg_id(pl[i]->L)) return;
        // fill_seg without the CTA bookkeeping
        PcSegArgs& sg = a.seg[i];
        sg.tw1 = tw1[i]; sg.tw2 = tw2[i]; sg.Hmid = H[i];
        sg.seg_start0 = pl[i]->seg_start0; sg.in_lo = pl[i]->seg_start0; sg.in_hi = c->N; sg.taps = pl[i]->taps;
        sg.gate0 = pl[i]->gate0; sg.g_end = pl[i]->gate0 + pl[i]->ngates; sg.valid = pl[i]->valid; sg.nblk = pl[i]->nblk;
        a.tw2_len[i] = (int)pl[i]->tw2.size();
        tab += a.tw2_len[i];
        for (int b = 0; b < pl[i]->nblk; ++b) items.push_back({pl[i]->L, i, b});
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.L > y.L; });
    int used[RSP_FUSED_MAX_ROUNDS] = {0, 0, 0, 0}, ngrp[RSP_FUSED_MAX_ROUNDS] = {0, 0, 0, 0}, work[RSP_FUSED_MAX_ROUNDS] = {0, 0, 0, 0};
    std::memset(a.warp_group, -1, sizeof a.warp_group);
    a.n_rounds = 0;
    bool narrow_placed = !narrow;
    for (const Item& it : items) {
        const int nw = it.L / 512;
        int r = 0;
        while (r < RSP_FUSED_MAX_ROUNDS && (used[r] + nw > RSP_FUSED_WARPS || ngrp[r] >= RSP_FUSED_MAX_GROUPS)) ++r;
        if (r == RSP_FUSED_MAX_ROUNDS) return;         // too many blocks per line for this kernel
        DbfPcGroup& g = a.grp[r][ngrp[r]];
        g.cfg = pc_cfg_id(it.L); g.seg = it.seg; g.blk = it.blk; g.warp0 = used[r]; g.work_off = work[r];
        g.narrow = (!narrow_placed && it.seg == 3 && it.blk == 0) ? 1 : 0;
        if (g.narrow) narrow_placed = true;
        for (int w = 0; w < nw; ++w) a.warp_group[r][used[r] + w] = (signed char)ngrp[r];
        used[r] += nw; work[r] += pc_smem_elems(it.L); ++ngrp[r];
        a.n_rounds = std::max(a.n_rounds, r + 1);
    }
    if (!narrow_placed) return;
    int work_elems = 0;
    for (int r = 0; r < a.n_rounds; ++r) work_elems = std::max(work_elems, work[r]);
    const size_t ring = (size_t)RSP_FUSED_STAGES * (RSP_FUSED_TILE / 16) * c->C * 128;
    a.line_bytes = (int)(((size_t)c->N * 8 + 127) & ~(size_t)127);
    a.work_bytes = (int)((std::max((size_t)work_elems * sizeof(float2), ring) + 127) & ~(size_t)127);
    c->dbf_pc_smem = (size_t)a.line_bytes + a.work_bytes + (size_t)tab * sizeof(float2) + 256 * sizeof(float) + 64;
    if (c->dbf_pc_smem > 227 * 1024) return;
    a.Wa = c->d_Wfrag_wa;
    a.C = c->C; a.B = c->B; a.P = c->P; a.N = c->N; a.ldb = c->ldb; a.ldg = c->ldg;
    a.tiles = (c->N + RSP_FUSED_TILE - 1) / RSP_FUSED_TILE;
    a.fir = c->d_fir; a.nfir = c->n_fir; a.fir_delay = c->prm.fir_delay;
    a.narrow_start0 = c->prm.seg_start[0] - 1; a.narrow_gates = c->prm.n_gates[0];
    { CUtensorMap probe; if (!c->dbf_pc_tma_rank4 && !encode_raw_tmap(c, c->d_raw, TMAP_GROUPS_2D, &probe)) return; }
    a.tma_rank4 = c->dbf_pc_tma_rank4;
    if (cudaFuncSetAttribute(dbf_pc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->dbf_pc_smem) != cudaSuccess) { cudaGetLastError(); return; }
    cudaFuncSetAttribute(dbf_pc_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaGetLastError();
    {   // clusters that can be resident at once = pulses per wave; a CTA prefetches the slice of the pulse one wave ahead
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(c->P * c->B)); cfg.blockDim = dim3(RSP_FUSED_THREADS); cfg.dynamicSmemBytes = c->dbf_pc_smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)c->B; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, dbf_pc_kernel<4>, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
        const char* e = getenv("RSP_FUSED_PREFETCH");
        a.prefetch_ahead = e ? atoi(e) : nc;
        c->dbf_pc_clusters = nc;
    }
    if (const char* e = getenv("RSP_FUSED_DEBUG")) {   // measurement aid (tools/fused_diag.py): per-CTA phase timestamps
        c->fused_dbg_flags = probe_env("RSP_FUSED_DEBUG", 0);
        (void)e;
        if (!c->d_fused_dbg && cudaMalloc(reinterpret_cast<void**>(&c->d_fused_dbg), (size_t)c->P * c->B * 8 * sizeof(long long)) != cudaSuccess) { cudaGetLastError(); c->d_fused_dbg = nullptr; }
    }
    c->dbf_pc_ok = true;
}

static const CUtensorMap* raw_tmap(rsp_ctx* c, const float2* raw, int mode = -1);
static int launch_dbf_pc(rsp_ctx* c, const float2* raw, int* det_count) {
    const CUtensorMap* map = raw_tmap(c, raw);
    if (!map) return fail(c, RSP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the raw cube at %p", raw);
    DbfPcArgs a = c->dbf_pc;
    a.pc = c->cur->pc;
    a.beam_out = c->keep_beam ? c->cur->beam : nullptr;
    a.det_count = det_count;
    a.dead = dead_amp(c);
    a.dbg = c->d_fused_dbg;
    a.dbg_flags = c->fused_dbg_flags > 0 ? c->fused_dbg_flags : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(c->P * c->B));
    cfg.blockDim = dim3(RSP_FUSED_THREADS);
    cfg.dynamicSmemBytes = c->dbf_pc_smem;
    cfg.stream = c->cur->s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)c->B; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    Timed t(c, K_DBF_PC);
    CU(c, cudaLaunchKernelEx(&cfg, dbf_pc_kernel<4>, *map, a));
    return RSP_OK;
}

static const CUtensorMap* raw_tmap(rsp_ctx* c, const float2* raw, int mode) {
    if (mode < 0) mode = c->dbf_pc_tma_rank4 ? TMAP_GROUPS_4D : TMAP_GROUPS_2D;
    for (auto& e : c->tmaps) if (e.ptr == raw && e.mode == mode) return &e.map;
    if (c->tmaps.size() >= 64) c->tmaps.erase(c->tmaps.begin());
    rsp_ctx::TmapEntry e;
    e.ptr = raw;
    e.mode = mode;
    if ((reinterpret_cast<uintptr_t>(raw) & 15) || !encode_raw_tmap(c, raw, mode, &e.map)) return nullptr;
    c->tmaps.push_back(e);
    return &c->tmaps.back().map;
}

static int launch_dbf_tma2(rsp_ctx* c, const float2* raw, int* det_count) {
    const CUtensorMap* map = raw_tmap(c, raw);
    if (!map) return fail(c, RSP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the raw cube at %p", raw);
    DbfTmaArgs a;
    a.beam = c->cur->beam; a.Wa = c->d_Wfrag_wa; a.det_count = det_count;
    a.C = c->C; a.B = c->B; a.N = c->N; a.ldb = c->ldb;
    a.tiles = (c->N + RSP_FUSED_TILE - 1) / RSP_FUSED_TILE;
    a.tiles_per_cta = c->dbf_tma2_tiles;
    a.tma_rank4 = c->dbf_pc_tma_rank4;
    a.dead = dead_amp(c);
    const size_t sm = (size_t)RSP_DBFT_STAGES * (RSP_FUSED_TILE / 16) * c->C * 128;
    static bool once = false;
    if (!once) { cudaFuncSetAttribute(dbf_tma2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); once = true; }
    dim3 grid((a.tiles + a.tiles_per_cta - 1) / a.tiles_per_cta, c->P);
    Timed t(c, K_DBF);
    dbf_tma2_kernel<4><<<grid, RSP_FUSED_THREADS, sm, c->cur->s>>>(*map, a);
    return RSP_OK;
}


// ------------------------------------------------------------------------------------------
// dbf_tc_kernel (rsp_dbf_tc.cuh): weight operands, shared-memory budget, launch
// ------------------------------------------------------------------------------------------
static std::vector<float> make_dbf_tc_weights(const double* W_ri /* [B][C][2] */, int B, int C, int Npad, int Cpad) {
    const int K = 2 * Cpad;
    std::vector<float> out((size_t)2 * Npad * K, 0.f);
    for (int n = 0; n < Npad; ++n)
        for (int kk = 0; kk < K; ++kk) {
            const int b = n >> 1, ro = n & 1, ch = kk >> 1, ri = kk & 1;
            float v = 0.f;
            if (b < B && ch < C) {
                const float wr = (float)W_ri[((size_t)b * C + ch) * 2], wi = (float)W_ri[((size_t)b * C + ch) * 2 + 1];
                v = ro == 0 ? (ri == 0 ? wr : wi) : (ri == 0 ? -wi : wr);      // x * conj(W), fsf:95
            }
            const size_t off = ((size_t)(kk / 4) * (Npad / 8) * 128 + (size_t)(n / 8) * 128 + (size_t)(n % 8) * 16 + (size_t)(kk % 4) * 4) / 4;
            const float hi = host_tf32(v);
            out[off] = hi;
            out[(size_t)Npad * K + off] = host_tf32(v - hi);
        }
    return out;
}

static int plan_dbf_tc(rsp_ctx* c, const rsp_constants* k) {
    c->dbf_tc = false;
    c->pulse_block = 0;
    const char* e = getenv("RSP_DBF");
    if (e && strcmp(e, "tc")) return RSP_OK;                      // another DBF kernel was asked for
    if (c->B > 16 || c->C > 32 || (c->N & 1) || !encode_tiled_fn()) return RSP_OK;
    { CUtensorMap probe; if (!encode_raw_tmap(c, c->d_raw, TMAP_ROWS_2D, &probe)) return RSP_OK; }
    DbfTcArgs& a = c->dbf_tc_args;
    a = DbfTcArgs{};
    a.C = c->C; a.B = c->B; a.P = c->P; a.N = c->N; a.ldb = c->ldb;
    a.Cpad = c->C <= 8 ? 8 : c->C <= 16 ? 16 : 32;
    a.Npad = 2 * c->B <= 16 ? 16 : 32;
    a.tiles_per_pulse = (c->N + RSP_TC_TILE - 1) / RSP_TC_TILE;
    const size_t raw_stage = (size_t)c->C * 1024, b_bytes = (size_t)a.Npad * 2 * a.Cpad * 4;
    const size_t fixed = 2 * b_bytes + (2 * RSP_TC_MAX_STAGES + 6) * 8 + 16;
    // Three stages and, for the small shapes, CTAs of 16 tiles: measured best inside the chain (a deep ring or one long-lived
    // CTA per SM is faster alone but keeps the other lanes' kernels off the SM; profiles/r2_tc_probe.txt)
    int ns = 3;
    if (const char* es = getenv("RSP_TC_STAGES")) ns = std::max(2, std::min(RSP_TC_MAX_STAGES, atoi(es)));
    a.chunk = c->C <= 16 ? 16 : 0;
    while (ns > 2 && fixed + ns * raw_stage > 220 * 1024) --ns;
    if (fixed + ns * raw_stage > 227 * 1024) return RSP_OK;
    a.stages = ns;
    a.dbg = probe_env("RSP_TC_DEBUG", 0);
    c->dbf_tc_smem = fixed + ns * raw_stage;
    CU(c, upload(&c->d_Bw_tc, make_dbf_tc_weights(reinterpret_cast<const double*>(k->dbf_weights), c->B, c->C, a.Npad, a.Cpad)));
    a.Bw = c->d_Bw_tc;
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->prm.device);
    int per_sm = 1;
    per_sm = std::max(1, probe_env("RSP_TC_CTAS_PER_SM", 1));
    c->dbf_tc_grid = std::min(nsm * per_sm, c->P * a.tiles_per_pulse);
    if (const char* ec = getenv("RSP_TC_CHUNK")) a.chunk = std::max(0, atoi(ec));
    if (a.chunk > 0) c->dbf_tc_grid = (c->P * a.tiles_per_pulse + a.chunk - 1) / a.chunk;
    cudaError_t err = cudaErrorInvalidValue;
#define RSP_TC_CASE(NP, CP) if (a.Npad == NP && a.Cpad == CP) { prefer_max_smem(dbf_tc_kernel<NP, CP>); err = cudaFuncSetAttribute(dbf_tc_kernel<NP, CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->dbf_tc_smem); }
    RSP_TC_CASE(16, 8) RSP_TC_CASE(16, 16) RSP_TC_CASE(16, 32) RSP_TC_CASE(32, 8) RSP_TC_CASE(32, 16) RSP_TC_CASE(32, 32)
#undef RSP_TC_CASE
    if (err != cudaSuccess) { cudaGetLastError(); return RSP_OK; }
    c->dbf_tc = true;
    // RSP_PULSE_BLOCK=n: opt-in pulse-blocked S5 -> S6 (enqueue_chain).  Measured slower at config 3 (470 us per CPI whole-cube,
    // 493 with groups of 16 pulses, 511 with 8: profiles/r2f_pulse_block_cfg3.txt), like the range-blocked chain.
    if (const char* e = getenv("RSP_PULSE_BLOCK")) c->pulse_block = std::max(0, atoi(e));
    return RSP_OK;
}

static int launch_dbf_tc(rsp_ctx* c, const float2* raw, int* det_count, int tile_lo = 0, int tile_hi = -1, int p_lo = 0, int p_hi = -1,
                         DiscardArgs dead_group = DiscardArgs{nullptr, 0}) {
    const CUtensorMap* map = raw_tmap(c, raw, TMAP_ROWS_2D);
    if (!map) return fail(c, RSP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the raw cube at %p", raw);
    DbfTcArgs a = c->dbf_tc_args;
    a.beam = c->cur->beam;
    a.det_count = det_count;
    a.dead = dead_amp(c);
    a.tile_lo = tile_lo; a.tile_hi = tile_hi < 0 ? a.tiles_per_pulse : tile_hi;
    a.p_lo = p_lo; a.p_hi = p_hi < 0 ? c->P : p_hi;
    int grid = c->dbf_tc_grid;
    if (tile_hi >= 0 || p_hi >= 0) {           // a chunk of the range-blocked path / a pulse group of the pulse-blocked path
        const int n_tiles = (a.p_hi - a.p_lo) * (a.tile_hi - a.tile_lo);
        if (!det_count) a.dead = dead_group;
        grid = a.chunk > 0 ? (n_tiles + a.chunk - 1) / a.chunk : std::min(grid, n_tiles);
    }
#ifdef RSP_PROBES
    // Experiment RSP_EXP_DBF_MULTI=K (tools/overlap_probe.py): ONE persistent launch forms the beams of K consecutive cubes of
    // the input pool (the launches of the other K - 1 CPIs are skipped), writing into a two-deep scratch ring, so that the
    // DBF CTAs stay resident while the kernels of another context come and go: do they share the SMs?
    static const int exp_multi = probe_env("RSP_EXP_DBF_MULTI", 0);
    static CUtensorMap multi_map;
    if (exp_multi > 1 && tile_hi < 0) {
        if ((c->exp_multi_count++ % exp_multi) != 0) return RSP_OK;
        const cuuint64_t dims[2] = {2 * (cuuint64_t)c->N, (cuuint64_t)c->C * c->P * exp_multi};
        const cuuint64_t strides[1] = {(cuuint64_t)c->N * 8};
        const cuuint32_t box[2] = {256u, (cuuint32_t)c->C};
        const cuuint32_t es[2] = {1, 1};
        if (encode_tiled_fn()(&multi_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(raw), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fail(c, RSP_ERR_CUDA, "multi-CPI tensor map");
        map = &multi_map;
        if (!c->exp_beam) cudaMalloc(reinterpret_cast<void**>(&c->exp_beam), (size_t)2 * c->P * c->B * c->ldb * sizeof(float2));
        a.beam = c->exp_beam;
        a.p_hi = c->P * exp_multi;
        a.p_wrap = 2 * c->P;
        a.chunk = 0;
        a.dead = DiscardArgs{nullptr, 0};
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->prm.device);
        grid = nsm;
    }
#endif
    Timed t(c, K_DBF);
#define RSP_TC_CASE(NP, CP) if (a.Npad == NP && a.Cpad == CP) dbf_tc_kernel<NP, CP><<<grid, RSP_TC_THREADS, c->dbf_tc_smem, c->cur->s>>>(*map, a);
    RSP_TC_CASE(16, 8) RSP_TC_CASE(16, 16) RSP_TC_CASE(16, 32) RSP_TC_CASE(32, 8) RSP_TC_CASE(32, 16) RSP_TC_CASE(32, 32)
#undef RSP_TC_CASE
    return RSP_OK;
}

template <int NB> static void launch_dbf(rsp_ctx* c, const float2* raw, int* det_count) {
    constexpr int SPT = 2, CU_ = 4;
    Timed t(c, K_DBF);
    dim3 grid((c->N + RSP_DBF_THREADS * SPT - 1) / (RSP_DBF_THREADS * SPT), c->P);
    dbf_kernel<NB, SPT, CU_><<<grid, RSP_DBF_THREADS, 0, c->cur->s>>>(raw, c->cur->beam, c->d_W, c->C, c->N, c->ldb, det_count, dead_amp(c));
}

template <int NT, int KS> static void launch_dbf_mma(rsp_ctx* c, const float2* raw, int* det_count) {
    Timed t(c, K_DBF);
    const int per_cta = (RSP_DBF_MMA_THREADS / 32) * 32;
    dim3 grid((c->N + per_cta - 1) / per_cta, c->P);
    const bool vec = (c->N % 2 == 0) && ((reinterpret_cast<uintptr_t>(raw) & 15) == 0);
    if (vec && c->dbf_wa) {
        constexpr int MT = (NT + 1) / 2;       // NT = ceil(B / 4) n-tiles of the old kernel  ->  ceil(B / 8) m-tiles
        constexpr int NQ = (KS <= 4 && MT == 1) ? 2 : 1;  // 16-sample warp tiles for the big shapes (registers)
        dim3 gg((c->N + per_cta * NQ / 2 - 1) / (per_cta * NQ / 2), c->P);
        dbf_mma2_kernel<MT, KS, NQ><<<gg, RSP_DBF_MMA_THREADS, 0, c->cur->s>>>(raw, c->cur->beam, c->d_Wfrag_wa, c->C, c->B, c->N, c->ldb, det_count, dead_amp(c));
        return;
    }
    if (vec) dbf_mma_kernel<NT, KS, true><<<grid, RSP_DBF_MMA_THREADS, 0, c->cur->s>>>(raw, c->cur->beam, c->d_Wfrag, c->C, c->B, c->N, c->ldb, det_count, dead_amp(c));
    else dbf_mma_kernel<NT, KS, false><<<grid, RSP_DBF_MMA_THREADS, 0, c->cur->s>>>(raw, c->cur->beam, c->d_Wfrag, c->C, c->B, c->N, c->ldb, det_count, dead_amp(c));
}

template <int MT, int KS> static void launch_dbf_synth(rsp_ctx* c, int* det_count) {
    constexpr int NQ = (KS <= 4 && MT == 1) ? 2 : 1;
    Timed t(c, K_DBF);
    const int per_cta = (RSP_DBF_MMA_THREADS / 32) * 16 * NQ;
    dim3 grid((c->N + per_cta - 1) / per_cta, c->P);
    dbf_synth_kernel<MT, KS, NQ><<<grid, RSP_DBF_MMA_THREADS, 0, c->cur->s>>>(*c->fused, c->cur->beam, c->d_Wfrag_wa, c->B, c->ldb, det_count, dead_amp(c));
}

static bool fused_synth_possible(const rsp_ctx* c, int n_tg) {
    return c->fuse_synth && c->dbf_nt && c->dbf_wa && (c->N % 2 == 0) && n_tg <= RSP_SYNTH_FUSED_T &&
           c->C <= 4 * c->dbf_ks;
}

static int launch_dbf_any(rsp_ctx* c, const float2* raw, int* det_count) {
    if (c->fused) {                        // S4 + S5 in one kernel (rsp_submit_targets)
        const int key = ((c->B + 7) / 8) * 10 + c->dbf_ks;
        switch (key) {
#define CASE(mt, ks) case mt * 10 + ks: launch_dbf_synth<mt, ks>(c, det_count); return RSP_OK;
            CASE(1, 4) CASE(2, 4) CASE(1, 8) CASE(2, 8)
#undef CASE
        }
        return fail(c, RSP_ERR_UNSUPPORTED, "no fused synthesis kernel for %d beams / %d channels", c->B, c->C);
    }
    // A context that synthesises its own echoes (rsp_set_waveform) takes the mma.sync DBF on every path: the fused S4 + S5
    // kernel of the pipelined frame path is built on that arithmetic, so a frame gives the same bits whether it goes
    // through rsp_process_targets, rsp_submit_targets or rsp_synthesize + rsp_process_cpi.
    const bool frame_mma = c->have_waveform && c->dbf_nt && c->dbf_wa && (c->N % 2 == 0);
    if (c->dbf_tc && !frame_mma && !(reinterpret_cast<uintptr_t>(raw) & 15)) return launch_dbf_tc(c, raw, det_count);
    if (c->dbf_tma2 && c->dbf_nt && c->B <= 8 && c->C <= 16 && c->N % 16 == 0 && !(c->ldb & 1) && encode_tiled_fn())
        return launch_dbf_tma2(c, raw, det_count);
    if (c->dbf_nt) {
        const int key = c->dbf_nt * 10 + c->dbf_ks;
        switch (key) {
#define CASE(nt, ks) case nt * 10 + ks: launch_dbf_mma<nt, ks>(c, raw, det_count); return RSP_OK;
            CASE(1, 4) CASE(2, 4) CASE(3, 4) CASE(4, 4) CASE(1, 8) CASE(2, 8) CASE(3, 8) CASE(4, 8)
#undef CASE
        }
    }
    switch (c->B) {
#define CASE(n) case n: launch_dbf<n>(c, raw, det_count); break;
        CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
        CASE(15) CASE(16)
#undef CASE
        default: return fail(c, RSP_ERR_UNSUPPORTED, "n_beams %d", c->B);
    }
    return RSP_OK;
}

static void fill_seg(const rsp_ctx* c, PcSegArgs& sg, const PcPlan& pl, const float2* tw1, const float2* tw2, const float2* H, int n_lines = -1) {
    sg.tw1 = tw1; sg.tw2 = tw2; sg.Hmid = H;
    sg.seg_start0 = pl.seg_start0; sg.in_lo = pl.seg_start0; sg.in_hi = c->N; sg.taps = pl.taps; sg.gate0 = pl.gate0; sg.g_end = pl.gate0 + pl.ngates; sg.valid = pl.valid;
    sg.nblk = pl.nblk;
    sg.n_items = pl.L ? (n_lines < 0 ? c->P * c->B : n_lines) * pl.nblk : 0;
    const int ng = pl.L ? RSP_PC_THREADS / pl.T : 1;
    sg.n_ctas = (sg.n_items + ng - 1) / ng;
}

// One launch covers the long segment (role 0), the medium segment and the narrow FIR (role 1).
// p_hi >= 0: only the lines of pulses [p_lo, p_hi) (pulse-blocked path; lines are [pulse][beam], so a pulse group is a
// contiguous run of lines of both cubes).
static void launch_pc(rsp_ctx* c, int p_lo = 0, int p_hi = -1) {
    const bool narrow = c->prm.n_gates[0] > 0;
    const bool fold = narrow && c->med.L > 0;        // the medium groups compute the narrow gates too
    const int line0 = p_lo * c->B, n_lines = ((p_hi < 0 ? c->P : p_hi) - p_lo) * c->B;
    const float2* beam = c->cur->beam + (size_t)line0 * c->ldb;
    float2* pc = c->cur->pc + (size_t)line0 * c->ldg;
    if (narrow && !fold) {
        Timed t(c, K_PC_NARROW);
        pc_narrow_kernel<<<n_lines, 256, 0, c->cur->s>>>(beam, pc, c->d_fir, c->n_fir, c->prm.fir_delay, c->N,
                                                         c->ldb, c->ldg, c->prm.seg_start[0] - 1, c->prm.n_gates[0]);
    }
    if (!c->med.L && !c->lng.L) return;
    PcKernelArgs a;
    a.beam = beam; a.pc = pc; a.N = c->N; a.ldb = c->ldb; a.ldg = c->ldg;
    fill_seg(c, a.seg[0], c->lng, c->d_lng_tw1, c->d_lng_tw2, c->d_lng_H, n_lines);
    fill_seg(c, a.seg[1], c->med, c->d_med_tw1, c->d_med_tw2, c->d_med_H, n_lines);
    a.do_narrow = fold ? 1 : 0;
    a.group_bar = c->pc_group_bar;
    a.fir = c->d_fir; a.nfir = c->n_fir; a.fir_delay = c->prm.fir_delay;
    a.narrow_start0 = c->prm.seg_start[0] - 1; a.narrow_gates = c->prm.n_gates[0];
    const int nctas = a.seg[0].n_ctas + a.seg[1].n_ctas;
    const int la = c->lng.L ? c->lng.L : 1024, lb = c->med.L ? c->med.L : 1024;
    Timed t(c, K_PC);
#ifdef RSP_PROBES
    static const int exp_merge = probe_env("RSP_EXP_MERGE", 0);
    if (exp_merge && la == 4096 && lb == 1024 && c->exp_raw && c->B <= 8 && c->C <= 16) {   // experiment: DBF CTAs interleaved with the PC CTAs
        if (!c->exp_beam) cudaMalloc(reinterpret_cast<void**>(&c->exp_beam), (size_t)c->P * c->B * c->ldb * sizeof(float2));
        MergeDbfArgs d;
        d.raw = c->exp_raw; d.beam = c->exp_beam; d.Wa = c->d_Wfrag_wa; d.C = c->C; d.NB = c->B; d.N = c->N; d.ldb = c->ldb; d.P = c->P;
        d.n_dbf = c->P * ((c->N + 255) / 256) / (exp_merge > 1 ? exp_merge : 1);
        d.n_total = nctas + d.n_dbf;
        static bool once = false;
        if (!once) { opt_in_smem(pc_dbf_merge_kernel<Pc4096, Pc1024>, pc_smem_pair<Pc4096, Pc1024>()); once = true; }
        pc_dbf_merge_kernel<Pc4096, Pc1024><<<d.n_total, RSP_PC_THREADS, pc_smem_pair<Pc4096, Pc1024>(), c->cur->s>>>(a, d);
    } else
#endif
    {
#define X(A, B) if (la == A::L && lb == B::L) pc_fft_kernel<A, B><<<nctas, RSP_PC_THREADS, pc_smem_pair<A, B>(), c->cur->s>>>(a);
    RSP_FOR_EACH_PC_PAIR(X)
#undef X
    }
    if (c->lngx[0].L) {            // the shorter blocks of a mixed long-segment plan
        fill_seg(c, a.seg[0], c->lngx[0], c->d_lngx_tw1[0], c->d_lngx_tw2[0], c->d_lngx_H[0], n_lines);
        fill_seg(c, a.seg[1], c->lngx[1], c->d_lngx_tw1[1], c->d_lngx_tw2[1], c->d_lngx_H[1], n_lines);
        a.do_narrow = 0;
        const int n2 = a.seg[0].n_ctas + a.seg[1].n_ctas;
        const int l0 = c->lngx[0].L, l1 = c->lngx[1].L ? c->lngx[1].L : 1024;
        c->launches++;
#define X(A, B) if (l0 == A::L && l1 == B::L) pc_fft_kernel<A, B><<<n2, RSP_PC_THREADS, pc_smem_pair<A, B>(), c->cur->s>>>(a);
        RSP_FOR_EACH_PC_PAIR(X)
#undef X
    }
}

static void launch_mtd(rsp_ctx* c, float2* rdm, int g_lo = 0, int g_hi = -1) {
    MtdArgs a;
    a.pc = c->cur->pc; a.rdm = rdm; a.amp = c->cur->amp; a.win = c->d_win; a.tw = c->d_dop_tw; a.perm = c->d_dop_perm;
    a.P = c->P; a.B = c->B; a.G = c->G; a.ldg = c->ldg;
    a.p_magic = (unsigned)((1ull << 32) / (unsigned)c->P + 1ull);
    a.dead = g_hi < 0 ? dead_beam(c) : DiscardArgs{nullptr, 0};
    a.g_lo = g_lo; a.g_hi = g_hi < 0 ? c->G : g_hi;
    const int ng = a.g_hi - a.g_lo;
    const int tg = c->mtd_tg;
    dim3 grid((ng + tg - 1) / tg, c->B);
    Timed t(c, K_MTD);
    if (c->pow2_doppler && c->mtd_mode == 2 && c->P == 64 && (int)c->h_win.size() == c->P) {
        MtdRegArgs r;
        r.m = a;
        for (int p = 0; p < 64; ++p) r.win[p] = c->h_win[p];
        dim3 sgrid((ng + 31) / 32, c->B);
        { static const bool once = (prefer_max_smem(mtd64_kernel<true>), prefer_max_smem(mtd64_kernel<false>), true); (void)once; }
        if (c->mtd_approx_sqrt) mtd64_kernel<true><<<sgrid, 256, 0, c->cur->s>>>(r);
        else mtd64_kernel<false><<<sgrid, 256, 0, c->cur->s>>>(r);
        return;
    }
    if (c->pow2_doppler) {
        switch (c->P) {
#define X(p, r0, r1, r2) case p: mtd_kernel<MtdCfg<p, r0, r1, r2>><<<grid, RSP_MTD_THREADS, c->mtd_smem, c->cur->s>>>(a); break;
            RSP_FOR_EACH_POW2_P(X)
#undef X
        }
    } else {
#define X(tgv, r, kt) if (tg == tgv && c->mtd_r == r && c->mtd_kt == kt) mtd_dft_kernel<tgv, r, kt><<<grid, RSP_MTD_THREADS, c->mtd_smem, c->cur->s>>>(a);
        RSP_FOR_EACH_DFT(X)
#undef X
    }
}

static bool cfar_testable(const rsp_ctx* c) {
    const int mR = c->prm.guard_r + c->prm.ref_r, mV = c->prm.guard_v + c->prm.ref_v;
    return c->G - 2 * mR > 0 && c->P - 2 * mV > 0;      // else fsf:192-193 loop ranges are empty
}

static void launch_cfar(rsp_ctx* c, const float2* rdm, int slot, int cut_lo = -1, int cut_hi = -1) {
    if (!cfar_testable(c)) return;
    CfarArgs a;
    a.amp = c->cur->amp; a.rdm = rdm;
    a.c.P = c->P; a.c.G = c->G; a.c.guard_r = c->prm.guard_r; a.c.guard_v = c->prm.guard_v;
    a.c.ref_r = c->prm.ref_r; a.c.ref_v = c->prm.ref_v; a.c.t_cfar = c->prm.t_cfar; a.c.pad_pitch = c->cfar_pad ? 1 : 0;
    a.count = c->d_counts + slot;
    a.raw = c->d_rawdet + (size_t)slot * c->prm.max_detections;
    a.cap = c->prm.max_detections;
    a.complex_mode = c->prm.monopulse_complex;
    a.dead = cut_lo < 0 ? dead_pc(c) : DiscardArgs{nullptr, 0};
    const int mR = c->prm.guard_r + c->prm.ref_r;
    a.cut_lo = cut_lo < 0 ? mR : cut_lo;
    a.cut_hi = cut_hi < 0 ? c->G - mR : cut_hi;
    if (a.cut_hi <= a.cut_lo) return;
    const int ncut = a.cut_hi - a.cut_lo, tg = c->cfar_tg;
    dim3 grid((ncut + tg - 1) / tg, c->B - 1);
    Timed t(c, K_CFAR);
#define LAUNCH(K) K<<<grid, RSP_CFAR_THREADS, c->cfar_smem, c->cur->s>>>(a);
    if (c->cfar5) { RSP_CFAR5_DISPATCH(LAUNCH) }
    else if (tg == 64) { RSP_CFAR_DISPATCH(64, LAUNCH) }
    else if (tg == 32) { RSP_CFAR_DISPATCH(32, LAUNCH) }
    else { RSP_CFAR_DISPATCH(16, LAUNCH) }
#undef LAUNCH
}

// S9 for slots [first, first + n) on stream st: one small launch per batch
static void launch_refine(rsp_ctx* c, int first, int n, cudaStream_t st) {
    if (!cfar_testable(c) || n <= 0) return;
    RefineArgs r;
    r.counts = c->d_counts; r.raw = c->d_rawdet; r.recs = c->d_recs; r.cap = c->prm.max_detections; r.first_slot = first;
    r.range_axis = c->d_range_axis; r.vel_axis = c->d_vel_axis; r.beam_angles = c->d_beam_angles; r.k_slopes = c->d_k_slopes;
    r.delta_r = c->delta_r; r.delta_v = c->delta_v; r.complex_mode = c->prm.monopulse_complex;
    c->launches++;
    refine_kernel<<<dim3(4, n), 128, 0, st>>>(r);
}


// ------------------------------------------------------------------------------------------
// Range-blocked chain.  Chunk 0 = narrow + medium segments (gates [0, g1 + g2)); then one chunk per overlap-save block of
// the long segment, in gate order.  Per chunk: DBF of the 128-sample tiles the block reads, the block's FFT . H . IFFT over
// all lines, the Doppler FFT of its gates, and the CFAR of the cells whose range windows are complete (they trail the
// Doppler stage by guard + ref gates).  The beam / pc / amplitude buffers keep their full size; only the chunk's part of
// each is touched between two HBM-bound stages, so it is still in L2 when the next kernel reads it.
// ------------------------------------------------------------------------------------------
static void plan_blocking(rsp_ctx* c) {
    c->blocked = false;
    c->chunks.clear();
    // Opt-in (RSP_BLOCK=1): measured at config 3, where beam + pc cubes are 500 MB, the chunked chain is SLOWER than the
    // whole-CPI kernels (631 vs 526 us per CPI on one lane, 541 on three; profiles/r2_block_cfg3.txt): every stage loses
    // more to its five short launches than it gains from reading its input out of L2.  Kept because it is exact
    // (bit-identical, tested) and is the scaffolding a fused per-chunk kernel would need.
    const char* e = getenv("RSP_BLOCK");
    const bool want = e && atoi(e) != 0;
    if (!want || !c->dbf_tc || c->dbf_pc_ok || !c->lng.L) return;
    { const char* el = getenv("RSP_BLOCK_LANES"); c->block_lanes = el ? std::max(1, std::min(8, atoi(el))) : 1; }
    const int mR = c->prm.guard_r + c->prm.ref_r;
    const int g12 = c->prm.n_gates[0] + c->prm.n_gates[1];
    auto tiles_of = [&](int s_lo, int s_hi, int& tl, int& th) {
        tl = std::max(0, s_lo) / RSP_TC_TILE;
        th = std::min((std::min(s_hi, c->N) + RSP_TC_TILE - 1) / RSP_TC_TILE, c->dbf_tc_args.tiles_per_pulse);
    };
    int cut = mR;
    auto push = [&](rsp_ctx::Chunk ch) {
        ch.cut_lo = cut;
        ch.cut_hi = std::max(cut, std::min(ch.g_hi == c->G ? c->G - mR : ch.g_hi - mR, c->G - mR));
        cut = ch.cut_hi;
        c->chunks.push_back(ch);
    };
    if (g12 > 0) {                                                          // narrow FIR + medium block(s)
        rsp_ctx::Chunk ch{};
        int s_hi = c->prm.seg_start[0] - 1 + c->prm.n_gates[0] + c->prm.fir_delay + c->n_fir;
        if (c->med.L) s_hi = std::max(s_hi, c->med.seg_start0 + c->med.gate0 - (c->med.taps - 1) + (c->med.nblk - 1) * c->med.valid + c->med.L);
        tiles_of(0, s_hi, ch.tile_lo, ch.tile_hi);
        ch.medium = true; ch.part = -1; ch.blk = 0; ch.g_lo = 0; ch.g_hi = g12;
        push(ch);
    }
    const PcPlan* parts[3] = {&c->lng, &c->lngx[0], &c->lngx[1]};
    for (int q = 0; q < 3; ++q) {
        const PcPlan& pl = *parts[q];
        if (!pl.L) continue;
        for (int b = 0; b < pl.nblk; ++b) {
            rsp_ctx::Chunk ch{};
            const int g0 = pl.gate0 + b * pl.valid;
            const int s0 = pl.seg_start0 + g0 - (pl.taps - 1);
            tiles_of(s0, s0 + pl.L, ch.tile_lo, ch.tile_hi);
            ch.medium = false; ch.part = q; ch.blk = b;
            ch.g_lo = g0; ch.g_hi = std::min(g0 + pl.valid, pl.gate0 + pl.ngates);
            push(ch);
        }
    }
    if (c->chunks.empty() || c->chunks.back().g_hi != c->G) { c->chunks.clear(); return; }
    c->blocked = true;
}

static void launch_pc_chunk(rsp_ctx* c, const rsp_ctx::Chunk& ch) {
    PcKernelArgs a;
    a.beam = c->cur->beam; a.pc = c->cur->pc; a.N = c->N; a.ldb = c->ldb; a.ldg = c->ldg;
    a.group_bar = c->pc_group_bar;
    a.fir = c->d_fir; a.nfir = c->n_fir; a.fir_delay = c->prm.fir_delay;
    a.narrow_start0 = c->prm.seg_start[0] - 1; a.narrow_gates = c->prm.n_gates[0];
    PcPlan none;
    int la = 1024, lb = 1024;
    if (ch.medium) {
        const bool narrow = c->prm.n_gates[0] > 0;
        if (narrow && !c->med.L) {
            Timed t(c, K_PC_NARROW);
            pc_narrow_kernel<<<c->P * c->B, 256, 0, c->cur->s>>>(c->cur->beam, c->cur->pc, c->d_fir, c->n_fir, c->prm.fir_delay, c->N,
                                                                 c->ldb, c->ldg, c->prm.seg_start[0] - 1, c->prm.n_gates[0]);
            return;
        }
        if (!c->med.L) return;
        fill_seg(c, a.seg[0], none, nullptr, nullptr, nullptr);
        fill_seg(c, a.seg[1], c->med, c->d_med_tw1, c->d_med_tw2, c->d_med_H);
        a.do_narrow = narrow ? 1 : 0;
        lb = c->med.L;
    } else {
        const PcPlan* parts[3] = {&c->lng, &c->lngx[0], &c->lngx[1]};
        const float2* tw1[3] = {c->d_lng_tw1, c->d_lngx_tw1[0], c->d_lngx_tw1[1]};
        const float2* tw2[3] = {c->d_lng_tw2, c->d_lngx_tw2[0], c->d_lngx_tw2[1]};
        const float2* H[3] = {c->d_lng_H, c->d_lngx_H[0], c->d_lngx_H[1]};
        const PcPlan& pl = *parts[ch.part];
        fill_seg(c, a.seg[0], pl, tw1[ch.part], tw2[ch.part], H[ch.part]);
        // one block of every line: block index 0 of a segment that starts at this block's first gate
        a.seg[0].gate0 = pl.gate0 + ch.blk * pl.valid;
        a.seg[0].nblk = 1;
        a.seg[0].n_items = c->P * c->B;
        a.seg[0].n_ctas = (a.seg[0].n_items + RSP_PC_THREADS / pl.T - 1) / (RSP_PC_THREADS / pl.T);
        fill_seg(c, a.seg[1], none, nullptr, nullptr, nullptr);
        a.do_narrow = 0;
        la = pl.L;
    }
    const int nctas = a.seg[0].n_ctas + a.seg[1].n_ctas;
    if (nctas == 0) return;
    Timed t(c, K_PC);
#define X(A, B) if (la == A::L && lb == B::L) pc_fft_kernel<A, B><<<nctas, RSP_PC_THREADS, pc_smem_pair<A, B>(), c->cur->s>>>(a);
    RSP_FOR_EACH_PC_PAIR(X)
#undef X
}

static int enqueue_chain_blocked(rsp_ctx* c, const float2* raw, float2* rdm, int slot) {
    bool first = true;
    for (const rsp_ctx::Chunk& ch : c->chunks) {
        if (int rc = launch_dbf_tc(c, raw, first ? c->d_counts + slot : nullptr, ch.tile_lo, ch.tile_hi)) return rc;
        first = false;
        launch_pc_chunk(c, ch);
        launch_mtd(c, rdm, ch.g_lo, ch.g_hi);
        if (cfar_testable(c)) launch_cfar(c, rdm, slot, ch.cut_lo, ch.cut_hi);
    }
    return RSP_OK;
}

static int kernels_per_cpi(const rsp_ctx* c) {
    const bool narrow = c->prm.n_gates[0] > 0;
    int n = 1 /*dbf*/ + (narrow && c->med.L == 0) + ((c->med.L > 0 || c->lng.L > 0) ? 1 : 0) + (c->lngx[0].L ? 1 : 0) + 1 /*mtd*/;
    if (c->dbf_pc_ok) n = 1 /*dbf_pc*/ + 1 /*mtd*/;
    const bool frame_ctx = c->have_waveform && c->dbf_nt && c->dbf_wa && (c->N % 2 == 0);
    if (!c->dbf_pc_ok && c->dbf_tc && !frame_ctx && c->pulse_block > 0 && c->pulse_block < c->P)   // pulse groups: each its own DBF + PC launches
        n = (n - 1) * ((c->P + c->pulse_block - 1) / c->pulse_block) + 1;
    if (c->blocked) return (int)c->chunks.size() * (cfar_testable(c) ? 4 : 3);
    if (cfar_testable(c)) n += 1;   // cfar (S9 is one refine launch per batch, not per CPI)
    return n;
}

// enqueue S5..S9 for one device-resident PCN cube on lane `lane`
static int enqueue_chain(rsp_ctx* c, const float2* raw, float2* rdm, int slot, int lane) {
    c->cur = &c->lanes[lane];
    c->exp_raw = raw;
    // RSP_STAGES (bit mask 1 = DBF, 2 = PC, 4 = MTD, 8 = CFAR) is a measurement aid for tools/stage_probe.py: it
    // leaves stages out so that the steady-state cost of each kernel on the lanes can be timed in isolation.
    const int stages = c->stages;
    int rc = RSP_OK;
    const bool frame_ctx = c->have_waveform && c->dbf_nt && c->dbf_wa && (c->N % 2 == 0);      // see launch_dbf_any
    if (c->blocked && !c->fused && !frame_ctx && stages == 15 && !c->keep_beam && !(reinterpret_cast<uintptr_t>(raw) & 15)) {
        rc = enqueue_chain_blocked(c, raw, rdm, slot);
        c->cur = &c->lanes[0];
        if (rc) return rc;
        CU(c, cudaGetLastError());
        return RSP_OK;
    }
    if (c->dbf_pc_ok && !c->fused && (stages & 3) == 3) {
        rc = launch_dbf_pc(c, raw, c->d_counts + slot);                   // S5 + S6 in one launch; zeroes the slot's counter
        if (rc) return rc;
    } else {
        // Opt-in (RSP_PULSE_BLOCK) pulse-blocked S5 -> S6: the beams of a group of pulses are formed and pulse-compressed before
        // the next group is touched, so that at the big shapes the beam cube is consumed out of L2 instead of costing a DRAM
        // write + read (config 3: 2 x 268 MB of 2.0 GB per CPI).  Same kernels, bit-identical results -- and slower: the short
        // launches lose more to ramp-up and tails than the DRAM traffic gives back (see plan_dbf_tc).
        const bool tc_ok = c->dbf_tc && !frame_ctx && !c->fused && !(reinterpret_cast<uintptr_t>(raw) & 15);
        if (c->pulse_block > 0 && c->pulse_block < c->P && tc_ok && stages == 15) {
            for (int p0 = 0; p0 < c->P && !rc; p0 += c->pulse_block) {
                const int p1 = std::min(c->P, p0 + c->pulse_block);
                // the beams of the previous group have been consumed by its pulse compression: drop their dirty lines from L2
                // instead of writing them back (the first group discards the previous CPI's amplitude map, like every DBF)
                const size_t per_pulse = (size_t)c->B * c->ldb * sizeof(float2);
                const DiscardArgs prev = p0 == 0 ? DiscardArgs{nullptr, 0}
                    : dead_buf(c, reinterpret_cast<char*>(c->cur->beam) + (size_t)(p0 - c->pulse_block) * per_pulse, (size_t)c->pulse_block * per_pulse);
                rc = launch_dbf_tc(c, raw, p0 == 0 ? c->d_counts + slot : nullptr, 0, -1, p0, p1, prev);
                if (!rc) launch_pc(c, p0, p1);
            }
            if (rc) return rc;
        } else {
        if (stages & 1) rc = launch_dbf_any(c, raw, c->d_counts + slot);  // dbf_kernel also zeroes the slot's counter
        if (rc) return rc;
        if (stages & 2) launch_pc(c);
        }
    }
    if (stages & 4) launch_mtd(c, rdm);
    if ((stages & 8) && cfar_testable(c)) launch_cfar(c, rdm, slot);
    c->cur = &c->lanes[0];
    CU(c, cudaGetLastError());
    return RSP_OK;
}

// RDM buffer of a lane for the pipelined paths: concurrent lanes must not share one
static float2* lane_rdm(rsp_ctx* c, int l) {
    if (l == 0) return c->d_rdm;
    rsp_ctx::Lane& ln = c->lanes[l];
    if (!ln.rdm && cudaMalloc(reinterpret_cast<void**>(&ln.rdm), (size_t)c->P * c->B * c->G * sizeof(float2)) != cudaSuccess) return nullptr;
    return ln.rdm;
}


static bool det_less(const rsp_detection& a, const rsp_detection& b) {
    if (a.pair_idx != b.pair_idx) return a.pair_idx < b.pair_idx;
    if (a.r_idx != b.r_idx) return a.r_idx < b.r_idx;
    return a.v_idx < b.v_idx;
}

// The reference's order (pair, then range, then Doppler: MATLAB find, fsf:215-221).  Dense frames (1600 records of 40 bytes)
// are sorted through one 64-bit key per record and gathered once; comparing and swapping whole records cost three times as much.
static void sort_detections_keyed(rsp_detection* dets, int n, std::vector<std::pair<uint64_t, int>>& keys, std::vector<rsp_detection>& tmp) {
    if (n < 64) { std::sort(dets, dets + n, det_less); return; }
    keys.resize((size_t)n);
    for (int i = 0; i < n; ++i)
        keys[i] = {((uint64_t)(uint32_t)dets[i].pair_idx << 48) | ((uint64_t)(uint32_t)dets[i].r_idx << 24) | (uint64_t)(uint32_t)dets[i].v_idx, i};
    std::sort(keys.begin(), keys.end());
    tmp.assign(dets, dets + n);
    for (int i = 0; i < n; ++i) dets[i] = tmp[(size_t)keys[i].second];
}
static void sort_detections_inplace(rsp_ctx* c, rsp_detection* dets, int n) { sort_detections_keyed(dets, n, c->sort_keys, c->sort_tmp); }

// bring the input cube to device PCN complex64; returns the device pointer to use
static int stage_input(rsp_ctx* c, const void* raw, rsp_layout layout, rsp_dtype dtype, rsp_mem mem, const float2** out) {
    const size_t n = (size_t)c->P * c->C * c->N;
    const size_t esz = dtype == RSP_C64 ? sizeof(float2) : sizeof(double2);
    if (layout == RSP_LAYOUT_PCN && dtype == RSP_C64) {
        if (mem == RSP_MEM_DEVICE) { *out = static_cast<const float2*>(raw); return RSP_OK; }
        CU(c, cudaMemcpyAsync(c->d_raw, raw, n * esz, cudaMemcpyHostToDevice, c->stream));
        *out = c->d_raw;
        return RSP_OK;
    }
    const void* src = raw;
    if (mem == RSP_MEM_HOST) {
        if (c->stage_bytes < n * esz) {
            if (c->d_stage) cudaFree(c->d_stage);
            c->d_stage = nullptr;
            CU(c, cudaMalloc(&c->d_stage, n * esz));
            c->stage_bytes = n * esz;
        }
        CU(c, cudaMemcpyAsync(c->d_stage, raw, n * esz, cudaMemcpyHostToDevice, c->stream));
        src = c->d_stage;
    }
    Timed t(c, K_CONVERT);
    if (layout == RSP_LAYOUT_PCN) {
        c128_to_c64_kernel<<<1184, 256, 0, c->stream>>>(static_cast<const double2*>(src), c->d_raw, n);
    } else {
        dim3 grid((c->P + 31) / 32, (c->N + 31) / 32, c->C), blk(32, 8);
        if (dtype == RSP_C64)
            matlab_to_pcn_kernel<float2><<<grid, blk, 0, c->stream>>>(static_cast<const float2*>(src), c->d_raw, c->P, c->N, c->C);
        else
            matlab_to_pcn_kernel<double2><<<grid, blk, 0, c->stream>>>(static_cast<const double2*>(src), c->d_raw, c->P, c->N, c->C);
    }
    CU(c, cudaGetLastError());
    *out = c->d_raw;
    return RSP_OK;
}

// A slot whose previous pipelined submission has not been fetched must not be reused: its pinned descriptor block /
// host cube may still be read by the asynchronous copy of that submission, and its counts and records would be replaced.
static int check_slot_free(rsp_ctx* c, int slot) {
    if (!c->slot_prefetched.empty() && c->slot_prefetched[slot])
        return fail(c, RSP_ERR_INVALID_ARG, "slot %d was submitted and not fetched yet (pair every submit with a fetch)", slot);
    return RSP_OK;
}

// end of a pipelined submission on lane l: prefetch the slot's result to pinned memory and mark the slot
static int finish_submit(rsp_ctx* c, int slot, int l) {
    cudaStream_t s = c->lanes[l].s;
    if (!c->h_slot_count) {
        c->prefetch_k = std::min(2048, c->prm.max_detections);   // a 64-target frame has ~1600 records: no second copy
        CU(c, cudaStreamCreateWithFlags(&c->fetch_stream, cudaStreamNonBlocking));
        CU(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_slot_count), (size_t)c->slots * sizeof(int)));
        CU(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_slot_recs), (size_t)c->slots * c->prefetch_k * sizeof(rsp_detection)));
        c->slot_done.resize((size_t)c->slots);
        for (auto& e : c->slot_done) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->slot_prefetched.assign((size_t)c->slots, 0);
    }
    CU(c, cudaMemcpyAsync(c->h_slot_count + slot, c->d_counts + slot, sizeof(int), cudaMemcpyDeviceToHost, s));
    CU(c, cudaMemcpyAsync(c->h_slot_recs + (size_t)slot * c->prefetch_k, c->d_recs + (size_t)slot * c->prm.max_detections,
                          (size_t)c->prefetch_k * sizeof(rsp_detection), cudaMemcpyDeviceToHost, s));
    CU(c, cudaEventRecord(c->slot_done[slot], s));
    c->slot_prefetched[slot] = 1;
    c->slot_lane[slot] = l;
    return RSP_OK;
}

static int fetch_slot(rsp_ctx* c, int slot, rsp_detection* dets, int32_t det_cap, int32_t* n_dets, bool sort = true) {
    const bool pre = !c->slot_prefetched.empty() && c->slot_prefetched[slot];
    int n;
    if (pre) {                            // pipelined submission: its result is already on its way to pinned memory
        CU(c, cudaEventSynchronize(c->slot_done[slot]));
        c->slot_prefetched[slot] = 0;
        c->slot_lane[slot] = -1;
        n = c->h_slot_count[slot];
    } else {
        CU(c, cudaMemcpyAsync(c->h_count, c->d_counts + slot, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        n = *c->h_count;
    }
    if (n_dets) *n_dets = n;
    if (n > c->prm.max_detections)
        return fail(c, RSP_ERR_OVERFLOW, "%d detections exceed max_detections=%d", n, c->prm.max_detections);
    if (n > det_cap) return fail(c, RSP_ERR_OVERFLOW, "%d detections exceed the caller's capacity %d", n, det_cap);
    if (n > 0) {
        if (pre && n <= c->prefetch_k) {
            std::memcpy(dets, c->h_slot_recs + (size_t)slot * c->prefetch_k, (size_t)n * sizeof(rsp_detection));
        } else if (pre) {
            // a dense frame: the records beyond the prefetched block come over on a stream of their own (the slot's kernels
            // have finished -- slot_done --, and the lanes' streams already hold the work of later frames)
            const size_t k0 = (size_t)c->prefetch_k;
            CU(c, cudaMemcpyAsync(c->h_recs, c->d_recs + (size_t)slot * c->prm.max_detections + k0, ((size_t)n - k0) * sizeof(rsp_detection),
                                  cudaMemcpyDeviceToHost, c->fetch_stream));
            std::memcpy(dets, c->h_slot_recs + (size_t)slot * k0, k0 * sizeof(rsp_detection));
            CU(c, cudaStreamSynchronize(c->fetch_stream));
            std::memcpy(dets + k0, c->h_recs, ((size_t)n - k0) * sizeof(rsp_detection));
        } else {
            CU(c, cudaMemcpyAsync(c->h_recs, c->d_recs + (size_t)slot * c->prm.max_detections, (size_t)n * sizeof(rsp_detection),
                                  cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            std::memcpy(dets, c->h_recs, (size_t)n * sizeof(rsp_detection));
        }
        if (sort) sort_detections_inplace(c, dets, n);
    }
    return RSP_OK;
}

extern "C" {

int rsp_process_cpi(rsp_ctx* c, const void* raw, rsp_layout layout, rsp_dtype dtype, rsp_mem raw_mem, void* rdm_out,
                    rsp_mem rdm_mem, rsp_detection* dets, int32_t det_cap, int32_t* n_dets) {
    if (!c || !raw || !n_dets || (det_cap > 0 && !dets)) return fail(c, RSP_ERR_INVALID_ARG, "null argument");
    if (!c->have_constants) return fail(c, RSP_ERR_NOT_READY, "rsp_upload_constants has not been called");
    CU(c, cudaSetDevice(c->prm.device));
    const float2* d_in = nullptr;
    int rc = stage_input(c, raw, layout, dtype, raw_mem, &d_in);
    if (rc) return rc;
    float2* rdm = (rdm_out && rdm_mem == RSP_MEM_DEVICE) ? static_cast<float2*>(rdm_out) : c->d_rdm;
    c->keep_beam = true;                       // rsp_get_beam: the fused kernel writes the beam lines out as well
    rc = enqueue_chain(c, d_in, rdm, 0, 0);
    c->keep_beam = false;
    if (rc) return rc;
    if (!c->slot_prefetched.empty()) c->slot_prefetched[0] = 0;
    launch_refine(c, 0, 1, c->stream);
    if (rdm_out && rdm_mem == RSP_MEM_HOST)
        CU(c, cudaMemcpyAsync(rdm_out, c->d_rdm, (size_t)c->P * c->B * c->G * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    c->ran = true;
    c->rdm_in_ctx = (rdm == c->d_rdm);
    return fetch_slot(c, 0, dets, det_cap, n_dets);
}

// The launches of one rsp_stream_enqueue call: fork the lanes from the caller's stream, deal the CPIs round-robin, join,
// one S9 launch for the batch.
static int stream_enqueue_body(rsp_ctx* c, const void* raw_dev, int raw_pool, void* rdm_dev, int rdm_pool, int n_cpi, int first_slot, int nl) {
    const size_t in_elems = (size_t)c->P * c->C * c->N, out_elems = (size_t)c->P * c->B * c->G;
    const bool own_rdm = !(rdm_dev && rdm_pool > 0);
    if (nl > 1) {                                    // fork: the extra lanes wait for work already on the caller's stream
        CU(c, cudaEventRecord(c->fork, c->stream));
        for (int l = 1; l < nl; ++l) CU(c, cudaStreamWaitEvent(c->lanes[l].s, c->fork, 0));
    }
    for (int i = 0; i < n_cpi; ++i) {
        const float2* in = static_cast<const float2*>(raw_dev) + (size_t)(i % raw_pool) * in_elems;
        float2* rdm = own_rdm ? lane_rdm(c, i % nl) : static_cast<float2*>(rdm_dev) + (size_t)(i % rdm_pool) * out_elems;
        if (!rdm) return fail(c, RSP_ERR_CUDA, "out of device memory for the lane's RDM");
        int rc = enqueue_chain(c, in, rdm, first_slot + i, i % nl);
        if (rc) return rc;
    }
    for (int l = 1; l < nl; ++l) {                   // join
        CU(c, cudaEventRecord(c->lanes[l].done, c->lanes[l].s));
        CU(c, cudaStreamWaitEvent(c->stream, c->lanes[l].done, 0));
    }
    launch_refine(c, first_slot, n_cpi, c->stream);  // S9 for the whole batch
    CU(c, cudaGetLastError());
    return RSP_OK;
}

// CUDA graph per batch.  A steady stream calls rsp_stream_enqueue with the same few argument sets over and over (pool
// pointers, ring slots); the second time a set is seen its launches are captured (stream capture of the very same code,
// lanes included) and from then on one cudaGraphLaunch replaces n_cpi x kernels_per_cpi launches plus the fork / join
// events: host cost per CPI drops from ~16 us to under 2 us (tools/enqueue_cost.py), which is what bounds the small
// shapes (config 1: ~20 us of device time per CPI).  Keyed on everything the kernel arguments depend on; any change of
// constants, stream or profiling state (caller_epoch) drops the cache.  RSP_GRAPH=0 keeps the direct launches.
static void drop_graphs(rsp_ctx* c) {
    for (auto& g : c->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}
static int stream_enqueue_graphed(rsp_ctx* c, const void* raw_dev, int raw_pool, void* rdm_dev, int rdm_pool, int n_cpi, int first_slot, int nl) {
    static const bool enabled = [] { const char* e = getenv("RSP_GRAPH"); return !(e && atoi(e) == 0); }();
    if (!enabled || c->profiling || n_cpi < 2)
        return stream_enqueue_body(c, raw_dev, raw_pool, rdm_dev, rdm_pool, n_cpi, first_slot, nl);
    if (c->graph_epoch != c->caller_epoch) { drop_graphs(c); c->graph_epoch = c->caller_epoch; }
    rsp_ctx::GraphEntry key{};
    key.raw = raw_dev; key.rdm = rdm_dev; key.raw_pool = raw_pool; key.rdm_pool = rdm_pool; key.n_cpi = n_cpi; key.first_slot = first_slot;
    key.flags = (c->discard ? 1 : 0) | (c->keep_beam ? 2 : 0) | (nl << 8);
    rsp_ctx::GraphEntry* hit = nullptr;
    for (auto& g : c->graphs)
        if (g.raw == key.raw && g.rdm == key.rdm && g.raw_pool == key.raw_pool && g.rdm_pool == key.rdm_pool && g.n_cpi == key.n_cpi &&
            g.first_slot == key.first_slot && g.flags == key.flags) { hit = &g; break; }
    if (!hit) {                                      // first sight: remember the set, launch directly
        if (c->graphs.size() >= 16) { if (c->graphs.front().exec) cudaGraphExecDestroy(c->graphs.front().exec); c->graphs.erase(c->graphs.begin()); }
        c->graphs.push_back(key);
        return stream_enqueue_body(c, raw_dev, raw_pool, rdm_dev, rdm_pool, n_cpi, first_slot, nl);
    }
    if (!hit->exec && !hit->failed) {                // second sight: capture
        const long before = c->launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            if (getenv("RSP_GRAPH_DEBUG")) fprintf(stderr, "librsp: cudaStreamBeginCapture failed: %s\n", cudaGetErrorString(cudaPeekAtLastError()));
            cudaGetLastError();
            hit->failed = true;
        }
        else {
            const int rc = stream_enqueue_body(c, raw_dev, raw_pool, rdm_dev, rdm_pool, n_cpi, first_slot, nl);
            const cudaError_t ee = cudaStreamEndCapture(c->stream, &graph);
            hit->launches = c->launches - before;
            c->launches = before;
            cudaError_t ei = cudaSuccess;
            if (rc != RSP_OK || ee != cudaSuccess || !graph || (ei = cudaGraphInstantiate(&hit->exec, graph, 0)) != cudaSuccess) {
                if (getenv("RSP_GRAPH_DEBUG")) fprintf(stderr, "librsp: graph capture failed: rc %d (%s), end-capture %s, instantiate %s\n", rc, c->err.c_str(), cudaGetErrorString(ee), cudaGetErrorString(ei));
                cudaGetLastError();
                hit->exec = nullptr;
                hit->failed = true;
            }
            if (graph) cudaGraphDestroy(graph);
        }
    }
    if (!hit->exec) return stream_enqueue_body(c, raw_dev, raw_pool, rdm_dev, rdm_pool, n_cpi, first_slot, nl);
    CU(c, cudaGraphLaunch(hit->exec, c->stream));
    c->launches += hit->launches;
    c->graph_launches++;
    return RSP_OK;
}

int rsp_stream_enqueue(rsp_ctx* c, const void* raw_dev, int32_t raw_pool, void* rdm_dev, int32_t rdm_pool, int32_t n_cpi,
                       int32_t first_slot) {
    if (!c || !raw_dev || raw_pool < 1 || n_cpi < 0) return fail(c, RSP_ERR_INVALID_ARG, "bad stream arguments");
    if (!c->have_constants) return fail(c, RSP_ERR_NOT_READY, "rsp_upload_constants has not been called");
    if (first_slot < 0 || first_slot + n_cpi > c->slots)
        return fail(c, RSP_ERR_INVALID_ARG, "slots [%d,%d) exceed the ring of %d", first_slot, first_slot + n_cpi, c->slots);
    CU(c, cudaSetDevice(c->prm.device));
    const int nl = std::min(c->blocked ? std::min(c->n_lanes, c->block_lanes) : c->n_lanes, std::max(n_cpi, 1));
    // Up to nl CPIs are in flight at once, one per lane: their range-Doppler maps must not share a buffer (MTD of CPI i + 1
    // would race with MTD and CFAR of CPI i).  A caller ring needs at least nl buffers, n_cpi > rdm_pool needs
    // rdm_pool % nl == 0 (CPI i goes to lane i % nl and to buffer i % rdm_pool); without a ring every lane uses its own map.
    const bool own_rdm = !(rdm_dev && rdm_pool > 0);
    if (!own_rdm && nl > 1 && (rdm_pool < nl || (n_cpi > rdm_pool && rdm_pool % nl != 0)))
        return fail(c, RSP_ERR_INVALID_ARG, "rdm_pool %d cannot serve %d concurrent lanes (need rdm_pool >= lanes, and a multiple of lanes when n_cpi > rdm_pool)", rdm_pool, nl);
    { const char* e = getenv("RSP_L2_DISCARD"); c->discard = !(e && atoi(e) == 0); }
    if (own_rdm)                                     // allocate before a capture could be open
        for (int l = 0; l < nl; ++l)
            if (!lane_rdm(c, l)) return fail(c, RSP_ERR_CUDA, "out of device memory for the lane's RDM");
    int rc = stream_enqueue_graphed(c, raw_dev, raw_pool, rdm_dev, rdm_pool, n_cpi, first_slot, nl);
    c->discard = false;
    if (rc) return rc;
    for (int i = 0; i < n_cpi; ++i) {
        c->slot_lane[first_slot + i] = -1;
        if (!c->slot_prefetched.empty()) c->slot_prefetched[first_slot + i] = 0;
    }
    return RSP_OK;
}

static int lane_fork(rsp_ctx* c, int l) {
    rsp_ctx::Lane& ln = c->lanes[l];
    if (l > 0 && ln.seen_epoch != c->caller_epoch) {
        CU(c, cudaEventRecord(c->fork, c->stream));
        CU(c, cudaStreamWaitEvent(ln.s, c->fork, 0));
    }
    ln.seen_epoch = c->caller_epoch;
    return RSP_OK;
}

int rsp_submit_cpi(rsp_ctx* c, const void* raw_host, void* rdm_dev, int32_t slot) {
    if (!c || !raw_host || slot < 0 || slot >= c->slots) return fail(c, RSP_ERR_INVALID_ARG, "bad submit arguments");
    if (!c->have_constants) return fail(c, RSP_ERR_NOT_READY, "rsp_upload_constants has not been called");
    CU(c, cudaSetDevice(c->prm.device));
    if (int rcs = check_slot_free(c, slot)) return rcs;
    const int l = slot % c->n_lanes;
    rsp_ctx::Lane& ln = c->lanes[l];
    const size_t bytes = (size_t)c->P * c->C * c->N * sizeof(float2);
    if (!ln.raw) CU(c, cudaMalloc(reinterpret_cast<void**>(&ln.raw), bytes));
    if (int rcf = lane_fork(c, l)) return rcf;
    CU(c, cudaMemcpyAsync(ln.raw, raw_host, bytes, cudaMemcpyHostToDevice, ln.s));
    float2* rdm = rdm_dev ? static_cast<float2*>(rdm_dev) : lane_rdm(c, l);
    if (!rdm) return fail(c, RSP_ERR_CUDA, "out of device memory for the lane's RDM");
    int rc = enqueue_chain(c, ln.raw, rdm, slot, l);
    if (rc) return rc;
    launch_refine(c, slot, 1, ln.s);
    return finish_submit(c, slot, l);
}

int rsp_stream_slots(const rsp_ctx* c) { return c ? c->slots : 0; }

int rsp_stream_device_buffers(rsp_ctx* c, void** counts_dev, void** records_dev) {
    if (!c) return RSP_ERR_INVALID_ARG;
    if (counts_dev) *counts_dev = c->d_counts;
    if (records_dev) *records_dev = c->d_recs;
    return RSP_OK;
}

int rsp_stream_fetch(rsp_ctx* c, int32_t slot, rsp_detection* dets, int32_t det_cap, int32_t* n_dets) {
    if (!c || slot < 0 || slot >= c->slots || !n_dets) return fail(c, RSP_ERR_INVALID_ARG, "bad slot");
    CU(c, cudaSetDevice(c->prm.device));
    return fetch_slot(c, slot, dets, det_cap, n_dets);
}

int rsp_sort_detections(rsp_detection* dets, int32_t n) {
    if (n < 0 || (n > 0 && !dets)) return RSP_ERR_INVALID_ARG;
    std::sort(dets, dets + n, det_less);
    return RSP_OK;
}

static int copy_out(rsp_ctx* c, void* dst, const void* src, size_t bytes) {
    if (!c || !dst) return RSP_ERR_INVALID_ARG;
    if (!c->ran) return fail(c, RSP_ERR_NOT_READY, "no CPI has been processed into the context buffers");
    CU(c, cudaSetDevice(c->prm.device));
    CU(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RSP_OK;
}

int rsp_get_beam(rsp_ctx* c, rsp_c64* dst) {
    if (!c || !dst) return RSP_ERR_INVALID_ARG;
    if (!c->ran) return fail(c, RSP_ERR_NOT_READY, "no CPI processed");
    CU(c, cudaSetDevice(c->prm.device));
    CU(c, cudaMemcpy2DAsync(dst, (size_t)c->N * sizeof(float2), c->lanes[0].beam, (size_t)c->ldb * sizeof(float2),
                            (size_t)c->N * sizeof(float2), (size_t)c->P * c->B, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RSP_OK;
}

int rsp_get_pc(rsp_ctx* c, rsp_c64* dst) {
    if (!c || !dst) return RSP_ERR_INVALID_ARG;
    if (!c->ran) return fail(c, RSP_ERR_NOT_READY, "no CPI processed");
    CU(c, cudaSetDevice(c->prm.device));
    CU(c, cudaMemcpy2DAsync(dst, (size_t)c->G * sizeof(float2), c->lanes[0].pc, (size_t)c->ldg * sizeof(float2),
                            (size_t)c->G * sizeof(float2), (size_t)c->P * c->B, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RSP_OK;
}

int rsp_get_rdm(rsp_ctx* c, rsp_c64* dst) {
    if (c && c->ran && !c->rdm_in_ctx) return fail(c, RSP_ERR_NOT_READY, "the last CPI wrote its RDM to a caller buffer");
    return copy_out(c, dst, c ? c->d_rdm : nullptr, c ? (size_t)c->P * c->B * c->G * sizeof(float2) : 0);
}

int rsp_get_amp(rsp_ctx* c, float* dst) {
    return copy_out(c, dst, c ? c->lanes[0].amp : nullptr, c ? (size_t)c->P * c->B * c->G * sizeof(float) : 0);
}

int rsp_process_frame(rsp_ctx* c, const void* raw, rsp_layout layout, rsp_dtype dtype, rsp_mem raw_mem,
                      const rsp_cluster_params* cp, rsp_target* final_targets, int32_t cap, int32_t* n_final) {
    if (!c || !cp || !n_final) return fail(c, RSP_ERR_INVALID_ARG, "null argument");
    std::vector<rsp_detection> dets((size_t)c->prm.max_detections);
    int32_t n = 0;
    int rc = rsp_process_cpi(c, raw, layout, dtype, raw_mem, nullptr, RSP_MEM_DEVICE, dets.data(), (int32_t)dets.size(), &n);
    if (rc) return rc;
    std::vector<rsp_target> fin((size_t)std::max(n, 1));
    int32_t nf = 0, n1 = 0;
    rc = rsp_cluster(dets.data(), n, cp, nullptr, &n1, fin.data(), &nf);
    if (rc) return fail(c, rc, "clustering failed");
    *n_final = nf;
    if (nf > cap) return fail(c, RSP_ERR_OVERFLOW, "%d targets exceed the caller's capacity %d", nf, cap);
    if (nf > 0 && !final_targets) return fail(c, RSP_ERR_INVALID_ARG, "null output");
    std::memcpy(final_targets, fin.data(), (size_t)nf * sizeof(rsp_target));
    return RSP_OK;
}

int rsp_stage2_configure(rsp_ctx* c, const rsp_stage2_config* cfg) {
    if (!c || !cfg) return RSP_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->prm.device));
    if (c->N < c->G) return fail(c, RSP_ERR_INVALID_ARG, "stage-2 contexts need n_samples >= sum(n_gates)");
    int gate0 = 0;
    for (int s = 0; s < 3; ++s) {
        rsp_ctx::S2Seg& sg = c->s2[s];
        const int ng = c->prm.n_gates[s], nt = cfg->n_pulse[s];
        sg.lo = gate0; sg.hi = gate0 + ng;
        sg.pl = PcPlan();
        if (ng > 0) {
            if (!cfg->pulse[s] || nt < 1 || nt > 4096) return fail(c, RSP_ERR_INVALID_ARG, "stage-2 pulse %d missing or too long", s);
            std::vector<zc> h(nt);                           // matched filter = conj(fliplr(pulse))
            for (int i = 0; i < nt; ++i) h[i] = std::conj(zc(cfg->pulse[s][nt - 1 - i].re, cfg->pulse[s][nt - 1 - i].im));
            const int L = choose_pc_len(nt, ng);
            // out[g] = sum_k h[k] y[(nt-1) + g - k]  ==  sum_k y[g + k] conj(pulse[k])
            if (!make_pc_plan(sg.pl, L, h.data(), nt, nt - 1, gate0, ng)) return fail(c, RSP_ERR_UNSUPPORTED, "no block plan for stage-2 segment %d", s);
            CU(c, upload(&sg.tw1, sg.pl.tw1)); CU(c, upload(&sg.tw2, sg.pl.tw2)); CU(c, upload(&sg.H, sg.pl.Hmid));
        }
        gate0 += ng;
    }
#define X(A, B) CU(c, opt_in_smem(pc_fft_kernel<A, B>, pc_smem_pair<A, B>()));
    RSP_FOR_EACH_PC_PAIR(X)
#undef X
    const int P = c->P;
    std::vector<float> win(P);
    c->pow2_doppler = make_doppler_plan(c->dop, P);
    for (int p = 0; p < P; ++p) {
        const double w = cfg->mtd_win ? cfg->mtd_win[p] : 1.0;
        win[p] = (float)(c->pow2_doppler ? w * ((p & 1) ? -1.0 : 1.0) : w);
    }
    CU(c, upload(&c->d_s2_win, win));
    c->h_s2_win = win;
    { const char* e = getenv("RSP_MTD"); c->mtd_mode = !e ? 2 : !strcmp(e, "tile") ? 0 : 2; }
    { const char* e = getenv("RSP_MTD_SQRT"); c->mtd_approx_sqrt = e && !strcmp(e, "approx"); }
    if (c->pow2_doppler) {
        CU(c, upload(&c->d_dop_tw, c->dop.tw));
        CU(c, upload(&c->d_dop_perm, c->dop.iperm));
        c->mtd_tg = RSP_MTD_TG;
        c->mtd_smem = ((size_t)P * (RSP_MTD_TG + 1) + c->dop.tw.size() + 1) * sizeof(float2);
        CU(c, mtd_opt_in(P, c->mtd_smem));
    } else {
        std::vector<float2> tw(P);
        for (int m = 0; m < P; ++m) {
            const double ang = -2.0 * kPi * (double)m / (double)P;
            tw[m] = make_float2((float)std::cos(ang), (float)std::sin(ang));
        }
        CU(c, upload(&c->d_dop_tw, tw));
        c->mtd_r = (P % 8 == 0) ? 8 : (P % 4 == 0) ? 4 : (P % 2 == 0) ? 2 : 1;
        c->mtd_tg = 0;
        if (!choose_dft_plan(P, c->mtd_r, &c->mtd_tg, &c->mtd_kt))
            return fail(c, RSP_ERR_UNSUPPORTED, "P=%d too large for the generic Doppler DFT kernel", P);
        c->mtd_smem = dft_smem_bytes(P, c->mtd_tg, c->mtd_r, c->mtd_kt);
#define X(tg, r, kt) if (c->mtd_tg == tg && c->mtd_r == r && c->mtd_kt == kt) CU(c, opt_in_smem(mtd_dft_kernel<tg, r, kt>, c->mtd_smem));
        RSP_FOR_EACH_DFT(X)
#undef X
    }
    c->s2_notch = cfg->zero_vel_bins < 0 ? 0 : cfg->zero_vel_bins;
    c->s2_ready = true;
    return RSP_OK;
}

static void launch_s2_pair(rsp_ctx* c, const rsp_ctx::S2Seg* a0, const rsp_ctx::S2Seg* a1) {
    PcKernelArgs a;
    a.beam = c->cur->beam; a.pc = c->cur->pc; a.N = c->G; a.ldb = c->ldb; a.ldg = c->ldg;
    const rsp_ctx::S2Seg* segs[2] = {a0, a1};
    int L[2] = {1024, 1024};
    for (int i = 0; i < 2; ++i) {
        PcSegArgs& sg = a.seg[i];
        std::memset(&sg, 0, sizeof sg);
        sg.nblk = 1;
        if (segs[i] && segs[i]->pl.L) {
            fill_seg(c, sg, segs[i]->pl, segs[i]->tw1, segs[i]->tw2, segs[i]->H);
            sg.in_lo = segs[i]->lo; sg.in_hi = segs[i]->hi;
            L[i] = segs[i]->pl.L;
        }
    }
    a.do_narrow = 0; a.fir = nullptr; a.nfir = 0; a.fir_delay = 0; a.narrow_start0 = 0; a.narrow_gates = 0;
    a.group_bar = c->pc_group_bar;
    const int nctas = a.seg[0].n_ctas + a.seg[1].n_ctas;
    if (nctas == 0) return;
    Timed t(c, K_PC);
#define X(A, B) if (L[0] == A::L && L[1] == B::L) pc_fft_kernel<A, B><<<nctas, RSP_PC_THREADS, pc_smem_pair<A, B>(), c->cur->s>>>(a);
    RSP_FOR_EACH_PC_PAIR(X)
#undef X
}

int rsp_stage2_mtd(rsp_ctx* c, const void* iq, rsp_dtype dtype, rsp_c128* mtd_out, rsp_c128* pc_out) {
    if (!c || !iq || !mtd_out) return fail(c, RSP_ERR_INVALID_ARG, "null argument");
    if (!c->s2_ready) return fail(c, RSP_ERR_NOT_READY, "rsp_stage2_configure has not been called");
    CU(c, cudaSetDevice(c->prm.device));
    const int P = c->P, G = c->G, B = c->B;
    const size_t n = (size_t)P * G * B;
    const size_t esz = dtype == RSP_C64 ? sizeof(float2) : sizeof(double2);
    if (c->stage_bytes < n * esz) {
        if (c->d_stage) cudaFree(c->d_stage);
        c->d_stage = nullptr;
        CU(c, cudaMalloc(&c->d_stage, n * esz));
        c->stage_bytes = n * esz;
    }
    if (!c->d_aux) CU(c, dev_alloc(&c->d_aux, n));
    c->cur = &c->lanes[0];
    CU(c, cudaMemcpyAsync(c->d_stage, iq, n * esz, cudaMemcpyHostToDevice, c->stream));
    {   // MATLAB [P,G,B] -> lines [p][b][g] in the lane's "beam" buffer (pitch ldb)
        Timed t(c, K_CONVERT);
        dim3 grid((P + 31) / 32, (G + 31) / 32, B), blk(32, 8);
        if (dtype == RSP_C64) pgb_to_pbg_kernel<float2><<<grid, blk, 0, c->stream>>>(static_cast<const float2*>(c->d_stage), c->cur->beam, P, G, B, c->ldb);
        else pgb_to_pbg_kernel<double2><<<grid, blk, 0, c->stream>>>(static_cast<const double2*>(c->d_stage), c->cur->beam, P, G, B, c->ldb);
    }
    launch_s2_pair(c, &c->s2[2], &c->s2[1]);      // long + medium
    launch_s2_pair(c, &c->s2[0], nullptr);        // narrow
    {   // MTD with the stage-2 window
        float* keep = c->d_win;
        c->d_win = c->d_s2_win;
        c->h_win.swap(c->h_s2_win);
        launch_mtd(c, c->d_rdm);
        c->h_win.swap(c->h_s2_win);
        c->d_win = keep;
    }
    if (c->s2_notch > 0) {
        const int ctr = P / 2;                    // zero Doppler after fftshift
        const int lo = std::max(0, ctr - c->s2_notch), hi = std::min(P - 1, ctr + c->s2_notch);
        Timed t(c, K_CONVERT);
        doppler_notch_kernel<<<592, 256, 0, c->stream>>>(c->d_rdm, (size_t)B * G, P, lo, hi);
    }
    std::vector<float2> h(n);
    auto to_host = [&](const float2* dev, rsp_c128* dst) -> int {
        CU(c, cudaMemcpyAsync(h.data(), dev, n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < n; ++i) dst[i] = rsp_c128{(double)h[i].x, (double)h[i].y};
        return RSP_OK;
    };
    int rc = to_host(c->d_rdm, mtd_out);          // rdm[b][g][v] is already MATLAB (v,g,b) byte order
    if (rc) return rc;
    if (pc_out) {
        Timed t(c, K_CONVERT);
        dim3 grid((P + 31) / 32, (G + 31) / 32, B), blk(32, 8);
        pbg_to_bgp_kernel<<<grid, blk, 0, c->stream>>>(c->cur->pc, c->d_aux, P, G, B, c->ldg);
        rc = to_host(c->d_aux, pc_out);
        if (rc) return rc;
    }
    CU(c, cudaGetLastError());
    c->ran = true;
    c->rdm_in_ctx = true;
    return RSP_OK;
}

static int cfar1d_fill(rsp_ctx* c, Cfar1dArgs& a, const rsp_cfar1d_params* p, int V, int R, int B) {
    if (!p || p->ref_cells < 1 || p->save_cells < 0 || p->ref_cells + p->save_cells > RSP_CFAR1D_MAXW || (p->method != 0 && p->method != 1))
        return fail(c, RSP_ERR_INVALID_ARG, "1-D CFAR: need ref >= 1, save >= 0, ref + save <= %d, method 0 or 1", RSP_CFAR1D_MAXW);
    const int W = p->ref_cells + p->save_cells;
    int g = 0;
    for (int s = 0; s < 3; ++s) {
        if (p->seg_len[s] < 0 || (p->seg_len[s] > 0 && p->seg_len[s] < 2 * W))       // every cell needs one window inside its segment
            return fail(c, RSP_ERR_UNSUPPORTED, "1-D CFAR: segment %d has %d gates, fewer than 2 (ref + save) = %d", s, p->seg_len[s], 2 * W);
        a.seg_lo[s] = g; g += p->seg_len[s]; a.seg_hi[s] = g;
    }
    if (g != R) return fail(c, RSP_ERR_INVALID_ARG, "1-D CFAR: segments add up to %d gates, the map has %d", g, R);
    a.V = V; a.R = R; a.B = B;
    a.ref = p->ref_cells; a.save = p->save_cells; a.method = p->method; a.t_cfar = p->t_cfar;
    const long center = (long)std::floor(V / 2.0 + 0.5) + 1;                         // round(V/2) + 1, 1-based (:446)
    const int n0 = p->zero_vel_bins < 0 ? 0 : p->zero_vel_bins;
    a.notch_lo = (int)std::max(1L, center - n0) - 1;
    a.notch_hi = (int)std::min((long)V, center + n0) - 1;
    return RSP_OK;
}

static void cfar1d_launch(const Cfar1dArgs& a, cudaStream_t s) {
    dim3 grid((a.V + 31) / 32, (a.R + RSP_CFAR1D_TG - 1) / RSP_CFAR1D_TG, a.B);
    cfar1d_kernel<<<grid, 256, 0, s>>>(a);
}

int rsp_cfar1d(int32_t device, const float* amp_host, int32_t V, int32_t R, int32_t B, const rsp_cfar1d_params* p,
               uint8_t* flags_host, float* thr_host) {
    if (!amp_host || !flags_host || V < 1 || R < 1 || B < 1) return fail(nullptr, RSP_ERR_INVALID_ARG, "1-D CFAR: null or empty map");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return fail(nullptr, RSP_ERR_NO_DEVICE, "no CUDA device visible; librsp has no CPU fallback"); }
    if (device < 0 || device >= ndev) return fail(nullptr, RSP_ERR_INVALID_ARG, "device %d of %d", device, ndev);
    Cfar1dArgs a{};
    if (int rc = cfar1d_fill(nullptr, a, p, V, R, B)) return rc;
    CU(nullptr, cudaSetDevice(device));
    const size_t n = (size_t)V * R * B;
    float *d_amp = nullptr, *d_thr = nullptr;
    unsigned char* d_flags = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_amp), n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_flags), n);
    if (e == cudaSuccess && thr_host) e = cudaMalloc(reinterpret_cast<void**>(&d_thr), n * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(d_amp, amp_host, n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        a.amp = d_amp; a.rdm = nullptr; a.flags = d_flags; a.thr = d_thr;
        cfar1d_launch(a, nullptr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(flags_host, d_flags, n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && thr_host) e = cudaMemcpy(thr_host, d_thr, n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_amp); cudaFree(d_flags); cudaFree(d_thr);
    if (e != cudaSuccess) return fail(nullptr, RSP_ERR_CUDA, "1-D CFAR failed: %s", cudaGetErrorString(e));
    return RSP_OK;
}

int rsp_stage2_cfar(rsp_ctx* c, const rsp_cfar1d_params* p, uint8_t* flags_host, float* thr_host) {
    if (!c || !flags_host) return fail(c, RSP_ERR_INVALID_ARG, "null argument");
    if (!c->s2_ready || !c->ran || !c->rdm_in_ctx) return fail(c, RSP_ERR_NOT_READY, "rsp_stage2_mtd has not produced a Doppler map in this context");
    Cfar1dArgs a{};
    if (int rc = cfar1d_fill(c, a, p, c->P, c->G, c->B)) return rc;
    CU(c, cudaSetDevice(c->prm.device));
    const size_t n = (size_t)c->P * c->G * c->B;
    unsigned char* d_flags = nullptr;
    float* d_thr = nullptr;
    CU(c, cudaMalloc(reinterpret_cast<void**>(&d_flags), n));
    if (thr_host && cudaMalloc(reinterpret_cast<void**>(&d_thr), n * sizeof(float)) != cudaSuccess) { cudaFree(d_flags); return fail(c, RSP_ERR_CUDA, "out of device memory"); }
    a.amp = nullptr; a.rdm = c->d_rdm; a.flags = d_flags; a.thr = d_thr;
    cfar1d_launch(a, c->stream);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(flags_host, d_flags, n, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && thr_host) e = cudaMemcpyAsync(thr_host, d_thr, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_flags); cudaFree(d_thr);
    if (e != cudaSuccess) return fail(c, RSP_ERR_CUDA, "1-D CFAR failed: %s", cudaGetErrorString(e));
    return RSP_OK;
}

int rsp_set_waveform(rsp_ctx* c, const rsp_waveform* w) {
    if (!c || !w || !w->tx_pulse) return RSP_ERR_INVALID_ARG;
    c->caller_epoch++;
    CU(c, cudaSetDevice(c->prm.device));
    std::vector<float2> tx(c->N);
    for (int i = 0; i < c->N; ++i) tx[i] = make_float2((float)w->tx_pulse[i].re, (float)w->tx_pulse[i].im);
    // the (at most three) non-zero stretches of the pulse train: only these are added per target
    int nseg = 0;
    for (int i = 0; i < c->N;) {
        if (tx[i].x == 0.f && tx[i].y == 0.f) { ++i; continue; }
        int j = i;
        while (j < c->N && !(tx[j].x == 0.f && tx[j].y == 0.f)) ++j;
        if (nseg == 3) { c->tx_seg_hi[2] = j; }              // more than three stretches: merge the tail
        else { c->tx_seg_lo[nseg] = i; c->tx_seg_hi[nseg] = j; ++nseg; }
        i = j;
    }
    for (int s2 = nseg; s2 < 3; ++s2) c->tx_seg_lo[s2] = c->tx_seg_hi[s2] = 0;
    CU(c, upload(&c->d_tx, tx));
    c->wf = *w;
    c->wf.tx_pulse = nullptr;
    c->have_waveform = true;
    return RSP_OK;
}

static long matlab_round(double x) { return x >= 0 ? (long)std::floor(x + 0.5) : -(long)std::floor(-x + 0.5); }

// frac(cycles) as a 0.64 fixed-point number (the double carries 53 of the 64 bits)
static unsigned long long cycle_fraction_fix(double cycles) {
    double f = cycles - std::floor(cycles);                   // [0, 1)
    if (!(f >= 0.0 && f < 1.0)) f = 0.0;                        // NaN / rounding to 1.0
    return (unsigned long long)std::ldexp(f, 64 - 11) << 11;   // f * 2^53 is an exact integer below 2^53
}

// targets(k) -> what the synthesis kernel needs (fsf:55-66); targets outside the range window are dropped
static int make_synth_targets(const rsp_ctx* c, const rsp_target_in* targets, int n_targets, SynthTarget* out) {
    const double kPiD = 3.14159265358979323846;
    int n = 0;
    for (int i = 0; i < n_targets; ++i) {
        const double ts = 1.0 / c->wf.fs;                                                               // fsf:18
        const long d = matlab_round((2.0 * targets[i].range / c->wf.c) / ts);                           // fsf:55-56: round(delay / ts)
        if (!(d > 0 && d < c->N)) continue;                                                              // fsf:66
        SynthTarget t;
        t.delay = (int)d;
        t.amp = (float)std::sqrt(std::pow(10.0, targets[i].snr_db / 10.0) / c->wf.p_signal_unscaled);   // fsf:61-63
        t.dop_fix = cycle_fraction_fix(2.0 * targets[i].velocity / c->wf.wavelength * c->wf.prt);       // fsf:57-58
        t.steer_fix = cycle_fraction_fix(c->wf.element_spacing * std::sin(targets[i].elevation_deg * kPiD / 180.0) / c->wf.wavelength);   // fsf:165
        out[n++] = t;
    }
    return n;
}

static void fill_synth_args(const rsp_ctx* c, SynthArgs& a, float2* raw, const SynthTarget* d_tg, int n_tg, double noise_power, uint64_t seed);

static void launch_synth(rsp_ctx* c, float2* raw, const SynthTarget* d_tg, int n_tg, double noise_power, uint64_t seed, cudaStream_t s) {
    SynthArgs a;
    fill_synth_args(c, a, raw, d_tg, n_tg, noise_power, seed);
    static const bool staged_only = [] { const char* e = getenv("RSP_SYNTH"); return e && !strcmp(e, "staged"); }();
    Timed t(c, K_SYNTH);
    if (n_tg <= RSP_SYNTH_GATHER_T && !staged_only)
        synth_gather_kernel<<<dim3(c->C, c->P, (c->N + RSP_SYNTH_GATHER_CHUNK - 1) / RSP_SYNTH_GATHER_CHUNK), 256, 0, s>>>(a);
    else
        synth_kernel<<<dim3(c->C, c->P, (c->N + RSP_SYNTH_CHUNK - 1) / RSP_SYNTH_CHUNK), 256, 0, s>>>(a);
}

static void fill_synth_args(const rsp_ctx* c, SynthArgs& a, float2* raw, const SynthTarget* d_tg, int n_tg, double noise_power, uint64_t seed) {
    a.raw = raw;
    a.tx = c->d_tx; a.tg = d_tg; a.n_targets = n_tg;
    a.P = c->P; a.C = c->C; a.N = c->N;
    for (int i = 0; i < 3; ++i) { a.seg_lo[i] = c->tx_seg_lo[i]; a.seg_hi[i] = c->tx_seg_hi[i]; }
    a.noise_sigma = noise_power > 0 ? (float)std::sqrt(noise_power / 2.0) : 0.f;
    a.seed = seed;
    for (int r = 0; r < 10; ++r)
        a.round_key[r] = make_uint2((unsigned)seed + (unsigned)r * 0x9E3779B9u, (unsigned)(seed >> 32) + (unsigned)r * 0xBB67AE85u);
    a.tx_lo = c->N; a.tx_hi = 0;
    for (int i = 0; i < 3; ++i)
        if (a.seg_hi[i] > a.seg_lo[i]) { a.tx_lo = std::min(a.tx_lo, a.seg_lo[i]); a.tx_hi = std::max(a.tx_hi, a.seg_hi[i]); }
    if (a.tx_hi < a.tx_lo) a.tx_lo = a.tx_hi = 0;
}

int rsp_synthesize(rsp_ctx* c, const rsp_target_in* targets, int32_t n_targets, double noise_power, uint64_t seed, void* raw_dev_out) {
    if (!c || n_targets < 0 || (n_targets > 0 && !targets)) return fail(c, RSP_ERR_INVALID_ARG, "bad synthesis arguments");
    if (!c->have_waveform) return fail(c, RSP_ERR_NOT_READY, "rsp_set_waveform has not been called");
    CU(c, cudaSetDevice(c->prm.device));
    std::vector<SynthTarget> tg((size_t)std::max(n_targets, 1));
    tg.resize((size_t)make_synth_targets(c, targets, n_targets, tg.data()));
    if ((int)tg.size() > c->tg_cap) {
        if (c->d_tg) cudaFree(c->d_tg);
        c->d_tg = nullptr;
        c->tg_cap = std::max<int>(64, (int)tg.size());
        CU(c, dev_alloc(&c->d_tg, (size_t)c->tg_cap));
    }
    if (!tg.empty()) CU(c, cudaMemcpyAsync(c->d_tg, tg.data(), tg.size() * sizeof(SynthTarget), cudaMemcpyHostToDevice, c->stream));
    c->cur = &c->lanes[0];
    launch_synth(c, raw_dev_out ? static_cast<float2*>(raw_dev_out) : c->d_raw, c->d_tg, (int)tg.size(), noise_power, seed, c->stream);
    CU(c, cudaGetLastError());
    // the pageable target vector goes out of scope: make sure its copy has been consumed
    if (!tg.empty()) CU(c, cudaStreamSynchronize(c->stream));
    return RSP_OK;
}

// Pipelined frame path: S4 on the device into the lane's own cube, then S5..S9 on the same lane; returns at once.
int rsp_submit_targets(rsp_ctx* c, const rsp_target_in* targets, int32_t n_targets, double noise_power, uint64_t seed, int32_t slot) {
    if (!c || n_targets < 0 || (n_targets > 0 && !targets) || slot < 0 || slot >= c->slots)
        return fail(c, RSP_ERR_INVALID_ARG, "bad submit arguments");
    if (n_targets > RSP_MAX_FRAME_TARGETS)
        return fail(c, RSP_ERR_INVALID_ARG, "%d targets exceed RSP_MAX_FRAME_TARGETS=%d", n_targets, RSP_MAX_FRAME_TARGETS);
    if (!c->have_constants) return fail(c, RSP_ERR_NOT_READY, "rsp_upload_constants has not been called");
    if (!c->have_waveform) return fail(c, RSP_ERR_NOT_READY, "rsp_set_waveform has not been called");
    CU(c, cudaSetDevice(c->prm.device));
    if (int rcs = check_slot_free(c, slot)) return rcs;
    if (!c->d_tg_ring) {
        CU(c, dev_alloc(&c->d_tg_ring, (size_t)c->slots * RSP_MAX_FRAME_TARGETS));
        CU(c, cudaMallocHost(reinterpret_cast<void**>(&c->h_tg_ring), (size_t)c->slots * RSP_MAX_FRAME_TARGETS * sizeof(SynthTarget)));
    }
    const int l = slot % c->n_lanes;
    rsp_ctx::Lane& ln = c->lanes[l];
    if (!ln.raw) CU(c, cudaMalloc(reinterpret_cast<void**>(&ln.raw), (size_t)c->P * c->C * c->N * sizeof(float2)));
    SynthTarget* h = c->h_tg_ring + (size_t)slot * RSP_MAX_FRAME_TARGETS;
    SynthTarget* d = c->d_tg_ring + (size_t)slot * RSP_MAX_FRAME_TARGETS;
    const int n_tg = make_synth_targets(c, targets, n_targets, h);
    if (int rcf = lane_fork(c, l)) return rcf;
    if (n_tg > 0) CU(c, cudaMemcpyAsync(d, h, (size_t)n_tg * sizeof(SynthTarget), cudaMemcpyHostToDevice, ln.s));
    c->cur = &ln;
    float2* rdm = lane_rdm(c, l);
    if (!rdm) return fail(c, RSP_ERR_CUDA, "out of device memory for the lane's RDM");
    int rc;
    if (fused_synth_possible(c, n_tg)) {       // S4 generated inside the DBF: the cube is never materialised
        SynthArgs a;
        fill_synth_args(c, a, nullptr, d, n_tg, noise_power, seed);
        c->fused = &a;
        rc = enqueue_chain(c, ln.raw, rdm, slot, l);
        c->fused = nullptr;
    } else {
        launch_synth(c, ln.raw, d, n_tg, noise_power, seed, ln.s);
        rc = enqueue_chain(c, ln.raw, rdm, slot, l);
    }
    if (rc) return rc;
    launch_refine(c, slot, 1, ln.s);
    if (l == 0) { c->ran = true; c->rdm_in_ctx = true; }     // lane 0 works in the context's own buffers (rsp_get_beam / rsp_get_rdm)
    return finish_submit(c, slot, l);
}

int rsp_fetch_targets(rsp_ctx* c, int32_t slot, const rsp_cluster_params* cp, rsp_target* final_targets, int32_t cap,
                      int32_t* n_final, rsp_detection* dets, int32_t det_cap, int32_t* n_dets) {
    if (!c || !cp || !n_final || slot < 0 || slot >= c->slots) return fail(c, RSP_ERR_INVALID_ARG, "bad fetch arguments");
    CU(c, cudaSetDevice(c->prm.device));
    std::vector<rsp_detection> own;
    if (!dets) { own.resize((size_t)c->prm.max_detections); dets = own.data(); det_cap = (int32_t)own.size(); }
    int32_t n = 0;
    int rc = fetch_slot(c, slot, dets, det_cap, &n);
    if (n_dets) *n_dets = n;
    if (rc) return rc;
    std::vector<rsp_target> fin((size_t)std::max(n, 1));
    int32_t nf = 0, n1 = 0;
    rc = rsp_cluster(dets, n, cp, nullptr, &n1, fin.data(), &nf);
    if (rc) return fail(c, rc, "clustering failed");
    *n_final = nf;
    if (nf > cap) return fail(c, RSP_ERR_OVERFLOW, "%d targets exceed the caller's capacity %d", nf, cap);
    if (nf > 0 && !final_targets) return fail(c, RSP_ERR_INVALID_ARG, "null output");
    if (nf > 0) std::memcpy(final_targets, fin.data(), (size_t)nf * sizeof(rsp_target));
    return RSP_OK;
}

// Many independent frames through the pipelined frame path with the host half of each frame (sorting the detection list
// into the reference's find order, S10 / S11 clustering) on a pool of worker threads: the submitting thread only enqueues
// frames `depth` ahead and copies finished detection lists out of pinned memory, so dense frames (64 targets: ~1600
// detections, 110 us of clustering each) no longer hold the GPU back.
int rsp_process_frames(rsp_ctx* c, const rsp_target_in* targets, const int32_t* n_targets, int32_t n_frames, double noise_power,
                       const uint64_t* seeds, const rsp_cluster_params* cp, int32_t depth, int32_t host_threads,
                       rsp_target* final_targets, int32_t cap, int32_t* n_final,
                       rsp_detection* dets, int64_t det_cap_total, int64_t* det_offsets) {
    if (!c || !cp || !n_targets || !seeds || !n_final || n_frames < 0 || cap < 0 || (cap > 0 && !final_targets) || (dets && !det_offsets))
        return fail(c, RSP_ERR_INVALID_ARG, "bad rsp_process_frames arguments");
    if (n_frames == 0) { if (det_offsets) det_offsets[0] = 0; return RSP_OK; }
    CU(c, cudaSetDevice(c->prm.device));
    depth = std::max(1, std::min(depth > 0 ? depth : 2 * c->n_lanes, c->slots));
    const int n_workers = std::max(1, std::min({host_threads > 0 ? host_threads : 4, 32, (n_frames + 3) / 4}));
    std::vector<int64_t> tg_off((size_t)n_frames + 1, 0);
    for (int i = 0; i < n_frames; ++i) {
        if (n_targets[i] < 0 || (n_targets[i] > 0 && !targets)) return fail(c, RSP_ERR_INVALID_ARG, "bad target count of frame %d", i);
        tg_off[i + 1] = tg_off[i] + n_targets[i];
    }
    struct Job { int frame; rsp_detection* d; int n; std::vector<rsp_detection> own; };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> queue;
    bool closed = false;
    int worker_rc = RSP_OK, worker_frame = -1, worker_nf = 0;
    auto work = [&]() {
        std::vector<std::pair<uint64_t, int>> keys;
        std::vector<rsp_detection> tmp;
        std::vector<rsp_target> fin;
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return closed || !queue.empty(); });
                if (queue.empty()) return;
                job = std::move(queue.front());
                queue.pop_front();
            }
            rsp_detection* d = job.d ? job.d : job.own.data();
            int32_t nf = 0, n1 = 0;
            int rc = RSP_OK;
            try {
                sort_detections_keyed(d, job.n, keys, tmp);
                fin.resize((size_t)std::max(job.n, 1));
                rc = rsp_cluster(d, job.n, cp, nullptr, &n1, fin.data(), &nf);
            } catch (...) {                           // out of host memory
                rc = RSP_ERR_INVALID_ARG;
            }
            n_final[job.frame] = nf;
            if (!rc && nf > cap) rc = RSP_ERR_OVERFLOW;
            if (!rc && nf > 0) std::memcpy(final_targets + (size_t)job.frame * cap, fin.data(), (size_t)nf * sizeof(rsp_target));
            if (rc) {
                std::lock_guard<std::mutex> lk(mu);
                if (!worker_rc) { worker_rc = rc; worker_frame = job.frame; worker_nf = nf; }
            }
        }
    };
    // No exception may cross the C ABI: a thread that cannot be created just leaves fewer workers (none: the calling thread
    // sorts and clusters after the last fetch).
    std::vector<std::thread> pool;
    try {
        for (int w = 0; w < n_workers; ++w) pool.emplace_back(work);
    } catch (...) {
    }
    int rc = RSP_OK, submitted = 0, fetched = 0;
    int64_t off = 0;
    if (det_offsets) det_offsets[0] = 0;
    for (int i = 0; i < n_frames + depth && !rc; ++i) {
        if (i >= depth) {
            const int j = i - depth;
            Job job;
            job.frame = j; job.d = nullptr; job.n = 0;
            int32_t n = 0;
            // peek at the count first (the slot's event has to be waited for anyway), then copy into the caller's flat buffer or
            // into a vector of the job's own
            CU(c, cudaEventSynchronize(c->slot_done[j % c->slots]));
            n = c->h_slot_count[j % c->slots];
            // (errors before the fetch leave the slot to the drain loop below, which clears it)
            if (n > c->prm.max_detections) { rc = fail(c, RSP_ERR_OVERFLOW, "frame %d: %d detections exceed max_detections=%d", j, n, c->prm.max_detections); break; }
            if (dets) {
                if (off + n > det_cap_total) { rc = fail(c, RSP_ERR_OVERFLOW, "frame %d: detections exceed the caller's capacity %lld", j, (long long)det_cap_total); break; }
                job.d = dets + off;
            } else {
                job.own.resize((size_t)std::max(n, 1));
            }
            rc = fetch_slot(c, j % c->slots, job.d ? job.d : job.own.data(), std::max(n, 1), &n, false);
            ++fetched;
            if (rc) break;
            job.n = n;
            off += n;
            if (det_offsets) det_offsets[j + 1] = off;
            {
                std::lock_guard<std::mutex> lk(mu);
                queue.push_back(std::move(job));
            }
            cv.notify_one();
        }
        if (i < n_frames) {
            rc = rsp_submit_targets(c, targets ? targets + tg_off[i] : nullptr, n_targets[i], noise_power, seeds[i], i % c->slots);
            if (!rc) ++submitted;
        }
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        closed = true;
    }
    cv.notify_all();
    if (pool.empty()) work();
    for (auto& t : pool) t.join();
    if (rc) {                                     // keep the submit / fetch pairing of the ring intact for the next caller
        const std::string msg = c->err;
        std::vector<rsp_detection> sink((size_t)c->prm.max_detections);
        for (int j = fetched; j < submitted; ++j) { int32_t n = 0; fetch_slot(c, j % c->slots, sink.data(), (int32_t)sink.size(), &n, false); }
        c->err = msg;
        return rc;
    }
    if (worker_rc == RSP_ERR_OVERFLOW) return fail(c, worker_rc, "frame %d: %d targets exceed the caller's capacity %d", worker_frame, worker_nf, cap);
    if (worker_rc) return fail(c, worker_rc, "clustering failed for frame %d", worker_frame);
    return RSP_OK;
}

int rsp_process_targets(rsp_ctx* c, const rsp_target_in* targets, int32_t n_targets, double noise_power, uint64_t seed,
                        const rsp_cluster_params* cp, rsp_target* final_targets, int32_t cap, int32_t* n_final,
                        rsp_detection* dets, int32_t det_cap, int32_t* n_dets) {
    if (!c || !cp || !n_final) return fail(c, RSP_ERR_INVALID_ARG, "null argument");
    int rc = rsp_synthesize(c, targets, n_targets, noise_power, seed, nullptr);
    if (rc) return rc;
    std::vector<rsp_detection> own;
    if (!dets) { own.resize((size_t)c->prm.max_detections); dets = own.data(); det_cap = (int32_t)own.size(); }
    int32_t n = 0;
    rc = rsp_process_cpi(c, c->d_raw, RSP_LAYOUT_PCN, RSP_C64, RSP_MEM_DEVICE, nullptr, RSP_MEM_DEVICE, dets, det_cap, &n);
    if (n_dets) *n_dets = n;
    if (rc) return rc;
    std::vector<rsp_target> fin((size_t)std::max(n, 1));
    int32_t nf = 0, n1 = 0;
    rc = rsp_cluster(dets, n, cp, nullptr, &n1, fin.data(), &nf);
    if (rc) return fail(c, rc, "clustering failed");
    *n_final = nf;
    if (nf > cap) return fail(c, RSP_ERR_OVERFLOW, "%d targets exceed the caller's capacity %d", nf, cap);
    if (nf > 0 && !final_targets) return fail(c, RSP_ERR_INVALID_ARG, "null output");
    std::memcpy(final_targets, fin.data(), (size_t)nf * sizeof(rsp_target));
    return RSP_OK;
}

int rsp_set_profiling(rsp_ctx* c, int enable) {
    if (!c) return RSP_ERR_INVALID_ARG;
    c->profiling = enable != 0;
    return RSP_OK;
}

int rsp_get_kernel_times(rsp_ctx* c, rsp_kernel_times* out) {
    if (!c || !out) return RSP_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->prm.device));
    if (int rcs = rsp_synchronize(c)) return rcs;       // spans of the pipelined paths sit on the other lanes' streams
    std::memset(out, 0, sizeof *out);
    out->n = K_NCLASS;
    for (int i = 0; i < K_NCLASS; ++i) out->name[i] = kKernelNames[i];
    for (auto& sp : c->spans) {
        float ms = 0.f;
        CU(c, cudaEventElapsedTime(&ms, sp.a, sp.b));
        out->total_ms[sp.cls] += ms;
        out->launches[sp.cls] += 1;
        c->event_pool.push_back(sp.a);
        c->event_pool.push_back(sp.b);
    }
    c->spans.clear();
    return RSP_OK;
}

int rsp_get_fused_trace(rsp_ctx* c, int64_t* dst, int32_t cap_ctas, int32_t* n_ctas) {
    if (!c || !dst || !n_ctas) return RSP_ERR_INVALID_ARG;
    *n_ctas = 0;
    if (!c->d_fused_dbg || !c->dbf_pc_ok) return RSP_OK;
    const int n = std::min(cap_ctas, c->P * c->B);
    CU(c, cudaSetDevice(c->prm.device));
    CU(c, rsp_synchronize(c) == RSP_OK ? cudaSuccess : cudaErrorUnknown);
    CU(c, cudaMemcpy(dst, c->d_fused_dbg, (size_t)n * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    *n_ctas = n;
    return RSP_OK;
}

int rsp_get_info(const rsp_ctx* c, rsp_info* info) {
    if (!c || !info) return RSP_ERR_INVALID_ARG;
    info->n_gates_total = c->G;
    info->fft_len_medium = c->med.L; info->fft_len_long = c->lng.L;
    info->blocks_medium = c->med.nblk; info->blocks_long = c->lng.nblk + c->lngx[0].nblk + c->lngx[1].nblk;
    info->kernels_per_cpi = c->have_constants ? kernels_per_cpi(c) : 0;
    info->algorithmic_bytes_per_cpi = 8LL * c->P * c->N * c->C + 8LL * c->B * c->P * c->G;
    info->launches_total = c->launches;
    info->lanes = c->n_lanes;
    info->graph_launches = (int32_t)c->graph_launches;
    return RSP_OK;
}

}  // extern "C"
