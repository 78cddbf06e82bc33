/*
 * rsp.h -- thin C ABI of librsp.so: the B200-native per-frame phased-array chain
 *          DBF -> pulse compression -> MTD -> GOCA-CFAR -> monopulse (+ host clustering).
 *
 * Drop-in boundary for the reference's MATLAB hot path (all paths under
 * /root/reference/Simulation/):
 *
 *   fun_process_single_frame.m:13
 *       final_targets = fun_process_single_frame(targets, config, cfar_params,
 *                                                cluster_params, precomputed_data, frame_idx)
 *   process_stage2_mtd.m:1
 *       [MTD_results, PC_results] = process_stage2_mtd(iq_data, angle, config)
 *
 * A MEX gateway (mex/ in this repo) or the ctypes binding (rsp_b200/_abi.py) unpacks the
 * MATLAB structs into the POD structs below and calls these entry points.  Plain pointers
 * and sizes only; no exceptions cross this boundary; every function returns an rsp_status
 * (0 = OK, negative = error, text via rsp_last_error).
 *
 * Index conventions: detection indices (v_idx, r_idx, pair_idx) are 1-based exactly like the
 * reference's all_raw_detections rows [v, r, pair, amp] (fun_process_single_frame.m:220).
 *
 * Layouts (complex64 = interleaved {float re, im}; complex128 likewise with doubles):
 *   RSP_LAYOUT_PCN      raw[p][c][n]   device-native: range sample fastest, then channel, then pulse
 *   RSP_LAYOUT_MATLAB   MATLAB [P,N,C] column-major, i.e. raw[c][n][p] (pulse fastest), the
 *                       in-memory order of raw_iq_data_noise (fun_process_single_frame.m:81)
 *   range-Doppler map   rdm[b][g][v]   Doppler fastest == MATLAB rdm_13beam(v,g,b) byte order
 *                       (fun_process_single_frame.m:131-136)
 *   pulse compression   pc[b][g][p]    == MATLAB pc_results_13beam(p,g,b) byte order (:101-127)
 *
 * The context is not thread-safe: one context per host thread / CUDA stream.
 */
#ifndef RSP_H_
#define RSP_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSP_ABI_VERSION 1
#define RSP_MAX_CHANNELS 32
#define RSP_MAX_BEAMS 16

typedef enum {
    RSP_OK = 0,
    RSP_ERR_INVALID_ARG = -1,
    RSP_ERR_UNSUPPORTED = -2,     /* shape outside what the kernels implement */
    RSP_ERR_CUDA = -3,            /* CUDA runtime failure, text in rsp_last_error */
    RSP_ERR_NO_DEVICE = -4,       /* no CUDA device: there is NO CPU fallback */
    RSP_ERR_OVERFLOW = -5,        /* detection list exceeded max_detections (never truncated silently) */
    RSP_ERR_NOT_READY = -6        /* constants not uploaded / stage not run yet */
} rsp_status;

typedef enum { RSP_LAYOUT_PCN = 0, RSP_LAYOUT_MATLAB = 1 } rsp_layout;
typedef enum { RSP_MEM_HOST = 0, RSP_MEM_DEVICE = 1 } rsp_mem;
typedef enum { RSP_C64 = 0, RSP_C128 = 1 } rsp_dtype;

typedef struct { float re, im; } rsp_c64;
typedef struct { double re, im; } rsp_c128;

/* Shape + detector parameters.  Field sources:
 *   config.Sig_Config.{channel_num, beam_num, prtNum, point_PRT, point_prt_segments}
 *       (fun_process_single_frame.m:17,47,92,175; main_simulate_echoes_with_array_v8_3.m:71-84)
 *   precomputed_data.{seg_start_*, N_gate_*, fir_delay}      (fun_process_single_frame.m:25,30-36)
 *   cfar_params.{T_CFAR, guardCells_*, refCells_*}           (fun_process_single_frame.m:177-179)
 */
typedef struct {
    int32_t abi_version;          /* RSP_ABI_VERSION */
    int32_t n_channels;           /* C */
    int32_t n_beams;              /* B */
    int32_t n_pulses;             /* P (prtNum) */
    int32_t n_samples;            /* N (point_PRT) */
    int32_t seg_start[3];         /* 1-based first sample of the narrow / medium / long segment */
    int32_t n_gates[3];           /* gates kept from each segment (sum = G) */
    int32_t fir_delay;            /* circshift of the narrow FIR output */
    float   t_cfar;
    int32_t guard_r, guard_v, ref_r, ref_v;
    int32_t max_detections;       /* capacity of the per-CPI detection list */
    int32_t monopulse_complex;    /* 0: amplitude ratio (fun_process_single_frame.m:282-285),
                                     1: complex ratio, real part (main_plot_snr_vs_angle_error.m:454-461) */
    int32_t device;               /* CUDA device ordinal */
} rsp_params;

/* Filters, windows, axes and calibration tables: precomputed_data
 * (main_simulate_echoes_with_array_v8_3.m:121-183), all double precision as MATLAB holds them.
 * The matched filters are passed as TIME-DOMAIN taps (precomputed_data.MF_medium_win /
 * MF_long_win, v8_3:146,148); the library builds its own block spectra from them. */
typedef struct {
    const rsp_c128* dbf_weights;  /* [B][C] row-major = DBF_coeffs_data_C (used conjugated, fsf:95) */
    const double*   fir;          /* MF_narrow taps */
    int32_t         n_fir;
    const rsp_c128* mf_medium;    /* time-domain matched filter, medium pulse */
    int32_t         n_mf_medium;
    const rsp_c128* mf_long;      /* time-domain matched filter, long pulse */
    int32_t         n_mf_long;
    const double*   mtd_win;      /* [P] MTD_win */
    const double*   range_axis;   /* [G] */
    const double*   velocity_axis;/* [P] */
    double          delta_r, delta_v;
    const double*   beam_angles_deg; /* [B] */
    const double*   k_slopes;     /* [B-1] k_slopes_LUT */
} rsp_constants;

/* One CFAR detection after S9 (fun_process_single_frame.m:220 and :293-297). */
typedef struct {
    int32_t v_idx, r_idx, pair_idx;   /* 1-based */
    float   power;                    /* S(v,r) = |rdm_A| + |rdm_B| */
    double  range, velocity, angle;
} rsp_detection;

/* One clustered target (fun_process_single_frame.m:340-350, :393-406). */
typedef struct { double range, velocity, angle, power; } rsp_target;

typedef struct {
    double max_range_sep, max_vel_sep, max_angle_sep;   /* cluster_params, fsf:328,381 */
} rsp_cluster_params;

typedef struct rsp_ctx rsp_ctx;

/* ---- lifecycle ---- */
int rsp_abi_version(void);
int rsp_device_count(void);
int rsp_create(const rsp_params* params, rsp_ctx** out);
void rsp_destroy(rsp_ctx* ctx);
const char* rsp_last_error(const rsp_ctx* ctx);          /* ctx may be NULL: last create() error */
int rsp_upload_constants(rsp_ctx* ctx, const rsp_constants* k);
/* Run on the caller's stream (a cudaStream_t handle; 0 is the legacy default stream, as PyTorch's
 * default stream reports).  RSP_STREAM_OWN switches back to the context's own non-blocking stream. */
#define RSP_STREAM_OWN ((void*)(intptr_t)-1)
int rsp_set_stream(rsp_ctx* ctx, void* cuda_stream);
int rsp_synchronize(rsp_ctx* ctx);

/* ---- the hot path: S5 -> S9 on one CPI (fun_process_single_frame.m:90-146) ----
 * raw: C*N*P complex samples in `layout`/`dtype`, in host or device memory.
 * rdm_out: optional (may be NULL) destination for the complex64 range-Doppler map [B][G][P]
 *          in `rdm_mem` space; NULL keeps it in the context (rsp_get_rdm).
 * dets: host array of capacity det_cap; *n_dets receives the count.  The list is returned in the
 *       reference's order (pair ascending, then range, then Doppler: MATLAB find, fsf:215-221). */
int rsp_process_cpi(rsp_ctx* ctx, const void* raw, rsp_layout layout, rsp_dtype dtype, rsp_mem raw_mem,
                    void* rdm_out, rsp_mem rdm_mem,
                    rsp_detection* dets, int32_t det_cap, int32_t* n_dets);

/* ---- device-resident CPI stream (throughput path; no host<->device traffic, asynchronous) ----
 * Processes n_cpi cubes: cube i is read from raw_dev + (i % raw_pool) * C*N*P complex64 (RSP_LAYOUT_PCN),
 * its RDM is written to rdm_dev + (i % rdm_pool) * B*G*P complex64, its detections to slot
 * (first_slot + i) of the context's device detection ring.  Returns after enqueueing.
 * Consecutive CPIs run concurrently on the context's lanes (rsp_info.lanes), so a caller ring must hold at least that many
 * maps, and a multiple of it when n_cpi > rdm_pool (else RSP_ERR_INVALID_ARG); rdm_dev = NULL keeps one map per lane. */
int rsp_stream_enqueue(rsp_ctx* ctx, const void* raw_dev, int32_t raw_pool, void* rdm_dev, int32_t rdm_pool,
                       int32_t n_cpi, int32_t first_slot);
/* Pipelined host-input path: enqueue the host->device copy of one PCN complex64 cube (pinned host memory
 * for true overlap) and its chain on lane (slot % lanes), and return at once; the copy of cube i+1 then
 * overlaps the kernels of cube i.  Collect with rsp_stream_fetch(ctx, slot, ...), which waits for that slot. */
/* Submit / fetch pairing: a slot must be fetched (rsp_stream_fetch / rsp_fetch_targets) before it is submitted again;
 * resubmitting an unfetched slot returns RSP_ERR_INVALID_ARG (its host cube / target block may still be in flight). */
int rsp_submit_cpi(rsp_ctx* ctx, const void* raw_host, void* rdm_dev /* may be NULL */, int32_t slot);
int rsp_stream_slots(const rsp_ctx* ctx);                 /* capacity of the detection ring (CPIs) */
/* Raw device pointers to the ring: counts[slot] (int32) and records[slot][max_detections] (unsorted). */
int rsp_stream_device_buffers(rsp_ctx* ctx, void** counts_dev, void** records_dev);
/* Copy one slot to the host and sort it into the reference order. */
int rsp_stream_fetch(rsp_ctx* ctx, int32_t slot, rsp_detection* dets, int32_t det_cap, int32_t* n_dets);
/* Sort an unsorted record list (e.g. gathered from another rank) into the reference order. */
int rsp_sort_detections(rsp_detection* dets, int32_t n);

/* ---- intermediates of the last rsp_process_cpi (parity / process_stage2_mtd outputs) ---- */
int rsp_get_beam(rsp_ctx* ctx, rsp_c64* dst_host);        /* [P][B][N]  iq_data_13beam, fsf:92-97 */
int rsp_get_pc(rsp_ctx* ctx, rsp_c64* dst_host);          /* [P][B][G]  device-native order */
int rsp_get_rdm(rsp_ctx* ctx, rsp_c64* dst_host);         /* [B][G][P]  rdm_13beam */
int rsp_get_amp(rsp_ctx* ctx, float* dst_host);           /* [B][G][P]  |rdm| */

/* ---- S10 + S11 on the host (fun_process_single_frame.m:302-407), order-dependent BFS ---- */
int rsp_cluster(const rsp_detection* dets, int32_t n_dets, const rsp_cluster_params* cp,
                rsp_target* stage1, int32_t* n_stage1,      /* capacity n_dets each; stage1 may be NULL */
                rsp_target* final_targets, int32_t* n_final);

/* ---- S5 -> S11: the whole fun_process_single_frame after echo synthesis ---- */
int rsp_process_frame(rsp_ctx* ctx, const void* raw, rsp_layout layout, rsp_dtype dtype, rsp_mem raw_mem,
                      const rsp_cluster_params* cp, rsp_target* final_targets, int32_t cap, int32_t* n_final);

/* ---- process_stage2_mtd.m:1 on an already beamformed + range-gated cube ----
 * The context is created with rsp_create (n_gates = the three segment lengths, e.g. 228/723/2453;
 * n_samples = their sum; the other chain parameters are unused) and configured once with the three
 * reference pulses (debug_simulated_data_processing_v2.m:309-317), the Doppler window and the width of
 * the zero-velocity notch (MTD_0v_num, main_test_with_simulated_data.m:124).
 * iq: MATLAB [P,G,B] column-major complex (dtype), host memory.  Outputs MATLAB-ordered [P,G,B]
 * complex128: mtd_out = Doppler map, pc_out = pulse-compressed cube.  The reference's callee
 * fun_MTD_produce is not shipped; the semantics are specified by this repo:
 *   pc(p, g, b)  = sum_k iq(p, g + k, b) * conj(pulse_s(k)),  g and g + k inside segment s;
 *   mtd(:, g, b) = fftshift(fft(pc(:, g, b) .* win));  rows within +-zero_vel_bins of zero Doppler = 0. */
typedef struct {
    const rsp_c128* pulse[3];     /* reference pulse of the narrow / medium / long segment */
    int32_t         n_pulse[3];
    const double*   mtd_win;      /* [P], NULL = rectangular */
    int32_t         zero_vel_bins;/* 0 = no notch */
} rsp_stage2_config;
int rsp_stage2_configure(rsp_ctx* ctx, const rsp_stage2_config* cfg);
int rsp_stage2_mtd(rsp_ctx* ctx, const void* iq, rsp_dtype dtype, rsp_c128* mtd_out, rsp_c128* pc_out);

/* ---- the per-segment 1-D range CFAR of the real-data path (debug_simulated_data_processing_v2.m:419-511:
 * local_execute_cfar / executeCFAR_2D / Function_CFAR1D_sub) ----
 * Each of the three pulse segments of a beam's Doppler map is detected on its own along range: two reference windows of
 * `ref_cells` gates `save_cells` away from the cell, a window that leaves the segment is replaced by the other one,
 * level = greater (method 0, GOCA) or smaller (method 1, SOCA) of the two means, flag = amplitude >= level * t_cfar;
 * Doppler rows within +-zero_vel_bins of row round(V/2)+1 are not tested (flag 0, threshold 0).
 * Arrays are [beam][gate][Doppler], Doppler fastest == MATLAB (Doppler, gate) maps per beam. */
typedef struct {
    int32_t ref_cells, save_cells;   /* config.cfar.refCells_R, saveCells_R */
    int32_t method;                  /* config.cfar.CFARmethod_R: 0 = greatest-of, 1 = smallest-of */
    int32_t zero_vel_bins;           /* config.cfar.MTD_0v_num */
    float   t_cfar;                  /* config.cfar.T_CFAR */
    int32_t seg_len[3];              /* gates of the narrow / medium / long segment (config.Sig_Config.point_prt(2:4)) */
} rsp_cfar1d_params;
/* Any amplitude map(s) in host memory (mtd_amplitude_map = abs(MTD) of n_beams beams). */
int rsp_cfar1d(int32_t device, const float* amp_host, int32_t n_doppler, int32_t n_gates, int32_t n_beams,
               const rsp_cfar1d_params* p, uint8_t* flags_host, float* thresholds_host /* may be NULL */);
/* The Doppler maps the last rsp_stage2_mtd left in the context (all beams, |MTD| computed on the device). */
int rsp_stage2_cfar(rsp_ctx* ctx, const rsp_cfar1d_params* p, uint8_t* flags_host, float* thresholds_host /* may be NULL */);

/* ---- S4 + S4.1 on the device (fun_process_single_frame.m:47-88): echo synthesis + noise ----
 * Fills a device-resident PCN complex64 cube: for every target the delayed transmit pulse scaled by the
 * SNR amplitude, the pulse-to-pulse Doppler phasor and the per-channel steering phasor (fsf:55-74), plus
 * (noise_power > 0) complex white Gaussian noise of that power from a counter-based Philox4x32-10
 * generator keyed by `seed` (the reference draws from MATLAB's global randn stream, fsf:84-85, which no
 * other generator can reproduce; parity of the echoes is exact, of the noise statistical).
 * rsp_set_waveform uploads precomputed_data.tx_pulse and the scalars the synthesis needs. */
typedef struct { double range, velocity, elevation_deg, snr_db; } rsp_target_in;   /* targets(k), fsf:51-61 */
typedef struct {
    const rsp_c128* tx_pulse;     /* [N] precomputed_data.tx_pulse (v8_3:132-138) */
    double c, fs, wavelength, prt, element_spacing;   /* config.Sig_Config.*, config.Array.* */
    double p_signal_unscaled;     /* precomputed_data.P_signal_unscaled (v8_3:139) */
} rsp_waveform;
int rsp_set_waveform(rsp_ctx* ctx, const rsp_waveform* w);
int rsp_synthesize(rsp_ctx* ctx, const rsp_target_in* targets, int32_t n_targets, double noise_power, uint64_t seed,
                   void* raw_dev_out /* NULL = the context's own cube */);
/* fun_process_single_frame end to end on the device: S4 (rsp_synthesize into the context's cube), S5..S9,
 * host clustering.  Only the target list crosses the bus. */
int rsp_process_targets(rsp_ctx* ctx, const rsp_target_in* targets, int32_t n_targets, double noise_power, uint64_t seed,
                        const rsp_cluster_params* cp, rsp_target* final_targets, int32_t cap, int32_t* n_final,
                        rsp_detection* dets /* may be NULL */, int32_t det_cap, int32_t* n_dets /* may be NULL */);

/* Pipelined frame path (the Monte-Carlo sweep and the multi-frame tracker submit many independent frames):
 * rsp_submit_targets enqueues S4 into the lane's own cube and S5..S9 behind it on lane (slot % lanes) and
 * returns at once, so frame i+1 is synthesised while frame i is processed; rsp_fetch_targets waits for that
 * slot, sorts its detections into the reference order and clusters them (S10/S11) on the host. */
#define RSP_MAX_FRAME_TARGETS 64
int rsp_submit_targets(rsp_ctx* ctx, const rsp_target_in* targets, int32_t n_targets, double noise_power, uint64_t seed,
                       int32_t slot);
int rsp_fetch_targets(rsp_ctx* ctx, int32_t slot, const rsp_cluster_params* cp, rsp_target* final_targets, int32_t cap,
                      int32_t* n_final, rsp_detection* dets /* may be NULL */, int32_t det_cap, int32_t* n_dets /* may be NULL */);

/* A whole block of independent frames (Monte-Carlo trials, the frame loop of main_simulate_echoes_with_array_v8_3.m:200-248)
 * in one call: frame i = targets[sum(n_targets[0..i)) ...], seeds[i]; frames are submitted `depth` ahead of the fetch
 * (0 = 2 x lanes) on the ring slots i % slots, and the host half of every frame -- sorting its detections into the reference's
 * find order, S10 / S11 clustering (fsf:302-407) -- runs on `host_threads` worker threads (0 = 4), so dense frames do not hold
 * the GPU back.  Results per frame as from rsp_fetch_targets: final_targets[i * cap ...], n_final[i]; optionally every frame's
 * sorted detections packed back to back in dets[det_offsets[i] .. det_offsets[i + 1]) (dets NULL: not returned).  More targets
 * than cap in a frame, or more detections than det_cap_total in all, is RSP_ERR_OVERFLOW.  No pipelined submission of this
 * context may be outstanding when it is called. */
int rsp_process_frames(rsp_ctx* ctx, const rsp_target_in* targets, const int32_t* n_targets, int32_t n_frames, double noise_power,
                       const uint64_t* seeds, const rsp_cluster_params* cp, int32_t depth, int32_t host_threads,
                       rsp_target* final_targets, int32_t cap, int32_t* n_final,
                       rsp_detection* dets /* may be NULL */, int64_t det_cap_total, int64_t* det_offsets /* [n_frames + 1] when dets */);

/* ---- introspection ---- */
typedef struct {
    int32_t n_gates_total;        /* G */
    int32_t fft_len_medium, fft_len_long, blocks_medium, blocks_long;
    int32_t kernels_per_cpi;      /* launches rsp_process_cpi / rsp_stream_enqueue issue per CPI */
    int64_t algorithmic_bytes_per_cpi;   /* 8*P*N*C + 8*B*P*G */
    int64_t launches_total;       /* kernels launched by this context so far */
    int32_t lanes;                /* concurrent CPI lanes of the stream / pipelined paths (RSP_LANES) */
    int32_t graph_launches;       /* rsp_stream_enqueue batches replayed as one CUDA graph (launches_total counts their kernels) */
} rsp_info;
int rsp_get_info(const rsp_ctx* ctx, rsp_info* info);
/* Measurement aid of the fused DBF + pulse-compression kernel: when the context was created under RSP_FUSED_DEBUG=<flags>,
 * the last launch recorded for every CTA {start, cluster up, DBF done, lines complete, round 0..2 done} in globaltimer ns
 * and its SM id; dst receives 8 int64 per CTA (CTA = pulse * B + beam).  *n_ctas = 0 when tracing is off. */
int rsp_get_fused_trace(rsp_ctx* ctx, int64_t* dst, int32_t cap_ctas, int32_t* n_ctas);

/* ---- per-kernel device timing (bench.py's roofline leg) ----
 * While enabled, every kernel of subsequently enqueued CPIs is bracketed by CUDA events on the
 * context's stream.  rsp_get_kernel_times synchronises, returns for each kernel class its name,
 * accumulated device milliseconds and launch count, and clears the accumulators. */
#define RSP_MAX_KERNEL_CLASSES 12
typedef struct {
    int32_t n;
    const char* name[RSP_MAX_KERNEL_CLASSES];
    double total_ms[RSP_MAX_KERNEL_CLASSES];
    int64_t launches[RSP_MAX_KERNEL_CLASSES];
} rsp_kernel_times;
int rsp_set_profiling(rsp_ctx* ctx, int enable);
int rsp_get_kernel_times(rsp_ctx* ctx, rsp_kernel_times* out);

#ifdef __cplusplus
}
#endif
#endif /* RSP_H_ */
