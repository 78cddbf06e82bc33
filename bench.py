#!/usr/bin/env python
"""Benchmark of the per-frame chain DBF -> PC -> MTD -> CFAR -> monopulse (BASELINE.json metric:
CPIs/s per B200 and at 1/2/4/8 GPUs; fraction of the HBM roofline; CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl reference]

One "step" = one pass of the hot path (S5..S9) over a batch of `--cpis-per-step` synthetic CPIs that
are already resident in HBM; CPIs are independent, so ranks share nothing but a final gather of the
detection lists (NCCL all_gather, weak scaling).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cpis_per_sec"
UNIT = "CPI/s"
FALLBACK_HBM_GBS = 6650.0       # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

CONFIG_DESC = {
    "cfg1": "single frame 16 ch x 13 beams x 32 pulses x 4096 samples (BASELINE configs[0])",
    "cfg2": "single-frame full chain DBF->PC->MTD->CFAR->monopulse, 16 ch x 8 beams x 64 pulses x 8192 bins "
            "(BASELINE configs[1])",
    "cfg3": "CPI stream 32 ch x 16 beams x 128 pulses x 16384 bins (BASELINE configs[2])",
    "native": "reference literal shape 16 ch x 13 beams x 332 pulses x 5819 samples",
}


def load_oracle():
    """The oracle may be executed only by the cpu_baseline leg and by --impl reference."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("rsp_oracle", os.path.join(ROOT, "oracle", "rsp_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rsp_oracle"] = mod
    spec.loader.exec_module(mod)
    return mod


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 50 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_pool(rsp, config, pd, n_cubes: int, seed0: int):
    """Distinct synthetic CPIs (targets T3 of SURVEY.md 8(d) + unit complex noise), complex64 PCN."""
    v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
    targets = [dict(Range=900.0, Velocity=0.15 * v_max, ElevationAngle=-5.0, SNR_dB=20.0),
               dict(Range=3000.0, Velocity=-0.10 * v_max, ElevationAngle=8.2, SNR_dB=10.0),
               dict(Range=8000.0, Velocity=0.05 * v_max, ElevationAngle=15.0, SNR_dB=10.0)]
    clean = rsp.synthesize_echo(targets, config, pd)
    cubes = []
    for i in range(n_cubes):
        cubes.append(rsp.add_noise(clean, np.random.default_rng(seed0 + i)).astype(np.complex64))
    return np.stack(cubes)


def _all_threads(n: int):
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host thread."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def cpu_reference_time(name: str, n_cpi: int, workers: int):
    """Time the fp64 NumPy/SciPy oracle (S5..S9, no clustering) on n_cpi cubes; best-effort all cores."""
    o = load_oracle()
    cfg, pre, raw = o.make_cube(name, 0)
    raw = raw.astype(np.complex128)
    o.process_cube(raw, cfg, pre, workers=workers, keep=False, cluster=False)        # warm-up (FFT plans)
    t0 = time.perf_counter()
    ndet = 0
    for _ in range(n_cpi):
        ndet = len(o.process_cube(raw, cfg, pre, workers=workers, keep=False, cluster=False).raw_detections)
    dt = time.perf_counter() - t0
    return n_cpi / dt, dt, ndet


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation of the path.  The reference is MATLAB and
    neither MATLAB nor Octave exists offline, so this is the fp64 NumPy/SciPy port (kind "port")."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    name = args.config
    o = load_oracle()
    cfg, pre, raw = o.make_cube(name, 0)
    raw = raw.astype(np.complex128)
    per_step = 1
    with _all_threads(cores):
        for _ in range(max(args.warmup, 1)):
            o.process_cube(raw, cfg, pre, workers=cores, keep=False, cluster=False)
        t0 = time.perf_counter()
        for _ in range(args.steps * per_step):
            o.process_cube(raw, cfg, pre, workers=cores, keep=False, cluster=False)
        dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CONFIG_DESC[name], "name": name, "cpis_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} CPIs of {name}, fp64 NumPy/SciPy oracle S5..S9, scipy.fft workers={cores}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is MATLAB; MATLAB/Octave unavailable offline, NumPy/SciPy port of fun_process_single_frame.m S5..S9 timed instead",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=list(CONFIG_DESC))
    ap.add_argument("--cpis-per-step", type=int, default=0)
    ap.add_argument("--pool", type=int, default=0, help="distinct input CPIs resident in HBM")
    ap.add_argument("--rdm-pool", type=int, default=0, help="distinct RDM output buffers")
    ap.add_argument("--e2e-cpis", type=int, default=24)
    ap.add_argument("--cpu-cpis", type=int, default=0, help="CPIs in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3          # timing rules: at least 3 warm-up steps

    import torch
    import torch.distributed as dist
    import rsp_b200 as rsp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    name = args.config
    config, cfar_params, cluster_params = rsp.named_config(name)
    pd = rsp.build_precomputed_data(config)
    chain = rsp.RadarChain(config, cfar_params, pd, device=local_rank)
    info = chain.info()
    P, C_, N, B, G = chain.P, chain.C, chain.N, chain.B, chain.G
    in_bytes, out_bytes = 8 * P * C_ * N, 8 * B * P * G
    alg_bytes = info["algorithmic_bytes_per_cpi"]
    assert alg_bytes == in_bytes + out_bytes
    l2_bytes = torch.cuda.get_device_properties(local_rank).L2_cache_size
    pool_n = args.pool or max(2, -(-2 * l2_bytes // in_bytes))             # inputs cycled: > 2x L2
    rdm_n = args.rdm_pool or max(2, -(-2 * l2_bytes // out_bytes))          # outputs cycled: > 2x L2
    slots = chain.stream_slots()
    cps = args.cpis_per_step or {"cfg1": 128, "cfg2": 64, "cfg3": 8, "native": 16}[name]
    cps = min(cps, slots)

    host_pool = make_pool(rsp, config, pd, pool_n, seed0=1000 * rank)
    pool = torch.from_numpy(host_pool).cuda()
    rdm_ring = torch.empty((rdm_n, B, G, P), dtype=torch.complex64, device="cuda")
    # everything timed runs on ONE explicit torch stream: the chain's kernels (rsp_set_stream), the
    # NCCL gather and the CUDA events that bracket the timed region
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    chain.set_stream(stream.cuda_stream)

    # detection ring as torch tensors (zero-copy) for the NCCL gather
    counts_ptr, recs_ptr = chain.stream_device_buffers()

    class _Raw:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    counts_t = torch.as_tensor(_Raw(counts_ptr, 4 * slots), device="cuda")[: 4 * cps]
    recs_t = torch.as_tensor(_Raw(recs_ptr, 40 * chain.max_detections * slots), device="cuda")
    gather_recs_cap = 512                                   # records gathered per CPI (lists are ~200 long)
    recs_view = recs_t.view(slots, chain.max_detections * 40)[:cps, : gather_recs_cap * 40]
    if world > 1:
        g_counts = torch.empty((world, 4 * cps), dtype=torch.uint8, device="cuda")
        g_recs = torch.empty((world, cps, gather_recs_cap * 40), dtype=torch.uint8, device="cuda")
        send_recs = torch.empty((cps, gather_recs_cap * 40), dtype=torch.uint8, device="cuda")

    def step():
        chain.stream_enqueue(pool.data_ptr(), pool_n, rdm_ring.data_ptr(), rdm_n, cps, 0)
        if world > 1:       # the path's one exchange step: detection lists -> every rank (rank 0 consumes)
            send_recs.copy_(recs_view)
            dist.all_gather_into_tensor(g_counts, counts_t)
            dist.all_gather_into_tensor(g_recs, send_recs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = chain.info()["launches_total"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        wall0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - wall0) * 1e3
    ms = ev0.elapsed_time(ev1)
    # the device-event time must agree with the host clock around the same (fully synchronised) region
    assert abs(wall_ms - ms) <= 0.1 * wall_ms + 1.0, f"event time {ms:.2f} ms vs wall {wall_ms:.2f} ms"
    launches = chain.info()["launches_total"] - launches0
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    total_cpis = world * args.steps * cps
    value = total_cpis / (ms_max / 1e3)

    # sanity: work was really done -- the last batch's slot 0 holds a plausible detection list
    d0 = chain.stream_fetch(0)
    assert 50 <= len(d0) <= chain.max_detections, f"implausible detection count {len(d0)}"

    # per-kernel device times: CUDA events around every launch on the launching stream, in a separate
    # pass (so the brackets do not perturb the headline number) that enqueues one CPI at a time, i.e. on a
    # single lane: each kernel runs alone, L2 warm, as in the ncu launch list under profiles/
    chain.set_profiling(True)
    for i in range(cps):
        chain.stream_enqueue(pool.data_ptr() + (i % pool_n) * in_bytes, 1, rdm_ring.data_ptr() + (i % rdm_n) * out_bytes, 1, 1, i)
    kt = chain.kernel_times()
    chain.set_profiling(False)
    per_cpi_ms = {k: v[0] / max(v[1], 1) for k, v in kt.items()}
    kern_sum = sum(per_cpi_ms.values())
    peak, peak_src = peaks()
    t_cpi_s = (ms / 1e3) / (args.steps * cps)                 # this rank's device time per CPI
    chain_achieved = alg_bytes / t_cpi_s / 1e9
    dominant = max(per_cpi_ms, key=per_cpi_ms.get) if per_cpi_ms else None
    # algorithmic bytes of each kernel's own launch: what it must read + write once
    beam_bytes, pc_bytes = 8 * P * B * N, 8 * P * B * G
    own_bytes = {"dbf": in_bytes + beam_bytes, "pc_fft": beam_bytes + pc_bytes, "pc_narrow": 0,
                 "mtd": pc_bytes + out_bytes + out_bytes // 2, "cfar_refine": 2 * (out_bytes // 2) * (B - 1) // B}
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(name, {}).get(dominant)
    except Exception:
        pass
    dom_ms = per_cpi_ms.get(dominant, 0.0)
    dom_achieved = own_bytes.get(dominant, 0) / (dom_ms / 1e3) / 1e9 if dom_ms else 0.0
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": dom_achieved, "peak": peak, "unit": "GB/s",
                "frac": dom_achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": own_bytes.get(dominant), "launch_ms": round(dom_ms, 5),
                "kernels_ms_per_cpi": {k: round(v, 5) for k, v in per_cpi_ms.items()},
                "kernels_share": {k: round(v / kern_sum, 4) for k, v in per_cpi_ms.items()} if kern_sum else {},
                "note": "dominant kernel class by device time, timed per CPI with CUDA events on its stream (pc_fft = both "
                        "launches of the mixed block plan); it works out of L2 and is bound by FP32 issue + the shared-memory "
                        "pipe (ncu: issue 52-57 %, L1 data pipe 34-59 %), see DESIGN.md sections 4-5; the chain-level figure is "
                        "chain_roofline"}
    chain_roofline = {"bound": "hbm", "achieved": chain_achieved, "peak": peak, "unit": "GB/s", "frac": chain_achieved / peak,
                      "algorithmic_bytes_per_cpi": alg_bytes, "kernels_per_cpi": info["kernels_per_cpi"],
                      "note": "8*P*N*C + 8*B*P*G bytes per CPI (SURVEY 8(d)) / device time per CPI of the timed region"}

    # end to end through the C ABI with HOST buffers: per CPI a 67 MB H2D copy of the pinned cube and a D2H
    # read of its sorted detection list, both inside the timed region.  rsp_submit_cpi pipelines the copy of
    # cube i+1 under the kernels of cube i (depth = lanes); rsp_stream_fetch collects in order.
    pinned = torch.from_numpy(host_pool[: min(pool_n, 4)]).pin_memory()
    cubes = [pinned[i].numpy() for i in range(len(pinned))]
    e2e_n = max(args.e2e_cpis, 2)
    depth = 3

    def e2e_pass(n):
        nd = 0
        for i in range(n + depth):
            if i < n:
                chain.submit_cpi(cubes[i % len(cubes)], i % slots)
            if i >= depth:
                nd += len(chain.stream_fetch((i - depth) % slots))
        return nd
    e2e_pass(4)
    barrier()
    t0 = time.perf_counter()
    nd = e2e_pass(e2e_n)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    t_e = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_n / float(t_e.item())
    d2h = 4 + 40 * min(512, chain.max_detections)      # the count and the first 512 records are prefetched to pinned memory

    # the reference's own signature: fun_process_single_frame(targets, ...) -> final_targets.  Only the
    # target list goes in and the clustered targets come out; S4 (echo synthesis + noise) runs on the GPU.
    chain.set_waveform(config, pd)
    v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
    tlist = [dict(Range=900.0, Velocity=0.15 * v_max, ElevationAngle=-5.0, SNR_dB=20.0),
             dict(Range=3000.0, Velocity=-0.10 * v_max, ElevationAngle=8.2, SNR_dB=10.0),
             dict(Range=8000.0, Velocity=0.05 * v_max, ElevationAngle=15.0, SNR_dB=10.0)]
    for i in range(2):
        chain.process_targets(tlist, cluster_params, 1.0, seed=i)
    barrier()
    n_t = max(args.e2e_cpis, 8)
    t0 = time.perf_counter()
    n_fin = 0
    for i in range(n_t):
        fin, _ = chain.process_targets(tlist, cluster_params, 1.0, seed=100 + i)
        n_fin += len(fin)
    torch.cuda.synchronize()
    e2e_targets_sync = world * n_t / (time.perf_counter() - t0)
    # the same call pipelined over the lanes (process_targets_batch = rsp_submit_targets / rsp_fetch_targets):
    # frame i+1 is synthesised while frame i runs; what the Monte-Carlo sweep and the tracker use
    n_tb = 8 * n_t
    chain.process_targets_batch([tlist] * 8, cluster_params, 1.0, list(range(8)))
    barrier()
    t0 = time.perf_counter()
    res = chain.process_targets_batch([tlist] * n_tb, cluster_params, 1.0, [1000 + i for i in range(n_tb)])
    torch.cuda.synchronize()
    e2e_targets = world * n_tb / (time.perf_counter() - t0)
    n_fin_b = sum(len(f) for f, _ in res)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_cpis or {"cfg1": 60, "cfg2": 32, "cfg3": 3, "native": 3}[name]      # ~10-20 s of CPU work
        with _all_threads(cores):
            v_all, dt_all, _ = cpu_reference_time(name, n_cpu, cores)
        v_one, dt_one, _ = cpu_reference_time(name, max(1, n_cpu // 4), 1)
        cpu_baseline = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n_cpu} CPIs of {name} ({dt_all:.1f} s), fp64 NumPy/SciPy oracle S5..S9 "
                                  f"(vectorised CFAR), scipy.fft workers={cores}; MATLAB/Octave unavailable offline",
                        "value_1_thread": v_one}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "wall_ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIG_DESC[name], "name": name, "C": C_, "B": B, "P": P, "N": N, "G": G,
                       "cpis_per_step": cps, "input_pool_cpis": pool_n, "rdm_pool": rdm_n,
                       "l2_policy": f"inputs ({pool_n * in_bytes / 1e6:.0f} MB) and RDM outputs "
                                    f"({rdm_n * out_bytes / 1e6:.0f} MB) cycled through pools larger than 2x L2 "
                                    f"({l2_bytes / 1e6:.0f} MB)",
                       "parallelism": f"cpi-sharded x{world}", "fft_len_medium": info["fft_len_medium"],
                       "fft_len_long": info["fft_len_long"], "blocks_long": info["blocks_long"],
                       "kernels_per_cpi": info["kernels_per_cpi"]},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": d2h,
                    "cpis": e2e_n, "note": "C ABI with host buffers: rsp_submit_cpi (pinned 67 MB cube H2D + chain, pipelined "
                                           "3 deep) and rsp_stream_fetch (sorted detection list D2H) per CPI; PCIe-bound"},
            "e2e_targets": {"value": e2e_targets, "unit": "frames/s", "frames": n_tb, "final_targets_per_frame": n_fin_b / n_tb,
                            "h2d_bytes_per_frame": 32 * len(tlist), "d2h_bytes_per_frame": d2h,
                            "one_frame_at_a_time": e2e_targets_sync, "one_at_a_time_targets_per_frame": n_fin / n_t,
                            "note": "the reference's own call signature fun_process_single_frame(targets, ...) -> "
                                    "final_targets: device echo synthesis + Philox noise (S4, fused into the DBF kernel on the pipelined path), S5..S9, host clustering; "
                                    "only target lists go in and detection lists come back.  value = frames pipelined "
                                    "over the lanes (rsp_submit_targets / rsp_fetch_targets); one_frame_at_a_time = "
                                    "synchronous rsp_process_targets"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "chain_roofline": chain_roofline,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    chain.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
