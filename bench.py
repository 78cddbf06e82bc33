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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                try:
                    pw.append(float(r[2]))
                except Exception:
                    pass
                for nm, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w": float(np.median(pw)) if pw else None}


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


def probe_reference_runtime():
    """BASELINE.md 5.3 / SURVEY 8(d): is there a MATLAB or an Octave on this box that could run the unmodified
    fun_process_single_frame.m (tools/ref_golden.m drives it)?"""
    import shutil
    found = {k: shutil.which(k) for k in ("octave", "octave-cli", "matlab")}
    return {k: v for k, v in found.items() if v} or None


def run_reference_runtime(exe: str, name: str, n_frames: int):
    """Time the reference's own fun_process_single_frame.m through tools/ref_golden.m; None when it cannot run here."""
    ref_dir = os.environ.get("RSP_REFERENCE_DIR", "/root/reference/Simulation")
    if not os.path.isdir(ref_dir):
        return None
    script = f"addpath('{os.path.join(ROOT, 'tools', 'ref_golden')}'); ref_golden('{ref_dir}', '', '{name}', {n_frames});"
    cmd = [exe, "--no-gui", "--eval", script] if "octave" in os.path.basename(exe) else [exe, "-batch", script]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900).stdout
        for line in out.splitlines():
            if line.startswith("REF_FRAMES_PER_SEC"):
                return float(line.split()[1])
    except Exception:
        pass
    return None


def t3_targets(config):
    v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
    return [dict(Range=900.0, Velocity=0.15 * v_max, ElevationAngle=-5.0, SNR_dB=20.0),
            dict(Range=3000.0, Velocity=-0.10 * v_max, ElevationAngle=8.2, SNR_dB=10.0),
            dict(Range=8000.0, Velocity=0.05 * v_max, ElevationAngle=15.0, SNR_dB=10.0)]


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores.  The reference is MATLAB;
    when neither MATLAB nor Octave is on the box (probed, reported) this is the fp64 NumPy/SciPy port (kind "port")."""
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
        # the reference's own call signature, the twin of the product's e2e_targets: targets -> S4 (echo synthesis + noise)
        # -> S5..S9 -> S10/S11 clustering, one frame per call
        tg = o.targets_t3(cfg, pre)
        n_fr = max(3, min(args.steps, 8))
        o.fun_process_single_frame(tg, cfg, pre, seed=0, workers=cores)          # warm-up
        t1 = time.perf_counter()
        n_fin = 0
        for i in range(n_fr):
            n_fin += len(o.fun_process_single_frame(tg, cfg, pre, seed=1 + i, workers=cores).final_targets)
        dt_fr = time.perf_counter() - t1
    value = args.steps * per_step / dt
    runtimes = probe_reference_runtime()
    ref_fps = None
    if runtimes:
        ref_fps = run_reference_runtime(next(iter(runtimes.values())), name, 3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": CONFIG_DESC[name], "name": name, "cpis_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} CPIs of {name}, fp64 NumPy/SciPy oracle S5..S9, scipy.fft workers={cores}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "e2e_targets": {"value": n_fr / dt_fr, "unit": "frames/s", "frames": n_fr, "final_targets_per_frame": n_fin / n_fr,
                        "note": "oracle fun_process_single_frame(targets, ...): S4 echo synthesis + NumPy noise, S5..S9, S10/S11 "
                                "clustering, one frame per call on all host cores -- the CPU twin of the product's e2e_targets"},
        "gpu_launches": 0,
        "reference_runtime": {"found": runtimes, "frames_per_sec": ref_fps,
                              "note": "command -v octave octave-cli matlab on this box; when one exists tools/ref_golden.m runs the "
                                      "unmodified fun_process_single_frame.m and its rate is reported here (kind reference-octave)"},
        "note": "reference is MATLAB; " + ("a MATLAB/Octave runtime was found, see reference_runtime" if runtimes else
                                           "no MATLAB/Octave on this box (probed)") + "; the timed arm is the NumPy/SciPy port of "
                "fun_process_single_frame.m S5..S9",
    }
    print(json.dumps(line), flush=True)


# warp instructions per CPI and DRAM bytes of every kernel, read from one `ncu --set full` capture of this build
# (profiles/traffic.json, written by tools/ncu_summary.py); None when the file has no entry for the config
def load_profile_numbers(name):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(name)
    except Exception:
        return None


def alg_flops(P, N, C, B, G, nfft_pts):
    """8PNCB (DBF) + 2 * 5 L log2 L per transformed block point (PC, forward + inverse) + 5 P log2 P G B (MTD) +
    30 P G (B - 1) (CFAR)."""
    import math
    return 8.0 * P * N * C * B + nfft_pts + 5.0 * P * math.log2(P) * G * B + 30.0 * P * G * (B - 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=list(CONFIG_DESC))
    ap.add_argument("--cpis-per-step", type=int, default=0, help="CPIs per step (a step is `batches` stream batches)")
    ap.add_argument("--cpis-per-batch", type=int, default=0, help="CPIs per rsp_stream_enqueue batch (at most half the ring slots)")
    ap.add_argument("--pool", type=int, default=0, help="distinct input CPIs resident in HBM")
    ap.add_argument("--rdm-pool", type=int, default=0, help="distinct RDM output buffers")
    ap.add_argument("--e2e-cpis", type=int, default=256)
    ap.add_argument("--cpu-cpis", type=int, default=0, help="CPIs in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--stream-cpis", type=int, default=1024, help="CPIs of the config-3 stream sub-record (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the e2e / frames / config-3 / config-4 sub-records (A/B runs)")
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
    from rsp_b200 import stream as rstream

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peak, peak_src = peaks()
    l2_bytes = torch.cuda.get_device_properties(local_rank).L2_cache_size
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count

    class _Raw:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stream_bench(name, steps, warmup, cps, batches, pool_arg=0, rdm_arg=0, clocks=False, total_cpis=None):
        """The device-resident CPI stream of one configuration on this rank: `steps` steps of `batches` batches of `cps` CPIs
        (or exactly total_cpis CPIs), each batch followed by the asynchronous gather of its detection lists.  Returns a dict."""
        config, cfar_params, cluster_params = rsp.named_config(name)
        pd = rsp.build_precomputed_data(config)
        chain = rsp.RadarChain(config, cfar_params, pd, device=local_rank)
        info = chain.info()
        P, C_, N, B, G = chain.P, chain.C, chain.N, chain.B, chain.G
        in_bytes, out_bytes = 8 * P * C_ * N, 8 * B * P * G
        alg_bytes = info["algorithmic_bytes_per_cpi"]
        assert alg_bytes == in_bytes + out_bytes
        lanes = info["lanes"]
        pool_n = pool_arg or max(2, -(-2 * l2_bytes // in_bytes))              # inputs cycled: > 2x L2
        rdm_n = rdm_arg or max(2, -(-2 * l2_bytes // out_bytes))               # outputs cycled: > 2x L2
        rdm_n = -(-rdm_n // lanes) * lanes                                       # a multiple of the lanes (rsp.h)
        slots = chain.stream_slots()
        cps = min(cps, slots // 2)                                               # two slot ranges alternate under the gather
        host_pool = make_pool(rsp, config, pd, pool_n, seed0=1000 * rank)
        pool = torch.from_numpy(host_pool).cuda()
        rdm_ring = torch.empty((rdm_n, B, G, P), dtype=torch.complex64, device="cuda")
        # everything timed runs on ONE explicit torch stream: the chain's kernels (rsp_set_stream) and the CUDA events
        # that bracket the timed region; the gather rides on a side stream behind an event (AsyncDetectionGather)
        st = torch.cuda.Stream()
        torch.cuda.set_stream(st)
        chain.set_stream(st.cuda_stream)
        counts_ptr, recs_ptr = chain.stream_device_buffers()
        counts_t = torch.as_tensor(_Raw(counts_ptr, 4 * slots), device="cuda")
        recs_t = torch.as_tensor(_Raw(recs_ptr, 40 * chain.max_detections * slots), device="cuda").view(slots, chain.max_detections * 40)
        gather_cap = min(1024, chain.max_detections)
        gat = rstream.AsyncDetectionGather(counts_t, recs_t, cps, gather_cap)
        state = {"k": 0, "pending": [None, None]}

        def batch():
            k = state["k"] & 1
            state["k"] += 1
            if state["pending"][k] is not None:
                gat.fence(state["pending"][k], st)           # the slots of two batches ago have been gathered
            chain.stream_enqueue(pool.data_ptr(), pool_n, rdm_ring.data_ptr(), rdm_n, cps, k * cps)
            state["pending"][k] = gat.launch(k * cps, st)    # the path's one exchange step, overlapped with the next batch
            return state["pending"][k]

        n_batches = steps * batches if total_cpis is None else -(-total_cpis // cps)
        for _ in range(warmup * (batches if total_cpis is None else 1)):
            batch()
        barrier()
        launches0 = chain.info()["launches_total"]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank) if clocks else None
        if sampler:
            sampler.__enter__()
        barrier()
        wall0 = time.perf_counter()
        ev0.record(st)
        last = None
        for _ in range(n_batches):
            last = batch()
        if gat.side is not None:
            st.wait_event(gat.done[last])                      # the timed region ends when the last gather has landed
        ev1.record(st)
        barrier()
        wall_ms = (time.perf_counter() - wall0) * 1e3
        if sampler:
            sampler.__exit__(None, None, None)
        ms = ev0.elapsed_time(ev1)
        assert abs(wall_ms - ms) <= 0.1 * wall_ms + 1.0, f"event time {ms:.2f} ms vs wall {wall_ms:.2f} ms"
        launches = chain.info()["launches_total"] - launches0
        t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms_max = float(t_ms.item())
        # rank 0 consumes the last batch: every rank's lists, overflow is an error (stream.AsyncDetectionGather.wait)
        lists = gat.wait(last)
        n0 = len(lists[0][0])
        assert 50 <= n0 <= gather_cap, f"implausible detection count {n0}"
        res = dict(chain=chain, config=config, cfar_params=cfar_params, cluster_params=cluster_params, pd=pd, info=info,
                   ms=ms, ms_max=ms_max, wall_ms=wall_ms, n_batches=n_batches, cps=cps, launches=launches,
                   pool=pool, pool_n=pool_n, rdm_ring=rdm_ring, rdm_n=rdm_n, in_bytes=in_bytes, out_bytes=out_bytes,
                   alg_bytes=alg_bytes, host_pool=host_pool, slots=slots, stream=st, gather_cap=gather_cap,
                   clocks=sampler.summary() if sampler else None, n_ranks_gathered=len(lists))
        return res

    name = args.config
    cps_default = {"cfg1": 64, "cfg2": 64, "cfg3": 8, "native": 16}[name]
    per_step_default = {"cfg1": 2560, "cfg2": 1024, "cfg3": 96, "native": 96}[name]     # >= 1 s of device time in 20 steps
    per_step = args.cpis_per_step or per_step_default
    cps = min(args.cpis_per_batch or cps_default, per_step)
    batches = max(1, per_step // cps)
    r = stream_bench(name, args.steps, args.warmup, cps, batches, args.pool, args.rdm_pool, clocks=True)
    chain, config, cfar_params, cluster_params, pd, info = r["chain"], r["config"], r["cfar_params"], r["cluster_params"], r["pd"], r["info"]
    P, C_, N, B, G = chain.P, chain.C, chain.N, chain.B, chain.G
    cps, in_bytes, out_bytes, alg_bytes = r["cps"], r["in_bytes"], r["out_bytes"], r["alg_bytes"]
    pool, pool_n, rdm_ring, rdm_n, slots, stream = r["pool"], r["pool_n"], r["rdm_ring"], r["rdm_n"], r["slots"], r["stream"]
    cpis_per_step = cps * batches
    total_cpis = world * r["n_batches"] * cps
    value = total_cpis / (r["ms_max"] / 1e3)
    t_cpi_s = (r["ms"] / 1e3) / (r["n_batches"] * cps)                 # this rank's device time per CPI

    # per-kernel device times: CUDA events around every launch on the launching stream, in a separate pass (so the
    # brackets do not perturb the headline number) that enqueues one CPI at a time, i.e. on a single lane: each kernel
    # runs alone, L2 warm, as in the ncu launch list under profiles/
    chain.set_profiling(True)
    for i in range(cps):
        chain.stream_enqueue(pool.data_ptr() + (i % pool_n) * in_bytes, 1, rdm_ring.data_ptr() + (i % rdm_n) * out_bytes, 1, 1, i)
    kt = chain.kernel_times()
    chain.set_profiling(False)
    per_cpi_ms = {k: v[0] / max(v[1], 1) * (v[1] / cps if k not in ("refine",) else 1.0) for k, v in kt.items()}
    kern_sum = sum(per_cpi_ms.values())
    prof = load_profile_numbers(name) or {}
    beam_bytes, pc_bytes = 8 * P * B * N, 8 * P * B * G
    # The roofline object is the DBF, the kernel that streams the input cube out of HBM: its own unavoidable bytes are the
    # raw cube in and the beam cube out.  The launch time is the event-bracketed one-at-a-time figure of this run.
    dbf_ms = per_cpi_ms.get("dbf", per_cpi_ms.get("dbf_pc", 0.0))
    dbf_bytes = in_bytes + beam_bytes
    dbf_achieved = dbf_bytes / (dbf_ms / 1e3) / 1e9 if dbf_ms else 0.0
    roofline = {"bound": "hbm", "kernel": "dbf (dbf_tc_kernel: tcgen05 + TMA)", "achieved": dbf_achieved, "peak": peak, "unit": "GB/s",
                "frac": dbf_achieved / peak, "traffic": (prof.get("dbf") or {}).get("dram_bytes") if isinstance(prof.get("dbf"), dict) else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dbf_bytes, "launch_ms": round(dbf_ms, 5),
                "kernels_ms_per_cpi": {k: round(v, 5) for k, v in per_cpi_ms.items()},
                "kernels_share": {k: round(v / kern_sum, 4) for k, v in per_cpi_ms.items()} if kern_sum else {},
                "kernel_bounds": {"dbf": "hbm", "pc_fft": "issue (fp32 + shared-memory pipe, works out of L2)", "mtd": "issue / latency",
                                  "cfar": "issue / L1", "refine": "latency"},
                "note": "the kernel that streams the input: raw cube read once + beam cube written once per launch; launch_ms is "
                        "the one-at-a-time event-bracketed time of this run (steady state inside the stream is shorter, see "
                        "profiles/); the chain-level figure, the one BASELINE.json's metric asks for, is chain_roofline"}
    chain_achieved = alg_bytes / t_cpi_s / 1e9
    chain_roofline = {"bound": "hbm", "achieved": chain_achieved, "peak": peak, "unit": "GB/s", "frac": chain_achieved / peak,
                      "algorithmic_bytes_per_cpi": alg_bytes, "kernels_per_cpi": info["kernels_per_cpi"],
                      "note": "8*P*N*C + 8*B*P*G bytes per CPI (SURVEY 8(d)) / device time per CPI of the timed region"}
    # the other two ceilings SURVEY 8(d) asks for: warp-instruction issue and fp32 arithmetic
    sm_hz = (r["clocks"] or {}).get("sm_mhz") or 1965.0
    issue_peak = n_sm * 4 * sm_hz * 1e6                                          # warp instructions per second
    winstr = prof.get("warp_instructions_per_cpi")
    roofline_issue = {"bound": "issue", "warp_instructions_per_cpi": winstr, "peak_warp_instr_per_s": issue_peak,
                      "frac": (winstr / t_cpi_s / issue_peak) if winstr else None,
                      "note": "executed warp instructions per CPI (ncu smsp__inst_executed.sum over the chain's kernels, "
                              "profiles/traffic.json) / device time per CPI / (SMs x 4 schedulers x SM clock)"}
    nfft_pts = 2 * 5.0 * P * B * sum(L * np.log2(L) * nb for L, nb in ((info["fft_len_medium"], info["blocks_medium"]),)) if info["fft_len_medium"] else 0.0
    long_pts = prof.get("long_block_points") or (info["fft_len_long"] * info["blocks_long"])
    nfft_pts += 2 * 5.0 * P * B * long_pts * np.log2(max(info["fft_len_long"], 2))
    flops = alg_flops(P, N, C_, B, G, nfft_pts)
    fp32_peak = n_sm * 128 * 2 * sm_hz * 1e6
    roofline_fp32 = {"bound": "fp32", "algorithmic_gflop_per_cpi": flops / 1e9, "achieved_tflops": flops / t_cpi_s / 1e12,
                     "peak_tflops": fp32_peak / 1e12, "frac": flops / t_cpi_s / fp32_peak,
                     "note": "SURVEY 8(d) flop count with this build's block plan / device time per CPI / (SMs x 128 lanes x 2 x SM clock)"}

    e2e = e2e_targets_rec = config4 = stream_cfg3 = None
    d2h = 4 + 40 * min(2048, chain.max_detections)     # the count and the first 2048 records are prefetched to pinned memory
    if not args.no_extras:
        # end to end through the C ABI with HOST buffers: per CPI a 67 MB H2D copy of the pinned cube and a D2H read of its
        # sorted detection list, both inside the timed region.  rsp_submit_cpi pipelines the copy of cube i+1 under the
        # kernels of cube i (depth = lanes); rsp_stream_fetch collects in order.
        host_pool = r["host_pool"]
        pinned = torch.from_numpy(host_pool[: min(pool_n, 4)]).pin_memory()
        cubes = [pinned[i].numpy() for i in range(len(pinned))]
        e2e_n = max(args.e2e_cpis, 2)
        depth = 3

        host_us = {"submit": 0.0, "fetch_wait": 0.0}

        def e2e_pass(n):
            nd = 0
            for i in range(n + depth):
                if i < n:
                    t1 = time.perf_counter()
                    chain.submit_cpi(cubes[i % len(cubes)], i % slots)
                    host_us["submit"] += time.perf_counter() - t1
                if i >= depth:
                    t1 = time.perf_counter()
                    nd += len(chain.stream_fetch((i - depth) % slots))
                    host_us["fetch_wait"] += time.perf_counter() - t1
            return nd
        e2e_pass(4 * info["lanes"])                       # every lane has its staging cube and its RDM before the clock starts
        # the link's own rate in this process, same pinned cubes, one cudaMemcpyAsync per cube: the ceiling of this path
        dev_cube = torch.empty(pinned[0].shape, dtype=pinned.dtype, device="cuda")
        for i in range(8):                        # warm-up: first copies into a fresh allocation run below the link rate
            dev_cube.copy_(pinned[i % len(cubes)], non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(96):
            dev_cube.copy_(pinned[i % len(cubes)], non_blocking=True)
        torch.cuda.synchronize()
        h2d_cubes_per_s = 96 / (time.perf_counter() - t0)
        del dev_cube
        barrier()
        host_us["submit"] = host_us["fetch_wait"] = 0.0
        t0 = time.perf_counter()
        e2e_pass(e2e_n)
        torch.cuda.synchronize()
        e2e_dt = time.perf_counter() - t0
        t_e = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * e2e_n / float(t_e.item()), "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": d2h,
               "cpis": e2e_n, "h2d_copy_ceiling_cpis_per_sec": world * h2d_cubes_per_s,
               "host_submit_us_per_cpi": host_us["submit"] / e2e_n * 1e6, "host_fetch_wait_us_per_cpi": host_us["fetch_wait"] / e2e_n * 1e6,
               "note": "C ABI with host buffers: rsp_submit_cpi (pinned cube H2D + chain, pipelined 3 deep) and "
                                      "rsp_stream_fetch (sorted detection list D2H) per CPI; PCIe-bound"}

        # the reference's own signature: fun_process_single_frame(targets, ...) -> final_targets.  Only the target list goes
        # in and the clustered targets come out; S4 (echo synthesis + noise) runs on the GPU.  Own context: a context with a
        # waveform takes the mma.sync DBF on every path (bit-identical frames whichever path they take).
        fchain = rsp.RadarChain(config, cfar_params, pd, device=local_rank)
        fchain.set_waveform(config, pd)
        tlist = t3_targets(config)
        for i in range(2):
            fchain.process_targets(tlist, cluster_params, 1.0, seed=i)
        barrier()
        n_t = 32
        t0 = time.perf_counter()
        n_fin = 0
        for i in range(n_t):
            fin, _ = fchain.process_targets(tlist, cluster_params, 1.0, seed=100 + i)
            n_fin += len(fin)
        torch.cuda.synchronize()
        e2e_targets_sync = world * n_t / (time.perf_counter() - t0)
        n_tb = 4096 if name in ("cfg1", "cfg2") else 64
        fchain.process_targets_batch([tlist] * 8, cluster_params, 1.0, list(range(8)))
        barrier()
        t0 = time.perf_counter()
        res = fchain.process_targets_batch([tlist] * n_tb, cluster_params, 1.0, [1000 + i for i in range(n_tb)], return_detections=False)
        torch.cuda.synchronize()
        dt_b = time.perf_counter() - t0
        n_fin_b = sum(len(f) for f, _ in res)
        t0 = time.perf_counter()
        fchain.process_targets_batch([tlist] * n_tb, cluster_params, 1.0, [1000 + i for i in range(n_tb)])
        torch.cuda.synchronize()
        dt_b_dets = time.perf_counter() - t0
        e2e_targets_rec = {"value": world * n_tb / dt_b, "unit": "frames/s", "frames": n_tb, "final_targets_per_frame": n_fin_b / n_tb,
                           "h2d_bytes_per_frame": 32 * len(tlist), "d2h_bytes_per_frame": d2h,
                           "with_detection_lists": world * n_tb / dt_b_dets,
                           "one_frame_at_a_time": e2e_targets_sync, "one_at_a_time_targets_per_frame": n_fin / n_t,
                           "note": "the reference's own call signature fun_process_single_frame(targets, ...) -> final_targets: "
                                   "device echo synthesis + Philox noise (S4, fused into the DBF kernel on the pipelined path), "
                                   "S5..S9, host clustering; only target lists go in and final targets come out, as from the reference function "
                                   "(with_detection_lists: the same call also handing every frame's sorted detection list to Python).  value = frames "
                                   "pipelined over the lanes by one native call per 256 frames (rsp_process_frames: submit / fetch + worker threads "
                                   "for sorting and clustering); one_frame_at_a_time = "
                                   "synchronous rsp_process_targets.  CPU twin: --impl reference, e2e_targets"}
        # BASELINE config 4: 64 targets per frame synthesised inside the DBF (R, V, El, SNR uniform, default_rng(1))
        rng = np.random.default_rng(1)
        dR = float(pd.deltaR) if hasattr(pd, "deltaR") else float(pd["deltaR"])
        v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
        vb = (P / 2 - 16) / P * v_max
        t64 = [dict(Range=float(rng.uniform(700 * dR, (G - 16) * dR)), Velocity=float(rng.uniform(-vb, vb)),
                    ElevationAngle=float(rng.uniform(-15.0, 60.0)), SNR_dB=float(rng.uniform(-10.0, 20.0))) for _ in range(64)]
        big = rsp.RadarChain(config, cfar_params, pd, device=local_rank, max_detections=32768)
        big.set_waveform(config, pd)
        n4 = 1024 if name in ("cfg1", "cfg2") else 16
        big.process_targets_batch([t64] * 6, cluster_params, 1.0, list(range(6)))
        barrier()
        t0 = time.perf_counter()
        res4t = big.process_targets_batch([t64] * n4, cluster_params, 1.0, [50 + i for i in range(n4)], return_detections=False)
        torch.cuda.synchronize()
        dt4 = time.perf_counter() - t0
        t0 = time.perf_counter()
        res4 = big.process_targets_batch([t64] * n4, cluster_params, 1.0, [50 + i for i in range(n4)])
        torch.cuda.synchronize()
        dt4_dets = time.perf_counter() - t0
        assert all(np.array_equal(a[0], b[0]) for a, b in zip(res4t, res4))
        config4 = {"value": world * n4 / dt4, "unit": "frames/s", "frames": n4, "targets_per_frame": 64,
                   "final_targets_per_frame": sum(len(f) for f, _ in res4) / n4, "detections_per_frame": sum(len(d) for _, d in res4) / n4,
                   "with_detection_lists": world * n4 / dt4_dets,
                   "note": "BASELINE configs[3]: 64-target echo synthesis (Philox noise) fused into the DBF ahead of the chain, "
                           "pipelined over the lanes (rsp_process_frames, host sorting + clustering on worker threads); the raw cube is never "
                           "written; value = final targets returned as by the reference function, with_detection_lists = the sorted detection "
                           "lists handed to Python as well"}
        big.close()
        fchain.close()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_cpis or {"cfg1": 60, "cfg2": 32, "cfg3": 3, "native": 3}[name]      # ~10-20 s of CPU work
        with _all_threads(cores):
            v_all, dt_all, _ = cpu_reference_time(name, n_cpu, cores)
        v_one, dt_one, _ = cpu_reference_time(name, max(1, n_cpu // 4), 1)
        runtimes = probe_reference_runtime()
        cpu_baseline = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{n_cpu} CPIs of {name} ({dt_all:.1f} s), fp64 NumPy/SciPy oracle S5..S9 "
                                  f"(vectorised CFAR), scipy.fft workers={cores}",
                        "value_1_thread": v_one, "matlab_or_octave_on_this_box": runtimes}

    main_clocks = r["clocks"]
    main_launches = int(r["launches"])
    main_ms_max, main_wall = r["ms_max"], r["wall_ms"]
    chain.close()
    del pool, rdm_ring, r
    torch.cuda.empty_cache()

    # BASELINE configs[2]: the 1024-CPI stream of 32 ch x 16 beams x 128 pulses x 16384 bins, CPI-sharded over the ranks
    # (stream.shard_range: contiguous blocks), measured at every N so that the driver sees it at 1/2/4/8 GPUs
    if args.stream_cpis > 0 and not args.no_extras and name != "cfg3":
        lo, hi = rstream.shard_range(args.stream_cpis, rank, world)
        r3 = stream_bench("cfg3", 0, 1, 8, 1, 2, 0, clocks=False, total_cpis=hi - lo)
        c3 = r3["chain"]
        done = r3["n_batches"] * r3["cps"]
        agg = world * done / (r3["ms_max"] / 1e3) if world > 1 else done / (r3["ms_max"] / 1e3)
        stream_cfg3 = {"cpis": args.stream_cpis, "cpis_this_rank": done, "cpis_per_sec": agg, "per_gpu": agg / world,
                       "ms_total": r3["ms_max"], "chain_frac": r3["alg_bytes"] / ((r3["ms"] / 1e3) / done) / 1e9 / peak,
                       "input_pool_cpis": r3["pool_n"], "kernels_per_cpi": r3["info"]["kernels_per_cpi"],
                       "note": "BASELINE configs[2]: CPIs sharded in contiguous blocks over the ranks (stream.shard_range), pool of 2 "
                               "distinct cubes per GPU (1.07 GB > 2x L2), detection lists gathered per batch of 8 CPIs on a side stream"}
        c3.close()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_ms_max / args.steps, "wall_ms_per_step": main_wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": CONFIG_DESC[name], "name": name, "C": C_, "B": B, "P": P, "N": N, "G": G,
                       "cpis_per_step": cpis_per_step, "cpis_per_batch": cps, "batches_per_step": batches,
                       "input_pool_cpis": pool_n, "rdm_pool": rdm_n,
                       "l2_policy": f"inputs ({pool_n * in_bytes / 1e6:.0f} MB) and RDM outputs "
                                    f"({rdm_n * out_bytes / 1e6:.0f} MB) cycled through pools larger than 2x L2 "
                                    f"({l2_bytes / 1e6:.0f} MB)",
                       "parallelism": f"cpi-sharded x{world}", "fft_len_medium": info["fft_len_medium"],
                       "fft_len_long": info["fft_len_long"], "blocks_long": info["blocks_long"],
                       "kernels_per_cpi": info["kernels_per_cpi"],
                       "exchange": "one packed all_gather per batch (count header + records) on a side stream, overlapped with the next batch"},
            "clocks": main_clocks,
            "e2e": e2e,
            "e2e_targets": e2e_targets_rec,
            "config4": config4,
            "stream_cfg3": stream_cfg3,
            "gpu_launches": main_launches,
            "roofline": roofline,
            "chain_roofline": chain_roofline,
            "roofline_issue": roofline_issue,
            "roofline_fp32": roofline_fp32,
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
