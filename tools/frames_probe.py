#!/usr/bin/env python
"""Where does the time of the targets -> final_targets path go?  (synthesis kernel, chain, host fetch + clustering)
    python tools/frames_probe.py [--config cfg2] [--frames 192]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rsp_b200 as rsp

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--frames", type=int, default=192)
ap.add_argument("--targets", type=int, default=0, help="K random targets per frame (SURVEY 8(d) config 4) instead of T3")
a = ap.parse_args()
config, cfar_params, cluster_params = rsp.named_config(a.config)
pd = rsp.build_precomputed_data(config)
chain = rsp.RadarChain(config, cfar_params, pd, max_detections=32768 if a.targets else 8192)
chain.set_waveform(config, pd)
v_max = config.Sig_Config.wavelength / (2 * config.Sig_Config.prt)
tl = [dict(Range=900.0, Velocity=0.15 * v_max, ElevationAngle=-5.0, SNR_dB=20.0),
      dict(Range=3000.0, Velocity=-0.10 * v_max, ElevationAngle=8.2, SNR_dB=10.0),
      dict(Range=8000.0, Velocity=0.05 * v_max, ElevationAngle=15.0, SNR_dB=10.0)]
if a.targets:
    import numpy as np
    rng = np.random.default_rng(1)
    dR = float(pd.deltaR)
    vb = (chain.P / 2 - 16) / chain.P * v_max
    tl = [dict(Range=float(rng.uniform(700 * dR, (chain.G - 16) * dR)), Velocity=float(rng.uniform(-vb, vb)),
               ElevationAngle=float(rng.uniform(-15.0, 60.0)), SNR_dB=float(rng.uniform(-10.0, 20.0))) for _ in range(a.targets)]
out = {"config": a.config, "lanes": chain.info()["lanes"], "targets_per_frame": len(tl)}
for _ in range(3):
    chain.process_targets(tl, cluster_params, 1.0, 1)
chain.set_profiling(True)
chain.kernel_times()
for i in range(8):
    chain.process_targets(tl, cluster_params, 1.0, i)
out["kernel_ms_per_frame_sync"] = {k: round(v[0] / 8, 5) for k, v in chain.kernel_times().items()}
chain.set_profiling(False)
n = a.frames
for depth in (1, 2, 3, 6, 12):
    chain.process_targets_batch([tl] * 8, cluster_params, 1.0, list(range(8)), depth=depth, native=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    chain.process_targets_batch([tl] * n, cluster_params, 1.0, list(range(n)), depth=depth, native=False)
    torch.cuda.synchronize()
    out[f"frames_per_s_depth{depth}"] = round(n / (time.perf_counter() - t0), 1)
# rsp_process_frames: submission, fetch and a pool of sorting + clustering threads inside librsp
for threads in (1, 2, 4, 8):
    for ret in (True, False):
        chain.process_targets_batch([tl] * 16, cluster_params, 1.0, list(range(16)), host_threads=threads, return_detections=ret)
        t0 = time.perf_counter()
        chain.process_targets_batch([tl] * (4 * n), cluster_params, 1.0, list(range(4 * n)), host_threads=threads, return_detections=ret)
        out[f"native_frames_per_s_threads{threads}" + ("" if ret else "_targets_only")] = round(4 * n / (time.perf_counter() - t0), 1)
# host side only: submit everything, wait, then time the fetch + sort + cluster of finished slots
slots = min(chain.stream_slots(), 96)
for i in range(slots):
    chain.submit_targets(tl, i, 1.0, i)
chain.synchronize()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(slots):
    chain.fetch_targets(i, cluster_params)
out["host_fetch_cluster_us_per_frame"] = round((time.perf_counter() - t0) / slots * 1e6, 1)
t0 = time.perf_counter()
for i in range(slots):
    chain.submit_targets(tl, i, 1.0, i)
out["host_submit_us_per_frame"] = round((time.perf_counter() - t0) / slots * 1e6, 1)
chain.synchronize()
torch.cuda.synchronize()
for i in range(slots):
    chain.fetch_targets(i, cluster_params)          # every submit is paired with a fetch (rsp.h)
# device only: submit all, one sync
t0 = time.perf_counter()
for i in range(slots):
    chain.submit_targets(tl, i, 1.0, i)
chain.synchronize()
torch.cuda.synchronize()
out["device_only_frames_per_s"] = round(slots / (time.perf_counter() - t0), 1)
for i in range(slots):
    chain.fetch_targets(i, cluster_params)
print(json.dumps(out))
