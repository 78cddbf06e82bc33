#!/usr/bin/env python
"""Monte-Carlo SNR sweep (BASELINE config 5) on 1..8 GPUs:
    python tools/mc_sweep.py --trials 100 [--config native]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mc_sweep.py --trials 480
Prints one JSON line (rank 0)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import rsp_b200 as rsp

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="native")
ap.add_argument("--trials", type=int, default=100)
ap.add_argument("--snr", default="-10:2:30")
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lo, st, hi = (float(x) for x in a.snr.split(":"))
snr = np.arange(lo, hi + 1e-9, st)
config, cfar_params, cluster_params = rsp.named_config(a.config)
pd = rsp.build_precomputed_data(config)
t0 = time.perf_counter()
chain = rsp.RadarChain(config, cfar_params, pd, device=local, monopulse_complex=True)      # mc:246-262 uses the complex ratio
chain.set_waveform(config, pd)
rsp.snr_vs_angle_error(config, cfar_params, cluster_params, pd, snr[:1], 8, device=local, chain=chain)   # warm-up (lane buffers)
torch.cuda.synchronize()
setup = time.perf_counter() - t0
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
res = rsp.snr_vs_angle_error(config, cfar_params, cluster_params, pd, snr, a.trials, device=local, rank=rank, world=world, chain=chain)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"config": a.config, "n_gpus": world, "trials_per_point": a.trials, "seconds": dt, "setup_seconds": setup,
                      "frames_per_sec": len(snr) * a.trials / dt, "snr_db": res["snr_db"],
                      "angle_error_std": [None if np.isnan(x) else round(float(x), 5) for x in res["angle_error_std"]],
                      "detection_probability": [round(float(x), 4) for x in res["detection_probability"]],
                      "theoretical_error_std": [round(float(x), 5) for x in res["theoretical_error_std"]]}))
if world > 1:
    dist.destroy_process_group()
