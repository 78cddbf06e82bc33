#!/usr/bin/env python
"""Host-buffer path (rsp_submit_cpi / rsp_stream_fetch) against the copy rate of the same pinned cubes in the same process.
    python tools/e2e_probe.py [--config cfg2] [--cpis 256]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rsp_b200 as rsp

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--cpis", type=int, default=256)
a = ap.parse_args()
config, cfar_params, cluster_params = rsp.named_config(a.config)
pd = rsp.build_precomputed_data(config)
chain = rsp.RadarChain(config, cfar_params, pd)
rng = np.random.default_rng(0)
host = (rng.standard_normal((4, chain.P, chain.C, chain.N, 2), dtype=np.float32) * 0.7071).view(np.complex64)[..., 0]
pinned = torch.from_numpy(host).pin_memory()
cubes = [pinned[i].numpy() for i in range(4)]
dev = torch.empty_like(pinned[0], device="cuda")
out = {"config": a.config, "lanes": chain.info()["lanes"], "cube_mb": round(pinned[0].numel() * 8 / 1e6, 1)}
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(64): dev.copy_(pinned[i % 4], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out["h2d_cubes_per_s"] = round(64 / dt, 1)
slots = chain.stream_slots()
def e2e_pass(n, depth):
    t_sub = t_fet = 0.0
    for i in range(n + depth):
        if i < n:
            t0 = time.perf_counter(); chain.submit_cpi(cubes[i % 4], i % slots); t_sub += time.perf_counter() - t0
        if i >= depth:
            t0 = time.perf_counter(); chain.stream_fetch((i - depth) % slots); t_fet += time.perf_counter() - t0
    return t_sub, t_fet
e2e_pass(12, 3)
for depth in (3, 6, 2, 3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ts, tf = e2e_pass(a.cpis, depth)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out[f"depth{depth}"] = {"cpis_per_s": round(a.cpis / dt, 1), "submit_us": round(ts / a.cpis * 1e6, 1), "fetch_wait_us": round(tf / a.cpis * 1e6, 1)}
print(json.dumps(out))
