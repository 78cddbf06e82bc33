#!/usr/bin/env python
"""Steady-state cost of each stage on the lanes: run the stream path with only some kernels enabled
(RSP_STAGES bit mask: 1 DBF, 2 PC, 4 MTD, 8 CFAR) and print us per CPI.   python tools/stage_probe.py [masks...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch, rsp_b200 as rsp
config, cfar_params, _ = rsp.named_config(os.environ.get("RSP_PROBE_CONFIG", "cfg2"))
chain = rsp.RadarChain(config, cfar_params, rsp.build_precomputed_data(config))
g = torch.Generator(device="cuda").manual_seed(0)
npool = int(os.environ.get("RSP_PROBE_POOL", "4"))
pool = torch.view_as_complex(torch.randn((npool, chain.P, chain.C, chain.N, 2), device="cuda", generator=g) * (0.5 ** 0.5)).contiguous()
rdm = torch.empty((12, chain.B, chain.G, chain.P), dtype=torch.complex64, device="cuda")
chain.set_stream(torch.cuda.current_stream().cuda_stream)
n = int(os.environ.get("RSP_PROBE_CPIS", "64"))
def run(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        chain.stream_enqueue(pool.data_ptr(), npool, rdm.data_ptr(), 12, n, 0)
    chain.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / (reps * n) * 1e6
run(2)
print("%%.2f" %% min(run(6), run(6)))
''' % ROOT
masks = [int(a) for a in sys.argv[1:]] or [15, 1, 2, 4, 8, 3, 12, 6, 7, 14]
names = {1: "dbf", 2: "pc", 4: "mtd", 8: "cfar"}
PROBES_LIB = os.path.join(ROOT, "radar-signal-simulation-and-target-detection_b200", "lib", "variants", "librsp_probes.so")
if "RSP_LIBRARY" not in os.environ:
    if not os.path.exists(PROBES_LIB):
        sys.exit("build the probes variant first: python radar-signal-simulation-and-target-detection_b200/build.py --variant probes -DRSP_PROBES")
    os.environ["RSP_LIBRARY"] = PROBES_LIB
for m in masks:
    env = dict(os.environ, RSP_STAGES=str(m))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    label = "+".join(v for k, v in names.items() if m & k)
    print(f"{label:20s} lanes={env.get('RSP_LANES', '3'):2s} {r.stdout.strip() or r.stderr[-300:]} us/CPI", flush=True)
