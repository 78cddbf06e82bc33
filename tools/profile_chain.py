#!/usr/bin/env python
"""Minimal driver for ncu: run n CPIs of one configuration through the device-resident stream path.

    python tools/profile_chain.py [--config cfg2] [--cpis 12]
Input cubes are unit complex Gaussian noise generated on the device (profiling only; parity and
the bench use the seeded target cubes)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import rsp_b200 as rsp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--cpis", type=int, default=12)
ap.add_argument("--pool", type=int, default=4)
ap.add_argument("--range", action="store_true", help="bracket the enqueue with cudaProfilerStart/Stop (ncu --replay-mode range)")
a = ap.parse_args()
config, cfar_params, _ = rsp.named_config(a.config)
pd = rsp.build_precomputed_data(config)
chain = rsp.RadarChain(config, cfar_params, pd)
g = torch.Generator(device="cuda").manual_seed(0)
pool = torch.view_as_complex(torch.randn((a.pool, chain.P, chain.C, chain.N, 2), device="cuda", generator=g)
                             * (0.5 ** 0.5)).contiguous()
nr = 2 * chain.info()["lanes"]
rdm = torch.empty((nr, chain.B, chain.G, chain.P), dtype=torch.complex64, device="cuda")
torch.cuda.synchronize()
chain.set_stream(torch.cuda.current_stream().cuda_stream)
if a.range:
    chain.stream_enqueue(pool.data_ptr(), a.pool, rdm.data_ptr(), nr, a.cpis, 0)      # warm-up outside the range
    chain.synchronize()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
chain.stream_enqueue(pool.data_ptr(), a.pool, rdm.data_ptr(), nr, a.cpis, 0)
chain.synchronize()
if a.range:
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done", chain.info())
