#!/usr/bin/env python
"""How much of each kernel's time is ramp-up / tail rather than steady state?

DBF and pulse compression treat pulses independently, so a cube with K x 64 pulses is exactly a batch of K
config-2 CPIs for those two kernels.  Per-kernel device time (one CPI at a time, one lane) divided by K, for
K = 1, 2, 4, shows what a batched launch would buy.
    python tools/batch_probe.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import rsp_b200 as rsp  # noqa: E402

out = {}
for K in (1, 2, 4):
    config, cfar_params, _ = rsp.default_config(channel_num=16, beam_num=8, prtNum=64 * K, point_PRT=8192)
    pd = rsp.build_precomputed_data(config)
    chain = rsp.RadarChain(config, cfar_params, pd)
    g = torch.Generator(device="cuda").manual_seed(0)
    npool = 3
    pool = torch.view_as_complex(torch.randn((npool, chain.P, chain.C, chain.N, 2), device="cuda", generator=g) * (0.5 ** 0.5)).contiguous()
    rdm = torch.empty((3, chain.B, chain.G, chain.P), dtype=torch.complex64, device="cuda")
    torch.cuda.synchronize()
    chain.set_stream(torch.cuda.current_stream().cuda_stream)
    in_b, out_b = pool[0].numel() * 8, rdm[0].numel() * 8
    for i in range(6):
        chain.stream_enqueue(pool.data_ptr() + (i % npool) * in_b, 1, rdm.data_ptr() + (i % 3) * out_b, 1, 1, i)
    chain.synchronize()
    chain.set_profiling(True)
    for i in range(12):
        chain.stream_enqueue(pool.data_ptr() + (i % npool) * in_b, 1, rdm.data_ptr() + (i % 3) * out_b, 1, 1, i)
    kt = chain.kernel_times()
    chain.set_profiling(False)
    per = {k: round(v[0] / max(v[1], 1) * 1e3 / K, 2) for k, v in kt.items()}
    # overlapped throughput
    n = 24
    chain.stream_enqueue(pool.data_ptr(), npool, rdm.data_ptr(), 3, n, 0)
    chain.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    chain.stream_enqueue(pool.data_ptr(), npool, rdm.data_ptr(), 3, n, 0)
    chain.synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[f"K={K}"] = {"us_per_cpi_equiv": per, "sum": round(sum(per.values()), 2), "stream_us_per_cpi_equiv": round(dt / n / K * 1e6, 2)}
    print(f"K={K}", out[f"K={K}"], flush=True)
    chain.close()
    del pool, rdm
print(json.dumps(out))
