#!/usr/bin/env python
"""Do kernels of two independent contexts overlap on the device?  Context A runs only stage mask MA, context B only MB
(RSP_STAGES is read at rsp_create); both are enqueued back to back on their own streams and timed together.
    python tools/overlap_probe.py 1 2      (DBF in one context, PC in the other)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import rsp_b200 as rsp  # noqa: E402

ma, mb = int(sys.argv[1]), int(sys.argv[2])
config, cfar_params, _ = rsp.named_config("cfg2")
pd = rsp.build_precomputed_data(config)
os.environ["RSP_STAGES"] = str(ma)
# per-context environment: A:KEY=VAL / B:KEY=VAL arguments (read at rsp_create)
envA = dict(a[2:].split("=", 1) for a in sys.argv[3:] if a.startswith("A:"))
envB = dict(a[2:].split("=", 1) for a in sys.argv[3:] if a.startswith("B:"))
os.environ.update(envA)
if "--a-high" in sys.argv:
    os.environ["RSP_STREAM_PRIO"] = "high"      # context A's streams get the greatest priority
A = rsp.RadarChain(config, cfar_params, pd)
os.environ.pop("RSP_STREAM_PRIO", None)
for k_ in envA:
    os.environ.pop(k_, None)
os.environ.update(envB)
os.environ["RSP_STAGES"] = str(mb)
B = rsp.RadarChain(config, cfar_params, pd)
g = torch.Generator(device="cuda").manual_seed(0)
npool = 4
pool = torch.view_as_complex(torch.randn((npool, A.P, A.C, A.N, 2), device="cuda", generator=g) * (0.5 ** 0.5)).contiguous()
rdmA = torch.empty((6, A.B, A.G, A.P), dtype=torch.complex64, device="cuda")
rdmB = torch.empty((6, A.B, A.G, A.P), dtype=torch.complex64, device="cuda")
n = 64


def run(which, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if "A" in which:
            A.stream_enqueue(pool.data_ptr(), npool, rdmA.data_ptr(), 6, n, 0)
        if "B" in which:
            B.stream_enqueue(pool.data_ptr(), npool, rdmB.data_ptr(), 6, n, 0)
    A.synchronize(); B.synchronize(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / (reps * n) * 1e6


for w in ("A", "B", "AB"):
    run(w, 2)
print(f"{' '.join(sys.argv[3:])} OCC_DBF={os.environ.get('RSP_OCC_DBF')} OCC_PC={os.environ.get('RSP_OCC_PC')} masks A={ma} B={mb} lanes={os.environ.get('RSP_LANES', '3')}: A alone {run('A'):.2f}  B alone {run('B'):.2f}  together {run('AB'):.2f} us per CPI pair")
