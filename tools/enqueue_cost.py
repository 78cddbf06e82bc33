"""How long does the host take to enqueue a batch?  (CPU launch-bound check)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rsp_b200 as rsp
config, cfar_params, _ = rsp.named_config("cfg2")
pd = rsp.build_precomputed_data(config)
chain = rsp.RadarChain(config, cfar_params, pd)
pool = torch.view_as_complex(torch.randn((4, chain.P, chain.C, chain.N, 2), device="cuda") * 0.7071).contiguous()
rdm = torch.empty((6, chain.B, chain.G, chain.P), dtype=torch.complex64, device="cuda")
s = torch.cuda.Stream(); torch.cuda.set_stream(s); chain.set_stream(s.cuda_stream)
for _ in range(3):
    chain.stream_enqueue(pool.data_ptr(), 4, rdm.data_ptr(), 6, 64, 0)
torch.cuda.synchronize()
for n in (64, 64, 128):
    t0 = time.perf_counter()
    for _ in range(n // 64):
        chain.stream_enqueue(pool.data_ptr(), 4, rdm.data_ptr(), 6, 64, 0)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{n} CPIs: host enqueue {1e6*(t1-t0)/n:.1f} us/CPI, total {1e6*(t2-t0)/n:.1f} us/CPI")
