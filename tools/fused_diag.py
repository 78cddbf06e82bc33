"""Diagnostic: where does the fused dbf_pc beam / pc cube differ from the two-kernel path?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import rsp_b200 as rsp
from conftest import oracle as o

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg, pre, raw = o.make_cube(name, 2)
config, cfar_params, cluster_params = rsp.named_config(name)
pd = rsp.build_precomputed_data(config)
out = {}
for fused in ("1", "0"):
    os.environ["RSP_FUSE_DBF_PC"] = fused
    chain = rsp.RadarChain(config, cfar_params, pd)
    for rep in range(2):
        chain.process_cpi(raw)
        out[(fused, rep)] = (chain.get_beam(), chain.get_pc())
    print("fused", fused, chain.info())
    chain.close()
for rep in range(2):
    for idx, what in ((0, "beam"), (1, "pc")):
        a, b = out[("1", rep)][idx], out[("0", 0)][idx]
        bad = np.argwhere(a != b)
        print(what, "rep", rep, "mismatches", len(bad), "of", a.size)
        if len(bad):
            ps, bs, ns = bad[:, 0], bad[:, 1], bad[:, 2]
            print("  pulses", np.unique(ps)[:20], "beams", np.unique(bs), "n range", ns.min(), ns.max())
            print("  n % 128 hist", np.bincount(ns % 128, minlength=128))
            print("  n // 128 uniq", np.unique(ns // 128)[:70])
            print("  first", bad[:5].tolist(), a[tuple(bad[0])], b[tuple(bad[0])])
            za = (a[tuple(bad.T)] == 0).sum()
            print("  fused value is zero at", za)


def trace_report(flags):
    os.environ["RSP_FUSE_DBF_PC"] = "1"
    os.environ["RSP_FUSED_DEBUG"] = str(flags)
    chain = rsp.RadarChain(config, cfar_params, pd)
    for rep in range(3):
        chain.process_cpi(raw)
    tr = chain.fused_trace().astype(np.float64)
    chain.close()
    os.environ.pop("RSP_FUSED_DEBUG")
    if not len(tr):
        print("no trace"); return
    t0 = tr[:, 0].min()
    names = ["start", "cluster_up", "dbf_done", "lines_ready", "round0", "round1", "round2"]
    print(f"--- trace flags={flags}: kernel span {(tr[:, 1:7].max() - t0) / 1e3:.1f} us, {len(np.unique(tr[:, 7]))} SMs, max CTAs/SM "
          f"{np.bincount(tr[:, 7].astype(int)).max()}")
    for i, nm in enumerate(names):
        col = tr[:, i]
        if col.max() == 0: continue
        print(f"  {nm:12s} min {(col.min() - t0) / 1e3:7.2f}  median {(np.median(col) - t0) / 1e3:7.2f}  max {(col.max() - t0) / 1e3:7.2f} us")
    d = tr[:, 1:7] - tr[:, 0:6]
    for i, nm in enumerate(["wait cluster", "dbf loop", "cluster barrier", "round 0", "round 1", "round 2"]):
        if tr[:, i + 1].max() == 0: continue
        print(f"  phase {nm:16s} median {np.median(d[:, i]) / 1e3:6.2f}  p90 {np.percentile(d[:, i], 90) / 1e3:6.2f} us")
    starts = np.sort(tr[:, 0] - t0) / 1e3
    print("  CTA start times (us) every 64th:", np.round(starts[::64], 1))


for flags in (0, 1):
    trace_report(flags)

for pf in ("0",):
    os.environ["RSP_FUSED_PREFETCH"] = pf
    print("prefetch_ahead =", pf)
    trace_report(0)
    os.environ.pop("RSP_FUSED_PREFETCH")
