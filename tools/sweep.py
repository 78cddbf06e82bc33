#!/usr/bin/env python
"""Run bench.py under several environment settings on the same box and print one line each.
   python tools/sweep.py "RSP_CPS_PC=1" "RSP_CPS_PC=2 RSP_LANES=4" ...   (use "-" for the defaults)"""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
extra = os.environ.get("SWEEP_BENCH_ARGS", "--no-cpu-baseline --e2e-cpis 2 --steps 40").split()
for spec in sys.argv[1:]:
    env = dict(os.environ)
    if spec != "-":
        for kv in spec.split():
            k, v = kv.split("=", 1)
            env[k] = v
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py")] + extra, env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{spec:45s} {d['value']:9.0f} CPI/s  frac {d['roofline']['frac']:.3f}  {d['roofline']['kernels_ms_per_cpi']}", flush=True)
    except Exception:
        print(f"{spec:45s} FAILED rc={r.returncode} {r.stderr[-300:]}", flush=True)
