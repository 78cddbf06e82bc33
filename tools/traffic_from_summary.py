#!/usr/bin/env python
"""profiles/traffic.json (read by bench.py: roofline.traffic, roofline_issue) from a tools/ncu_summary.py CSV of ONE CPI's launches:
    python tools/traffic_from_summary.py profiles/r2h_ncu_full_summary_cfg2.csv cfg2"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, cfg = sys.argv[1], sys.argv[2]
rows = list(csv.DictReader(open(src)))
def num(v):
    v = v.strip()
    mult = {"u": 1e-6, "m": 1e-3, "K": 1e3, "M": 1e6, "G": 1e9}.get(v[-1], None) if v and not v[-1].isdigit() else None
    return float(v[:-1]) * mult if mult else float(v)
classes = {"dbf": "dbf", "pc_fft": "pc_fft", "mtd": "mtd", "cfar": "cfar", "refine": "refine", "synth": "synth"}
out = {}
for r in rows:
    name = r["kernel"].replace("rsp::", "")
    k = next((c for key, c in classes.items() if name.startswith(key)), None)
    if not k:
        continue
    e = out.setdefault(k, {"dram_bytes": 0, "warp_instructions": 0, "launches": 0, "cold_us": 0.0})
    e["dram_bytes"] += int(num(r["dramR"]) + num(r["dramW"]))
    e["warp_instructions"] += int(num(r["winst"]))
    e["launches"] += 1
    e["cold_us"] = round(e["cold_us"] + num(r["dur"]) * 1e6, 2)
out["warp_instructions_per_cpi"] = sum(e["warp_instructions"] for k, e in out.items() if isinstance(e, dict) and k != "refine")
path = os.path.join(ROOT, "profiles", "traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[cfg] = out
data["_note"] = (f"per CPI and kernel class from {os.path.relpath(src, ROOT)} (ncu --set full --clock-control none, one CPI on one lane, cold cache: every "
                 "replay pass flushes L2, so intermediates that are L2 hits in the running chain count as DRAM reads here); dram_bytes = "
                 "dram__bytes_read.sum + dram__bytes_write.sum, warp_instructions = smsp__inst_executed.sum (mbarrier polls of the tcgen05 DBF's role "
                 "warps included); pc_fft = both launches of the mixed block plan; refine is one launch per batch and not part of the per-CPI sum")
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
