#!/usr/bin/env python
"""Opcode mix + hottest SASS lines of one kernel from an .ncu-rep: python tools/ncu_opmix.py rep kernel_regex [top]"""
import collections, csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
h = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) > 10 and r[0].startswith("0x")]
iS, iN, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
tot, totE = sum(int(r[iN]) for r in data), sum(int(r[iE]) for r in data)
mix, smp = collections.Counter(), collections.Counter()
for r in data:
    parts = r[iS].strip().split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    mix[op] += int(r[iE]); smp[op] += int(r[iN])
print(f"{pat}: sass lines {len(data)}, warp instr {totE}, samples {tot}")
for op, c in mix.most_common(top):
    print(f"  {op:10s} exec {c / totE * 100:5.1f}%  samples {smp[op] / max(tot, 1) * 100:5.1f}%")
print("hottest lines:")
for r in sorted(data, key=lambda r: -int(r[iN]))[:top]:
    print(f"  {int(r[iN]) / max(tot, 1) * 100:5.1f}%  {r[iS].strip()[:100]}")
