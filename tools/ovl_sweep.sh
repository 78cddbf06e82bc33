#!/bin/bash
# co-residency experiment: a persistent multi-CPI tcgen05 DBF (context A) beside PC + MTD + CFAR of another context (B)
V=radar-signal-simulation-and-target-detection_b200/lib/variants
export RSP_EXP_DBF_MULTI=4
for lib in probes80 probes; do
export RSP_LIBRARY=$V/librsp_$lib.so
echo "== $lib"
for co in 0 100 86; do
echo "-- carveout $co"
RSP_CARVEOUT=$co python tools/overlap_probe.py 1 14 A:RSP_LANES=1 2>&1 | tail -1
RSP_CARVEOUT=$co python tools/overlap_probe.py 1 14 A:RSP_LANES=1 --a-high 2>&1 | tail -1
RSP_CARVEOUT=$co python tools/overlap_probe.py 1 14 A:RSP_LANES=1 A:RSP_TC_STAGES=2 2>&1 | tail -1
RSP_CARVEOUT=$co python tools/overlap_probe.py 1 2 A:RSP_LANES=1 2>&1 | tail -1
done
done
