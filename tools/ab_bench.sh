#!/bin/bash
# usage: [CONFIG=cfg3] [STEPS=6] tools/ab_bench.sh "ENV1=a ENV2=b" "ENV1=c" ...   -> one line per variant: CPI/s, us/CPI, per-kernel ms
for v in "$@"; do
  env $v python bench.py --config ${CONFIG:-cfg2} --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline ${EXTRA:---no-extras} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('${CONFIG:-cfg2}', '$v', '| CPI/s %.0f | us/CPI %.2f | frac %.3f |' % (d['value'], 1e6/d['value'], d['chain_roofline']['frac']), d['roofline'].get('kernels_ms_per_cpi'), '| frames/s %.0f' % ((d.get('e2e_targets') or {}).get('value',0)))
"
done
