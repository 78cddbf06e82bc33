for cps in 16 32 64 128 256; do
RSP_STREAM_SLOTS=512 python bench.py --config cfg2 --steps 8 --warmup 3 --no-cpu-baseline --no-extras --cpis-per-step 1024 --cpis-per-batch $cps 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cps', d['config']['cpis_per_batch'], 'CPI/s %.0f us/CPI %.2f' % (d['value'], 1e6/d['value']), d['clocks']['sm_mhz'])"
done
