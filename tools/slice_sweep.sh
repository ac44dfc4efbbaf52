#!/bin/bash
for k in 2 4 6 8; do
  MLT_SLICES=$k timeout 200 python bench.py --steps 60 --cu-frames 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('slices $k: device %.0f CTU/s  %.3f ms  clk %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))"
done
