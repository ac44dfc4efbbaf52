#!/bin/bash
# e2e (host-buffer) throughput of the CTU path for several chunk schedules (MLT_CHUNKS override); n = 3840
for sch in "" "480,1440,1920" "960,960,1920" "1920,1920" "640,1280,1920" "240,600,1080,1920" "480,960,1200,1200"; do
  MLT_CHUNKS=$sch timeout 200 python bench.py --steps 40 --cu-frames 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$sch'.ljust(24), 'device %.0f  pipelined %.0f  sync %.0f  clk %s' % (d['value'], d['e2e']['value'], d['e2e']['sync_call_value'], d['clocks']['sm_mhz']))"
done
