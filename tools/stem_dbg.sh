#!/bin/bash
python - <<'PY'
import sys; sys.path.insert(0, '.')
from fastintercu_vvc_b200.pack_weights import write_blob
from fastintercu_vvc_b200.synth import make_state_dict
write_blob(make_state_dict(10), '/tmp/dbg.mltw')
PY
for st in 8 4; do echo "== n=900 stagers $st"; MLT_STEM5_STAGERS=$st timeout -s KILL 25 python tools/stem_dbg.py 900 /tmp/dbg.mltw 2>&1 | tail -1; done
STEM_AB_VARIANTS="MLT_STEM5_STAGERS=8;MLT_STEM5_STAGERS=4;MLT_STEM5_STAGERS=8,MLT_STEM5_CTAS=1" timeout -s KILL 200 python tools/stem_ab.py
