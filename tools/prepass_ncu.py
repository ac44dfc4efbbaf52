"""The pre-pass gather kernels on a 3840x2160 picture (480 CTUs / 32,400 16-px CUs), for an ncu launch list:
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:picture --csv python tools/prepass_ncu.py"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg  # noqa: E402
from fastintercu_vvc_b200.synth import make_cu_state_dict, make_state_dict  # noqa: E402

rng = np.random.RandomState(0)
w, h = 3840, 2160
org = rng.randint(0, 1024, (h, w)).astype(np.int16)
ref = rng.randint(0, 1024, (h, w)).astype(np.int16)
blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
pkg.write_blob(make_state_dict(10), blob)
with pkg.MltPredictor(blob, max_batch=480) as p:
    p.begin_picture(org, 1)
    n = p.picture_ctu_count()
    for mv in (None, rng.randint(-16, 17, (n, 2)).astype(np.int16), (rng.randint(-2, 3, (n, 2)) * 8).astype(np.int16)):
        assert len(p.predict_picture(ref, 32, mv=mv)) == n
for size in (64, 16):
    pkg.write_cu_blob(make_cu_state_dict(10, size), size, blob)
    n = (w // size) * (h // size)
    with pkg.MltCuPredictor(blob, size, max_batch=n) as p:
        for mv in (None, rng.randint(-16, 17, (n, 2)).astype(np.int16)):
            assert len(p.predict_picture(org, ref, 1, 32, mv=mv)) == n
os.unlink(blob)
print("ok")
