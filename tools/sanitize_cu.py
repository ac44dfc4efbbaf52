"""Small end-to-end run of every kernel (CTU + the three CU sizes) for compute-sanitizer: python tools/sanitize_cu.py"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg
from fastintercu_vvc_b200.synth import make_cu_state_dict, make_state_dict, synth_ctus, synth_cus
blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
pkg.write_blob(make_state_dict(10), blob)
op, pq = synth_ctus(5, 5)
with pkg.MltPredictor(blob, max_batch=8) as p:
    r = p.predict_batch_dense(op, pq)
    p.submit_batch_dense(op, pq); r2 = p.collect()
    assert r.tobytes() == r2.tobytes()
print("ctu ok", r["split_l3"].tolist())
for size in (64, 32, 16):
    pkg.write_cu_blob(make_cu_state_dict(10, size), size, blob)
    cus, cq = synth_cus(70, size, 3)  # ragged against every images-per-tile count
    with pkg.MltCuPredictor(blob, size, max_batch=96) as p:
        r = p.predict_batch_dense(cus, cq)
        one = p.predict(cus[69, 0], cus[69, 1], cq[69, 0], cq[69, 1])
        assert one.tobytes() == r[69].tobytes()
    print("cu", size, "ok", r["split"][:6, 0].tolist())
os.unlink(blob)
