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
    # round 2: 10-bit packed transport, the one-CTU call (tiny-batch channel splits), a pair
    from fastintercu_vvc_b200.capi import pack10
    assert p.predict_batch_packed10(pack10(op), pq).tobytes() == r.tobytes()
    p.submit_batch_packed10(pack10(op), pq); assert p.collect().tobytes() == r.tobytes()
    assert p.predict_ctu(op[1, 0], op[1, 1], int(pq[1, 0]), int(pq[1, 1])).tobytes() == r[1].tobytes()
    assert p.predict_batch_dense(op[2:4], pq[2:4]).tobytes() == r[2:4].tobytes()
print("ctu ok", r["split_l3"].tolist())
if os.environ.get("MLT_CHAIN"):
    print("(cluster chain kernel was used for the 1- and 2-CTU calls)")
# frame-level pre-pass: gather kernels with MVs that hang over every picture border, odd width pitch, strided planes
rng = np.random.RandomState(1)
w, h = 264, 136  # 2 x 1 CTUs + a partial column / row; pitch 264
obuf, rbuf = rng.randint(0, 1024, (h + 4, w + 11)).astype(np.int16), rng.randint(0, 1024, (h + 4, w + 11)).astype(np.int16)
org, ref = obuf[2 : 2 + h, 3 : 3 + w], rbuf[2 : 2 + h, 3 : 3 + w]
with pkg.MltPredictor(blob, max_batch=8) as p:
    p.begin_picture(org, 2)
    for mv in (None, np.array([[-300, -200], [301, 199]], np.int16), np.array([[7, -1], [-9, 135]], np.int16)):
        r = p.predict_picture(ref, 30, mv=mv)
        assert len(r) == 2
    p.submit_batch_dense(op, pq); p.submit_batch_dense(op, pq); p.collect(); p.collect()
print("ctu picture pre-pass ok")
for size in (64, 32, 16):
    pkg.write_cu_blob(make_cu_state_dict(10, size), size, blob)
    cus, cq = synth_cus(70, size, 3)  # ragged against every images-per-tile count
    with pkg.MltCuPredictor(blob, size, max_batch=96) as p:
        r = p.predict_batch_dense(cus, cq)
        one = p.predict(cus[69, 0], cus[69, 1], cq[69, 0], cq[69, 1])
        assert one.tobytes() == r[69].tobytes()
        n = (w // size) * (h // size)
        if n <= 96:
            mv = rng.randint(-40, 41, (n, 2)).astype(np.int16)
            mv[0], mv[-1] = (-400, -300), (400, 300)
            assert len(p.predict_picture(org, ref, 2, 30, mv=mv)) == n and len(p.predict_picture(org, ref, 2, 30)) == n
        p.submit_batch_dense(cus, cq); p.submit_batch_dense(cus[:33], cq[:33])
        a, b = p.collect().copy(), p.collect().copy()
        assert a.tobytes() == r.tobytes() and b.tobytes() == r[:33].tobytes()
    print("cu", size, "ok", r["split"][:6, 0].tolist())
os.unlink(blob)
