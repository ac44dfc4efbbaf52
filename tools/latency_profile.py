"""Per-kernel device times of the CTU path at small batches (CUDA events inside the library): python tools/latency_profile.py"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg
from fastintercu_vvc_b200.synth import make_state_dict, synth_ctus
blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
pkg.write_blob(make_state_dict(10), blob)
base, pq = synth_ctus(16, 5)
with pkg.MltPredictor(blob, max_batch=960) as p:
    for n in (1, 8, 120, 480, 960):
        idx = np.arange(n) % 16
        op, q = np.ascontiguousarray(base[idx]), np.ascontiguousarray(pq[idx])
        for _ in range(3):
            p.predict_batch_dense(op, q)
        p.set_profiling(True)
        acc = np.zeros(18)
        for _ in range(10):
            p.predict_batch_dense(op, q)
            acc += p.get_profile()
        p.set_profiling(False)
        acc /= 10
        print(f"n={n:4d} total {acc.sum()*1e3:7.1f} us | stem {acc[0]*1e3:6.1f} | L0 {' '.join('%5.1f' % (x*1e3) for x in acc[2:5])} | L1 {' '.join('%5.1f' % (x*1e3) for x in acc[5:9])} | "
              f"L2 {' '.join('%5.1f' % (x*1e3) for x in acc[9:13])} | L3 {' '.join('%5.1f' % (x*1e3) for x in acc[13:17])} | head {acc[17]*1e3:5.1f}")
