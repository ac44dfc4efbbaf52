"""Minimal stem5 run under a timeout (debugging aid): python tools/stem_dbg.py <n_ctus>"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg
from fastintercu_vvc_b200.synth import make_state_dict, synth_ctus
n = int(sys.argv[1])
blob = sys.argv[2]
base, pq = synth_ctus(8, 5)
idx = np.arange(n) % 8
with pkg.MltPredictor(blob, max_batch=max(n, 8)) as p:
    r = p.predict_batch_dense(np.ascontiguousarray(base[idx]), np.ascontiguousarray(pq[idx]))
    print("n", n, "ok", r["split_l3"][:4].tolist(), flush=True)
