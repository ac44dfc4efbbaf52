import os, sys, tempfile, numpy as np
sys.path.insert(0, os.getcwd())
import fastintercu_vvc_b200 as pkg
from fastintercu_vvc_b200.pack_weights import write_blob
from fastintercu_vvc_b200.synth import make_state_dict, synth_ctus
blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
write_blob(make_state_dict(10), blob)
n = 960
base, pq = synth_ctus(16, 5)
idx = np.arange(n) % 16
op, pq = np.ascontiguousarray(base[idx]), np.ascontiguousarray(pq[idx])
with pkg.MltPredictor(blob, max_batch=n) as p:
    p.predict_batch_dense(op, pq)
    os.environ["MLT_TRACE_DUMP"] = "1"
    p.predict_batch_dense(op, pq)
