"""CTU model through engine 2 (standalone conv1 + conv 0 via the generic tcgen05 kernel) at a large batch."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg
from fastintercu_vvc_b200.synth import make_state_dict, synth_ctus
n = int(sys.argv[1]); eng = int(sys.argv[2])
blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
pkg.write_blob(make_state_dict(10), blob)
base, pq = synth_ctus(16, 5)
idx = np.arange(n) % 16
op, pq = np.ascontiguousarray(base[idx]), np.ascontiguousarray(pq[idx])
with pkg.MltPredictor(blob, max_batch=n) as p:
    ref = p.predict_batch_dense(op[:16], pq[:16])
    p.set_engine(eng)
    r = p.predict_batch_dense(op[:900], pq[:900]) if n > 900 else None
    r = p.predict_batch_dense(op, pq)
print("ok", n, eng, bool((r["logits"][:16] == r["logits"][16:32]).all()), float(np.abs(r["logits"][:16] - ref["logits"]).max()))
