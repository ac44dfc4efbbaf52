"""Debug helper: run one batch with MLT_TRACE_LAYER set and print the per-tile pipeline timeline of CTA 0."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
layer = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 960
os.environ["MLT_TRACE_LAYER"] = str(layer)
os.environ["MLT_TRACE_FILE"] = f"gpurun_out/trace_l{layer}.txt"
import fastintercu_vvc_b200 as pkg
from fastintercu_vvc_b200.pack_weights import write_blob
from fastintercu_vvc_b200.synth import make_state_dict, synth_ctus
blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
write_blob(make_state_dict(10), blob)
base, pq = synth_ctus(16, 5)
idx = np.arange(n) % 16
op, pq = np.ascontiguousarray(base[idx]), np.ascontiguousarray(pq[idx])
with pkg.MltPredictor(blob, max_batch=n) as p:
    for _ in range(3):
        p.predict_batch_dense(op, pq)
rows = [l.split() for l in open(os.environ["MLT_TRACE_FILE"]) if not l.startswith("#")]
T = {}
for r in rows:
    T[(int(r[0]), int(r[1]))] = [int(x) for x in r[2:]]
t0 = min(v[0] for v in T.values() if v[0] > 0)
print("tile | prod: wait_start wait_done issued | mma: start accEmpty_ok fullA_ok committed | epi: start accFull_ok done")
for t in range(0, 40):
    pr, mm, ep = T[(0, t)], T[(1, t)], T[(2 + (t & 1), t)]
    f = lambda v: " ".join(f"{(x - t0) if x else -1:7d}" for x in v)
    print(f"{t:3d} | {f(pr[:3])} | {f(mm)} | {f(ep[:3])}")
