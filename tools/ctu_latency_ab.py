"""Single-CTU hook latency (mlt_predict_ctu, the in-encoder call) + bit-identity of the call against the batch path.
Used in round 1 to A/B a CUDA-graph replay of the n = 1 call (switch MLT_NO_GRAPH; 159.5 vs 160.1 us, no gain -- the graph
path was removed again, see profiles/r01/README.md); kept as the latency probe."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg  # noqa: E402
from fastintercu_vvc_b200.synth import make_state_dict, synth_ctus  # noqa: E402

blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
pkg.write_blob(make_state_dict(10), blob)
ctus, pq = synth_ctus(8, 10)
with pkg.MltPredictor(blob, device=0, max_batch=160) as p:
    ref = p.predict_batch_dense(ctus, pq)
    for i in range(8):
        one = p.predict_ctu(ctus[i, 0], ctus[i, 1], int(pq[i, 0]), int(pq[i, 1]))
        assert one.tobytes() == ref[i].tobytes(), i
    for rep in range(3):
        t0 = time.perf_counter()
        for k in range(500):
            i = k & 7
            p.predict_ctu(ctus[i, 0], ctus[i, 1], int(pq[i, 0]), int(pq[i, 1]))
        print("graph" if not os.environ.get("MLT_NO_GRAPH") else "stream", f"{(time.perf_counter() - t0) / 500 * 1e6:.1f} us per mlt_predict_ctu", "launches", p.launch_count)
os.unlink(blob)
