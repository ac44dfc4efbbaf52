"""Per-logit mean / spread of the product path's error against the fp32 cross-check engine, for packer / kernel variants
(run on the B200 box): python tools/logit_error_stats.py [n=4096]"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg
from fastintercu_vvc_b200 import pack_weights as pw
from oracle import ref_arch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sd = ref_arch.make_state_dict(10)
ctus, pq = ref_arch.synth_ctus(n, 555000)
ref = None
for name, corr in (("bias correction ON", True), ("bias correction OFF", False)):
    blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
    secs_fn = pw.build_sections
    pw.build_sections = lambda s, calib=None, correct_bias=True, _f=secs_fn, _c=corr: _f(s, calib, _c)
    pw.write_blob(sd, blob)
    pw.build_sections = secs_fn
    with pkg.MltPredictor(blob, device=0, max_batch=2048) as p:
        if ref is None:
            p.set_engine(1)
            ref = np.concatenate([p.predict_batch_dense(ctus[i:i + 2048], pq[i:i + 2048])["logits"] for i in range(0, n, 2048)])
            p.set_engine(0)
        got = np.concatenate([p.predict_batch_dense(ctus[i:i + 2048], pq[i:i + 2048])["logits"] for i in range(0, n, 2048)])
    os.unlink(blob)
    d = (got - ref).astype(np.float64)
    flips = [(got[:, a:b].argmax(1) != ref[:, a:b].argmax(1)).sum() for a, b in ((0, 2), (2, 5), (5, 9))]
    print(f"{name} [{os.environ.get('MLT_STEM_OLD') and 'old stem' or 'stem5'}]: mean x1e4 {np.round(d.mean(0) * 1e4, 2)}\n   std x1e4 {np.round(d.std(0) * 1e4, 2)}  flips {flips}/{n}", flush=True)
