"""Compare weight-rounding modes of the packer on the GPU against the fp32 C oracle (run on the B200 box)."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg
from fastintercu_vvc_b200 import pack_weights as pw
from oracle import ref_arch
from tests.oracle_lib import OracleModel

def softmax_levels(lg):
    out = np.empty_like(lg)
    for a, b in ((0, 2), (2, 5), (5, 9)):
        e = np.exp(lg[:, a:b] - lg[:, a:b].max(1, keepdims=True)); out[:, a:b] = e / e.sum(1, keepdims=True)
    return out

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sd = ref_arch.make_state_dict(10)
orgpred, pocqp = ref_arch.synth_ctus(n, 4242)
lg, sp = OracleModel(sd).predict_batch(orgpred, pocqp)
for mode in ("nearest", "diffused"):
    pw.ROUNDING = mode
    with tempfile.NamedTemporaryFile(suffix=".mltw") as f:
        pw.write_blob(sd, f.name)
        with pkg.MltPredictor(f.name, device=0, max_batch=n) as p:
            res = p.predict_batch_dense(orgpred, pocqp)
    dl = np.abs(res["logits"] - lg)
    dp = np.abs(res["probs"] - softmax_levels(lg))
    flips = int((res["split_l3"] != sp).sum())
    print(f"{mode:9s}: |dlogit| mean {dl.mean():.3e} p99 {np.percentile(dl, 99):.3e} max {dl.max():.3e} | |dprob| mean {dp.mean():.3e} p99.9 {np.percentile(dp, 99.9):.3e} max {dp.max():.3e} | L3 flips {flips}/{n}")
