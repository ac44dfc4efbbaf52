"""Probability / decision parity of the CU models on a larger sample (run on the B200 box): python tools/precision_cu.py [n]"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg
from oracle import ref_arch
import torch
torch.set_num_threads(len(os.sched_getaffinity(0)))
LEVELS = ((0, 2), (2, 5), (5, 9), (9, 15))
def softmax_levels(lg):
    out = np.empty_like(lg)
    for a, b in LEVELS:
        e = np.exp(lg[:, a:b] - lg[:, a:b].max(1, keepdims=True)); out[:, a:b] = e / e.sum(1, keepdims=True)
    return out
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for size in (64, 32, 16):
    sd = ref_arch.make_cu_state_dict(10, size)
    cus, pq = ref_arch.synth_cus(n, size, 4242)
    blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
    pkg.write_cu_blob(sd, size, blob)
    with pkg.MltCuPredictor(blob, size, max_batch=n) as p:
        res = p.predict_batch_dense(cus, pq)
    os.unlink(blob)
    lg = ref_arch.forward_cu_logits(ref_arch.build_cu_model(sd), ref_arch.stage_numpy(cus), pq)
    dl = np.abs(res["logits"] - lg); dp = np.abs(res["probs"] - softmax_levels(lg))
    flips = [int((res["split"][:, l] != lg[:, a:b].argmax(1)).sum()) for l, (a, b) in enumerate(LEVELS)]
    print(f"size {size}: n={n} |dlogit| mean {dl.mean():.3e} p99 {np.percentile(dl, 99):.3e} max {dl.max():.3e} | |dprob| mean {dp.mean():.3e} "
          f"p99.9 {np.percentile(dp, 99.9):.3e} max {dp.max():.3e} | flips per level {flips}/{n}", flush=True)
