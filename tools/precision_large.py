"""Split-decision agreement and probability error of the tcgen05 product path at scale ("measure
agreement on >= 100 k random CTUs"): N seeded synthetic CTUs through engine 0 (product) and engine 1 (the fp32 CUDA-core
cross-check engine, itself anchored to the fp32 C oracle on the first 256 CTUs of the run).
usage (B200 box): python tools/precision_large.py [n_ctus=51200]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastintercu_vvc_b200 as pkg  # noqa: E402
from oracle import ref_arch  # noqa: E402
from tests.oracle_lib import OracleModel  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 51200
B = 2048
sd = ref_arch.make_state_dict(10)
blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
pkg.write_blob(sd, blob)
levels = ((0, 2, "split_l1"), (2, 5, "split_l2"), (5, 9, "split_l3"))
max_dp, sum_dp, cnt = 0.0, 0.0, 0
flips = {k: 0 for _, _, k in levels}
tie_flips = {k: 0 for _, _, k in levels}  # flips whose fp32 top-2 logit margin is below 5e-3
worst = []
t0 = time.time()
with pkg.MltPredictor(blob, device=0, max_batch=B) as p:
    done = 0
    while done < N:
        n = min(B, N - done)
        ctus, pq = ref_arch.synth_ctus(n, 900000 + done)
        p.set_engine(0)
        a = p.predict_batch_dense(ctus, pq).copy()
        p.set_engine(1)
        r = p.predict_batch_dense(ctus, pq).copy()
        if done == 0:
            lg, _ = OracleModel(sd).predict_batch(ctus[:256], pq[:256])
            print(f"anchor: fp32 GPU engine vs C oracle on 256 CTUs: max |dlogit| {np.abs(r['logits'][:256] - lg).max():.2e}")
        dp = np.abs(a["probs"] - r["probs"])
        max_dp = max(max_dp, float(dp.max()))
        sum_dp += float(dp.sum())
        cnt += dp.size
        for lo, hi, key in levels:
            bad = np.nonzero(a[key] != r[key])[0]
            flips[key] += len(bad)
            for i in bad:
                top2 = np.sort(r["logits"][i, lo:hi])[-2:]
                margin = float(top2[1] - top2[0])
                tie_flips[key] += margin < 5e-3
                worst.append((margin, key, done + int(i)))
        done += n
        print(f"{done:7d} CTUs  max |dprob| {max_dp:.3e}  flips L1/L2/L3 {flips['split_l1']}/{flips['split_l2']}/{flips['split_l3']}  ({time.time() - t0:.0f}s)", flush=True)
os.unlink(blob)
print(f"== {N} CTUs: max |dprob| {max_dp:.3e}, mean {sum_dp / cnt:.3e}")
for _, _, key in levels:
    print(f"   {key}: {flips[key]} flips = {100.0 * (1 - flips[key] / N):.4f} % agreement; {tie_flips[key]} of them inside a 5e-3 fp32 logit margin")
worst.sort(reverse=True)
print("   largest fp32 top-2 margins among the flips:", [(f"{m:.2e}", k, i) for m, k, i in worst[:5]])
