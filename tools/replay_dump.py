#!/usr/bin/env python
"""Replay the predictor inputs an encoder dumped (MLT_DUMP_INPUTS: records {int32 poc, int32 qp, int16 org[128*128], int16 pred[128*128]})
through the fp32 C oracle and through libmltcnn.so, next to the decisions the encoder itself traced (MLT_TRACE).  This is how a
differing decision inside a real encode is classified: the inputs up to the first difference are the same in both encoders, so
evaluating ONE encoder's dump with both arithmetics shows whether the difference is an fp32 tie.

  python tools/replay_dump.py oracle/_ref/data/ref_cpu_1080p_q27_inputs.bin oracle/_ref/data/ref_cpu_1080p_q27.json [--out x.json]
(test infrastructure: imports the oracle; needs a B200 for the library arm)"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

REC = 8 + 2 * 2 * 128 * 128


def load_dump(path: str):
    raw = np.fromfile(path, np.uint8)
    n = len(raw) // REC
    assert n * REC == len(raw), "truncated dump"
    recs = raw.reshape(n, REC)
    pocqp = recs[:, :8].copy().view(np.int32).reshape(n, 2)
    blocks = recs[:, 8:].copy().view(np.int16).reshape(n, 2, 128, 128)
    return blocks, pocqp


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("encode_json", help="vtm_run.run_encode result of the encoder that wrote the dump (its 'trace')")
    ap.add_argument("--out")
    a = ap.parse_args()
    import fastintercu_vvc_b200 as pkg
    from oracle import ref_arch
    from tests.oracle_lib import OracleModel

    blocks, pocqp = load_dump(a.dump)
    enc = json.load(open(a.encode_json))
    traced = np.array([t[4] for t in enc["trace"]])
    n = len(blocks)
    assert len(traced) == n, (len(traced), n)
    assert np.array_equal(pocqp[:, 0], [t[0] for t in enc["trace"]]) and np.array_equal(pocqp[:, 1], [t[3] for t in enc["trace"]])
    sd = ref_arch.make_state_dict(10)
    lg, sp = OracleModel(sd).predict_batch(blocks, pocqp)
    blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
    pkg.write_blob(sd, blob)
    with pkg.MltPredictor(blob, device=0, max_batch=max(n, 8)) as p:
        batch = p.predict_batch_dense(blocks, pocqp)
        single = np.array([p.predict_ctu(blocks[i, 0], blocks[i, 1], int(pocqp[i, 0]), int(pocqp[i, 1]))["split_l3"] for i in range(n)])
    os.unlink(blob)
    srt = np.sort(lg[:, 5:9], 1)
    margin = srt[:, -1] - srt[:, -2]
    rows = []
    for i in np.nonzero((batch["split_l3"] != traced) | (sp != traced))[0]:
        rows.append({"call": int(i), "poc": int(pocqp[i, 0]), "x": enc["trace"][i][1], "y": enc["trace"][i][2], "qp": int(pocqp[i, 1]),
                     "encoder_libtorch": int(traced[i]), "oracle_fp32": int(sp[i]), "libmltcnn": int(batch["split_l3"][i]),
                     "oracle_l3_logits": [float(v) for v in lg[i, 5:9]], "libmltcnn_l3_logits": [float(v) for v in batch["logits"][i, 5:9]],
                     "fp32_top2_margin": float(margin[i])})
    doc = {"calls": n, "encoder": enc["encoder"], "clip": enc["clip"], "qp": enc["qp"],
           "libmltcnn_equals_encoder": int((batch["split_l3"] == traced).sum()), "oracle_equals_encoder": int((sp == traced).sum()),
           "single_call_equals_batch": bool(np.array_equal(single, batch["split_l3"])),
           "max_abs_dlogit_vs_oracle": float(np.abs(batch["logits"] - lg).max()),
           "smallest_fp32_margins": [float(v) for v in np.sort(margin)[:5]], "differences": rows,
           "note": "inputs are the ones the reference-hook encoder saw; every listed call differs between at least two of "
                   "{libtorch inside the encoder, fp32 C oracle, libmltcnn}"}
    s = json.dumps(doc, indent=1)
    print(s)
    if a.out:
        open(a.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
