#!/usr/bin/env python
"""Generate tests/golden/* by running the REFERENCE's own code in this container.

  * logits_seed10.npz : outputs of the reference architecture file
        /root/reference/mlt-cnn-python/codes/models/archs/mlt_ctu_or_pq_arch.py
        (`GapBigMltCtuORPQ()`, imported by path, eval mode, torch CPU fp32) on seeded synthetic CTUs
        with the seeded parameters of oracle/ref_arch.make_state_dict(10); plus the traced
        TorchScript module's outputs (traced exactly as model2torchScript.py:37-48) at B=1 with int
        poc/qp tensors as the hook passes them (EncCu.cpp:881-882).
  * stage_kat.npz : OpenCV's `Mat::convertTo(CV_32F, 1/1023)` on all 1024 10-bit codes (and a few
        out-of-range 16-bit codes), reached through cv2.normalize(..., NORM_MINMAX, dtype=CV_32F)
        (SURVEY.md section 8c), i.e. the arithmetic of EncCu.cpp:835-838.

The reference cannot travel to the GPU box, so the vectors are committed.  Re-run:
    python tools/gen_golden.py
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_arch  # noqa: E402

REF_ARCH = "/root/reference/mlt-cnn-python/codes/models/archs/mlt_ctu_or_pq_arch.py"
N_GOLDEN = 24
SEED = 10


def load_reference_arch():
    spec = importlib.util.spec_from_file_location("ref_mlt_ctu_or_pq_arch", REF_ARCH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    mod = load_reference_arch()
    sd = ref_arch.make_state_dict(SEED)
    net = mod.GapBigMltCtuORPQ()
    missing = net.load_state_dict(ref_arch.to_torch_state_dict(sd), strict=True)
    print("reference arch loaded:", missing)
    net.eval()

    orgpred, pocqp = ref_arch.synth_ctus(N_GOLDEN, SEED)
    x = ref_arch.stage_numpy(orgpred)
    logits = ref_arch.forward_logits(net, x, pocqp, batch=1)

    # the restatement must be the same function
    mine = ref_arch.forward_logits(ref_arch.build_model(sd), x, pocqp, batch=1)
    print("restated arch vs reference arch: max|d| =", np.abs(mine - logits).max())
    assert np.array_equal(mine, logits), "oracle/ref_arch.py diverges from the reference arch"

    # TorchScript trace exactly as model2torchScript.py:37-48 (float example inputs), called with
    # int tensors as the hook does
    ex = (torch.cat((torch.rand(1, 1, 128, 128), torch.rand(1, 1, 128, 128)), 1), torch.rand(1), torch.rand(1))
    traced = torch.jit.trace(net, ex)
    tl = []
    with torch.no_grad():
        for i in range(N_GOLDEN):
            o = traced(torch.from_numpy(x[i : i + 1]), torch.tensor([int(pocqp[i, 0])]), torch.tensor([int(pocqp[i, 1])]))
            tl.append(torch.cat(o, 1).numpy())
    tl = np.concatenate(tl, 0)
    print("traced vs eager: max|d| =", np.abs(tl - logits).max())
    split = logits[:, 5:9].argmax(1).astype(np.int32)
    print("level-3 class histogram:", np.bincount(split, minlength=4))

    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(
        os.path.join(out, "logits_seed10.npz"),
        seed=np.int32(SEED),
        n=np.int32(N_GOLDEN),
        pocqp=pocqp,
        logits=logits.astype(np.float32),
        logits_traced=tl.astype(np.float32),
        split=split,
        orgpred_sha256=np.frombuffer(hashlib.sha256(orgpred.tobytes()).digest(), np.uint8),
        staged_sha256=np.frombuffer(hashlib.sha256(x.tobytes()).digest(), np.uint8),
        params_sha256=np.frombuffer(
            hashlib.sha256(b"".join(np.ascontiguousarray(sd[k]).tobytes() for k in sorted(sd))).digest(), np.uint8
        ),
        torch_version=np.bytes_(torch.__version__),
    )

    # ---- OpenCV convertTo KAT (EncCu.cpp:835-838)
    import cv2

    codes = np.arange(1024, dtype=np.uint16)
    kat = []
    for width in (1024, 1021, 7):  # SIMD body and scalar tails
        src = np.resize(codes, (1, width)).astype(np.uint16)
        # normalize(min 0, max 1023 -> 0..1) calls src.convertTo(dst, CV_32F, 1/1023, 0)
        if src.min() != 0 or src.max() != 1023:
            src = np.concatenate([src, np.array([[0, 1023]], np.uint16)], axis=1)
        dst = cv2.normalize(src, None, 0, 1, cv2.NORM_MINMAX, dtype=cv2.CV_32F)
        kat.append((src.ravel().copy(), dst.ravel().copy()))
    # all widths must agree per code
    table = np.full(1024, np.nan, np.float32)
    for s, d in kat:
        for c, v in zip(s, d):
            if np.isnan(table[c]):
                table[c] = v
            assert table[c] == v
    assert not np.isnan(table).any()
    formula = codes.astype(np.float32) * np.float32(1.0 / 1023)
    print("cv2 convertTo == float(v)*float32(1/1023) on all 1024 codes:", np.array_equal(table, formula))
    print("alpha bits:", hex(np.float32(1.0 / 1023).view(np.uint32)))
    np.savez_compressed(
        os.path.join(out, "stage_kat.npz"),
        codes=codes,
        cv2_convert_to=table,
        cv2_version=np.bytes_(cv2.__version__),
    )
    print("wrote", out)


if __name__ == "__main__":
    main()
