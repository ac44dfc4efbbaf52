#!/usr/bin/env python
"""Generate tests/golden/cu_logits_seed10.npz by running the REFERENCE's own smaller-CU architecture file here.

  /root/reference/mlt-cnn-python/codes/models/archs/mlt_cu_or_pq_arch.py : `GapBigMltCuORPQ()` (the model
  model2torchScript.py:22 exports for the 64 / 32 / 16-px CUs), imported by path, eval mode, torch CPU fp32, on
  seeded synthetic CUs with the seeded per-size parameters of synth.make_cu_state_dict(10, size); plus the traced
  TorchScript module's outputs (traced with 128x128 example inputs exactly as model2torchScript.py:37-48, then
  called at the CU's own size with int poc/qp like the hook, EncCu.cpp:881-882,909).

The reference cannot travel to the GPU box, so the vectors are committed.  Re-run: python tools/gen_golden_cu.py
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

REF_ARCH = "/root/reference/mlt-cnn-python/codes/models/archs/mlt_cu_or_pq_arch.py"
N_GOLDEN = 48
SEED = 10


def main():
    torch.set_num_threads(8)
    spec = importlib.util.spec_from_file_location("ref_mlt_cu_or_pq_arch", REF_ARCH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {"seed": np.int32(SEED), "n": np.int32(N_GOLDEN), "torch_version": np.bytes_(torch.__version__)}
    for size in ref_arch.CU_SIZES:
        sd = ref_arch.make_cu_state_dict(SEED, size)
        net = mod.GapBigMltCuORPQ()
        net.load_state_dict(ref_arch.to_torch_state_dict(sd), strict=True)
        net.eval()
        orgpred, pocqp = ref_arch.synth_cus(N_GOLDEN, size, SEED)
        x = ref_arch.stage_numpy(orgpred)
        logits = ref_arch.forward_cu_logits(net, x, pocqp, batch=1)
        mine = ref_arch.forward_cu_logits(ref_arch.build_cu_model(sd), x, pocqp, batch=1)
        assert np.array_equal(mine, logits), "oracle/ref_arch.MltCuNet diverges from the reference arch"
        ex = (torch.cat((torch.rand(1, 1, 128, 128), torch.rand(1, 1, 128, 128)), 1), torch.rand(1), torch.rand(1))
        traced = torch.jit.trace(net, ex)
        tl = []
        with torch.no_grad():
            for i in range(N_GOLDEN):
                o = traced(torch.from_numpy(x[i : i + 1]), torch.tensor([int(pocqp[i, 0])]), torch.tensor([int(pocqp[i, 1])]))
                tl.append(torch.cat(o, 1).numpy())
        tl = np.concatenate(tl, 0)
        print(size, "traced vs eager max|d| =", np.abs(tl - logits).max(), "level-1 histogram", np.bincount(logits[:, :2].argmax(1), minlength=2))
        out[f"logits_{size}"] = logits.astype(np.float32)
        out[f"logits_traced_{size}"] = tl.astype(np.float32)
        out[f"pocqp_{size}"] = pocqp
        out[f"split_{size}"] = logits[:, :2].argmax(1).astype(np.int32)  # what the hook uses below 128: elements()[0] (EncCu.cpp:916-919)
        out[f"orgpred_sha256_{size}"] = np.frombuffer(hashlib.sha256(orgpred.tobytes()).digest(), np.uint8)
        out[f"params_sha256_{size}"] = np.frombuffer(
            hashlib.sha256(b"".join(np.ascontiguousarray(sd[k]).tobytes() for k in sorted(sd))).digest(), np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "cu_logits_seed10.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
