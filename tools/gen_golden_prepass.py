"""Golden vectors of the frame-level pre-pass (gather + block matching) from the CPU oracle restatement -> tests/golden/prepass_seed10.npz.
The reference has no such pass, so these pin the restatement and the CUDA path against
regressions only; the planes are regenerated from the seed by the tests (prepass_planes below)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.oracle_lib import picture_ctus, picture_me, picture_pred  # noqa: E402


def prepass_planes(seed: int = 10, w: int = 392, h: int = 264):
    rng = np.random.RandomState(seed)
    coarse = rng.randint(0, 1024, (h // 8 + 2, w // 8 + 2)).repeat(8, 0).repeat(8, 1)
    org = (coarse[:h, :w] * 3 // 4 + rng.randint(0, 256, (h, w))).astype(np.int16)
    ref = np.clip(np.roll(org, (3, -2), (0, 1)).astype(np.int32) + rng.randint(-4, 5, (h, w)), 0, 1023).astype(np.int16)
    return org, ref


if __name__ == "__main__":
    org, ref = prepass_planes()
    h, w = org.shape
    xy = picture_ctus(w, h)
    out = {"xy": xy}
    for R in (0, 4, 9):
        res = [picture_me(org, ref, x, y, R) for x, y in xy]
        out[f"mv_r{R}"] = np.array([m for m, _ in res], np.int16)
        out[f"cost_r{R}"] = np.array([c for _, c in res], np.uint32)
    mvs = np.array([[0, 0], [-2, 3], [7, -1], [-300, 200], [129, 129], [-8, 16]], np.int16)[: len(xy)]
    out["mv_fixed"] = mvs
    out["pred_sha256"] = np.array([hashlib.sha256(picture_pred(ref, x, y, *mv).tobytes()).hexdigest() for (x, y), mv in zip(xy, mvs)])
    out["cu16_pred_sha256"] = np.array(hashlib.sha256(b"".join(picture_pred(ref, (i % (w // 16)) * 16, (i // (w // 16)) * 16, i % 7 - 3, i % 5 - 2, size=16).tobytes()
                                                               for i in range((w // 16) * (h // 16)))).hexdigest())
    path = os.path.join(ROOT, "tests", "golden", "prepass_seed10.npz")
    np.savez_compressed(path, **out)
    print(path, {k: (v.shape, v.dtype) for k, v in out.items()})
