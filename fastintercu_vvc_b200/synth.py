"""Seeded synthetic data for tests and benches (plain numpy; no model code, no oracle code).

  * make_state_dict(seed): random parameters with the reference's state_dict keys and shapes
    (mlt_ctu_or_pq_arch.py:239-263; the trained MLTORPQ_splitMode_128.pt is not distributed);
  * synth_ctus(n, seed): synthetic (org, pred, poc, qp) CTUs in the spirit of BASELINE config 2;
  * make_cu_state_dict(seed, size) / synth_cus(n, size, seed): the same for the smaller-CU models
    (64 / 32 / 16-px `GapBigMltCuORPQ`, mlt_cu_or_pq_arch.py:59-130).

numpy's legacy RandomState is bit-stable across versions and machines, so the committed golden vectors
(tests/golden/) stay valid on the GPU box.
"""
from __future__ import annotations

import numpy as np

PLANES = (32, 64, 128, 256)
FC_IN = (64 + 2, 128 + 2, 256 + 2)
FC_OUT = (2, 3, 4)
HEAD_CENTER = ((-0.39, -2.15), (2.03, -1.08, -3.66), (-1.03, -1.96, -3.60, -2.98))


# --------------------------------------------------------------------------- seeded parameters


def conv_keys():
    """(state_dict prefix, cin, cout, k, stride) for every conv in forward order."""
    out = [("conv1", 2, 32, 3, 1)]
    cin = 32
    for i, planes in enumerate(PLANES):
        for b, stride in enumerate((2, 1)):
            p = f"layer{i}.{b}"
            out.append((f"{p}.conv1", cin, planes, 3, stride))
            out.append((f"{p}.conv2", planes, planes, 3, 1))
            if b == 0:
                out.append((f"{p}.shortcut.0", cin, planes, 1, stride))
            cin = planes
    return out


def bn_for(conv_prefix: str) -> str:
    if conv_prefix.endswith("shortcut.0"):
        return conv_prefix[:-1] + "1"
    return conv_prefix.replace("conv", "bn")


def make_state_dict(seed: int = 10) -> dict:
    """Seeded random parameters with the reference's state_dict keys (numpy legacy RNG:
    bit-stable across numpy versions and machines, unlike torch.manual_seed across builds).

    Conv: kaiming-normal fan_out like arch.py:258-260.  BN: randomised gamma/beta/running stats so
    that BN folding is non-trivial.  FC: uniform(+-1/sqrt(fan_in)) like nn.Linear, rescaled and
    re-centred so that every class of every level occurs on the synthetic CTUs.
    """
    rng = np.random.RandomState(seed)
    sd = {}
    for prefix, cin, cout, k, _ in conv_keys():
        std = np.sqrt(2.0 / (cout * k * k))
        sd[f"{prefix}.weight"] = (rng.standard_normal((cout, cin, k, k)) * std).astype(np.float32)
        if prefix != "conv1":
            bn = bn_for(prefix)
            sd[f"{bn}.weight"] = rng.uniform(0.5, 1.5, cout).astype(np.float32)
            sd[f"{bn}.bias"] = rng.uniform(-0.3, 0.3, cout).astype(np.float32)
            sd[f"{bn}.running_mean"] = rng.uniform(-0.3, 0.3, cout).astype(np.float32)
            sd[f"{bn}.running_var"] = rng.uniform(0.5, 2.0, cout).astype(np.float32)
    # top-level bn1: present in the reference state_dict, never used in forward
    sd["bn1.weight"] = np.ones(32, np.float32)
    sd["bn1.bias"] = np.zeros(32, np.float32)
    sd["bn1.running_mean"] = np.zeros(32, np.float32)
    sd["bn1.running_var"] = np.ones(32, np.float32)
    for i in range(3):
        bound = 1.0 / np.sqrt(FC_IN[i])
        w = rng.uniform(-bound, bound, (FC_OUT[i], FC_IN[i])).astype(np.float32)
        b = rng.uniform(-bound, bound, FC_OUT[i]).astype(np.float32)
        # raw poc (0..32) / qp (22..45) enter un-normalised: keep their columns small so the
        # pooled features, not qp alone, decide the class
        w[:, -2:] *= 0.05
        w[:, :-2] *= 4.0
        # centre the logits (constants measured once on synth_ctus(256, 10)) so every class wins somewhere
        b = (b - np.asarray(HEAD_CENTER[i], np.float32)).astype(np.float32)
        sd[f"branch{i + 1}.weight"] = w
        sd[f"branch{i + 1}.bias"] = b
    return sd



# --------------------------------------------------------------------------- synthetic CTUs


def synth_ctus(n: int, seed: int = 10):
    """Synthetic (org, pred, poc, qp) in the spirit of BASELINE config 2:
    org = 10-bit luma (smooth base + optional edge + texture), pred = clip(org shifted by (dy,dx) +
    N(0,sigma)), poc in [1,31], qp in [22,45]; per-CTU brightness / contrast / noise vary widely so the
    pooled features (and the split classes) vary.  Returns int16 [n,2,128,128] (org, pred), int32 [n,2]."""
    rng = np.random.RandomState(seed)
    out = np.empty((n, 2, 128, 128), np.int16)
    yy, xx = np.mgrid[0:132, 0:132].astype(np.float32)
    for i in range(n):
        f = rng.uniform(0.005, 0.35, 4)
        ph = rng.uniform(0, 6.28, 4)
        amp = rng.uniform(0, 300, 2) * (rng.uniform() < 0.8)
        base = rng.uniform(64, 960) + amp[0] * np.sin(f[0] * xx + ph[0]) * np.cos(f[1] * yy + ph[1])
        base += amp[1] * 0.5 * np.sin(f[2] * (xx + yy) + ph[2])
        if rng.uniform() < 0.35:  # a hard edge through the CTU
            a, b, c = rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(20, 110)
            base += rng.uniform(-400, 400) * ((a * (xx - c) + b * (yy - c)) > 0)
        tex = rng.standard_normal((132, 132)).astype(np.float32) * rng.uniform(0, 60) * (rng.uniform() < 0.85)
        img = np.clip(np.rint(base + tex), 0, 1023)
        dy, dx = rng.randint(0, 3, 2)
        org = img[2:130, 2:130]
        sigma = rng.uniform(0, 30) * (rng.uniform() < 0.9)
        pred = np.clip(np.rint(img[2 - dy : 130 - dy, 2 - dx : 130 - dx] + rng.standard_normal((128, 128)) * sigma), 0, 1023)
        out[i, 0] = org.astype(np.int16)
        out[i, 1] = pred.astype(np.int16)
    rng2 = np.random.RandomState(seed + 7919)  # separate stream: synth_ctus(n)[:k] == synth_ctus(k)
    pq = np.empty((n, 2), np.int32)
    for i in range(n):
        pq[i] = (rng2.randint(1, 32), rng2.randint(22, 46))
    return out, pq




# --------------------------------------------------------------------------- smaller-CU models (64 / 32 / 16 px)

CU_PLANES = (32, 64, 96, 128, 256)  # mlt_cu_or_pq_arch.py:69-82 (layer0..layer4)
CU_FC_IN = (64 + 2, 96 + 2, 128 + 2, 256 + 2)  # branch1..4 hang off layer1..layer4 (:71,74,77,80)
CU_FC_OUT = (2, 3, 4, 6)
CU_SIZES = (64, 32, 16)
# logit centres measured once on synth_cus(256, size, 10) with the un-centred seed-10 parameters
CU_HEAD_CENTER = {
    64: ((0.28, -2.22), (-1.39, 0.82, 5.25), (9.82, 2.10, -7.04, 2.84), (1.42, -10.05, 0.05, 0.99, -2.53, 1.01)),
    32: ((-0.99, -2.40), (1.25, 3.71, 2.90), (-1.54, 0.98, -0.95, -1.62), (-0.89, -0.63, 1.92, 3.01, -1.10, 1.45)),
    16: ((0.75, 0.38), (-1.41, 1.59, 0.80), (-0.18, 1.38, -2.45, 1.02), (-0.65, 0.57, 2.86, 2.69, -1.08, -0.02)),
}


def cu_conv_keys():
    """(state_dict prefix, cin, cout, k, stride) for every conv of MltCnnL4ORPQv4 in forward order
    (BasicBlock [2,2,2,2,2], every stage starts with a stride-2 block: mlt_cu_or_pq_arch.py:69-80,129-130)."""
    out = [("conv1", 2, 32, 3, 1)]
    cin = 32
    for i, planes in enumerate(CU_PLANES):
        for b, stride in enumerate((2, 1)):
            p = f"layer{i}.{b}"
            out.append((f"{p}.conv1", cin, planes, 3, stride))
            out.append((f"{p}.conv2", planes, planes, 3, 1))
            if b == 0:
                out.append((f"{p}.shortcut.0", cin, planes, 1, stride))
            cin = planes
    return out


def make_cu_state_dict(seed: int = 10, size: int = 64, center: bool = True) -> dict:
    """Seeded random parameters with the state_dict keys of `GapBigMltCuORPQ` (one parameter set per CU size, like
    the reference's per-size MLTORPQ_splitMode_<cuw>.pt files, EncCu.cpp:899).  Same recipe as make_state_dict."""
    assert size in CU_SIZES
    rng = np.random.RandomState(seed * 1000 + size)
    sd = {}
    for prefix, cin, cout, k, _ in cu_conv_keys():
        std = np.sqrt(2.0 / (cout * k * k))
        sd[f"{prefix}.weight"] = (rng.standard_normal((cout, cin, k, k)) * std).astype(np.float32)
        if prefix != "conv1":
            bn = bn_for(prefix)
            sd[f"{bn}.weight"] = rng.uniform(0.5, 1.5, cout).astype(np.float32)
            sd[f"{bn}.bias"] = rng.uniform(-0.3, 0.3, cout).astype(np.float32)
            sd[f"{bn}.running_mean"] = rng.uniform(-0.3, 0.3, cout).astype(np.float32)
            sd[f"{bn}.running_var"] = rng.uniform(0.5, 2.0, cout).astype(np.float32)
    sd["bn1.weight"] = np.ones(32, np.float32)  # defined, never used in forward (mlt_cu_or_pq_arch.py:68 vs :104-105)
    sd["bn1.bias"] = np.zeros(32, np.float32)
    sd["bn1.running_mean"] = np.zeros(32, np.float32)
    sd["bn1.running_var"] = np.ones(32, np.float32)
    for i in range(4):
        bound = 1.0 / np.sqrt(CU_FC_IN[i])
        w = rng.uniform(-bound, bound, (CU_FC_OUT[i], CU_FC_IN[i])).astype(np.float32)
        b = rng.uniform(-bound, bound, CU_FC_OUT[i]).astype(np.float32)
        w[:, -2:] *= 0.05
        w[:, :-2] *= 4.0
        if center:
            b = (b - np.asarray(CU_HEAD_CENTER[size][i], np.float32)).astype(np.float32)
        sd[f"branch{i + 1}.weight"] = w
        sd[f"branch{i + 1}.bias"] = b
    return sd


def synth_cus(n: int, size: int, seed: int = 10):
    """Synthetic (org, pred, poc, qp) CUs of size x size luma samples: windows of synth_ctus CTUs at seeded
    positions (what the partitioner hands the hook below the CTU level).  int16 [n,2,size,size], int32 [n,2]."""
    assert size in CU_SIZES
    per = (128 // size) ** 2
    ctus, pq = synth_ctus((n + per - 1) // per, seed + size)
    out = np.empty((n, 2, size, size), np.int16)
    pocqp = np.empty((n, 2), np.int32)
    k = 128 // size
    for i in range(n):
        c, j = divmod(i, per)
        y, x = (j // k) * size, (j % k) * size
        out[i] = ctus[c, :, y : y + size, x : x + size]
        pocqp[i] = pq[c]
    return out, pocqp
