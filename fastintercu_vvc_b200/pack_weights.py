"""Offline weight packer: reference state_dict -> MLTW blob loaded once by mlt_create().

Replaces the TorchScript container the reference hook re-loads per CTU (EncCu.cpp:894-900; exported by
mlt-cnn-python/codes/model2torchScript.py:22-48).  Input is the same state_dict that script reads
(`torch.load(path)['params']`, optional `module.` prefixes, :23-32) or an in-memory dict of numpy arrays.

What it does
  * folds every eval-mode BatchNorm (eps 1e-5) into the preceding bias-free conv:
        w' = w * gamma / sqrt(var + eps),  b' = beta - mean * gamma / sqrt(var + eps)
    (the top-level `bn1.*` keys are ignored: defined but unused, arch.py:247 vs :277-278);
  * writes the folded 3x3 / 1x1 conv weights as fp16 in the tcgen05 shared-memory operand layout used by
    csrc/conv_umma.cuh: K-major, no-swizzle core matrices, [cin_group][tap][G/8][Cout][8] so that every
    (cin_group, tap) slab is one contiguous bulk copy;
  * writes, for the second conv of every BasicBlock, the extra K-slab operand [xc/gx][gx/8][Cout][8]: the folded
    1x1 stride-2 shortcut weights, or the identity matrix for an identity residual;
  * writes conv1 as a tcgen05 operand with the staging scale folded in (conv1_operand);
  * writes each conv's (shortcut-fused) bias as a K=16 fp16 operand (hi/lo split) that one extra MMA adds
    into the accumulator;
  * writes fp32 copies ([tap][Cin][Cout]) for the fp32 CUDA-core cross-check engine, the fp32 biases, the
    fp32 conv1 ([tap][2][32], consumed by the fused staging+conv1 kernel) and the three FC heads.

The smaller-CU models (64 / 32 / 16-px `GapBigMltCuORPQ`, mlt_cu_or_pq_arch.py:59-130; exported per size by
model2torchScript.py:14-22) use the same container with the header's arch field = CU size: the operand layouts follow
the per-size tile shapes of csrc/cu_net.cuh, so a blob is packed for ONE size (`cu_conv_table(size)`).

  * corrects every tcgen05 conv's bias for the SYSTEMATIC part of its fp16 weight-rounding error, E[(w_q - w) * x], with per-tap
    input means from one exact fp32 forward over calibration blocks (bias_correction; seeded synthetic blocks by default,
    `--calib blocks.npy` for real ones); the fp32 cross-check sections are left exact.

CLI:  python -m fastintercu_vvc_b200.pack_weights model.pth out.mltw            (128x128 CTU model)
      python -m fastintercu_vvc_b200.pack_weights --cu 64 model.pth out.mltw    (64 / 32 / 16-px CU model)
      python -m fastintercu_vvc_b200.pack_weights --calib blocks.npy model.pth out.mltw
"""
from __future__ import annotations

import struct
import sys

import numpy as np

MAGIC = 0x57544C4D  # "MLTW"
VERSION = 2
ARCH_CTU128 = 128
BN_EPS = 1e-5
PLANES = (32, 64, 128, 256)

SEC_CONV1_F32 = 0x001
SEC_CONV1_UMMA = 0x002
SEC_STEM_CONV1 = 0x003
SEC_STEM5_W = 0x004     # composite stem operands for csrc/stem5_umma.cu (stem5_operands)
SEC_STEM5_CORR = 0x005  # fp32 border-correction weights of the same kernel
SEC_W_F16 = 0x100
SEC_BIAS_FUSED = 0x200
SEC_W_F32 = 0x300
SEC_BIAS = 0x400
SEC_SC_W_F32 = 0x600
SEC_SC_BIAS = 0x700
SEC_BIAS_MMA = 0xA00
SEC_FC_W = 0x800
SEC_FC_B = 0x900
SEC_X_W_F16 = 0xB00
SEC_W_SPLIT = 0xC00  # layer3 weights again, output channels split 4 ways: [split][cin_group][tap][G/8][Cout/4][8]
SEC_X_SPLIT = 0xD00  # and their extra-operand slabs: [split][xc/gx][(hi, lo)][gx/8][Cout/4][8]
SEC_W_SPLIT8 = 0xE00  # layer3 split 8 ways (tiny batches)
SEC_X_SPLIT8 = 0xF00
SPLIT_LAYERS, SPLIT_WAYS = (8, 9, 10, 11, 12, 13, 14, 15), 4  # layer2 and layer3, = CONV_SPLIT_* / CONV_L2SPLIT_* of csrc/mlt_internal.h
SPLIT8_LAYERS, SPLIT8_WAYS = (12, 13, 14, 15), 8
ROUNDING = "diffused"  # "nearest" = independent round-to-nearest (tools/precision_ab.py compares the two)
ALPHA = np.float32(1.0 / 1023)  # (float)(1.0/1023), bits 0x3A802008: cv::Mat::convertTo's alpha at EncCu.cpp:835-838


def conv_table():
    """3x3 convs after conv1, forward order: (prefix, cin, cout, stride, hout, cin_group, shortcut_idx|-1)."""
    rows = []
    cin, hin = 32, 128
    for L, planes in enumerate(PLANES):
        hout = hin // 2
        p = f"layer{L}"
        g = 32 if planes >= 256 else min(planes, 64)  # = ConvCfg::G of csrc/conv_umma.cuh
        rows.append((f"{p}.0.conv1", cin, planes, 2, hout, 16 if planes >= 256 else 32, -1))
        rows.append((f"{p}.0.conv2", planes, planes, 1, hout, g, L))
        rows.append((f"{p}.1.conv1", planes, planes, 1, hout, g, -1))
        rows.append((f"{p}.1.conv2", planes, planes, 1, hout, g, -1))
        cin, hin = planes, hout
    return rows


def _np(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v)


def normalise_state_dict(sd: dict) -> dict:
    """Accepts the checkpoint dict, its 'params' entry, torch tensors or numpy arrays; strips 'module.'."""
    if "params" in sd and not any(k.endswith(".weight") for k in sd):
        sd = sd["params"]
    out = {}
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[7:]
        out[k] = _np(v)
    return out


def fold_bn(w: np.ndarray, sd: dict, bn: str):
    g = sd[f"{bn}.weight"].astype(np.float64)
    b = sd[f"{bn}.bias"].astype(np.float64)
    m = sd[f"{bn}.running_mean"].astype(np.float64)
    v = sd[f"{bn}.running_var"].astype(np.float64)
    s = g / np.sqrt(v + BN_EPS)
    return (w.astype(np.float64) * s[:, None, None, None]).astype(np.float32), (b - m * s).astype(np.float32)


def quantize_fp16_diffused(w: np.ndarray) -> np.ndarray:
    """Folded OIHW fp32 -> fp16 with ERROR DIFFUSION instead of independent round-to-nearest.

    fp16 weight rounding is the dominant error of this fp16-operand / fp32-accumulate network (measured: mean |dlogit|
    8e-4 from weights alone vs 1e-4 from activation rounding), and it is systematic: the same rounding errors hit every
    pixel, so they do not average out in the global pools.  Each weight is therefore rounded to the fp16 value nearest to
    (weight + error carried from the previous one): along the 9 taps of every (cout, cin) pair for 3x3 kernels -- the
    input of one channel is locally smooth, so sum_taps(dw) ~ 0 cancels the error to first order -- and along cin for the
    1x1 shortcuts.  Every stored value is still within one fp16 ulp of the exact weight; it halves the logit error at no
    run-time cost."""
    cout, cin, kh, kw = w.shape
    w64 = w.astype(np.float64)
    if ROUNDING == "nearest":  # A/B switch for tools/precision_ab.py
        return w.astype(np.float16)
    if kh * kw > 1:
        flat = w64.reshape(cout, cin, kh * kw)
        out = np.empty(flat.shape, np.float16)
        carry = np.zeros((cout, cin))
        for t in range(kh * kw):
            v = flat[:, :, t] + carry
            out[:, :, t] = v.astype(np.float16)
            carry = v - out[:, :, t].astype(np.float64)
    else:
        flat = w64.reshape(cout, cin)
        out = np.empty(flat.shape, np.float16)
        carry = np.zeros(cout)
        for c in range(cin):
            v = flat[:, c] + carry
            out[:, c] = v.astype(np.float16)
            carry = v - out[:, c].astype(np.float64)
    return out.reshape(w.shape)


# --------------------------------------------------------------------------- calibrated bias correction


def stage_input(orgpred: np.ndarray) -> np.ndarray:
    """int16 [n][2][S][S] (org, pred) -> float32 [n][2][S][S] (org, |org - pred|) * (float)(1/1023), clamped: EncCu.cpp:810-867."""
    o = orgpred[:, 0].astype(np.uint16).astype(np.int32)
    q = orgpred[:, 1].astype(np.uint16).astype(np.int32)
    x = np.stack([o, np.abs(o - q)], 1).astype(np.float32) * ALPHA
    return np.clip(x, 0.0, 1.0)


def tap_means(x, k: int, stride: int, pad: int) -> np.ndarray:
    """Mean, over images and output positions, of the zero-padded input sample each kernel tap sees: [C][k][k] (torch [n][C][H][W])."""
    import torch.nn.functional as F

    n, c, h, w = x.shape
    xp = F.pad(x, (pad, pad, pad, pad))
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out = np.zeros((c, k, k))
    for a in range(k):
        for b in range(k):
            out[:, a, b] = xp[:, :, a : a + stride * (ho - 1) + 1 : stride, b : b + stride * (wo - 1) + 1 : stride].double().mean((0, 2, 3)).numpy()
    return out


def calibrate_tap_means(sd: dict, orgpred: np.ndarray) -> dict:
    """One exact fp32 forward of the network (torch CPU, folded BN) over calibration blocks; returns, per 3x3 conv prefix, the mean
    input sample under each kernel tap [cin][3][3] (zero padding included), plus 'stem5': the same for the composite 5x5 stride-2
    stem conv over the two input planes [2][5][5].  Works for the CTU network (4 stages) and the CU networks (5 stages)."""
    import torch
    import torch.nn.functional as F

    sd = normalise_state_dict(sd)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32))
    means = {}
    with torch.no_grad():
        x = T(stage_input(orgpred))
        means["stem5"] = tap_means(x, 5, 2, 2)
        y = F.conv2d(x, T(sd["conv1.weight"]), padding=1)
        L = 0
        while f"layer{L}.0.conv1.weight" in sd:
            for b in range(2):
                p = f"layer{L}.{b}"
                w1, b1 = fold_bn(sd[f"{p}.conv1.weight"], sd, f"{p}.bn1")
                w2, b2 = fold_bn(sd[f"{p}.conv2.weight"], sd, f"{p}.bn2")
                stride = 2 if b == 0 else 1
                means[f"{p}.conv1"] = tap_means(y, 3, stride, 1)
                t = F.relu(F.conv2d(y, T(w1), T(b1), stride=stride, padding=1))
                means[f"{p}.conv2"] = tap_means(t, 3, 1, 1)
                o = F.conv2d(t, T(w2), T(b2), padding=1)
                if b == 0:
                    ws, bs = fold_bn(sd[f"{p}.shortcut.0.weight"], sd, f"{p}.shortcut.1")
                    o = o + F.conv2d(y, T(ws), T(bs), stride=stride)
                else:
                    o = o + y
                y = F.relu(o)
            L += 1
    # fp16-rounded means: the correction must not depend on the last bits of this machine's fp32 conv kernels
    return {k: v.astype(np.float16).astype(np.float64) for k, v in means.items()}


def bias_correction(exact: np.ndarray, quantized: np.ndarray, tapmean: np.ndarray) -> np.ndarray:
    """E[sum_k (w_q - w) x_k] per output channel for weights [cout][cin][k][k] and tap means [cin][k][k]: the systematic part of
    the fp16 weight-rounding error.  It is the same at every pixel of every block, so it survives the global average pools and is
    the largest single error of the fp16-operand path (tools/emulate_ctu_precision.py: per-logit offsets up to 6e-4 against a
    sample-to-sample spread of 1e-4); subtracting it from the layer's bias costs nothing at run time."""
    dw = quantized.astype(np.float64) - exact.astype(np.float64)
    return np.einsum("ncab,cab->n", dw, tapmean)


CALIB_SEED, CALIB_CTUS, CALIB_CUS = 77001, 16, 96


def default_calibration(arch: int) -> np.ndarray:
    """Calibration blocks when the caller has none (e.g. dumped from a real encode, hook/dataset_dump): seeded synthetic ones."""
    from . import synth

    if arch == ARCH_CTU128:
        return synth.synth_ctus(CALIB_CTUS, CALIB_SEED)[0]
    return synth.synth_cus(CALIB_CUS, arch, CALIB_SEED)[0]


def pack_umma_b(w: np.ndarray, group: int) -> np.ndarray:
    """Folded OIHW fp32 -> fp16 [cin/group][kh*kw][group/8][cout][8] (error-diffused rounding, quantize_fp16_diffused)."""
    cout, cin, kh, kw = w.shape
    assert cin % group == 0 and group % 16 == 0
    t = quantize_fp16_diffused(w).reshape(cout, cin // group, group // 8, 8, kh * kw)  # [n][cg][j][e][t]
    t = t.transpose(1, 4, 2, 0, 3)  # [cg][t][j][n][e]
    return np.ascontiguousarray(t)


def split_cout(packed: np.ndarray, ways: int) -> np.ndarray:
    """[..., cout, 8] operand slabs -> [ways][..., cout / ways, 8]: the layout of the channel-split kernels (ConvCfg::NSPLIT),
    where every CTA streams only its own quarter of each slab as one contiguous bulk copy."""
    cout = packed.shape[-2]
    t = packed.reshape(*packed.shape[:-2], ways, cout // ways, 8)
    return np.ascontiguousarray(np.moveaxis(t, -3, 0))


def bias_operand(b: np.ndarray) -> np.ndarray:
    """fp32 bias [cout] -> fp16 tcgen05 B operand [2][cout][8]: k=0 hi(b), k=1 lo(b) = b - hi(b), rest 0.
    Multiplied by an A operand whose k=0,1 columns are 1.0 it initialises the accumulator to b (to ~2^-22)."""
    hi = b.astype(np.float16)
    lo = (b.astype(np.float32) - hi.astype(np.float32)).astype(np.float16)
    out = np.zeros((2, b.shape[0], 8), np.float16)
    out[0, :, 0] = hi
    out[0, :, 1] = lo
    return out


def conv1_operand(w: np.ndarray) -> np.ndarray:
    """conv1 OIHW fp32 [32][2][3][3] -> fp16 tcgen05 B operand [hi, lo][4 chunks][32 cout][8] (csrc/stage_conv1.cu).

    chunk kh (0..2), element dx*2 + ch = w[cout][ch][kh][dx] * alpha * 2^10 for dx < 3, else 0; chunk 3 = 0.
    The A operand is the integer sample * 2^-10 (exact in fp16), so A x (hi + lo) reproduces
    w * (v * (float)(1/1023)) to ~2^-22 relative -- the staging multiply is folded into the weights."""
    ws = w.astype(np.float64) * float(ALPHA) * 1024.0
    op = np.zeros((4, 32, 8), np.float64)
    for kh in range(3):
        for dx in range(3):
            for ch in range(2):
                op[kh, :, dx * 2 + ch] = ws[:, ch, kh, dx]
    hi = op.astype(np.float16)
    lo = (op - hi.astype(np.float64)).astype(np.float16)
    return np.stack([hi, lo], 0)


def stem_conv1_operand(w: np.ndarray) -> np.ndarray:
    """conv1 for the fused stem kernel (csrc/stem_umma.cu): fp16 [py][2 MMAs][2 chunks][32 cout][8].

    Chunk kh = one kernel row: element dx*2 + ch = w[cout][ch][kh][dx] * (float)(1/1023) * 2^10 for dx < 3, else 0
    (the A operand is the integer sample * 2^-10, exact in fp16, so the staging multiply is folded into the weights);
    fp16 with error diffusion over the nine taps of every (cout, ch) kernel (quantize_fp16_diffused).  Per output-row
    parity py the stem issues two K=16 MMAs whose two chunks are kernel rows at a constant shared-memory distance:
      py = 0: [kh1 | kh0], [0 | kh2]        py = 1: [kh0 | kh1], [0 | kh2]"""
    ws = (w.astype(np.float64) * float(ALPHA) * 1024.0).astype(np.float32)
    q = quantize_fp16_diffused(ws)  # [32][2][3][3] fp16
    chunk = np.zeros((4, 32, 8), np.float16)  # chunk 3 = zeros
    for kh in range(3):
        for dx in range(3):
            for ch in range(2):
                chunk[kh, :, dx * 2 + ch] = q[:, ch, kh, dx]
    z = 3
    order = (((1, 0), (z, 2)), ((0, 1), (z, 2)))
    out = np.zeros((2, 2, 2, 32, 8), np.float16)
    for py in range(2):
        for m in range(2):
            for k in range(2):
                out[py, m, k] = chunk[order[py][m][k]]
    return out


def stem5_composite(w1: np.ndarray, w0f: np.ndarray):
    """conv1 (arch.py:278: 2 -> 32, 3x3, pad 1, NO BatchNorm, NO activation) feeds layer0.0.conv1 (32 -> 32, 3x3, stride 2, pad 1,
    folded BN) and layer0.0's shortcut directly, so conv1 o layer0.0.conv1 is ONE linear map of the input: a 5x5 stride-2 conv
    2 -> 32 -- except where layer0.0.conv1's zero padding replaces conv1 outputs at row / column -1 (output row 0 / column 0).
    Returns float64
      W5   [32][2][5][5]  W5[co][ci][a+kh][b+kw] += w0f[co][cm][a][b] * w1[cm][ci][kh][kw]; z = W5 * in(2oy-2.., 2ox-2..) (zero-padded in)
      Wtop [32][2][5]     the part of W5[dy = 2] that came through conv1 row -1 (a = 0, kh = 2): subtract Wtop * in(0, 2ox-2+e) at oy = 0
      Wleft[32][2][5]     the same through conv1 column -1 (b = 0, kw = 2): subtract Wleft * in(2oy-2+e, 0) at ox = 0
      Wc   [32][2]        conv1(-1,-1)'s share (a = b = 0, kh = kw = 2), contained in both: add back Wc * in(0, 0) at (0, 0)."""
    w1 = w1.astype(np.float64)
    w0f = w0f.astype(np.float64)
    W5 = np.zeros((32, 2, 5, 5))
    Wtop = np.zeros((32, 2, 5))
    Wleft = np.zeros((32, 2, 5))
    for a in range(3):
        for b in range(3):
            for kh in range(3):
                for kw in range(3):
                    t = w0f[:, :, a, b] @ w1[:, :, kh, kw]  # [co][ci]
                    W5[:, :, a + kh, b + kw] += t
                    if a == 0 and kh == 2:
                        Wtop[:, :, b + kw] += t
                    if b == 0 and kw == 2:
                        Wleft[:, :, a + kh] += t
    Wc = w0f[:, :, 0, 0] @ w1[:, :, 2, 2]
    return W5, Wtop, Wleft, Wc


def stem5_operands(w1: np.ndarray, w0f: np.ndarray):
    """Operands of csrc/stem5_umma.cu.  The A operand rows are output pixels; a 16-byte K chunk is {org, res} of four horizontally
    adjacent input samples (sample * 2^-10, exact in fp16), so all weights carry (float)(1/1023) * 2^10 (the staging multiply of
    EncCu.cpp:835-838 folded in, like conv1_operand).
      fp16 [7][2 chunks][32 cout][8]:
        m = 0..4  composite 5x5 stride-2 conv, kernel row dy = m: chunk 0 element dx*2 + ch = W5[co][ch][dy][dx] (dx < 4),
                  chunk 1 elements 0, 1 = W5[co][ch][dy][4], rest 0
        m = 5     conv1 at even positions (the shortcut's input): chunks = kernel rows kh = 0 and kh = 2, element dx*2 + ch
        m = 6     chunk 0 = kernel row kh = 1, chunk 1 = 0
      fp32 [2][5][2][32] + [2][32]: Wtop / Wleft as [e][ch][co] and Wc as [ch][co], scaled the same way (CUDA-core border fix)."""
    scale = float(ALPHA) * 1024.0
    W5, Wtop, Wleft, Wc = stem5_composite(w1, w0f)
    op = np.zeros((7, 2, 32, 8), np.float64)
    for dy in range(5):
        for dx in range(4):
            for ch in range(2):
                op[dy, 0, :, dx * 2 + ch] = W5[:, ch, dy, dx] * scale
        for ch in range(2):
            op[dy, 1, :, ch] = W5[:, ch, dy, 4] * scale
    w1d = w1.astype(np.float64) * scale
    for m, c, kh in ((5, 0, 0), (5, 1, 2), (6, 0, 1)):
        for dx in range(3):
            for ch in range(2):
                op[m, c, :, dx * 2 + ch] = w1d[:, ch, kh, dx]
    corr = np.concatenate([(Wtop * scale).transpose(2, 1, 0).reshape(-1), (Wleft * scale).transpose(2, 1, 0).reshape(-1),
                           (Wc * scale).transpose(1, 0).reshape(-1)]).astype(np.float32)
    return op.astype(np.float16), corr


def stem5_sections(w1: np.ndarray, w0f: np.ndarray, b0f: np.ndarray, tapmean5: np.ndarray | None):
    """(SEC_STEM5_W array, SEC_STEM5_CORR array): the operands of stem5_operands + the stem kernel's own bias [32] at the end of
    the fp32 section, corrected for the fp16 rounding of the composite weights when tap means are given."""
    s5w, s5c = stem5_operands(w1, w0f)
    b5 = b0f.astype(np.float64)
    if tapmean5 is not None:  # the composite weights as the kernel sees them (fp16, without the staging scale) against the exact ones
        W5 = stem5_composite(w1, w0f)[0]
        W5q = np.zeros_like(W5)
        sc = float(ALPHA) * 1024.0
        for dy in range(5):
            for ch in range(2):
                for dx in range(4):
                    W5q[:, ch, dy, dx] = s5w[dy, 0, :, dx * 2 + ch].astype(np.float64) / sc
                W5q[:, ch, dy, 4] = s5w[dy, 1, :, ch].astype(np.float64) / sc
        b5 = b5 - bias_correction(W5, W5q, tapmean5)
    return s5w, np.concatenate([s5c, b5.astype(np.float32)])


def extra_operand_hilo(ws: np.ndarray, gx: int) -> np.ndarray:
    """[cout][xc] fp32 folded 1x1 shortcut weights -> fp16 [xc/gx][hi, lo][gx/8][cout][8]: hi = fp16(w), lo = fp16(w - hi).
    The conv kernel runs the extra-operand stage twice (ConvCfg::XLO), so the shortcut is applied at ~2^-22 precision."""
    cout, xc = ws.shape
    assert xc % gx == 0 and gx % 16 == 0
    hi = ws.astype(np.float16)
    lo = (ws.astype(np.float64) - hi.astype(np.float64)).astype(np.float16)
    t = np.stack([hi, lo], 0).reshape(2, cout, xc // gx, gx // 8, 8).transpose(2, 0, 3, 1, 4)
    return np.ascontiguousarray(t)


def extra_operand(ws: np.ndarray, gx: int) -> np.ndarray:
    """[cout][xc] fp32 (folded 1x1 shortcut weights, or the identity) -> fp16 [xc/gx][gx/8][cout][8]."""
    cout, xc = ws.shape
    assert xc % gx == 0 and gx % 16 == 0
    q = quantize_fp16_diffused(ws.reshape(cout, xc, 1, 1)).reshape(cout, xc)  # exact for the identity
    t = q.reshape(cout, xc // gx, gx // 8, 8).transpose(1, 2, 0, 3)
    return np.ascontiguousarray(t)


def build_sections(sd: dict, calib: np.ndarray | None = None, correct_bias: bool = True) -> list:
    """calib: int16 [n][2][128][128] calibration CTUs for the bias correction (default: seeded synthetic ones)."""
    sd = normalise_state_dict(sd)
    secs = []
    tm = calibrate_tap_means(sd, default_calibration(ARCH_CTU128) if calib is None else calib) if correct_bias else None

    def add(sid, arr, dtype):
        secs.append((sid, np.ascontiguousarray(arr, dtype)))

    w = sd["conv1.weight"].astype(np.float32)  # [32][2][3][3], no BN / bias
    add(SEC_CONV1_F32, w.transpose(2, 3, 1, 0).reshape(9, 2, 32), np.float32)
    add(SEC_CONV1_UMMA, conv1_operand(w), np.float16)
    add(SEC_STEM_CONV1, stem_conv1_operand(w), np.float16)
    w0f, b0f = fold_bn(sd["layer0.0.conv1.weight"], sd, "layer0.0.bn1")
    s5w, s5c = stem5_sections(w, w0f, b0f, tm["stem5"] if tm is not None else None)
    add(SEC_STEM5_W, s5w, np.float16)
    add(SEC_STEM5_CORR, s5c, np.float32)  # border-term weights + the stem kernel's own (corrected) bias [32]
    for li, (prefix, cin, cout, stride, hout, group, sc) in enumerate(conv_table()):
        bn = prefix.replace("conv", "bn")
        wf, bf = fold_bn(sd[f"{prefix}.weight"], sd, bn)
        assert wf.shape == (cout, cin, 3, 3), (prefix, wf.shape)
        packed = pack_umma_b(wf, group)
        add(SEC_W_F16 + li, packed, np.float16)
        corr = bias_correction(wf, quantize_fp16_diffused(wf), tm[prefix]) if tm is not None else 0.0
        if li in SPLIT_LAYERS:
            add(SEC_W_SPLIT + li, split_cout(packed, SPLIT_WAYS), np.float16)
        if li in SPLIT8_LAYERS:
            add(SEC_W_SPLIT8 + li, split_cout(packed, SPLIT8_WAYS), np.float16)
        add(SEC_W_F32 + li, wf.transpose(2, 3, 1, 0).reshape(9, cin, cout), np.float32)
        add(SEC_BIAS + li, bf, np.float32)
        fused = bf
        if sc >= 0:
            sp = prefix.rsplit(".", 1)[0] + ".shortcut"
            ws, bs = fold_bn(sd[f"{sp}.0.weight"], sd, f"{sp}.1")
            csc = ws.shape[1]
            xop = extra_operand_hilo(ws.reshape(cout, csc), min(csc, group))
            add(SEC_X_W_F16 + li, xop, np.float16)
            if li in SPLIT_LAYERS:
                add(SEC_X_SPLIT + li, split_cout(xop, SPLIT_WAYS), np.float16)
            if li in SPLIT8_LAYERS:
                add(SEC_X_SPLIT8 + li, split_cout(xop, SPLIT8_WAYS), np.float16)
            add(SEC_SC_W_F32 + sc, ws.reshape(cout, csc).T, np.float32)
            add(SEC_SC_BIAS + sc, bs, np.float32)
            fused = (bf.astype(np.float32) + bs.astype(np.float32)).astype(np.float32)
        elif li & 1:
            # identity residual of the second block (arch.py:44-57): one more K-slab with Wx = I, exact in fp16
            xop = extra_operand(np.eye(cout, dtype=np.float32), min(cout, group))
            add(SEC_X_W_F16 + li, xop, np.float16)
            if li in SPLIT_LAYERS:
                add(SEC_X_SPLIT + li, split_cout(xop, SPLIT_WAYS), np.float16)
            if li in SPLIT8_LAYERS:
                add(SEC_X_SPLIT8 + li, split_cout(xop, SPLIT8_WAYS), np.float16)
        fused = (fused.astype(np.float64) - corr).astype(np.float32)  # tcgen05 path only: SEC_BIAS (fp32 engine, exact weights) stays exact
        add(SEC_BIAS_FUSED + li, fused, np.float32)
        add(SEC_BIAS_MMA + li, bias_operand(fused), np.float16)
    for i in range(3):
        add(SEC_FC_W + i, sd[f"branch{i + 1}.weight"], np.float32)
        add(SEC_FC_B + i, sd[f"branch{i + 1}.bias"], np.float32)
    return secs


def cu_conv_table(size: int):
    """3x3 convs of the `size`-px CU network after conv1, forward order -- mirrors csrc/cu_net.cuh CuSel / ConvCfg:
    (prefix, cin, cout, stride as executed, hout, cin_group G, xc, gx, kind) with kind 0..3 = position in the stage."""
    planes_all = (32, 64, 96, 128, 256)
    rows = []
    cin_stage, hin = 32, size
    for L, planes in enumerate(planes_all):
        fake_s2 = hin == 1  # stride-2 conv on a 1x1 map == the stride-1 conv (reads the same samples)
        hout = max(hin // 2, 1)
        flat = hout <= 4
        for k in range(4):
            cin = cin_stage if k == 0 else planes
            stride = 2 if (k == 0 and not fake_s2) else 1
            if stride == 2 and (planes >= 256 or flat):
                g = 16
            elif stride == 2 or planes >= 256 or flat:
                g = 32
            else:
                g = 32 if cin % 64 else 64
            xc = cin_stage if k == 1 else (planes if k == 3 else 0)
            gx = 0 if xc == 0 else (xc if xc < g else (g if xc % g == 0 else 32))
            prefix = f"layer{L}.{k // 2}.conv{k % 2 + 1}"
            rows.append((prefix, cin, planes, stride, hout, g, xc, gx, k))
        cin_stage, hin = planes, hout
    return rows


def build_cu_sections(sd: dict, size: int, calib: np.ndarray | None = None, correct_bias: bool = True) -> list:
    sd = normalise_state_dict(sd)
    secs = []
    tm = calibrate_tap_means(sd, default_calibration(size) if calib is None else calib) if correct_bias else None

    def add(sid, arr, dtype):
        secs.append((sid, np.ascontiguousarray(arr, dtype)))

    w = sd["conv1.weight"].astype(np.float32)  # [32][2][3][3], no BN / bias (mlt_cu_or_pq_arch.py:105)
    add(SEC_CONV1_F32, w.transpose(2, 3, 1, 0).reshape(9, 2, 32), np.float32)
    add(SEC_CONV1_UMMA, conv1_operand(w), np.float16)
    add(SEC_STEM_CONV1, stem_conv1_operand(w), np.float16)  # round-1 fused stem (csrc/stem_umma.cu; MLT_STEM_OLD=1)
    w0f, b0f = fold_bn(sd["layer0.0.conv1.weight"], sd, "layer0.0.bn1")
    s5w, s5c = stem5_sections(w, w0f, b0f, tm["stem5"] if tm is not None else None)  # the 64- / 32-px networks run csrc/stem5_umma.cu
    add(SEC_STEM5_W, s5w, np.float16)
    add(SEC_STEM5_CORR, s5c, np.float32)
    for li, (prefix, cin, cout, stride, hout, group, xc, gx, kind) in enumerate(cu_conv_table(size)):
        wf, bf = fold_bn(sd[f"{prefix}.weight"], sd, prefix.replace("conv", "bn"))
        assert wf.shape == (cout, cin, 3, 3), (prefix, wf.shape)
        add(SEC_W_F16 + li, pack_umma_b(wf, group), np.float16)
        add(SEC_W_F32 + li, wf.transpose(2, 3, 1, 0).reshape(9, cin, cout), np.float32)
        add(SEC_BIAS + li, bf, np.float32)
        # (tap means include the zero padding, so the centre-tap-only layers of 1x1 maps are covered: their other taps have mean 0)
        corr = bias_correction(wf, quantize_fp16_diffused(wf), tm[prefix]) if tm is not None else 0.0
        fused = bf
        if kind == 1:  # 1x1 stride-2 shortcut conv + BN of the stage's first block, as hi + lo fp16 K-slabs
            sp = prefix.rsplit(".", 1)[0] + ".shortcut"
            ws, bs = fold_bn(sd[f"{sp}.0.weight"], sd, f"{sp}.1")
            assert ws.shape[1] == xc
            add(SEC_X_W_F16 + li, extra_operand_hilo(ws.reshape(cout, xc), gx), np.float16)
            add(SEC_SC_W_F32 + li // 4, ws.reshape(cout, xc).T, np.float32)
            add(SEC_SC_BIAS + li // 4, bs, np.float32)
            fused = (bf.astype(np.float32) + bs.astype(np.float32)).astype(np.float32)
        elif kind == 3:  # identity residual of the second block
            add(SEC_X_W_F16 + li, extra_operand(np.eye(cout, dtype=np.float32), gx), np.float16)
        fused = (fused.astype(np.float64) - corr).astype(np.float32)
        add(SEC_BIAS_FUSED + li, fused, np.float32)
        add(SEC_BIAS_MMA + li, bias_operand(fused), np.float16)
    for i in range(4):
        add(SEC_FC_W + i, sd[f"branch{i + 1}.weight"], np.float32)
        add(SEC_FC_B + i, sd[f"branch{i + 1}.bias"], np.float32)
    return secs


def pack(sd: dict, arch: int = ARCH_CTU128, calib: np.ndarray | None = None) -> bytes:
    secs = build_sections(sd, calib) if arch == ARCH_CTU128 else build_cu_sections(sd, arch, calib)
    table_bytes = 24 * len(secs)
    off = (32 + table_bytes + 255) // 256 * 256
    table, blobs = [], []
    for sid, arr in secs:
        raw = arr.tobytes()
        table.append(struct.pack("<IIQQ", sid, 1 if arr.dtype == np.float16 else 0, off, len(raw)))
        pad = (-len(raw)) % 256
        blobs.append(raw + b"\0" * pad)
        off += len(raw) + pad
    head = struct.pack("<IIIIQQ", MAGIC, VERSION, arch, len(secs), off, 0)
    body = head + b"".join(table)
    body += b"\0" * ((-len(body)) % 256)
    out = body + b"".join(blobs)
    assert len(out) == off
    return out


def write_blob(sd: dict, path: str, calib: np.ndarray | None = None) -> int:
    data = pack(sd, calib=calib)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def write_cu_blob(sd: dict, size: int, path: str, calib: np.ndarray | None = None) -> int:
    assert size in (64, 32, 16)
    data = pack(sd, size, calib)
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def read_sections(path: str, arch: int = ARCH_CTU128) -> dict:
    """Parse an MLTW blob back into {section id: numpy array} (round-trip checks / tooling)."""
    raw = open(path, "rb").read()
    magic, ver, barch, nsec, total, _ = struct.unpack_from("<IIIIQQ", raw, 0)
    if magic != MAGIC or ver != VERSION or barch != arch or total != len(raw):
        raise ValueError(f"not an MLTW v2 blob for arch {arch}")
    out = {}
    for i in range(nsec):
        sid, dt, off, nb = struct.unpack_from("<IIQQ", raw, 32 + 24 * i)
        out[sid] = np.frombuffer(raw, np.float16 if dt == 1 else np.float32, nb // (2 if dt == 1 else 4), off)
    return out


def load_checkpoint(path: str) -> dict:
    """Reads what the reference produces: a BasicSR checkpoint `.pth` whose state_dict sits
    under 'params' with optional 'module.' prefixes (model2torchScript.py:23-32), a bare state_dict, or the traced
    TorchScript `.pt` the hook loads (MLTORPQ_splitMode_<cuw>.pt, model2torchScript.py:46-48; EncCu.cpp:899)."""
    import zipfile

    import torch

    is_script = False
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            is_script = any(n.endswith("constants.pkl") or "/code/" in n for n in z.namelist())
    if is_script:
        sd = torch.jit.load(path, map_location="cpu").state_dict()
    else:
        sd = torch.load(path, map_location="cpu")
    sd = normalise_state_dict(sd)
    return {k: v for k, v in sd.items() if not k.endswith("num_batches_tracked")}


def export_state_dict(sd: dict, path: str) -> None:
    """Writes a checkpoint in the reference's own container ({'params': state_dict}, mlt_base_model.py save format) so that
    blobs can be round-tripped and seeded parameters handed to the reference's scripts."""
    import torch

    torch.save({"params": {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in normalise_state_dict(sd).items()}}, path)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    size, calib = 0, None
    while len(argv) >= 2 and argv[0] in ("--cu", "--calib"):
        if argv[0] == "--cu":
            size = int(argv[1])
        else:  # int16 [n][2][S][S] .npy of (org, pred) blocks, e.g. collected with hook/dataset_dump from a real encode
            calib = np.load(argv[1])
        argv = argv[2:]
    if len(argv) != 2 or size not in (0, 64, 32, 16):
        print(__doc__)
        return 2
    sd = load_checkpoint(argv[0])
    n = write_cu_blob(sd, size, argv[1], calib) if size else write_blob(sd, argv[1], calib)
    print(f"wrote {argv[1]}: {n} bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
