"""GPU parity tests (run on the B200 box: python -m pytest tests -m gpu).

Everything goes through the C ABI of libmltcnn.so (fastintercu_vvc_b200.capi).  The checker is the CPU
oracle (oracle/mltcnn_oracle.c via tests/oracle_lib.py) and the committed golden vectors produced by the
reference's own architecture file (tests/golden/, tools/gen_golden.py).

Bars (BASELINE.json north_star): integer staging bit-exact; probabilities max |diff| <= 1e-3 vs fp32;
>= 99.9 % split-decision agreement.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from oracle import ref_arch
from tests.oracle_lib import OracleModel

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PROB_TOL = 1e-3  # north_star: max abs diff of probabilities vs the fp32 reference


def softmax_levels(lg):
    out = np.empty_like(lg)
    for a, b in ((0, 2), (2, 5), (5, 9)):
        e = np.exp(lg[:, a:b] - lg[:, a:b].max(1, keepdims=True))
        out[:, a:b] = e / e.sum(1, keepdims=True)
    return out


@pytest.fixture(scope="module")
def sd():
    return ref_arch.make_state_dict(10)


@pytest.fixture(scope="module")
def blob(sd):
    from fastintercu_vvc_b200.pack_weights import write_blob

    with tempfile.NamedTemporaryFile(suffix=".mltw", delete=False) as f:
        path = f.name
    write_blob(sd, path)
    yield path
    os.unlink(path)


@pytest.fixture(scope="module")
def pred(blob):
    from fastintercu_vvc_b200 import MltPredictor

    p = MltPredictor(blob, device=0, max_batch=160)
    yield p
    p.close()


@pytest.fixture(scope="module")
def oracle(sd):
    m = OracleModel(sd)
    yield m
    m.close()


@pytest.fixture(scope="module")
def ctus():
    return ref_arch.synth_ctus(24, 10)


def as_list(orgpred, pocqp):
    return [(orgpred[i, 0], orgpred[i, 1], int(pocqp[i, 0]), int(pocqp[i, 1])) for i in range(len(orgpred))]


# ------------------------------------------------------------------------------------------- tcgen05 descriptor probe


def test_umma_descriptor_probe():
    """The smem-descriptor conventions conv_umma.cuh relies on hold on this GPU."""
    exe = os.path.join(ROOT, "tests", "cuda", "umma_probe.bin")
    if not os.path.exists(exe):
        subprocess.check_call(
            ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-I", os.path.join(ROOT, "fastintercu_vvc_b200", "csrc"),
             os.path.join(ROOT, "tests", "cuda", "umma_probe.cu"), "-o", exe]
        )
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout


# ------------------------------------------------------------------------------------------- staging (integer path)


def test_stage_bit_exact_vs_oracle_and_opencv_kat(pred, oracle, ctus):
    orgpred, pocqp = ctus
    got = pred.debug_stage(as_list(orgpred[:8], pocqp[:8]))
    for i in range(8):
        want = oracle.stage(orgpred[i, 0], orgpred[i, 1])
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32))
    kat = np.load(os.path.join(GOLD, "stage_kat.npz"))
    org = np.resize(kat["codes"].astype(np.int16), (128, 128))
    z = np.zeros((128, 128), np.int16)
    g = pred.debug_stage([(org, z, 0, 0)])[0]
    want = kat["cv2_convert_to"][org.astype(np.int64)]
    assert np.array_equal(g[0].view(np.uint32), want.view(np.uint32))
    assert np.array_equal(g[1].view(np.uint32), want.view(np.uint32))


def test_stage_edge_cases_strided_negative_extremes(pred, oracle):
    rng = np.random.RandomState(3)
    big = rng.randint(-300, 1400, (140, 200)).astype(np.int16)  # negative Pel -> uint16 > 1023 -> clamp to 1
    org = big[5:133, 17:145]
    prd = rng.randint(0, 1024, (131, 135)).astype(np.int16)[2:130, 3:131]
    z = np.zeros((128, 128), np.int16)
    f = np.full((128, 128), 1023, np.int16)
    cases = [(org, prd, 1, 30), (z, z, 0, 0), (f, z, 0, 0), (z, f, 0, 0), (prd, org, 5, 22)]
    got = pred.debug_stage(cases)
    for i, (o, p, _, _) in enumerate(cases):
        want = oracle.stage(o, p)
        assert np.array_equal(got[i].view(np.uint32), want.view(np.uint32)), i
    assert got.min() >= 0.0 and got.max() <= 1.0


# ------------------------------------------------------------------------------------------- network


def test_fp32_engine_matches_oracle(pred, oracle, ctus):
    """GPU fp32 cross-check engine vs C oracle: pins BN folding, shortcut and residual semantics."""
    orgpred, pocqp = ctus
    pred.set_engine(1)
    try:
        res = pred.predict_batch(as_list(orgpred[:6], pocqp[:6]))
    finally:
        pred.set_engine(0)
    lg, sp = oracle.predict_batch(orgpred[:6], pocqp[:6])
    err = np.abs(res["logits"] - lg).max()
    print("fp32 engine vs oracle max|dlogit| =", err)
    assert err < 3e-4
    assert np.array_equal(res["split_l3"], sp)


def test_tcgen05_layers_vs_fp32_engine(pred, ctus):
    """Layer-by-layer comparison on the device: tcgen05 (fp16 operands, fp32 accumulate) vs fp32 engine."""
    orgpred, pocqp = ctus
    n = 3  # odd: exercises the half-empty layer3 tile
    batch = as_list(orgpred[:n], pocqp[:n])
    pred.set_engine(1)
    pred.predict_batch(batch)
    ref = [pred.debug_activation(a, n) for a in range(17)]
    pred.set_engine(0)
    pred.predict_batch(batch)
    worst = 0.0
    for a in range(17):
        got = pred.debug_activation(a, n)
        scale = np.abs(ref[a]).max()
        err = np.abs(got - ref[a]).max() / max(scale, 1e-6)
        print(f"act {a:2d} shape {got.shape}: max|d|/max|ref| = {err:.3e} (max|ref| {scale:.3f})")
        worst = max(worst, err)
        assert err < 2e-2, f"activation {a} diverges: {err}"
    assert worst > 0  # the two engines really are different arithmetic


def test_fused_stem_matches_unfused_tcgen05_engine(pred, ctus):
    """Product path (staging + conv1 + layer0.0.conv1 fused in one kernel, conv1 output kept on chip) vs the unfused
    tcgen05 engine (standalone conv1 kernel + the generic conv kernel): same arithmetic up to fp32 summation order."""
    orgpred, pocqp = ctus
    n = 5
    batch = as_list(orgpred[:n], pocqp[:n])
    pred.set_engine(2)
    try:
        ref = pred.predict_batch(batch)
        ref_act1 = pred.debug_activation(1, n)
        ref_act2 = pred.debug_activation(2, n)
    finally:
        pred.set_engine(0)
    got = pred.predict_batch(batch)
    act1 = pred.debug_activation(1, n)
    act2 = pred.debug_activation(2, n)
    e1 = np.abs(act1 - ref_act1).max() / np.abs(ref_act1).max()
    e2 = np.abs(act2 - ref_act2).max() / np.abs(ref_act2).max()
    dl = np.abs(got["logits"] - ref["logits"]).max()
    print(f"fused vs unfused stem: act1 {e1:.2e}, act2 (uses the shortcut quarter of conv1) {e2:.2e}, max|dlogit| {dl:.2e}")
    assert e1 < 2e-3 and e2 < 2e-3  # an fp16 ulp here and there (the stem uses single fp16 conv1 weights, engine 2 hi + lo)
    assert dl < 2e-3
    assert np.array_equal(got["split_l3"], ref["split_l3"])


def test_tcgen05_matches_reference_golden(pred, ctus):
    """Product path vs logits of the reference's own arch file (tests/golden/logits_seed10.npz)."""
    gold = np.load(os.path.join(GOLD, "logits_seed10.npz"))
    orgpred, pocqp = ctus
    assert np.array_equal(pocqp, gold["pocqp"])
    res = pred.predict_batch(as_list(orgpred, pocqp))
    dl = np.abs(res["logits"] - gold["logits"]).max()
    dp = np.abs(res["probs"] - softmax_levels(gold["logits"])).max()
    print(f"tcgen05 vs reference golden: max|dlogit|={dl:.3e} max|dprob|={dp:.3e}")
    assert dp <= PROB_TOL
    # decisions: allow a flip only where the fp32 reference itself is within 2e-3 of a tie
    srt = np.sort(gold["logits"][:, 5:9], 1)
    clear = (srt[:, -1] - srt[:, -2]) > 2e-3
    assert np.array_equal(res["split_l3"][clear], gold["split"][clear])
    # probs are a softmax of the logits; flags follow the argmaxes
    assert np.abs(softmax_levels(res["logits"]) - res["probs"]).max() < 1e-6
    assert np.array_equal(res["split_l3"], res["logits"][:, 5:9].argmax(1))
    assert np.array_equal(res["split_l2"], res["logits"][:, 2:5].argmax(1))
    assert np.array_equal(res["split_l1"], res["logits"][:, 0:2].argmax(1))
    assert np.array_equal((res["flags"] >> 3) & 1, (res["split_l3"] == 1).astype(np.uint32))


def test_agreement_on_larger_sample(pred, oracle):
    """>= 99.9 % split agreement and <= 1e-3 probability error on 160 fresh CTUs (oracle on all host cores)."""
    orgpred, pocqp = ref_arch.synth_ctus(160, 2024)
    res = pred.predict_batch_dense(orgpred, pocqp)
    lg, sp = oracle.predict_batch(orgpred, pocqp)
    dp = np.abs(res["probs"] - softmax_levels(lg)).max()
    agree = (res["split_l3"] == sp).mean()
    srt = np.sort(lg[:, 5:9], 1)
    margin = srt[:, -1] - srt[:, -2]
    print(f"n=160: max|dprob|={dp:.3e}, agreement={agree:.4f}, min fp32 margin={margin.min():.2e}")
    assert dp <= PROB_TOL
    bad = res["split_l3"] != sp
    assert np.all(margin[bad] < 2e-3), "a disagreement away from a numerical tie"
    assert agree >= 0.999
    assert set(sp.tolist()) == {0, 1, 2, 3}


def test_agreement_2048_ctus_vs_fp32_oracle(blob, oracle):
    """BASELINE north_star bars on a bigger sample: max |dprob| <= 1e-3 vs the fp32 oracle and >= 99.9 % split-decision
    agreement on 2048 fresh synthetic CTUs (the fp32 C oracle runs on all host cores, ~20 s)."""
    from fastintercu_vvc_b200 import MltPredictor

    orgpred, pocqp = ref_arch.synth_ctus(2048, 4242)
    with MltPredictor(blob, device=0, max_batch=2048) as p:
        res = p.predict_batch_dense(orgpred, pocqp)  # 4 chunks, both activation sets
    lg, sp = oracle.predict_batch(orgpred, pocqp)
    dp = np.abs(res["probs"] - softmax_levels(lg)).max()
    agree = [(res[k] == lg[:, a:b].argmax(1)).mean() for k, (a, b) in (("split_l1", (0, 2)), ("split_l2", (2, 5)), ("split_l3", (5, 9)))]
    srt = np.sort(lg[:, 5:9], 1)
    margin = srt[:, -1] - srt[:, -2]
    bad = res["split_l3"] != sp
    print(f"n=2048: max|dprob|={dp:.3e}, agreement L1/L2/L3={agree[0]:.4f}/{agree[1]:.4f}/{agree[2]:.4f}, "
          f"L3 flips={int(bad.sum())} (fp32 margins {np.round(margin[bad], 5).tolist()}), classes={np.bincount(sp, minlength=4).tolist()}")
    assert dp <= PROB_TOL
    assert agree[2] >= 0.999 and min(agree) >= 0.998
    assert np.all(margin[bad] < 2e-3), "a disagreement away from a numerical tie"


def test_agreement_at_scale_20480_ctus(blob, oracle):
    """The north_star bars where they are tight: 20,480 fresh synthetic CTUs through the product path (tcgen05, fp16 operands) and
    through the fp32 CUDA-core cross-check engine, itself anchored to the fp32 C oracle on the first 256: max |dprob| <= 1e-3 and
    >= 99.9 % split-decision agreement at EVERY level; every disagreement is an fp32 tie (top-2 logit margin < 2e-3).  Deterministic:
    seeded inputs, bit-reproducible kernels -- the same counts on every run of a given build."""
    from fastintercu_vvc_b200 import MltPredictor

    N, B = 20480, 2048
    levels = (("split_l1", 0, 2), ("split_l2", 2, 5), ("split_l3", 5, 9))
    flips = {k: 0 for k, _, _ in levels}
    max_dp, worst_margin = 0.0, 0.0
    with MltPredictor(blob, device=0, max_batch=B) as p:
        for done in range(0, N, B):
            ctus_, pq = ref_arch.synth_ctus(B, 700000 + done)
            p.set_engine(0)
            a = p.predict_batch_dense(ctus_, pq).copy()
            p.set_engine(1)
            r = p.predict_batch_dense(ctus_, pq).copy()
            if done == 0:
                lg, _ = oracle.predict_batch(ctus_[:256], pq[:256])
                assert np.abs(r["logits"][:256] - lg).max() < 3e-4  # the comparison engine is the oracle's arithmetic
            max_dp = max(max_dp, float(np.abs(a["probs"] - r["probs"]).max()))
            for k, lo, hi in levels:
                bad = np.nonzero(a[k] != r[k])[0]
                flips[k] += len(bad)
                for i in bad:
                    top2 = np.sort(r["logits"][i, lo:hi])[-2:]
                    worst_margin = max(worst_margin, float(top2[1] - top2[0]))
    agree = {k: 1.0 - v / N for k, v in flips.items()}
    print(f"n={N}: max|dprob|={max_dp:.3e}, flips L1/L2/L3={flips['split_l1']}/{flips['split_l2']}/{flips['split_l3']}, "
          f"agreement {agree['split_l1']:.5f}/{agree['split_l2']:.5f}/{agree['split_l3']:.5f}, largest fp32 margin among the flips {worst_margin:.2e}")
    assert max_dp <= PROB_TOL
    assert min(agree.values()) >= 0.999
    assert worst_margin < 2e-3, "a disagreement away from a numerical tie"


def test_product_stem_follows_the_uint16_cast_semantics(pred, oracle):
    """Pel values outside 0..1023 through the PRODUCT stem (stem5_umma.cu stagers, not the staging probe): negative samples become
    large uint16 values and clamp to 1.0, |org - pred| is taken on the casts (EncCu.cpp:816,827,833,848-867).  Logits and
    probabilities must follow the oracle, and the stem's two outputs the fp32 engine's, as for in-range CTUs."""
    rng = np.random.RandomState(17)
    n = 12
    orgpred = rng.randint(-300, 1400, (n, 2, 128, 128)).astype(np.int16)
    orgpred[0] = -1          # every sample casts to 65535
    orgpred[1, 0] = 1023
    orgpred[1, 1] = -32768   # |1023 - 32768| clamps
    orgpred[2, 0] = 1024     # just past the clamp
    orgpred[2, 1] = 0
    orgpred[3, 0, :, :64] = 0  # a hard vertical edge through the block at the work-unit boundary
    pocqp = np.stack([rng.randint(0, 33, n), rng.randint(20, 50, n)], 1).astype(np.int32)
    res = pred.predict_batch_dense(orgpred, pocqp)
    act1 = pred.debug_activation(1, n)
    act2 = pred.debug_activation(2, n)  # consumes the stem's conv1 quarter through layer0.0's shortcut
    lg, sp = oracle.predict_batch(orgpred, pocqp)
    dp = np.abs(res["probs"] - softmax_levels(lg)).max()
    pred.set_engine(1)
    try:
        pred.predict_batch_dense(orgpred, pocqp)
        ref1, ref2 = pred.debug_activation(1, n), pred.debug_activation(2, n)
    finally:
        pred.set_engine(0)
    e1 = np.abs(act1 - ref1).max() / np.abs(ref1).max()
    e2 = np.abs(act2 - ref2).max() / np.abs(ref2).max()
    print(f"out-of-range Pel through the product stem: act1 {e1:.2e}, act2 {e2:.2e}, max|dprob| {dp:.2e}")
    assert e1 < 2e-3 and e2 < 2e-3
    assert dp <= PROB_TOL
    srt = np.sort(lg[:, 5:9], 1)
    bad = res["split_l3"] != sp
    assert np.all((srt[:, -1] - srt[:, -2])[bad] < 2e-3)


def test_batch_shapes_and_entry_points_agree(pred, ctus):
    """Ragged / odd batches and every entry point give bit-identical results per CTU."""
    orgpred, pocqp = ctus
    full = pred.predict_batch_dense(orgpred, pocqp)
    for n in (1, 2, 3, 5, 23):
        part = pred.predict_batch(as_list(orgpred[:n], pocqp[:n]))
        assert np.array_equal(part["logits"].view(np.uint32), full["logits"][:n].view(np.uint32)), n
    one = pred.predict_ctu(orgpred[7, 0], orgpred[7, 1], pocqp[7, 0], pocqp[7, 1])
    assert np.array_equal(one["logits"].view(np.uint32), full["logits"][7].view(np.uint32))
    assert pred.predict_batch([]).shape == (0,)
    again = pred.predict_batch_dense(orgpred, pocqp)
    assert again.tobytes() == full.tobytes()  # deterministic run to run


def test_picture_staging_path(pred):
    """mlt_begin_picture + per-CTU pred upload == explicit org pointer path (EncSlice-level staging)."""
    rng = np.random.RandomState(5)
    w, h, margin = 416, 240, 16
    buf = rng.randint(0, 1024, (h + 2 * margin, w + 2 * margin + 3)).astype(np.int16)
    pic = buf[margin : margin + h, margin : margin + w]  # strided view like a VTM PelStorage with margins
    pred.begin_picture(pic, poc=7)
    for (x, y) in ((0, 0), (128, 0), (256, 0)):  # the 3 eligible CTUs of a 416x240 picture (EncCu.cpp:755)
        p = rng.randint(0, 1024, (128, 128)).astype(np.int16)
        a = pred.predict_ctu_in_picture(x, y, p, qp=37)
        b = pred.predict_ctu(pic[y : y + 128, x : x + 128], p, 7, 37)
        assert a.tobytes() == b.tobytes()
    from fastintercu_vvc_b200 import MltError

    with pytest.raises(MltError):
        pred.predict_ctu_in_picture(384, 0, p, 37)  # partial CTU: outside the picture
    with pytest.raises(MltError):
        pred.predict_ctu_in_picture(0, 128, p, 37)


def test_cpp_hook_drives_the_library_like_the_encoder(pred, blob):
    """The C++ hook mirror (what INTEGRATION.md patches into EncCu.cpp) run over a picture's CTU raster: gate, per-CTU
    predict() with picture-strided pointers and the per-picture staging path give the ctypes binding's decisions."""
    hook = os.path.join(ROOT, "fastintercu_vvc_b200", "hook")
    exe = os.path.join(hook, "hook_encode_sim.bin")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", hook])
    rng = np.random.RandomState(11)
    w, h, stride, poc = 416, 240, 416 + 2 * 80 + 8, 9  # VTM-like margins; 3 eligible CTUs (EncCu.cpp:755)
    base, _ = ref_arch.synth_ctus(3, 77)
    org = rng.randint(0, 1024, (h, stride)).astype(np.int16)
    for i in range(3):
        org[0:128, 128 * i : 128 * i + 128] = base[i, 0]
    preds = [base[i, 1] for i in range(3)]
    qps = [37, 22, 41]
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        f.write(np.array([w, h, stride, poc, 3], np.int32).tobytes())
        f.write(org.tobytes())
        for q, p in zip(qps, preds):
            f.write(np.array([q], np.int32).tobytes())
            f.write(np.ascontiguousarray(p).tobytes())
        path = f.name
    try:
        env = dict(os.environ, MLT_WEIGHTS=blob, MLT_DEVICE="0")
        r = subprocess.run([exe, path], capture_output=True, text=True, timeout=120, env=env)
    finally:
        os.unlink(path)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = [list(map(int, l.split())) for l in r.stdout.strip().splitlines()]
    assert [(x, y) for x, y, _, _ in rows] == [(0, 0), (128, 0), (256, 0)]
    for (x, y, a, b), q, p in zip(rows, qps, preds):
        want = pred.predict_ctu(org[y : y + 128, x : x + 128], p, poc, q)["split_l3"]
        assert a == want and b == want


# ------------------------------------------------------------------------------- frame-level pre-pass (section 8f rank 2)
def _prepass_picture(w, h, seed):
    """org + reference luma as strided views (VTM PelStorage with margins) and per-CTU integer MVs that cover: zero,
    aligned, odd / even sub-word shifts, and windows hanging over every picture border."""
    from tests.oracle_lib import picture_ctus

    rng = np.random.RandomState(seed)
    base = (rng.randint(0, 1024, (h // 8 + 1, w // 8 + 1)).repeat(8, 0).repeat(8, 1)[:h, :w] * 3 // 4 + rng.randint(0, 256, (h, w))).astype(np.int16)
    margin = 16
    obuf = np.zeros((h + 2 * margin, w + 2 * margin + 3), np.int16)
    rbuf = np.zeros_like(obuf)
    org = obuf[margin : margin + h, margin : margin + w]
    ref = rbuf[margin : margin + h, margin : margin + w]
    org[:] = base
    ref[:] = np.clip(np.roll(base, (1, 2), (0, 1)).astype(np.int32) + rng.randint(-12, 13, (h, w)), 0, 1023).astype(np.int16)
    xy = picture_ctus(w, h)
    n = len(xy)
    mv = rng.randint(-24, 25, (n, 2)).astype(np.int16)
    special = [(0, 0), (8, 0), (-8, 16), (1, 0), (2, 0), (7, -1), (-3, 5), (-150, -140), (160, 150), (-129, 0), (0, 2000), (4000, -4000)]
    for i, v in enumerate(special[:n]):
        mv[i] = v
    return org, ref, xy, mv


@pytest.mark.parametrize("w,h", [(416, 240), (1920, 1080)])
def test_picture_prepass_gather_is_bit_exact(blob, oracle, w, h):
    """The device-built prediction blocks of mlt_predict_picture == the oracle's integer-MV gather, every sample."""
    from fastintercu_vvc_b200 import MltPredictor
    from tests.oracle_lib import picture_pred

    org, ref, xy, mv = _prepass_picture(w, h, 21)
    with MltPredictor(blob, device=0, max_batch=len(xy)) as p:
        p.begin_picture(org, poc=5)
        assert p.picture_ctu_count() == len(xy)
        for mvs in (None, mv):
            res = p.predict_picture(ref, 37, mv=mvs)
            assert len(res) == len(xy)
            got = p.debug_picture_pred()
            for i, (x, y) in enumerate(xy):
                mx, my = (0, 0) if mvs is None else mvs[i]
                assert np.array_equal(got[i], picture_pred(ref, x, y, mx, my)), (i, x, y, mx, my)


def test_picture_prepass_equals_per_ctu_calls(blob):
    """One frame-level batch == the drop-in per-CTU call fed the same prediction blocks: bit-identical results, in the
    raster order of EncSlice::encodeCtus; per-CTU QPs honoured; errors reported, never computed around."""
    from fastintercu_vvc_b200 import MltError, MltPredictor
    from tests.oracle_lib import picture_pred

    w, h = 1920, 1080
    org, ref, xy, mv = _prepass_picture(w, h, 22)
    n = len(xy)
    assert n == 120
    qps = np.random.RandomState(1).randint(22, 46, n).astype(np.int32)
    with MltPredictor(blob, device=0, max_batch=160) as p:
        with pytest.raises(MltError):
            p.predict_picture(ref, 32)  # no picture begun
        p.begin_picture(org, poc=9)
        before = p.launch_count
        res = p.predict_picture(ref, 0, mv=mv, ctu_qp=qps)
        assert p.launch_count - before == 18  # gather + stem + 15 convs + head: one batch, not 120 calls
        ctus = [(org[y : y + 128, x : x + 128], picture_pred(ref, x, y, *mv[i]), 9, int(qps[i])) for i, (x, y) in enumerate(xy)]
        want = p.predict_batch(ctus)
        assert res.tobytes() == want.tobytes()
        one = p.predict_ctu(*ctus[77])
        assert one.tobytes() == res[77].tobytes()
        same_qp = p.predict_picture(ref, 31)
        want0 = p.predict_batch([(o, picture_pred(ref, x, y), 9, 31) for (o, _, _, _), (x, y) in zip(ctus, xy)])
        assert same_qp.tobytes() == want0.tobytes()
        # page-locked picture buffers (mlt_pin_host_buffer, what a patched VTM does once per Picture): same bytes out
        for buf in (org.base, ref.base):
            p.pin_host_buffer(buf)
            p.pin_host_buffer(buf)  # pinning twice is not an error
        p.begin_picture(org, poc=9)
        assert p.predict_picture(ref, 0, mv=mv, ctu_qp=qps).tobytes() == res.tobytes()
        for buf in (org.base, ref.base):
            p.unpin_host_buffer(buf)
        with pytest.raises(MltError):
            p.unpin_host_buffer(ref.base)  # not pinned any more
    with MltPredictor(blob, device=0, max_batch=64) as small:
        small.begin_picture(org, poc=9)
        with pytest.raises(MltError) as e:
            small.predict_picture(ref, 32)
        assert e.value.rc == -7  # MLT_E_BATCH: 120 eligible CTUs > max_batch


@pytest.mark.parametrize("w,h,rng_", [(416, 240, 8), (416, 240, 16), (1920, 1080, 3), (416, 240, 0)])
def test_picture_block_matching_is_bit_exact(blob, w, h, rng_):
    """mlt_estimate_picture_mv (integer full search on the device) == the oracle's restatement: same MV and same cost for
    every eligible CTU (ties included: flat blocks are planted), and the MVs feed mlt_predict_picture with the
    reference plane reused on the device."""
    from fastintercu_vvc_b200 import MltError, MltPredictor
    from tests.oracle_lib import picture_me

    org, ref, xy, _ = _prepass_picture(w, h, 31)
    n = len(xy)
    ref[:] = np.clip(np.roll(org, (3, -2), (0, 1)).astype(np.int32) + np.random.RandomState(2).randint(-3, 4, org.shape), 0, 1023).astype(np.int16)
    org[0:128, 0:128] = 300  # a flat CTU against a flat reference area: every candidate inside it ties
    ref[0:140, 0:140] = 300
    with MltPredictor(blob, device=0, max_batch=max(n, 8)) as p:
        p.begin_picture(org, poc=5)
        with pytest.raises(MltError):
            p.estimate_picture_mv(None, rng_)  # no reference plane uploaded for this picture yet
        with pytest.raises(MltError):
            p.estimate_picture_mv(ref, 17)
        mv, cost = p.estimate_picture_mv(ref, rng_)
        assert mv.shape == (n, 2)
        for i, (x, y) in enumerate(xy):
            want_mv, want_cost = picture_me(org, ref, x, y, rng_)
            assert (int(mv[i, 0]), int(mv[i, 1])) == want_mv and int(cost[i]) == want_cost, (i, x, y)
        if rng_ >= 3:
            assert sum(1 for v in mv if tuple(v) == (-2, 3)) >= n // 2  # the planted motion is found
        mv2, cost2 = p.estimate_picture_mv(None, rng_)  # plane reused, run-to-run identical
        assert np.array_equal(mv, mv2) and np.array_equal(cost, cost2)
        a = p.predict_picture(None, 33, mv=mv)
        b = p.predict_picture(ref, 33, mv=mv)
        assert a.tobytes() == b.tobytes()
        p.begin_picture(org, poc=6)
        with pytest.raises(MltError):
            p.predict_picture(None, 33)  # the reference plane belonged to the previous picture


def test_prepass_matches_committed_golden(blob):
    """The CUDA pre-pass against tests/golden/prepass_seed10.npz: block-matching MVs / costs for three search ranges and the
    checksums of the gathered prediction blocks."""
    import hashlib

    from fastintercu_vvc_b200 import MltPredictor
    from tools.gen_golden_prepass import prepass_planes

    g = np.load(os.path.join(GOLD, "prepass_seed10.npz"))
    org, ref = prepass_planes()
    with MltPredictor(blob, device=0, max_batch=8) as p:
        p.begin_picture(org, poc=3)
        assert p.picture_ctu_count() == len(g["xy"])
        for R in (0, 4, 9):
            mv, cost = p.estimate_picture_mv(ref, R)
            assert np.array_equal(mv, g[f"mv_r{R}"]) and np.array_equal(cost, g[f"cost_r{R}"]), R
        p.predict_picture(None, 30, mv=g["mv_fixed"])
        for blk, want in zip(p.debug_picture_pred(), g["pred_sha256"]):
            assert hashlib.sha256(blk.tobytes()).hexdigest() == str(want)


def test_cpp_hook_prepass_drives_the_library(pred, blob):
    """The C++ hook's prepassPicture / pictureSplit pair over a 416x240 picture: decisions of
    the eligible CTUs equal the per-CTU drop-in call on the same prediction, partial CTUs carry none."""
    from tests.oracle_lib import picture_pred

    hook = os.path.join(ROOT, "fastintercu_vvc_b200", "hook")
    exe = os.path.join(hook, "hook_prepass_sim.bin")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", hook])
    w, h, poc, qp = 416, 240, 6, 38
    org, ref, xy, mv = _prepass_picture(w, h, 23)
    stride = org.strides[0] // 2
    full_o = np.zeros((h, stride), np.int16)
    full_r = np.zeros((h, stride), np.int16)
    full_o[:, :w] = org
    full_r[:, :w] = ref
    pred.begin_picture(org, poc)
    est_mv, _ = pred.estimate_picture_mv(ref, 6)
    for has_mv in (0, 1, -6):
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
            f.write(np.array([w, h, stride, poc, qp, has_mv], np.int32).tobytes())
            f.write(full_o.tobytes())
            f.write(full_r.tobytes())
            if has_mv == 1:
                f.write(mv.tobytes())
            path = f.name
        try:
            env = dict(os.environ, MLT_WEIGHTS=blob, MLT_DEVICE="0")
            r = subprocess.run([exe, path], capture_output=True, text=True, timeout=120, env=env)
        finally:
            os.unlink(path)
        assert r.returncode == 0, r.stdout + r.stderr
        rows = [tuple(map(int, l.split())) for l in r.stdout.strip().splitlines()]
        assert len(rows) == 4 * 2  # the whole CTU raster, partial CTUs included
        elig = [row for row in rows if row[2]]
        assert [(x, y) for x, y, _, _ in elig] == [tuple(v) for v in xy]
        assert all(s == -1 for _, _, e, s in rows if not e)
        for i, (x, y, _, s) in enumerate(elig):
            m = mv[i] if has_mv == 1 else (est_mv[i] if has_mv < 0 else (0, 0))
            want = pred.predict_ctu(org[y : y + 128, x : x + 128], picture_pred(ref, x, y, *m), poc, qp)["split_l3"]
            assert s == want


def test_device_resident_batch_via_torch(pred, ctus):
    import torch

    orgpred, pocqp = ctus
    host = pred.predict_batch_dense(orgpred, pocqp)
    d_in = torch.from_numpy(orgpred).cuda()
    d_pq = torch.from_numpy(pocqp).cuda()
    d_out = torch.zeros(len(orgpred) * 88, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    pred.predict_batch_device(len(orgpred), d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), st.cuda_stream)
    st.synchronize()
    from fastintercu_vvc_b200.capi import RESULT_DTYPE

    dev = np.frombuffer(d_out.cpu().numpy().tobytes(), RESULT_DTYPE)
    assert dev.tobytes() == host.tobytes()


def test_sliced_device_batch_matches_small_batches(blob, ctus):
    """n >= 480 device-resident batches run as two slices on two streams / activation sets (odd split included):
    every row must be bit-identical to the same CTU computed in a small single-slice batch."""
    import torch

    from fastintercu_vvc_b200 import MltPredictor
    from fastintercu_vvc_b200.capi import RESULT_DTYPE

    orgpred, pocqp = ctus
    n = 601
    idx = np.random.RandomState(7).randint(0, len(orgpred), n)
    big, big_pq = np.ascontiguousarray(orgpred[idx]), np.ascontiguousarray(pocqp[idx])
    with MltPredictor(blob, device=0, max_batch=640) as p:
        base = p.predict_batch_dense(orgpred, pocqp)  # 24 CTUs, one slice
        d_in, d_pq = torch.from_numpy(big).cuda(), torch.from_numpy(big_pq).cuda()
        d_out = torch.zeros(n * RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        st = torch.cuda.current_stream()
        for _ in range(2):  # twice: the second pass reuses both activation sets
            p.predict_batch_device(n, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), st.cuda_stream)
        st.synchronize()
        dev = np.frombuffer(d_out.cpu().numpy().tobytes(), RESULT_DTYPE)
        host = p.predict_batch_dense(big, big_pq)  # n < 1024: one chunk, one slice
    assert np.array_equal(dev["logits"].view(np.uint32), base["logits"][idx].view(np.uint32))
    assert dev.tobytes() == host.tobytes()


def test_device_calls_on_different_streams_and_host_calls_do_not_race(blob, ctus):
    """mlt_predict_batch_device is asynchronous on the caller's stream but uses the context's activation buffers: two device calls
    on DIFFERENT streams and a host-path call issued right behind them, with no synchronisation by the caller, must all give the
    small-batch results (the library orders them with an event); uncollected submitted batches refuse the device entry."""
    import torch

    from fastintercu_vvc_b200 import MltPredictor
    from fastintercu_vvc_b200.capi import RESULT_DTYPE, MltError

    orgpred, pocqp = ctus
    n = 601
    rs = np.random.RandomState(11)
    ia, ib = rs.randint(0, len(orgpred), n), rs.randint(0, len(orgpred), n)
    with MltPredictor(blob, device=0, max_batch=640) as p:
        base = p.predict_batch_dense(orgpred, pocqp)
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        bufs = []
        for idx in (ia, ib):
            bufs.append((torch.from_numpy(np.ascontiguousarray(orgpred[idx])).cuda(), torch.from_numpy(np.ascontiguousarray(pocqp[idx])).cuda(),
                         torch.zeros(n * RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")))
        torch.cuda.synchronize()
        for rep in range(3):
            for (d_in, d_pq, d_out), st in zip(bufs, (sa, sb)):
                p.predict_batch_device(n, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), st.cuda_stream)
            host = p.predict_batch_dense(orgpred, pocqp)  # host path right behind the two device calls, no sync in between
            one = p.predict_ctu(orgpred[3, 0], orgpred[3, 1], int(pocqp[3, 0]), int(pocqp[3, 1]))
            torch.cuda.synchronize()
            assert host.tobytes() == base.tobytes() and one.tobytes() == base[3].tobytes()
            for (_, _, d_out), idx in zip(bufs, (ia, ib)):
                dev = np.frombuffer(d_out.cpu().numpy().tobytes(), RESULT_DTYPE)
                assert np.array_equal(dev["logits"].view(np.uint32), base["logits"][idx].view(np.uint32)), rep
        p.submit_batch_dense(orgpred, pocqp)
        d_in, d_pq, d_out = bufs[0]
        with pytest.raises(MltError):
            p.predict_batch_device(n, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), sa.cuda_stream)
        assert p.collect().tobytes() == base.tobytes()


def test_packed10_transport_is_byte_identical_to_the_int16_path(blob, ctus):
    """10-bit packed transport (mlt_pack10 on the host, unpack10_kernel on the device): the blocking and the pipelined packed entry
    points must return exactly what mlt_predict_batch_dense returns on the same samples -- small batch, a chunked batch (>= 1024
    CTUs: growing chunk schedule, both activation sets) and two batches in flight."""
    from fastintercu_vvc_b200 import MltPredictor
    from fastintercu_vvc_b200.capi import pack10

    orgpred, pocqp = ctus
    n = 1100
    idx = np.random.RandomState(5).randint(0, len(orgpred), n)
    big, big_pq = np.ascontiguousarray(orgpred[idx]), np.ascontiguousarray(pocqp[idx])
    with MltPredictor(blob, device=0, max_batch=n) as p:
        base = p.predict_batch_dense(orgpred, pocqp)
        l0 = p.launch_count
        got = p.predict_batch_packed10(pack10(orgpred), pocqp)
        assert p.launch_count - l0 == 18  # unpack + stem + 15 convs + head
        assert got.tobytes() == base.tobytes()
        want = p.predict_batch_dense(big, big_pq)
        packed = pack10(big)
        assert p.predict_batch_packed10(packed, big_pq).tobytes() == want.tobytes()
        p.submit_batch_packed10(packed, big_pq)
        p.submit_batch_packed10(pack10(orgpred), pocqp)
        assert p.collect().tobytes() == want.tobytes()
        assert p.collect().tobytes() == base.tobytes()
        assert np.array_equal(want["logits"].view(np.uint32), base["logits"][idx].view(np.uint32))


def test_fp32_engine_large_host_batch_is_chunk_safe(blob, oracle):
    """The fp32 cross-check engine has one set of fp32 buffers: a host batch large enough to be chunked (>= 1024 CTUs) must
    not spread its chunks over the two compute streams (found by tools/precision_large.py: garbage at n = 2048)."""
    from fastintercu_vvc_b200 import MltPredictor

    ctus, pq = ref_arch.synth_ctus(1100, 77)
    with MltPredictor(blob, device=0, max_batch=1100) as p:
        p.set_engine(1)
        big = p.predict_batch_dense(ctus, pq).copy()
        small = np.concatenate([p.predict_batch_dense(ctus[i : i + 275], pq[i : i + 275]) for i in range(0, 1100, 275)])
        assert big.tobytes() == small.tobytes()
        lg, sp = oracle.predict_batch(ctus[:64], pq[:64])
        assert np.abs(big["logits"][:64] - lg).max() < 2e-4
        p.set_engine(0)
        prod = p.predict_batch_dense(ctus, pq)
        assert np.abs(prod["probs"] - big["probs"]).max() <= PROB_TOL


def test_errors(pred, blob, ctus):
    from fastintercu_vvc_b200 import MltError, MltPredictor

    orgpred, pocqp = ctus
    with pytest.raises(MltError) as e:
        MltPredictor("/nonexistent/weights.mltw")
    assert e.value.rc == -2
    with tempfile.NamedTemporaryFile(suffix=".mltw") as f:
        f.write(b"not a blob" * 100)
        f.flush()
        with pytest.raises(MltError) as e:
            MltPredictor(f.name)
        assert e.value.rc == -3
    big = np.zeros((161, 2, 128, 128), np.int16)
    with pytest.raises(MltError) as e:
        pred.predict_batch_dense(big, np.zeros((161, 2), np.int32))
    assert e.value.rc == -7
    with pytest.raises(MltError):
        pred.set_engine(5)
    # the context is still usable after errors
    r = pred.predict_ctu(orgpred[0, 0], orgpred[0, 1], pocqp[0, 0], pocqp[0, 1])
    assert 0 <= r["split_l3"] <= 3


def test_full_size_batch_properties():
    """BASELINE config 5 size (4096 CTUs): duplicates give identical rows, a sample matches the oracle."""
    from fastintercu_vvc_b200 import MltPredictor
    from fastintercu_vvc_b200.pack_weights import write_blob

    sd = ref_arch.make_state_dict(10)
    with tempfile.NamedTemporaryFile(suffix=".mltw") as f:
        write_blob(sd, f.name)
        with MltPredictor(f.name, max_batch=4096) as p:
            base, pq = ref_arch.synth_ctus(64, 99)
            idx = np.random.RandomState(1).randint(0, 64, 4096)
            idx[:64] = np.arange(64)
            orgpred = np.ascontiguousarray(base[idx])
            pocqp = np.ascontiguousarray(pq[idx])
            res = p.predict_batch_dense(orgpred, pocqp)
            first = res["logits"][:64]
            assert np.array_equal(res["logits"].view(np.uint32), first[idx].view(np.uint32))
            m = OracleModel(sd)
            lg, sp = m.predict_batch(base[:16], pq[:16])
            assert np.abs(softmax_levels(first[:16]) - softmax_levels(lg)).max() <= PROB_TOL


def test_pipelined_submit_collect_matches_the_synchronous_call():
    """mlt_submit_batch_dense / mlt_collect: two batches in flight, results bit-identical to mlt_predict_batch_dense,
    FIFO order, state errors reported (never a silent overwrite of a buffer in flight)."""
    from fastintercu_vvc_b200 import MltError, MltPredictor
    from fastintercu_vvc_b200.pack_weights import write_blob

    with tempfile.NamedTemporaryFile(suffix=".mltw", delete=False) as f:
        path = f.name
    write_blob(ref_arch.make_state_dict(10), path)
    base, pq = ref_arch.synth_ctus(24, 91)
    batches = []
    for k, n in enumerate((1100, 7, 1500, 1024)):  # chunked and single-pass batches mixed
        idx = (np.arange(n) * (k + 3)) % 24
        batches.append((np.ascontiguousarray(base[idx]), np.ascontiguousarray(pq[idx])))
    with MltPredictor(path, device=0, max_batch=1500) as p:
        want = [p.predict_batch_dense(o, q).copy() for o, q in batches]
        with pytest.raises(MltError) as e:
            p.collect()
        assert e.value.rc == -8
        p.submit_batch_dense(*batches[0])
        p.submit_batch_dense(*batches[1])
        with pytest.raises(MltError) as e:
            p.submit_batch_dense(*batches[2])  # two already in flight
        assert e.value.rc == -8
        with pytest.raises(MltError) as e:
            p.predict_batch_dense(*batches[2])  # synchronous calls are refused while batches are in flight
        assert e.value.rc == -8
        got0 = p.collect().copy()
        p.submit_batch_dense(*batches[2])
        got1 = p.collect().copy()
        p.submit_batch_dense(*batches[3])
        got2 = p.collect().copy()
        got3 = p.collect().copy()
        for g, w in zip((got0, got1, got2, got3), want):
            assert g.tobytes() == w.tobytes()
        assert p.predict_batch_dense(*batches[1]).tobytes() == want[1].tobytes()  # synchronous path usable again
    os.unlink(path)


# ------------------------------------------------------------------------------------------- inside the encoder


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "EncoderApp_mlt")), reason="oracle/_ref encoders not built")
def test_patched_vtm_encoder_matches_the_reference_hook_encoder(tmp_path):
    """The hot path where the reference runs it: VTM-11.0's EncCu::xCompressCU.  A short RA encode (1 I + 2 B pictures of 416x240,
    6 predictor calls) with (i) the reference's own hook translation unit on libtorch CPU and (ii) the patched encoder calling
    libmltcnn.so through hook/mlt_hook.cpp must take the same split decisions, write the same bitstream and reconstruction, and the
    bitstream must decode to that reconstruction (oracle/vtm/Makefile builds both from /root/reference; tools/vtm_run.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import vtm_run

    work = str(tmp_path)
    model_dir, blob = vtm_run.make_weights(work)
    clip = {"name": "t416", "w": 416, "h": 240, "bits": 8, "frames": 3, "path": os.path.join(work, "t.yuv")}
    vtm_run.synth_clip(clip["path"], 416, 240, 3, 8)
    res = [vtm_run.run_encode(e, clip, 32, work, model_dir, blob, 0, "2.1") for e in ("ref_cpu", "mlt", "staged")]
    for r in res:
        assert r["rc"] == 0 and r["decode_matches_recon"], r
        assert r["error_lines"] == 0 and r["hello_lines"] == 0, r
    ref, mlt, staged = res
    assert len(ref["trace"]) == 6 and all(t[4] in (0, 1, 2, 3) for t in ref["trace"])
    assert mlt["trace"] == ref["trace"] and staged["trace"] == ref["trace"]
    assert mlt["bitstream_md5"] == ref["bitstream_md5"] == staged["bitstream_md5"]
    assert mlt["recon_md5"] == ref["recon_md5"]
    assert mlt["hook_calls"] == 6


def test_cluster_chain_kernel_is_bit_identical_to_the_per_layer_kernels(blob, ctus, tmp_path):
    """MLT_CHAIN=1 runs convs 1..15 of a one- / two-CTU call as ONE kernel on a 16-CTA cluster (cluster barriers between the layers,
    mbarriers re-initialised per layer, TMEM allocated once): same tiles, same MMA order -> byte-identical results, 3 launches per call."""
    orgpred, pocqp = ctus
    np.save(str(tmp_path / "chain_ctus.npy"), orgpred[:6])
    np.save(str(tmp_path / "chain_pq.npy"), pocqp[:6])
    code = f"""
import numpy as np, sys
sys.path.insert(0, {ROOT!r})
import fastintercu_vvc_b200 as pkg
o, q = np.load({str(tmp_path / 'chain_ctus.npy')!r}), np.load({str(tmp_path / 'chain_pq.npy')!r})
with pkg.MltPredictor({blob!r}, device=0, max_batch=8) as p:
    full = p.predict_batch_dense(o, q)
    l0 = p.launch_count
    one = [p.predict_ctu(o[i, 0], o[i, 1], int(q[i, 0]), int(q[i, 1])).tobytes() for i in range(6)]
    per_call = (p.launch_count - l0) / 6
    pair = p.predict_batch_dense(o[2:4], q[2:4]).tobytes()
print(per_call, all(one[i] == full[i].tobytes() for i in range(6)), pair == full[2:4].tobytes())
"""
    out = {}
    for tag, env in (("chain", {"MLT_CHAIN": "1"}), ("layers", {})):
        e = {k: v for k, v in os.environ.items() if k not in ("MLT_CHAIN", "MLT_NO_CHAIN")}
        e.update(env)
        r = subprocess.run(["timeout", "-s", "KILL", "120", sys.executable, "-c", code], env=e, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-500:]
        out[tag] = r.stdout.split()
    assert out["chain"] == ["3.0", "True", "True"], out
    assert out["layers"] == ["17.0", "True", "True"], out


@pytest.mark.parametrize("stagers", ["4", "8"])
def test_stem_pipeline_survives_many_units_per_cta(blob, ctus, stagers, tmp_path):
    """Regression for a parity-aliasing deadlock in stem5_umma_kernel: the border-term barrier (corr_full) was not gated on its
    consumer, so its producer could run two phases ahead and the epilogue's parity wait never returned.  It needed ~25+ units per
    CTA and fast stagers (two stager groups, MLT_STEM5_STAGERS=8) to show.  900 CTUs = 49 units per CTA, in a subprocess under a
    hard timeout; results must equal the small-batch ones."""
    orgpred, pocqp = ctus
    np.save(str(tmp_path / "c.npy"), orgpred)
    np.save(str(tmp_path / "q.npy"), pocqp)
    code = f"""
import numpy as np, sys
sys.path.insert(0, {ROOT!r})
import fastintercu_vvc_b200 as pkg
o, q = np.load({str(tmp_path / 'c.npy')!r}), np.load({str(tmp_path / 'q.npy')!r})
idx = np.arange(900) % len(o)
with pkg.MltPredictor({blob!r}, device=0, max_batch=900) as p:
    small = p.predict_batch_dense(o, q)
    for _ in range(3):
        big = p.predict_batch_dense(np.ascontiguousarray(o[idx]), np.ascontiguousarray(q[idx]))
print(bool(np.array_equal(big["logits"].view(np.uint32), small["logits"][idx].view(np.uint32))))
np.save(sys.argv[1], big["logits"])
"""
    got = {}
    for bw in ("1", "0"):  # border terms on their own warp (default) / on the stager warps: the same fp32 expressions, so bit-equal
        e = dict(os.environ, MLT_STEM5_STAGERS=stagers, MLT_STEM5_BW=bw)
        out = str(tmp_path / f"l{bw}.npy")
        r = subprocess.run(["timeout", "-s", "KILL", "120", sys.executable, "-c", code, out], env=e, capture_output=True, text=True)
        assert r.returncode == 0, f"bw={bw} rc {r.returncode} (killed by the timeout = the pipeline hung): {r.stderr[-400:]}"
        assert r.stdout.split() == ["True"], r.stdout
        got[bw] = np.load(out)
    assert np.array_equal(got["1"].view(np.uint32), got["0"].view(np.uint32))
