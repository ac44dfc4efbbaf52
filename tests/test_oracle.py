"""The oracle is pinned before it is trusted (CPU-only tests).

Golden vectors come from the REFERENCE's own architecture file run in the build container
(tools/gen_golden.py); the OpenCV convertTo KAT comes from cv2 (same cvt_32f arithmetic the hook
calls at EncCu.cpp:835-838)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import ref_arch
from tests.oracle_lib import OracleModel

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF_ARCH = "/root/reference/mlt-cnn-python/codes/models/archs/mlt_ctu_or_pq_arch.py"


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "logits_seed10.npz"))


@pytest.fixture(scope="module")
def sd():
    return ref_arch.make_state_dict(10)


@pytest.fixture(scope="module")
def model(sd):
    m = OracleModel(sd)
    yield m
    m.close()


def test_seeded_inputs_and_params_are_the_golden_ones(gold, sd):
    orgpred, pocqp = ref_arch.synth_ctus(int(gold["n"]), int(gold["seed"]))
    assert hashlib.sha256(orgpred.tobytes()).digest() == gold["orgpred_sha256"].tobytes()
    assert np.array_equal(pocqp, gold["pocqp"])
    h = hashlib.sha256(b"".join(np.ascontiguousarray(sd[k]).tobytes() for k in sorted(sd))).digest()
    assert h == gold["params_sha256"].tobytes()


def test_stage_matches_opencv_kat():
    """C oracle staging == cv2 convertTo on every 10-bit code, bit for bit."""
    kat = np.load(os.path.join(GOLD, "stage_kat.npz"))
    codes = kat["codes"].astype(np.int16)
    org = np.resize(codes, (128, 128)).astype(np.int16)  # all 1024 codes, 16 times
    pred = np.zeros((128, 128), np.int16)
    m = OracleModel(ref_arch.make_state_dict(10))
    x = m.stage(org, pred)
    want = kat["cv2_convert_to"][org.astype(np.int64)]
    assert np.array_equal(x[0].view(np.uint32), want.view(np.uint32))
    assert np.array_equal(x[1].view(np.uint32), want.view(np.uint32))  # |org-0| == org
    assert np.float32(1.0 / 1023).view(np.uint32) == 0x3A802008


def test_stage_bit_exact_vs_numpy_and_golden_hash(gold, model):
    orgpred, _ = ref_arch.synth_ctus(int(gold["n"]), int(gold["seed"]))
    want = ref_arch.stage_numpy(orgpred)
    assert hashlib.sha256(want.tobytes()).digest() == gold["staged_sha256"].tobytes()
    for i in range(len(orgpred)):
        got = model.stage(orgpred[i, 0], orgpred[i, 1])
        assert np.array_equal(got.view(np.uint32), want[i].view(np.uint32))


def test_stage_edge_cases(model):
    """Strided buffers, org<pred, negative Pel (cast to uint16 > 1023 -> clamp to 1), extremes."""
    rng = np.random.RandomState(3)
    big = rng.randint(-300, 1400, (140, 200)).astype(np.int16)  # includes negatives and >1023
    org = big[5:133, 17:145]
    pred = np.ascontiguousarray(rng.randint(0, 1024, (128, 128)).astype(np.int16))
    x = model.stage(org, pred)
    o = org.astype(np.uint16).astype(np.int64)
    p = pred.astype(np.uint16).astype(np.int64)
    a = np.float32(1.0 / 1023)
    assert np.array_equal(x[0], np.clip(o.astype(np.float32) * a, 0, 1).astype(np.float32))
    assert np.array_equal(x[1], np.clip(np.abs(o - p).astype(np.float32) * a, 0, 1).astype(np.float32))
    assert x.min() >= 0.0 and x.max() <= 1.0
    z = np.zeros((128, 128), np.int16)
    f = np.full((128, 128), 1023, np.int16)
    assert np.all(model.stage(z, z) == 0)
    x = model.stage(f, z)
    assert np.all(x == 1.0)
    x = model.stage(z, f)
    assert np.all(x[0] == 0) and np.all(x[1] == 1.0)


def test_c_oracle_matches_reference_golden_logits(gold, model):
    """C restatement vs the reference arch's logits (fp32 reassociation only)."""
    orgpred, pocqp = ref_arch.synth_ctus(int(gold["n"]), int(gold["seed"]))
    lg, sp = model.predict_batch(orgpred, pocqp)
    err = np.abs(lg - gold["logits"]).max()
    assert err < 2e-4, err
    assert np.array_equal(sp, gold["split"])
    assert set(gold["split"].tolist()) == {0, 1, 2, 3}  # every setNewModeList branch is reachable
    assert np.abs(gold["logits_traced"] - gold["logits"]).max() == 0.0


def test_torch_restatement_matches_golden(gold, sd):
    orgpred, pocqp = ref_arch.synth_ctus(8, int(gold["seed"]))
    net = ref_arch.build_model(sd)
    lg = ref_arch.forward_logits(net, ref_arch.stage_numpy(orgpred), pocqp, batch=1)
    assert np.abs(lg - gold["logits"][:8]).max() < 1e-5


@pytest.mark.skipif(not os.path.exists(REF_ARCH), reason="reference mount absent (GPU box)")
def test_torch_restatement_equals_reference_arch_file(sd):
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_arch_file", REF_ARCH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = mod.GapBigMltCtuORPQ()
    ref.load_state_dict(ref_arch.to_torch_state_dict(sd), strict=True)
    ref.eval()
    orgpred, pocqp = ref_arch.synth_ctus(3, 77)
    x = ref_arch.stage_numpy(orgpred)
    a = ref_arch.forward_logits(ref, x, pocqp, batch=1)
    b = ref_arch.forward_logits(ref_arch.build_model(sd), x, pocqp, batch=1)
    assert np.array_equal(a, b)
    assert sum(p.numel() for p in ref.parameters()) == 2797403


# ------------------------------------------------------------------------------- frame-level pre-pass (section 8f rank 2)
def test_picture_ctu_raster_follows_the_gate():
    """Eligible CTUs per picture = the counts SURVEY.md section 8d derives from EncCu.cpp:755, raster order."""
    from tests.oracle_lib import picture_ctus

    for (w, h), want in (((416, 240), 3), ((1920, 1080), 120), ((3840, 2160), 480), ((127, 500), 0), ((128, 128), 1)):
        xy = picture_ctus(w, h)
        assert len(xy) == want
        ref = [(x, y) for y in range(0, h, 128) for x in range(0, w, 128) if x + 128 <= w and y + 128 <= h]
        assert [tuple(v) for v in xy] == ref


def test_picture_pred_equals_padded_reference_window():
    """Integer-MV prediction from a border-replicated plane: the C restatement against numpy's edge padding, which is
    what Picture::extendPicBorder (Picture.cpp:1117) builds around every reference picture."""
    from tests.oracle_lib import picture_pred

    rng = np.random.RandomState(3)
    w, h, m = 416, 240, 300
    buf = rng.randint(0, 1024, (h, w + 24)).astype(np.int16)
    ref = buf[:, 5 : 5 + w]  # strided view
    padded = np.pad(ref, m, mode="edge")
    for (x, y, mvx, mvy) in ((0, 0, 0, 0), (128, 0, 3, -2), (256, 0, -7, 5), (0, 0, -200, -150), (256, 0, 290, 250), (128, 0, 8, 16), (256, 0, 33, 113)):
        got = picture_pred(ref, x, y, mvx, mvy)
        want = padded[m + y + mvy : m + y + mvy + 128, m + x + mvx : m + x + mvx + 128]
        assert np.array_equal(got, want), (x, y, mvx, mvy)
    for size, (x, y, mvx, mvy) in ((64, (320, 128, 9, -3)), (32, (384, 192, 40, 60)), (16, (400, 224, -1, 1)), (16, (0, 0, -17, -33))):
        got = picture_pred(ref, x, y, mvx, mvy, size=size)
        want = padded[m + y + mvy : m + y + mvy + size, m + x + mvx : m + x + mvx + size]
        assert np.array_equal(got, want), (size, x, y, mvx, mvy)


def test_picture_me_finds_planted_motion_and_prefers_zero_mv():
    """Block matching restatement: a reference that is the original shifted by (dx, dy) gives MV (dx, dy) at cost 0 for
    interior CTUs; a flat picture (every candidate ties) gives the zero MV; cost == numpy's SAD at the winner."""
    from tests.oracle_lib import picture_me, picture_pred

    rng = np.random.RandomState(9)
    w, h = 416, 240
    org = rng.randint(0, 1024, (h, w)).astype(np.int16)
    for (dx, dy) in ((3, -2), (-5, 4), (0, 0), (6, 6)):
        # ref(x + dx, y + dy) == org(x, y) wherever the shifted position is inside the picture
        ref = np.roll(org, (dy, dx), (0, 1))
        mv, cost = picture_me(org, ref, 128, 0, 6)
        if 0 + dy >= 0:  # the CTU at (128, 0) only sees un-wrapped rows when dy >= 0
            assert mv == (dx, dy) and cost == 0, (dx, dy, mv, cost)
        blk = org[0:128, 128:256].astype(np.int64)
        assert cost == np.abs(blk - picture_pred(ref, 128, 0, *mv).astype(np.int64)).sum()
    flat = np.full((h, w), 512, np.int16)
    assert picture_me(flat, flat, 256, 0, 4) == ((0, 0), 0)


def test_prepass_oracle_matches_committed_golden():
    """tests/golden/prepass_seed10.npz (tools/gen_golden_prepass.py): eligible-CTU raster, block-matching MVs / costs for three
    search ranges, checksums of the gathered prediction blocks (CTU and 16-px forms)."""
    import hashlib

    from tests.oracle_lib import picture_ctus, picture_me, picture_pred
    from tools.gen_golden_prepass import prepass_planes

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "prepass_seed10.npz"))
    org, ref = prepass_planes()
    h, w = org.shape
    xy = picture_ctus(w, h)
    assert np.array_equal(xy, g["xy"]) and len(xy) == 6
    for R in (0, 4, 9):
        res = [picture_me(org, ref, x, y, R) for x, y in xy]
        assert np.array_equal(np.array([m for m, _ in res], np.int16), g[f"mv_r{R}"])
        assert np.array_equal(np.array([c for _, c in res], np.uint32), g[f"cost_r{R}"])
    assert (g["mv_r9"] == (-2, 3)).all(1).sum() >= 4  # the planted (-2, 3) motion
    for (x, y), mv, want in zip(xy, g["mv_fixed"], g["pred_sha256"]):
        assert hashlib.sha256(picture_pred(ref, x, y, *mv).tobytes()).hexdigest() == str(want)
    cu = b"".join(picture_pred(ref, (i % (w // 16)) * 16, (i // (w // 16)) * 16, i % 7 - 3, i % 5 - 2, size=16).tobytes() for i in range((w // 16) * (h // 16)))
    assert hashlib.sha256(cu).hexdigest() == str(g["cu16_pred_sha256"])


def test_picture_pred_property_random_geometry():
    """Property check over random picture sizes, block positions and MVs (including far outside): the C restatement ==
    numpy edge padding."""
    from tests.oracle_lib import picture_pred

    rng = np.random.RandomState(123)
    for _ in range(40):
        w, h = int(rng.randint(16, 300)), int(rng.randint(16, 300))
        size = int(rng.choice([16, 32, 64, 128]))
        if size > w or size > h:
            size = 16
        ref = rng.randint(-2000, 2000, (h, w + 3)).astype(np.int16)[:, 1 : 1 + w]
        x, y = int(rng.randint(0, w - size + 1)), int(rng.randint(0, h - size + 1))
        mvx, mvy = (int(v) for v in rng.randint(-400, 401, 2))
        m = 420
        padded = np.pad(ref, m, mode="edge")
        want = padded[m + y + mvy : m + y + mvy + size, m + x + mvx : m + x + mvx + size]
        assert np.array_equal(picture_pred(ref, x, y, mvx, mvy, size=size), want), (w, h, size, x, y, mvx, mvy)


def test_cpp_libtorch_runner_matches_the_oracle(model, tmp_path):
    """oracle/ref_libtorch.bin -- the hook's libtorch call sequence (EncCu.cpp:806-921) as a C++ program, the CPU baseline
    timer of bench.py -- gives the C oracle's split decisions and logits on seeded CTUs, in both load modes."""
    import subprocess

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "ref_libtorch.bin")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(root, "oracle")])
    sd = ref_arch.make_state_dict(10)
    traced = torch.jit.trace(ref_arch.build_model(sd), (torch.rand(1, 2, 128, 128), torch.rand(1), torch.rand(1)))
    pt, ctu_file = str(tmp_path / "m.pt"), str(tmp_path / "c.bin")
    traced.save(pt)
    ctus, pq = ref_arch.synth_ctus(8, 10)
    with open(ctu_file, "wb") as f:
        f.write(np.int32(8).tobytes())
        for i in range(8):
            f.write(pq[i].astype(np.int32).tobytes() + np.ascontiguousarray(ctus[i, 0]).tobytes() + np.ascontiguousarray(ctus[i, 1]).tobytes())
    lg, sp = model.predict_batch(ctus, pq)
    for mode in ("0", "1"):
        r = subprocess.run([exe, pt, ctu_file, "0.2", "2", mode], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        rows = [l.split() for l in r.stdout.splitlines() if l.startswith("split")]
        assert len(rows) == 8
        for i, row in enumerate(rows):
            assert int(row[2]) == int(sp[i])
            assert np.abs(np.array(row[3:7], np.float64) - lg[i, 5:9]).max() < 2e-5
        n, dt, thr = r.stdout.strip().splitlines()[-1].split()
        assert int(n) >= 8 and float(dt) > 0 and int(thr) == 2
