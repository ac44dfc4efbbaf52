"""The smaller-CU oracle (64 / 32 / 16 px) is pinned before it is trusted (CPU-only).

Golden vectors: tests/golden/cu_logits_seed10.npz, produced by tools/gen_golden_cu.py from the REFERENCE's own
mlt_cu_or_pq_arch.py (`GapBigMltCuORPQ`, eager and traced as model2torchScript.py:37-48)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import ref_arch
from tests.oracle_lib import OracleCuModel

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF_ARCH = "/root/reference/mlt-cnn-python/codes/models/archs/mlt_cu_or_pq_arch.py"


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "cu_logits_seed10.npz"))


@pytest.mark.parametrize("size", ref_arch.CU_SIZES)
def test_seeded_cu_inputs_and_params_are_the_golden_ones(gold, size):
    orgpred, pocqp = ref_arch.synth_cus(int(gold["n"]), size, int(gold["seed"]))
    assert orgpred.shape == (int(gold["n"]), 2, size, size)
    assert hashlib.sha256(orgpred.tobytes()).digest() == gold[f"orgpred_sha256_{size}"].tobytes()
    assert np.array_equal(pocqp, gold[f"pocqp_{size}"])
    sd = ref_arch.make_cu_state_dict(int(gold["seed"]), size)
    h = hashlib.sha256(b"".join(np.ascontiguousarray(sd[k]).tobytes() for k in sorted(sd))).digest()
    assert h == gold[f"params_sha256_{size}"].tobytes()


@pytest.mark.parametrize("size", ref_arch.CU_SIZES)
def test_c_oracle_matches_reference_cu_golden_logits(gold, size):
    """Plain-C restatement vs the reference arch's logits (fp32 reassociation only); level-1 decisions identical."""
    sd = ref_arch.make_cu_state_dict(int(gold["seed"]), size)
    orgpred, pocqp = ref_arch.synth_cus(int(gold["n"]), size, int(gold["seed"]))
    m = OracleCuModel(sd, size)
    lg = m.predict_batch(orgpred, pocqp)
    err = np.abs(lg - gold[f"logits_{size}"]).max()
    assert err < 2e-4, err
    assert np.array_equal(lg[:, :2].argmax(1), gold[f"split_{size}"])
    assert set(gold[f"split_{size}"].tolist()) == {0, 1}
    assert np.abs(gold[f"logits_traced_{size}"] - gold[f"logits_{size}"]).max() == 0.0


@pytest.mark.parametrize("size", ref_arch.CU_SIZES)
def test_cu_stage_bit_exact_vs_numpy(size):
    rng = np.random.RandomState(size)
    big = rng.randint(-300, 1400, (size + 9, size + 21)).astype(np.int16)  # negatives and > 1023: cast + clamp
    org = big[4 : 4 + size, 13 : 13 + size]
    pred = np.ascontiguousarray(rng.randint(0, 1024, (size, size)).astype(np.int16))
    m = OracleCuModel(ref_arch.make_cu_state_dict(10, size), size)
    x = m.stage(org, pred)
    want = ref_arch.stage_numpy(np.stack([np.ascontiguousarray(org), pred])[None])[0]
    assert np.array_equal(x.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("size", ref_arch.CU_SIZES)
def test_torch_cu_restatement_matches_golden(gold, size):
    sd = ref_arch.make_cu_state_dict(int(gold["seed"]), size)
    orgpred, pocqp = ref_arch.synth_cus(8, size, int(gold["seed"]))
    lg = ref_arch.forward_cu_logits(ref_arch.build_cu_model(sd), ref_arch.stage_numpy(orgpred), pocqp, batch=1)
    assert np.abs(lg - gold[f"logits_{size}"][:8]).max() < 1e-5


@pytest.mark.skipif(not os.path.exists(REF_ARCH), reason="reference mount absent (GPU box)")
def test_torch_cu_restatement_equals_reference_arch_file():
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_cu_arch_file", REF_ARCH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for size in ref_arch.CU_SIZES:
        sd = ref_arch.make_cu_state_dict(3, size)
        ref = mod.GapBigMltCuORPQ()
        ref.load_state_dict(ref_arch.to_torch_state_dict(sd), strict=True)
        ref.eval()
        orgpred, pocqp = ref_arch.synth_cus(5, size, 77)
        x = ref_arch.stage_numpy(orgpred)
        a = ref_arch.forward_cu_logits(ref, x, pocqp, batch=1)
        b = ref_arch.forward_cu_logits(ref_arch.build_cu_model(sd), x, pocqp, batch=1)
        assert np.array_equal(a, b)
