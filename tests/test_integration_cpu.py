"""CPU-only tests of the in-VTM integration tooling: the patch applier (anchors against the reference tree when it is present),
the ed-form patch file, the synthetic clip generator, BD-rate, and -- when oracle/_ref was built here -- the reference's own hook
statements (oracle/_ref/ref_hook_tu_cpu) against the C oracle on the same CTUs."""
import hashlib
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ENC = "/root/reference/vtm-mlt-cpp/source/Lib/EncoderLib"
REF_TU = os.path.join(ROOT, "oracle", "_ref", "ref_hook_tu_cpu")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


patcher = _load(os.path.join(ROOT, "integration", "apply_vtm_patch.py"), "apply_vtm_patch")
vtm_run = _load(os.path.join(ROOT, "tools", "vtm_run.py"), "vtm_run_mod")


def test_ed_patch_file_is_what_the_applier_emits():
    assert open(os.path.join(ROOT, "integration", "vtm_mlt.patch")).read() == patcher.emit_ed()


def test_patch_edits_are_disjoint_and_ordered():
    for name, edits in patcher.FILES.items():
        spans = sorted((f, max(l, f - 1)) for f, l, *_ in edits)
        for (a0, a1), (b0, _b1) in zip(spans, spans[1:]):
            assert a1 < b0, (name, a0, a1, b0)


def test_patch_applies_to_a_synthetic_file_and_rejects_drift():
    n = 1000
    lines = [f"line {i + 1}\n" for i in range(n)]
    for f, l, a0, a1, _ in patcher.ENC_CU_EDITS:
        lines[f - 1] = f"x {a0} x\n"
        if l >= f:
            lines[l - 1] = f"y {a1} y\n" if l != f else lines[f - 1]
    out = patcher.apply_edits(lines, patcher.ENC_CU_EDITS, "EncCu.cpp")
    text = "".join(out)
    assert '#include "mlt_hook.h"' in text and "predictAt(" in text and "if (predictedSplitMode >= 0)" in text
    assert "line 59\n" in text and "line 66\n" in text and "line 60\n" not in text and "line 900\n" not in text  # only the ranges went
    assert "line 927\n" in text and "line 929\n" in text
    lines[803] = "something else\n"  # drift outside anchors is fine ...
    patcher.apply_edits(lines, patcher.ENC_CU_EDITS, "EncCu.cpp")
    lines[802] = "drifted\n"  # ... on an anchor line it must fail loudly
    with pytest.raises(SystemExit):
        patcher.apply_edits(lines, patcher.ENC_CU_EDITS, "EncCu.cpp")


@pytest.mark.skipif(not os.path.isdir(REF_ENC), reason="reference tree not present (GPU box)")
def test_patch_anchors_hold_on_the_reference_tree(tmp_path):
    for name, edits in patcher.FILES.items():
        src = open(os.path.join(REF_ENC, name)).readlines()
        out = patcher.apply_edits(src, edits, name)
        text = "".join(out)
        assert "torch" not in text.replace("// C ABI host mirror (fastintercu_vvc_b200/hook); no libtorch, no OpenCV", "") or name != "EncCu.cpp"
        assert "opencv2" not in text
        removed = sum(max(l, f - 1) - f + 1 for f, l, *_ in edits)
        added = sum(len(r) for *_, r in edits)
        assert len(out) == len(src) - removed + added
    src = open(os.path.join(REF_ENC, "EncCu.cpp")).readlines()
    for dev in ("cpu", "cuda"):
        out = patcher.apply_reference_edits(src, dev)
        assert len(out) == len(src) + 10  # the trace line and the input dump (oracle build only); everything else token edits
        changed = [i for i, (a, b) in enumerate(zip(src[:926], out[:926])) if a != b]
        assert changed == ([803, 898] if dev == "cpu" else [898])  # 0-based: EncCu.cpp:804 (device), :899 (model directory)
        blk = patcher.hook_block(src, dev)
        assert len(blk) == 926 - 803 + 1


def test_synth_clip_is_deterministic_and_has_the_yuv420_layout(tmp_path):
    p8, p8b, p10 = (str(tmp_path / n) for n in ("a.yuv", "b.yuv", "c.yuv"))
    vtm_run.synth_clip(p8, 64, 48, 3, 8)
    vtm_run.synth_clip(p8b, 64, 48, 3, 8)
    vtm_run.synth_clip(p10, 64, 48, 3, 10)
    assert os.path.getsize(p8) == 3 * 64 * 48 * 3 // 2 and os.path.getsize(p10) == 2 * os.path.getsize(p8)
    assert open(p8, "rb").read() == open(p8b, "rb").read()
    y8 = np.frombuffer(open(p8, "rb").read(), np.uint8)[: 64 * 48].astype(np.int32)
    y10 = np.frombuffer(open(p10, "rb").read(), "<u2")[: 64 * 48].astype(np.int32)
    assert y10.max() <= 1023 and np.abs(y10 - 4 * y8).max() <= 4  # the 10-bit clip is the same picture at 4x scale
    u = np.frombuffer(open(p8, "rb").read(), np.uint8)[64 * 48 : 64 * 48 + 32 * 24]
    assert np.all(u == 128)


def test_bd_rate_known_answers():
    r = np.array([1000.0, 2000.0, 4000.0, 8000.0])
    p = np.array([32.0, 35.0, 38.0, 41.0])
    assert abs(vtm_run.bd_rate(r, p, r, p)) < 1e-9
    assert abs(vtm_run.bd_rate(r, p, 1.05 * r, p) - 5.0) < 1e-6  # 5 % more rate at equal PSNR
    assert vtm_run.bd_rate(r, p, r, p + 0.5) < 0  # better quality at equal rate = negative BD-rate
    assert vtm_run.bd_rate(r[:3], p[:3], r[:3], p[:3]) is None


def test_compare_reports_decisions_bitstreams_and_time():
    def res(enc, md5, trace, t):
        return {"encoder": enc, "clip": "c", "qp": 32, "rc": 0, "bitstream_md5": md5, "recon_md5": md5, "decoded_md5": md5,
                "decode_matches_recon": True, "trace": trace, "total_time_elapsed_s": t, "bitrate_kbps": 100.0, "psnr_y": 30.0}
    tr = [[1, 0, 0, 33, 1], [1, 128, 0, 33, 3]]
    c = vtm_run.compare([res("ref_cpu", "a", tr, 30.0), res("mlt", "a", tr, 20.0), res("anchor", "b", [[1, 0, 0, 33, -1]], 40.0)])
    assert c["decode_ok"] and c["pairs"][0]["all_decisions_equal"] and c["pairs"][0]["bitstream_equal"]
    ets = {t["encoder"]: t["ets_pct"] for t in c["time"]}
    assert ets == {"mlt": 50.0, "ref_cpu": 25.0}
    tr2 = [[1, 0, 0, 33, 1], [1, 128, 0, 33, 2]]
    c = vtm_run.compare([res("ref_cpu", "a", tr, 30.0), res("mlt", "z", tr2, 20.0)])
    assert not c["pairs"][0]["all_decisions_equal"] and c["pairs"][0]["decisions_equal"] == 1 and not c["pairs"][0]["bitstream_equal"]


@pytest.mark.skipif(not os.path.exists(REF_TU), reason="oracle/_ref not built (needs /root/reference; python __graft_entry__.py)")
def test_reference_hook_statements_agree_with_the_c_oracle(tmp_path):
    """The reference's own lines EncCu.cpp:803-926 (compiled verbatim into oracle/_ref/ref_hook_tu_cpu) and the C restatement
    (oracle/mltcnn_oracle.c) must take the same split decision on the same CTUs, under the same seeded weights."""
    import torch

    from oracle import ref_arch
    from tests.oracle_lib import OracleModel

    sd = ref_arch.make_state_dict(10)
    n = 12
    orgpred, pocqp = ref_arch.synth_ctus(n, 10)
    torch.manual_seed(0)
    traced = torch.jit.trace(ref_arch.build_model(sd), (torch.rand(1, 2, 128, 128), torch.rand(1), torch.rand(1)))
    traced.save(str(tmp_path / "MLTORPQ_splitMode_128.pt"))
    with open(tmp_path / "ctus.bin", "wb") as f:
        f.write(np.int32(n).tobytes())
        for i in range(n):
            f.write(pocqp[i].astype(np.int32).tobytes() + np.ascontiguousarray(orgpred[i, 0]).tobytes() + np.ascontiguousarray(orgpred[i, 1]).tobytes())
    env = {k: v for k, v in os.environ.items() if k != "MLT_REF_LOAD_PER_CALL"}
    env["MLT_REF_MODEL_DIR"] = str(tmp_path)
    r = subprocess.run([REF_TU, str(tmp_path / "ctus.bin"), "0.1", "4"], capture_output=True, text=True, env=env, timeout=300, check=True)
    got = [int(ln.split()[2]) for ln in r.stdout.splitlines() if ln.startswith("split")][:n]
    logits, split = OracleModel(sd).predict_batch(orgpred, pocqp)
    l3 = np.sort(logits[:, 5:9], 1)
    clear = (l3[:, -1] - l3[:, -2]) > 1e-3  # away from fp32 ties the two CPU paths must agree exactly
    assert len(got) == n and all(g == int(s) for g, s, c in zip(got, split, clear) if c)
    assert clear.sum() >= n - 2
