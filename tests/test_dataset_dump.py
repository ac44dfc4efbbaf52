"""Dataset dump hook: the PNG / CSV files written by the C++ hook are exactly what the
reference's loader reads (mlt-cnn-python/codes/data/mlt_ctu_or_pq_dataset.py:13-15,46-69).  CPU only."""
import csv
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_arch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOOK = os.path.join(ROOT, "fastintercu_vvc_b200", "hook")
REF_CODES = "/root/reference/mlt-cnn-python/codes"


def _dump(tmp_path, size, n, seq="BasketballDrive", base_qp=32):
    subprocess.check_call(["make", "-s", "-C", HOOK])
    if size == 128:
        orgpred, pocqp = ref_arch.synth_ctus(n, 3)
    else:
        orgpred, pocqp = ref_arch.synth_cus(n, size, 3)
    orgpred = orgpred.copy()
    orgpred[0, 0, 0, :4] = [-1, -300, 1023, 0]  # negative Pel: the hook's (uint16_t) cast wraps them (EncCu.cpp:816)
    rng = np.random.RandomState(1)
    meta = np.stack([pocqp[:, 0], rng.randint(0, 30, n) * size, rng.randint(0, 16, n) * size, rng.randint(0, 4, n), pocqp[:, 1]], 1).astype(np.int32)
    path = tmp_path / "in.bin"
    with open(path, "wb") as f:
        f.write(np.array([n, size], np.int32).tobytes())
        for i in range(n):
            f.write(meta[i].tobytes())
            f.write(orgpred[i, 0].tobytes())
            f.write(orgpred[i, 1].tobytes())
    root, csv_path = tmp_path / "data", tmp_path / "train.csv"
    r = subprocess.run([os.path.join(HOOK, "dump_tool.bin"), str(path), str(root), seq, str(base_qp), str(csv_path)], capture_output=True, text=True)
    assert r.returncode == 0 and f"dumped {n}" in r.stdout, r.stdout + r.stderr
    return orgpred, meta, root, csv_path, seq, base_qp


@pytest.mark.parametrize("size", [128, 64, 16])
def test_dump_files_decode_to_the_blocks_the_hook_saw(tmp_path, size):
    from PIL import Image

    n = 5
    orgpred, meta, root, csv_path, seq, base_qp = _dump(tmp_path, size, n)
    rows = list(csv.reader(open(csv_path)))
    assert len(rows) == n
    for i, row in enumerate(rows):
        assert row == [seq, str(base_qp), str(meta[i, 0]), str(meta[i, 1]), str(meta[i, 2]), str(meta[i, 3]), str(meta[i, 4])]
        name = f"{row[1]}_{row[2]}_{row[3]}_{row[4]}.png"  # dataset.py:52
        for plane, sub in ((0, "org"), (1, "pred")):
            img = np.asarray(Image.open(os.path.join(root, seq, sub, name)))
            assert img.dtype == np.uint16 and img.shape == (size, size)
            assert np.array_equal(img, orgpred[i, plane].astype(np.uint16))
    import cv2

    img = cv2.imread(os.path.join(root, seq, "org", f"{base_qp}_{meta[0, 0]}_{meta[0, 1]}_{meta[0, 2]}.png"), cv2.IMREAD_UNCHANGED)
    assert img.dtype == np.uint16 and np.array_equal(img, orgpred[0, 0].astype(np.uint16))


@pytest.mark.skipif(not os.path.exists(REF_CODES), reason="reference mount absent (GPU box)")
def test_reference_dataset_class_reads_the_dump(tmp_path):
    """The reference's own MltCtuORPQDataset, imported by path, loads the dump; its (org, resi) differ from the C++ hook's
    staging only where the training code deviates from the encoder (saturating subtract)."""
    import importlib.util

    n = 4
    orgpred, meta, root, csv_path, seq, base_qp = _dump(tmp_path, 128, n)
    spec = importlib.util.spec_from_file_location("ref_dataset", os.path.join(REF_CODES, "data", "mlt_ctu_or_pq_dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ds = mod.MltCtuORPQDataset({"dataroot_org": str(root), "dataroot_pred": str(root), "data_csv": str(csv_path)})
    assert len(ds) == n
    for i in range(1, n):  # (item 0 holds the wrapped negative samples, outside the 10-bit range the loader assumes)
        item = ds[i]
        o = orgpred[i, 0].astype(np.float64)
        p = orgpred[i, 1].astype(np.float64)
        assert item["poc"] == meta[i, 0] and item["qp"] == meta[i, 4] and item["l3_gt"] == meta[i, 3]
        assert np.allclose(item["org"].numpy()[0], (o / 1023).astype(np.float32))
        assert np.allclose(item["resi"].numpy()[0], (np.maximum(o - p, 0) / 1023).astype(np.float32))  # saturating subtract (dataset.py:58)
        staged = ref_arch.stage_numpy(orgpred[i : i + 1])[0]
        assert np.abs(item["org"].numpy()[0] - staged[0]).max() < 1e-6  # 1-ulp: float64 /1023 vs fp32 * (float)(1/1023)
