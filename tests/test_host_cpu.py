"""CPU-only tests: weight packer, C-ABI surface, host-side hook mirror, multi-GPU sharding (gloo, world_size 2)."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from fastintercu_vvc_b200 import pack_weights as pw
from fastintercu_vvc_b200 import shard
from fastintercu_vvc_b200.synth import make_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------- packer


@pytest.fixture(scope="module")
def sd():
    return make_state_dict(10)


def test_bn_folding_matches_torch_eval(sd):
    """w*gamma/sqrt(var+eps), beta-mean*gamma/sqrt(var+eps) == conv followed by eval-mode BatchNorm2d."""
    import torch
    import torch.nn.functional as F

    w, b = pw.fold_bn(sd["layer1.0.conv1.weight"], sd, "layer1.0.bn1")
    x = torch.randn(2, 32, 16, 16, generator=torch.Generator().manual_seed(0))
    y = F.conv2d(x, torch.from_numpy(sd["layer1.0.conv1.weight"]), stride=2, padding=1)
    y = F.batch_norm(y, torch.from_numpy(sd["layer1.0.bn1.running_mean"]), torch.from_numpy(sd["layer1.0.bn1.running_var"]),
                     torch.from_numpy(sd["layer1.0.bn1.weight"]), torch.from_numpy(sd["layer1.0.bn1.bias"]), False, 0.0, 1e-5)
    z = F.conv2d(x, torch.from_numpy(w), torch.from_numpy(b), stride=2, padding=1)
    assert (y - z).abs().max() < 2e-5


def test_blob_roundtrip_and_operand_layout(sd):
    with tempfile.NamedTemporaryFile(suffix=".mltw") as f:
        n = pw.write_blob(sd, f.name)
        assert n == os.path.getsize(f.name)
        secs = pw.read_sections(f.name)
    table = pw.conv_table()
    assert len(table) == 16
    for li, (prefix, cin, cout, stride, hout, group, sc) in enumerate(table):
        wf, bf = pw.fold_bn(sd[f"{prefix}.weight"], sd, prefix.replace("conv", "bn"))
        packed = secs[pw.SEC_W_F16 + li].reshape(cin // group, 9, group // 8, cout, 8)
        # element (cg, tap, j, n, e) == folded weight [n][cg*G + j*8 + e][kh][kw]
        for (cg, t, j, nn, e) in ((0, 0, 0, 0, 0), (cin // group - 1, 8, group // 8 - 1, cout - 1, 7), (0, 4, 1, 5, 3)):
            ci = cg * group + j * 8 + e
            want = wf[nn, ci, t // 3, t % 3]
            ulp = float(np.spacing(np.float16(np.abs(wf[nn, ci]).max())))  # ulp of the largest tap of this (cout, cin) kernel
            assert abs(float(packed[cg, t, j, nn, e]) - float(want)) <= ulp * 1.001
        # error-diffused rounding: per (cout, cin) the nine rounding errors cancel to within half an ulp of the last tap
        dense = packed.transpose(3, 0, 2, 4, 1).reshape(cout, cin, 9).astype(np.float64)  # [n][cg][j][e][t] -> [cout][cin][tap]
        resid = np.abs((dense - wf.reshape(cout, cin, 9).astype(np.float64)).sum(2))
        ulp_last = np.spacing(np.abs(dense[:, :, 8]).astype(np.float16)).astype(np.float64)
        assert np.all(resid <= 0.5001 * np.maximum(ulp_last, float(np.spacing(np.float16(2.0 ** -14)))))
        rn = np.abs((wf.astype(np.float16).astype(np.float64) - wf.astype(np.float64)).reshape(cout, cin, 9).sum(2))
        assert resid.mean() < 0.6 * rn.mean()  # independent round-to-nearest leaves ~3x more
        assert np.array_equal(secs[pw.SEC_W_F32 + li].reshape(9, cin, cout)[4, 1, 2], wf[2, 1, 1, 1])
        bias_op = secs[pw.SEC_BIAS_MMA + li].reshape(2, cout, 8).astype(np.float32)
        fused = secs[pw.SEC_BIAS_FUSED + li]
        assert np.abs(bias_op[0, :, 0] + bias_op[0, :, 1] - fused).max() < 1e-6  # hi + lo split is ~fp32 exact
        assert not bias_op[1].any() and not bias_op[0, :, 2:].any()
        if li & 1:  # second conv of a block: extra K-slab operand = folded shortcut weights, or the identity
            xc = table[li - 1][1] if sc >= 0 else cout
            gx = min(xc, group)
            if sc >= 0:  # folded 1x1 shortcut weights as an fp16 hi + lo pair
                xop = secs[pw.SEC_X_W_F16 + li].reshape(xc // gx, 2, gx // 8, cout, 8).astype(np.float64)
                dense = xop.sum(1).transpose(2, 0, 1, 3).reshape(cout, xc)  # hi + lo, [cout][xc]
                sp = prefix.rsplit(".", 1)[0] + ".shortcut"
                ws, _ = pw.fold_bn(sd[f"{sp}.0.weight"], sd, f"{sp}.1")
                assert np.abs(dense - ws.reshape(cout, xc)).max() <= np.abs(ws).max() * 2.0 ** -20
            else:
                xop = secs[pw.SEC_X_W_F16 + li].reshape(xc // gx, gx // 8, cout, 8).astype(np.float32)
                dense = xop.transpose(2, 0, 1, 3).reshape(cout, xc)  # [cout][xc]
                assert np.array_equal(dense, np.eye(cout, dtype=np.float32))
        else:
            assert (pw.SEC_X_W_F16 + li) not in secs
    # conv1 as a tcgen05 operand: hi + lo reproduces w * (float)(1/1023) * 2^10; 4th pixel and 4th chunk are zero
    c1 = secs[pw.SEC_CONV1_UMMA].reshape(2, 4, 32, 8).astype(np.float64)
    w1 = sd["conv1.weight"].astype(np.float64) * float(np.float32(1.0 / 1023)) * 1024.0
    for kh in range(3):
        for dx in range(3):
            for ch in range(2):
                got = c1[0, kh, :, dx * 2 + ch] + c1[1, kh, :, dx * 2 + ch]
                assert np.abs(got - w1[:, ch, kh, dx]).max() <= np.abs(w1).max() * 2.0 ** -20
    assert not c1[:, 3].any() and not c1[:, :, :, 6:].any()
    # 'params' wrapper and 'module.' prefixes (model2torchScript.py:23-32) are accepted
    wrapped = {"params": {"module." + k: v for k, v in sd.items()}}
    assert pw.pack(wrapped) == pw.pack(sd)
    assert secs[pw.SEC_CONV1_F32].reshape(9, 2, 32)[5, 1, 7] == sd["conv1.weight"][7, 1, 1, 2]


# ------------------------------------------------------------------------------------------- C ABI surface


def test_library_exports_every_declared_symbol():
    from fastintercu_vvc_b200 import capi

    hdr = open(os.path.join(ROOT, "include", "mltcnn.h")).read() + open(os.path.join(ROOT, "include", "mltcnn_cu.h")).read()
    declared = set(re.findall(r"MLT_API[^;(]*?\b(mlt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18 + 11, declared
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported"
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", capi.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (mlt_[a-z0-9_]+)", out))
    assert exported == declared, exported ^ declared  # nothing else leaks out of the library
    assert lib.mlt_abi_version() == 1
    assert b"batch" in lib.mlt_strerror(-7)
    import ctypes as C

    assert C.sizeof(capi.MltResult) == 88 and C.sizeof(capi.CtuDesc) == 32


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fastintercu_vvc_b200 import MltError, MltPredictor

    with pytest.raises(MltError) as e:
        MltPredictor("/nonexistent.mltw")
    assert e.value.rc == -6  # MLT_E_NODEVICE: fails loudly, never computes on the CPU


def test_no_cpu_fallback_for_the_cu_models_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fastintercu_vvc_b200 import MltCuPredictor, MltError

    with pytest.raises(MltError) as e:
        MltCuPredictor("/nonexistent.mltw", 64)
    assert e.value.rc == -6
    with pytest.raises(MltError) as e:
        MltCuPredictor("/nonexistent.mltw", 48)  # only 64 / 32 / 16 exist (EncCu.cpp:754)
    assert e.value.rc == -1


def test_cu_packer_layout_agrees_with_the_kernel_tables(tmp_path):
    """The per-size operand layouts (weight-slab channel groups) the packer writes are the ones the tcgen05 kernels
    were instantiated with (csrc/cu_net.cuh), read back through mlt_cu_layer_info -- no GPU needed."""
    from fastintercu_vvc_b200 import capi, synth

    lib = capi.load_library()
    planes = (32, 64, 96, 128, 256)
    for size in (64, 32, 16):
        table = pw.cu_conv_table(size)
        assert len(table) == 20
        flops = 2 * size * size * 32 * 18
        for li, (prefix, cin, cout, stride, hout, g, xc, gx, kind) in enumerate(table):
            info = np.zeros(10, np.int32)
            assert lib.mlt_cu_layer_info(size, li, info.ctypes.data) == 0
            assert (cin, cout, stride, hout, xc) == tuple(info[:5]), (size, li)
            assert (g, gx) == tuple(info[8:10]), (size, li)
            assert cout == planes[li // 4] and hout == max(size >> (li // 4 + 1), 1)
            assert cin % g == 0 and (xc == 0 or xc % gx == 0)
            nb, flat = int(info[6]), int(info[7])
            assert flat == (hout <= 4)
            if flat:  # NB images x HOUT rows x (HOUT + halo) positions fit the 128 accumulator rows
                halo = 0 if (hout == 1 and stride == 1) else 1  # a 3x3 conv on a 1x1 map is its centre tap: no halo, 128 images per tile
                assert nb * hout * (hout + (2 * halo if stride == 1 else 1)) <= 128
                assert nb == 128 // (hout * (hout + (2 * halo if stride == 1 else 1)))
        assert lib.mlt_cu_layer_info(size, 20, np.zeros(10, np.int32).ctypes.data) == -1
        path = str(tmp_path / f"cu{size}.mltw")
        sd = synth.make_cu_state_dict(10, size)
        pw.write_cu_blob(sd, size, path)
        secs = pw.read_sections(path, size)
        with pytest.raises(ValueError):
            pw.read_sections(path)  # not a CTU blob
        # spot-check one packed element of a streamed 96-channel layer: conv 9 = layer2.0.conv2 (96 -> 96)
        prefix, cin, cout, stride, hout, g, xc, gx, kind = table[9]
        wf, _ = pw.fold_bn(sd[f"{prefix}.weight"], sd, prefix.replace("conv", "bn"))
        packed = secs[pw.SEC_W_F16 + 9].reshape(cin // g, 9, g // 8, cout, 8)
        co, ci, kh, kw = 77, 50, 2, 1
        assert abs(float(packed[ci // g, kh * 3 + kw, (ci % g) // 8, co, ci % 8]) - float(wf[co, ci, kh, kw])) <= 2e-3 * abs(float(wf[co, ci, kh, kw])) + 1e-6
        assert secs[pw.SEC_X_W_F16 + 9].size == 2 * xc * cout  # shortcut weights as hi + lo
    assert lib.mlt_cu_layer_info(48, 0, np.zeros(10, np.int32).ctypes.data) == -1


def test_product_code_never_touches_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "fastintercu_vvc_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(base, fn)).read()
                for pat in ("import oracle", "from oracle", "oracle/", "oracle.", "mlto_", "libmltcnn_oracle"):
                    assert pat not in src, f"{fn} references the oracle ({pat})"


# ------------------------------------------------------------------------------------------- hook mirror (C++)


def test_hook_mirror_cpp():
    hook = os.path.join(ROOT, "fastintercu_vvc_b200", "hook")
    subprocess.check_call(["make", "-s", "-C", hook])
    r = subprocess.run([os.path.join(hook, "test_hook.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "test_hook: OK" in r.stdout
    assert "Hello" in r.stdout  # the reference's pred < 0 branch (EncModeCtrl.cpp:147-148)


# ------------------------------------------------------------------------------------------- sharding


def test_shard_range_and_lpt():
    for n in (0, 1, 7, 8, 120, 121):
        for w in (1, 2, 3, 8):
            parts = [list(shard.shard_range(n, w, r)) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    costs = [416 * 240, 832 * 480, 1280 * 720, 1920 * 1080] * 2
    for w in (1, 2, 4, 8):
        plan = shard.assign_encodes(costs, w)
        assert sorted(sum(plan, [])) == list(range(8))
        loads = [sum(costs[i] for i in p) for p in plan]
        assert max(loads) <= sum(costs) / w + max(costs)


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from fastintercu_vvc_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 37
rows = np.array([[i, (i * 2654435761) % 4] for i in shard.shard_range(n, world, rank)], np.int64)  # stand-in "decisions"
allrows = shard.gather_rows(rows, world, rank)
assert allrows.shape == (n, 2) and np.array_equal(allrows[:, 0], np.arange(n))
assert np.array_equal(allrows[:, 1], (np.arange(n) * 2654435761) % 4)
dist.barrier()
if rank == 0:
    print("GLOO_OK", world)
dist.destroy_process_group()
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29541", str(script), ROOT],
        capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK 2" in r.stdout


# ------------------------------------------------------------------------------------------- weight tooling (section 8f rank 3)


def test_packer_reads_the_reference_containers(tmp_path):
    """`.pth` with 'params' + 'module.' prefixes (model2torchScript.py:23-32) and the traced TorchScript `.pt` the hook
    loads (model2torchScript.py:46-48) pack to byte-identical blobs; the CLI takes both, for the CTU and the CU models."""
    import torch

    from fastintercu_vvc_b200 import synth
    from oracle import ref_arch

    for size in (128, 16):
        sd = make_state_dict(10) if size == 128 else synth.make_cu_state_dict(10, size)
        want = pw.pack(sd) if size == 128 else pw.pack(sd, size)
        # BasicSR-style checkpoint with DataParallel prefixes
        pth = str(tmp_path / f"net_{size}.pth")
        torch.save({"params": {"module." + k: torch.from_numpy(v) for k, v in sd.items()}}, pth)
        # TorchScript traced exactly like model2torchScript.py:37-48
        net = ref_arch.build_model(sd) if size == 128 else ref_arch.build_cu_model(sd)
        ex = (torch.cat((torch.rand(1, 1, 128, 128), torch.rand(1, 1, 128, 128)), 1), torch.rand(1), torch.rand(1))
        pt = str(tmp_path / f"MLTORPQ_splitMode_{size}.pt")
        torch.jit.trace(net, ex).save(pt)
        for src in (pth, pt):
            got = pw.load_checkpoint(src)
            assert set(got) == set(sd)
            assert all(np.array_equal(got[k], sd[k]) for k in sd)
            out = str(tmp_path / "out.mltw")
            argv = ([] if size == 128 else ["--cu", str(size)]) + [src, out]
            assert pw.main(argv) == 0
            assert open(out, "rb").read() == want
        # export in the reference's container and back
        rt = str(tmp_path / "rt.pth")
        pw.export_state_dict(sd, rt)
        assert all(np.array_equal(pw.load_checkpoint(rt)[k], sd[k]) for k in sd)
    assert pw.main(["only-one-arg"]) == 2


def test_pack10_host_packer_roundtrip_and_range_check():
    """mlt_pack10 (host side of the 10-bit packed transport): little-endian bit stream of 10-bit fields, 4 samples per 5 bytes;
    samples outside [0, 1023] are counted, not silently saturated."""
    from fastintercu_vvc_b200 import capi

    rs = np.random.RandomState(3)
    x = rs.randint(0, 1024, (3, 2, 128, 128)).astype(np.int16)
    x[0, 0, 0, :4] = (0, 1023, 512, 1)
    p = capi.pack10(x)
    assert p.shape == (3, capi.PACKED10_BYTES) and capi.PACKED10_BYTES == 2 * 128 * 128 * 10 // 8
    bits = np.unpackbits(p.reshape(-1), bitorder="little").reshape(-1, 10)
    back = (bits.astype(np.int32) << np.arange(10)).sum(1).astype(np.int16).reshape(x.shape)
    assert np.array_equal(back, x)
    assert p[0, :5].tolist() == [0x00, 0xFC, 0x0F, 0x60, 0x00]  # 0 | 1023<<10 | 512<<20 | 1<<30, known answer
    L = capi.load_library()
    out = np.empty(capi.PACKED10_BYTES, np.uint8)
    for bad_val, count in ((-1, 1), (1024, 1), (-32768, 1)):
        y = x[1].copy()
        y[1, 5, 7] = bad_val
        assert L.mlt_pack10(y.ctypes.data, y.size, out.ctypes.data) == count
        with pytest.raises(ValueError):
            capi.pack10(y[None])


def test_pack10_vector_and_scalar_paths_agree_at_every_length():
    """mlt_pack10 takes an AVX2 path for all but the last >= 24 samples (its 16-byte stores overlap) and a scalar loop for the rest:
    every length around the switch-over must give the same bit stream, count out-of-range samples in both parts, and never write
    past count * 10 / 8 bytes."""
    from fastintercu_vvc_b200 import capi

    L = capi.load_library()
    rs = np.random.RandomState(11)
    for count in [4, 8, 20, 36, 40, 44, 52, 56, 60, 64, 72, 100, 1000, 4096 + 12, 2 * 128 * 128]:
        x = rs.randint(0, 1024, count).astype(np.int16)
        nbytes = count * 10 // 8
        out = np.full(nbytes + 32, 0xAB, np.uint8)
        assert L.mlt_pack10(x.ctypes.data, count, out.ctypes.data) == 0
        bits = np.unpackbits(out[:nbytes], bitorder="little").reshape(-1, 10)
        assert np.array_equal((bits.astype(np.int32) << np.arange(10)).sum(1).astype(np.int16), x), count
        assert (out[nbytes:] == 0xAB).all(), f"wrote past the end at count {count}"
        y = x.copy()
        y[0], y[count - 1] = 1024, -1  # one in the vector part (when there is one), one in the scalar tail
        assert L.mlt_pack10(y.ctypes.data, count, out.ctypes.data) == 2


def test_stem5_composite_is_the_two_reference_convs(sd):
    """conv1 (arch.py:278, no BN / activation) followed by layer0.0.conv1 (3x3 stride 2, folded BN) == ONE 5x5 stride-2 conv of the
    input minus the 1-D border terms at output row 0 / column 0 (pack_weights.stem5_composite), to float64 rounding; and the packed
    tcgen05 operands hold exactly those weights (fp16-rounded, staging scale folded in) in the chunk layout stem5_umma.cu reads."""
    import torch
    import torch.nn.functional as F

    w1 = sd["conv1.weight"]
    w0f, b0f = pw.fold_bn(sd["layer0.0.conv1.weight"], sd, "layer0.0.bn1")
    W5, Wtop, Wleft, Wc = pw.stem5_composite(w1, w0f)
    x = torch.rand(2, 2, 32, 32, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    ref = F.conv2d(F.conv2d(x, torch.from_numpy(w1).double(), padding=1), torch.from_numpy(w0f).double(), stride=2, padding=1)
    z = F.conv2d(x, torch.from_numpy(W5), stride=2, padding=2)
    xp = F.pad(x, (2, 2, 2, 2))
    z[:, :, 0, :] -= F.conv1d(xp[:, :, 2, :], torch.from_numpy(Wtop), stride=2)
    z[:, :, :, 0] -= F.conv1d(xp[:, :, :, 2], torch.from_numpy(Wleft), stride=2)
    z[:, :, 0, 0] += x[:, :, 0, 0] @ torch.from_numpy(Wc).T
    assert (z - ref).abs().max() < 1e-12
    op, corr = pw.stem5_operands(w1, w0f)
    assert op.shape == (7, 2, 32, 8) and op.dtype == np.float16 and corr.shape == (2 * 5 * 2 * 32 + 2 * 32,)
    scale = float(pw.ALPHA) * 1024.0
    for dy, dx, ch, co in ((0, 0, 0, 0), (4, 3, 1, 31), (2, 4, 0, 7), (1, 4, 1, 19)):
        got = float(op[dy, 0, co, dx * 2 + ch]) if dx < 4 else float(op[dy, 1, co, ch])
        assert abs(got - W5[co, ch, dy, dx] * scale) <= 2.0 ** -11 * abs(W5[co, ch, dy, dx] * scale) + 1e-7
    assert not op[:5, 1, :, 2:].any() and not op[6, 1].any()  # unused K slots carry zero weights
    assert float(op[5, 1, 3, 2 * 2 + 1]) == np.float16(w1[3, 1, 2, 2] * scale)  # conv1 quarter: MMA 5 chunk 1 = kernel row 2
    assert np.allclose(corr[: 5 * 2 * 32].reshape(5, 2, 32), (Wtop * scale).transpose(2, 1, 0), rtol=1e-6)
    secs = dict(pw.build_sections(sd))
    assert secs[pw.SEC_STEM5_CORR].shape == (2 * 5 * 2 * 32 + 2 * 32 + 32,)  # ... + the stem kernel's own (bias-corrected) bias
    assert np.abs(secs[pw.SEC_STEM5_CORR][-32:] - b0f).max() < 5e-3  # the correction is a small shift of the folded bias


def test_bias_correction_is_the_mean_rounding_error():
    """bias_correction == E[sum (w_q - w) x] for inputs whose per-tap means are the calibration means (exact for constant inputs)."""
    rs = np.random.RandomState(0)
    w = rs.standard_normal((8, 4, 3, 3)).astype(np.float32) * 0.1
    q = pw.quantize_fp16_diffused(w)
    mu = rs.uniform(0, 2, (4, 3, 3))
    got = pw.bias_correction(w, q, mu)
    want = ((q.astype(np.float64) - w.astype(np.float64)) * mu[None]).sum((1, 2, 3))
    assert np.allclose(got, want, rtol=0, atol=1e-15)
    assert np.abs(got).max() < 9 * 4 * 2.0 ** -11 * 0.5 * 2  # bounded by the fp16 rounding of 36 weights times the largest mean
