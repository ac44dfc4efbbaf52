#!/usr/bin/env python
"""Run the encoders built by oracle/vtm/Makefile on synthetic clips and compare them.

Encoders (oracle/_ref/):
  ref_cpu   the reference hook's own translation unit, libtorch CPU     (EncoderApp_ref_cpu)
  ref_cuda  the same as shipped (at::kCUDA), libtorch CUDA              (EncoderApp_ref_cuda; needs a GPU)
  mlt       patched encoder, C ABI of libmltcnn.so                      (EncoderApp_mlt; needs a B200)
  anchor    EncoderApp_mlt with MLT_DISABLE=1 = stock VTM-11.0 RDO, no predictor
  prepass0 / prepass8   EncoderApp_mlt with MLT_PREPASS=1 and MLT_PREPASS_RANGE 0 / 8 (frame-level pre-pass, changes decisions)
  staged    EncoderApp_mlt with MLT_PICTURE_STAGING=1 (org plane uploaded once per picture; same decisions as mlt)

For every (clip, QP, encoder): bitstream md5, encoder-recon md5, decoder-recon md5 (oracle/_ref/DecoderApp), the
"poc x y qp split" trace of every predictor call, bitrate / PSNR from VTM's summary line (Analyze.h:249-252) and
`Total Time` (encmain.cpp:335).  Comparisons: decisions and bitstreams of `mlt` vs `ref_cpu`, encode-time saving vs `anchor`,
BD-rate (Bjontegaard, piecewise-cubic like VTM's reporting sheets use) between any two encoders over the QP set.

  python tools/vtm_run.py --config 1 --encoders ref_cpu,mlt,anchor --out profiles/r02/vtm_config1.json
  python tools/vtm_run.py --size 1920x1080 --bits 10 --frames 5 --qps 22,27,32,37 --encoders ref_cpu,mlt,anchor --jobs 12 --out ...
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFDIR = os.path.join(ROOT, "oracle", "_ref")
CFG = os.path.join(REFDIR, "cfg", "encoder_randomaccess_vtm.cfg")

ENCODERS = {
    "ref_cpu": ("EncoderApp_ref_cpu", {}),
    "ref_cuda": ("EncoderApp_ref_cuda", {}),
    "mlt": ("EncoderApp_mlt", {}),
    "staged": ("EncoderApp_mlt", {"MLT_PICTURE_STAGING": "1"}),
    "anchor": ("EncoderApp_mlt", {"MLT_DISABLE": "1"}),
    "prepass0": ("EncoderApp_mlt", {"MLT_PREPASS": "1", "MLT_PREPASS_RANGE": "0"}),
    "prepass8": ("EncoderApp_mlt", {"MLT_PREPASS": "1", "MLT_PREPASS_RANGE": "8"}),
    # the smaller-CU models inside the encoder: the reference's own cuw < 128 branch with the size gate of
    # EncCu.cpp:754 un-commented (64 / 32 / 16 px, per-size model files, level-1 argmax) against the patched encoder with MLT_CU_SIZES
    "ref_cpu_cu": ("EncoderApp_ref_cpu_cu", {}),
    "mlt_cu": ("EncoderApp_mlt", {"MLT_CU_SIZES": "64,32,16"}),
}


def md5(path: str) -> str:
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def synth_clip(path: str, w: int, h: int, frames: int, bits: int, seed: int = 10) -> None:
    """SURVEY.md section 8d config 1 / 3 generator: blurred-noise base + fine texture, global translation (2,1) px per frame,
    sigma 1.5 (8-bit scale) sensor noise, flat mid-grey chroma; planar 4:2:0, 1 B/sample (8-bit) or 2 B LE (VideoIOYuv.cpp:257-300)."""
    rng = np.random.default_rng(seed)
    pad = 2 * frames + 8
    H, W = h + pad, w + 2 * pad
    base = rng.standard_normal((H, W)).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-24, 25) / 8.0) ** 2).astype(np.float32)
    k /= k.sum()
    for ax in (0, 1):  # separable Gaussian blur, sigma 8
        base = np.apply_along_axis(lambda v: np.convolve(v, k, mode="same"), ax, base)
    base = (base - base.min()) / (base.max() - base.min())
    # a few moving-object-like blobs with hard edges so that all split classes are plausible
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    for _ in range(max(4, (w * h) // 40000)):
        cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(12, 70)
        base += rng.uniform(-0.35, 0.35) * (((yy - cy) ** 2 + (xx - cx) ** 2) < r * r)
    tex = rng.standard_normal((H, W)).astype(np.float32)
    img = 128 + 150 * (base - 0.5) + 6 * tex
    scale = 1 << (bits - 8)
    dt = np.uint8 if bits == 8 else np.dtype("<u2")
    cw, ch = w // 2, h // 2
    chroma = np.full((ch, cw), 128 * scale, dt).tobytes()
    with open(path, "wb") as f:
        for t in range(frames):
            y0, x0 = pad // 2 + t, pad + 2 * t
            fr = img[y0 : y0 + h, x0 : x0 + w] + 1.5 * rng.standard_normal((h, w)).astype(np.float32)
            f.write(np.clip(np.rint(fr * scale), 0, 255 * scale + scale - 1).astype(dt).tobytes())
            f.write(chroma)
            f.write(chroma)


def make_weights(d: str, cu_models: bool = False) -> tuple[str, str]:
    """Seeded weights (oracle.ref_arch.make_state_dict(10)) in both containers: TorchScript traced exactly as
    model2torchScript.py:37-48 for the reference hook, MLTW blob for libmltcnn.so."""
    import torch

    from fastintercu_vvc_b200.pack_weights import write_blob
    from oracle import ref_arch

    sd = ref_arch.make_state_dict(10)
    net = ref_arch.build_model(sd)
    torch.manual_seed(0)
    traced = torch.jit.trace(net, (torch.rand(1, 2, 128, 128), torch.rand(1), torch.rand(1)))
    traced.save(os.path.join(d, "MLTORPQ_splitMode_128.pt"))
    blob = os.path.join(d, "MLTORPQ_splitMode_128.mltw")
    write_blob(sd, blob)
    if cu_models:  # per-size models: MLTORPQ_splitMode_<cuw>.pt (EncCu.cpp:899) and MLT_WEIGHTS_<cuw> blobs
        from fastintercu_vvc_b200.pack_weights import write_cu_blob

        for size in (64, 32, 16):
            csd = ref_arch.make_cu_state_dict(10, size)
            torch.manual_seed(0)
            tr = torch.jit.trace(ref_arch.build_cu_model(csd), (torch.rand(1, 2, size, size), torch.rand(1), torch.rand(1)))
            tr.save(os.path.join(d, f"MLTORPQ_splitMode_{size}.pt"))
            write_cu_blob(csd, size, os.path.join(d, f"MLTORPQ_splitMode_{size}.mltw"))
    return d, blob


SUMMARY_RE = re.compile(r"^\s+(\d+)\s+a\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", re.M)
TIME_RE = re.compile(r"Total Time:\s+([\d.]+) sec\. \[user\]\s+([\d.]+) sec\. \[elapsed\]")


def run_encode(enc: str, clip: dict, qp: int, work: str, model_dir: str, blob: str, device: int | None, level: str, extra_env=None) -> dict:
    exe, env_add = ENCODERS[enc]
    tag = f"{clip['name']}_q{qp}_{enc}"
    bs, rec, dec, trace, log = (os.path.join(work, f"{tag}.{e}") for e in ("bin", "rec.yuv", "dec.yuv", "trace", "log"))
    for p in (trace,):
        if os.path.exists(p):
            os.remove(p)
    env = dict(os.environ)
    env.update({"MLT_REF_MODEL_DIR": model_dir, "MLT_WEIGHTS": blob, "MLT_TRACE": trace, "MLT_STATS": "1"})
    for size in (64, 32, 16):
        env[f"MLT_WEIGHTS_{size}"] = os.path.join(model_dir, f"MLTORPQ_splitMode_{size}.mltw")
    env.update(env_add)
    env.update(extra_env or {})
    if device is not None:
        env["CUDA_VISIBLE_DEVICES"] = str(device)
    cmd = [os.path.join(REFDIR, exe), "-c", CFG, "-i", clip["path"], "-wdt", str(clip["w"]), "-hgt", str(clip["h"]), "-fr", "30",
           "-f", str(clip["frames"]), "-q", str(qp), f"--InputBitDepth={clip['bits']}", f"--Level={level}", "-b", bs, "-o", rec]
    t0 = time.time()
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    open(log, "w").write(p.stdout + "\n--- stderr ---\n" + p.stderr)
    r = {"encoder": enc, "clip": clip["name"], "qp": qp, "rc": p.returncode, "wall_s": round(wall, 2), "device": device}
    if p.returncode != 0:
        r["error"] = (p.stderr or p.stdout)[-400:]
        return r
    m = SUMMARY_RE.search(p.stdout)
    t = TIME_RE.search(p.stdout)
    if m:
        r.update(frames=int(m.group(1)), bitrate_kbps=float(m.group(2)), psnr_y=float(m.group(3)), psnr_u=float(m.group(4)),
                 psnr_v=float(m.group(5)), psnr_yuv=float(m.group(6)))
    if t:
        r.update(total_time_user_s=float(t.group(1)), total_time_elapsed_s=float(t.group(2)))
    st = re.search(r"mlt_hook: (\d+) predictor calls, ([\d.]+) ms total", p.stderr)
    if st:
        r.update(hook_calls=int(st.group(1)), hook_ms_total=float(st.group(2)))
    r["hello_lines"] = p.stdout.count("Hello")
    r["error_lines"] = p.stderr.count("error")
    r["bitstream_bytes"] = os.path.getsize(bs)
    r["bitstream_md5"] = md5(bs)
    r["recon_md5"] = md5(rec)
    d = subprocess.run([os.path.join(REFDIR, "DecoderApp"), "-b", bs, "-o", dec, "-d", "0"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    r["decode_rc"] = d.returncode
    r["decoded_md5"] = md5(dec) if d.returncode == 0 and os.path.exists(dec) else None
    r["decode_matches_recon"] = r["decoded_md5"] == r["recon_md5"]
    r["trace"] = [list(map(int, ln.split())) for ln in open(trace)] if os.path.exists(trace) else []
    for p_ in (rec, dec):
        if os.path.exists(p_):
            os.remove(p_)
    return r


def bd_rate(r1, p1, r2, p2) -> float | None:
    """Bjontegaard delta rate (%) of curve 2 against curve 1: piecewise-cubic (PCHIP) interpolation of log-rate over PSNR,
    integrated over the common PSNR interval -- the JVET reporting-sheet method.  Needs >= 4 points with increasing PSNR."""
    from scipy.interpolate import PchipInterpolator

    a = sorted(zip(p1, np.log10(r1)))
    b = sorted(zip(p2, np.log10(r2)))
    if len(a) < 4 or len(b) < 4:
        return None
    pa, la = map(np.asarray, zip(*a))
    pb, lb = map(np.asarray, zip(*b))
    if np.any(np.diff(pa) <= 0) or np.any(np.diff(pb) <= 0):
        return None
    lo, hi = max(pa[0], pb[0]), min(pa[-1], pb[-1])
    if hi <= lo:
        return None
    ia = PchipInterpolator(pa, la).integrate(lo, hi)
    ib = PchipInterpolator(pb, lb).integrate(lo, hi)
    return float((10 ** ((ib - ia) / (hi - lo)) - 1) * 100)


def compare(results: list[dict]) -> dict:
    by = {(r["clip"], r["qp"], r["encoder"]): r for r in results if r.get("rc") == 0}
    clips = sorted({r["clip"] for r in results})
    qps = sorted({r["qp"] for r in results})
    encs = sorted({r["encoder"] for r in results})
    out = {"decode_ok": all(r.get("decode_matches_recon") for r in by.values()), "pairs": [], "bd_rate": [], "time": []}
    for c in clips:
        for q in qps:
            for ref_name, e in (("ref_cpu_cu", "mlt_cu"),):
                ref, got = by.get((c, q, ref_name)), by.get((c, q, e))
                if ref and got:
                    tr, tg = ref["trace"], got["trace"]
                    out["pairs"].append({"clip": c, "qp": q, "a": ref_name, "b": e, "calls_a": len(tr), "calls_b": len(tg),
                                         "decisions_equal": sum(1 for x, y in zip(tr, tg) if x == y), "all_decisions_equal": tr == tg,
                                         "bitstream_equal": ref["bitstream_md5"] == got["bitstream_md5"],
                                         "recon_equal": ref["recon_md5"] == got["recon_md5"]})
            ref = by.get((c, q, "ref_cpu")) or by.get((c, q, "ref_cuda"))
            for e in ("mlt", "staged", "ref_cuda"):
                got = by.get((c, q, e))
                if ref and got and got is not ref:
                    tr, tg = ref["trace"], got["trace"]
                    same = sum(1 for a, b in zip(tr, tg) if a == b)
                    out["pairs"].append({"clip": c, "qp": q, "a": ref["encoder"], "b": e, "calls_a": len(tr), "calls_b": len(tg),
                                         "decisions_equal": same, "all_decisions_equal": tr == tg,
                                         "bitstream_equal": ref["bitstream_md5"] == got["bitstream_md5"],
                                         "recon_equal": ref["recon_md5"] == got["recon_md5"]})
            anc = by.get((c, q, "anchor"))
            for e in encs:
                got = by.get((c, q, e))
                if anc and got and e != "anchor" and "total_time_elapsed_s" in got and "total_time_elapsed_s" in anc:
                    out["time"].append({"clip": c, "qp": q, "encoder": e, "total_time_s": got["total_time_elapsed_s"],
                                        "anchor_total_time_s": anc["total_time_elapsed_s"],
                                        "ets_pct": round(100 * (1 - got["total_time_elapsed_s"] / anc["total_time_elapsed_s"]), 2)})
        for base, test in (("ref_cpu", "mlt"), ("anchor", "mlt"), ("anchor", "ref_cpu"), ("mlt", "prepass0"), ("mlt", "prepass8"),
                           ("anchor", "prepass0"), ("anchor", "prepass8"), ("mlt", "staged")):
            A = [by.get((c, q, base)) for q in qps]
            B = [by.get((c, q, test)) for q in qps]
            if all(A) and all(B) and len(qps) >= 4:
                bd = bd_rate([r["bitrate_kbps"] for r in A], [r["psnr_y"] for r in A], [r["bitrate_kbps"] for r in B], [r["psnr_y"] for r in B])
                out["bd_rate"].append({"clip": c, "anchor": base, "test": test, "bd_rate_y_pct": None if bd is None else round(bd, 4),
                                       "time_ratio": round(sum(r["total_time_elapsed_s"] for r in B) / sum(r["total_time_elapsed_s"] for r in A), 4)})
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, choices=[1], help="1 = BASELINE config 1: 416x240 8-bit, 8 frames, QP 32")
    ap.add_argument("--size", action="append", default=[], help="WxH (repeatable)")
    ap.add_argument("--bits", type=int, default=10)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--qps", default="32")
    ap.add_argument("--encoders", default="ref_cpu,mlt,anchor")
    ap.add_argument("--jobs", type=int, default=1, help="encodes run side by side (timings are only comparable at equal load)")
    ap.add_argument("--gpus", type=int, default=1, help="GPUs to spread GPU encoders over (CUDA_VISIBLE_DEVICES per process)")
    ap.add_argument("--ref-threads", type=int, default=0, help="OMP_NUM_THREADS for the reference-hook encoders (0 = libtorch default = all cores)")
    ap.add_argument("--work", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--keep-traces", action="store_true")
    a = ap.parse_args()
    if a.config == 1:
        a.size, a.bits, a.frames, a.qps = ["416x240"], 8, 8, "32"
    work = a.work or tempfile.mkdtemp(prefix="vtm_run_")
    os.makedirs(work, exist_ok=True)
    encs = a.encoders.split(",")
    model_dir, blob = make_weights(work, cu_models=any(e.endswith("_cu") for e in encs))
    clips = []
    for s in a.size:
        w, h = map(int, s.split("x"))
        c = {"name": f"synth_{w}x{h}_{a.bits}b_{a.frames}f", "w": w, "h": h, "bits": a.bits, "frames": a.frames,
             "path": os.path.join(work, f"synth_{w}x{h}_{a.bits}b.yuv")}
        synth_clip(c["path"], w, h, a.frames, a.bits)
        c["md5"] = md5(c["path"])
        clips.append(c)
    qps = [int(q) for q in a.qps.split(",")]
    encs = a.encoders.split(",")
    level = {416: "2.1", 832: "3.1", 1280: "4", 1920: "4.1", 3840: "5.1"}
    jobs = [(e, c, q) for c in clips for q in sorted(qps) for e in encs]  # low QPs (the long encodes) start first
    t0 = time.time()
    gpu_rr = [0]

    def one(job):
        e, c, q = job
        dev = None
        if e not in ("ref_cpu", "ref_cpu_cu", "anchor"):
            dev = gpu_rr[0] % a.gpus
            gpu_rr[0] += 1
        extra = {"OMP_NUM_THREADS": str(a.ref_threads)} if (a.ref_threads and e.startswith("ref_")) else None
        return run_encode(e, c, q, work, model_dir, blob, dev, level.get(c["w"], "5.1"), extra)

    with ThreadPoolExecutor(max_workers=a.jobs) as ex:
        results = list(ex.map(one, jobs))
    doc = {"host_cores": os.cpu_count(), "jobs_side_by_side": a.jobs, "ref_threads": a.ref_threads or "all", "gpus": a.gpus, "wall_s": round(time.time() - t0, 1),
           "clips": [{k: v for k, v in c.items() if k != "path"} for c in clips], "qps": qps, "encoders": encs,
           "weights": "oracle.ref_arch.make_state_dict(10): seeded random (the trained .pt is not in the reference, .MISSING_LARGE_BLOBS)",
           "comparison": compare(results)}
    if not a.keep_traces:
        for r in results:
            tr = r.get("trace", [])
            r["trace_calls"] = len(tr)
            r["split_histogram"] = {str(k): sum(1 for t in tr if t[4] == k) for k in (-1, 0, 1, 2, 3)}
            if len(tr) > 64:
                r["trace"] = tr[:64]
    doc["results"] = results
    s = json.dumps(doc, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(s + "\n")
    brief = {k: doc["comparison"][k] for k in ("decode_ok", "pairs", "time", "bd_rate")}
    print(json.dumps(brief, indent=1))
    for r in results:
        print({k: r.get(k) for k in ("encoder", "clip", "qp", "rc", "bitrate_kbps", "psnr_y", "total_time_elapsed_s", "hook_calls", "hook_ms_total",
                                     "bitstream_md5", "decode_matches_recon", "trace_calls", "split_histogram", "error")})


if __name__ == "__main__":
    main()
