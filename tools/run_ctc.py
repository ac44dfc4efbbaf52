#!/usr/bin/env python
"""CTC-style sweep launcher: independent encodes (clip x QP) partitioned over the GPUs
of one box by `fastintercu_vvc_b200.shard.assign_encodes` (longest-processing-time first), ONE encoder process per GPU at a
time with its own mlt_ctx, device chosen by CUDA_VISIBLE_DEVICES -- how the reference's authors ran their sweeps
(vtm-mlt-cpp/script_128/archive: `CUDA_VISIBLE_DEVICES=1 ./X_enc.sh`).  No collective, no NCCL: results are gathered from the
per-encode logs.  Reports the sweep's wall time at each GPU count, the per-GPU assignment and that every bitstream is identical
across GPU counts (the partition must not change results).

  python tools/run_ctc.py --gpus 1,2,4,8 --sizes 416x240 --clips 2 --frames 3 --qps 22,27,32,37 --out profiles/r02/ctc_sweep.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import vtm_run  # noqa: E402
from fastintercu_vvc_b200.shard import assign_encodes  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2", help="GPU counts to run the whole sweep at, e.g. 1,2,4,8")
    ap.add_argument("--sizes", default="416x240")
    ap.add_argument("--clips", type=int, default=2, help="distinct synthetic clips (seeds) per size")
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--qps", default="22,27,32,37")
    ap.add_argument("--encoder", default="mlt")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    work = tempfile.mkdtemp(prefix="ctc_")
    model_dir, blob = vtm_run.make_weights(work)
    level = {416: "2.1", 832: "3.1", 1280: "4", 1920: "4.1", 3840: "5.1"}
    jobs = []
    for s in a.sizes.split(","):
        w, h = map(int, s.split("x"))
        for k in range(a.clips):
            c = {"name": f"synth{k}_{w}x{h}", "w": w, "h": h, "bits": a.bits, "frames": a.frames, "path": os.path.join(work, f"s{k}_{w}x{h}.yuv")}
            vtm_run.synth_clip(c["path"], w, h, a.frames, a.bits, seed=10 + k)
            for q in map(int, a.qps.split(",")):
                # cost model for the partition: samples x frames, lower QP = more RDO work (~15 % per 5 QP steps)
                jobs.append({"clip": c, "qp": q, "cost": w * h * a.frames * (1.0 + 0.03 * (37 - q))})
    runs = []
    for n in map(int, a.gpus.split(",")):
        plan = assign_encodes([j["cost"] for j in jobs], n)
        out: list = [None] * len(jobs)

        def worker(rank: int, mine: list[int]):
            for i in mine:  # one encoder process at a time on this GPU
                j = jobs[i]
                out[i] = vtm_run.run_encode(a.encoder, j["clip"], j["qp"], os.path.join(work, f"n{n}"), model_dir, blob, rank, level.get(j["clip"]["w"], "5.1"))

        os.makedirs(os.path.join(work, f"n{n}"), exist_ok=True)
        t0 = time.time()
        th = [threading.Thread(target=worker, args=(r, plan[r])) for r in range(n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        wall = time.time() - t0
        runs.append({"gpus": n, "wall_s": round(wall, 2), "assignment": plan,
                     "per_gpu_busy_s": [round(sum(out[i]["wall_s"] for i in plan[r]), 2) for r in range(n)],
                     "encodes": [{k: o.get(k) for k in ("clip", "qp", "device", "rc", "wall_s", "total_time_elapsed_s", "bitrate_kbps", "psnr_y",
                                                        "bitstream_md5", "decode_matches_recon", "hook_calls", "hook_ms_total")} for o in out]})
        print(f"{n} GPU(s): sweep wall {wall:.1f}s, busy per GPU {runs[-1]['per_gpu_busy_s']}", flush=True)
    base = [e["bitstream_md5"] for e in runs[0]["encodes"]]
    doc = {"host_cores": os.cpu_count(), "encoder": a.encoder, "n_encodes": len(jobs),
           "jobs": [{"clip": j["clip"]["name"], "qp": j["qp"], "cost": j["cost"]} for j in jobs],
           "runs": runs,
           "bitstreams_identical_across_gpu_counts": all([e["bitstream_md5"] for e in r["encodes"]] == base for r in runs),
           "all_ok": all(e["rc"] == 0 and e["decode_matches_recon"] for r in runs for e in r["encodes"]),
           "speedup_vs_first": [round(runs[0]["wall_s"] / r["wall_s"], 3) for r in runs],
           "note": "encodes are CPU-bound (RDO); a GPU serves its encoder process's predictor calls (0.2 ms each). One encoder process per GPU at a time."}
    s = json.dumps(doc, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(s + "\n")
    print(json.dumps({k: doc[k] for k in ("n_encodes", "bitstreams_identical_across_gpu_counts", "all_ok", "speedup_vs_first")}))


if __name__ == "__main__":
    main()
