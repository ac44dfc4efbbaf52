#!/usr/bin/env python
"""bench.py -- MLT-CNN split inference throughput (CTUs/s) on 1..8 B200, beside the CPU libtorch baseline.

Metric / config: BASELINE.json -- "MLT-CNN split CTUs/s", batched split inference on all eligible CTUs of
synthetic 1920x1080 frames (CTU 128 -> 120 CTUs per frame).  One STEP = one pass
of the hot path (staging + conv stack + head) over `--frames` independent 1080p frames (default 32 frames =
3840 CTUs = 240 MiB of int16 input, larger than the 126 MB L2, so no L2 flush is needed between steps).

  value     : CTUs/s with the int16 inputs already resident in HBM (device-resident C-ABI entry point,
              CUDA events on the stream the kernels run on, max over ranks).
  e2e       : the same through the host-buffer C-ABI call (mlt_predict_batch_dense): pinned host int16 in,
              H2D + kernels + D2H of the mlt_result array inside the timed region.
  roofline  : tensor roofline of the tcgen05 kernels (stem_umma_kernel + 15 conv_umma_kernel launches per step = all
              21 convolutions): algorithmic FLOPs (1,134,559,232 per CTU) / device time measured live with CUDA
              events around each launch; peak = MEASURED_PEAKS.json bf16 sustained.
  cpu_baseline : oracle/ref_arch.py (torch CPU fp32 = the libtorch backend the reference hook calls) timed on a
              bounded sample on this box's host cores.

Multi-GPU (torchrun, one process per GPU): frames are sharded over ranks, no data-path collective
(weak scaling: `--frames` is per GPU).  `--impl reference` times the reference's CPU path (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CTUS_PER_FRAME = 120  # 1920x1080, CTU 128: 15 x 8 full CTUs (EncCu.cpp:755 excludes the partial bottom row)
FLOP_PER_CTU = 1_134_562_340  # SURVEY.md section 8a (2*MAC, all 21 convs + 3 FC)
FLOP_CONV1 = 18_874_368
FLOP_FC = 3_108
FLOP_UMMA_PER_CTU = FLOP_PER_CTU - FLOP_CONV1 - FLOP_FC  # the 16 tcgen05 conv launches
NCU_DRAM_BYTES_PER_CTU = 4_827_100  # measured: 18.536 GB per 3840-CTU step (profiles/r02/ncu_full_d_summary.csv; r01: 18.581 GB)
METRIC = "mlt_cnn_split_ctus_per_s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def in_window(self) -> int:
        return sum(1 for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[-1] <= self.t1)

    def start(self):
        # NVML polled every ~5 ms from a thread (a K-step timed region lasts ~0.1 s: nvidia-smi's 50 ms loop gives it 1-2 samples);
        # nvidia-smi -lms as the fallback when NVML is not usable
        self._halt = False
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(self.index)
            self._h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
            self._nvml = pynvml
            self.source = "nvml"
            threading.Thread(target=self._pump_nvml, daemon=True).start()
            return
        except Exception:
            self._nvml = None
        self.source = "nvidia-smi"
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump_nvml(self):
        nv, h = self._nvml, self._h
        bits = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self._halt:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                rs = get(h)
                self.rows.append([str(sm), str(mx), "0"] + ["Active" if rs & b else "Not Active" for b, _ in bits] + [time.time()])
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")] + [time.time()])

    def stop(self):
        self._halt = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if self.t0 is None or self.t1 is None or self.t0 <= r[-1] <= self.t1]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "source": getattr(self, "source", "nvidia-smi")}


def synth_frames(n_ctus: int, seed: int):
    """Cheap synthetic int16 CTUs with the distribution of oracle.ref_arch.synth_ctus (64 distinct CTUs tiled
    with per-copy noise so every CTU differs); generated without importing oracle/ (bench product arm)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:128, 0:128].astype(np.float32)
    base = np.empty((64, 2, 128, 128), np.float32)
    for i in range(64):
        f, ph = rng.uniform(0.005, 0.35, 3), rng.uniform(0, 6.28, 3)
        img = rng.uniform(64, 960) + rng.uniform(0, 300) * np.sin(f[0] * xx + ph[0]) * np.cos(f[1] * yy + ph[1])
        img += rng.standard_normal((128, 128)) * rng.uniform(0, 60)
        base[i, 0] = img
        base[i, 1] = np.roll(img, (1, 1), (0, 1)) + rng.standard_normal((128, 128)) * rng.uniform(0, 30)
    out = np.empty((n_ctus, 2, 128, 128), np.int16)
    for s in range(0, n_ctus, 64):
        k = min(64, n_ctus - s)
        noise = rng.randint(-3, 4, (k, 2, 128, 128)).astype(np.float32)
        out[s : s + k] = np.clip(np.rint(base[:k] + noise), 0, 1023).astype(np.int16)
    pocqp = np.stack([rng.randint(1, 32, n_ctus), rng.randint(22, 46, n_ctus)], 1).astype(np.int32)
    return out, pocqp


# ----------------------------------------------------------------------------------------------- CPU baseline


REF_RUNNER = os.path.join(ROOT, "oracle", "ref_libtorch.bin")  # C++ libtorch runner of the hook's call sequence (oracle/ref_libtorch.cpp)
# The reference's OWN statements for the path (EncCu.cpp:803-926 compiled verbatim into a timing harness, oracle/vtm/ref_hook_tu.cpp;
# built by oracle/vtm/Makefile where /root/reference exists and shipped to the GPU box under oracle/_ref/): preferred over the port
REF_TU = os.path.join(ROOT, "oracle", "_ref", "ref_hook_tu_cpu")
REF_TU_CUDA = os.path.join(ROOT, "oracle", "_ref", "ref_hook_tu_cuda")
HOW_TU = ("the reference's own hook statements EncCu.cpp:803-926 compiled verbatim (oracle/_ref/ref_hook_tu_cpu; edits: at::kCPU, model "
          "directory, module kept across calls), libtorch CPU fp32")
HOW_PORT = "C++ libtorch runner of the hook's call sequence (oracle/ref_libtorch.cpp, libtorch CPU fp32)"


def ref_kind() -> str:
    return "reference" if os.path.exists(REF_TU) else "port"


def _runner_inputs(n_ctus: int):
    """TorchScript file (traced like model2torchScript.py:37-48, seeded weights) + a CTU file for oracle/ref_libtorch.bin."""
    import torch

    from oracle import ref_arch

    net = ref_arch.build_model(ref_arch.make_state_dict(10))
    traced = torch.jit.trace(net, (torch.rand(1, 2, 128, 128), torch.rand(1), torch.rand(1)))
    d = tempfile.mkdtemp(prefix="mlt_ref_")
    pt, ctu_file = os.path.join(d, "MLTORPQ_splitMode_128.pt"), os.path.join(d, "ctus.bin")
    traced.save(pt)
    orgpred, pocqp = ref_arch.synth_ctus(n_ctus, 10)
    with open(ctu_file, "wb") as f:
        f.write(np.int32(n_ctus).tobytes())
        for i in range(n_ctus):
            f.write(pocqp[i].astype(np.int32).tobytes())
            f.write(np.ascontiguousarray(orgpred[i, 0]).tobytes())
            f.write(np.ascontiguousarray(orgpred[i, 1]).tobytes())
    return d, pt, ctu_file


def _run_runner(pt: str, ctu_file: str, budget_s: float, threads: int, mode: int, exe: str | None = None):
    """-> (ctus, seconds, threads) measured by the C++ runner itself (steady_clock around its per-CTU loop)."""
    import subprocess

    env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "MLT_REF_LOAD_PER_CALL")}  # torchrun exports OMP_NUM_THREADS=1
    if os.path.exists(exe or REF_TU):  # mode 1 = torch::jit::load on every call, the hook as written (EncCu.cpp:894-900)
        env["MLT_REF_MODEL_DIR"] = os.path.dirname(pt)
        if mode == 1:
            env["MLT_REF_LOAD_PER_CALL"] = "1"
        cmd = [exe or REF_TU, ctu_file, f"{budget_s:.3f}", str(threads)]
    else:
        cmd = [REF_RUNNER, pt, ctu_file, f"{budget_s:.3f}", str(threads), str(mode)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=budget_s * 4 + 180, env=env, check=True)
    n, dt, thr = r.stdout.strip().splitlines()[-1].split()
    return int(n), float(dt), int(thr)


def cpu_reference_rate(budget_s: float, threads: int | None = None):
    """Reference CPU path, one CTU per call like the hook (EncCu.cpp:806-921), model loaded once.  Preferred: the C++ libtorch
    runner oracle/ref_libtorch.bin (the hook's own call sequence: staging loops, from_blob / cat / permute, forward, argmax);
    if it was not built, the same traced TorchScript module driven from Python (same libtorch kernels).
    Returns (CTU/s, threads, CTUs, seconds, how)."""
    import shutil

    threads = threads or len(os.sched_getaffinity(0))
    if os.path.exists(REF_TU) or os.path.exists(REF_RUNNER):
        d, pt, ctu_file = _runner_inputs(16)
        try:
            n, dt, thr = _run_runner(pt, ctu_file, budget_s, threads, 0)
        finally:
            shutil.rmtree(d, ignore_errors=True)
        return n / dt, thr, n, dt, HOW_TU if os.path.exists(REF_TU) else HOW_PORT
    import torch

    from oracle import ref_arch

    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU arm)
    torch.set_num_threads(threads)
    sd = ref_arch.make_state_dict(10)
    net = ref_arch.build_model(sd)
    ex = (torch.rand(1, 2, 128, 128), torch.rand(1), torch.rand(1))
    traced = torch.jit.trace(net, ex)  # as model2torchScript.py:37-48
    orgpred, pocqp = ref_arch.synth_ctus(16, 10)
    x = torch.from_numpy(ref_arch.stage_numpy(orgpred))
    with torch.no_grad():
        for i in range(3):
            traced(x[i : i + 1], torch.tensor([int(pocqp[i, 0])]), torch.tensor([int(pocqp[i, 1])]))
        n, t0 = 0, time.perf_counter()
        while True:
            i = n % 16
            traced(x[i : i + 1], torch.tensor([int(pocqp[i, 0])]), torch.tensor([int(pocqp[i, 1])]))
            n += 1
            dt = time.perf_counter() - t0
            if (dt >= budget_s and n >= 16) or n >= 100000:
                break
    return n / dt, torch.get_num_threads(), n, dt, "traced TorchScript driven from Python (torch CPU fp32 = libtorch)"


def cpu_reference_variants(budget_s: float = 3.0):
    """The other two CPU figures SURVEY.md section 8d asks for: (ii) the hook AS WRITTEN -- torch::jit::load of the TorchScript
    file on every call (EncCu.cpp:894-900) -- and a B = 120 frame batch through the same module (not something the
    reference does; the best case for the CPU)."""
    import torch

    from oracle import ref_arch

    torch.set_num_threads(len(os.sched_getaffinity(0)))
    net = ref_arch.build_model(ref_arch.make_state_dict(10))
    traced = torch.jit.trace(net, (torch.rand(1, 2, 128, 128), torch.rand(1), torch.rand(1)))
    path = tempfile.NamedTemporaryFile(suffix=".pt", delete=False).name
    traced.save(path)
    orgpred, pocqp = ref_arch.synth_ctus(CTUS_PER_FRAME, 10)
    x = torch.from_numpy(ref_arch.stage_numpy(orgpred))
    poc, qp = torch.from_numpy(pocqp[:, 0].copy()), torch.from_numpy(pocqp[:, 1].copy())
    with torch.no_grad():
        if os.path.exists(REF_TU) or os.path.exists(REF_RUNNER):  # load-per-call mode: torch::jit::load inside the per-CTU loop
            import shutil

            d, pt, ctu_file = _runner_inputs(16)
            try:
                rn, rdt, _ = _run_runner(pt, ctu_file, budget_s, len(os.sched_getaffinity(0)), 1)
            finally:
                shutil.rmtree(d, ignore_errors=True)
            as_written = rn / rdt
        else:
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < budget_s or n < 3:
                m = torch.jit.load(path)
                m.eval()
                m(x[n % 120 : n % 120 + 1], poc[n % 120 : n % 120 + 1], qp[n % 120 : n % 120 + 1])
                n += 1
            as_written = n / (time.perf_counter() - t0)
        traced(x, poc, qp)
        k, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < budget_s or k < 2:
            traced(x, poc, qp)
            k += 1
        b120 = k * CTUS_PER_FRAME / (time.perf_counter() - t0)
    os.unlink(path)
    return as_written, b120


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warm = max(args.steps, 1), args.warmup
    per_step_budget = min(8.0, 120.0 / (steps + warm))
    rates, cores, n_tot, how = [], 0, 0, ""
    prepared = _runner_inputs(16) if (os.path.exists(REF_TU) or os.path.exists(REF_RUNNER)) else None  # TorchScript + CTU files written once
    try:
        for i in range(warm + steps):
            if prepared:
                n, dt, cores = _run_runner(prepared[1], prepared[2], per_step_budget, len(os.sched_getaffinity(0)), 0)
                how = HOW_TU if os.path.exists(REF_TU) else HOW_PORT
            else:
                _, cores, n, dt, how = cpu_reference_rate(per_step_budget)
            if i >= warm:
                rates.append((n, dt))
                n_tot += n
    finally:
        if prepared:
            import shutil

            shutil.rmtree(prepared[0], ignore_errors=True)
    tot_n = sum(n for n, _ in rates)
    tot_t = sum(t for _, t in rates)
    v = tot_n / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "CTU/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * tot_t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "1080p frames (120 CTUs/frame), CPU path: one CTU per forward as EncCu.cpp:869-921, model loaded once"},
        "cpu_baseline": {"value": v, "unit": "CTU/s", "cores": cores, "kind": ref_kind(),
                         "sample": f"{tot_n} CTUs at B=1 through {how} in {tot_t:.1f}s"},
        "e2e": {"value": v, "unit": "CTU/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- smaller-CU models


def cu_flops(size: int) -> int:
    """Algorithmic FLOPs (2 * MAC) of one `size`-px CU through GapBigMltCuORPQ (mlt_cu_or_pq_arch.py:59-130)."""
    from fastintercu_vvc_b200.pack_weights import cu_conv_table

    f = 2 * size * size * 32 * 18
    for _, cin, cout, _, hout, _, xc, _, kind in cu_conv_table(size):
        f += 2 * hout * hout * cout * cin * 9
        if kind == 1:
            f += 2 * hout * hout * cout * xc  # 1x1 stride-2 shortcut conv
    return f + 2 * (66 * 2 + 98 * 3 + 130 * 4 + 258 * 6)


def synth_cu_batch(n: int, size: int, seed: int):
    per = (128 // size) ** 2
    ctus, pq = synth_frames((n + per - 1) // per, seed)
    k = 128 // size
    cus = ctus.reshape(-1, 2, k, size, k, size).transpose(0, 2, 4, 1, 3, 5).reshape(-1, 2, size, size)[:n]
    return np.ascontiguousarray(cus), np.ascontiguousarray(np.repeat(pq, per, 0)[:n])


def cpu_cu_reference_rate(size: int, budget_s: float):
    """Reference CPU path of a smaller-CU model: traced TorchScript of `GapBigMltCuORPQ` on torch CPU fp32 (libtorch), one CU
    per forward like the hook's cuw != 128 branch (EncCu.cpp:869-921), model loaded once.  Bounded by `budget_s` seconds."""
    import torch

    from oracle import ref_arch

    torch.set_num_threads(len(os.sched_getaffinity(0)))
    net = ref_arch.build_cu_model(ref_arch.make_cu_state_dict(10, size))
    traced = torch.jit.trace(net, (torch.rand(1, 2, 128, 128), torch.rand(1), torch.rand(1)))  # as model2torchScript.py:37-48
    cus, pq = ref_arch.synth_cus(16, size, 10)
    x = torch.from_numpy(ref_arch.stage_numpy(cus))
    with torch.no_grad():
        for i in range(3):
            traced(x[i : i + 1], torch.tensor([int(pq[i, 0])]), torch.tensor([int(pq[i, 1])]))
        n, t0 = 0, time.perf_counter()
        while True:
            i = n % 16
            traced(x[i : i + 1], torch.tensor([int(pq[i, 0])]), torch.tensor([int(pq[i, 1])]))
            n += 1
            dt = time.perf_counter() - t0
            if (dt >= budget_s and n >= 16) or n >= 100000:
                break
    return n / dt, torch.get_num_threads(), n, dt


def bench_cu_models(local: int, frames: int, steps: int, warm: int, cpu_budget: float = 0.0, sizes=(64, 32, 16)):
    """Device-resident and host-buffer throughput of the 64 / 32 / 16-px CU models on all
    same-size CUs of `frames` 1080p frames per step.  Secondary numbers: the headline metric stays the CTU model."""
    import torch

    import fastintercu_vvc_b200 as pkg
    from fastintercu_vvc_b200.capi import CU_RESULT_DTYPE
    from fastintercu_vvc_b200.synth import make_cu_state_dict

    out = {}
    stream = torch.cuda.current_stream()
    for size in sizes:
        n = frames * CTUS_PER_FRAME * (128 // size) ** 2
        blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
        pkg.write_cu_blob(make_cu_state_dict(10, size), size, blob)
        pred = pkg.MltCuPredictor(blob, size, device=local, max_batch=n)
        os.unlink(blob)
        cus, pq = synth_cu_batch(n, size, 2000 + size)
        h_in, h_pq = torch.from_numpy(cus).pin_memory(), torch.from_numpy(pq).pin_memory()
        d_in, d_pq = h_in.cuda(), h_pq.cuda()
        d_out = torch.zeros(n * CU_RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        h_out = np.zeros(n, CU_RESULT_DTYPE)
        for _ in range(warm):
            pred.predict_batch_device(n, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        l0 = pred.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            pred.predict_batch_device(n, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = pred.launch_count - l0
        hin, hpq = h_in.numpy(), h_pq.numpy()
        pred.predict_batch_dense(hin, hpq, h_out)
        t0 = time.perf_counter()
        for _ in range(max(steps // 2, 1)):
            pred.predict_batch_dense(hin, hpq, h_out)
        e2e_sync_s = (time.perf_counter() - t0) / max(steps // 2, 1)
        # pipelined host API: two batches in flight, the H2D of batch k + 1 under batch k's kernels
        reps = max(steps, 4)
        pred.submit_batch_dense(hin, hpq)
        pred.submit_batch_dense(hin, hpq)  # allocates the second input slot (first use), outside the timed region
        pred.collect(h_out)
        pred.collect(h_out)
        t0 = time.perf_counter()
        pred.submit_batch_dense(hin, hpq)
        for _ in range(reps - 1):
            pred.submit_batch_dense(hin, hpq)
            pred.collect(h_out)
        pred.collect(h_out)
        e2e_s = (time.perf_counter() - t0) / reps
        o1, p1 = np.ascontiguousarray(hin[0, 0]), np.ascontiguousarray(hin[0, 1])
        for _ in range(10):
            pred.predict(o1, p1, int(hpq[0, 0]), int(hpq[0, 1]))
        t0 = time.perf_counter()
        for _ in range(100):
            pred.predict(o1, p1, int(hpq[0, 0]), int(hpq[0, 1]))
        one_us = (time.perf_counter() - t0) / 100 * 1e6
        # frame-level pre-pass at this CU size: all size x size blocks of ONE 1080p picture, planes in page-locked host memory
        prng = np.random.RandomState(size)
        p_org = torch.from_numpy(prng.randint(0, 1024, (1080, 1920)).astype(np.int16)).pin_memory().numpy()
        p_ref = torch.from_numpy(np.clip(np.roll(p_org, (1, 1), (0, 1)).astype(np.int32) + prng.randint(-8, 9, p_org.shape), 0, 1023).astype(np.int16)).pin_memory().numpy()
        n_pic = (1920 // size) * (1080 // size)
        p_mv = prng.randint(-16, 17, (n_pic, 2)).astype(np.int16)
        for _ in range(3):
            pred.predict_picture(p_org, p_ref, 3, 32, mv=p_mv)
        t0 = time.perf_counter()
        for _ in range(20):
            pred.predict_picture(p_org, p_ref, 3, 32, mv=p_mv)
        pic_ms = (time.perf_counter() - t0) / 20 * 1e3
        fl = cu_flops(size)
        out[str(size)] = {"cus_per_step": n, "ms_per_step": ms, "cus_per_s": n / (ms * 1e-3), "e2e_cus_per_s": n / e2e_s,
                          "e2e_api": "mlt_cu_submit_batch_dense + mlt_cu_collect (two batches in flight)", "e2e_sync_call_cus_per_s": n / e2e_sync_s,
                          "flop_per_cu": fl, "tflops": n * fl / (ms * 1e-3) / 1e12, "gpu_launches_per_step": launches // steps,
                          "cu_latency_us": one_us, "h2d_bytes_per_step": int(n * (4 * size * size + 8)),
                          "d2h_bytes_per_step": int(n * CU_RESULT_DTYPE.itemsize),
                          "picture_prepass": {"cus": n_pic, "ms_per_picture": pic_ms, "serial_hook_calls_ms": one_us * n_pic * 1e-3,
                                              "note": "mlt_cu_predict_picture: every block of one 1920x1080 picture in one batch (planes pinned, "
                                                      "gather on the device) vs one blocking mlt_cu_predict per CU"}}
        pred.close()
        del d_in, d_pq, d_out
        torch.cuda.empty_cache()
        if cpu_budget > 0:
            r, cores, cn, cdt = cpu_cu_reference_rate(size, cpu_budget)
            out[str(size)]["cpu_baseline"] = {"value": r, "unit": "CU/s", "cores": cores, "kind": "port",
                                              "sample": f"{cn} CUs at B=1 through traced TorchScript (torch CPU fp32 = libtorch) in {cdt:.1f}s"}
    return out


def bind_to_gpu_numa(local: int):
    """One process per GPU: pin this rank's host threads (and, by first touch, its page-locked input buffers) to the CPUs
    the driver reports as local to its GPU, the way one encoder process per GPU would be started under numactl.  Returns a
    short description for the JSON line; never fails the bench (VMs often expose no topology)."""
    if os.environ.get("MLT_BENCH_NO_BIND"):
        return "off (MLT_BENCH_NO_BIND)"
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use or use == allowed:
            return f"no-op ({len(use)} of {len(allowed)} CPUs local)"
        os.sched_setaffinity(0, use)
        return f"{len(use)} of {len(allowed)} CPUs (GPU-local)"
    except Exception as e:  # noqa: BLE001
        return f"unavailable ({type(e).__name__})"


# ----------------------------------------------------------------------------------------------- product arm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--frames", type=int, default=32, help="independent 1080p frames per step per GPU (120 CTUs each)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline work (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sustain-s", type=float, default=2.5, help="seconds of the sustained device-resident run (clocks sampled inside it)")
    ap.add_argument("--cu-frames", type=int, default=8, help="1080p frames whose 64 / 32 / 16-px CUs form one step of the CU-model lines (0 = skip)")
    ap.add_argument("--cu-only", action="store_true", help="only the smaller-CU models (tuning runs)")
    ap.add_argument("--cu-sizes", default="64,32,16", help="CU sizes of the CU-model lines (profiling runs pick one)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch

    import fastintercu_vvc_b200 as pkg
    from fastintercu_vvc_b200.capi import RESULT_DTYPE
    from fastintercu_vvc_b200.pack_weights import write_blob

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else "single process"
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps, warm = args.steps, max(args.warmup, 3)
    n = args.frames * CTUS_PER_FRAME
    if args.cu_only:
        print(json.dumps({"metric": "mlt_cnn_cu_split_cus_per_s",
                          "cu_models": bench_cu_models(local, args.cu_frames, max(steps // 5, 3), warm, sizes=tuple(int(x) for x in args.cu_sizes.split(",")))}))
        return 0

    # seeded random weights of the exact architecture (the trained .pt is not distributed with the reference)
    from fastintercu_vvc_b200.synth import make_state_dict

    blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
    write_blob(make_state_dict(10), blob)
    pred = pkg.MltPredictor(blob, device=local, max_batch=n)
    os.unlink(blob)

    orgpred, pocqp = synth_frames(n, 1000 + rank)
    h_in = torch.from_numpy(orgpred).pin_memory()
    h_pq = torch.from_numpy(pocqp).pin_memory()
    d_in = h_in.cuda(non_blocking=True)
    d_pq = h_pq.cuda(non_blocking=True)
    d_out = torch.zeros(n * RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    h_out = np.zeros(n, RESULT_DTYPE)
    stream = torch.cuda.current_stream()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def device_step():
        pred.predict_batch_device(n, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), stream.cuda_stream)

    # ---- device-resident throughput (value)
    sampler = ClockSampler(local)
    sampler.start()  # nvidia-smi takes a moment to start: launch it before the warm-up, keep only in-window samples
    for _ in range(warm):
        device_step()
    barrier()
    sampler.mark_begin()
    l0 = pred.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        device_step()
    e1.record(stream)
    barrier()
    sampler.mark_end()
    launches = pred.launch_count - l0
    dev_ms = e0.elapsed_time(e1)
    clocks_note = "sampled during the timed region"
    if sampler.in_window() < 3:
        # timed region shorter than a few sampling periods: sample the identical loop again, untimed, for ~1 s
        sampler.mark_begin()
        t_end = time.time() + 1.0
        while time.time() < t_end:
            device_step()
            torch.cuda.synchronize()
        sampler.mark_end()
        clocks_note = "timed region too short for nvidia-smi; sampled during an identical untimed loop right after it"
    clocks = sampler.stop()
    clocks["note"] = clocks_note

    # ---- sustained measurement: the identical device-resident loop for >= 2 s with nvidia-smi sampled INSIDE the window
    # (the K-step region above lasts only K x ~5.5 ms; the part is power-capped, so the long run is the honest clock)
    sus_sampler = ClockSampler(local)
    sus_sampler.start()
    for _ in range(3):
        device_step()
    barrier()
    sus_sampler.mark_begin()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ms = torch.tensor([dev_ms / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
    sus_steps = max(steps, int(np.ceil(args.sustain_s * 1e3 / float(step_ms))))  # the same count on every rank
    s0.record(stream)
    for i in range(sus_steps):
        device_step()
        if i % 64 == 63:
            torch.cuda.current_stream().synchronize()  # keep the launch queue bounded
    s1.record(stream)
    barrier()
    sus_sampler.mark_end()
    sus_ms = s0.elapsed_time(s1)
    sus_clocks = sus_sampler.stop()

    # ---- per-kernel device times (roofline), measured live with CUDA events around each launch
    pred.set_profiling(True)
    prof = np.zeros(18, np.float64)
    for _ in range(max(3, min(steps, 10))):
        device_step()
        torch.cuda.synchronize()
        prof += pred.get_profile()
    prof /= max(3, min(steps, 10))
    pred.set_profiling(False)

    # ---- end to end through the host-buffer C-ABI call (pinned host in, H2D + compute + D2H out)
    orgpred_pinned = h_in.numpy()
    pocqp_pinned = h_pq.numpy()
    for _ in range(2):
        pred.predict_batch_dense(orgpred_pinned, pocqp_pinned, h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        pred.predict_batch_dense(orgpred_pinned, pocqp_pinned, h_out)
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    # the same through the pipelined public API (mlt_submit_batch_dense / mlt_collect, two batches in flight): every
    # step's inputs still cross PCIe from pinned host memory and every step's results are read back inside the timed
    # region; the H2D of step k + 1 overlaps the kernels of step k
    for _ in range(2):  # warm-up with two batches in flight: the second input slot is allocated on its first use
        pred.submit_batch_dense(orgpred_pinned, pocqp_pinned)
        pred.submit_batch_dense(orgpred_pinned, pocqp_pinned)
        pred.collect(h_out)
        pred.collect(h_out)
    barrier()
    t0 = time.perf_counter()
    pred.submit_batch_dense(orgpred_pinned, pocqp_pinned)
    for _ in range(steps - 1):
        pred.submit_batch_dense(orgpred_pinned, pocqp_pinned)
        pred.collect(h_out)
    pred.collect(h_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # ---- the same pipelined API over the 10-bit packed transport (mlt_submit_batch_packed10): 40 KiB instead of 64 KiB per CTU cross
    # PCIe, unpacked on the device.  The producers (encoder processes) pack their own blocks with mlt_pack10; its cost is reported.
    from fastintercu_vvc_b200.capi import pack10

    h_packed = torch.empty((n, 40960), dtype=torch.uint8).pin_memory()
    packed_np = h_packed.numpy()
    t0 = time.perf_counter()
    pack10(orgpred_pinned, packed_np)
    pack_us_per_ctu = (time.perf_counter() - t0) / n * 1e6
    for _ in range(2):
        pred.submit_batch_packed10(packed_np, pocqp_pinned)
        pred.submit_batch_packed10(packed_np, pocqp_pinned)
        pred.collect(h_out)
        pred.collect(h_out)
    barrier()
    t0 = time.perf_counter()
    pred.submit_batch_packed10(packed_np, pocqp_pinned)
    for _ in range(steps - 1):
        pred.submit_batch_packed10(packed_np, pocqp_pinned)
        pred.collect(h_out)
    pred.collect(h_out)
    torch.cuda.synchronize()
    e2e_packed_s = time.perf_counter() - t0

    # ---- diagnostic: what this box's PCIe link gives the same pinned buffer (an e2e below compute speed is then explained)
    d_probe = torch.empty_like(d_in)
    h_probe = h_in
    d_probe.copy_(h_probe, non_blocking=True)
    barrier()  # every rank copies at the same time: the probe sees the host's aggregate H2D contention
    t0 = time.perf_counter()
    for _ in range(6):
        d_probe.copy_(h_probe, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbps = 6 * orgpred_pinned.nbytes / (time.perf_counter() - t0) / 1e9
    barrier()
    del d_probe

    # ---- single-frame latency (the 120 CTUs of one 1080p frame), host buffers, for the record
    f_in, f_pq, f_out = orgpred_pinned[:CTUS_PER_FRAME], pocqp_pinned[:CTUS_PER_FRAME], h_out[:CTUS_PER_FRAME]
    for _ in range(3):
        pred.predict_batch_dense(f_in, f_pq, f_out)
    t0 = time.perf_counter()
    for _ in range(20):
        pred.predict_batch_dense(f_in, f_pq, f_out)
    frame_ms = (time.perf_counter() - t0) / 20 * 1e3

    # the same frame device-resident (BASELINE config 2 as one launch sequence: 120 CTUs, inputs already in HBM)
    for _ in range(5):
        pred.predict_batch_device(CTUS_PER_FRAME, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    f0.record(stream)
    for _ in range(50):
        pred.predict_batch_device(CTUS_PER_FRAME, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
    f1.record(stream)
    torch.cuda.synchronize()
    frame_dev_ms = f0.elapsed_time(f1) / 50

    # ---- frame-level pre-pass (section 8f rank 2): one 1080p picture = org plane upload + reference plane upload + on-device
    # integer-MV gather + ONE 120-CTU batch (mlt_begin_picture + mlt_predict_picture), vs 120 serial hook calls
    prng = np.random.RandomState(7)
    pic_org = prng.randint(0, 1024, (1080, 1920)).astype(np.int16)
    pic_ref = np.clip(np.roll(pic_org, (1, 1), (0, 1)).astype(np.int32) + prng.randint(-8, 9, pic_org.shape), 0, 1023).astype(np.int16)
    pic_mv = prng.randint(-16, 17, (CTUS_PER_FRAME, 2)).astype(np.int16)
    for _ in range(3):
        pred.begin_picture(pic_org, 3)
        pred.predict_picture(pic_ref, 32, mv=pic_mv)
    t0 = time.perf_counter()
    for _ in range(20):
        pred.begin_picture(pic_org, 3)
        pred.predict_picture(pic_ref, 32, mv=pic_mv)
    prepass_ms = (time.perf_counter() - t0) / 20 * 1e3
    pred.pin_host_buffer(pic_org)  # what a patched VTM does once per Picture buffer (mlt_pin_host_buffer)
    pred.pin_host_buffer(pic_ref)
    for _ in range(3):
        pred.begin_picture(pic_org, 3)
        pred.predict_picture(pic_ref, 32, mv=pic_mv)
    t0 = time.perf_counter()
    for _ in range(20):
        pred.begin_picture(pic_org, 3)
        pred.predict_picture(pic_ref, 32, mv=pic_mv)
    prepass_pinned_ms = (time.perf_counter() - t0) / 20 * 1e3
    # the same with the MVs estimated on the device first (integer full search, +-8 samples), reference plane uploaded once
    for _ in range(3):
        pred.begin_picture(pic_org, 3)
        me_mv, _ = pred.estimate_picture_mv(pic_ref, 8)
        pred.predict_picture(None, 32, mv=me_mv)
    t0 = time.perf_counter()
    for _ in range(20):
        pred.begin_picture(pic_org, 3)
        me_mv, _ = pred.estimate_picture_mv(pic_ref, 8)
        pred.predict_picture(None, 32, mv=me_mv)
    prepass_me_ms = (time.perf_counter() - t0) / 20 * 1e3
    pred.unpin_host_buffer(pic_org)
    pred.unpin_host_buffer(pic_ref)

    # ---- single-CTU latency: the in-encoder hook call (mlt_predict_ctu: 64 KiB H2D, 18 launches, 88 B D2H, synchronous)
    o1, p1 = np.ascontiguousarray(orgpred_pinned[0, 0]), np.ascontiguousarray(orgpred_pinned[0, 1])
    for _ in range(20):
        pred.predict_ctu(o1, p1, int(pocqp_pinned[0, 0]), int(pocqp_pinned[0, 1]))
    t0 = time.perf_counter()
    for _ in range(200):
        pred.predict_ctu(o1, p1, int(pocqp_pinned[0, 0]), int(pocqp_pinned[0, 1]))
    ctu_us = (time.perf_counter() - t0) / 200 * 1e6

    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, e2e_sync_s, sus_ms, -h2d_gbps, e2e_packed_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_sync_s, sus_ms, h2d_gbps, e2e_packed_s = (float(t[0]), float(t[1]), float(t[2]), float(t[3]), -float(t[4]),
                                                                     float(t[5]))  # h2d: the slowest rank's link
    total_ctus = n * world * steps
    value = total_ctus / (dev_ms * 1e-3)
    e2e = total_ctus / e2e_s
    e2e_sync = total_ctus / e2e_sync_s
    e2e_packed = total_ctus / e2e_packed_s

    if rank == 0:
        sustained, burst, hbm, how = load_peaks()
        umma_ms = float(prof[0:17].sum())  # stem (staging + conv1 + layer0.0.conv1) + the 15 other tcgen05 convs
        ach = n * (FLOP_PER_CTU - FLOP_FC) / (umma_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "CTU/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": f"{args.frames} independent synthetic 1920x1080 frames per step per GPU, 120 CTUs of 128x128 each "
                                   f"({n} CTUs/step/GPU), MLT-CNN GapBigMltCtuORPQ, seeded random weights",
                       "frames_per_step": args.frames, "ctus_per_step_per_gpu": n,
                       "l2": f"inputs {n * 65536 / 2**20:.0f} MiB + activations > 126 MB L2, no flush needed",
                       "parallelism": f"replicas x{world} (frames sharded, no collectives)", "host_binding": numa},
            # ONE public API is the e2e value at every N: the pipelined pair over the 10-bit packed transport (every Pel of a 10-bit
            # encode is in [0, 1023]; 40 KiB per CTU cross PCIe instead of 64 KiB and are unpacked on the device); the same pair over
            # int16 buffers (the hook's own format) and the blocking int16 call are reported next to it
            "e2e": {"value": e2e_packed, "unit": "CTU/s", "h2d_bytes_per_step": int(n * (40960 + 8)), "d2h_bytes_per_step": int(n * RESULT_DTYPE.itemsize),
                    "api": "mlt_submit_batch_packed10 + mlt_collect (pinned host buffers of 10-bit packed Pel in, mlt_result out, two batches in flight)",
                    "host_pack_us_per_ctu_per_core": pack_us_per_ctu,
                    "host_pack_note": "mlt_pack10 on one host core, done by the producers outside the timed region (an encoder process spends seconds of RDO per CTU)",
                    "int16_transport_value": e2e, "int16_transport_api": "mlt_submit_batch_dense + mlt_collect (pinned host int16 in, 64 KiB per CTU)",
                    "sync_call_value": e2e_sync, "sync_call_api": "mlt_predict_batch_dense (one blocking call per step)",
                    "h2d_link_gbps": h2d_gbps, "h2d_bound_ctus_per_s": h2d_gbps * 1e9 / (40960 + 8) * world,
                    "h2d_link_note": "slowest rank's link with every rank copying at once (barrier in front)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "sustained": {"seconds": sus_ms * 1e-3, "steps": sus_steps, "value": n * world * sus_steps / (sus_ms * 1e-3), "unit": "CTU/s",
                          "ms_per_step": sus_ms / sus_steps, "clocks": sus_clocks,
                          "tflops_whole_net": n * sus_steps / (sus_ms * 1e-3) * FLOP_PER_CTU / 1e12,
                          "frac_of_sustained_peak": n * sus_steps / (sus_ms * 1e-3) * FLOP_PER_CTU / 1e12 / sustained,
                          "frac_of_burst_peak": n * sus_steps / (sus_ms * 1e-3) * FLOP_PER_CTU / 1e12 / burst,
                          "note": "the same device-resident step repeated back to back for >= --sustain-s seconds; nvidia-smi sampled inside the window"},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": sustained, "unit": "TFLOP/s", "frac": ach / sustained,
                         "peak_burst": burst, "frac_vs_burst": ach / burst,
                         "traffic": NCU_DRAM_BYTES_PER_CTU * n, "traffic_note": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one 3840-CTU step (profiles/r02/ncu_full_d_summary.csv), scaled to this step; algorithmic minimum is 65,536 B/CTU -- the rest is inter-layer fp16 activations",
                         "kernel": "stem5_umma_kernel + conv_umma_kernel x15 (every tcgen05 launch of a step: all 21 convs)",
                         "frac_note": "achieved = algorithmic FLOPs / sum of the per-launch CUDA-event times, launches serialised on one stream; the timed step "
                                      "overlaps two half-batches on two streams and is faster: whole_step_* = value x FLOP/CTU",
                         "whole_step_achieved": value / world * (FLOP_PER_CTU - FLOP_FC) / 1e12,
                         "whole_step_frac": value / world * (FLOP_PER_CTU - FLOP_FC) / 1e12 / sustained,
                         "whole_step_frac_vs_burst": value / world * (FLOP_PER_CTU - FLOP_FC) / 1e12 / burst,
                         "peak_source": f"{how} bf16 sustained", "kernel_ms_per_step": umma_ms, "stem_ms": float(prof[0]),
                         "head_ms": float(prof[17]), "per_layer_ms": [round(float(x), 4) for x in prof[2:17]]},
            "frame_latency_ms": frame_ms,
            "single_frame": {"ctus": CTUS_PER_FRAME, "host_buffers_ms": frame_ms, "host_buffers_ctus_per_s": CTUS_PER_FRAME / (frame_ms * 1e-3),
                             "device_resident_ms": frame_dev_ms, "device_resident_ctus_per_s": CTUS_PER_FRAME / (frame_dev_ms * 1e-3),
                             "note": "BASELINE config 2 taken literally: ONE 1080p frame (120 CTUs) per call, back to back"},
            "ctu_latency_us": ctu_us,
            "picture_prepass": {"ctus": CTUS_PER_FRAME, "ms_per_picture": prepass_pinned_ms, "ctus_per_s": CTUS_PER_FRAME / (prepass_pinned_ms * 1e-3),
                                "ms_per_picture_pageable": prepass_ms, "ms_per_picture_with_block_matching_r8": prepass_me_ms,
                                "h2d_bytes_per_picture": int(2 * 1920 * 1080 * 2 + CTUS_PER_FRAME * (16 + 32)),
                                "serial_hook_calls_ms": ctu_us * CTUS_PER_FRAME * 1e-3,
                                "note": "section 8f rank 2: mlt_begin_picture + mlt_predict_picture on one 1920x1080 picture, host planes page-locked once "
                                        "with mlt_pin_host_buffer (org + reference luma, per-CTU integer MVs; _pageable = without the pin), vs 120 blocking mlt_predict_ctu calls"},
            "tflops_whole_net": value / world * FLOP_PER_CTU / 1e12,
        }
        if args.cu_frames > 0 and world == 1:
            pred.close()
            line["cu_models"] = bench_cu_models(local, args.cu_frames, max(steps // 5, 3), warm,
                                                0.0 if args.no_cpu_baseline else min(3.0, args.cpu_budget / 4))
            line["cu_models"]["note"] = (f"secondary (SURVEY.md section 8f rank 1): 64 / 32 / 16-px GapBigMltCuORPQ on all same-size CUs of "
                                         f"{args.cu_frames} 1080p frames per step, device-resident and host-buffer (e2e) CUs/s")
        if not args.no_cpu_baseline and world == 1 and os.path.exists(REF_TU_CUDA):
            # the reference AS SHIPPED runs its hook on libtorch-CUDA (at::kCUDA, EncCu.cpp:804): the same statements on this GPU
            import shutil

            d, pt, ctu_file = _runner_inputs(16)
            try:
                gn, gdt, _ = _run_runner(pt, ctu_file, 3.0, len(os.sched_getaffinity(0)), 0, exe=REF_TU_CUDA)
                wn, wdt, _ = _run_runner(pt, ctu_file, 3.0, len(os.sched_getaffinity(0)), 1, exe=REF_TU_CUDA)
                line["reference_as_shipped_cuda"] = {
                    "load_once_ctus_per_s": gn / gdt, "as_written_load_per_call_ctus_per_s": wn / wdt, "unit": "CTU/s",
                    "note": "oracle/_ref/ref_hook_tu_cuda: the reference's own hook statements (EncCu.cpp:803-926, at::kCUDA as shipped) through "
                            "libtorch 2.11 CUDA on this B200, B = 1 like the hook; as_written reloads the TorchScript file on every call (:894-900)"}
            except Exception as e:  # noqa: BLE001 -- diagnostic only
                line["reference_as_shipped_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
            finally:
                shutil.rmtree(d, ignore_errors=True)
        if not args.no_cpu_baseline and world == 1:
            r, cores, cn, cdt, how = cpu_reference_rate(args.cpu_budget)
            aw, b120 = cpu_reference_variants(3.0)
            line["cpu_baseline"] = {"value": r, "unit": "CTU/s", "cores": cores, "kind": ref_kind(),
                                    "sample": f"{cn} CTUs at B=1 through {how} in {cdt:.1f}s",
                                    "as_written_load_per_call": aw, "b120_frame_batch": b120,
                                    "variants_note": "as_written = torch.jit.load on every call like EncCu.cpp:894-900; b120 = one 120-CTU frame per forward (CTU/s)"}
        print(json.dumps(line))
    pred.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
