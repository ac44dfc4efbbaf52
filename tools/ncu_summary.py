#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` (one row per launch, ~2400 columns) into the per-kernel table kept under
profiles/.  usage: python tools/ncu_summary.py raw.csv > summary.csv"""
import csv
import re
import sys

COLS = [
    ("time[ms]", "gpu__time_duration.sum", 1e-6),
    ("dram_rd[GB]", "dram__bytes_read.sum", None),
    ("dram_wr[GB]", "dram__bytes_write.sum", None),
    ("dram_pct[%]", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("tensor_pipe_active_pct[%]", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("sm_clock[GHz]", "sm__cycles_elapsed.avg.per_second", None),
    ("tc_smem_wavefront_pct[%]", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
    ("lsu_smem_wavefront_pct[%]", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1),
    ("l2_hit[%]", "lts__t_sector_hit_rate.pct", 1),
    ("regs", "launch__registers_per_thread", 1),
    ("grid", "launch__grid_size", 1),
    ("block", "launch__block_size", 1),
]
UNIT = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def short(name):
    m = re.search(r"ConvCfg<([^>]*)>", name)
    if m:
        return "conv_umma<" + m.group(1).replace(", ", "_").replace("(int)", "") + ">"
    return re.sub(r"\(.*", "", name).split("::")[-1].replace("void ", "")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("kernel," + ",".join(c[0] for c in COLS))
    tot = [0.0, 0.0, 0.0]
    for r in data:
        out = [short(r[idx["Kernel Name"]])]
        for k, (label, metric, _) in enumerate(COLS):
            if metric not in idx:
                out.append("")
                continue
            v = float(r[idx[metric]].replace(",", "")) if r[idx[metric]] not in ("", "n/a") else float("nan")
            u = units[idx[metric]]
            if label.endswith("[ms]") or label.endswith("[GB]") or label.endswith("[GHz]"):
                v *= UNIT.get(u, 1.0)
            if k < 3:
                tot[k] += v
            out.append(f"{v:.6g}")
        print(",".join(out))
    print(f"# total: {tot[0]:.3f} ms, DRAM read {tot[1]:.3f} GB + write {tot[2]:.3f} GB = {tot[1] + tot[2]:.3f} GB over {len(data)} launches")


if __name__ == "__main__":
    main()
