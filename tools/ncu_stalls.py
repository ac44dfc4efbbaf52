#!/usr/bin/env python
"""Rank source lines of one kernel by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda`.
usage: ncu -i rep --page source --csv --print-source cuda --launch-skip K --launch-count 1 > src.csv; python tools/ncu_stalls.py src.csv [N]"""
import csv
import sys


def num(x):
    try:
        return int(x.replace(",", ""))
    except Exception:
        return 0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    for k, hi in enumerate(hdr_idx):
        end = hdr_idx[k + 1] if k + 1 < len(hdr_idx) else len(rows)
        hdr = rows[hi]
        col = {n: i for i, n in enumerate(hdr)}
        if "# Samples" not in col:
            continue
        samp = col["# Samples"]
        data = [r for r in rows[hi + 1 : end] if len(r) > samp and num(r[samp]) > 0]
        tot = sum(num(r[samp]) for r in data)
        if not tot:
            continue
        print(f"== view {k} ({rows[hi-2][1] if hi >= 2 else ''}) total samples {tot}")
        stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        data.sort(key=lambda r: -num(r[samp]))
        for r in data[:top_n]:
            st = {n[6:]: num(r[col[n]]) for n in stall_cols if num(r[col[n]])}
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print(f"{num(r[samp]):7d} {100*num(r[samp])/tot:5.1f}%  L{r[0]:>4s} {r[1].strip()[:90]:90s} {top}")


if __name__ == "__main__":
    main()
