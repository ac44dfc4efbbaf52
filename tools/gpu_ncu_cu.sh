#!/bin/bash
# ncu --set full of one whole pass (23 launches) of a CU model: bash tools/gpu_ncu_cu.sh <tag> <size>
tag=${1:-ncu}; size=${2:-64}
mkdir -p gpurun_out
# skip the first warm-up pass (23 launches), capture the second one
timeout 1200 ncu --set full --clock-control none --import-source on -s 23 -c 23 -f -o gpurun_out/${tag}_cu${size}_full \
    python bench.py --cu-only --cu-sizes $size --steps 5 --warmup 3 > gpurun_out/${tag}_cu${size}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${tag}_cu${size}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_cu${size}_full_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_cu${size}_full.ncu-rep
ls -la gpurun_out/${tag}_cu${size}_full_raw.csv
