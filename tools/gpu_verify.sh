#!/bin/bash
# One gpurun call: GPU parity tests, both bench arms, ncu launch list.  Outputs under gpurun_out/<tag>_*.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_verify.sh <tag> [full]'
tag=${1:-verify}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${tag}_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
if [ "$2" = "full" ]; then
  # skip the 3 warm-up steps (3 x 35 launches: descs + 2 slices x 17), capture one whole step
  timeout 1200 ncu --set full --clock-control none --import-source on -s 105 -c 35 -f -o gpurun_out/${tag}_full \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
  echo "ncu full rc=$?"
  ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
  ls -la gpurun_out/${tag}_full.ncu-rep
fi
