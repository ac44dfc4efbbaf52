#!/bin/bash
# One gpurun call: ncu --set full of one whole CTU-model step (35 launches) + time-only launch lists of the CU models.
tag=${1:-ncu}
mkdir -p gpurun_out
# 3 warm-up steps x 35 launches (descs + 2 slices x 17) are skipped; one whole step is captured
timeout 1500 ncu --set full --clock-control none --import-source on -s 105 -c 35 -f -o gpurun_out/${tag}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cu-frames 0 > gpurun_out/${tag}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_full.ncu-rep gpurun_out/${tag}_full_raw.csv
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"conv_umma|cu_conv1|cu_head|cu_dense" --csv --log-file gpurun_out/${tag}_cu_launches.csv \
    python bench.py --cu-only --steps 5 --warmup 3 > gpurun_out/${tag}_cu_bench_under_ncu.log 2>&1
echo "ncu cu rc=$?"
