#!/bin/bash
# Shorter end-of-round check (after the border-warp change): memcheck of every kernel, GPU tests, both bench arms, ncu launch list.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_final2.sh <tag>'
tag=${1:-final2}
mkdir -p gpurun_out
timeout -s KILL 400 compute-sanitizer --tool memcheck python tools/sanitize_cu.py > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/${tag}_sanitizer_memcheck.log
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
timeout -s KILL 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['ctu_latency_us'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])
print({k: (round(v['cus_per_s'] / 1e6, 3), round(v['tflops'], 1)) for k, v in d['cu_models'].items()})"
timeout -s KILL 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
tail -c 400 gpurun_out/${tag}_bench_ref.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cu-frames 0 --sustain-s 0.1 > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
