#!/bin/bash
# One gpurun call for the end-of-round evidence: replay of a dumped encode, compute-sanitizer, GPU tests, both bench arms, ncu launch list.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_final.sh <tag>'
tag=${1:-final}
mkdir -p gpurun_out
if [ -f oracle/_ref/data/ref_cpu_1080p_q27_inputs.bin ]; then
  timeout 200 python tools/replay_dump.py oracle/_ref/data/ref_cpu_1080p_q27_inputs.bin oracle/_ref/data/ref_cpu_1080p_q27.json \
      --out gpurun_out/${tag}_replay_1080p_q27.json > gpurun_out/${tag}_replay.log 2>&1
  echo "replay rc=$?"; grep -A12 '"differences"' gpurun_out/${tag}_replay_1080p_q27.json | head -30
fi
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_cu.py > gpurun_out/${tag}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/${tag}_sanitizer_memcheck.log
MLT_CHAIN=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_cu.py > gpurun_out/${tag}_sanitizer_memcheck_chain.log 2>&1
echo "memcheck (chain) rc=$?"; tail -2 gpurun_out/${tag}_sanitizer_memcheck_chain.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['sustained']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d['ctu_latency_us'], d['e2e']['value'], d['cpu_baseline']['value'], d.get('reference_as_shipped_cuda'))"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cu-frames 0 --sustain-s 0.1 > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
