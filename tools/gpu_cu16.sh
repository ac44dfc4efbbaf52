#!/bin/bash
# CU-network checks after the 16-px composed stem: parity tests, 16-px bench A/B (fused vs MLT_CU_UNFUSED), in-VTM run of the CU-size models
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cu_parity.py -x -q 2>&1 | tail -6
show='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1])["cu_models"]["16"]; print({k:d[k] for k in ("cus_per_s","e2e_cus_per_s","tflops","gpu_launches_per_step","cu_latency_us")})'
echo "== 16 px fused";   timeout 200 python bench.py --cu-only --cu-sizes 16 --steps 30 2>/dev/null | python -c "$show"
echo "== 16 px unfused"; MLT_CU_UNFUSED=1 timeout 200 python bench.py --cu-only --cu-sizes 16 --steps 30 2>/dev/null | python -c "$show"
timeout 500 python tools/vtm_run.py --size 416x240 --bits 8 --frames 4 --qps 27,32 --encoders ref_cpu_cu,mlt_cu --jobs 4 --ref-threads 4 \
    --out gpurun_out/r02j_vtm_cu.json > gpurun_out/r02j_vtm_cu.log 2>&1
grep -A10 '"pairs"' gpurun_out/r02j_vtm_cu.log | head -30; tail -4 gpurun_out/r02j_vtm_cu.log | cut -c1-420
