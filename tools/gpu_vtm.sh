#!/bin/bash
# One gpurun call: the in-VTM runs (SURVEY.md section 8d config 1 and a cut of config 3) with the encoders of oracle/_ref/.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_vtm.sh <tag> [config1|cut3 ...]'
tag=${1:-vtm}; shift
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt
ldd oracle/_ref/EncoderApp_mlt | grep -i "mltcnn\|not found" >> gpurun_out/${tag}_host.txt
for what in "$@"; do
  case $what in
    config1)  # 416x240 8-bit, 8 frames, QP 32: every encoder, two at a time
      timeout 900 python tools/vtm_run.py --config 1 --encoders ref_cpu,ref_cuda,mlt,staged,anchor --jobs 2 \
          --out gpurun_out/${tag}_config1.json > gpurun_out/${tag}_config1.log 2>&1
      echo "config1 rc=$?"; head -60 gpurun_out/${tag}_config1.log ;;
    config1seq)  # the three timed encoders one at a time (clean Total Time)
      timeout 900 python tools/vtm_run.py --config 1 --encoders ref_cpu,mlt,anchor --jobs 1 \
          --out gpurun_out/${tag}_config1seq.json > gpurun_out/${tag}_config1seq.log 2>&1
      echo "config1seq rc=$?"; head -40 gpurun_out/${tag}_config1seq.log ;;
    cut3)  # affordable cut of config 3: 1920x1080 10-bit, few frames, 4 QPs, all encodes side by side
      timeout ${CUT3_TIMEOUT:-2400} python tools/vtm_run.py --size ${CUT3_SIZE:-1920x1080} --bits 10 --frames ${CUT3_FRAMES:-5} --qps 22,27,32,37 \
          --encoders ${CUT3_ENCODERS:-ref_cpu,mlt,anchor,prepass0,prepass8} --jobs ${CUT3_JOBS:-12} --ref-threads ${CUT3_REF_THREADS:-2} \
          --out gpurun_out/${tag}_cut3.json > gpurun_out/${tag}_cut3.log 2>&1
      echo "cut3 rc=$?"; head -120 gpurun_out/${tag}_cut3.log ;;
  esac
done
