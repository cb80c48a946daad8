#!/bin/bash
# split-precision loop: kernel tests, scorer parity, per-op table (usage: tools/gpu_x3.sh [tag] [chunk sizes...])
mkdir -p gpurun_out
tag=${1:-x3}; shift
timeout -s KILL 400 python -m pytest tests/test_split_gpu.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -90 | tee gpurun_out/test_split_$tag.log
for c in ${@:-2}; do
  SEMDIFF_X3_CHUNK_KB=$c timeout -s KILL 300 python tools/x3_accuracy.py 2>&1 | grep -v Warning | tee -a gpurun_out/x3_accuracy_$tag.log
done
timeout -s KILL 600 python -m pytest tests/test_scorer_gpu.py -m gpu -q -s -k "fp16x3" 2>&1 | grep "parity\|passed\|failed\|Error" | tee gpurun_out/test_scorer_$tag.log
timeout -s KILL 300 python tools/profile_ops.py --pairs 256 --microbatch 256 --precision fp16x3 --steps 3 > gpurun_out/ops_$tag.txt 2>&1
head -4 gpurun_out/ops_$tag.txt
exit 0
