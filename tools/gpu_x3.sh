#!/bin/bash
# split-precision loop: kernel tests, scorer parity, per-op table (usage: tools/gpu_x3.sh [tag])
mkdir -p gpurun_out
tag=${1:-x3}
timeout -s KILL 400 python -m pytest tests/test_split_gpu.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -70 | tee gpurun_out/test_split_$tag.log
timeout -s KILL 600 python -m pytest tests/test_scorer_gpu.py -m gpu -q -s -k "fp16x3" 2>&1 | tail -30 | tee gpurun_out/test_scorer_$tag.log
timeout -s KILL 300 python tools/profile_ops.py --pairs 256 --microbatch 256 --precision fp16x3 --steps 3 > gpurun_out/ops_$tag.txt 2>&1
head -70 gpurun_out/ops_$tag.txt
exit 0
