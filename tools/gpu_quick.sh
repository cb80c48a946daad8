#!/bin/bash
# quick loop: conv kernel tests, then the per-op tables of both trunks (usage: tools/gpu_quick.sh [pytest -k expression])
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_conv_tc_gpu.py tests/test_conv_chain_gpu.py -m gpu -q -x ${1:+-k "$1"} 2>&1 | tail -8 | tee gpurun_out/test_quick.log
for t in resnet50 resnet50_clip.openai; do
  timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --trunk $t --steps 5 > gpurun_out/ops_quick_$t.txt 2>&1
  grep "=== micro" gpurun_out/ops_quick_$t.txt
done
sed -n 2,6p gpurun_out/ops_quick_resnet50.txt; sed -n 2,8p gpurun_out/ops_quick_resnet50_clip.openai.txt
exit 0
