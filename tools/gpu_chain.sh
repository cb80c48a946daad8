#!/bin/bash
# chained block-boundary convs (csrc/conv_chain.cu): parity tests, then per-op CUDA-event tables with and without the fusion
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_conv_chain_gpu.py tests/test_conv_tc_gpu.py -m gpu -q -x -k "chain or maxpool" 2>&1 | tail -25 | tee gpurun_out/test_chain.log
timeout -s KILL 300 python -m pytest tests/test_scorer_gpu.py -m gpu -q -x -s -k "chained" 2>&1 | tail -8 | tee gpurun_out/test_chain_scorer.log
for t in resnet50 resnet50_clip.openai; do
  timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --trunk $t --steps 5 > gpurun_out/ops_chain_$t.txt 2>&1
  grep -i "step\|total" gpurun_out/ops_chain_$t.txt | head -5
done
for r in ${CHAIN_RINGS:-}; do
  SEMDIFF_CHAIN_RING=$r timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --trunk resnet50 --steps 5 > gpurun_out/ops_chain_ring$r.txt 2>&1
  echo "ring $r"; grep -i "step\|total" gpurun_out/ops_chain_ring$r.txt | head -3; sed -n 6,13p gpurun_out/ops_chain_ring$r.txt
done
SEMDIFF_NO_POOL_FUSION=1 timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --trunk resnet50 --steps 5 > gpurun_out/ops_nopool.txt 2>&1
echo nopool; sed -n 1,4p gpurun_out/ops_nopool.txt
echo default; sed -n 1,13p gpurun_out/ops_chain_resnet50.txt
exit 0
