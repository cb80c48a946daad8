#!/bin/bash
# One gpurun call: box facts + GPU tests, each stage under its own timeout so that a hung kernel cannot eat the lease.
# usage (from the repo root, on the GPU box): bash tools/gpu_round.sh [stage ...]   (default: info kernels tc scorer)
mkdir -p gpurun_out
stages="${@:-info kernels tc scorer}"
for s in $stages; do
  case $s in
    info)
      { nvidia-smi; nvidia-smi topo -m; lscpu | head -25; nproc; free -g | head -2; ls /root/reference 2>&1 | head -3; \
        python -c "import timm" 2>&1 | tail -1; } > gpurun_out/box_info.txt 2>&1 ;;
    kernels) timeout -s KILL 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/test_kernels.log ;;
    tc)      timeout -s KILL 240 python -m pytest tests/test_conv_tc_gpu.py tests/test_conv_chain_gpu.py -m gpu -q 2>&1 | tail -60 | tee gpurun_out/test_tc.log ;;
    scorer)  timeout -s KILL 400 python -m pytest tests/test_scorer_gpu.py -m gpu -q -s 2>&1 | tail -80 | tee gpurun_out/test_scorer.log ;;
    all)     timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/test_all.log ;;
    smoke)   timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/smoke.log ;;
    bench)   timeout -s KILL 900 python bench.py 2>&1 | tail -5 | tee gpurun_out/bench.log ;;
  esac
done
exit 0
