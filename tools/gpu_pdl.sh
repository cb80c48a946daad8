#!/bin/bash
# programmatic dependent launch: full GPU suite, then step time with / without it
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/test_all.log
for t in resnet50 resnet50_clip.openai; do
  timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --trunk $t --steps 5 > gpurun_out/ops_pdl_$t.txt 2>&1
  SEMDIFF_NO_PDL=1 SKIP=1 true || timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --trunk $t --steps 5 > gpurun_out/ops_nopdl_$t.txt 2>&1
  grep -i "=== micro" gpurun_out/ops_pdl_$t.txt gpurun_out/ops_nopdl_$t.txt
done
exit 0
