#!/bin/bash
# programmatic dependent launch A/B (opt-in SEMDIFF_PDL=1): step time with / without it, both trunks (profiles/r1_pdl.txt)
mkdir -p gpurun_out
for t in resnet50 resnet50_clip.openai; do
  SEMDIFF_PDL=1 timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --trunk $t --steps 10 > gpurun_out/ops_pdl_$t.txt 2>&1
  timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --trunk $t --steps 10 > gpurun_out/ops_nopdl_$t.txt 2>&1
  grep -i "=== micro" gpurun_out/ops_pdl_$t.txt gpurun_out/ops_nopdl_$t.txt
done
exit 0
