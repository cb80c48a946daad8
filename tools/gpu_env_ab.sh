#!/bin/bash
# A/B of one environment knob on the per-op table: usage tools/gpu_env_ab.sh VAR=value [VAR=value ...]
mkdir -p gpurun_out
timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --steps 5 > gpurun_out/ops_ab_base.txt 2>&1
grep "=== micro" gpurun_out/ops_ab_base.txt
for kv in "$@"; do
  env $kv timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --steps 5 > gpurun_out/ops_ab_$kv.txt 2>&1
  echo "$kv"; grep "=== micro" gpurun_out/ops_ab_$kv.txt
  diff <(cut -c1-50 gpurun_out/ops_ab_base.txt) <(cut -c1-50 gpurun_out/ops_ab_$kv.txt) | grep "^[<>]" | awk '{print}' | head -40
done
exit 0
