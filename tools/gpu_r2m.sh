#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/test_all_r2m.log
timeout -s KILL 600 python bench.py --workload unet --modes fp16x3,bf16 > gpurun_out/r2_unet_1gpu.json 2> gpurun_out/unet.err; tail -2 gpurun_out/unet.err; cut -c1-1500 gpurun_out/r2_unet_1gpu.json
for kb in 6 8 10; do echo "N256_MIN_KB=$kb"; SEMDIFF_N256_MIN_KB=$kb timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --precision fp16x3 --steps 3 2>&1 | grep "=== micro"; done
timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --precision bf16 --steps 5 2>&1 | grep "=== micro"
timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --precision bf16 --trunk resnet50_clip.openai --steps 5 2>&1 | grep "=== micro"
timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --precision fp16x3 --trunk resnet50_clip.openai --steps 3 2>&1 | grep "=== micro"
exit 0
