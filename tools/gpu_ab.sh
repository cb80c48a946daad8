for v in 0 1; do echo "DEEP_RING=$v"; SEMDIFF_X3_DEEP_RING=$v timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --precision fp16x3 --steps 3 2>&1 | grep -v "Warn\|model = " > gpurun_out/ops_deep$v.txt; grep "=== micro" gpurun_out/ops_deep$v.txt; done
timeout -s KILL 300 python -m pytest tests/test_split_gpu.py -m gpu -q 2>&1 | tail -2
timeout -s KILL 300 python -m pytest tests/test_scorer_gpu.py -m gpu -q -k "fp16x3" 2>&1 | tail -2
