#!/bin/bash
# last call of a round: full GPU suite, smoke, bench of both trunks (no ncu)
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/test_all.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout -s KILL 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout -s KILL 200 python bench.py --trunk resnet50_clip.openai --no-cpu-baseline > gpurun_out/bench_clip_final.json 2> gpurun_out/bench_clip.err; echo "clip rc=$?"
exit 0
