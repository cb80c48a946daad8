#!/bin/bash
# round-end evidence: full GPU test suite, smoke, bench (ours + reference arm), ncu launch list of the bench command
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/test_all.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_final.json 2>&1
timeout -s KILL 300 python bench.py --trunk resnet50_clip.openai --no-cpu-baseline > gpurun_out/bench_clip_final.json 2> gpurun_out/bench_clip.err
timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/bench_launches.csv
exit 0
