#!/bin/bash
# bench + ncu launch list + one full capture of the top conv kernel (run each under ncu only after the plain run exits 0)
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --steps 1 --warmup 3 --pairs 64 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --pairs 64 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
exit 0
