#!/bin/bash
# Round-2 ncu evidence in one call: (1) launch list of the bench command, (2) light metric set over one forward in bf16
# (stamped with the kernel-source digest) and in fp16x3, (3) --set full + source capture of four conv_tc_split launches.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-x3 > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-x3 > gpurun_out/r2_bench_ncu.log 2>&1
echo "launch list rc=$?"
python -c "import bench; print(bench.kernel_sources_digest())" > gpurun_out/r2_light_bf16.digest
cp gpurun_out/r2_light_bf16.digest gpurun_out/r2_light_x3.digest
python tools/ncu_target.py > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r2_light_bf16.csv python tools/ncu_target.py > gpurun_out/ncu_light.log 2>&1
echo "light bf16 rc=$?"
python tools/ncu_target.py --precision fp16x3 > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r2_light_x3.csv python tools/ncu_target.py --precision fp16x3 > gpurun_out/ncu_light.log 2>&1
echo "light x3 rc=$?"
ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:conv_tc_split" --launch-skip 24 --launch-count 4 -f -o gpurun_out/r2_split_full \
    python tools/ncu_target.py --precision fp16x3 > gpurun_out/ncu_sel.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/r2_split_full.ncu-rep --page raw --csv > gpurun_out/r2_split_full_raw.csv 2>/dev/null
ls -la gpurun_out/ | grep r2_
exit 0
