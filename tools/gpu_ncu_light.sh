#!/bin/bash
# light metric set over every kernel of one forward (a few replays per kernel) + CSV export
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active
python -c "import bench; print(bench.kernel_sources_digest())" > gpurun_out/light.digest
python tools/ncu_target.py $@ > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/light.csv python tools/ncu_target.py $@ > gpurun_out/ncu_light.log 2>&1
echo "light rc=$?"
exit 0
