#!/bin/bash
# Round-end evidence in one call: full GPU suite, smoke, bench lines (RN50, CLIP, reference arm), per-op tables, ncu launch
# list of the bench command + light metric sets (bf16 stamped with the kernel-source digest, fp16x3).
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2_final_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -5 | tee gpurun_out/r2_final_smoke.log
timeout -s KILL 600 python bench.py --detail > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -2 gpurun_out/r2_final_bench.err
timeout -s KILL 600 python bench.py --detail --trunk resnet50_clip.openai --no-cpu-baseline > gpurun_out/r2_final_bench_clip.json 2>> gpurun_out/r2_final_bench.err
timeout -s KILL 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2_final_bench_reference.json 2>> gpurun_out/r2_final_bench.err
for prec in bf16 fp16x3; do for t in resnet50 resnet50_clip.openai; do
  timeout -s KILL 200 python tools/profile_ops.py --pairs 256 --microbatch 256 --precision $prec --trunk $t --steps 5 2>&1 | grep -v "Warn\|model = " > gpurun_out/r2_final_ops_${prec}_$t.txt
  grep "=== micro" gpurun_out/r2_final_ops_${prec}_$t.txt
done; done
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
exit 0
