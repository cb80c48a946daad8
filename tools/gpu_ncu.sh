#!/bin/bash
# launch list (time per kernel) + full-metric capture of every kernel of one forward pass (exported to CSV on the box;
# the .ncu-rep of all 62 kernels is > 64 MiB) + a small .ncu-rep with source for the first kernels of the trunk
mkdir -p gpurun_out
python tools/ncu_target.py > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/ncu_target.py > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
python tools/ncu_target.py > gpurun_out/plain2.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none -f -o /tmp/prof_full \
    python tools/ncu_target.py > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
ncu -i /tmp/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2> gpurun_out/ncu_export.log
python tools/ncu_target.py > gpurun_out/plain3.log 2>&1 && \
ncu --profile-from-start off --set full --import-source on --clock-control none --launch-skip 1 --launch-count 6 -f -o gpurun_out/prof_head6 \
    python tools/ncu_target.py > gpurun_out/ncu3.log 2>&1
echo "head6 capture rc=$?"; ls -la gpurun_out/
exit 0
