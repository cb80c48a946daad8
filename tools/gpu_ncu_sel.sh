#!/bin/bash
# full-metric + source capture of selected conv_tc launches: usage tools/gpu_ncu_sel.sh <kernel regex> <skip> <count> <out name>
mkdir -p gpurun_out
python tools/ncu_target.py > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$1" --launch-skip $2 --launch-count $3 -f -o gpurun_out/$4 \
    python tools/ncu_target.py > gpurun_out/ncu_sel.log 2>&1
echo "rc=$?"; ls -la gpurun_out/$4.ncu-rep
ncu -i gpurun_out/$4.ncu-rep --page raw --csv > gpurun_out/$4_raw.csv 2>/dev/null
exit 0
