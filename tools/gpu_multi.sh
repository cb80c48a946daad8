#!/bin/bash
# One launch, all three workloads of BASELINE.json on N GPUs of the box: headline (configs[1]), 10k sweep (configs[3]),
# 1024x1024 pairs (configs[4]).  usage: tools/gpu_multi.sh N [extra bench.py args]   -> gpurun_out/r2_multi_${N}gpu.jsonl
N=${1:-1}; shift
mkdir -p gpurun_out
out=gpurun_out/r2_multi_${N}gpu
nvidia-smi topo -m > ${out}_topo.txt 2>&1; nproc >> ${out}_topo.txt; numactl -H >> ${out}_topo.txt 2>&1
if [ "$N" = "1" ]; then
  timeout -s KILL 1500 python bench.py --detail --workload pairs224,sweep10k,hires1024 --steps 20 --warmup 5 "$@" > $out.jsonl 2> $out.err
else
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --detail --gpus $N --workload pairs224,sweep10k,hires1024 --steps 20 --warmup 5 "$@" > $out.jsonl 2> $out.err
fi
echo "rc=$?"; tail -5 $out.err; cut -c1-1500 $out.jsonl
exit 0
