#!/bin/bash
# After tools/gpu_final_r2.sh: turn gpurun_out/ scratch into the tracked round-2 records under profiles/.
set -e
python tools/summarize_ncu_light.py gpurun_out/r2_light_bf16.csv profiles/r2_ncu_light_rn50_256pairs.md "ImageNet RN50 trunk, bf16, final round-2 build" profiles/roofline_traffic.json | tail -3
python tools/summarize_ncu_light.py gpurun_out/r2_light_x3.csv profiles/r2_ncu_light_rn50_x3_256pairs.md "ImageNet RN50 trunk, fp16x3 (split precision), final round-2 build" | tail -3
sed -i 's/(512 images, 224x224, bf16), `tools\/gpu_ncu_light.sh`/(512 images, 224x224, fp16x3 = hi + lo fp16 pairs, 4 bytes per element), `tools\/gpu_final_r2.sh`/; s/DRAM \([0-9.]*\) GB (algorithmic 1.54 GB)/DRAM \1 GB (algorithmic 3.08 GB: hi and lo halves of every tap)/' profiles/r2_ncu_light_rn50_x3_256pairs.md
sed -i 's/`tools\/gpu_ncu_light.sh`/`tools\/gpu_final_r2.sh`/' profiles/r2_ncu_light_rn50_256pairs.md
cp gpurun_out/r2_bench_launches.csv profiles/r2_ncu_launches_bench.csv
for f in bf16_resnet50 bf16_resnet50_clip.openai fp16x3_resnet50 fp16x3_resnet50_clip.openai; do cp gpurun_out/r2_final_ops_$f.txt profiles/r2_per_op_$f.txt; done
python - <<'PY'
import json, subprocess
head = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
for src, dst in (('r2_final_bench.json', 'r2_bench_1gpu.json'), ('r2_final_bench_clip.json', 'r2_bench_1gpu_clip.json'),
                 ('r2_final_bench_reference.json', 'r2_bench_reference_arm.json')):
    d = json.loads([l for l in open('gpurun_out/' + src) if l.startswith('{')][0]); d['commit'] = head
    json.dump(d, open('profiles/' + dst, 'w'), indent=1)
    if 'e2e' in d and 'roofline' in d:
        print(dst, round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'], 3), 'burst', round(d['roofline']['frac_of_burst_peak'], 3),
              'traffic', d['roofline']['traffic'], 'x3', round(d['fp16x3']['value']), round(d['fp16x3']['e2e']['value']))
PY
python -c "import bench, json; print('digest ok:', bench.kernel_sources_digest() == json.load(open('profiles/roofline_traffic.json'))['kernel_sources_sha256'])"
