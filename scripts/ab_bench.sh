#!/bin/bash
# A/B of two prebuilt libraries inside one gpurun call (same box, same clocks): scripts/ab_bench.sh ab_libs/old.so ab_libs/new.so [rounds]
A=$1; B=$2; R=${3:-2}
for r in $(seq 1 $R); do
  for lib in $A $B; do
    cp $lib vit_rpe_rope_b200/lib/libvrr_b200.so
    timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['hot_path_kernels']
print('$lib', round(d['value']), round(d['ms_per_step'],2), 'fc1_fwd', round(k['fc1_fwd']['avg_launch_ms']*1e3,1), 'fc2_dx', round(k['fc2_dx']['avg_launch_ms']*1e3,1), 'fc2_fwd', round(k['fc2_fwd']['avg_launch_ms']*1e3,1), 'qkv', round(k['qkv_rope_fwd']['avg_launch_ms']*1e3,1), d['clocks']['sm_mhz'])"
  done
done
