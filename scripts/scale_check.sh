#!/bin/bash
# Multi-GPU bench lines of the three data-parallel configs (run through gpurun --gpus N): scripts/scale_check.sh N [workloads...]
N=${1:-8}; shift
mkdir -p gpurun_out
wls="${*:-vitb16-224-rope-mixed-bf16 vitl16-384-rope-axial-bf16 vitb16-512-rope-mixed-infer-bf16}"
port=29511
for wl in $wls; do
  port=$((port + 1))
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-reference-gpu \
    > gpurun_out/scale_${wl}_n$N.json 2> gpurun_out/scale_${wl}_n$N.err
  echo "$wl N=$N exit: $?"; tail -2 gpurun_out/scale_${wl}_n$N.err | cut -c1-300; cut -c1-420 gpurun_out/scale_${wl}_n$N.json
done
