#!/bin/bash
# Round-trip check on a B200 box (run through gpurun): GPU parity tests, smoke, short benches.
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
# tensor-core kernels first, each under a timeout: a broken one is reported and the run continues on SIMT
for probe in attn qkv bwd patch; do
  timeout 300 python scripts/tc_probe.py $probe > gpurun_out/probe_$probe.log 2>&1
  rc=$?; echo "probe $probe exit: $rc" >> gpurun_out/probe_$probe.log
  if [ $rc -ne 0 ]; then export VRR_IMPL=simt; echo "PROBE $probe FAILED -> VRR_IMPL=simt"; tail -5 gpurun_out/probe_$probe.log; fi
done
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --timeout=300 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
for wl in vit-tiny-rope-axial-fp32 vit-tiny-polynomial-fp32 vitb16-224-rope-mixed-bf16; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "bench $wl exit: $?" >> gpurun_out/smoke.log
done
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench_*.json
