#!/bin/bash
# gpurun helper: probe the CTA-pair GEMM under a timeout first; only if it passes run the rest.
mkdir -p gpurun_out
timeout 180 python scripts/gemm_probe.py check > gpurun_out/gemm_check.log 2>&1; rc=$?
echo "gemm check exit: $rc" >> gpurun_out/gemm_check.log
tail -25 gpurun_out/gemm_check.log
if [ $rc -ne 0 ]; then grep -c BAD gpurun_out/gemm_check.log; grep BAD gpurun_out/gemm_check.log | head -40; exit 1; fi
timeout 180 python scripts/gemm_probe.py time > gpurun_out/gemm_time.log 2>&1; echo "gemm time exit: $?" >> gpurun_out/gemm_time.log
cat gpurun_out/gemm_time.log
bash scripts/gpu_check.sh "$@"
