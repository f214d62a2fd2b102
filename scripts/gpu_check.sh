#!/bin/bash
# Round-trip check on a B200 box (run through gpurun): GPU parity tests, smoke, short benches.
# Everything lands in gpurun_out/.   usage: scripts/gpu_check.sh [tests] [smoke] [bench] [kbench] [launches]
mkdir -p gpurun_out
what="${*:-tests smoke bench kbench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
rm -f gpurun_out/parity_report.txt
for w in $what; do
  case $w in
    tests)
      timeout 1700 python -m pytest tests -m gpu -q --maxfail=30 --timeout=300 -x 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
      echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
      tail -8 gpurun_out/pytest_gpu.log ;;
    tests_all)
      timeout 1700 python -m pytest tests -m gpu -q --maxfail=60 --timeout=300 2>&1 | tail -120 > gpurun_out/pytest_gpu.log
      echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
      tail -30 gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
      tail -6 gpurun_out/smoke.log ;;
    bench)
      timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
      echo "bench exit: $?"; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json ;;
    bench_all)
      for wl in vit-tiny-rope-axial-fp32 vit-tiny-polynomial-fp32 vit-tiny-relative-fp32 vitl16-384-rope-axial-bf16 vitb16-512-rope-mixed-infer-bf16; do
        timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
        echo "bench $wl exit: $?"; tail -2 gpurun_out/bench_$wl.err; cat gpurun_out/bench_$wl.json
      done ;;
    sweep5)  # BASELINE configs[4]: batch sweep of the 1025-token inference config (per GPU; the 8-GPU run is 8 replicas)
      for bs in 1 8 32 64 128; do
        timeout 300 python bench.py --workload vitb16-512-rope-mixed-infer-bf16 --batch $bs --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu \
          > gpurun_out/bench_cfg5_b$bs.json 2> gpurun_out/bench_cfg5_b$bs.err
        echo "cfg5 batch $bs exit: $?"; cut -c1-160 gpurun_out/bench_cfg5_b$bs.json
      done ;;
    kbench)
      for geo in vitb vitl x512; do timeout 300 python scripts/kbench.py $geo > gpurun_out/kbench_$geo.log 2>&1; cat gpurun_out/kbench_$geo.log; done ;;
    launches)
      timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_plain.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_ncu.log 2>&1
      echo "launch list exit: $?" ;;
  esac
done
