#!/bin/bash
OUT=${1:-gpurun_out/r1d}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
timeout 600 python scripts/kernel_bench.py > $OUT/kernel_bench.log 2>&1; echo "kernel_bench exit $?"; cat $OUT/kernel_bench.log
timeout 900 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -c 3000 $OUT/bench.log
timeout 400 python scripts/cpu_threads.py > $OUT/cpu_threads.log 2>&1; cat $OUT/cpu_threads.log
