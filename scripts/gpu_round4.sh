#!/bin/bash
OUT=${1:-gpurun_out/r1i}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kernel_bench_gemm.log 2>&1; cat $OUT/kernel_bench_gemm.log | cut -c1-150
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench.log
