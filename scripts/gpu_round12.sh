#!/bin/bash
OUT=${1:-gpurun_out/r1r}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -20
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench.log
MMG_GEMM_STAGED=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_nostage.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_nostage.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench_nostage.log
timeout 300 python scripts/kernel_bench.py --only gemm,sample > $OUT/kernel_bench.log 2>&1; cut -c1-150 $OUT/kernel_bench.log
