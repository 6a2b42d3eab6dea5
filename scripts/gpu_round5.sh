#!/bin/bash
OUT=${1:-gpurun_out/r1k}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "FAILED\|Error\|error" $OUT/*.log | head -40
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench.log
