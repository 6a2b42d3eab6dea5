#!/bin/bash
OUT=${1:-gpurun_out/r1d}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
timeout 600 python scripts/kernel_bench.py > $OUT/kernel_bench.log 2>&1; echo "kernel_bench exit $?"; cat $OUT/kernel_bench.log
timeout 900 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -c 3000 $OUT/bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logits_sample -s 2 -c 1 -f -o $OUT/prof_sample python scripts/kernel_bench.py --only sample --iters 1 > $OUT/ncu_sample.log 2>&1; echo "ncu sample exit $?"
