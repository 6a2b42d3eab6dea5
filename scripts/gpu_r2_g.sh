#!/bin/bash
OUT=${1:-gpurun_out/r2g}
mkdir -p $OUT
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tc_logits_kernel -s 2 -c 1 -f -o $OUT/prof_lf python scripts/kernel_bench.py --only fused --iters 1 > $OUT/ncu_lf.log 2>&1; echo "ncu lf exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:logits_finish -s 2 -c 1 -f -o $OUT/prof_fin python scripts/kernel_bench.py --only fused --iters 1 > $OUT/ncu_fin.log 2>&1; echo "ncu fin exit $?"
ls -la $OUT
