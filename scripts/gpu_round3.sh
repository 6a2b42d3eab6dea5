#!/bin/bash
OUT=${1:-gpurun_out/r1g}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_pdl.log 2>&1; echo "bench pdl exit $?"; tail -c 1300 $OUT/bench_pdl.log
MMG_PDL=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_nopdl.log 2>&1; echo "bench nopdl exit $?"; tail -c 400 $OUT/bench_nopdl.log | head -c 400; grep -o '"value": [0-9.]*' $OUT/bench_nopdl.log | head -2
timeout 300 python scripts/kernel_bench.py --only sample,attn,vq > $OUT/kernel_bench.log 2>&1; cat $OUT/kernel_bench.log | cut -c1-150
