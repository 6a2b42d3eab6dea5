#!/bin/bash
OUT=${1:-gpurun_out/qkvt}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT > /dev/null 2>&1; cat $OUT/summary.txt; grep -h "^FAILED\|^ERROR" $OUT/*.log | head
echo "== default (QKVT on)"; timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb.log 2>&1; grep "qkv" $OUT/kb.log | cut -c1-140
echo "== QKVT off"; MMG_GEMM_QKVT=0 timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_off.log 2>&1; grep "qkv" $OUT/kb_off.log | cut -c1-140
echo "== QKVT + pair"; MMG_GEMM_PAIR=1 timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_pair.log 2>&1; grep "qkv" $OUT/kb_pair.log | cut -c1-140
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench.log
MMG_GEMM_QKVT=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_off.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_off.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench_off.log
