#!/bin/bash
OUT=${1:-gpurun_out/r1s}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "linear or lnfold or residual" > $OUT/k_red.log 2>&1; echo "kernels(red) exit $?: $(tail -1 $OUT/k_red.log)"; grep -h "^FAILED\|^ERROR" $OUT/k_red.log | head
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_red.log 2>&1; cut -c1-150 $OUT/kb_red.log | grep -v logits
MMG_GEMM_RED=0 timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_nored.log 2>&1; cut -c1-150 $OUT/kb_nored.log | grep "resid"
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu --no-header -p no:cacheprovider > $OUT/models.log 2>&1; echo "models exit $?: $(tail -1 $OUT/models.log)"; grep -h "^FAILED\|^ERROR" $OUT/models.log | head
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench.log
MMG_GEMM_RED=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_nored.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_nored.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench_nored.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline --global-batch 8 > $OUT/bench_b8.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_b8.log
MMG_GEMM_RED=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline --global-batch 8 > $OUT/bench_b8_nored.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_b8_nored.log
