#!/bin/bash
OUT=${1:-gpurun_out/r1t}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "linear or lnfold or residual or conv" > $OUT/k.log 2>&1; echo "kernels exit $?: $(tail -1 $OUT/k.log)"; grep -h "^FAILED\|^ERROR" $OUT/k.log | head
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb.log 2>&1; cut -c1-150 $OUT/kb.log | grep -v logits
MMG_GEMM_NFAST=0 timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_mfast.log 2>&1; cut -c1-150 $OUT/kb_mfast.log | grep -v "logits\|square"
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench.log
MMG_GEMM_NFAST=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_mfast.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_mfast.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench_mfast.log
