#!/bin/bash
OUT=${1:-gpurun_out/r1v}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "linear or lnfold or residual" > $OUT/k.log 2>&1; echo "kernels exit $?: $(tail -1 $OUT/k.log)"; grep -h "^FAILED\|^ERROR\|Error" $OUT/k.log | head
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb.log 2>&1; cut -c1-150 $OUT/kb.log | grep -v "logits\|square"
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu --no-header -p no:cacheprovider > $OUT/models.log 2>&1; echo "models exit $?: $(tail -1 $OUT/models.log)"; grep -h "^FAILED\|^ERROR" $OUT/models.log | head
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench.log
MMG_GEMM_RED=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_nored.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_nored.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench_nored.log
MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace.log 2>&1; echo "trace exit $?"; grep -A12 "resid f32" $OUT/trace.log | grep -v "whole kernel\|MMA: issue span\|EPI: wait\|EPI: MMA last"
