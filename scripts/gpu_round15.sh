#!/bin/bash
OUT=${1:-gpurun_out/r1u}
mkdir -p $OUT
echo "== default (RED 32 cols, 3 stages)"
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb.log 2>&1; cut -c1-150 $OUT/kb.log | grep "resid"
echo "== RED 16 cols, 4 stages"
MMG_LIB=scripts/_build/libmmg_red16.so timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb16.log 2>&1; cut -c1-150 $OUT/kb16.log | grep "resid"
MMG_LIB=scripts/_build/libmmg_red16.so timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "linear or lnfold or residual" > $OUT/k16.log 2>&1; echo "kernels(red16) exit $?: $(tail -1 $OUT/k16.log)"
MMG_LIB=scripts/_build/libmmg_red16.so timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench16.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench16.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench16.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench.log
echo "== trace (default build flags + trace)"
MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace.log 2>&1; echo "trace exit $?"; grep -A12 "resid f32" $OUT/trace.log | grep -v "whole kernel\|MMA: issue span\|EPI: wait\|EPI: MMA last"
