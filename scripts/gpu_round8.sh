#!/bin/bash
OUT=${1:-gpurun_out/r1n}
mkdir -p $OUT
echo "== pair forced"; MMG_GEMM_PAIR=1 timeout 120 python scripts/pair_check.py 2>&1 | tail -12; echo "exit $?"
echo "== pair off";    MMG_GEMM_PAIR=0 timeout 120 python scripts/pair_check.py 2>&1 | tail -8
nvidia-smi --query-gpu=name --format=csv,noheader || exit 1
echo "== kernel tests, pair forced"
MMG_GEMM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x --no-header -p no:cacheprovider > $OUT/k_pair.log 2>&1; echo "exit $?: $(tail -1 $OUT/k_pair.log)"; grep -h "FAILED\|Error" $OUT/k_pair.log | head
echo "== trace, pair"
MMG_GEMM_PAIR=1 MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace_pair.log 2>&1; echo "trace exit $?"; cat $OUT/trace_pair.log | grep -v "whole kernel"
echo "== bench default policy"
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench.log
MMG_GEMM_PAIR=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_nopair.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_nopair.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench_nopair.log
