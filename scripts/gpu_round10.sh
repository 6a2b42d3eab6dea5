#!/bin/bash
OUT=${1:-gpurun_out/r1p}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider > $OUT/kernels.log 2>&1; echo "kernels exit $?: $(tail -1 $OUT/kernels.log)"
grep -h "^FAILED\|^ERROR" $OUT/kernels.log | head -20
MMG_GEMM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "linear or conv or vq" > $OUT/kernels_pair.log 2>&1; echo "kernels(pair) exit $?: $(tail -1 $OUT/kernels_pair.log)"
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu --no-header -p no:cacheprovider > $OUT/models.log 2>&1; echo "models exit $?: $(tail -1 $OUT/models.log)"
grep -h "^FAILED\|^ERROR" $OUT/models.log | head -20
echo "== trace"
MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace.log 2>&1; echo "trace exit $?"; grep -v "whole kernel\|MMA: issue span\|EPI: wait\|EPI: MMA last" $OUT/trace.log
echo "== trace, pair"
MMG_GEMM_PAIR=1 MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace_pair.log 2>&1; echo "trace exit $?"; grep -v "whole kernel\|MMA: issue span\|EPI: wait\|EPI: MMA last" $OUT/trace_pair.log
echo "== bench"
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench.log
MMG_GEMM_PAIR=1 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_pair.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_pair.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench_pair.log
