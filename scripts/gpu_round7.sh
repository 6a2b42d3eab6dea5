#!/bin/bash
OUT=${1:-gpurun_out/r1m}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -q -m gpu -k "vq or vae" --no-header -p no:cacheprovider > $OUT/vq_tests.log 2>&1; echo "tests exit $?: $(tail -1 $OUT/vq_tests.log)"
grep -h "FAILED\|Error" $OUT/vq_tests.log | head
MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace_gemm.log 2>&1; echo "trace exit $?"; cat $OUT/trace_gemm.log
