#!/bin/bash
OUT=${1:-gpurun_out/r1o}
mkdir -p $OUT
timeout 120 scripts/_build/stbench 2>&1 | tee $OUT/stbench.log
echo "== trace, pair"
MMG_GEMM_PAIR=1 MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace_pair.log 2>&1; echo "trace exit $?"; cat $OUT/trace_pair.log | grep -v "whole kernel"
