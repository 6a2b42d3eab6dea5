#!/bin/bash
OUT=${1:-gpurun_out/r2i}
mkdir -p $OUT
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_gemm_default.log 2>&1; cut -c1-170 $OUT/kb_gemm_default.log
MMG_GEMM_PAIR=1 timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_gemm_pair.log 2>&1; cut -c1-170 $OUT/kb_gemm_pair.log
( MMG_GEMM_PAIR=1 timeout 600 python bench.py --no-extras ) > $OUT/bench_pair.log 2>&1; echo "bench pair exit $?"; grep "^{" $OUT/bench_pair.log | cut -c1-200
