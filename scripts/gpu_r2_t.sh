#!/bin/bash
# round 2, call T: parity suite after the two-chunk TMEM preload, bench, GEMM timeline + microbenchmark
OUT=${1:-gpurun_out/r2t}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log > $OUT/bench.json; python -c "
import json
d=json.load(open('$OUT/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"
MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace_gemm.log 2>&1; echo "trace exit $?"; grep -A4 "^--- qkv\|^--- geglu" $OUT/trace_gemm.log | cut -c1-120
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_gemm.log 2>&1; cut -c1-170 $OUT/kb_gemm.log
( timeout 300 python bench.py --no-extras --global-batch 8 ) > $OUT/bench_b8.log 2>&1; grep "^{" $OUT/bench_b8.log | cut -c1-200
