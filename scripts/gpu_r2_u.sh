#!/bin/bash
# round 2, call U: parity suite with double-buffered QKV tiles, bench, forced CTA pairs (conv-transposes) A/B, GEMM timeline
OUT=${1:-gpurun_out/r2u}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
pr() { python -c "
import json
d=json.load(open('$1')); print('$2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"; }
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log > $OUT/bench.json; pr $OUT/bench.json default
( MMG_GEMM_PAIR=1 timeout 600 python bench.py --no-extras ) > $OUT/bench_pair1.log 2>&1; grep "^{" $OUT/bench_pair1.log > $OUT/bench_pair1.json; pr $OUT/bench_pair1.json pair1
MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace_gemm.log 2>&1; echo "trace exit $?"; grep -A11 "^--- qkv" $OUT/trace_gemm.log | cut -c1-120
