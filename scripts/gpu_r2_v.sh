#!/bin/bash
# round 2, call V: A-resident CTA-pair GEMM (ARES): targeted parity first (short timeouts), then the suite, then A/B timings
OUT=${1:-gpurun_out/r2v}
mkdir -p $OUT
timeout -k 5 240 python -m pytest tests/test_gpu_full_config.py tests/test_gpu_kernels.py -q -m gpu -x --no-header -p no:cacheprovider -k "gemm_generate_shape_qkv or gemm_generate_shape_ff_geglu or qkv_epilogue or geglu or bitwise" > $OUT/targeted.log 2>&1; rc=$?; echo "targeted exit $rc: $(tail -1 $OUT/targeted.log)"
if [ $rc -ne 0 ]; then grep -n "Error\|error\|FAILED\|assert" $OUT/targeted.log | head -20; export MMG_GEMM_ARES=0; echo "ARES disabled for the rest of this call"; fi
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
pr() { python -c "
import json
d=json.load(open('$1')); print('$2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"; }
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_gemm.log 2>&1; grep "qkv\|ff1" $OUT/kb_gemm.log | cut -c1-170
MMG_GEMM_ARES=0 timeout 300 python scripts/kernel_bench.py --only gemm 2>&1 | grep "qkv\|ff1" | cut -c1-170
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log > $OUT/bench.json; pr $OUT/bench.json ares
( MMG_GEMM_ARES=0 timeout 600 python bench.py --no-extras ) > $OUT/bench_noares.log 2>&1; grep "^{" $OUT/bench_noares.log > $OUT/bench_noares.json; pr $OUT/bench_noares.json noares
MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace_gemm.log 2>&1; echo "trace exit $?"; grep -A11 "^--- qkv\|^--- geglu" $OUT/trace_gemm.log | cut -c1-120
