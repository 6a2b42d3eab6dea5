#!/bin/bash
# round 2, call B: fused logits + sampling tail — parity, microbench, end-to-end
OUT=${1:-gpurun_out/r2b}
mkdir -p $OUT
run() { local name=$1; local to=$2; shift 2; timeout -k 10 $to python -m pytest "$@" -q -rP -m gpu --no-header -p no:cacheprovider > $OUT/$name.log 2>&1; echo "$name exit $?: $(tail -1 $OUT/$name.log)"; }
run fused_tail 600 tests/test_gpu_fused_tail.py
grep -h "^FAILED\|^ERROR\|fallback rows" $OUT/fused_tail.log | head -30
run full_config 600 tests/test_gpu_full_config.py
grep -h "^FAILED\|^ERROR\|flip rate" $OUT/full_config.log | head
run models 600 tests/test_gpu_models.py
run kernels_sample 300 tests/test_gpu_kernels.py -k "sample or lfq or vq"
timeout 300 python scripts/kernel_bench.py --only fused > $OUT/kernel_bench_fused.log 2>&1; cut -c1-200 $OUT/kernel_bench_fused.log
( timeout 600 python bench.py --no-extras ) > $OUT/bench_fused.log 2>&1; echo "bench fused exit $?"; grep "^{" $OUT/bench_fused.log | cut -c1-1800
( MMG_FUSED_TAIL=0 timeout 600 python bench.py --no-extras ) > $OUT/bench_unfused.log 2>&1; echo "bench unfused exit $?"; grep "^{" $OUT/bench_unfused.log | cut -c1-300
