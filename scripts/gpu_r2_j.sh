#!/bin/bash
# round 2, call J: whole parity suite on the pair-default build, default bench with extras, secondary configs, launch lists, compute-sanitizer
OUT=${1:-gpurun_out/r2j}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -20
( time timeout 900 python bench.py ) > $OUT/bench_default.log 2>&1; echo "bench default exit $?"; grep "^{" $OUT/bench_default.log | cut -c1-400
for c in C2 C4 C5; do ( timeout 600 python bench.py --config $c ) > $OUT/bench_$c.log 2>&1; echo "bench $c exit $?"; grep "^{" $OUT/bench_$c.log | cut -c1-260; done
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches b64 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b64.csv > $OUT/launches_b64.txt; head -16 $OUT/launches_b64.txt | cut -c1-150
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b8.csv python scripts/profile_step.py 8 > $OUT/ncu_launches8.log 2>&1; echo "ncu launches b8 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b8.csv > $OUT/launches_b8.txt; head -16 $OUT/launches_b8.txt | cut -c1-150
# compute-sanitizer on the hand-rolled mbarrier / TMEM / cluster kernels (subset: sanitizer slows kernels down by 10-100x)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_tail.py -q -m gpu -x --no-header -p no:cacheprovider -k "linear or attention or sample or fused_equals or conv2d or lfq" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?: $(tail -3 $OUT/sanitizer_memcheck.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_tail.py -q -m gpu -x --no-header -p no:cacheprovider -k "test_linear_store or test_linear_residual_inplace or test_linear_qkv_epilogue_tma_tiles or (test_attention and 256-257) or (test_fused_equals_materialised_philox and 300)" > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?: $(tail -3 $OUT/sanitizer_racecheck.log | tr '\n' ' ')"
