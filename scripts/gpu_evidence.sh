#!/bin/bash
# round 2, final evidence of the last build: parity suite, default bench with extras, launch lists (b64 with DRAM bytes, b8), sanitizer
OUT=${1:-gpurun_out/r2final}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
grep -h "explicit codebook\|C3 full-config" $OUT/*.log | cut -c1-260
( time timeout 900 python bench.py ) > $OUT/bench_default.log 2>&1; echo "bench default exit $?"; grep "^{" $OUT/bench_default.log > $OUT/bench_default.json; python -c "
import json
d=json.load(open('$OUT/bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'], d['roofline']['achieved']); print(d.get('hbm_kernels')); print(d.get('parity'))"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches b64 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b64.csv $OUT/traffic.json > $OUT/launches_b64.txt; head -24 $OUT/launches_b64.txt | cut -c1-150
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b8.csv python scripts/profile_step.py 8 > $OUT/ncu_launches8.log 2>&1; echo "ncu launches b8 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b8.csv > $OUT/launches_b8.txt; head -22 $OUT/launches_b8.txt | cut -c1-150
( timeout 300 python bench.py --no-extras --global-batch 8 ) > $OUT/bench_b8.log 2>&1; grep "^{" $OUT/bench_b8.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_tail.py -q -m gpu -x --no-header -p no:cacheprovider -k "linear or attention or sample or fused_equals or conv2d or lfq or split3" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?: $(tail -3 $OUT/sanitizer_memcheck.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_tail.py -q -m gpu -x --no-header -p no:cacheprovider -k "test_linear_store or test_linear_residual_inplace or test_linear_qkv_epilogue_tma_tiles or test_linear_geglu_lnfold_pair or (test_attention and 256-257) or (test_attention and 256-33) or (test_fused_equals_materialised_philox and 300)" > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?: $(tail -3 $OUT/sanitizer_racecheck.log | tr '\n' ' ')"
du -sh $OUT
