#!/bin/bash
# round 2, call S: parity suite, finisher register-budget A/B, VQ lookup with / without programmatic dependent launch, compute-sanitizer memcheck + racecheck
OUT=${1:-gpurun_out/r2s}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
for v in 4 6; do ( MMG_FINISH_MINB=$v timeout 300 python bench.py --no-extras ) > $OUT/bench_minb$v.log 2>&1; grep "^{" $OUT/bench_minb$v.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('minb $v', d['value'], d['ms_per_step'], d['roofline']['by_entry_point_ms'].get('mmg_logits_fused'))"; done
timeout 300 python scripts/vq_bench.py > $OUT/vq.log 2>&1; tail -1 $OUT/vq.log | cut -c1-600
MMG_PDL=1 timeout 300 python scripts/vq_bench.py > $OUT/vq_pdl.log 2>&1; tail -1 $OUT/vq_pdl.log | cut -c1-600
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_tail.py -q -m gpu -x --no-header -p no:cacheprovider -k "linear or attention or sample or fused_equals or conv2d or lfq or split3" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?: $(tail -3 $OUT/sanitizer_memcheck.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fused_tail.py -q -m gpu -x --no-header -p no:cacheprovider -k "test_linear_store or test_linear_residual_inplace or test_linear_qkv_epilogue_tma_tiles or test_linear_geglu_lnfold_pair or (test_attention and 256-257) or (test_fused_equals_materialised_philox and 300)" > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?: $(tail -3 $OUT/sanitizer_racecheck.log | tr '\n' ' ')"
du -sh $OUT
