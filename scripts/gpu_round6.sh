#!/bin/bash
OUT=${1:-gpurun_out/r1l}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention or critic" --no-header -p no:cacheprovider > $OUT/attn_tests.log 2>&1; echo "tests exit $?: $(tail -1 $OUT/attn_tests.log)"
grep -h "FAILED\|Error" $OUT/attn_tests.log | head
for kb in 64 128 160; do
  echo "== MMG_ATTN_KB=$kb"
  MMG_ATTN_KB=$kb timeout 300 python scripts/kernel_bench.py --only attention 2>&1 | cut -c1-150
  MMG_ATTN_KB=$kb timeout 600 python bench.py --steps 3 --no-cpu-baseline > $OUT/bench_kb$kb.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_kb$kb.log; grep -o '"mmg_attention": [0-9.]*' $OUT/bench_kb$kb.log
done
