#!/bin/bash
# round 2, call Z: attention with the null key outside the MMA blocks: parity suite, microbenchmark, bench
OUT=${1:-gpurun_out/r2z}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
grep -h "C3 full-config\|C4 super" $OUT/full_config.log | cut -c1-250
timeout 300 python scripts/kernel_bench.py --only attn > $OUT/kb_attn.log 2>&1; cut -c1-170 $OUT/kb_attn.log
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log > $OUT/bench.json; python -c "
import json
d=json.load(open('$OUT/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"
( timeout 300 python bench.py --no-extras --global-batch 8 ) > $OUT/bench_b8.log 2>&1; grep "^{" $OUT/bench_b8.log | cut -c1-200
