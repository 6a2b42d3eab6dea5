#!/bin/bash
# round 2, call W: parity suite (attention liveness ballot / unrounded denominator), attention microbenchmark, bench, secondary configs
OUT=${1:-gpurun_out/r2w}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
grep -h "C3 full-config" $OUT/full_config.log | cut -c1-250
timeout 300 python scripts/kernel_bench.py --only attn > $OUT/kb_attn.log 2>&1; cut -c1-170 $OUT/kb_attn.log
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log > $OUT/bench.json; python -c "
import json
d=json.load(open('$OUT/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"
for c in C2 C4 C5; do ( timeout 600 python bench.py --config $c ) > $OUT/bench_$c.log 2>&1; echo "bench $c exit $?"; grep "^{" $OUT/bench_$c.log > $OUT/bench_$c.json; cut -c1-330 $OUT/bench_$c.json; done
