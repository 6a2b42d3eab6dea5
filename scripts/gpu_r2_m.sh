#!/bin/bash
OUT=${1:-gpurun_out/r2m}
mkdir -p $OUT
run() { local name=$1; local to=$2; shift 2; timeout -k 10 $to python -m pytest "$@" -q -rP -m gpu --no-header -p no:cacheprovider > $OUT/$name.log 2>&1; echo "$name exit $?: $(tail -1 $OUT/$name.log)"; }
run kernels 600 tests/test_gpu_kernels.py
run models 600 tests/test_gpu_models.py
grep -h "^FAILED\|^ERROR" $OUT/*.log | head
timeout 300 python scripts/kernel_bench.py --only ln > $OUT/kb_ln.log 2>&1; cut -c1-170 $OUT/kb_ln.log
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -4 $OUT/smoke.log
