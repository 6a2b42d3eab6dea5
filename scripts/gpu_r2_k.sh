#!/bin/bash
OUT=${1:-gpurun_out/r2k}
mkdir -p $OUT
run() { local name=$1; local to=$2; shift 2; timeout -k 10 $to python -m pytest "$@" -q -rP -m gpu --no-header -p no:cacheprovider > $OUT/$name.log 2>&1; echo "$name exit $?: $(tail -1 $OUT/$name.log)"; }
run fused_tail 600 tests/test_gpu_fused_tail.py
run kernels 600 tests/test_gpu_kernels.py
run models 600 tests/test_gpu_models.py
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/fused_launches.csv python scripts/kernel_bench.py --only fused --iters 1 > $OUT/ncu_fused.log 2>&1; echo "ncu exit $?"
python scripts/ncu_traffic.py $OUT/fused_launches.csv > $OUT/fused_launches.txt; head -8 $OUT/fused_launches.txt | cut -c1-200
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"
