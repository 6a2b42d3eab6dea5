#!/bin/bash
# round 2, call C: fused tail after the gather / split fixes — parity, per-kernel breakdown (ncu launch list), end-to-end
OUT=${1:-gpurun_out/r2h}
mkdir -p $OUT
run() { local name=$1; local to=$2; shift 2; timeout -k 10 $to python -m pytest "$@" -q -rP -m gpu --no-header -p no:cacheprovider > $OUT/$name.log 2>&1; echo "$name exit $?: $(tail -1 $OUT/$name.log)"; }
run fused_tail 600 tests/test_gpu_fused_tail.py
grep -h "^FAILED\|^ERROR" $OUT/fused_tail.log | head -30
timeout 300 python scripts/kernel_bench.py --only fused > $OUT/kernel_bench_fused.log 2>&1; cut -c1-200 $OUT/kernel_bench_fused.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/fused_launches.csv python scripts/kernel_bench.py --only fused --iters 1 > $OUT/ncu_fused.log 2>&1; echo "ncu exit $?"
python scripts/ncu_traffic.py $OUT/fused_launches.csv > $OUT/fused_launches.txt; cat $OUT/fused_launches.txt | cut -c1-200
( timeout 600 python bench.py --no-extras ) > $OUT/bench_fused.log 2>&1; echo "bench fused exit $?"; grep "^{" $OUT/bench_fused.log | cut -c1-2500
