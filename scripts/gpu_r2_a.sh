#!/bin/bash
# round 2, call A: parity suite incl. the full-config tests, the reworked bench (all arms / configs), launch list with DRAM bytes
OUT=${1:-gpurun_out/r2a}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -20
grep -h "C2 full\|C3 full\|C4 super\|C5 \|flip rate" $OUT/full_config.log | cut -c1-400
( time timeout 900 python bench.py ) > $OUT/bench_default.log 2>&1; echo "bench default exit $?"; tail -n 4 $OUT/bench_default.log | cut -c1-3000
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_reference.log 2>&1; echo "bench reference exit $?"; tail -n 4 $OUT/bench_reference.log | cut -c1-800
for c in C2 C4 C5; do ( time timeout 600 python bench.py --config $c ) > $OUT/bench_$c.log 2>&1; echo "bench $c exit $?"; tail -n 4 $OUT/bench_$c.log | cut -c1-700; done
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 300 python scripts/kernel_bench.py --only vq,sample > $OUT/kernel_bench.log 2>&1; cut -c1-170 $OUT/kernel_bench.log
