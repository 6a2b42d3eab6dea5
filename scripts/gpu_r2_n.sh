#!/bin/bash
# round 2, call N (re-entry baseline): whole parity suite, default bench with extras, launch lists b64 / b8 with DRAM bytes (text only comes back)
OUT=${1:-gpurun_out/r2n}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -20
( time timeout 900 python bench.py ) > $OUT/bench_default.log 2>&1; echo "bench default exit $?"; grep "^{" $OUT/bench_default.log > $OUT/bench_default.json; cut -c1-400 $OUT/bench_default.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches b64 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b64.csv > $OUT/launches_b64.txt; head -22 $OUT/launches_b64.txt | cut -c1-150
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b8.csv python scripts/profile_step.py 8 > $OUT/ncu_launches8.log 2>&1; echo "ncu launches b8 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b8.csv > $OUT/launches_b8.txt; head -22 $OUT/launches_b8.txt | cut -c1-150
( timeout 300 python bench.py --no-extras --global-batch 8 ) > $OUT/bench_b8.log 2>&1; grep "^{" $OUT/bench_b8.log | cut -c1-300
du -sh $OUT
