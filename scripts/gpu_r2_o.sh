#!/bin/bash
# round 2, call O: whole parity suite on the current build (fp32 products as 3-way bf16 splits on tcgen05, GEGLU epilogue through TMA tiles, cheaper GELU),
# default bench with extras, A/B of the GEGLU tile epilogue, launch lists b64 / b8 with DRAM bytes (text only comes back)
OUT=${1:-gpurun_out/r2o}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
( time timeout 900 python bench.py ) > $OUT/bench_default.log 2>&1; echo "bench default exit $?"; grep "^{" $OUT/bench_default.log > $OUT/bench_default.json; cut -c1-400 $OUT/bench_default.json
( MMG_GEMM_GEGLUT=0 timeout 300 python bench.py --no-extras ) > $OUT/bench_nogeglut.log 2>&1; grep "^{" $OUT/bench_nogeglut.log | cut -c1-200
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_gemm.log 2>&1; cut -c1-170 $OUT/kb_gemm.log
MMG_GEMM_GEGLUT=0 timeout 300 python scripts/kernel_bench.py --only gemm 2>&1 | grep geglu | cut -c1-170
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches b64 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b64.csv > $OUT/launches_b64.txt; head -24 $OUT/launches_b64.txt | cut -c1-150
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b8.csv python scripts/profile_step.py 8 > $OUT/ncu_launches8.log 2>&1; echo "ncu launches b8 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b8.csv > $OUT/launches_b8.txt; head -24 $OUT/launches_b8.txt | cut -c1-150
( timeout 300 python bench.py --no-extras --global-batch 8 ) > $OUT/bench_b8.log 2>&1; grep "^{" $OUT/bench_b8.log | cut -c1-300
du -sh $OUT
