#!/bin/bash
# evidence only (the parity suite ran in gpu_final.sh): bench lines, ncu launch lists, condensed full captures
OUT=${1:-gpurun_out/final}
mkdir -p $OUT
( time timeout 900 python bench.py ) > $OUT/bench_default.log 2>&1; echo "bench default exit $?"
( time timeout 900 python bench.py --impl reference ) > $OUT/bench_reference.log 2>&1; echo "bench reference exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b8.csv python scripts/profile_step.py 8 > $OUT/ncu_launches8.log 2>&1; echo "ncu launches b8 exit $?"
timeout 600 ncu --set full --clock-control none -k regex:logits_sample -s 2 -c 1 -f -o /tmp/prof_sample python scripts/kernel_bench.py --only sample --iters 1 > $OUT/ncu_sample.log 2>&1; echo "ncu sample exit $?"
python scripts/ncu_summary.py /tmp/prof_sample.ncu-rep > $OUT/ncu_logits_sample.txt 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tc_gemm -c 40 -f -o /tmp/prof_gemm python scripts/kernel_bench.py --only gemm --iters 1 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
python scripts/ncu_summary.py /tmp/prof_gemm.ncu-rep > $OUT/ncu_gemm_shapes.txt 2>&1
timeout 300 python scripts/kernel_bench.py > $OUT/kernel_bench.log 2>&1
ls -la $OUT
