#!/bin/bash
# Round-end validation + evidence on one B200: parity suite, the default bench line (with cpu_baseline), the reference arm,
# ncu launch list of one eager generate(), full ncu captures of the sampler and of the GEMM shapes.
OUT=${1:-gpurun_out/final}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -20
( time timeout 900 python bench.py ) > $OUT/bench_default.log 2>&1; echo "bench default exit $?"; tail -n 5 $OUT/bench_default.log | cut -c1-600
( time timeout 900 python bench.py --impl reference ) > $OUT/bench_reference.log 2>&1; echo "bench reference exit $?"; tail -n 5 $OUT/bench_reference.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b8.csv python scripts/profile_step.py 8 > $OUT/ncu_launches8.log 2>&1; echo "ncu launches b8 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logits_sample -s 2 -c 1 -f -o $OUT/prof_sample python scripts/kernel_bench.py --only sample --iters 1 > $OUT/ncu_sample.log 2>&1; echo "ncu sample exit $?"
timeout 900 ncu --set full --clock-control none -k regex:tc_gemm -c 30 -f -o $OUT/prof_gemm python scripts/kernel_bench.py --only gemm --iters 1 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 300 python scripts/kernel_bench.py > $OUT/kernel_bench.log 2>&1; cut -c1-150 $OUT/kernel_bench.log
