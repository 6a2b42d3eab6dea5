#!/bin/bash
# One GPU session: parity suite, kernel microbenchmarks, bench line, ncu launch list + full captures of the top kernels.
OUT=${1:-gpurun_out/r1c}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
timeout 600 python scripts/kernel_bench.py > $OUT/kernel_bench.log 2>&1; echo "kernel_bench exit $?"
timeout 900 python bench.py --steps 5 > $OUT/bench.log 2>&1; echo "bench exit $?"; tail -c 2500 $OUT/bench.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_step.py 16 > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
for k in "sample:logits_sample_kernel" "ln:layernorm_vec_kernel" "attn:attention_tc_kernel" "gemm:tc_gemm_kernel"; do
  only=${k%%:*}; kern=${k##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 2 -f -o $OUT/prof_$only python scripts/kernel_bench.py --only $only --iters 1 > $OUT/ncu_$only.log 2>&1; echo "ncu $only exit $?"
done
ls -la $OUT
