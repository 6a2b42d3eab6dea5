#!/bin/bash
# validation + evidence: parity suite, bench line, ncu launch list of one generate(), full ncu capture of the sampler and the logits GEMM
OUT=${1:-gpurun_out/r1q}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -20
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; echo "bench exit $?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"by_entry_point_ms": {[^}]*}' $OUT/bench.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logits_sample -s 2 -c 1 -f -o $OUT/prof_sample python scripts/kernel_bench.py --only sample --iters 1 > $OUT/ncu_sample.log 2>&1; echo "ncu sample exit $?"
timeout 300 python scripts/kernel_bench.py --only sample,attn,vq,ln > $OUT/kernel_bench.log 2>&1; cut -c1-150 $OUT/kernel_bench.log
