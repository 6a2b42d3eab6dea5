#!/bin/bash
# round 2, call P: parity suite (split attention, finisher v2, final_embed), quick bench, per-tile GEMM timeline (pair-aware), fused-tail / vq microbenchmarks, launch list
OUT=${1:-gpurun_out/r2p}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
grep -h "C3 full-config\|max |err|" $OUT/*.log | cut -c1-300
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log > $OUT/bench.json; python -c "
import json
d=json.load(open('$OUT/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"
MMG_LIB=scripts/_build/libmmg_trace.so timeout 300 python scripts/trace_gemm.py > $OUT/trace_gemm.log 2>&1; echo "trace exit $?"; cat $OUT/trace_gemm.log | cut -c1-120
timeout 300 python scripts/kernel_bench.py --only fused,vq > $OUT/kb_fused_vq.log 2>&1; cut -c1-200 $OUT/kb_fused_vq.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches b64 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b64.csv > $OUT/launches_b64.txt; head -24 $OUT/launches_b64.txt | cut -c1-150
du -sh $OUT
