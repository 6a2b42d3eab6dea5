#!/bin/bash
# round 2, call Q: parity suite (deterministic row statistics, streaming finisher), models suite repeated (flakiness check), default bench with extras, fused-tail microbenchmark
OUT=${1:-gpurun_out/r2q}
mkdir -p $OUT
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_models.py -q -m gpu --no-header -p no:cacheprovider -k "graph_equals_eager or decode_step_c_abi or deterministic or golden" > $OUT/models_rep$i.log 2>&1; echo "models rep $i: $(tail -1 $OUT/models_rep$i.log)"; done
( time timeout 900 python bench.py ) > $OUT/bench_default.log 2>&1; echo "bench default exit $?"; grep "^{" $OUT/bench_default.log > $OUT/bench_default.json; python -c "
import json
d=json.load(open('$OUT/bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac']); print(d.get('hbm_kernels')); print(d.get('parity')); print(d.get('gpu_eager_baseline'))"
timeout 300 python scripts/kernel_bench.py --only fused > $OUT/kb_fused.log 2>&1; cut -c1-200 $OUT/kb_fused.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_b64.csv python scripts/profile_step.py 64 > $OUT/ncu_launches.log 2>&1; echo "ncu launches b64 exit $?"
python scripts/ncu_traffic.py $OUT/launches_b64.csv > $OUT/launches_b64.txt; head -14 $OUT/launches_b64.txt | cut -c1-150
du -sh $OUT
