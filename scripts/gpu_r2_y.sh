#!/bin/bash
# round 2, call Y: L2 persistence window for the residual stream A/B, conv-transpose RGB as CTA pairs A/B, smoke()
OUT=${1:-gpurun_out/r2y}
mkdir -p $OUT
pr() { python -c "
import json
d=json.load(open('$1')); print('$2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_entry_point_ms'], d['roofline']['frac'])"; }
( timeout 600 python bench.py --no-extras ) > $OUT/bench.log 2>&1; echo "bench exit $?"; grep "^{" $OUT/bench.log > $OUT/bench.json; pr $OUT/bench.json l2persist
( MMG_L2_PERSIST=0 timeout 600 python bench.py --no-extras ) > $OUT/bench_nopersist.log 2>&1; grep "^{" $OUT/bench_nopersist.log > $OUT/bench_nopersist.json; pr $OUT/bench_nopersist.json nopersist
( MMG_CONVT_RGB_PAIR=1 timeout 600 python bench.py --no-extras ) > $OUT/bench_rgbpair.log 2>&1; grep "^{" $OUT/bench_rgbpair.log > $OUT/bench_rgbpair.json; pr $OUT/bench_rgbpair.json rgbpair
MMG_CONVT_RGB_PAIR=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_full_config.py -q -m gpu --no-header -p no:cacheprovider -k "convt or conv_transpose" > $OUT/rgbpair_tests.log 2>&1; echo "rgb pair tests: $(tail -1 $OUT/rgbpair_tests.log)"
( timeout 300 python bench.py --no-extras --global-batch 8 ) > $OUT/bench_b8.log 2>&1; grep "^{" $OUT/bench_b8.log | cut -c1-200
( MMG_L2_PERSIST=0 timeout 300 python bench.py --no-extras --global-batch 8 ) > $OUT/bench_b8_np.log 2>&1; grep "^{" $OUT/bench_b8_np.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -4 $OUT/smoke.log
bash scripts/gpu_tests.sh $OUT
grep -h "^FAILED\|^ERROR" $OUT/*.log | head -30
