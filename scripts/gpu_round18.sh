#!/bin/bash
OUT=${1:-gpurun_out/r1x}
mkdir -p $OUT
MMG_GEMM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "linear or conv" > $OUT/k_pair.log 2>&1; echo "kernels(pair) exit $?: $(tail -1 $OUT/k_pair.log)"; grep -h "^FAILED\|^ERROR\|Error" $OUT/k_pair.log | head
echo "== default"; timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb.log 2>&1; cut -c1-150 $OUT/kb.log
echo "== pair forced"; MMG_GEMM_PAIR=1 timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_pair.log 2>&1; cut -c1-150 $OUT/kb_pair.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench.log
MMG_GEMM_PAIR=1 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_pair.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_pair.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench_pair.log
