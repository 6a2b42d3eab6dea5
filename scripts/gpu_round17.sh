#!/bin/bash
OUT=${1:-gpurun_out/r1w}
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "linear" > $OUT/k.log 2>&1; echo "kernels exit $?: $(tail -1 $OUT/k.log)"; grep -h "^FAILED\|^ERROR\|Error" $OUT/k.log | head
timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb.log 2>&1; cut -c1-150 $OUT/kb.log | grep "logits"
MMG_GEMM_TSTORE=0 timeout 300 python scripts/kernel_bench.py --only gemm > $OUT/kb_staged.log 2>&1; cut -c1-150 $OUT/kb_staged.log | grep "logits"
timeout 900 python -m pytest tests/test_gpu_models.py -q -m gpu --no-header -p no:cacheprovider > $OUT/models.log 2>&1; echo "models exit $?: $(tail -1 $OUT/models.log)"; grep -h "^FAILED\|^ERROR" $OUT/models.log | head
timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench.log
MMG_GEMM_TSTORE=0 timeout 600 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_staged.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_staged.log; grep -o '"mmg_linear": [0-9.]*' $OUT/bench_staged.log
echo "== b=8: default / RED off"
timeout 600 python bench.py --steps 8 --no-cpu-baseline --global-batch 8 > $OUT/bench_b8.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_b8.log
MMG_GEMM_RED=0 timeout 600 python bench.py --steps 8 --no-cpu-baseline --global-batch 8 > $OUT/bench_b8_nored.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_b8_nored.log
MMG_GEMM_NFAST=0 timeout 600 python bench.py --steps 8 --no-cpu-baseline --global-batch 8 > $OUT/bench_b8_mfast.log 2>&1; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_b8_mfast.log
