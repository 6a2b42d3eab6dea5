#!/bin/bash
OUT=${1:-gpurun_out/r2l}
mkdir -p $OUT
run() { local name=$1; local to=$2; shift 2; timeout -k 10 $to python -m pytest "$@" -q -rP -m gpu --no-header -p no:cacheprovider > $OUT/$name.log 2>&1; echo "$name exit $?: $(tail -1 $OUT/$name.log)"; }
run models 600 tests/test_gpu_models.py -k "non_default or pack_cache or any_topk"
timeout 900 ncu --set full --clock-control none -k regex:tc_gemm -c 30 -f -o $OUT/prof_gemm python scripts/kernel_bench.py --only gemm --iters 1 > $OUT/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
python scripts/ncu_summary.py $OUT/prof_gemm.ncu-rep 3 2 > $OUT/ncu_gemm_shapes.txt; cut -c1-330 $OUT/ncu_gemm_shapes.txt
timeout 600 ncu --set full --clock-control none -k regex:attention_tc -c 6 -f -o $OUT/prof_attn python scripts/kernel_bench.py --only attn --iters 1 > $OUT/ncu_attn.log 2>&1; echo "ncu attn exit $?"
python scripts/ncu_summary.py $OUT/prof_attn.ncu-rep > $OUT/ncu_attn.txt; cut -c1-330 $OUT/ncu_attn.txt
timeout 600 ncu --set full --clock-control none -k regex:layernorm_vec -c 3 -f -o $OUT/prof_ln python scripts/kernel_bench.py --only ln --iters 1 > $OUT/ncu_ln.log 2>&1; echo "ncu ln exit $?"
python scripts/ncu_summary.py $OUT/prof_ln.ncu-rep > $OUT/ncu_ln.txt; cut -c1-330 $OUT/ncu_ln.txt
