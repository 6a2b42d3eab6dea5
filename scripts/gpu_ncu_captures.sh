#!/bin/bash
# round 2, call R: ncu --set full captures (text summaries + raw csv come back): finisher, LFQ lookup GEMM, attention, the block GEMMs
OUT=${1:-gpurun_out/r2r}
mkdir -p $OUT
cap() { # name, kernel regex, count, kernel_bench selector
  timeout 600 ncu --set full --clock-control none -k regex:$2 -c $3 -f -o /tmp/prof_$1 python scripts/kernel_bench.py --only $4 --iters 1 > $OUT/ncu_$1.log 2>&1; echo "ncu $1 exit $?"
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $OUT/ncu_$1_raw.csv 2>/dev/null
  python scripts/ncu_stalls.py /tmp/prof_$1.ncu-rep 12 > $OUT/ncu_$1_stalls.txt 2>&1; cut -c1-400 $OUT/ncu_$1_stalls.txt
}
cap finish logits_finish 3 fused
cap logits tc_logits 3 fused
cap vq "tc_gemm_kernel<64" 3 vq
cap attn attention_tc 6 attn
cap gemm tc_gemm 21 gemm
du -sh $OUT
