#!/bin/bash
# what bounds the fused logits kernel: mainloop alone (dbg 2), + softmax (dbg 1), full (dbg 0); A-resident vs streaming
OUT=${1:-gpurun_out/r2f}
mkdir -p $OUT
for ares in 1 0; do for dbg in 2 1 0; do
  MMG_LOGITS_ARES=$ares MMG_LOGITS_DBG=$dbg timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed.sum.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none -k regex:tc_logits --csv --log-file $OUT/lf_a${ares}_d${dbg}.csv python scripts/kernel_bench.py --only fused --iters 1 > $OUT/lf_a${ares}_d${dbg}.log 2>&1
  echo "== ares=$ares dbg=$dbg"; python - <<PY
import csv
rows=list(csv.reader(open("$OUT/lf_a${ares}_d${dbg}.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]; h=rows[hi]
ki,mi,vi,gi=h.index("Kernel Name"),h.index("Metric Name"),h.index("Metric Value"),h.index("Grid Size")
cur={}
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    cur.setdefault(r[0],{"g":r[gi]})[r[mi]]=r[vi]
for k,d in list(cur.items())[2::3][:4]:
    print(d["g"], {m.split("__")[1][:28]:v for m,v in d.items() if m!="g"})
PY
done; done
