#!/bin/bash
# What bounds the fused logits GEMM: ncu time + tensor-pipe activity of tc_logits_kernel at 16 384 / 10 240 rows with the full epilogue (MMG_LOGITS_DBG unset),
# without the candidate lists (1) and with an epilogue that only drains TMEM (2).  Text only.
OUT=${1:-gpurun_out/fused_bounds}
mkdir -p $OUT
for d in 0 1 2; do
  MMG_LOGITS_DBG=$d timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_write.sum --clock-control none -k regex:tc_logits -c 4 --csv --log-file $OUT/dbg$d.csv python scripts/kernel_bench.py --only fused --iters 1 > $OUT/dbg$d.log 2>&1
  echo "MMG_LOGITS_DBG=$d"; python - <<PY
import csv
rows=list(csv.reader(open("$OUT/dbg$d.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]; h=rows[hi]
per={}
for r in rows[hi+1:]:
    if len(r)>h.index("Metric Value"): per.setdefault(r[h.index("ID")],{})[r[h.index("Metric Name")]]=(r[h.index("Metric Value")],r[h.index("Metric Unit")])
for k,v in per.items(): print("  launch",k,{m:" ".join(x) for m,x in v.items()})
PY
done
