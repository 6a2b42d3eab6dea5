#!/bin/bash
# round 2: programmatic dependent launch and finisher register budget at 8 images per GPU / batch 64 on the final build
OUT=${1:-gpurun_out/r2pdl}
mkdir -p $OUT
b() { ( env $1 timeout 300 python bench.py --no-extras $2 ) > $OUT/$3.log 2>&1; grep "^{" $OUT/$3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$3', d['value'], d['ms_per_step'])"; }
b "X=1" "--global-batch 8" b8_default
b "MMG_PDL=1" "--global-batch 8" b8_pdl
b "MMG_FINISH_MINB=6" "--global-batch 8" b8_minb6
b "MMG_PDL=1 MMG_FINISH_MINB=6" "--global-batch 8" b8_pdl_minb6
b "X=1" "" b64_default
b "MMG_PDL=1" "" b64_pdl
