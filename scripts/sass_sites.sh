#!/bin/bash
# Evidence that the hot kernels are tcgen05 / TMA code: count and list the Blackwell SASS mnemonics per kernel of libmmg.so
# (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce, UTCBAR = tcgen05.commit).
# usage: scripts/sass_sites.sh > profiles/rN_sass_sites.txt
LIB=muse_maskgit_pytorch_b200/libmmg.so
cuobjdump -sass $LIB | awk '
  /Function :/ { fn=$3; next }
  /UTCHMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTCBAR|UTMAPF|SYNCS/ {
    m=""; if ($0 ~ /UTCHMMA/) m="UTCHMMA"; else if ($0 ~ /LDTM/) m="LDTM"; else if ($0 ~ /STTM/) m="STTM"; else if ($0 ~ /UTMALDG/) m="UTMALDG";
    else if ($0 ~ /UTMASTG/) m="UTMASTG"; else if ($0 ~ /UTMAREDG/) m="UTMAREDG"; else if ($0 ~ /UTCBAR/) m="UTCBAR"; else if ($0 ~ /UTMAPF/) m="UTMAPF"; else m="SYNCS";
    c[fn" "m]++ }
  END { for (k in c) print k, c[k] }' | sort | c++filt | awk '{n=$NF; $NF=""; print n, $0}' | sort -k2
echo
echo "--- first tcgen05.mma / TMA sites of the GEMM, attention and fused-logits kernels (cuobjdump -sass, address : instruction)"
cuobjdump -sass $LIB | awk '/Function :/ {fn=$3; shown=0} /UTCHMMA|UTMALDG|UTMAREDG|UTMASTG|LDTM/ { if (fn ~ /tc_gemm_kernelILi256ELb0ELb1ELi2E|attention_tc|tc_logits_kernelILb1E/ && shown < 12) { print fn ": " $0; shown++ } }' | c++filt | cut -c1-230
