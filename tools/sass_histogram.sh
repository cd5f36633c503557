#!/bin/bash
# SASS mnemonic histogram of the in-tree library: the instructions that prove the Blackwell-native path
# (UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, SYNCS = mbarrier).
so=${1:-musev_b200/_lib/libmusevb200.so}
echo "# cuobjdump -sass $so | mnemonic histogram ($(date -u +%Y-%m-%dT%H:%MZ))"
cuobjdump -sass "$so" | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?[A-Z0-9_.]+" | awk '{print $NF}' | sed -E 's/\.(.*)//' > /tmp/_sass_ops.txt
echo "## tensor / TMA / TMEM / barrier instructions (full mnemonics)"
cuobjdump -sass "$so" | grep -oE "(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|UTCATOMSWS|UTMACCTL|UTMACMDFLUSH|SYNCS|UBLKCP|FENCE\.VIEW|HMMA|MUFU|F2FP|FMNMX3|FFMA2|FADD2|FMUL2|LDSM|ELECT|UCGABAR|BAR)[A-Z0-9_.]*" | sort | uniq -c | sort -rn
echo "## all base mnemonics"
sort /tmp/_sass_ops.txt | uniq -c | sort -rn | head -60
echo "## kernels"
cuobjdump -sass "$so" | grep -oE "Function : [A-Za-z0-9_]+" | sed 's/Function : //' | c++filt | sed -E 's/\(.*//' | sort | uniq -c | sort -rn | head -40
