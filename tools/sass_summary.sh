#!/bin/bash
# Counts the Blackwell-only SASS mnemonics per kernel of the in-tree library (the PTX names never appear in SASS:
# tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG / UTMASTG, tcgen05.commit -> UTCBAR) and
# prints the instruction lines of the first occurrence of each, per kernel.   bash tools/sass_summary.sh > profiles/r2_sass_summary.txt
LIB="$(dirname "$0")/../jpd-se_b200/libjpdse_b200.so"
echo "# cuobjdump -sass $(basename "$LIB") ($(date -u +%Y-%m-%d)), sm_100a; counts of tcgen05 / TMEM / TMA instructions per kernel"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
  /Function :/ { fn=$3; next }
  { for (i = 1; i <= NF; ++i) if ($i ~ /^(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UBLKCP|UTMAPF|HMMA)/) { sub(/;$/, "", $i); cnt[fn "\t" $i]++; if (!((fn,$i) in first)) first[fn,$i]=$0 } }
  END { for (k in cnt) print cnt[k] "\t" k }' | sort -t$'\t' -k2,2 -k3,3 | awk -F'\t' '{ if ($2 != last) { print ""; print $2; last=$2 } printf "    %4d x %s\n", $1, $3 }'
echo
echo "# legacy tensor path (HMMA = mma.sync / wmma): $(cuobjdump -sass "$LIB" 2>/dev/null | grep -cE '(^|[^A-Z])HMMA' ) occurrences"
echo "# excerpt: the MMA issue loop of pair_conv3x3_kernel (tcgen05.mma.cta_group::2) and the first TMA load / TMEM read"
cuobjdump -sass "$LIB" 2>/dev/null | awk '/Function : .*pair_conv3x3_kernel/{on=1} /Function :/{ if (!/pair_conv3x3_kernel/) on=0 } on && /(UTCHMMA|UTMALDG|LDTM|UTCBAR)/' | head -24
