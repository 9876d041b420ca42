#!/bin/bash
# Trimmed SASS listing of the shipped library: the TMA tile load, the mbarrier wait, the dp4a window loops, the
# staging copy and the 16-byte flush of kf_fused<9,0>.   usage: profiles/sass_excerpt.sh > profiles/r2_sass_excerpt.txt
LIB=${1:-sickle_b200/libsickle_b200.so}
TMP=$(mktemp -d)
( cd $TMP && cuobjdump -xelf all $OLDPWD/$LIB > /dev/null 2>&1 )
CUBIN=$(ls $TMP/*.cubin | head -1)
nvdisasm --print-line-info $CUBIN 2>/dev/null | awk '
  /\.text\._ZN2sk8kf_fusedILi9ELi0E/ {on=1} /\.text\./ && !/kf_fusedILi9ELi0E/ {on=0}
  on {print}' > $TMP/k.txt
echo "# kf_fused<9,0> of $LIB: $(grep -cE '^\s+/\*[0-9a-f]{4,6}\*/' $TMP/k.txt) SASS instructions"
echo "# mnemonic histogram (top 30)"
grep -E '^\s+/\*[0-9a-f]{4,6}\*/' $TMP/k.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+ )?//' | awk '{print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn | head -30
for pat in "UBLKCP" "SYNCS" "IDP.4A" "STG.E.128|STG.E.EF.128|STG.E.128.STRONG" "LDS.128" "STS.128" "ATOMG|RED" "BAR.SYNC" "CCTL|PREFETCH|LDGDEPBAR"; do
  echo; echo "# --- lines with $pat (first 12, with the source line they come from)"
  awk -v pat="$pat" '/\/\/## File/ {src=$0} $0 ~ pat && /\/\*[0-9a-f]+\*\// {print src; print $0}' $TMP/k.txt | head -24 | cut -c1-170
done
rm -rf $TMP
