#!/bin/bash
# SASS evidence for the in-tree library: which tensor-core / TMA / async instructions the sm_100a build contains, per kernel.
# usage: scripts/sass_grep.sh > profiles/r2_sass_grep.txt
LIB=dsgan_b200/libdsgan_b200.so
S=$(mktemp)
cuobjdump -sass $LIB > $S
echo "cuobjdump -sass $LIB | grep -c <mnemonic>   (in-tree library, sm_100a only)"
for m in UTCHMMA UTMALDG UTMASTG UTCBAR LDTM "SYNCS.PHASECHK.TRANS64.TRYWAIT" "NANOSLEEP.SYNCS" "REDG.E.ADD.F32x4" "STG.E.ENL2.256" "LDG.E.ENL2.256" LDGSTS; do
  echo "$m $(grep -c "$m" $S)"
done
echo "HMMA.16816.F32.BF16 (mma.sync m16n8k16 of the depthwise / narrow-channel kernels; lines matching \"[^C]HMMA\") $(grep -c '[^C]HMMA' $S)"
echo "LDSM (ldmatrix) $(grep -c 'LDSM' $S)"
echo "HGMMA $(grep -c HGMMA $S)"
echo "FFMA2 $(grep -c FFMA2 $S)"
echo
echo "kernels containing UTCHMMA (tcgen05.mma):"
awk '/Function :/ {f=$3} /UTCHMMA/ {c[f]++} END {for (k in c) print c[k], k}' $S | sort -rn | cut -c1-150
echo
echo "kernels containing HMMA (mma.sync): depthwise Toeplitz kernels (dwconv_mma.cu) and narrow-channel convolutions (nm_conv.cu) only"
awk '/Function :/ {f=$3} /[^C]HMMA/ {c[f]++} END {for (k in c) print c[k], k}' $S | sort -rn | cut -c1-150
echo
echo "kernels containing UTMASTG (TMA store):"
awk '/Function :/ {f=$3} /UTMASTG/ {c[f]++} END {for (k in c) print c[k], k}' $S | sort -rn | cut -c1-150
rm -f $S
