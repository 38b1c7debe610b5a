#!/bin/bash
# SASS evidence of the built library (no GPU needed): counts of the tcgen05 / TMEM / TMA mnemonics per kernel family.
# Usage: bash tools/sass_summary.sh > profiles/r02_sass_summary.txt
LIB=i-dccrn-vae_b200/libidv_b200.so
echo "# cuobjdump -sass $LIB (sm_100a), built $(date -u +%Y-%m-%dT%H:%MZ) from $(git rev-parse --short HEAD)"
echo "# mnemonic counts over the whole library"
cuobjdump -sass $LIB > /tmp/idv_sass.txt
for m in UTCHMMA "UTCHMMA.2CTA" UTCBAR LDTM UTMALDG UTMASTG UTMAPF SYNCS "ATOM\|RED" ; do
  printf "%-14s %s\n" "$m" "$(grep -c "$m" /tmp/idv_sass.txt)"
done
echo
echo "# per kernel (Function : name): UTCHMMA / UTCHMMA.2CTA / LDTM / UTMALDG / UTMASTG"
awk '/Function :/ {name=$3} /UTCHMMA/ {a[name]++} /UTCHMMA.2CTA/ {b[name]++} /LDTM/ {c[name]++} /UTMALDG/ {d[name]++} /UTMASTG/ {e[name]++}
     END {for (n in a) printf "%-110s %4d %4d %4d %4d %4d\n", n, a[n], b[n]+0, c[n]+0, d[n]+0, e[n]+0}' /tmp/idv_sass.txt | sort
