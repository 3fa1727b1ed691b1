#!/bin/bash
# SASS evidence of the Blackwell-native pieces: TMA loads, mbarrier waits, packed FMAs, DSMEM async stores.
#   tools/sass_counts.sh > profiles/r02_sass_counts.txt
set -e
LIB="$(dirname "$0")/../fast-poisson-image-editing_b200/fpie_b200/libfpie_b200.so"
SASS="$(mktemp)"
trap 'rm -f "$SASS"' EXIT
cuobjdump -sass "$LIB" > "$SASS"
echo "cuobjdump -sass libfpie_b200.so (sm_100a); instruction mnemonics per kernel"
awk '
  /Function :/ { fn=$3; sub(/^_ZN4fpie[0-9]*/, "", fn); next }
  /\/\*[0-9a-f]+\*\// {
    op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op)
    if (op ~ /^UTMALDG/) t[fn]++
    if (op ~ /^UTMAPF|^UTMACCTL/) pf[fn]++
    if (op ~ /^SYNCS/) s[fn]++
    if (op ~ /^FFMA2/) f2[fn]++
    else if (op ~ /^FFMA/) f[fn]++
    if (op ~ /^SHFL/) sh[fn]++
    if (op ~ /^STAS|^ST\.E.*ASYNC|^STS.*ASYNC/) sa[fn]++
    if (op ~ /^UCGABAR|^CGABAR|^MAPA|^UMAPA/) cg[fn]++
    n[fn]++
  }
  END {
    printf "%-78s %6s %7s %6s %6s %6s %6s %6s %6s\n", "kernel", "insts", "UTMALDG", "SYNCS", "FFMA", "FFMA2", "SHFL", "STAS", "CGA"
    for (k in n) if (t[k] + f2[k] + sa[k] + cg[k] > 0 || k ~ /sweep|patch/)
      printf "%-78s %6d %7d %6d %6d %6d %6d %6d %6d\n", substr(k, 1, 78), n[k], t[k], s[k], f[k], f2[k], sh[k], sa[k], cg[k]
  }' "$SASS" | sort
