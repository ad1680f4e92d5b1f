#!/bin/bash
# SASS opcode histogram of the shipped library: proves which Blackwell paths the build really contains
# (tcgen05 = UTCHMMA/UTCQMMA, TMEM = LDTM/STTM, tcgen05.commit = UTCBAR, bulk TMA = UBLKCP, tensor TMA = UTMALDG,
# cp.async = LDGSTS, cluster DSMEM stores = ST*.*cluster / mapa = MAPA, REDUX = warp reductions).
# Usage: tools/sass_histogram.sh [lib] > profiles/rNN_sass_histogram.txt
LIB="${1:-kd_pointcloud_b200/libkdpc.so}"
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%dT%H:%MZ), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
TMP=$(mktemp)
cuobjdump -sass "$LIB" > "$TMP"
echo "# per-opcode counts, whole library"
for op in UTCHMMA UTCQMMA UTCBAR LDTM STTM UTCCP UBLKCP UTMALDG UTMASTG LDGSTS SYNCS MAPA REDUX ELECT UCGABAR_ARV MEMBAR FFMA HMMA IMMA SHFL LDS STS LDG STG ATOM RED BAR; do
  printf "%-12s %8d\n" "$op" "$(grep -c "^\s*/\*[0-9a-f]*\*/\s*\(@!\?U\?P[0-9T]* \)\?$op" "$TMP")"
done
echo "# per-kernel tcgen05 / TMA / cp.async counts"
awk '/Function :/ {fn=$3} /UTCHMMA/ {a[fn]++} /LDTM/ {b[fn]++} /UBLKCP/ {c[fn]++} /LDGSTS/ {d[fn]++} /UTMALDG/ {e[fn]++} END {for (f in a) printf "%s UTCHMMA=%d LDTM=%d UBLKCP=%d LDGSTS=%d UTMALDG=%d\n", f, a[f], b[f], c[f], d[f], e[f]}' "$TMP" | c++filt | sort
rm -f "$TMP"
