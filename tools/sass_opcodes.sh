#!/bin/bash
# SASS evidence that the hot path is tcgen05 / TMEM / TMA (B200_PROFILING.md "What proves a Blackwell-native kernel"):
# per kernel of libduodiff_b200.so, the number of UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTCCP
# (tcgen05.cp), UTMALDG / UTMASTG (TMA load / store), HMMA (legacy mma.sync) instructions.
#   tools/sass_opcodes.sh > profiles/r02_sass_opcodes.txt
cd "$(dirname "$0")/.."
SO=duodiff_b200/libduodiff_b200.so
echo "# $(date -u +%Y-%m-%dT%H:%MZ)  cuobjdump -sass $SO  ($(python -c 'from duodiff_b200 import _lib; print(_lib.load().ddb_version().decode())'))"
echo "# git $(git rev-parse --short HEAD)  (+ working tree)"
printf "%-92s %8s %6s %6s %6s %8s %8s %6s\n" kernel UTCHMMA LDTM STTM UTCCP UTMALDG UTMASTG HMMA
cuobjdump -sass $SO | awk '
  /Function :/ { if (name != "") emit(); name=$3; a=b=c=d=e=f=g=0; next }
  /UTCHMMA/ {a++} /LDTM/ {b++} /STTM/ {c++} /UTCCP/ {d++} /UTMALDG/ {e++} /UTMASTG/ {f++} / HMMA/ {g++}
  function emit() { if (a+b+c+d+e+f+g > 0) printf "%s %d %d %d %d %d %d %d\n", name, a, b, c, d, e, f, g }
  END { emit() }' | while read n a b c d e f g; do
    printf "%-92s %8s %6s %6s %6s %8s %8s %6s\n" "$(echo $n | c++filt | cut -c1-90)" $a $b $c $d $e $f $g
  done | sort
