#!/bin/bash
# SASS evidence that the hot path is Blackwell-native (B200_PROFILING.md "What proves a Blackwell-native kernel"): counts of
# the tcgen05 / TMEM / TMA mnemonics in the built library, per kernel family. Writes profiles/sass_opcount.txt.
cd "$(dirname "$0")/.."
SO=visual-rag-toolkit_b200/visual_rag_b200/libvrag_b200.so
OUT=${1:-profiles/sass_opcount.txt}
cuobjdump -sass $SO > /tmp/vrag_sass.txt
{
  echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%dT%H:%MZ), $(git rev-parse --short HEAD))"
  echo "# mnemonic counts over the whole library"
  for m in UTCHMMA UTCQMMA LDTM STTM UTMALDG UTMASTG UBLKCP UTCBAR SYNCS HMMA HGMMA; do
    printf "%-10s %6d\n" $m $(grep -c "\b$m" /tmp/vrag_sass.txt)
  done
  echo
  echo "# per scan-kernel instantiation: UTCHMMA (tcgen05.mma) / LDTM (tcgen05.ld) / UTMALDG (cp.async.bulk.tensor)"
  awk '/Function : /{name=$3} /UTCHMMA/{a[name]++} /LDTM/{b[name]++} /UTMALDG/{c[name]++} END{for(n in a) printf "%4d %4d %4d  %s\n", a[n], b[n], c[n], n}' /tmp/vrag_sass.txt | sort -k4 | sed 's/_ZN4vrag18maxsim_scan_kernel/maxsim_scan_kernel/'
} > $OUT
cat $OUT | head -40
