#!/bin/bash
# Per-kernel SASS evidence for profiles/: counts of the tcgen05 / TMEM / TMA mnemonics, registers and spills.
# usage: tools/sass_summary.sh [lib.so] > profiles/rNN_sass_summary.txt
LIB=${1:-speech_diarization_b200/csrc/libsd_b200.so}
echo "# cuobjdump -sass $LIB (sm_100a) — per-kernel mnemonic counts"
echo "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,"
echo "# SYNCS = mbarrier, F2FP.SATFINITE = saturating f32->f16x2 conversion, STL/LDL = spills"
cuobjdump -sass "$LIB" | awk '
/Function : /{ if (name!="") flush(); name=$3; delete c; n=0; next }
/^[ \t]+\/\*[0-9a-f]+\*\//{ n++;
  if ($0 ~ /UTCHMMA\.2CTA/) c["UTCHMMA.2CTA"]++; else if ($0 ~ /UTCHMMA/) c["UTCHMMA"]++;
  if ($0 ~ /LDTM/) c["LDTM"]++;
  if ($0 ~ /UTMALDG/) { c["UTMALDG"]++; if ($0 ~ /MULTICAST/) c["UTMALDG.MULTICAST"]++ }
  if ($0 ~ /UTMASTG/) c["UTMASTG"]++;
  if ($0 ~ /UTCBAR/) c["UTCBAR"]++;
  if ($0 ~ /SYNCS/) c["SYNCS"]++;
  if ($0 ~ /F2FP.*SATFINITE/) c["F2FP.SATFINITE"]++;
  if ($0 ~ / STL/) c["STL"]++; if ($0 ~ / LDL/) c["LDL"]++;
  if ($0 ~ /SHFL/) c["SHFL"]++; if ($0 ~ /MUFU/) c["MUFU"]++;
  if ($0 ~ /HMMA|IMMA/ && $0 !~ /UTCHMMA/) c["legacy-MMA"]++;
}
function flush(   k, s) { s=""; for (k in c) s = s sprintf(" %s=%d", k, c[k]); printf "%-6d instr %s :%s\n", n, name, s }
END{ flush() }' | c++filt | sort -k3
echo
echo "# cuobjdump -res-usage"
cuobjdump -res-usage "$LIB" 2>/dev/null | c++filt | grep -A1 "Function" | grep -v "^--" | paste - - | sed 's/Fatbin elf code://' | awk '{$1=$1};1'
