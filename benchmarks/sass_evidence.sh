#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove Blackwell-native code (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
# LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync.
cuobjdump -sass "${1:-office_person_detection_vit_b200/libopd_b200.so}" 2>/dev/null | awk '
/Function :/ { fn=$3; sub(/^_ZN[0-9a-z_A-Z]*GLOBAL__N__[0-9a-f]*_[0-9]*_/, "", fn); next }
{ for (i = 1; i <= NF; i++) if ($i ~ /^(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|HMMA|FFMA2|FADD2|SYNCS\.ARRIVE|SYNCS\.PHASECHK)/) { split($i, a, ";"); c[fn "  " a[1]]++ } }
END { for (k in c) printf "%5d  %s\n", c[k], k }' | sort -k2
