#!/bin/bash
# opcode histogram of the shipped library (run here, no GPU needed): the tcgen05 / TMA / bulk-reduce mnemonics of
# /opt/skills/guides/B200_PROFILING.md.  Usage: bash tools/sass_histogram.sh > profiles/rNN_sass_histogram.txt
SO=dsen2_b200/_lib/libdsen2_b200.so
echo "# cuobjdump -sass $SO | opcode histogram ($(date -u +%Y-%m-%d), $(git rev-parse --short HEAD))"
echo "## tensor core / TMEM / TMA / bulk copies"
cuobjdump -sass $SO | grep -oE "\b(UTCHMMA|UTCQMMA|UTCOMMA|LDTM|STTM|UTCBAR|UTCCP|UTMALDG|UTMASTG|UTMAPF|UTMACCTL|UBLKCP|UBLKRED|UBLKPF|SYNCS|FENCE|ELECT|UGETNEXTWORKID)[A-Z0-9_.]*" | sort | uniq -c | sort -rn
echo "## kernels"
cuobjdump -sass $SO | grep -oE "Function : .*" | sed 's/Function : //' | c++filt | sed 's/(.*//' | sort | uniq -c | sort -rn
