#!/bin/bash
# SASS mnemonics that prove the TMA / tcgen05 paths (B200_PROFILING.md): run after `make`, writes profiles/r02_sass_*.txt
set -e
cd "$(dirname "$0")/.."
cuobjdump -sass build/obj/hex8_tc32.o | grep -E "UTCHMMA|LDTM|UTCBAR|UTMALDG|UTCATOMSWS" | awk '{$1=$1};1' | sed -E 's/\/\*[0-9a-f]+\*\/ //; s/ \/\*.*//' | sed -E 's/UR[0-9]+/UR/g; s/R[0-9]+/R/g' | sort | uniq -c | sort -rn > profiles/r02_sass_tcgen05.txt
cuobjdump -sass build/obj/hex8.o | awk '/Function : .*k_hex8_zm4ILi16ELi1ELb0/{f=1} f&&/Function :/&&!/k_hex8_zm4ILi16ELi1ELb0/{f=0} f' | grep -E "UTMALDG|SYNCS|LDCU|DFMA|DADD|DMUL" | awk '{$1=$1};1' | sed -E 's/\/\*[0-9a-f]+\*\/ //; s/ \/\*.*//' | awk '{print $1}' | sed -E 's/^@!?U?P[0-9T]+//' | sort | uniq -c | sort -rn > profiles/r02_sass_zm4_epi_dot.txt
head -8 profiles/r02_sass_tcgen05.txt; echo ...; head -8 profiles/r02_sass_zm4_epi_dot.txt
