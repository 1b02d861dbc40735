#!/usr/bin/env bash
# Writes profiles/<round>/sass_production.txt: cuobjdump listings of the production kernels of the
# in-tree libb200va.so (which is git-ignored), so the SASS evidence travels with the history.
#   bash tools/sass_evidence.sh r02
set -eu
exec 2> >(grep -v 'not found' >&2)   # the fatbin holds two ELF images; only one has the kernels
R=${1:-r02}
LIB=k8s-gpu-hpa_b200/libb200va.so
OUT=profiles/$R/sass_production.txt
mkdir -p profiles/$R
{
  echo "# SASS evidence for $LIB  (built $(date -u +%F) by make -C k8s-gpu-hpa_b200; nvcc $(nvcc --version | grep -o 'V[0-9][0-9.]*' | tail -1))"
  echo "# arch list:"; cuobjdump -lelf $LIB | sed 's/^/#   /'; cuobjdump -lptx $LIB | sed 's/^/#   /'
  echo "# kernels in the library: $(cuobjdump -sass $LIB | grep -c 'Function :')   sha256 $(sha256sum $LIB | cut -c1-16)"
  echo "# mnemonic census over the whole library:"
  cuobjdump -sass $LIB | grep -oE '\b(LDG\.E[.A-Z0-9]*\.(128|256)|STG\.E[.A-Z0-9]*\.(128|256)|UBLKCP\.S\.G|UBLKCP\.G\.S|UGETNEXTWORKID\.SELFCAST|UBLKPF\.L2|SYNCS\.[A-Z.0-9]+|ACQBULK|PREEXIT|HMMA[.A-Z0-9]*|UTCHMMA[.A-Z0-9]*)' | sort | uniq -c | sed 's/^/#   /'
  for f in \
    _ZN6b200va8vadd_vecILi4ELi1ELi0ELi1ELi0EEEvPKfS2_Pfmmmm \
    _ZN6b200va8vadd_vecILi4ELi1ELi0ELi1ELi1EEEvPKfS2_Pfmmmm \
    _ZN6b200va8vadd_vecILi4ELi1ELi0ELi1ELi2EEEvPKfS2_Pfmmmm \
    _ZN6b200va8vadd_vecILi4ELi2ELi3ELi0ELi1EEEvPKfS2_Pfmmmm \
    _ZN6b200va12vadd_vec_clcILi4ELi2ELi0ELi1ELi0EEEvPKfS2_Pfmmmm \
    _ZN6b200va12vadd_tma_clcILb0ELi1EEEvPKfS2_Pfmmmjj \
    _ZN6b200va8vadd_vecILi8ELi1ELi0ELi1ELi0EEEvPKfS2_Pfmmmm ; do
    echo; echo "==================== $(echo $f | c++filt)"
    cuobjdump -sass -fun $f $LIB | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's@\s+/\* 0x[0-9a-f]+ \*/\s*$@@'
  done
} > $OUT
wc -l $OUT
