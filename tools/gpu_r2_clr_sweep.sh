#!/bin/bash
# clr_mag_kernel: CTA size x resident CTAs per SM (rebuilds mp2_score.o on the box), isolated call times
for v in "1024 1" "512 2" "512 3" "256 4" "256 6"; do
  set -- $v
  make -C magpurify2_b200/csrc -B EXTRA="-DMP2_CLRM_THREADS=$1 -DMP2_CLRM_MINB=$2" mp2_score.o > /dev/null 2>&1 && make -C magpurify2_b200/csrc > /dev/null 2>&1 || echo BUILD FAILED
  echo "threads $1 min CTAs/SM $2: $(python tools/score_probe.py 2>/dev/null | grep clr_centroid)"
done
