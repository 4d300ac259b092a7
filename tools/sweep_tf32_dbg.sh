#!/bin/bash
# needs a library built with -DCGGP_DEBUG_KNOBS (the shipped build ignores CGGP_TF32_DBG)
# which part of the tcgen05 gram kernel binds?  (CGGP_TF32_DBG bits: 1 no epilogue math, 2 no MMAs, 4 no TMA copies,
# 8 no global loads of the column scalars, 16 no tcgen05.ld)
for d in ${DBGS:-0 1 2 4 7 15 23 31}; do
  echo "== CGGP_TF32_DBG=$d"
  CGGP_TF32_DBG=$d timeout 120 python tools/bench_matvec.py c5 2>&1 | grep "tcgen05"
done
