#!/bin/bash
# tuning: exp-table / sqrt variants of the pipelined matvec (CGGP_PIPE_ET: 0 baseline, 100 cubic sqrt, 110 smem table, 10 both)
for et in 0 100 110 10; do
  echo "== CGGP_PIPE_ET=$et"
  CGGP_PIPE_ET=$et timeout 300 python tools/bench_matvec.py c3 c2 2>&1 | grep -v "^c.: simple"
done
