#!/bin/bash
# tuning: exp-table / sqrt variants of the pipelined matvec (CGGP_PIPE_ET: 0 = 32-entry shuffle table + two Newton
# steps, 10 = 1024-entry shared-memory table + third-order sqrt step, the default)
for et in 0 10; do
  echo "== CGGP_PIPE_ET=$et"
  CGGP_PIPE_ET=$et timeout 300 python tools/bench_matvec.py ${@:-c3 c2} 2>&1 | grep -v "^c.: simple"
done
