#!/bin/bash
# development aid: time the pipelined matvec for the phase-2 interleave depths
for hb in 6 5 4 3; do echo "HB=$hb"; CGGP_PIPE_HB=$hb timeout 200 python tools/bench_matvec.py c3 c2 2>&1 | grep "v3\|diff"; done
