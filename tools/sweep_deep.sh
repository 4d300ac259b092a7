#!/bin/bash
# development aid: the pipelined matvec with two 48-row K buffers (CGGP_PIPE_DEEP=0) vs three 32-row buffers (=1)
for d in 0 1; do echo "DEEP=$d"; CGGP_PIPE_DEEP=$d timeout 300 python tools/bench_matvec.py ${@:-c3 c2 c1 c4} 2>&1 | grep "v3"; done
