#!/bin/bash
# development aid (needs a library built with -DCGGP_DEBUG_KNOBS): time the pipelined matvec with parts switched off
# (results are wrong when CGGP_PIPE_DBG != 0; the shipped build ignores the variable)
for d in 0 1 2 3; do echo "DBG=$d"; CGGP_PIPE_DBG=$d timeout 200 python tools/bench_matvec.py ${@:-c3} 2>&1 | grep "v3"; done
