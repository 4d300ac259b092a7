set -x
mkdir -p gpurun_out
CGGP_CT_TRACE=1 timeout 300 python tools/covertree_bench.py 2000000 2 0.25 > gpurun_out/r2_ctb37b.log 2>&1
CGGP_CT_TRACE=1 CGGP_CT_CLUSTER=1 timeout 300 python tools/covertree_bench.py 2000000 2 0.25 > gpurun_out/r2_ctb37b1.log 2>&1
CGGP_CT_TRACE=1 timeout 300 python tools/covertree_bench.py 434874 3 0.8 > gpurun_out/r2_ctb37c.log 2>&1
CGGP_CT_TRACE=1 CGGP_CT_CLUSTER=1 timeout 300 python tools/covertree_bench.py 434874 3 0.8 > gpurun_out/r2_ctb37c1.log 2>&1
for f in b b1 c c1; do echo "== $f"; grep -n "greedy pass\|built in\|^device\|root\|Voronoi\|host" gpurun_out/r2_ctb37$f.log | awk -F: '$1>20 && $1<48' | cut -c1-170; done
