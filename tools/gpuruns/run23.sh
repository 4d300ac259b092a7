set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_ab_*.log
for i in 1 2; do
CGGP_PIPE_CLAMP=1 timeout 300 python tools/bench_matvec.py c3 c2 c4 2>&1 | grep "fused matvec" >> gpurun_out/r2_ab_clamp.log
timeout 300 python tools/bench_matvec.py c3 c2 c4 2>&1 | grep "fused matvec" >> gpurun_out/r2_ab_noclamp.log
done
sort gpurun_out/r2_ab_clamp.log; sort gpurun_out/r2_ab_noclamp.log
