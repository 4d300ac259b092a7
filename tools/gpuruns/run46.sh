set -x
mkdir -p gpurun_out
cat > /tmp/ct_case.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import cggp_b200 as cb
g = torch.Generator(device="cuda").manual_seed(1)
n = 2_000_000
x = torch.rand((n, 2), dtype=torch.float64, device="cuda", generator=g) * 20.0 - 10.0
y = torch.randn((n, 1), dtype=torch.float64, device="cuda", generator=g)
t = cb.CoverTree(None, (x, y), spatial_resolution=0.25)
m, c = t.cluster_mean_and_counts
torch.cuda.synchronize()
print([t.level_size(l) for l in range(t.num_levels)], float(c.sum()))
PY
python /tmp/ct_case.py > gpurun_out/plain_ct.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_ct_launches.csv python /tmp/ct_case.py > gpurun_out/ncu_ct.log 2>&1
tail -2 gpurun_out/plain_ct.log; wc -l gpurun_out/r2_ct_launches.csv
