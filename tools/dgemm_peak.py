import torch, time
torch.backends.cuda.matmul.allow_tf32 = False
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): (a @ b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"cuBLAS DGEMM n={n}: {2*n**3/best/1e9:.2f} TFLOP/s ({best:.2f} ms)")
# sustained 3 s
n = 8192
t0 = time.time(); cnt = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5): c = a @ b
    cnt += 5; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print(f"cuBLAS DGEMM sustained: {2*n**3*cnt/e0.elapsed_time(e1)/1e9:.2f} TFLOP/s")
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
torch.backends.cuda.matmul.allow_tf32 = True
for _ in range(2): (a @ b)
torch.cuda.synchronize(); best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print(f"cuBLAS TF32 GEMM n={n}: {2*n**3/best/1e9:.2f} TFLOP/s")
