"""one dense product for ncu (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cggp_b200 as cb
g = torch.Generator(device="cuda").manual_seed(0)
M, B = 4096, 4096
A = torch.randn(M, M, dtype=torch.float64, device="cuda", generator=g); A = A + A.t()
V = torch.randn(B, M, dtype=torch.float64, device="cuda", generator=g)
op = cb.DenseOperator(A)
for _ in range(4):
    Y = op.matmul(V)
torch.cuda.synchronize()
print(float((Y - V @ A).abs().max()))
