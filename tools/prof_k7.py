"""Dev tool for ncu: one encode + a few K7 thin products (NN, TN) + the K1 medium GEMM (M=N=chi D^2, K=chi) at D=8 chi=256."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops

n, q, chi = 16384, 258, 256
dev = torch.device("cuda")
torch.manual_seed(0)
Q = torch.rand(n, n, dtype=torch.float64, device=dev) - 0.3
Y = torch.randn(n, q, dtype=torch.float64, device=dev)
enc = ops.i8_encode(Q)
for _ in range(2):
    enc = ops.i8_encode(Q, storage=enc.storage)
    out = ops.i8_matmul(enc, Y)
    out = ops.i8_matmul(enc, Y, adjoint=True)
A = torch.randn(n, chi, dtype=torch.float64, device=dev)
B = torch.randn(chi, n, dtype=torch.float64, device=dev)
for _ in range(2):
    ops.matmul(A, B, out=Q)
torch.cuda.synchronize()
print("done")
