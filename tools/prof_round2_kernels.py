"""Dev tool for ncu: one launch each of K1 on the 2 chi^3 D^4 class (16384 x 16384 x 256), the fused orthonormalisation (1024 x 66)
and the single-CTA Jacobi SVD (66 x 66, QR-preconditioned core as in the rSVD pipeline), after a warm-up launch of each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
m, k = 16384, 256
At = torch.randn(k, m, dtype=torch.float64, device=dev)
B = torch.randn(k, m, dtype=torch.float64, device=dev)
C = torch.empty(m, m, dtype=torch.float64, device=dev)
Y = torch.randn(1024, 66, dtype=torch.float64, device=dev)
G = torch.randn(66, 66, dtype=torch.float64, device=dev) * torch.logspace(0, -9, 66, dtype=torch.float64, device=dev)[None, :]
R = torch.linalg.qr(G).R.contiguous()
for _ in range(2):
    ops.matmul(At, B, transpose_a=True, out=C)
    ops.orthonormalize(Y.clone())
    ops.jacobi_svd(R, chi=64, cutoff=1e-12)
    torch.cuda.synchronize()
