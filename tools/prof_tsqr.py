"""Dev tool for ncu: two orthonormalisations (K4: BCGS2 + Householder TSQR) of a 16384 x 258 matrix and one Jacobi SVD core."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops

torch.manual_seed(0)
dev = torch.device("cuda")
for _ in range(2):
    Y = torch.randn(16384, 258, dtype=torch.float64, device=dev)
    Q = ops.orthonormalize(Y)
R = torch.randn(258, 258, dtype=torch.float64, device=dev)
for _ in range(2):
    out = ops.jacobi_svd(R.clone(), chi=256, cutoff=1e-12)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    Q = ops.orthonormalize(torch.randn(16384, 258, dtype=torch.float64, device=dev))
e1.record()
torch.cuda.synchronize()
print("orthonormalize 16384x258: %.3f ms" % (e0.elapsed_time(e1) / 5))
e0.record()
for _ in range(5):
    out = ops.jacobi_svd(R.clone(), chi=256, cutoff=1e-12)
e1.record()
torch.cuda.synchronize()
print("jacobi_svd 258: %.3f ms" % (e0.elapsed_time(e1) / 5))
