"""Dev tool: sweeps K5 needs on the triangular core R of a tall matrix vs on its transpose (Drmac-Veselic: one-sided Jacobi converges
faster on one of the two), for spectra like the rSVD's (graded over several decades) and for a flat one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
for q in (66, 258):
    for decades in (0, 4, 12):
        m = 4096
        U = torch.linalg.qr(torch.randn(m, q, dtype=torch.float64, generator=g)).Q.to(dev)
        V = torch.linalg.qr(torch.randn(q, q, dtype=torch.float64, generator=g)).Q.to(dev)
        s = torch.logspace(0, -decades, q, dtype=torch.float64, device=dev)
        B = (U * s) @ V.T
        R = torch.linalg.qr(B).R.contiguous()
        ref = torch.linalg.svdvals(R)
        out = []
        order = torch.argsort(R.norm(dim=1), descending=True)
        cols = torch.argsort(R.norm(dim=0), descending=True)
        for name, M in (("R (upper)", R), ("rows sorted by norm", R[order].contiguous()), ("rows ascending", R[order.flip(0)].contiguous()),
                        ("cols sorted", R[:, cols].contiguous())):
            S, Wt, Jt, info = ops.jacobi_svd(M)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            S, Wt, Jt, info = ops.jacobi_svd(M)
            e1.record()
            torch.cuda.synchronize()
            err = float((S - ref).abs().max() / ref[0])
            out.append(f"{name}: sweeps {int(info[1])}, {e0.elapsed_time(e1):.3f} ms, |ds|/s0 {err:.1e}")
        print(f"q={q} decades={decades}: " + " | ".join(out), flush=True)
