"""Dev tool: exactly one CTMRG site-move (projector pair + three absorptions) on the B200 backend, for ncu launch lists."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops
from oracle import ctmrg_oracle as orc

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=8)
ap.add_argument("--chi", type=int, default=256)
ap.add_argument("--d", type=int, default=2)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--engine", default="i8", choices=["i8", "dmma"])
args = ap.parse_args()
D, chi, d = args.D, args.chi, args.d
dev = torch.device("cuda")
torch.manual_seed(0)
site = orc.random_site(D, d, chi)
A = site.A.to(dev)
C = [c.to(dev) for c in site.C]
E = [e.to(dev) for e in site.E]
for _ in range(args.reps):
    Q1, _ = ops.quarter_tensor(C[0], E[0], E[3], A)
    Q4, _ = ops.quarter_tensor(C[3], E[3], E[2], A.permute(3, 0, 1, 2, 4))
    omega = torch.randn(Q4.shape[1], chi + 2, dtype=torch.float64, device=dev)
    encs = [ops.i8_encode(Q1), ops.i8_encode(Q4)] if args.engine == "i8" else None
    mx = torch.ones(2, dtype=torch.float64, device=dev)
    _, S, V, info, AtQ, Wt = ops.rsvd([Q1, Q4], omega, niter=2, chi=chi, cutoff=1e-12, want_u=False, want_atq=True, encs=encs)
    keep = int(info[0])
    p1, p2 = ops.projectors_from_usv(Q1, Q4, None, V, S, keep, qmax1=mx[0:1], qmax4=mx[1:2], AtQ=AtQ, Wt=Wt,
                                     enc4=encs[1] if encs else None)
    p1 = p1.view(chi, D, D, keep)
    p2 = p2.view(chi, D, D, keep)
    c1 = ops.absorb_corner1(C[3], E[2], p1)
    c2 = ops.absorb_corner2(C[0], E[0], p2)
    e = ops.absorb_edge(E[3], A, p2, p1)
    torch.cuda.synchronize()
print("keep", keep, "launches", ops.launch_count())
