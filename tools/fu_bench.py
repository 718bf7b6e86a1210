"""Dev tool: one full-update bond update (FullUpdater.tensor_update) at the benchmarking_full_update.py shape
(BASELINE config 3: D=6, chi=144): acetn_b200.evolution.full_update_bond vs the reference torch path on the same GPU
(oracle port: torch.einsum / linalg.qr / eigh / pinv / svd -> cuBLAS + cuSOLVER)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import evolution as evo
from acetn_b200.ipeps import Ipeps
from oracle import ctmrg_oracle as orc

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=6)
ap.add_argument("--chi", type=int, default=144)
ap.add_argument("--d", type=int, default=2)
args = ap.parse_args()


class Cfg:
    als_niter, als_tol, als_method, als_epsilon = 10, -float("inf"), "cholesky", 1e-12     # benchmarking_full_update.py:142-161
    use_gauge_fix, gauge_fix_atol, positive_approx_cutoff = True, 1e-12, 1e-12


dev = torch.device("cuda")
cell = orc.random_cell(2, 2, args.D, args.chi, args.d, seed=0)
ip = Ipeps.from_plain(cell)
bond = cell.bond_list[0]
s1, s2, k = bond
perm = [(i + k) % 4 for i in range(4)] + [4]
a1 = cell[s1].A.permute(perm).contiguous().to(dev)
a2 = cell[s2].A.permute(perm).contiguous().to(dev)
gate = torch.linalg.matrix_exp(-0.01 * orc.heisenberg_bond_hamiltonian(1.0)).reshape(2, 2, 2, 2).to(dev) if args.d == 2 else \
    torch.eye(args.d ** 2, dtype=torch.float64).reshape(args.d, args.d, args.d, args.d).to(dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


t_b200, (o1, o2) = timed(lambda: evo.full_update_bond(ip, bond, a1, a2, gate, Cfg))
gcell = cell.clone()
for s in gcell.site_list:
    st = gcell[s]
    st.A = st.A.to(dev); st.C = [c.to(dev) for c in st.C]; st.E = [e.to(dev) for e in st.E]
t_ref, (r1, r2) = timed(lambda: orc.full_update_bond(gcell, bond, a1, a2, gate, als_niter=10, als_tol=-float("inf")))
th, ref = orc.bond_theta(o1, o2), orc.bond_theta(r1, r2)
print("D=%d chi=%d: full_update_bond b200 %.1f ms, reference torch path on the same GPU %.1f ms (x%.2f); theta rel diff %.1e" % (
    args.D, args.chi, t_b200, t_ref, t_ref / t_b200, float((th - ref).norm() / ref.norm())))
