"""Dev tool: kernel breakdown (torch.profiler) of one full-update bond update (evolution.full_update_bond) at D=8 chi=256."""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from acetn_b200 import evolution as evo
from acetn_b200.ipeps import Ipeps
from oracle import ctmrg_oracle as orc


class Cfg:
    als_niter, als_tol, als_method, als_epsilon = 10, -float("inf"), "cholesky", 1e-12
    use_gauge_fix, gauge_fix_atol, positive_approx_cutoff = True, 1e-12, 1e-12


D, chi, d = 8, 256, 2
dev = torch.device("cuda")
cell = orc.random_cell(2, 2, D, chi, d, seed=0)
ip = Ipeps.from_plain(cell)
bond = cell.bond_list[0]
a1 = cell[bond[0]].A.contiguous().to(dev)
a2 = cell[bond[1]].A.contiguous().to(dev)
gate = torch.linalg.matrix_exp(-0.01 * orc.heisenberg_bond_hamiltonian(1.0)).reshape(2, 2, 2, 2).to(dev)
for _ in range(2):
    evo.full_update_bond(ip, bond, a1, a2, gate, Cfg)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    evo.full_update_bond(ip, bond, a1, a2, gate, Cfg)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        name = re.sub(r"\(.*", "", e.name.replace("(anonymous namespace)::", "")).replace("void ", "")
        ev.append((e.time_range.start, e.time_range.end, name))
ev.sort()
t0 = ev[0][0]
print("bond update wall %.2f ms, kernel sum %.2f ms, %d launches" % ((ev[-1][1] - t0) / 1e3, sum(b - a for a, b, _ in ev) / 1e3, len(ev)))
agg = collections.defaultdict(lambda: [0, 0.0])
for a, b, n in ev:
    agg[n][0] += 1
    agg[n][1] += b - a
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("  %8.2f ms x%4d  %s" % (t / 1e3, c, n[:80]))
# coarse timeline: gaps and segments after the norm tensor
last = t0
for a, b, n in ev:
    if b - a > 1000 or "jacobi_block" in n or "als" in n:
        print("   @%7.2f ms  %6.2f ms  %s" % ((a - t0) / 1e3, (b - a) / 1e3, n[:60]))
