"""Dev tool: time the B200 measure path (site + bond RDMs, energy) and the norm tensor at benchmark shapes."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acetn_b200.ipeps import Ipeps
from acetn_b200.measurement import RDM, measure
from acetn_b200.evolution import build_norm_tensor
from oracle import ctmrg_oracle as orc

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=8); ap.add_argument("--chi", type=int, default=256); ap.add_argument("--d", type=int, default=2)
a = ap.parse_args()
cell = orc.random_cell(2, 2, a.D, a.chi, a.d, seed=0)
ip = Ipeps.from_plain(cell)
rdm = RDM(ip)
def timed(f):
    f(); torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); return time.perf_counter() - t, r
res = {"D": a.D, "chi": a.chi}
res["site_rdm_s"], _ = timed(lambda: rdm[(0, 0)])
res["bond_rdm_s"], _ = timed(lambda: rdm[ip.bond_list[0]])
measure(ip, orc.heisenberg_bond_hamiltonian(1.0)); torch.cuda.synchronize()      # warm-up: every kernel variant of all bond directions resident
t = time.perf_counter(); out = measure(ip, orc.heisenberg_bond_hamiltonian(1.0)); torch.cuda.synchronize(); res["measure_s"] = time.perf_counter() - t
res["energy"] = float(out["Energy"])
nD = min(a.D ** 3, a.d * a.D)
g = torch.Generator().manual_seed(1)
a1q = torch.linalg.qr(torch.randn(a.D ** 3, nD, dtype=torch.float64, generator=g)).Q.reshape(a.D, a.D, a.D, nD).cuda()
a2q = torch.linalg.qr(torch.randn(a.D ** 3, nD, dtype=torch.float64, generator=g)).Q.reshape(a.D, a.D, a.D, nD).cuda()
res["norm_tensor_s"], _ = timed(lambda: build_norm_tensor(ip, ip.bond_list[0], a1q, a2q))
res["max_mem_gib"] = torch.cuda.max_memory_allocated() / 2 ** 30
print(json.dumps(res))
