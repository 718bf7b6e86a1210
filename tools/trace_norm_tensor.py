"""Dev tool: kernel breakdown (torch.profiler) of build_norm_tensor and of one whole bond update at D=8 chi=256."""
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

D, chi, d = 8, 256, 2
dev = torch.device("cuda")
cell = orc.random_cell(2, 2, D, chi, d, seed=0)
ip = Ipeps.from_plain(cell)
bond = cell.bond_list[0]
nD = min(D ** 3, d * D)
g = torch.Generator().manual_seed(1)
a1q = torch.linalg.qr(torch.randn(D ** 3, nD, dtype=torch.float64, generator=g)).Q.reshape(D, D, D, nD).to(dev)
a2q = torch.linalg.qr(torch.randn(D ** 3, nD, dtype=torch.float64, generator=g)).Q.reshape(D, D, D, nD).to(dev)
for _ in range(2):
    n12 = evo.build_norm_tensor(ip, bond, a1q, a2q)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    n12 = evo.build_norm_tensor(ip, bond, a1q, a2q)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        name = re.sub(r"\(.*", "", e.name.replace("(anonymous namespace)::", "")).replace("void ", "")
        ev.append((e.time_range.start, e.time_range.end, name))
ev.sort()
print("build_norm_tensor wall %.2f ms, kernel sum %.2f ms" % ((ev[-1][1] - ev[0][0]) / 1e3, sum(b - a for a, b, _ in ev) / 1e3))
for a, b, n in ev:
    if b - a > 200:
        print("  %8.2f ms  %s" % ((b - a) / 1e3, n[:90]))
