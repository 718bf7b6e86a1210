"""Dev tool: peak device memory of the N=1 bench workload (two sweeps at D=8 chi=256, 2x2 cell)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200.ipeps import CTMRGConfig
from acetn_b200.renormalization import DirectionalMover, ctmrg
from acetn_b200.synthetic import random_ipeps

dev = torch.device("cuda", 0)
cfg = CTMRGConfig(steps=1)
ip = random_ipeps(2, 2, 8, 256, 2, seed=0, ctmrg=cfg, device=dev)
mover = DirectionalMover(cfg)
torch.manual_seed(1)
for _ in range(2):
    ctmrg(ip, cfg, mover)
torch.cuda.synchronize()
print("peak allocated %.1f GiB, reserved %.1f GiB" % (torch.cuda.max_memory_allocated() / 2 ** 30, torch.cuda.max_memory_reserved() / 2 ** 30))
