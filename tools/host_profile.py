"""Dev tool: where the HOST time of a launch-bound sweep goes (cProfile of `ctmrg` at a small size, eager schedule)."""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200.ipeps import CTMRGConfig
from acetn_b200.renormalization import DirectionalMover, ctmrg
from acetn_b200.synthetic import random_ipeps

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=4)
ap.add_argument("--chi", type=int, default=64)
ap.add_argument("--sweeps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda", 0)
cfg = CTMRGConfig(steps=1)
ip = random_ipeps(2, 2, args.D, args.chi, 2, seed=0, ctmrg=cfg, device=dev)
mover = DirectionalMover(cfg)
torch.manual_seed(1)
for _ in range(5):
    ctmrg(ip, cfg, mover)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.sweeps):
    ctmrg(ip, cfg, mover)
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"{args.sweeps} sweeps: {(t1 - t0) / args.sweeps * 1e3:.2f} ms per sweep (no profiler)")
pr = cProfile.Profile()
pr.enable()
for _ in range(args.sweeps):
    ctmrg(ip, cfg, mover)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
