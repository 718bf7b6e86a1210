"""Dev tool: the reference README quickstart (Heisenberg 2x2, D=2, chi=20, evolve(0.01, 100) + measure) on backend='b200' and on the
reference's torch path on the same GPU, each run twice in one process (the second run excludes first-use warm-up: module load, kernel
load, CUDA-graph capture)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.dropin_util import setup

Ipeps = setup()
base = {"dtype": "float64", "device": "cuda", "TN": {"nx": 2, "ny": 2, "dims": {"phys": 2, "bond": 2, "chi": 20}},
        "model": {"name": "heisenberg", "params": {"J": 1.0}}}
for backend in ("b200", "torch"):
    for run in range(2):
        cfg = dict(base)
        cfg["evolution"] = {"backend": backend, "disable_progressbar": True}
        cfg["ctmrg"] = {"disable_progressbar": True}
        torch.manual_seed(0)
        ip = Ipeps(cfg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ip.evolve(dtau=0.01, steps=100)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        e = float(ip.measure()["Energy"])
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{backend} run {run}: evolve {t1 - t0:.2f} s, measure {t2 - t1:.2f} s, E = {e:.13f}", flush=True)
