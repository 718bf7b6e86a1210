"""Dev tool: CTMRG sweeps/s of the launch-bound BASELINE configs (1: D=2 chi=20, 2: D=4 chi=64; 2x2 cell) with and without CUDA-graph
replay of the phases, next to the reference's torch path on the same GPU (oracle/_ref, when vendored)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200.ipeps import CTMRGConfig
from acetn_b200.renormalization import DirectionalMover, ctmrg
from acetn_b200.synthetic import random_ipeps


def sweeps_per_s(D, chi, graphs, nsweep=40):
    os.environ["ACETN_B200_GRAPHS"] = graphs
    cfg = CTMRGConfig(steps=1)
    ip = random_ipeps(2, 2, D, chi, 2, seed=0, ctmrg=cfg, device="cuda")
    mover = DirectionalMover(cfg)
    torch.manual_seed(1)
    for _ in range(6):
        ctmrg(ip, cfg, mover)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(nsweep):
        ctmrg(ip, cfg, mover)
    torch.cuda.synchronize()
    return nsweep / (time.perf_counter() - t0), mover.graph_replays


def reference_sweeps_per_s(D, chi, nsweep=5):
    import argparse
    import bench
    args = argparse.Namespace(D=D, chi=chi, d=2, seed=0)
    ip = bench.reference_ipeps(args, 2, 2, "cuda:0")
    if ip is None:
        return None
    gen = bench.reference_site_moves(ip, torch.cuda.synchronize)
    with torch.no_grad():
        for _ in range(16 * 3):
            next(gen)
        t = sum(next(gen) for _ in range(16 * nsweep))
    return nsweep / t


res = []
for D, chi in ((2, 20), (4, 64)):
    eager, _ = sweeps_per_s(D, chi, "0")
    graphed, rep = sweeps_per_s(D, chi, "auto")
    ref = reference_sweeps_per_s(D, chi)
    res.append({"D": D, "chi": chi, "eager_sweeps_s": eager, "graph_sweeps_s": graphed, "graph_replays": rep,
                "reference_torch_gpu_sweeps_s": ref, "speedup_graph_over_eager": graphed / eager,
                "speedup_over_reference_torch_gpu": graphed / ref if ref else None})
    print(json.dumps(res[-1]), flush=True)
