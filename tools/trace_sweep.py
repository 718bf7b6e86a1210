"""Dev tool: kernel timeline of ONE CTMRG sweep (bench.py's N=1 workload) through torch.profiler (CUPTI activity records,
so kernels on all four streams keep their real overlap -- unlike ncu, which serialises).  Prints, per kernel name, the
summed duration and the wall time ATTRIBUTED to it (each instant of the timeline is split evenly among the kernels
running at that instant), plus GPU idle time inside the sweep.  Writes gpurun_out/trace_sweep.json.

    python tools/trace_sweep.py [--D 8 --chi 256] [--out gpurun_out/trace_sweep.json]
"""
import argparse
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from acetn_b200.ipeps import CTMRGConfig
from acetn_b200.renormalization import DirectionalMover, ctmrg
from acetn_b200.synthetic import random_ipeps

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=8)
ap.add_argument("--chi", type=int, default=256)
ap.add_argument("--d", type=int, default=2)
ap.add_argument("--nx", type=int, default=2)
ap.add_argument("--ny", type=int, default=2)
ap.add_argument("--out", default="gpurun_out/trace_sweep.json")
args = ap.parse_args()

dev = torch.device("cuda", 0)
cfg = CTMRGConfig(steps=1)
ip = random_ipeps(args.nx, args.ny, args.D, args.chi, args.d, seed=0, ctmrg=cfg, device=dev)
mover = DirectionalMover(cfg)
torch.manual_seed(1)
for _ in range(3):
    ctmrg(ip, cfg, mover)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ctmrg(ip, cfg, mover)
    torch.cuda.synchronize()

chrome = os.path.splitext(args.out)[0] + "_chrome.json"
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
prof.export_chrome_trace(chrome)
os.system(f"gzip -f {chrome}")
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        name = e.name.replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void ", "", name)
        ev.append((e.time_range.start, e.time_range.end, name))
ev.sort()
t0, t1 = min(s for s, _, _ in ev), max(e for _, e, _ in ev)
points = sorted(set([s for s, _, _ in ev] + [e for _, e, _ in ev]))
# sweep line
import heapq
attributed = collections.defaultdict(float)
summed = collections.defaultdict(float)
count = collections.Counter()
for s, e, n in ev:
    summed[n] += e - s
    count[n] += 1
bounds = []
for i, (s, e, n) in enumerate(ev):
    bounds.append((s, 1, i))
    bounds.append((e, 0, i))
bounds.sort()
active = set()
idle = 0.0
prev = bounds[0][0]
for t, kind, i in bounds:
    dt = t - prev
    if dt > 0:
        if active:
            share = dt / len(active)
            for j in active:
                attributed[ev[j][2]] += share
        else:
            idle += dt
    prev = t
    if kind == 1:
        active.add(i)
    else:
        active.discard(i)
wall = t1 - t0
# time during which no throughput-bound kernel (any launch longer than 0.3 ms: the big DGEMMs, K2, the K7 GEMM, the encodings) runs:
# what the latency-bound chains (TSQR, Jacobi, CRT, ...) fail to hide
big = sorted((s, e) for s, e, n in ev if e - s > 300.0)
covered, cur_s, cur_e = 0.0, None, None
for s, e in big:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            covered += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
if cur_e is not None:
    covered += cur_e - cur_s
no_big = wall - covered
print(f"time with no throughput-bound kernel running: {no_big / 1e3:.2f} ms ({100 * no_big / wall:.1f} % of the sweep)")
rows = sorted(attributed.items(), key=lambda kv: -kv[1])
print(f"wall {wall / 1e3:.2f} ms, idle {idle / 1e3:.2f} ms ({100 * idle / wall:.1f} %), kernels {len(ev)}")
print(f"{'kernel':70s} {'n':>6s} {'sum ms':>9s} {'attr ms':>9s} {'attr %':>7s}")
for n, a in rows[:40]:
    print(f"{n[:70]:70s} {count[n]:6d} {summed[n] / 1e3:9.2f} {a / 1e3:9.2f} {100 * a / wall:7.1f}")
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump({"wall_ms": wall / 1e3, "idle_ms": idle / 1e3, "no_throughput_kernel_ms": no_big / 1e3, "n_kernels": len(ev), "cell": f"{args.nx}x{args.ny}",
           "inflight": os.environ.get("ACETN_B200_INFLIGHT", "4"),
           "kernels": [{"name": n, "launches": count[n], "sum_ms": summed[n] / 1e3, "attributed_ms": a / 1e3} for n, a in rows]},
          open(args.out, "w"), indent=1)
