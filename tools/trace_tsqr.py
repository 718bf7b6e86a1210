"""Dev tool: per-kernel timeline of ONE orthonormalisation (16384 x 258) and ONE Jacobi core (258) through torch.profiler."""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from acetn_b200 import ops

dev = torch.device("cuda")
torch.manual_seed(0)
Y0 = torch.randn(16384, 258, dtype=torch.float64, device=dev)
R = torch.randn(258, 258, dtype=torch.float64, device=dev)
for _ in range(3):
    ops.orthonormalize(Y0.clone())
    ops.jacobi_svd(R.clone(), chi=256, cutoff=1e-12)
torch.cuda.synchronize()
for what in ("orthonormalize", "jacobi"):
    Y = Y0.clone()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        if what == "orthonormalize":
            ops.orthonormalize(Y)
        else:
            ops.jacobi_svd(R.clone(), chi=256, cutoff=1e-12)
        torch.cuda.synchronize()
    ev = sorted((e.time_range.start, e.time_range.end, re.sub(r"\(.*", "", e.name.replace("void ", ""))) for e in prof.events()
                if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None)
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    busy = sum(e[1] - e[0] for e in ev)
    agg, cnt = collections.defaultdict(float), collections.Counter()
    for s, e, n in ev:
        agg[n] += e - s
        cnt[n] += 1
    print(f"== {what}: wall {(t1 - t0) / 1e3:.3f} ms, kernels {len(ev)}, sum of kernel durations {busy / 1e3:.3f} ms, gaps {(t1 - t0 - busy) / 1e3:.3f} ms")
    for n, t in sorted(agg.items(), key=lambda kv: -kv[1])[:12]:
        print(f"   {n[:80]:80s} n={cnt[n]:4d}  total {t / 1e3:7.3f} ms  avg {t / cnt[n]:7.1f} us")
    if what == "orthonormalize":
        print("   first 40 launches (us):", " ".join(f"{n.split('::')[-1][:12]}={e - s_:.0f}" for s_, e, n in ev[:40]))
