"""Dev tool: kernel timeline of ONE site-move running alone (the N=8 situation: one task per rank and phase) through
torch.profiler; prints the serial chain compressed by kernel name and writes gpurun_out/trace_site_move_chrome.json.gz."""
import gzip
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TRACE_CHILD"] = "1"
import torch
from torch.profiler import ProfilerActivity, profile

sys.argv = [sys.argv[0], "--reps", "2"]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "one_site_move.py")).read()
exec(compile(src, "one_site_move.py", "exec"))          # warm-up (2 reps)
sys.argv = [sys.argv[0], "--reps", "1"]
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    exec(compile(src, "one_site_move.py", "exec"))
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
path = "gpurun_out/trace_site_move_chrome.json"
prof.export_chrome_trace(path)
d = json.load(open(path))
subprocess.call(["gzip", "-f", path])
ev = [e for e in d["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]


def nm(e):
    n = e["name"].replace("(anonymous namespace)::", "")
    return re.sub(r"\(.*", "", n).replace("void ", "").replace("ab200::", "")[:48]


out = []
for e in ev:
    n = nm(e)
    if out and out[-1][0] == n:
        out[-1][2] += e["dur"]; out[-1][3] += 1; out[-1][4] = e["ts"] + e["dur"] - t0
    else:
        out.append([n, e["ts"] - t0, e["dur"], 1, e["ts"] + e["dur"] - t0])
print("wall %.2f ms, kernel busy %.2f ms, launches %d" % ((ev[-1]["ts"] + ev[-1]["dur"] - t0) / 1e3, sum(e["dur"] for e in ev) / 1e3, len(ev)))
for o in out:
    print(f"{o[1] / 1e3:8.2f} -> {o[4] / 1e3:8.2f}  span {(o[4] - o[1]) / 1e3:6.2f}  busy {o[2] / 1e3:6.2f} ms x{o[3]:4d} {o[0]}")
