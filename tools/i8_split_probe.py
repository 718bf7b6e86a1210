"""Dev tool: time the K7 thin product (forward + adjoint) at D=8 chi=256 for the MMA column split / stage count given by
the env knobs ACETN_B200_I8_N1 / ACETN_B200_I8_STAGES (read once per process), and check the result against K1 (DMMA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops

n, q = 16384, 258
dev = torch.device("cuda")
torch.manual_seed(0)
Q = torch.rand(n, n, dtype=torch.float64, device=dev) - 0.3
Y = torch.randn(n, q, dtype=torch.float64, device=dev)
enc = ops.i8_encode(Q)
out = torch.empty(n, q, dtype=torch.float64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for adj in (False, True):
    for _ in range(3):
        ops.i8_matmul(enc, Y, adjoint=adj, out=out)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.i8_matmul(enc, Y, adjoint=adj, out=out)
    e1.record()
    torch.cuda.synchronize()
    ref = ops.matmul(Q, Y, transpose_a=adj)
    err = float((out - ref).abs().max() / ref.abs().max())
    res["adjoint" if adj else "forward"] = (e0.elapsed_time(e1) / 20, err)
print("N1=%s STAGES=%s  forward %.4f ms (err %.1e)  adjoint %.4f ms (err %.1e)" % (
    os.environ.get("ACETN_B200_I8_N1", "-"), os.environ.get("ACETN_B200_I8_STAGES", "-"),
    res["forward"][0], res["forward"][1], res["adjoint"][0], res["adjoint"][1]), flush=True)
