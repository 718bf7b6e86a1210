"""Dev tool: CUDA-event timing of the stages of one CTMRG site-move on the B200 backend (synthetic random tensors)."""
import argparse
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops
from oracle import ctmrg_oracle as orc

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=8)
ap.add_argument("--chi", type=int, default=256)
ap.add_argument("--d", type=int, default=2)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--out", default="gpurun_out/site_move_profile.json")
args = ap.parse_args()
D, chi, d = args.D, args.chi, args.d
dev = torch.device("cuda")
torch.manual_seed(0)
site = orc.random_site(D, d, chi)
A = site.A.to(dev)
C = [c.to(dev) for c in site.C]
E = [e.to(dev) for e in site.E]
m = chi * D * D
q = chi + 2


def timed(fn, reps=args.reps, warm=1):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


res = {"D": D, "chi": chi, "d": d}
ak = A.permute(0, 1, 2, 3, 4)
ak3 = A.permute(3, 0, 1, 2, 4)
t, (Q1, qd) = timed(lambda: ops.quarter_tensor(C[0], E[0], E[3], ak))
res["quarter_ms"] = t
res["quarter_tflops"] = (2 * chi**3 * D**2 + 2 * chi**3 * D**4 + 4 * chi**2 * D**6 * d) / t * 1e-9
_, (Q4, _) = timed(lambda: ops.quarter_tensor(C[3], E[3], E[2], ak3), reps=1, warm=0)
omega = torch.randn(m, q, dtype=torch.float64, device=dev)
t, Y = timed(lambda: ops.matmul(Q1, omega))
res["thin_gemm_nn_ms"] = t
res["thin_gemm_nn_tflops"] = 2.0 * m * m * q / t * 1e-9
t, _ = timed(lambda: ops.matmul(Q1, omega, transpose_a=True))
res["thin_gemm_tn_ms"] = t
res["thin_gemm_tn_tflops"] = 2.0 * m * m * q / t * 1e-9
for tile in (1, 2, 3):
    for sk in (1, 2, 3, 4, 5, 6, 8):
        try:
            t, _ = timed(lambda: ops.matmul(Q1, omega, force_tile=tile, force_splitk=sk), reps=2)
            res[f"thin_nn_tile{tile}_sk{sk}_tflops"] = 2.0 * m * m * q / t * 1e-9
        except Exception as ex:  # noqa
            res[f"thin_nn_tile{tile}_sk{sk}_tflops"] = str(ex)[:80]
t, _ = timed(lambda: torch.matmul(Q1, omega))
res["cublas_thin_nn_tflops"] = 2.0 * m * m * q / t * 1e-9
t, _ = timed(lambda: ops.orthonormalize(Y.clone()))
res["orthonormalize_ms"] = t
Rm = torch.linalg.qr(torch.randn(q, q, dtype=torch.float64).mul(torch.logspace(0, -8, q, dtype=torch.float64))).R.to(dev).contiguous()
t, (S_, W_, J_, info) = timed(lambda: ops.jacobi_svd(Rm))
res["jacobi_ms"] = t
res["jacobi_sweeps"] = int(info[1])
t, (U, S, V, info) = timed(lambda: ops.rsvd([Q1, Q4], omega, niter=2, chi=chi, cutoff=1e-12), reps=1)
res["rsvd_ms"] = t
res["rsvd_gemm_tflops_equiv"] = (12 * 2.0 * m * m * q) / t * 1e-9
res["rsvd_jacobi_sweeps"] = int(info[1])
keep = int(info[0])
res["keep"] = keep
t, (p1, p2) = timed(lambda: ops.projectors_from_usv(Q1, Q4, U, V, S, keep), reps=1)
res["projectors_ms"] = t
p1 = p1.view(chi, D, D, keep)
p2 = p2.view(chi, D, D, keep)
t, _ = timed(lambda: ops.absorb_corner1(C[3], E[2], p1))
res["corner1_ms"] = t
t, _ = timed(lambda: ops.absorb_corner2(C[0], E[0], p2))
res["corner2_ms"] = t
t, _ = timed(lambda: ops.absorb_edge(E[3], ak, p2, p1))
res["edge_ms"] = t
res["edge_tflops"] = (4 * chi**3 * D**4 + 4 * chi**2 * D**6 * d) / t * 1e-9
tot = 2 * res["quarter_ms"] + res["rsvd_ms"] + res["projectors_ms"] + res["corner1_ms"] + res["corner2_ms"] + res["edge_ms"]
res["site_move_ms"] = tot
res["site_move_tflops"] = orc.flops_site_move(D, chi, d) / tot * 1e-9
res["sweep_2x2_s"] = 16 * tot * 1e-3
res["max_mem_gib"] = torch.cuda.max_memory_allocated() / 2**30
print(json.dumps(res, indent=1))
os.makedirs(os.path.dirname(args.out), exist_ok=True)
open(args.out, "w").write(json.dumps(res, indent=1))
