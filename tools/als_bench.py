"""Dev tool: time the persistent ALS kernel (K6) against the reference loop (oracle port) on the same GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acetn_b200 import ops
from oracle import ctmrg_oracle as orc

res = {}
for (D, d) in [(6, 2), (8, 2)]:
    nD, bD, pD = min(D ** 3, d * D), D, d
    g = torch.Generator().manual_seed(0)
    M = torch.randn(nD * nD, nD * nD, dtype=torch.float64, generator=g)
    n12 = (M @ M.T / (nD * nD) + 0.1 * torch.eye(nD * nD, dtype=torch.float64)).reshape(nD, nD, nD, nD).permute(0, 1, 2, 3).contiguous()
    a1 = torch.randn(nD, bD, pD, dtype=torch.float64, generator=g)
    a2 = torch.randn(nD, bD, pD, dtype=torch.float64, generator=g)
    a12g = torch.einsum("yup,xuq->yxpq", a1, a2) + 0.05 * torch.randn(nD, nD, pD, pD, dtype=torch.float64, generator=g)
    n12g = torch.einsum("yxYX,yxpq->YXpq", n12, a12g)
    a1r0, a2r0 = orc.als_initial_guess(a12g, (nD, bD, pD))
    dev = "cuda"
    args = [t.to(dev) for t in (a1r0, a2r0, n12g, n12, a12g)]
    for niter in (10, 100):
        ops.als_solve(*args, niter=niter, tol=-1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b1, b2, info = ops.als_solve(*args, niter=niter, tol=-1.0); e1.record(); torch.cuda.synchronize()
        t_k6 = e0.elapsed_time(e1)
        t0 = time.perf_counter(); r1, r2, it = orc.als_solve(*args, niter=niter, tol=-1.0); torch.cuda.synchronize(); t_ref = (time.perf_counter() - t0) * 1e3
        err = float((b1 - r1).norm() / r1.norm())
        res[f"D{D}_niter{niter}"] = {"k6_ms": round(t_k6, 3), "torch_gpu_ms": round(t_ref, 2), "us_per_iter_k6": round(1e3 * t_k6 / niter, 1), "rel_diff": err,
                                     "chol_fail": int(info[1])}
print(json.dumps(res, indent=1))
