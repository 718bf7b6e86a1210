"""Dev tool: K1 on the headline thin-GEMM shapes (Q (m x m) times thin (m x q)), per tile config / split-K, vs cuBLAS."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=16384)
ap.add_argument("--q", type=int, default=258)
ap.add_argument("--tiles", default="0,1,2,3")
ap.add_argument("--splitk", default="0")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--once", action="store_true", help="single launch per config (for ncu)")
ap.add_argument("--trans", default="0,1")
ap.add_argument("--square", action="store_true", help="also time the 16384x16384x256 (Q.s2-like) shape")
args = ap.parse_args()
dev = torch.device("cuda")
m, q = args.m, args.q
Q = torch.randn(m, m, dtype=torch.float64, device=dev)
X = torch.randn(m, q, dtype=torch.float64, device=dev)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {}
for tr in [int(t) for t in args.trans.split(",")]:
    for tile in [int(t) for t in args.tiles.split(",")]:
        for sk in [int(s) for s in args.splitk.split(",")]:
            fn = lambda: ops.matmul(Q, X, transpose_a=bool(tr), force_tile=tile, force_splitk=sk)
            if args.once:
                fn(); torch.cuda.synchronize(); continue
            try:
                t = timed(fn, args.reps)
                res[f"{'TN' if tr else 'NN'}_tile{tile}_sk{sk}"] = round(2.0 * m * m * q / t * 1e-9, 2)
            except Exception as ex:
                res[f"{'TN' if tr else 'NN'}_tile{tile}_sk{sk}"] = str(ex)[:100]
if not args.once:
    t = timed(lambda: torch.matmul(Q, X), args.reps)
    res["cublas_NN"] = round(2.0 * m * m * q / t * 1e-9, 2)
    t = timed(lambda: torch.matmul(Q.T, X), args.reps)
    res["cublas_TN"] = round(2.0 * m * m * q / t * 1e-9, 2)
    if args.square:
        B = torch.randn(256, m, dtype=torch.float64, device=dev)
        At = torch.randn(256, m, dtype=torch.float64, device=dev)
        for tile in [int(t) for t in args.tiles.split(",")]:
            t = timed(lambda: ops.matmul(At, B, transpose_a=True, force_tile=tile), 3)
            res[f"square_TN_16384x16384x256_tile{tile}"] = round(2.0 * m * m * 256 / t * 1e-9, 2)
        t = timed(lambda: torch.matmul(At.T, B), 3)
        res["square_cublas"] = round(2.0 * m * m * 256 / t * 1e-9, 2)
    print(json.dumps(res, indent=1))
