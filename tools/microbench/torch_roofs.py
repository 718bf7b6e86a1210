"""Bench-only: cuBLAS DGEMM roof + cuSOLVER QR/SVD latencies on the GPU box (reference GPU torch path building blocks)."""
import torch, time, json
dev = torch.device("cuda")
def t(f, reps=5, warm=2):
    for _ in range(warm): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
res = {}
for (m, n, k) in [(8192, 8192, 8192), (16384, 258, 16384), (16384, 264, 16384), (16384, 16384, 256), (256, 256, 16384), (16384, 256, 16384)]:
    a = torch.randn(m, k, dtype=torch.float64, device=dev); b = torch.randn(k, n, dtype=torch.float64, device=dev)
    ms = t(lambda: a @ b)
    res[f"dgemm_{m}x{n}x{k}"] = {"ms": ms, "tflops": 2.0 * m * n * k / ms * 1e-9}
    if m == 16384 and k == 16384 and n == 258:
        ms = t(lambda: a.mH @ b)
        res[f"dgemm_T_{m}x{n}x{k}"] = {"ms": ms, "tflops": 2.0 * m * n * k / ms * 1e-9}
    del a, b
y = torch.randn(16384, 258, dtype=torch.float64, device=dev)
res["qr_16384x258_ms"] = t(lambda: torch.linalg.qr(y), reps=3, warm=1)
bt = torch.randn(258, 16384, dtype=torch.float64, device=dev)
res["svd_258x16384_ms"] = t(lambda: torch.linalg.svd(bt, full_matrices=False), reps=2, warm=1)
r = torch.randn(258, 258, dtype=torch.float64, device=dev)
res["svd_258x258_ms"] = t(lambda: torch.linalg.svd(r), reps=3, warm=1)
x = torch.randn(256, 8, 8, 256, 8, 8, dtype=torch.float64, device=dev)
aa = torch.randn(8, 8, 8, 8, 2, dtype=torch.float64, device=dev)
res["einsum_Qs3_ms"] = t(lambda: torch.einsum("cuUelL,LURDP->cuelRDP", x, aa), reps=2, warm=1)
x2 = torch.randn(256, 8, 256, 8, 8, 8, 2, dtype=torch.float64, device=dev)
res["einsum_Qs4_ms"] = t(lambda: torch.einsum("lurdp,cuelRDp->crRedD", aa, x2), reps=2, warm=1)
print(json.dumps(res, indent=1))
open("gpurun_out/torch_roofs.json", "w").write(json.dumps(res, indent=1))
