"""Dev tool: stage-by-stage check of the INT8 tensor-core exact product (K7, acetn_b200/csrc/i8crt.cu) on a B200.

Small shapes: every stage (exponents, residues of both operands, the per-modulus INT8 GEMM, CRT reconstruction) is compared
bit-for-bit with an integer restatement in torch/python.  Full size: result vs FP64 matmul + timings vs the DMMA GEMM (K1)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import _lib, ops

MODS = [256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193]
NONE = -1000000
dev = torch.device("cuda", 0)


def bal(x, m):
    lo = -(m // 2)
    r = torch.remainder(x, m)
    return torch.where(r > lo + m - 1, r - m, r)


def expo(x):
    """e with |x| < 2^e (frexp exponent); NONE for zeros"""
    _, e = torch.frexp(x)
    return torch.where(x == 0, torch.full_like(e, NONE), e).to(torch.int64)


def operand_bits(kmax):
    lg = 0
    while (1 << lg) < kmax:
        lg += 1
    return min(54, int((125.375636 - 2.0 - lg) / 2.0))


def ld128(n):
    return (n + 127) // 128 * 128


def r256(n):
    return (n + 255) // 256 * 256


def check_small(rows, cols, q, adjoint, seed, graded=False):
    tag = f"[{rows}x{cols} q={q} {'TN' if adjoint else 'NN'}{' graded' if graded else ''}]"
    g = torch.Generator().manual_seed(seed)
    Q = torch.randn(rows, cols, dtype=torch.float64, generator=g)
    if graded:
        Q = Q * torch.logspace(0, -9, rows, dtype=torch.float64)[:, None] * torch.logspace(0, -7, cols, dtype=torch.float64)[None, :]
        Q[3, :] = 0.0
        Q[:, 5] = 0.0
    k = rows if adjoint else cols
    m_out = cols if adjoint else rows
    Y = torch.randn(k, q, dtype=torch.float64, generator=g)
    P = operand_bits(max(rows, cols))
    ok = True
    # ---- reference encoding -------------------------------------------------------------------------------------
    re = expo(Q).amax(dim=1)
    ce = (expo(Q) - re[:, None]).masked_fill(Q == 0, NONE).amax(dim=0)
    sh = P - re[:, None] - ce[None, :]
    Ai = torch.where((re[:, None] > NONE) & (ce[None, :] > NONE), torch.ldexp(Q, sh.clamp(-4000, 4000).to(torch.int32)).round(), torch.zeros_like(Q)).to(torch.int64)
    rowshift = re if adjoint else ce
    Ys = torch.ldexp(Y, rowshift.clamp(-4000, 4000).to(torch.int32)[:, None])
    Ys = torch.where(rowshift[:, None] > NONE, Ys, torch.zeros_like(Ys))
    ez = expo(Ys).amax(dim=0)
    Bi = torch.where(ez[None, :] > NONE, torch.ldexp(Ys, (P - ez).clamp(-4000, 4000).to(torch.int32)[None, :]).round(), torch.zeros_like(Ys)).to(torch.int64)
    # ---- device ----------------------------------------------------------------------------------------------------
    Qd, Yd = Q.to(dev), Y.to(dev)
    enc = ops.i8_encode(Qd)
    torch.cuda.synchronize()
    ld = ld128(cols)
    st = enc.storage.cpu()
    off = 0
    res = st[off:off + 16 * rows * ld].view(16, rows, ld); off += r256(16 * rows * ld)      # uint8: A residues in [0, m)
    rexp = st[off:off + rows * 4].view(torch.int32); off += r256(rows * 4)
    cexp = st[off:off + cols * 4].view(torch.int32)
    e1 = bool((rexp.to(torch.int64) == re).all()); e2 = bool((cexp.to(torch.int64) == ce).all())
    print(tag, "rowexp", "PASS" if e1 else "FAIL", "colexp", "PASS" if e2 else "FAIL")
    ok &= e1 and e2
    bad = 0
    for l, m in enumerate(MODS):
        bad += int((res[l, :, :cols].to(torch.int64) != torch.remainder(Ai, m)).sum())
        bad += int((res[l, :, cols:] != 0).sum())
    print(tag, "A residues", "PASS" if bad == 0 else f"FAIL ({bad} mismatches)")
    ok &= bad == 0
    out = ops.i8_matmul(enc, Yd, adjoint=adjoint)
    torch.cuda.synchronize()
    npad = (q + 15) // 16 * 16
    ldk = ld128(k)
    nb = _lib.load().acetn_b200_i8_matmul_workspace_bytes(rows, cols, q)
    ws = ops._ws(dev, nb).cpu()
    off = 0
    Bres = ws[off:off + 16 * npad * ldk].view(torch.int8).view(16, npad, ldk); off += r256(16 * npad * ldk)
    Cres = ws[off:off + 16 * m_out * npad].view(torch.int8).view(16, m_out, npad); off += r256(16 * m_out * npad)
    ezd = ws[off:off + npad * 4].view(torch.int32)
    e3 = bool((ezd[:q].to(torch.int64) == ez).all())
    print(tag, "thin colexp", "PASS" if e3 else "FAIL")
    ok &= e3
    bad = 0
    for l, m in enumerate(MODS):
        bad += int((Bres[l, :q, :k].to(torch.int64) != bal(Bi, m).T).sum())
        bad += int((Bres[l, q:, :] != 0).sum()) + int((Bres[l, :, k:] != 0).sum())
    print(tag, "B residues", "PASS" if bad == 0 else f"FAIL ({bad} mismatches)")
    ok &= bad == 0
    # per-modulus GEMM, from the DEVICE residues (isolates the tensor-core kernel)
    bad = 0
    first = None
    for l, m in enumerate(MODS):
        Al = res[l, :, :cols].to(torch.int64)
        Bl = Bres[l, :, :k].to(torch.int64)
        Cl = (Al.T if adjoint else Al) @ Bl.T
        exp = bal(Cl, m)
        got = Cres[l].to(torch.int64)
        neq = got != exp
        nb_ = int(neq.sum())
        if nb_ and first is None:
            idx = neq.nonzero()[:6]
            first = [(l, int(i), int(z), int(got[i, z]), int(exp[i, z])) for i, z in idx]
            colsbad = neq.any(dim=0).nonzero().flatten().tolist()
            rowsbad = neq.any(dim=1).nonzero().flatten().tolist()
            print(tag, f"  plane {l}: {nb_} bad; bad cols {colsbad[:8]}..{colsbad[-3:]} ({len(colsbad)}), bad rows {rowsbad[:8]}..{rowsbad[-3:]} ({len(rowsbad)})")
        bad += nb_
    print(tag, "INT8 GEMM", "PASS" if bad == 0 else f"FAIL ({bad} mismatches) first (l,i,z,got,exp): {first}")
    ok &= bad == 0
    # final result: exact integer product, scaled
    Cx = (Ai.T if adjoint else Ai).to(torch.float64)  # not exact for big ints; use python ints on a sample instead
    ref = (Q.T if adjoint else Q) @ Y
    o = out.cpu()
    scale = ref.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    err = ((o - ref).abs() / scale).max().item()
    # exact check on a sample of entries with python integers
    eo = ce if adjoint else re
    worst = 0.0
    for (i, z) in [(0, 0), (1, q - 1), (m_out - 1, 0), (m_out // 2, q // 2), (7, 3)]:
        a = (Ai[:, i] if adjoint else Ai[i, :]).tolist()
        b = Bi[:, z].tolist()
        x = sum(int(u) * int(v) for u, v in zip(a, b))
        if eo[i] > NONE and ez[z] > NONE:
            import math
            val = math.ldexp(float(x), int(eo[i]) + int(ez[z]) - 2 * P) if x != 0 else 0.0
        else:
            val = 0.0
        d = abs(o[i, z].item() - val)
        worst = max(worst, d / (abs(val) + 1e-300) if val != 0 else d)
    print(tag, f"CRT sample rel err vs exact integers {worst:.2e}; max |out-ref|/rowmax(ref) = {err:.2e}", "PASS" if worst < 1e-15 and err < 1e-13 else "FAIL")
    ok &= worst < 1e-15 and err < 1e-13
    return ok


def check_full(n=16384, q=258):
    torch.manual_seed(0)
    Q = torch.rand(n, n, dtype=torch.float64, device=dev) - 0.3
    Y = torch.randn(n, q, dtype=torch.float64, device=dev)
    res = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    enc = ops.i8_encode(Q)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(); enc = ops.i8_encode(Q, storage=enc.storage); e1.record(); torch.cuda.synchronize()
    res["encode_ms"] = e0.elapsed_time(e1)
    for adj in (False, True):
        ref = (Q.T if adj else Q) @ Y
        out = ops.i8_matmul(enc, Y, adjoint=adj)
        torch.cuda.synchronize()
        scale = ref.abs().amax(dim=1, keepdim=True)
        res[f"err_rowmax_{'TN' if adj else 'NN'}"] = ((out - ref).abs() / scale).max().item()
        res[f"err_fro_{'TN' if adj else 'NN'}"] = ((out - ref).norm() / ref.norm()).item()
        d = ops.matmul(Q, Y, transpose_a=adj)
        res[f"dmma_err_fro_{'TN' if adj else 'NN'}"] = ((d - ref).norm() / ref.norm()).item()
        for name, fn in (("i8", lambda: ops.i8_matmul(enc, Y, adjoint=adj, out=out)), ("dmma", lambda: ops.matmul(Q, Y, transpose_a=adj, out=d))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record(); torch.cuda.synchronize()
            res[f"{name}_ms_{'TN' if adj else 'NN'}"] = e0.elapsed_time(e1) / 10
    print(json.dumps(res, indent=1))
    return res


if __name__ == "__main__":
    ok = True
    for (r, c, q, adj, gr) in [(256, 384, 40, False, False), (256, 384, 40, True, False), (384, 256, 258, False, False), (384, 256, 258, True, False),
                               (300, 200, 37, False, True), (300, 200, 37, True, True), (1024, 1024, 258, False, False), (1024, 1024, 258, True, False)]:
        try:
            ok &= check_small(r, c, q, adj, seed=r + c + q, graded=gr)
        except Exception as e:   # keep going: later stages may still be informative
            print("EXC", r, c, q, adj, repr(e))
            ok = False
    print("SMALL:", "ALL PASS" if ok else "FAILURES")
    if "--full" in sys.argv:
        check_full()
