"""TEST INFRASTRUCTURE -- a CPU torch stand-in for the tensor-level wrappers of `acetn_b200.ops` (one function per C-ABI
entry point, same arguments / returns), so that the HOST logic above the C ABI (renormalization.py, evolution.py,
measurement.py, distributed.py, integration.py -- schedules, index descriptors, reference-object handling, routing) can be
driven end to end in a container that has no GPU.  Nothing in the product imports this; the product path still raises
without a CUDA device (tests/test_cabi_cpu.py::test_no_cpu_fallback).  Activated only through `emulated()`.

Arithmetic: the oracle's / plain torch float64; `gemm_ex` evaluates the two-level index descriptors
{div, s_hi, s_lo} literally (offset tables), so descriptor mistakes in the host code show up as wrong numbers.
"""
import contextlib
import os

import torch

from oracle import ctmrg_oracle as orc


class _NullStream:
    cuda_stream = 0

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


def _require(*tensors):
    for t in tensors:
        if t.dtype != torch.float64:
            raise RuntimeError(f"acetn_b200: only float64 is supported on the b200 backend (got {t.dtype})")
    return tensors[0].device


def _offsets(n, div, s_hi, s_lo):
    i = torch.arange(n, dtype=torch.int64)
    if div:
        return (i // div) * s_hi + (i % div) * s_lo
    return i * s_lo


def gemm_ex(M, N, K, batch, A, B, C, idx, alpha=1.0, beta=0.0, force_tile=0, force_splitk=0):
    g = [tuple(int(v) for v in idx[3 * i:3 * i + 3]) for i in range(9)]
    am, ak, ab, bk, bn, bb, cm, cn, cb = [_offsets(n, *t) for n, t in zip((M, K, batch, K, N, batch, M, N, batch), g)]

    def flat(t):
        # the operand as the library sees it: a base pointer + element offsets (works for any view with a storage behind it)
        return torch.as_strided(t, (t.untyped_storage().nbytes() // 8 - t.storage_offset(),), (1,))
    Af, Bf, Cf = flat(A), flat(B), flat(C)
    for b in range(batch):
        a = Af[(am[:, None] + ak[None, :] + ab[b]).reshape(-1)].reshape(M, K)
        bm = Bf[(bk[:, None] + bn[None, :] + bb[b]).reshape(-1)].reshape(K, N)
        ci = (cm[:, None] + cn[None, :] + cb[b]).reshape(-1)
        r = alpha * (a @ bm)
        if beta != 0.0:
            r = r + beta * Cf[ci].reshape(M, N)
        Cf[ci] = r.reshape(-1)
    return C


def quarter_tensor(C, E2, E1, A_view, normalize=True, stream=None, absmax=None, out=None, enc_storage=None):
    t = torch.einsum("ab,bcuU->acuU", C, E2)
    t = torch.einsum("acuU,ealL->cuUelL", t, E1)
    t = torch.einsum("cuUelL,LURDP->cuelRDP", t, A_view)
    t = torch.einsum("lurdp,cuelRDp->crRedD", A_view, t)
    shp = tuple(t.shape)
    Q = t.reshape(shp[0] * shp[1] * shp[2], -1)
    mx = Q.abs().max()
    if absmax is not None:
        absmax.fill_(float(mx))
    if normalize:
        Q = Q / mx
    if out is not None:
        o = out[:Q.numel()].view(Q.shape)
        o.copy_(Q)
        Q = o
    return Q.contiguous(), shp


def orthonormalize(Y):
    Y.copy_(torch.linalg.qr(Y).Q)
    return Y


def jacobi_svd(R, chi=None, cutoff=0.0):
    U, S, Vh = torch.linalg.svd(R)
    q = R.shape[0]
    kept = int((S / S[0] > cutoff).sum()) if S[0] > 0 else 0
    info = torch.tensor([min(q if chi is None else chi, kept), 1], dtype=torch.int32)
    return S, Vh.contiguous(), U.t().contiguous(), info


def rsvd(mats, omega, niter=2, reorth_adjoint=False, chi=None, cutoff=1e-12, stream=None, want_u=True, want_atq=False, encs=None, info=None):
    def fwd(Y):
        for M in reversed(mats):
            Y = M @ Y
        return Y

    def adj(Y):
        for M in mats:
            Y = M.t() @ Y
        return Y
    q = omega.shape[1]
    Y = fwd(omega)
    for _ in range(niter):
        Y = torch.linalg.qr(Y).Q
        Y = adj(Y)
        if reorth_adjoint:
            Y = torch.linalg.qr(Y).Q
        Y = fwd(Y)
    Qy = torch.linalg.qr(Y).Q
    AtQ = mats[0].t() @ Qy
    Bt = AtQ.t()
    for M in mats[1:]:
        Bt = Bt @ M
    Ub, S, Vh = torch.linalg.svd(Bt, full_matrices=False)
    kept = int((S / S[0] > cutoff).sum())
    info = torch.tensor([min(q if chi is None else chi, kept), 1], dtype=torch.int32)
    U = (Qy @ Ub) if want_u else None
    if want_atq:
        return U, S, Vh.t().contiguous(), info, AtQ.contiguous(), Ub.t().contiguous()
    return U, S, Vh.t().contiguous(), info


def projectors_from_usv(Q1, Q4, U, V, S, keep, stream=None, qmax1=None, qmax4=None, AtQ=None, Wt=None, enc1=None, enc4=None):
    w = 1.0 / torch.sqrt(S[:keep] / S[0])
    if AtQ is not None:
        p1 = AtQ @ (Wt.t()[:, :keep] * w)
    else:
        p1 = Q1.t() @ (U[:, :keep] * w)
    p2 = Q4 @ (V[:, :keep] * w)
    if qmax1 is not None:
        p1 = p1 / qmax1
    if qmax4 is not None:
        p2 = p2 / qmax4
    return p1.contiguous(), p2.contiguous()


def absorb_corner1(ci, ei, proj):
    return orc.absorb_corner1(ci, ei, proj).contiguous()


def absorb_corner2(ci, ei, proj):
    return orc.absorb_corner2(ci, ei, proj).contiguous()


def _edge_t3(ei, A_view, proj1):
    t = torch.einsum("ablL,buUx->alLuUx", ei, proj1)
    t = torch.einsum("LURDP,alLuUx->RDPalux", A_view, t)
    return torch.einsum("lurdp,RDpalux->adDxrR", A_view, t)


def absorb_edge(ei, A_view, proj2, proj1, normalize=True):
    t = torch.einsum("adDxrR,adDy->yxrR", _edge_t3(ei, A_view, proj1), proj2)
    return (t / t.norm() if normalize else t).contiguous()


def absorb_edge_begin(ei, A_view, proj1):
    xa, D = ei.shape[0], ei.shape[2]
    return _edge_t3(ei, A_view, proj1).reshape(xa * D * D, -1).contiguous()


def absorb_edge_finish(T3, proj2, D, normalize=True):
    xa, xy = proj2.shape[0], proj2.shape[3]
    t = torch.einsum("adDxrR,adDy->yxrR", T3.reshape(xa, D, D, -1, D, D), proj2)
    return (t / t.norm() if normalize else t).contiguous()


def permute_copy(view):
    return view.contiguous()


def absmax(x, out):
    out.fill_(max(float(out), float(x.abs().max())))
    return out


def frob_normalize(x):
    x.div_(x.norm())
    return x


def als_solve(a1r, a2r, n12g, n12, a12g, niter=100, tol=1e-15, epsilon=1e-12, method="cholesky"):
    a1, a2, it = orc.als_solve(a1r.clone(), a2r.clone(), n12g, n12, a12g, niter=niter, tol=tol, epsilon=epsilon, method=method)
    return a1.contiguous(), a2.contiguous(), torch.tensor([it, 0], dtype=torch.int32)


def i8_supported(rows, cols, q):
    return False


def _plain_site(st):
    return orc.Site(st['A'], list(st['C']), list(st['E']))


def site_rdm(C, E, A):
    return orc.site_rdm(orc.Cell(1, 1, {}, {(0, 0): orc.Site(A, list(C), list(E))}), (0, 0)).contiguous()


def bond_rdm(site1, site2, k):
    cell = orc.Cell(2, 1, {}, {(0, 0): _plain_site(site1), (1, 0): _plain_site(site2)})
    return orc.bond_rdm(cell, ((0, 0), (1, 0), k)).contiguous()


def norm_tensor(site1, site2, k, a1q, a2q):
    cell = orc.Cell(2, 1, {}, {(0, 0): _plain_site(site1), (1, 0): _plain_site(site2)})
    return orc.norm_tensor(cell, ((0, 0), (1, 0), k), a1q, a2q).contiguous()


_EMULATED = ["gemm_ex", "quarter_tensor", "orthonormalize", "jacobi_svd", "rsvd", "projectors_from_usv", "absorb_corner1",
             "absorb_corner2", "absorb_edge", "absorb_edge_begin", "absorb_edge_finish", "permute_copy", "absmax", "frob_normalize",
             "als_solve", "i8_supported", "site_rdm", "bond_rdm", "norm_tensor"]


@contextlib.contextmanager
def emulated():
    """Route acetn_b200.ops through the CPU stand-ins (sequential schedule: one stream).  ops.matmul / ops.contract keep
    their product implementation -- they are built on gemm_ex / permute_copy, so their descriptor logic is exercised."""
    from acetn_b200 import ops
    saved = {n: getattr(ops, n) for n in _EMULATED + ["_require_cuda", "require_cuda_device"]}
    env = {k: os.environ.get(k) for k in ("ACETN_B200_STREAMS", "ACETN_B200_STAGGER")}
    cur, sync = torch.cuda.current_stream, torch.cuda.synchronize
    launches = {"n": 0}

    def counted(fn):
        def wrap(*a, **kw):
            launches["n"] += 1
            return fn(*a, **kw)
        return wrap
    try:
        for n in _EMULATED:
            setattr(ops, n, counted(globals()[n]))
        ops._require_cuda = _require
        ops.require_cuda_device = lambda device: None
        os.environ["ACETN_B200_STREAMS"] = "1"
        os.environ["ACETN_B200_STAGGER"] = "0"
        torch.cuda.current_stream = lambda device=None: _NullStream()
        torch.cuda.synchronize = lambda device=None: None
        yield launches
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        torch.cuda.current_stream, torch.cuda.synchronize = cur, sync
        for k, v in env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
