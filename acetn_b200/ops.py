"""Tensor-level wrappers around the C ABI: allocate outputs/workspace with torch, pass raw pointers, extents,
strides and the current CUDA stream.  No torch types cross the boundary."""
import contextlib
import ctypes
import math

import torch

from . import _lib

_workspaces = {}
_initialised = set()


def _require_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("acetn_b200: tensors must live on a CUDA (B200) device; there is no CPU path for backend='b200'")
        if t.dtype != torch.float64:
            raise RuntimeError(f"acetn_b200: only float64 is supported on the b200 backend (got {t.dtype})")
    dev = tensors[0].device
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _initialised:
        with torch.cuda.device(idx):
            _lib.check(_lib.load().acetn_b200_init(idx), "init")
        _initialised.add(idx)
    return dev


def require_cuda_device(device):
    """The host mirrors call this before scheduling work: backend='b200' has no CPU path."""
    if torch.device(device).type != "cuda":
        raise RuntimeError("acetn_b200: tensors must live on a CUDA (B200) device; there is no CPU path for backend='b200'")


_NULL_CTX = contextlib.nullcontext()


def _dev_index(dev):
    return dev.index if dev.index is not None else torch.cuda.current_device()


def _on(dev):
    """Context that makes `dev` the current CUDA device for the C call -- a no-op object when it already is (the common case: the
    torch.cuda.device context manager costs ~10 us, a third of the host time of a launch-bound sweep went into it and current_stream)."""
    idx = dev.index
    if idx is None or torch.cuda.current_device() == idx:
        return _NULL_CTX
    return torch.cuda.device(idx)


def _raw_stream(dev):
    """cudaStream_t (int) of torch's current stream on `dev`."""
    try:
        return torch._C._cuda_getCurrentRawStream(_dev_index(dev))
    except AttributeError:      # private torch API not available: the public (slower) route
        return torch.cuda.current_stream(dev).cuda_stream


def _ws(dev, nbytes, stream=None):
    """Per-(device, stream) grow-only scratch buffer (torch-allocated; the library owns no device memory)."""
    # one scratch buffer per CUDA stream: work enqueued on different streams may run concurrently
    sid = stream.cuda_stream if stream is not None else _raw_stream(dev)
    key = (dev.type, _dev_index(dev), sid)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            torch.cuda.synchronize(dev)      # rare (grow-only): kernels of any stream may still use the old buffer
        _workspaces[key] = None
        buf = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device=dev)
        _workspaces[key] = buf
    return buf


def release_workspace():
    _workspaces.clear()


def workspace_high_water(dev):
    """Largest scratch buffer this process has needed on `dev` so far (bytes)."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return max([b.numel() for k, b in _workspaces.items() if b is not None and k[1] == idx] + [0])


def reserve_workspace(dev, stream, nbytes):
    """Make sure the scratch buffer of (dev, stream) holds nbytes: CUDA-graph capture must not meet the grow path (which synchronises)."""
    return _ws(dev, nbytes, stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream(dev, stream=None):
    """cudaStream_t handed to the library: an explicit side stream, or torch's current stream."""
    if stream is not None:
        return ctypes.c_void_p(stream.cuda_stream)
    return ctypes.c_void_p(_raw_stream(dev))


_replayed_launches = 0


def launch_count():
    """Kernels of libacetn_b200.so launched so far: stream launches counted by the library + kernel nodes of replayed CUDA graphs."""
    return int(_lib.load().acetn_b200_launch_count()) + _replayed_launches


def note_replayed_launches(n):
    """A CUDA graph holding `n` kernels of the library was replayed (renormalization.MoveGraph)."""
    global _replayed_launches
    _replayed_launches += int(n)


def reset_launch_count():
    global _replayed_launches
    _replayed_launches = 0
    _lib.load().acetn_b200_reset_launch_count()


# ---------------------------------------------------------------------------------------------------------------
def gemm_ex(M, N, K, batch, A, B, C, idx, alpha=1.0, beta=0.0, force_tile=0, force_splitk=0):
    """Raw K1 call. idx: 27 ints = {div,s_hi,s_lo} for A(m,k,batch), B(k,n,batch), C(m,n,batch)."""
    dev = _require_cuda(A, B, C)
    lib = _lib.load()
    ix = _lib.i64_array(idx)
    nb = lib.acetn_b200_gemm_workspace_bytes(M, N, K, batch, ix, force_tile, force_splitk)
    ws = _ws(dev, nb)
    with _on(dev):
        st = lib.acetn_b200_gemm(M, N, K, batch, _p(A), _p(B), _p(C), ix, float(alpha), float(beta), force_tile, force_splitk,
                                 _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "gemm")
    return C


def matmul(A, B, transpose_a=False, force_tile=0, force_splitk=0, out=None):
    """C = op(A) @ B for 2-D row-major tensors through K1.  out: optional contiguous (M, N) destination (e.g. a row block
    of a larger row-major matrix)."""
    A = A.contiguous()
    B = B.contiguous()
    if transpose_a:
        K, M = A.shape
        a_idx = [0, 0, 1, 0, 0, M, 0, 0, 0]
    else:
        M, K = A.shape
        a_idx = [0, 0, K, 0, 0, 1, 0, 0, 0]
    N = B.shape[1]
    if out is None:
        C = torch.empty(M, N, dtype=A.dtype, device=A.device)
    else:
        if tuple(out.shape) != (M, N) or not out.is_contiguous():
            raise ValueError("matmul: out must be a contiguous (M, N) tensor")
        C = out
    idx = a_idx + [0, 0, N, 0, 0, 1, 0, 0, 0] + [0, 0, N, 0, 0, 1, 0, 0, 0]
    return gemm_ex(M, N, K, 1, A, B, C, idx, force_tile=force_tile, force_splitk=force_splitk)


def quarter_tensor(C, E2, E1, A_view, normalize=True, stream=None, absmax=None, out=None, enc_storage=None):
    """projectors.py:36-60.  C (xa,xb), E2 (xb,xc,D,D), E1 (xe,xa,D,D), A_view = bond_permute(k) (strided view).
    absmax: optional 1-element device tensor receiving max|Q| of the un-normalised tensor.
    out: optional flat FP64 device buffer (>= numel of Q) that receives Q (a task slot of the phase scheduler).
    enc_storage: uint8 device buffer (>= i8_encoded_bytes) -> the K7 residue encoding of Q is produced in the same call
    (acetn_b200_quarter_tensor_enc; requires normalize=False) and an I8Encoded is returned as third value."""
    dev = _require_cuda(C, E2, E1, A_view)
    C, E2, E1 = C.contiguous(), E2.contiguous(), E1.contiguous()
    xa, xb = C.shape
    xc, D = E2.shape[1], E2.shape[2]
    xe = E1.shape[0]
    d = A_view.shape[4]
    if E2.shape[0] != xb or E1.shape[1] != xa:
        raise ValueError(f"quarter_tensor: inconsistent chi legs C{tuple(C.shape)} E2{tuple(E2.shape)} E1{tuple(E1.shape)}")
    lib = _lib.load()
    if out is not None:
        Q = out[:xc * D * D * xe * D * D].view(xc * D * D, xe * D * D)
    else:
        Q = torch.empty(xc * D * D, xe * D * D, dtype=torch.float64, device=dev)
    nb = lib.acetn_b200_quarter_tensor_workspace_bytes(xa, xb, xc, xe, D, d)
    ws = _ws(dev, nb, stream)
    if enc_storage is not None:
        if normalize:
            raise ValueError("quarter_tensor: the K7 encoding is taken from the un-normalised tensor (normalize=False)")
        with _on(dev):
            st = lib.acetn_b200_quarter_tensor_enc(_p(C), _p(E2), _p(E1), _p(A_view), _lib.i64_array(A_view.stride()), xa, xb, xc, xe, D, d,
                                                   _p(Q), _p(absmax) if absmax is not None else None, _p(enc_storage), enc_storage.numel(),
                                                   _p(ws), ws.numel(), _stream(dev, stream))
        _lib.check(st, "quarter_tensor_enc")
        return Q, (xc, D, D, xe, D, D), I8Encoded(enc_storage, Q.shape[0], Q.shape[1])
    with _on(dev):
        st = lib.acetn_b200_quarter_tensor(_p(C), _p(E2), _p(E1), _p(A_view), _lib.i64_array(A_view.stride()), xa, xb, xc, xe, D, d,
                                           1 if normalize else 0, _p(Q), _p(absmax) if absmax is not None else None, _p(ws), ws.numel(),
                                           _stream(dev, stream))
    _lib.check(st, "quarter_tensor")
    return Q, (xc, D, D, xe, D, D)


def orthonormalize(Y):
    """In-place orthonormal basis of range(Y) (torch.linalg.qr(Y).Q up to column signs/rotations)."""
    dev = _require_cuda(Y)
    assert Y.dim() == 2 and Y.is_contiguous()
    m, q = Y.shape
    lib = _lib.load()
    nb = lib.acetn_b200_orthonormalize_workspace_bytes(m, q)
    ws = _ws(dev, nb)
    with _on(dev):
        st = lib.acetn_b200_orthonormalize(_p(Y), m, q, q, _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "orthonormalize")
    return Y


def jacobi_svd(R, chi=None, cutoff=0.0):
    """R (q,q) = Jt^T diag(S) Wt. Returns S, Wt, Jt, info(int32[2])."""
    dev = _require_cuda(R)
    R = R.contiguous()
    q = R.shape[0]
    lib = _lib.load()
    S = torch.empty(q, dtype=torch.float64, device=dev)
    Wt = torch.empty(q, q, dtype=torch.float64, device=dev)
    Jt = torch.empty(q, q, dtype=torch.float64, device=dev)
    info = torch.zeros(2, dtype=torch.int32, device=dev)
    nb = lib.acetn_b200_jacobi_svd_workspace_bytes(q)
    ws = _ws(dev, nb)
    with _on(dev):
        st = lib.acetn_b200_jacobi_svd(_p(R), q, _p(S), _p(Wt), _p(Jt), q if chi is None else chi, float(cutoff), _p(info),
                                       _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "jacobi_svd")
    return S, Wt, Jt, info


def rsvd(mats, omega, niter=2, reorth_adjoint=False, chi=None, cutoff=1e-12, stream=None, want_u=True, want_atq=False, encs=None, info=None):
    """Randomized SVD of mats[0] @ ... @ mats[-1] with the caller's test matrix omega (n, q).
    Returns U (m,q), S (q), V (n,q), info (int32[2] on device: [kept, jacobi sweeps]).
    want_atq=True additionally returns (AtQ, Wt) = (mats[0]^T Q, core left vectors as rows) and want_u=False skips U
    (then U is None): the half-system projector pipeline needs only AtQ, Wt and V.
    encs: optional list (one entry per factor) of I8Encoded / None; an encoded factor's thin products run on the INT8
    tensor cores (K7) and its FP64 entry in mats may be None."""
    nmat = len(mats)
    encs = list(encs) if encs is not None else [None] * nmat
    live = [m for m in mats if m is not None]
    dev = _require_cuda(*live, omega)
    mats = [m.contiguous() if m is not None else None for m in mats]
    omega = omega.contiguous()
    rows = [m.shape[0] if m is not None else e.rows for m, e in zip(mats, encs)]
    cols = [m.shape[1] if m is not None else e.cols for m, e in zip(mats, encs)]
    n, q = omega.shape
    if n != cols[-1]:
        raise ValueError("rsvd: omega rows must equal the column count of the last factor")
    m = rows[0]
    lib = _lib.load()
    U = torch.empty(m, q, dtype=torch.float64, device=dev) if want_u else None
    S = torch.empty(q, dtype=torch.float64, device=dev)
    V = torch.empty(n, q, dtype=torch.float64, device=dev)
    AtQ = torch.empty(cols[0], q, dtype=torch.float64, device=dev) if want_atq else None
    Wt = torch.empty(q, q, dtype=torch.float64, device=dev) if want_atq else None
    if info is None:
        info = torch.empty(2, dtype=torch.int32, device=dev)      # written by the library (no torch kernel on another stream)
    r_arr, c_arr = _lib.i64_array(rows), _lib.i64_array(cols)
    use = (ctypes.c_int32 * nmat)(*[1 if e is not None else 0 for e in encs])
    nb = lib.acetn_b200_rsvd_enc_workspace_bytes(nmat, r_arr, c_arr, q, use)
    ws = _ws(dev, nb, stream)
    ptrs = (ctypes.c_void_p * nmat)(*[t.data_ptr() if t is not None else None for t in mats])
    eptrs = (ctypes.c_void_p * nmat)(*[e.storage.data_ptr() if e is not None else None for e in encs])
    with _on(dev):
        st = lib.acetn_b200_rsvd_enc(nmat, ptrs, eptrs, r_arr, c_arr, _p(omega), q, int(niter), 1 if reorth_adjoint else 0,
                                     q if chi is None else int(chi), float(cutoff), _p(U) if want_u else None, _p(S), _p(V), _p(info),
                                     _p(AtQ) if want_atq else None, _p(Wt) if want_atq else None, _p(ws), ws.numel(),
                                     _stream(dev, stream))
    _lib.check(st, "rsvd")
    if want_atq:
        return U, S, V, info, AtQ, Wt
    return U, S, V, info


def projectors_from_usv(Q1, Q4, U, V, S, keep, stream=None, qmax1=None, qmax4=None, AtQ=None, Wt=None, enc1=None, enc4=None):
    """projectors.py:166-173. Q1 (m1,n1), Q4 (m4,n4), U (m1,q), V (n4,q). Returns proj1 (n1,keep), proj2 (m4,keep).
    qmax1/qmax4: max|Q| scalars of un-normalised Q1/Q4 (see acetn_b200.h).  enc1/enc4: I8Encoded Q1/Q4 (K7); the FP64
    tensor of an encoded factor may be None."""
    dev = _require_cuda(V, S)
    m1, n1 = Q1.shape if Q1 is not None else (enc1.rows, enc1.cols)
    m4, n4 = Q4.shape if Q4 is not None else (enc4.rows, enc4.cols)
    lib = _lib.load()
    p1 = torch.empty(n1, keep, dtype=torch.float64, device=dev)
    p2 = torch.empty(m4, keep, dtype=torch.float64, device=dev)
    use_enc = enc1 is not None or enc4 is not None
    nb = lib.acetn_b200_projectors_enc_workspace_bytes(m1, n1, m4, n4, keep, 1 if use_enc else 0)
    ws = _ws(dev, nb, stream)

    def ptr(t):
        return _p(t) if t is not None else None

    with _on(dev):
        st = lib.acetn_b200_projectors_from_usv_enc(ptr(Q1), ptr(enc1.storage) if enc1 is not None else None, m1, n1,
                                                    ptr(Q4), ptr(enc4.storage) if enc4 is not None else None, m4, n4,
                                                    ptr(U), U.stride(0) if U is not None else 0, _p(V), V.stride(0), _p(S), keep,
                                                    ptr(qmax1), ptr(qmax4), ptr(AtQ), ptr(Wt), Wt.shape[0] if Wt is not None else 0,
                                                    _p(p1), _p(p2), _p(ws), ws.numel(), _stream(dev, stream))
    _lib.check(st, "projectors_from_usv")
    return p1, p2


def absorb_corner1(ci, ei, proj):
    """directional_mover.py:306-323 : out[a,x]."""
    dev = _require_cuda(ci, ei, proj)
    ci, ei, proj = ci.contiguous(), ei.contiguous(), proj.contiguous()
    xa, xb, D = ei.shape[0], ei.shape[1], ei.shape[2]
    xc, xx = ci.shape[1], proj.shape[3]
    if ci.shape[0] != xb or proj.shape[0] != xc:
        raise ValueError("absorb_corner1: inconsistent chi legs")
    lib = _lib.load()
    out = torch.empty(xa, xx, dtype=torch.float64, device=dev)
    ws = _ws(dev, lib.acetn_b200_absorb_corner_workspace_bytes(xa, xb, xc, xx, D))
    with _on(dev):
        st = lib.acetn_b200_absorb_corner1(_p(ci), _p(ei), _p(proj), xa, xb, xc, xx, D, _p(out), _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "absorb_corner1")
    return out


def absorb_corner2(ci, ei, proj):
    """directional_mover.py:325-343 : out[x,c]."""
    dev = _require_cuda(ci, ei, proj)
    ci, ei, proj = ci.contiguous(), ei.contiguous(), proj.contiguous()
    xa, xb = ci.shape
    xc, D = ei.shape[1], ei.shape[2]
    xx = proj.shape[3]
    if ei.shape[0] != xb or proj.shape[0] != xa:
        raise ValueError("absorb_corner2: inconsistent chi legs")
    lib = _lib.load()
    out = torch.empty(xx, xc, dtype=torch.float64, device=dev)
    ws = _ws(dev, lib.acetn_b200_absorb_corner_workspace_bytes(xa, xb, xc, xx, D))
    with _on(dev):
        st = lib.acetn_b200_absorb_corner2(_p(ci), _p(ei), _p(proj), xa, xb, xc, xx, D, _p(out), _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "absorb_corner2")
    return out


def absorb_edge(ei, A_view, proj2, proj1, normalize=True):
    """directional_mover.py:345-366 : out[y,x,r,R].  normalize=False: un-normalised (partial) result."""
    dev = _require_cuda(ei, A_view, proj2, proj1)
    ei, proj1, proj2 = ei.contiguous(), proj1.contiguous(), proj2.contiguous()
    xa, xb, D = ei.shape[0], ei.shape[1], ei.shape[2]
    d = A_view.shape[4]
    xx, xy = proj1.shape[3], proj2.shape[3]
    if proj1.shape[0] != xb or proj2.shape[0] != xa:
        raise ValueError("absorb_edge: inconsistent chi legs")
    lib = _lib.load()
    out = torch.empty(xy, xx, D, D, dtype=torch.float64, device=dev)
    ws = _ws(dev, lib.acetn_b200_absorb_edge_workspace_bytes(xa, xb, xx, xy, D, d))
    with _on(dev):
        st = lib.acetn_b200_absorb_edge(_p(ei), _p(A_view), _lib.i64_array(A_view.stride()), _p(proj2), _p(proj1), xa, xb, xx, xy, D, d,
                                        1 if normalize else 0, _p(out), _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "absorb_edge")
    return out


def absorb_edge_begin(ei, A_view, proj1):
    """First stage of absorb_edge (needs only proj1 of the neighbouring task): T3[a,(d,Dd),x,(r,R)], see include/acetn_b200.h."""
    dev = _require_cuda(ei, A_view, proj1)
    ei, proj1 = ei.contiguous(), proj1.contiguous()
    xa, xb, D = ei.shape[0], ei.shape[1], ei.shape[2]
    d = A_view.shape[4]
    xx = proj1.shape[3]
    if proj1.shape[0] != xb:
        raise ValueError("absorb_edge_begin: inconsistent chi legs")
    lib = _lib.load()
    T3 = torch.empty(xa * D * D, xx * D * D, dtype=torch.float64, device=dev)
    ws = _ws(dev, lib.acetn_b200_absorb_edge_begin_workspace_bytes(xa, xb, xx, D, d))
    with _on(dev):
        st = lib.acetn_b200_absorb_edge_begin(_p(ei), _p(A_view), _lib.i64_array(A_view.stride()), _p(proj1), xa, xb, xx, D, d, _p(T3),
                                              _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "absorb_edge_begin")
    return T3


def absorb_edge_finish(T3, proj2, D, normalize=True):
    """Second stage of absorb_edge: out[y,x,r,R] = sum proj2[a,d,Dd,y] T3[a,(d,Dd),x,(r,R)] (+ Frobenius normalisation)."""
    dev = _require_cuda(T3, proj2)
    proj2 = proj2.contiguous()
    xa, xy = proj2.shape[0], proj2.shape[3]
    xx = T3.shape[1] // (D * D)
    if T3.shape[0] != xa * D * D:
        raise ValueError("absorb_edge_finish: inconsistent chi legs")
    lib = _lib.load()
    out = torch.empty(xy, xx, D, D, dtype=torch.float64, device=dev)
    ws = _ws(dev, lib.acetn_b200_absorb_edge_finish_workspace_bytes(xa, xx, xy, D))
    with _on(dev):
        st = lib.acetn_b200_absorb_edge_finish(_p(proj2), _p(T3), xa, xx, xy, D, 1 if normalize else 0, _p(out), _p(ws), ws.numel(),
                                               _stream(dev))
    _lib.check(st, "absorb_edge_finish")
    return out


# ---------------------------------------------------------------------------------------------------------------
# generic pairwise contraction (transpose-transpose-GEMM): used by the measure / norm-tensor paths
# ---------------------------------------------------------------------------------------------------------------
def permute_copy(view):
    """Contiguous copy of an arbitrarily strided <= 8-d view, by the library's gather kernel."""
    dev = _require_cuda(view)
    out = torch.empty(view.shape, dtype=view.dtype, device=dev)
    if out.numel() == 0:
        return out
    # merge nothing, just pass dims/strides (drop size-1 dims to stay within 8)
    dims = [s for s in view.shape if s != 1] or [1]
    strides = [st for s, st in zip(view.shape, view.stride()) if s != 1] or [1]
    if len(dims) > 8:
        raise RuntimeError("permute_copy: more than 8 non-trivial dims")
    with _on(dev):
        st = _lib.load().acetn_b200_permute(_p(out), _p(view), len(dims), _lib.i64_array(dims), _lib.i64_array(strides), _stream(dev))
    _lib.check(st, "permute")
    return out


def _as_layout(t, labels, first, second):
    """Return (tensor, order) with t laid out contiguously as [batch.., first.., second..] or [batch.., second.., first..]
    without a copy when its memory already has one of the two orders; otherwise one permute copy into the first."""
    def perm(order):
        return t.permute([labels.index(c) for c in order])
    for tag, order in (("fs", first[0] + first[1] + second), ("sf", first[0] + second + first[1])):
        v = perm(order)
        if v.is_contiguous():
            return v, tag
    return permute_copy(perm(first[0] + first[1] + second)), "fs"


def contract(spec, A, B):
    """C = einsum(spec, A, B) for a pairwise contraction, executed as (at most one gather per operand) + one batched K1
    DGEMM.  Returns a (possibly permuted) view with the requested output leg order."""
    lhs, out = spec.replace(" ", "").split("->")
    sa, sb = lhs.split(",")
    dev = _require_cuda(A, B)
    ext = {}
    for labels, t in ((sa, A), (sb, B)):
        if len(labels) != t.dim():
            raise ValueError(f"contract: spec {spec} does not match operand rank {t.dim()}")
        for c, n in zip(labels, t.shape):
            if ext.setdefault(c, n) != n:
                raise ValueError(f"contract: extent mismatch on leg {c} in {spec}")
    batch = "".join(c for c in out if c in sa and c in sb)
    K = "".join(c for c in sa if c in sb and c not in out)
    M = "".join(c for c in out if c in sa and c not in sb)
    N = "".join(c for c in out if c in sb and c not in sa)
    if set(batch + M + N) != set(out) or set(batch + M + K) != set(sa) or set(batch + K + N) != set(sb):
        raise ValueError(f"contract: unsupported spec {spec} (traces / outer sums are not handled)")
    prod = lambda s: math.prod(ext[c] for c in s)   # noqa: E731
    nb, m, n, k = prod(batch), prod(M), prod(N), prod(K)
    Av, ta = _as_layout(A, sa, (batch, M), K)        # [b, M, K] ("fs") or [b, K, M] ("sf")
    Bv, tb = _as_layout(B, sb, (batch, K), N)        # [b, K, N] ("fs") or [b, N, K] ("sf")
    C = torch.empty([ext[c] for c in batch + M + N], dtype=A.dtype, device=dev)
    a_idx = ([0, 0, k, 0, 0, 1] if ta == "fs" else [0, 0, 1, 0, 0, m]) + [0, 0, m * k if nb > 1 else 0]
    b_idx = ([0, 0, n, 0, 0, 1] if tb == "fs" else [0, 0, 1, 0, 0, k]) + [0, 0, k * n if nb > 1 else 0]
    c_idx = [0, 0, n, 0, 0, 1, 0, 0, m * n if nb > 1 else 0]
    gemm_ex(m, n, k, nb, Av, Bv, C, a_idx + b_idx + c_idx)
    order = batch + M + N
    return C.permute([order.index(c) for c in out])


# ---------------------------------------------------------------------------------------------------------------
# environment contractions behind the C ABI (acetn_b200/csrc/environment.cu)
# ---------------------------------------------------------------------------------------------------------------
def _chi2(*tensors):
    out = []
    for t in tensors:
        out += [t.shape[0], t.shape[1]]
    return out


def site_rdm(C, E, A):
    """RDM.build_site_rdm (rdm.py:35-67): C, E = the four corners / edges of the site, A = site['A'] (any strides) -> rho (d,d)."""
    dev = _require_cuda(*C, *E, A)
    C = [c.contiguous() for c in C]
    E = [e.contiguous() for e in E]
    D, d = A.shape[0], A.shape[4]
    chi = _lib.i64_array(_chi2(*C, *E))
    lib = _lib.load()
    rho = torch.empty(d, d, dtype=torch.float64, device=dev)
    ws = _ws(dev, lib.acetn_b200_site_rdm_workspace_bytes(chi, D, d))
    with _on(dev):
        st = lib.acetn_b200_site_rdm(_p(C[0]), _p(C[1]), _p(C[2]), _p(C[3]), _p(E[0]), _p(E[1]), _p(E[2]), _p(E[3]), _p(A),
                                     _lib.i64_array(A.stride()), chi, D, d, _p(rho), _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "site_rdm")
    return rho


def _bond_boundary(a, b, k):
    """The ten boundary tensors of the bond (s1, s2, k) in the argument order of acetn_b200_bond_rdm / _norm_tensor
    (rdm.py:84-96, full_update.py:184-196); a, b = the two site-tensor objects."""
    return [a['C'][(k + 1) % 4], a['E'][(k + 1) % 4], a['E'][k % 4], a['C'][(k + 2) % 4], a['E'][(k + 2) % 4],
            b['C'][k % 4], b['E'][k % 4], b['E'][(k + 3) % 4], b['C'][(k + 3) % 4], b['E'][(k + 2) % 4]]


def bond_rdm(site1, site2, k):
    """RDM.build_bond_rdm (rdm.py:69-154) of the bond (s1, s2, k): site1 / site2 = the reference's (or acetn_b200's) SiteTensor
    objects -> rho (d,d,d,d) [P,Q,p,q]."""
    a1, a2 = site1.bond_permute(k), site2.bond_permute(k)
    bt = [t.contiguous() for t in _bond_boundary(site1, site2, k)]
    dev = _require_cuda(*bt, a1, a2)
    D, d = a1.shape[0], a1.shape[4]
    chi = _lib.i64_array(_chi2(*bt))
    lib = _lib.load()
    rho = torch.empty(d, d, d, d, dtype=torch.float64, device=dev)
    ws = _ws(dev, lib.acetn_b200_bond_rdm_workspace_bytes(chi, D, d))
    with _on(dev):
        st = lib.acetn_b200_bond_rdm(_p(bt[0]), _p(bt[1]), _p(bt[2]), _p(bt[3]), _p(bt[4]), _p(a1), _lib.i64_array(a1.stride()),
                                     _p(bt[5]), _p(bt[6]), _p(bt[7]), _p(bt[8]), _p(bt[9]), _p(a2), _lib.i64_array(a2.stride()),
                                     chi, D, d, _p(rho), _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "bond_rdm")
    return rho


def norm_tensor(site1, site2, k, a1q, a2q):
    """build_norm_tensor (full_update.py:163-227) of the bond (s1, s2, k) with the QR-reduced site factors a1q, a2q (D,D,D,nD)
    -> N12 (nD,nD,nD,nD) [y,x,Y,X]."""
    bt = [t.contiguous() for t in _bond_boundary(site1, site2, k)]
    a1q, a2q = a1q.contiguous(), a2q.contiguous()
    dev = _require_cuda(*bt, a1q, a2q)
    D, nD = a1q.shape[0], a1q.shape[3]
    chi = _lib.i64_array(_chi2(*bt))
    lib = _lib.load()
    n12 = torch.empty(nD, nD, nD, nD, dtype=torch.float64, device=dev)
    ws = _ws(dev, lib.acetn_b200_norm_tensor_workspace_bytes(chi, D, nD))
    with _on(dev):
        st = lib.acetn_b200_norm_tensor(_p(bt[0]), _p(bt[1]), _p(bt[2]), _p(bt[3]), _p(bt[4]), _p(a1q), _p(bt[5]), _p(bt[6]), _p(bt[7]),
                                        _p(bt[8]), _p(bt[9]), _p(a2q), chi, D, nD, _p(n12), _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "norm_tensor")
    return n12


def absmax(x, out):
    """out[0] = max(out[0], max|x|)  (out: 1-element device tensor, zero it first)."""
    dev = _require_cuda(x, out)
    x = x.contiguous()
    with _on(dev):
        st = _lib.load().acetn_b200_absmax(_p(x), x.numel(), _p(out), _stream(dev))
    _lib.check(st, "absmax")
    return out


def frob_normalize(x):
    """x /= ||x||_F in place (deterministic two-stage reduction)."""
    dev = _require_cuda(x)
    assert x.is_contiguous()
    ws = _ws(dev, 8192 * 8)
    with _on(dev):
        st = _lib.load().acetn_b200_frob_normalize(_p(x), x.numel(), _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "frob_normalize")
    return x


ALS_METHODS = {"cholesky": 0, "pinv": 1}


def als_solve(a1r, a2r, n12g, n12, a12g, niter=100, tol=1e-15, epsilon=1e-12, method="cholesky"):
    """als_solver.py:55-82 / csrc/evolution/als_solve.cpp:107-137, method "cholesky" or "pinv" (als_solver.py:218-229).  Returns
    (a1r, a2r, info) with info = device int32[2] {iterations run, non-positive Cholesky pivots}."""
    if method not in ALS_METHODS:
        raise ValueError(f"Invalid als_method: {method} provided.")
    dev = _require_cuda(a1r, a2r, n12g, n12, a12g)
    a1 = a1r.contiguous().clone()
    a2 = a2r.contiguous().clone()
    n12g, n12, a12g = n12g.contiguous(), n12.contiguous(), a12g.contiguous()
    nD, bD, pD = a1.shape
    lib = _lib.load()
    info = torch.zeros(2, dtype=torch.int32, device=dev)
    ws = _ws(dev, lib.acetn_b200_als_workspace_bytes(nD, bD, pD))
    with _on(dev):
        st = lib.acetn_b200_als_solve(_p(a1), _p(a2), _p(n12g), _p(n12), _p(a12g), nD, bD, pD, int(niter), float(tol), float(epsilon),
                                      ALS_METHODS[method], _p(info), _p(ws), ws.numel(), _stream(dev))
    _lib.check(st, "als_solve")
    return a1, a2, info


# ---------------------------------------------------------------------------------------------------------------
# K7: INT8 tensor-core exact products (acetn_b200/csrc/i8crt.cu)
class I8Encoded:
    """A big FP64 matrix encoded once as 16 planes of int8 residues (+ row/column exponents); opaque device storage."""

    def __init__(self, storage, rows, cols):
        self.storage, self.rows, self.cols = storage, rows, cols


def i8_supported(rows, cols, q):
    return bool(_lib.load().acetn_b200_i8_supported(rows, cols, q))


def i8_encoded_bytes(rows, cols):
    return int(_lib.load().acetn_b200_i8_encoded_bytes(rows, cols))


def i8_encode(Q, stream=None, storage=None):
    dev = _require_cuda(Q)
    assert Q.dim() == 2 and Q.stride(1) == 1
    rows, cols = Q.shape
    lib = _lib.load()
    nb = lib.acetn_b200_i8_encoded_bytes(rows, cols)
    if storage is None or storage.numel() < nb:
        storage = torch.empty(nb, dtype=torch.uint8, device=dev)
    with _on(dev):
        st = lib.acetn_b200_i8_encode(_p(Q), rows, cols, Q.stride(0), _p(storage), storage.numel(), _stream(dev, stream))
    _lib.check(st, "i8_encode")
    return I8Encoded(storage, rows, cols)


def i8_matmul(enc, Y, adjoint=False, stream=None, out=None):
    """out = Q @ Y (adjoint=False) or Q.T @ Y (adjoint=True) for an I8Encoded Q; FP64 in/out, integer arithmetic inside."""
    dev = _require_cuda(Y)
    Y = Y.contiguous()
    k, q = Y.shape
    m = enc.cols if adjoint else enc.rows
    if k != (enc.rows if adjoint else enc.cols):
        raise ValueError("i8_matmul: inner dimensions differ")
    lib = _lib.load()
    if out is None:
        out = torch.empty(m, q, dtype=torch.float64, device=dev)
    nb = lib.acetn_b200_i8_matmul_workspace_bytes(enc.rows, enc.cols, q)
    ws = _ws(dev, nb, stream)
    with _on(dev):
        st = lib.acetn_b200_i8_matmul(_p(enc.storage), enc.rows, enc.cols, 1 if adjoint else 0, _p(Y), q, q, _p(out), out.stride(0),
                                      _p(ws), ws.numel(), _stream(dev, stream))
    _lib.check(st, "i8_matmul")
    return out
