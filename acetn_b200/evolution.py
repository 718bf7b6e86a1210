"""B200 counterparts of the full-update pieces: build_norm_tensor (acetn/evolution/full_update.py:163-227), the ALS inner
solver (acetn/evolution/als_solver.py) and -- SURVEY.md 8f-1 -- the callers around them in one bond update: QR split
(tensor_update.py:53-75), positive_approx / gauge_fix (full_update.py:262-343), finalize_reduced_tensors (:121-161).

Everything dense runs on the library's own kernels (no cuSOLVER): contractions = gather + batched K1 DGEMM (`ops.contract`),
thin QR = K4 Householder TSQR + one K1 product for R, symmetric eigen-decomposition / SVD / pseudo-inverse of the small
matrices (<= 256 x 256) = K5 one-sided Jacobi, ALS loop = K6.  `ALSSolver` keeps the reference's constructor/solve() shape so
it can replace `ALSSolver(n12, a12g, ar_shape, config).solve()` when `config.backend == "b200"` (als_solver.py:48-51);
`full_update_bond` is `FullUpdater.tensor_update` (full_update.py:34-63)."""
import torch

from . import ops
from .ops import contract


def _ix(*triples):
    """27 ints of ops.gemm_ex: {div, s_hi, s_lo} for A(m, k, batch), B(k, n, batch), C(m, n, batch); a plain int is a stride."""
    out = []
    for t in triples:
        out += [0, 0, t] if isinstance(t, int) else list(t)
    return out


def _norm_half(cA, eA, eB, cB, eC, aq, left):
    """One half of the norm tensor (full_update.py:205-218 / 220-227):
        right (left=False): n1[f,e,Y,y] = sum tmp_r2[a,f,d,D] ( ((c12 e12) e11) conj(a1q) a1q )[a,e,D,Y,d,y]
        left  (left=True) : n2[f,c,X,x] = sum tmp_l2[b,f,d,D] ( ((c21 e21) e24) conj(a2q) a2q )[c,b,X,D,x,d]
    Every leg permutation of the reference's einsum chain is folded into the two-level index descriptors of K1, so none of the
    2 - 8 GiB intermediates is ever re-laid out in HBM (the generic gather + GEMM form spent 42 % of its time in gathers)."""
    dev, dt = cA.device, cA.dtype
    x0, x1 = cA.shape                      # cA[a,b]
    xc, D = eA.shape[1], eA.shape[2]       # eA[b,c,.,.]
    xe = eB.shape[0]                       # eB[e,a,.,.]
    nD = aq.shape[3]
    D2 = D * D
    # 1. t1[a,(c,p,P)] = cA[a,b] eA[b,(c,p,P)]
    t1 = ops.matmul(cA.contiguous(), eA.reshape(x1, xc * D2))
    # 2. t[(c,p,P),(e,q,Q)] = sum_a t1[a,(c,p,P)] eB[e,a,(q,Q)]
    m = xc * D2
    n = xe * D2
    t = torch.empty(m, n, dtype=dt, device=dev)
    ops.gemm_ex(m, n, x0, 1, t1, eB.contiguous(), t, _ix(1, m, 0, D2, (D2, x0 * D2, 1), 0, n, 1, 0))
    del t1
    # 3. t2[c,p,e,q,(n1,n2)] = sum_{P,Q} t[c,p,P,e,q,Q] aq[..]   (K = the two bra legs)
    #    right: conj(a1q)[R,D,U,Y]: k=(R,U) n=(D,Y)      left: conj(a2q)[D,L,U,X]: k=(U,L) n=(X,D)
    M3 = xc * D * xe * D
    N3 = D * nD
    t2 = torch.empty(M3, N3, dtype=dt, device=dev)
    a_m = (xe * D, D * n, D)               # hi = (c,p): stride D*n (= one p step), lo = (e,q): stride D
    a_k = (D, n, 1)                        # hi = P: stride n, lo = Q: stride 1
    # the 64 KiB site factor as a plain (k, n) matrix for both steps (conj is the identity: FP64 real):
    #   right: a1q[R,D,U,Y] -> [(R,U)][(D,Y)]      left: a2q[D,L,U,X] -> [(U,L)][(X,D)]
    bq = (aq.permute(0, 2, 1, 3) if not left else aq.permute(2, 1, 3, 0)).contiguous()
    ops.gemm_ex(M3, N3, D2, 1, t, bq, t2, _ix(a_m, a_k, 0, N3, 1, 0, N3, 1, 0))
    del t
    # 4. per (c,e): t3[(n1,n2),(n3,n4)] = sum_{p,q} t2[c,p,e,q,(n1,n2)] aq[..]   (K = the two ket legs), written straight into
    #    the layout step 5 wants: [k-leg chi][d][D][n-leg chi][nD][nD]
    #    right: a1q[r,d,u,y]: k=(r,u) n=(d,y), t3[c][d][D][e][Y][y]     left: a2q[d,l,u,x]: k=(u,l) n=(d,x), t3[e][d][D][c][X][x]
    t3 = torch.empty(xc if not left else xe, D, D, xe if not left else xc, nD, nD, dtype=dt, device=dev)
    s_e4, s_c4 = D * N3, D * xe * D * N3                              # strides of e and c in t2[c,p,e,q,N3]
    a_b = (xe, s_c4, s_e4)                                            # batch = (c,e)
    a_k4 = (D, xe * D * N3, N3)                                       # k = (p,q)
    nn = nD * nD
    if not left:
        kl, nl = xc, xe
        s_d, s_D, s_n, s_Y = D * nl * nn, nl * nn, nn, nD             # t3[c][d][D][e][Y][y]
        c_b = (xe, D * D * nl * nn, s_n)                              # c -> k leg, e -> n leg
        c_m = (nD, s_D, s_Y)                                          # m = (D,Y)
        c_n = (nD, s_d, 1)                                            # n = (d,y)
    else:
        kl, nl = xe, xc
        s_d, s_D, s_n, s_X = D * nl * nn, nl * nn, nn, nD             # t3[e][d][D][c][X][x]
        c_b = (xe, s_n, D * D * nl * nn)                              # c -> n leg, e -> k leg
        c_m = (D, s_X, s_D)                                           # m = (X,D)
        c_n = (nD, s_d, 1)                                            # n = (d,x): x innermost => coalesced stores of the 8 GiB t3
    bq4 = bq if not left else aq.permute(2, 1, 0, 3).contiguous()      # left: a2q[d,l,u,x] -> [(u,l)][(d,x)]
    ops.gemm_ex(N3, N3, D2, xc * xe, t2, bq4, t3, _ix(1, a_k4, a_b, N3, 1, 0, c_m, c_n, c_b))
    del t2
    # 5. out[f,(n-leg chi, nD, nD)] = sum_{k-leg chi, d, D} tmp2[k-leg chi, f, d, D] t3[(k-leg chi, d, D), (...)]
    tmp2 = ops.matmul(cB.contiguous(), eC.reshape(cB.shape[1], -1)) if not left else None
    if not left:
        # tmp_r2[a,f,d,D] = c13[a,b] e13[b,f,d,D]
        xf = eC.shape[1]
        A5 = tmp2.reshape(cB.shape[0], xf, D2).permute(1, 0, 2).contiguous()          # [f][a][(d,D)]
    else:
        # tmp_l2[b,f,d,D] = c24[a,b] e23[f,a,d,D]
        xf = eC.shape[0]
        A5 = contract("ab,fadD->fbdD", cB, eC).contiguous()                             # [f][b][(d,D)]
    out = ops.matmul(A5.reshape(xf, kl * D2), t3.reshape(kl * D2, nl * nn))
    return out.reshape(xf, nl, nD, nD)


def build_norm_tensor(ipeps, bond, a1q, a2q):
    """full_update.py:163-227 : N12[y,x,Y,X]; bond = (s1, s2, k); a1q/a2q (D,D,D,nD)."""
    s1, s2, k = bond
    a, b = ipeps[s1], ipeps[s2]
    c12, e12, e11 = a['C'][(k + 1) % 4], a['E'][(k + 1) % 4], a['E'][k % 4]
    c13, e13 = a['C'][(k + 2) % 4], a['E'][(k + 2) % 4]
    c21, e21, e24 = b['C'][k % 4], b['E'][k % 4], b['E'][(k + 3) % 4]
    c24, e23 = b['C'][(k + 3) % 4], b['E'][(k + 2) % 4]
    a1q, a2q = a1q.contiguous(), a2q.contiguous()
    n1 = _norm_half(c12, e12.contiguous(), e11, c13, e13.contiguous(), a1q, left=False)      # [f,e,Y,y]
    n2 = _norm_half(c21, e21.contiguous(), e24, c24, e23.contiguous(), a2q, left=True)       # [f,c,X,x]
    return contract("fcYy,fcXx->yxYX", n1, n2).contiguous()


class ALSSolver:
    """als_solver.py:6-82 with the iteration loop in libacetn_b200.so (one cooperative kernel, convergence on device)."""

    def __init__(self, n12, a12g, ar_shape, config):
        self.niter = config.als_niter
        self.tol = config.als_tol
        self.method = config.als_method
        self.epsilon = config.als_epsilon
        self.n12, self.a12g, self.ar_shape = n12, a12g, ar_shape
        self.info = None
        if self.method != "cholesky":
            raise NotImplementedError("backend='b200': als_method must be 'cholesky'")

    def initialize_tensors(self):
        """als_solver.py:112-146 (SVD of the (nD pD) x (nD pD) gate-tensor product on K5)."""
        nD, bD, pD = self.ar_shape
        m = self.a12g.permute(0, 2, 1, 3).reshape(nD * pD, nD * pD)
        U, S, Vh = svd_small(m)
        V = Vh.t()
        S = torch.sqrt(S[:bD] / S[0])
        a1r = (U[:, :bD].reshape(nD, pD, bD) * S).permute(0, 2, 1).contiguous()
        a2r = (V[:, :bD].reshape(nD, pD, bD) * S).permute(0, 2, 1).contiguous()
        n12g = contract("yxYX,yxpq->YXpq", self.n12, self.a12g).contiguous()
        return a1r, a2r, n12g

    def solve(self):
        a1r, a2r, n12g = self.initialize_tensors()
        a1r, a2r, self.info = ops.als_solve(a1r, a2r, n12g, self.n12, self.a12g, niter=self.niter, tol=self.tol, epsilon=self.epsilon)
        return a1r, a2r


# ---- small dense linear algebra on the library's kernels ------------------------------------------------------------
def svd_small(M):
    """torch.linalg.svd(M) of a small square matrix: M = U diag(S) Vh, S descending (K5: M = Jt^T diag(S) Wt)."""
    S, Wt, Jt, _ = ops.jacobi_svd(M.contiguous())
    return Jt.t().contiguous(), S, Wt


def eigh_sym(N):
    """Eigen-decomposition of a symmetric matrix, N = V diag(w) V^T (torch.linalg.eigh up to the order of the pairs): the left
    singular vectors of K5 are orthonormal eigenvectors, the eigenvalues are their Rayleigh quotients (signed)."""
    N = N.contiguous()
    _, _, Jt, _ = ops.jacobi_svd(N)
    w = (ops.matmul(Jt, N) * Jt).sum(dim=1)
    return w, Jt.t().contiguous()


def pinv_small(R, atol=1e-12):
    """torch.linalg.pinv(R, atol=atol) (singular values <= atol are dropped; rtol = 0 as in torch when atol is given)."""
    S, Wt, Jt, _ = ops.jacobi_svd(R.contiguous())
    inv = torch.where(S > atol, 1.0 / S, torch.zeros_like(S))
    return ops.matmul((Wt * inv[:, None]).contiguous(), Jt, transpose_a=True)


def qr_small(M):
    """torch.linalg.qr(M) of a tall matrix up to the signs of the columns of Q / rows of R: Q = K4 (Householder TSQR), R = Q^T M."""
    M = M.contiguous()
    Q = ops.orthonormalize(M.clone())
    return Q, ops.matmul(Q, M, transpose_a=True)


# ---- the callers around the norm tensor / ALS --------------------------------------------------------
def decompose_site_tensors(a1, a2):
    """tensor_update.py:53-68."""
    bD, pD = a1.shape[3:]
    nD = min(bD ** 3, pD * bD)
    a1q, a1r = qr_small(a1.permute(2, 3, 1, 0, 4).reshape(bD ** 3, pD * bD))        # "lurdp->rdulp"
    a2q, a2r = qr_small(a2.permute(3, 0, 1, 2, 4).reshape(bD ** 3, pD * bD))        # "lurdp->dlurp"
    return a1q.reshape(bD, bD, bD, nD), a1r.reshape(nD, bD, pD), a2q.reshape(bD, bD, bD, nD), a2r.reshape(nD, bD, pD)


def recompose_site_tensors(a1q, a1r, a2q, a2r):
    """tensor_update.py:70-75."""
    return contract("rdux,xlp->lurdp", a1q, a1r).contiguous(), contract("dlux,xrp->lurdp", a2q, a2r).contiguous()


def positive_approx(n12, cutoff=1e-12):
    """full_update.py:262-293."""
    nD = n12.shape[0]
    N = n12.reshape(nD ** 2, nD ** 2).clone()
    # torch.linalg.eigh (UPLO='L') reads the lower triangle only; on a not-yet-converged environment the norm tensor is visibly
    # non-symmetric, so the matrix the reference actually decomposes is tril(N) mirrored
    N = torch.tril(N) + torch.tril(N, -1).t()
    nw, nz = eigh_sym(N)
    lo = float(nw.min())
    while lo < cutoff:
        N += 2 * max(cutoff, abs(lo)) * torch.eye(nD ** 2, dtype=N.dtype, device=N.device)
        nw, nz = eigh_sym(N)
        lo = float(nw.min())
    return nz.reshape(nD, nD, nD ** 2) * torch.sqrt(nw)


def gauge_fix(nz, a12g, atol=1e-12):
    """full_update.py:296-343."""
    nD = a12g.shape[0]
    _, nzyr = qr_small(nz.permute(2, 1, 0).reshape(nD ** 3, nD))        # "yxz->zxy"
    _, nzxr = qr_small(nz.permute(2, 0, 1).reshape(nD ** 3, nD))        # "yxz->zyx"
    nzyr_inv = pinv_small(nzyr, atol=atol)
    nzxr_inv = pinv_small(nzxr, atol=atol)
    nz = contract("yxz,xw->yzw", nz, nzxr_inv)
    nz = contract("yzw,yv->zvw", nz, nzyr_inv).contiguous()
    n12 = contract("zvw,zVW->vwVW", nz, nz).contiguous()
    a12g = contract("zx,yxpq->yzpq", nzxr, a12g)
    a12g = contract("wy,yzpq->wzpq", nzyr, a12g).contiguous()
    return n12, a12g, nzxr_inv, nzyr_inv


def finalize_reduced_tensors(a1r, a2r, nzxr_inv=None, nzyr_inv=None):
    """full_update.py:121-161."""
    if nzyr_inv is not None:
        a1r = contract("yz,zup->yup", nzyr_inv, a1r)
        a2r = contract("xw,wvq->xvq", nzxr_inv, a2r)
    nD, bD, pD = a1r.shape
    q1, r1 = qr_small(a1r.permute(0, 2, 1).reshape(nD * pD, bD))
    q2, r2 = qr_small(a2r.permute(0, 2, 1).reshape(nD * pD, bD))
    U, s, Vh = svd_small(ops.matmul(r1, r2.t().contiguous()))
    s = torch.sqrt(s[:bD] / s.norm())
    r1 = (U[:, :bD] * s).contiguous()                       # "ab,b->ab"
    r2 = (Vh[:bD, :] * s[:, None]).contiguous()             # "ba,b->ba"
    a1r = contract("ypa,au->yup", q1.reshape(nD, pD, bD), r1).contiguous()
    a2r = contract("xqb,vb->xvq", q2.reshape(nD, pD, bD), r2).contiguous()
    return a1r, a2r


def full_update_bond(ipeps, bond, a1, a2, gate, config):
    """FullUpdater.tensor_update (full_update.py:34-96): a1, a2 in the bond frame, gate (d,d,d,d) -> updated, normalised a1, a2."""
    a1q, a1r, a2q, a2r = decompose_site_tensors(a1, a2)
    n12 = build_norm_tensor(ipeps, bond, a1q, a2q)
    a12g = contract("yup,xuq->yxpq", a1r, a2r)
    a12g = contract("yxpq,pqrs->yxrs", a12g, gate.contiguous()).contiguous()
    nz = positive_approx(n12, cutoff=config.positive_approx_cutoff)
    inv = (None, None)
    if config.use_gauge_fix:
        n12, a12g, nzxr_inv, nzyr_inv = gauge_fix(nz, a12g, atol=config.gauge_fix_atol)
        inv = (nzxr_inv, nzyr_inv)
    else:
        n12 = contract("xyz,XYz->xyXY", nz.contiguous(), nz.contiguous()).contiguous()
    b1, b2 = ALSSolver(n12, a12g, tuple(a1r.shape), config).solve()
    b1, b2 = finalize_reduced_tensors(b1, b2, *inv)
    a1n, a2n = recompose_site_tensors(a1q, b1, a2q, b2)
    ops.frob_normalize(a1n)
    ops.frob_normalize(a2n)
    return a1n, a2n
