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


import os as _os

_SMALL_K_TILE = int(_os.environ.get("ACETN_B200_ENV_TILE", "0"))      # K1 tile of the K = D^2 steps (0 = planner's choice)


def _ix(*triples):
    """27 ints of ops.gemm_ex: {div, s_hi, s_lo} for A(m, k, batch), B(k, n, batch), C(m, n, batch); a plain int is a stride."""
    out = []
    for t in triples:
        out += [0, 0, t] if isinstance(t, int) else list(t)
    return out


def env_front(cA, eA, eB):
    """t[(c,p,P),(e,q,Q)] = sum_{a,b} cA[a,b] eA[b,c,p,P] eB[e,a,q,Q]: the boundary part of a bond-environment half (full_update.py
    :205-206 / :220-221, rdm.py:95-96 / :100-101), shared by everything that is contracted with site tensors afterwards."""
    x0, x1 = cA.shape
    xc, D = eA.shape[1], eA.shape[2]
    xe = eB.shape[0]
    D2 = D * D
    t1 = ops.matmul(cA.contiguous(), eA.contiguous().reshape(x1, xc * D2))            # t1[a,(c,p,P)]
    m, n = xc * D2, xe * D2
    t = torch.empty(m, n, dtype=cA.dtype, device=cA.device)
    ops.gemm_ex(m, n, x0, 1, t1, eB.contiguous(), t, _ix(1, m, 0, D2, (D2, x0 * D2, 1), 0, n, 1, 0))
    return t, (xc, xe, D)


def env_back(t, dims, A5, bra, ket, Yb, yk, left):
    """out[f, n-leg chi, Yb, yk] from t = env_front(...), the closing boundary A5[f][k-leg chi][(d,D)] and the site factors as
    plain matrices:  bra[(P,Q)][Nb], ket[(p,q)][(d,yk)]  with  Nb enumerated (D,Yb) (right half) or (Yb,D) (left half).
        right: k-leg = c (first chi leg of t), n-leg = e        left: k-leg = e, n-leg = c
    Every leg permutation of the reference's einsum chain is folded into the two-level index descriptors of K1, so none of the
    2 - 8 GiB intermediates is re-laid out in HBM (the generic gather + GEMM form spent 42 % of its time in gathers)."""
    xc, xe, D = dims
    dev, dt = t.device, t.dtype
    D2 = D * D
    n = xe * D2
    Nb, Nk = D * Yb, D * yk
    # 3. t2[c,p,e,q,Nb] = sum_{P,Q} t[c,p,P,e,q,Q] bra[(P,Q),Nb]
    M3 = xc * D * xe * D
    t2 = torch.empty(M3, Nb, dtype=dt, device=dev)
    a_m = (xe * D, D * n, D)               # hi = (c,p): stride D*n, lo = (e,q): stride D
    a_k = (D, n, 1)                        # hi = P: stride n, lo = Q: stride 1
    ops.gemm_ex(M3, Nb, D2, 1, t, bra, t2, _ix(a_m, a_k, 0, Nb, 1, 0, Nb, 1, 0), force_tile=_SMALL_K_TILE)
    # 4. per (c,e): t3[Nb,(d,yk)] = sum_{p,q} t2[c,p,e,q,Nb] ket[(p,q),(d,yk)], written straight into the layout step 5 wants:
    #    [k-leg chi][d][D][n-leg chi][Yb][yk]
    kl, nl = (xc, xe) if not left else (xe, xc)
    nn = Yb * yk
    t3 = torch.empty(kl, D, D, nl, Yb, yk, dtype=dt, device=dev)
    s_e4, s_c4 = D * Nb, D * xe * D * Nb                              # strides of e and c in t2[c,p,e,q,Nb]
    a_b = (xe, s_c4, s_e4)                                            # batch = (c,e)
    a_k4 = (D, xe * D * Nb, Nb)                                       # k = (p,q)
    s_d, s_D, s_n, s_Y = D * nl * nn, nl * nn, nn, yk
    if not left:
        c_b = (xe, D * D * nl * nn, s_n)                              # c -> k leg, e -> n leg
        c_m = (Yb, s_D, s_Y)                                          # m = (D,Yb)
    else:
        c_b = (xe, s_n, D * D * nl * nn)                              # c -> n leg, e -> k leg
        c_m = (D, s_Y, s_D)                                           # m = (Yb,D)
    c_n = (yk, s_d, 1)                                                # n = (d,yk): yk innermost => coalesced stores
    ops.gemm_ex(Nb, Nk, D2, xc * xe, t2, ket, t3, _ix(1, a_k4, a_b, Nk, 1, 0, c_m, c_n, c_b), force_tile=_SMALL_K_TILE)
    del t2
    # 5. out[f,(n-leg chi, Yb, yk)] = sum_{k-leg chi, d, D} A5[f,(k-leg chi, d, D)] t3[(k-leg chi, d, D), (...)]
    xf = A5.shape[0]
    out = ops.matmul(A5.reshape(xf, kl * D2), t3.reshape(kl * D2, nl * nn))
    return out.reshape(xf, nl, Yb, yk)


def closing_right(cB, eC):
    """tmp_r2[a,f,d,D] = cB[a,b] eC[b,f,d,D] (full_update.py:207, rdm.py:97) as A5[f][a][(d,D)]."""
    D2 = eC.shape[2] * eC.shape[3]
    xf = eC.shape[1]
    return ops.matmul(cB.contiguous(), eC.contiguous().reshape(cB.shape[1], xf * D2)).reshape(cB.shape[0], xf, D2).permute(1, 0, 2).contiguous()


def closing_left(cB, eC):
    """tmp_l2[b,f,d,D] = cB[a,b] eC[f,a,d,D] (full_update.py:222, rdm.py:102) as A5[f][b][(d,D)]."""
    return contract("ab,fadD->fbdD", cB, eC).contiguous()


def build_norm_tensor(ipeps, bond, a1q, a2q):
    """full_update.py:163-227 : N12[y,x,Y,X]; bond = (s1, s2, k); a1q/a2q (D,D,D,nD).
        right: n1[f,e,Y,y] = sum tmp_r2[a,f,d,D] ( ((c12 e12) e11) conj(a1q)[R,D,U,Y] a1q[r,d,u,y] )[a,e,D,Y,d,y]
        left : n2[f,c,X,x] = sum tmp_l2[b,f,d,D] ( ((c21 e21) e24) conj(a2q)[D,L,U,X] a2q[d,l,u,x] )[c,b,X,D,x,d]"""
    s1, s2, k = bond
    a, b = ipeps[s1], ipeps[s2]
    c12, e12, e11 = a['C'][(k + 1) % 4], a['E'][(k + 1) % 4], a['E'][k % 4]
    c13, e13 = a['C'][(k + 2) % 4], a['E'][(k + 2) % 4]
    c21, e21, e24 = b['C'][k % 4], b['E'][k % 4], b['E'][(k + 3) % 4]
    c24, e23 = b['C'][(k + 3) % 4], b['E'][(k + 2) % 4]
    D, nD = a1q.shape[0], a1q.shape[3]
    # the 64 KiB site factors as plain (k, n) matrices (conj is the identity: FP64 real)
    q1 = a1q.permute(0, 2, 1, 3).contiguous().reshape(D * D, D * nD)                    # [(R,U)][(D,Y)] = [(r,u)][(d,y)]
    t, dims = env_front(c12, e12, e11)
    n1 = env_back(t, dims, closing_right(c13, e13), q1, q1, nD, nD, left=False)        # [f,e,Y,y]
    del t
    bra2 = a2q.permute(2, 1, 3, 0).contiguous().reshape(D * D, nD * D)                  # [(U,L)][(X,D)]
    ket2 = a2q.permute(2, 1, 0, 3).contiguous().reshape(D * D, D * nD)                  # [(u,l)][(d,x)]
    t, dims = env_front(c21, e21, e24)
    n2 = env_back(t, dims, closing_left(c24, e23), bra2, ket2, nD, nD, left=True)      # [f,c,X,x]
    del t
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
    singular vectors (K4 + K5) are orthonormal eigenvectors, the eigenvalues are their Rayleigh quotients (signed)."""
    N = N.contiguous()
    # QR first: one-sided Jacobi on the triangular factor converges in ~6 sweeps instead of 13-18 on the symmetric matrix itself
    Q, R = qr_small(N)
    _, _, Jt, _ = ops.jacobi_svd(R)
    V = ops.matmul(Q, Jt.t().contiguous())                   # left singular vectors of N = Q R = (Q Jt^T) S Wt
    w = (ops.matmul(N, V) * V).sum(dim=0)
    return w, V


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
        # the reference adds shift * I and decomposes again (full_update.py:288-290): N + shift I has the eigenvectors of N and
        # its eigenvalues moved by shift, so the second decomposition is not needed (nz nz^T is the same to rounding)
        shift = 2 * max(cutoff, abs(lo))
        nw = nw + shift
        lo += shift
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
