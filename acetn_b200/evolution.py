"""B200 counterparts of the full-update pieces: build_norm_tensor (acetn/evolution/full_update.py:163-227), the ALS inner
solver (acetn/evolution/als_solver.py) and -- SURVEY.md 8f-1 -- the callers around them in one bond update: QR split
(tensor_update.py:53-75), positive_approx / gauge_fix (full_update.py:262-343), finalize_reduced_tensors (:121-161).

Everything dense runs on the library's own kernels (no cuSOLVER): the norm tensor behind acetn_b200_norm_tensor, the small
contractions = gather + batched K1 DGEMM (`ops.contract`),
thin QR = K4 Householder TSQR + one K1 product for R, symmetric eigen-decomposition / SVD / pseudo-inverse of the small
matrices (<= 256 x 256) = K5 one-sided Jacobi, ALS loop = K6.  `ALSSolver` keeps the reference's constructor/solve() shape so
it can replace `ALSSolver(n12, a12g, ar_shape, config).solve()` when `config.backend == "b200"` (als_solver.py:48-51);
`full_update_bond` is `FullUpdater.tensor_update` (full_update.py:34-63)."""
import torch

from . import ops
from .ops import contract


def build_norm_tensor(ipeps, bond, a1q, a2q):
    """full_update.py:163-227 : N12[y,x,Y,X]; bond = (s1, s2, k); a1q/a2q (D,D,D,nD).  One C-ABI call (acetn_b200_norm_tensor,
    acetn_b200/csrc/environment.cu): every leg permutation of the reference's einsum chain is folded into K1's two-level index
    descriptors, so the 2 / 4 / 8 GiB intermediates are written once in the layout the next GEMM reads."""
    s1, s2, k = bond
    return ops.norm_tensor(ipeps[s1], ipeps[s2], k, a1q, a2q)


class ALSSolver:
    """als_solver.py:6-82.  Both methods run the whole iteration loop in libacetn_b200.so (K6, one cooperative kernel, convergence
    test on the device): "cholesky" (default) and "pinv" (als_solver.py:226-228 = csrc/evolution/als_solve.cpp:47-50; symmetric
    eigen-decomposition by two-sided Jacobi inside the kernel).  Normal matrices too large for the kernel's shared memory
    (nD * bD > PINV_KERNEL_MAX, i.e. D >= 9 at d = 2) take the reference's host-driven pinv loop on the library's kernels (K1
    contractions, K4 + K5 eigen-decomposition, one host read of the cost per iteration like the reference, als_solver.py:78-79)."""
    PINV_KERNEL_MAX = 159

    def __init__(self, n12, a12g, ar_shape, config):
        self.niter = config.als_niter
        self.tol = config.als_tol
        self.method = config.als_method
        self.epsilon = config.als_epsilon
        self.n12, self.a12g, self.ar_shape = n12, a12g, ar_shape
        self.info = None
        if self.method not in ("cholesky", "pinv"):
            raise ValueError(f"Invalid als_method: {self.method} provided.")

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
        nD, bD, _ = self.ar_shape
        if self.method == "pinv" and nD * bD > self.PINV_KERNEL_MAX:
            return self.solve_pinv(a1r, a2r, n12g)
        a1r, a2r, self.info = ops.als_solve(a1r, a2r, n12g, self.n12, self.a12g, niter=self.niter, tol=self.tol, epsilon=self.epsilon,
                                            method=self.method)
        return a1r, a2r

    # ---- method "pinv" ------------------------------------------------------------------------------------------------
    def solve_ar_pinv(self, R, S):
        """als_solver.py:213-229, pinv branch: ar = pinv(R_sym, hermitian=True, rcond=epsilon) @ S."""
        nD, bD, pD = S.shape
        n = nD * bD
        R = R.reshape(n, n)
        R = (0.5 * (R + R.t())).contiguous()
        w, V = eigh_sym(R)
        keep = w.abs() > self.epsilon * w.abs().max()
        winv = torch.where(keep, 1.0 / torch.where(keep, w, torch.ones_like(w)), torch.zeros_like(w))
        y = ops.matmul(V, S.reshape(n, pD).contiguous(), transpose_a=True)          # V^T S
        return ops.matmul(V, (y * winv[:, None]).contiguous()).reshape(nD, bD, pD)

    def calculate_cost(self, a1r, a2r):
        """als_solver.py:246-257."""
        a12n = contract("yup,xuq->yxpq", a1r, a2r).contiguous()
        d2 = contract("yxYX,yxpq->YXpq", self.n12, a12n)
        d3 = contract("yxYX,yxpq->YXpq", self.n12, self.a12g)
        return ((d2 - 2.0 * d3) * a12n).sum()

    def solve_pinv(self, a1r, a2r, n12g):
        """als_solver.py:55-82 with solve_ar's pinv branch."""
        n12 = self.n12
        d1 = abs(float(self.calculate_cost(a1r, a2r)))
        it = 0
        for i in range(self.niter):
            it = i + 1
            S = contract("YXpQ,XUQ->YUp", n12g, a2r).contiguous()
            R = contract("yxYX,xuq->yYXuq", n12, a2r)
            R = contract("yYXuQ,XUQ->YUyu", R.contiguous(), a2r).contiguous()
            a1r = self.solve_ar_pinv(R, S)
            S = contract("YXPq,YVP->XVq", n12g, a1r).contiguous()
            R = contract("yxYX,yvp->xYXvp", n12, a1r)
            R = contract("xYXvP,YVP->XVxv", R.contiguous(), a1r).contiguous()
            a2r = self.solve_ar_pinv(R, S)
            d2 = float(self.calculate_cost(a1r, a2r))       # the reference's one host read per iteration (als_solver.py:78-79)
            if abs(d2 - d1) / abs(d1) < self.tol and i > 1:
                break
            d1 = d2
        self.info = torch.tensor([it, 0], dtype=torch.int32)
        return a1r, a2r


# ---- small dense linear algebra on the library's kernels ------------------------------------------------------------
def svd_small(M):
    """torch.linalg.svd(M) of a small square matrix: M = U diag(S) Vh, S descending (K5: M = Jt^T diag(S) Wt)."""
    S, Wt, Jt, _ = ops.jacobi_svd(M.contiguous())
    return Jt.t().contiguous(), S, Wt


def eigh_sym(N):
    """Eigen-decomposition of a symmetric matrix, N = V diag(w) V^T (torch.linalg.eigh up to the order of the pairs).

    The singular vectors of N are eigenvectors only where |lambda| is simple: a pair +lambda / -lambda (noise eigenvalues of either sign
    on a norm matrix that is positive only up to truncation error) is ONE singular value, and an SVD may return any rotation of the two
    eigenvectors -- Rayleigh quotients near 0 instead of +-lambda.  So the SVD is taken of N + sigma I with sigma = 1.01 ||N||_F >= the
    spectral radius: every eigenvalue becomes positive, singular vectors = eigenvectors, and what is lost is relative accuracy of tiny
    eigenvalues below eps * ||N|| -- which a symmetric eigensolver (LAPACK syevd behind torch.linalg.eigh) does not have either.
    The eigenvalues are the Rayleigh quotients with the unshifted N (signed)."""
    N = N.contiguous()
    n = N.shape[0]
    sigma = 1.01 * torch.linalg.matrix_norm(N)                   # device scalar, no host read
    Ns = N + sigma * torch.eye(n, dtype=N.dtype, device=N.device)
    # QR first: one-sided Jacobi on the triangular factor converges in a few sweeps
    Q, R = qr_small(Ns)
    _, _, Jt, _ = ops.jacobi_svd(R)
    V = ops.matmul(Q, Jt.t().contiguous())                   # left singular vectors of N + sigma I = Q R = (Q Jt^T) S Wt
    w = (ops.matmul(N, V) * V).sum(dim=0)
    # ascending like torch.linalg.eigh: the callers' results are gauge invariant in exact arithmetic only -- gauge_fix inverts factors with
    # condition numbers ~1e3, so the ORDER of the eigenpairs is visible at ~1e-9 in the updated tensors; keep the reference's order
    w, order = torch.sort(w, stable=True)
    return w, V[:, order].contiguous()


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
