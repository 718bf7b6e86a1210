"""B200 counterparts of the full-update pieces on the hot path: build_norm_tensor
(acetn/evolution/full_update.py:163-227) and the ALS inner solver (acetn/evolution/als_solver.py).

The callers around them (QR split, positive_approx, gauge_fix, finalize, gates -- SURVEY.md 8f-1) stay in the
reference; `ALSSolver` keeps the reference's constructor/solve() shape so it can replace
`ALSSolver(n12, a12g, ar_shape, config).solve()` when `config.backend == "b200"` (als_solver.py:48-51)."""
import torch

from . import ops
from .ops import contract


def build_norm_tensor(ipeps, bond, a1q, a2q):
    """full_update.py:163-227 : N12[y,x,Y,X]; bond = (s1, s2, k); a1q/a2q (D,D,D,nD)."""
    s1, s2, k = bond
    a, b = ipeps[s1], ipeps[s2]
    c12, e12, e11 = a['C'][(k + 1) % 4], a['E'][(k + 1) % 4], a['E'][k % 4]
    c13, e13 = a['C'][(k + 2) % 4], a['E'][(k + 2) % 4]
    c21, e21, e24 = b['C'][k % 4], b['E'][k % 4], b['E'][(k + 3) % 4]
    c24, e23 = b['C'][(k + 3) % 4], b['E'][(k + 2) % 4]
    # right half
    t = contract("ab,bcrR->acrR", c12, e12)
    t = contract("acrR,eauU->crReuU", t, e11)
    t = contract("crReuU,RDUY->creuDY", t, a1q.conj())
    t = contract("creuDY,rduy->ceDYdy", t, a1q)
    n1 = contract("ab,bfdD->afdD", c13, e13)
    n1 = contract("afdD,aeDYdy->feYy", n1, t)
    # left half
    t = contract("ab,bcuU->acuU", c21, e21)
    t = contract("acuU,ealL->cuUelL", t, e24)
    t = contract("cuUelL,DLUX->cuelXD", t, a2q.conj())
    t = contract("cuelXD,dlux->ceXDxd", t, a2q)
    n2 = contract("ab,fadD->bfdD", c24, e23)
    n2 = contract("bfdD,cbXDxd->fcXx", n2, t)
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
        """als_solver.py:112-146 (a 2nD x 2nD SVD: host-side small dense LA)."""
        nD, bD, pD = self.ar_shape
        m = self.a12g.permute(0, 2, 1, 3).reshape(nD * pD, nD * pD)
        U, S, Vh = torch.linalg.svd(m)
        V = Vh.mH
        S = torch.sqrt(S[:bD] / S[0])
        a1r = (U[:, :bD].reshape(nD, pD, bD) * S).permute(0, 2, 1).contiguous()
        a2r = (V[:, :bD].reshape(nD, pD, bD) * S).permute(0, 2, 1).contiguous()
        n12g = contract("yxYX,yxpq->YXpq", self.n12, self.a12g).contiguous()
        return a1r, a2r, n12g

    def solve(self):
        a1r, a2r, n12g = self.initialize_tensors()
        a1r, a2r, self.info = ops.als_solve(a1r, a2r, n12g, self.n12, self.a12g, niter=self.niter, tol=self.tol, epsilon=self.epsilon)
        return a1r, a2r
