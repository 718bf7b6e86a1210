"""Row-sharded half-system projector: one projector task computed cooperatively by a group of G ranks
(SURVEY.md 8e "2x2 on 8 GPUs", hard part 7 -- no reference counterpart).

Rank g of the group builds only its row block of the two quarter tensors -- the rows (c,r,R) of Q_k with the chi leg c
in block g, which is exactly `make_quarter_tensor` applied to the slice E[k][:, c-block] (projectors.py:52-59) -- so the
O(chi^3 D^4 + chi^2 D^6 d) construction and every large product of the randomized SVD are split G ways:

    forward products  Y = Q1 X      : each rank computes its row block of Y           -> summed into the full thin matrix
    adjoint products  Z = Q1^T Y    : each rank computes Q1[g]^T Y[rows g] (partial)  -> summed (all-reduce)

Both exchanges are all-reduces of one (chi D^2) x (chi+2) thin matrix (34 MB at D=8, chi=256: disjoint row supports sum
exactly, x + 0 = x, so the result is bit-identical on every rank); 13 of them per projector, < 3 % of the halved GEMM
time over NVLink.  The latency-bound stages (TSQR orthonormalisation, Jacobi core) are replicated: every rank of the
group runs the identical deterministic kernels on identical data, so no further exchange is needed.  Omega is the
caller's (drawn in the canonical order by every rank).

The linear-algebra backend is injected (`la`) so that the exchange logic is testable on CPU/gloo with torch kernels.
"""
import torch
import torch.distributed as dist


class _EncodedRows:
    """Row block of a quarter tensor together with its K7 encoding (int8 residue planes); quacks like the matrix for `.shape`."""

    def __init__(self, Q, enc):
        self.Q, self.enc, self.shape = Q, enc, Q.shape


class B200LinAlg:
    """libacetn_b200.so kernels.  use_i8 (default: env ACETN_B200_COOP_I8 == "1", EXPERIMENTAL -- written for the next round, not
    yet measured on multi-GPU boxes): the row blocks are encoded once and every big x thin product of the cooperative rSVD
    runs on K7 (exact integer arithmetic on the INT8 tensor cores) like in the single-rank path."""

    def __init__(self, use_i8=None):
        import os
        from . import ops
        self.ops = ops
        self.use_i8 = (os.environ.get("ACETN_B200_COOP_I8", "0") == "1") if use_i8 is None else bool(use_i8)

    def quarter_rows(self, site_tensor, k, c0, c1, absmax):
        ak = site_tensor.bond_permute(k)
        ck = site_tensor['C'][k % 4]
        ek1 = site_tensor['E'][(3 + k) % 4]
        ek2 = site_tensor['E'][k % 4][:, c0:c1].contiguous()
        Q, shp = self.ops.quarter_tensor(ck, ek2, ek1, ak, normalize=False, absmax=absmax)
        if self.use_i8 and min(Q.shape) >= 4096 and self.ops.i8_supported(Q.shape[0], Q.shape[1], 272):
            return _EncodedRows(Q, self.ops.i8_encode(Q))
        return Q

    def matmul(self, A, B, transpose_a=False, out=None):
        if isinstance(A, _EncodedRows):
            if B.shape[1] <= 272:
                return self.ops.i8_matmul(A.enc, B.contiguous(), adjoint=transpose_a, out=out)
            A = A.Q
        return self.ops.matmul(A, B, transpose_a=transpose_a, out=out)

    def orthonormalize(self, Y):
        return self.ops.orthonormalize(Y)

    def core_svd(self, R, chi, cutoff):
        S, Wt, Jt, info = self.ops.jacobi_svd(R, chi=chi, cutoff=cutoff)
        return S, Wt, Jt, info

    def keep_of(self, info):
        return int(info[0].item())


def _blocks(n, G):
    """G contiguous blocks of [0, n) (sizes differ by at most one)."""
    base, rem = divmod(n, G)
    out, start = [], 0
    for g in range(G):
        size = base + (1 if g < rem else 0)
        out.append((start, start + size))
        start += size
    return out


class ShardedHalfSystemProjector:
    def __init__(self, config, group, group_rank, group_size, la=None):
        self.cfg = config
        self.group, self.g, self.G = group, group_rank, group_size
        self.la = la if la is not None else B200LinAlg()
        self.spectra = None

    # ---- exchange -------------------------------------------------------------------------------------------------
    def _sum(self, t):
        if self.G > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _max(self, t):
        if self.G > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def _forward(self, Qg, rows, X, nrows_total):
        """full (nrows_total x q) = Q X from this rank's row block Qg = Q[rows]."""
        if self.G > 1 and nrows_total % self.G == 0 and rows[1] - rows[0] == nrows_total // self.G:
            # equal row blocks: plain all-gather of the blocks (half the traffic of the zero-padded all-reduce)
            out = torch.empty(nrows_total, X.shape[1], dtype=X.dtype, device=X.device)
            blk = self.la.matmul(Qg, X, out=out[rows[0]:rows[1]])
            dist.all_gather_into_tensor(out, blk, group=self.group)
            return out
        out = torch.zeros(nrows_total, X.shape[1], dtype=X.dtype, device=X.device)
        if rows[1] > rows[0]:
            self.la.matmul(Qg, X, out=out[rows[0]:rows[1]])
        return self._sum(out)

    def _adjoint(self, Qg, rows, Y):
        """full Q^T Y from the partial Qg^T Y[rows]."""
        if rows[1] > rows[0]:
            part = self.la.matmul(Qg, Y[rows[0]:rows[1]].contiguous(), transpose_a=True)
        else:
            part = torch.zeros(Qg.shape[1], Y.shape[1], dtype=Y.dtype, device=Y.device)
        return self._sum(part)

    # ---- the projector pair ---------------------------------------------------------------------------------------------
    def begin(self, ipeps, sites, k, omega):
        cfg = self.cfg
        chi = ipeps.dims["chi"]
        D = ipeps.dims["bond"]
        D2 = D * D
        st1, st4 = ipeps[sites[0]], ipeps[sites[3]]
        xc1 = st1['E'][k % 4].shape[1]                 # rows of Q1 = xc1 * D^2
        xc4 = st4['E'][(k + 3) % 4].shape[1]           # rows of Q4
        b1, b4 = _blocks(xc1, self.G)[self.g], _blocks(xc4, self.G)[self.g]
        rows1, rows4 = (b1[0] * D2, b1[1] * D2), (b4[0] * D2, b4[1] * D2)
        m, kdim = xc1 * D2, xc4 * D2
        mx = torch.zeros(2, dtype=omega.dtype, device=omega.device)
        Q1g = self.la.quarter_rows(st1, k, b1[0], b1[1], mx[0:1])
        Q4g = self.la.quarter_rows(st4, k + 3, b4[0], b4[1], mx[1:2])
        self._max(mx)
        if Q1g.shape[1] != kdim:
            raise ValueError("sharded projector: inner dimensions of Q1 and Q4 differ")
        q = omega.shape[1]
        Y = self._forward(Q1g, rows1, self._forward(Q4g, rows4, omega, kdim), m)
        for _ in range(cfg.rsvd_niter):
            self.la.orthonormalize(Y)
            Z = self._adjoint(Q4g, rows4, self._adjoint(Q1g, rows1, Y))
            Y = self._forward(Q1g, rows1, self._forward(Q4g, rows4, Z, kdim), m)
        self.la.orthonormalize(Y)
        AtQ = self._adjoint(Q1g, rows1, Y)                     # Q1^T Qy   (kdim x q)
        Z = self._adjoint(Q4g, rows4, AtQ)                      # Bt^T = Q4^T Q1^T Qy   (n x q)
        Qb = Z.clone()
        self.la.orthonormalize(Qb)
        R = self.la.matmul(Qb, Z, transpose_a=True)             # (q x q) core, replicated
        S, Wt, Jt, info = self.la.core_svd(R, chi, cfg.svd_cutoff)
        V = self.la.matmul(Qb, Jt.t().contiguous())             # (n x q)
        return {"Q4g": Q4g, "rows4": rows4, "kdim": kdim, "AtQ": AtQ, "Wt": Wt, "V": V, "S": S, "info": info, "mx": mx,
                "shape1": (st1['E'][(3 + k) % 4].shape[0], D, D), "shape4": (xc4, D, D)}

    def finish(self, pd):
        keep = self.la.keep_of(pd["info"])
        S = pd["S"]
        if self.spectra is not None:
            self.spectra.append((S / S[0]).detach().cpu())
        w = 1.0 / torch.sqrt(S[:keep] / S[0])
        Ub = (pd["Wt"][:keep].t() * (w / pd["mx"][0])).contiguous()          # U_B diag(w) / max|Q1|   (q x keep)
        Vs = (pd["V"][:, :keep] * (w / pd["mx"][1])).contiguous()            # (n x keep)
        p1 = self.la.matmul(pd["AtQ"], Ub)                                   # replicated small product
        p2 = self._forward(pd["Q4g"], pd["rows4"], Vs, pd["kdim"])
        return p1.view(*pd["shape1"], keep), p2.view(*pd["shape4"], keep)
