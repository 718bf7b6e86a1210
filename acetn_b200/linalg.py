"""B200 counterparts of acetn/linalg: svd_lowrank, fused_matmul_svd_lowrank, fused_3matmul_svd_lowrank.

Same signatures and return convention as the reference (U, S, V with V not transposed).  The Gaussian test matrix
is drawn here with torch.randn on the tensors' device exactly as the reference does
(acetn/linalg/fused_matmul_svd_lowrank.py:32) so that both paths consume the same random stream; everything else
(thin DGEMM chain, TSQR orthonormalisation, Jacobi core SVD) runs in libacetn_b200.so."""
import torch

from . import ops

_omega_source = None


def set_omega_source(fn):
    """Override the test-matrix generator: fn(n, q, dtype, device) -> (n, q) tensor.  None restores torch.randn.
    Used by the parity tests to replay the oracle's Omega tape."""
    global _omega_source
    _omega_source = fn


def _omega(n, q, dtype, device):
    if _omega_source is not None:
        return _omega_source(n, q, dtype, device).to(device=device, dtype=dtype)
    return torch.randn(n, q, dtype=dtype, device=device)


last_info = None   # device int32[2] of the most recent call: [kept (uncapped by chi unless given), jacobi sweeps]


def _run(mats, q, niter, reorth, chi=None, cutoff=1e-12):
    global last_info
    m, n = mats[0].shape[0], mats[-1].shape[1]
    q = min(q, m, n)
    omega = _omega(n, q, mats[0].dtype, mats[0].device)
    U, S, V, info = ops.rsvd(list(mats), omega, niter=niter, reorth_adjoint=reorth, chi=chi, cutoff=cutoff)
    last_info = info
    return U, S, V


def svd_lowrank(A, q=6, niter=2):
    """acetn/linalg/svd_lowrank.py:4-44."""
    return _run([A], q, niter, False)


def fused_matmul_svd_lowrank(A, B, q=6, niter=2):
    """acetn/linalg/fused_matmul_svd_lowrank.py:4-52 (rSVD of A @ B without forming it)."""
    return _run([A, B], q, niter, False)


def fused_3matmul_svd_lowrank(A, B, C, D, q=6, niter=2):
    """acetn/linalg/fused_3matmul_svd_lowrank.py:4-56 (rSVD of (A@B)@(C@D); re-orthonormalises mid power step)."""
    return _run([A, B, C, D], q, niter, True)
