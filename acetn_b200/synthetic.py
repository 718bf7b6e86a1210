"""Synthetic random-init iPEPS tensors and the algorithmic flop model of the CTMRG sweep.

Per site, after torch.manual_seed(seed): A = torch.rand(D,D,D,D,d, float64) - 0.5, normalised; boundary tensors exactly
as the reference's 'random' initial condition draws them -- C[k] = torch.rand(chi,chi), E[k] = torch.rand(chi,chi,D,D) on
the CPU in float32, then cast (acetn/ipeps/site_tensor.py:166-168,192-194 with the cast of __setitem__ :77-83).
Generated on the CPU and moved, so CPU and GPU runs share inputs."""
import torch

from .ipeps import CTMRGConfig, Ipeps, SiteTensor


def random_site(D, d, chi):
    A = torch.rand(D, D, D, D, d, dtype=torch.float64) - 0.5
    A = A / A.norm()
    C = [torch.rand(chi, chi).to(torch.float64) for _ in range(4)]
    E = [torch.rand(chi, chi, D, D).to(torch.float64) for _ in range(4)]
    return SiteTensor(A, C, E)


def random_ipeps(nx, ny, D, chi, d=2, seed=0, ctmrg=None, device="cuda"):
    torch.manual_seed(seed)
    sites = {}
    for x in range(nx):
        for y in range(ny):
            sites[(x, y)] = random_site(D, d, chi)
    return Ipeps(nx, ny, {"phys": d, "bond": D, "chi": chi}, sites, ctmrg or CTMRGConfig(), device)


def flops_site_move(D, chi, d=2, niter=2, p=2, chi_new=None):
    """Algorithmic flops of one site-move as the reference executes it: two quarter tensors, the
    (4+4 niter) + ... thin products of the fused randomized SVD, two projector products, three absorptions."""
    m = chi * D * D
    q = min(chi + p, m)
    xn = chi if chi_new is None else chi_new
    FQ = 2 * chi ** 3 * D ** 2 + 2 * chi ** 3 * D ** 4 + 4 * chi ** 2 * D ** 6 * d
    FR = (4 + 4 * niter) * 2 * m * m * q + 2 * m * q * q
    FP = 4 * m * m * xn
    FA = 8 * chi ** 3 * D ** 2 + 4 * chi ** 3 * D ** 4 + 4 * chi ** 2 * D ** 6 * d
    return 2 * FQ + FR + FP + FA


def flops_sweep(nx, ny, D, chi, d=2, niter=2, p=2):
    return 4 * nx * ny * flops_site_move(D, chi, d, niter, p)
