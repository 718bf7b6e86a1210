"""B200 counterparts of acetn/renormalization: ProjectorCalculator, DirectionalMover, ctmrg.

Class/method names, argument meaning, tensor layouts, mutation semantics and error behaviour follow the reference
(acetn/renormalization/{projectors,directional_mover,ctmrg}.py) so this module drops in behind
`Ipeps.renormalize()` (see acetn_b200.integration).  All contractions, the randomized SVD and the normalisations
run in libacetn_b200.so; torch only allocates and draws Omega."""
import torch

from . import linalg, ops


class ProjectorCalculator:
    """acetn/renormalization/projectors.py:6-236."""

    def __init__(self, config):
        self.projectors = config.projectors
        self.svd_type = config.svd_type
        self.svd_cutoff = config.svd_cutoff
        self.rsvd_niter = config.rsvd_niter
        self.rsvd_oversampling = config.rsvd_oversampling
        self.spectra = None          # optional recorder: list receiving the normalised spectrum of every projector
        self.set_calculate()

    def set_calculate(self):
        if self.projectors is None or self.projectors == "full-system":
            self.calculate = self.calculate_full_system
        elif self.projectors == "half-system":
            self.calculate = self.calculate_half_system
        else:
            raise ValueError(f"Invalid ctmrg projector type: {self.projectors} provided.")

    @staticmethod
    def make_quarter_tensor(site_tensor, k):
        """projectors.py:36-60 -> (Q matrix (chi D^2, chi D^2), 6-tuple shape)."""
        ak = site_tensor.bond_permute(k)
        ck = site_tensor['C'][(0 + k) % 4]
        ek1 = site_tensor['E'][(3 + k) % 4]
        ek2 = site_tensor['E'][(0 + k) % 4]
        return ops.quarter_tensor(ck, ek2, ek1, ak, normalize=True)

    def _truncate(self, S, info, chi):
        # projectors.py:163-164 : s/=s[0]; chi' = min(chi, #{s > cutoff}) -- the one host sync per projector
        keep = int(info[0].item())
        if self.spectra is not None:
            self.spectra.append((S / S[0]).detach().cpu())
        return keep

    def calculate_half_system(self, ipeps, sites, k):
        """projectors.py:138-174."""
        if self.svd_type == "full-rank":
            raise NotImplementedError("backend='b200': svd_type='full-rank' is not implemented (use 'rsvd')")
        s1, s4 = sites[0], sites[3]
        Q1, q1D = self.make_quarter_tensor(ipeps[s1], k)
        Q4, q4D = self.make_quarter_tensor(ipeps[s4], k + 3)
        chi = ipeps.dims["chi"]
        q = min(chi + self.rsvd_oversampling, Q1.shape[0], Q4.shape[1])
        omega = linalg._omega(Q4.shape[1], q, Q1.dtype, Q1.device)
        U, S, V, info = ops.rsvd([Q1, Q4], omega, niter=self.rsvd_niter, reorth_adjoint=False, chi=chi, cutoff=self.svd_cutoff)
        keep = self._truncate(S, info, chi)
        p1, p2 = ops.projectors_from_usv(Q1, Q4, U, V, S, keep)
        return p1.view(*q1D[3:], keep), p2.view(*q4D[:3], keep)

    def calculate_full_system(self, ipeps, sites, k):
        """projectors.py:176-217 (rsvd branch): rSVD of (Q2 Q1)(Q4 Q3), proj1 = Q1^H (Q2^H conj(U)), proj2 = Q4 (Q3 V)."""
        if self.svd_type == "full-rank":
            raise NotImplementedError("backend='b200': svd_type='full-rank' is not implemented (use 'rsvd')")
        s1, s2, s3, s4 = sites
        Q1, q1D = self.make_quarter_tensor(ipeps[s1], k)
        Q2, _ = self.make_quarter_tensor(ipeps[s2], k + 1)
        Q3, _ = self.make_quarter_tensor(ipeps[s3], k + 2)
        Q4, q4D = self.make_quarter_tensor(ipeps[s4], k + 3)
        chi = ipeps.dims["chi"]
        q = min(chi + self.rsvd_oversampling, Q2.shape[0], Q3.shape[1])
        omega = linalg._omega(Q3.shape[1], q, Q1.dtype, Q1.device)
        U, S, V, info = ops.rsvd([Q2, Q1, Q4, Q3], omega, niter=self.rsvd_niter, reorth_adjoint=True, chi=chi, cutoff=self.svd_cutoff)
        keep = self._truncate(S, info, chi)
        w = 1.0 / torch.sqrt(S[:keep] / S[0])
        Us = (U[:, :keep] * w).contiguous()
        Vs = (V[:, :keep] * w).contiguous()
        p1 = ops.matmul(Q1, ops.matmul(Q2, Us, transpose_a=True), transpose_a=True)
        p2 = ops.matmul(Q4, ops.matmul(Q3, Vs))
        return p1.view(*q1D[3:], keep), p2.view(*q4D[:3], keep)


class DirectionalMover:
    """acetn/renormalization/directional_mover.py:5-366 (non-distributed moves)."""

    def __init__(self, config):
        self.projector_calculator = ProjectorCalculator(config)
        self.calculate_projectors = self.projector_calculator.calculate

    # ---- the four moves (directional_mover.py:23-97) -----------------------------------------------------------
    def left_move(self, ipeps, xi):
        proj1, proj2 = {}, {}
        for yi in range(ipeps.ny):
            proj1[yi], proj2[yi] = self.calculate_left_projectors(ipeps, xi, yi)
        for yi in range(ipeps.ny):
            xj = (xi + 1) % ipeps.nx
            yj = (yi + 1) % ipeps.ny
            self.renormalize_boundary(ipeps, proj1, proj2, (xi, yi), (xj, yi), yi, yj, k=0)

    def up_move(self, ipeps, yi):
        proj1, proj2 = {}, {}
        for xi in range(ipeps.nx):
            proj1[xi], proj2[xi] = self.calculate_up_projectors(ipeps, xi, yi)
        for xi in range(ipeps.nx):
            xj = (xi + 1) % ipeps.nx
            yj = (yi - 1 + ipeps.ny) % ipeps.ny
            self.renormalize_boundary(ipeps, proj1, proj2, (xi, yi), (xi, yj), xi, xj, k=1)

    def right_move(self, ipeps, xi):
        proj1, proj2 = {}, {}
        for yi in range(ipeps.ny):
            proj1[yi], proj2[yi] = self.calculate_right_projectors(ipeps, xi, yi)
        for yi in range(ipeps.ny):
            xj = (xi - 1 + ipeps.nx) % ipeps.nx
            yj = (yi - 1 + ipeps.ny) % ipeps.ny
            self.renormalize_boundary(ipeps, proj1, proj2, (xi, yi), (xj, yi), yi, yj, k=2)

    def down_move(self, ipeps, yi):
        proj1, proj2 = {}, {}
        for xi in range(ipeps.nx):
            proj1[xi], proj2[xi] = self.calculate_down_projectors(ipeps, xi, yi)
        for xi in range(ipeps.nx):
            xj = (xi - 1 + ipeps.nx) % ipeps.nx
            yj = (yi + 1) % ipeps.ny
            self.renormalize_boundary(ipeps, proj1, proj2, (xi, yi), (xi, yj), xi, xj, k=3)

    # ---- plaquette pickers (directional_mover.py:99-181) ----------------------------------------------------------
    def calculate_left_projectors(self, ipeps, xi, yi):
        xj, yj = (xi + 1) % ipeps.nx, (yi - 1 + ipeps.ny) % ipeps.ny
        return self.calculate_projectors(ipeps, [(xi, yi), (xj, yi), (xj, yj), (xi, yj)], k=0)

    def calculate_right_projectors(self, ipeps, xi, yi):
        xj, yj = (xi - 1 + ipeps.nx) % ipeps.nx, (yi + 1) % ipeps.ny
        return self.calculate_projectors(ipeps, [(xi, yi), (xj, yi), (xj, yj), (xi, yj)], k=2)

    def calculate_up_projectors(self, ipeps, xi, yi):
        xj, yj = (xi - 1 + ipeps.nx) % ipeps.nx, (yi - 1 + ipeps.ny) % ipeps.ny
        return self.calculate_projectors(ipeps, [(xi, yi), (xi, yj), (xj, yj), (xj, yi)], k=1)

    def calculate_down_projectors(self, ipeps, xi, yi):
        xj, yj = (xi + 1) % ipeps.nx, (yi + 1) % ipeps.ny
        return self.calculate_projectors(ipeps, [(xi, yi), (xi, yj), (xj, yj), (xj, yi)], k=3)

    # ---- absorption (directional_mover.py:273-366) ------------------------------------------------------------------
    def renormalize_boundary(self, ipeps, proj1, proj2, s1, s2, i, j, k):
        src, dst = ipeps[s1], ipeps[s2]
        dst['C'][(3 + k) % 4] = self.renormalize_cj1(src['C'][(3 + k) % 4], src['E'][(2 + k) % 4], proj1[i])
        dst['C'][k] = self.renormalize_cj2(src['C'][k], src['E'][k], proj2[j])
        dst['E'][(3 + k) % 4] = self.renormalize_ej(src['E'][(3 + k) % 4], src.bond_permute(k), proj2[i], proj1[j])

    @staticmethod
    def renormalize_cj1(ci, ei, proj):
        return ops.absorb_corner1(ci, ei, proj)

    @staticmethod
    def renormalize_cj2(ci, ei, proj):
        return ops.absorb_corner2(ci, ei, proj)

    @staticmethod
    def renormalize_ej(ei, ai, proj2, proj1):
        return ops.absorb_edge(ei, ai, proj2, proj1)


def ctmrg(ipeps, config, mover=None):
    """acetn/renormalization/ctmrg.py:4-31 (non-distributed ordering; the sharded schedule lives in
    acetn_b200.distributed)."""
    mover = mover or DirectionalMover(config)
    for _ in range(config.steps):
        for xi in range(ipeps.nx):
            mover.left_move(ipeps, xi)
            mover.right_move(ipeps, (ipeps.nx - xi + 1) % ipeps.nx)
        for yi in range(ipeps.ny):
            mover.up_move(ipeps, (ipeps.ny - yi + 1) % ipeps.ny)
            mover.down_move(ipeps, yi)
    return mover
