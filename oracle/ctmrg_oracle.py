"""
CPU oracle for the CTMRG hot path of ace-tn/ace-tn  --  TEST INFRASTRUCTURE ONLY.

This module restates, on CPU torch float64, the algorithm of the reference's CTMRG
renormalization path so that the CUDA product path (acetn_b200/) can be checked
against it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  The product never routes through this file.

Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
  (i)  the reference's own known-answer energies (tests/integration/ipeps_gs/energies.csv,
       rel 1e-10, reference test tests/integration/test_ground_states.py:30) on the two
       converged states shipped by the reference (converted to plain tensors by
       tests/golden/make_golden.py), and
  (ii) golden vectors produced by importing the reference itself in the build container
       (tests/golden/make_golden.py: quarter tensors, rSVD spectra with recorded Omega,
       projectors, absorbed C/E, RDMs, energies after sweeps), and
  (iii) the reference's norm tensor, ALS iterates and whole FullUpdater.tensor_update outputs
       (tests/golden/make_golden_fu.py -> ref_vectors_fu.pt; reproduced bit for bit).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Conventions: A[l,u,r,d,p]; k: 0=left 1=up 2=right
3=down; C[k] (chi,chi); E[k] (chi,chi,D_ket,D_bra).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch

F64 = torch.float64


# --------------------------------------------------------------------------------------
# data model (plain containers; acetn/ipeps/site_tensor.py:41-150, tensor_network.py:153-193)
# --------------------------------------------------------------------------------------
class Site:
    """One unit-cell site: A (D,D,D,D,d), C = 4 corners, E = 4 edges.
    Mirrors the reference's SiteTensor accessors (site_tensor.py:41-60,138-150)."""

    def __init__(self, A, C, E):
        self.A = A
        self.C = list(C)
        self.E = list(E)

    def __getitem__(self, key):
        if key == "A":
            return self.A
        if key == "C":
            return self.C
        if key == "E":
            return self.E
        raise ValueError(f"Invalid key: '{key}' provided.")

    def bond_permute(self, k):
        # site_tensor.py:138-150 : strided view, never copied
        return self.A.permute([(i + k) % 4 for i in range(4)] + [4])

    def clone(self):
        return Site(self.A.clone(), [c.clone() for c in self.C], [e.clone() for e in self.E])


class Cell:
    """nx x ny unit cell of Sites; site_list / bond_list as tensor_network.py:153-193."""

    def __init__(self, nx, ny, dims, sites: Dict[Tuple[int, int], Site]):
        self.nx, self.ny = nx, ny
        self.dims = dict(dims)
        self.sites = sites
        self.site_list = [(x, y) for x in range(nx) for y in range(ny)]
        self.bond_list = [((x, y), ((x + 1) % nx, y), 2) for x in range(nx) for y in range(ny)] + \
                         [((x, y), (x, (y + 1) % ny), 1) for y in range(ny) for x in range(nx)]

    def __getitem__(self, site):
        s = self.sites.get(tuple(site))
        if s is None:
            raise ValueError(f"Site tensor not defined at site {site}.")
        return s

    def clone(self):
        return Cell(self.nx, self.ny, self.dims, {k: v.clone() for k, v in self.sites.items()})


def product_state_site(D, d, chi, site_state=(1.0, 0.0), noise=1e-2, dtype=F64):
    """site_tensor.py:120-136 (A = noise*rand(float32 CPU) + product state, normalised) and
    :152-202 (C,E from the double-layer of A)."""
    A = (noise * torch.rand(D, D, D, D, d)).to(dtype)
    for n, v in enumerate(site_state):
        A[0, 0, 0, 0, n] += v
    A = A / A.norm()
    s = Site(A, [], [])
    for k in range(4):
        ak = s.bond_permute(k)
        ck = torch.einsum("lurdp,luRDp->dDrR", ak, ak.conj()).reshape(D * D, D * D)
        s.C.append(ck / ck.norm())
        ek = torch.einsum("lurdp,LuRDp->lLrRdD", ak, ak.conj()).reshape(D * D, D * D, D, D)
        s.E.append(ek / ek.norm())
    return s


def random_site(D, d, chi, dtype=F64):
    """Synthetic benchmark input of SURVEY.md 8(d): A = rand-0.5 normalised (float64 draw);
    C,E from the reference's 'random' branch: float32 CPU torch.rand then cast
    (site_tensor.py:166-168,192-194 with the cast of __setitem__ :77-83)."""
    A = torch.rand(D, D, D, D, d, dtype=dtype) - 0.5
    A = A / A.norm()
    C = [torch.rand(chi, chi).to(dtype) for _ in range(4)]
    E = [torch.rand(chi, chi, D, D).to(dtype) for _ in range(4)]
    return Site(A, C, E)


def random_cell(nx, ny, D, chi, d=2, seed=0):
    torch.manual_seed(seed)
    sites = {}
    for x in range(nx):
        for y in range(ny):
            sites[(x, y)] = random_site(D, d, chi)
    return Cell(nx, ny, {"phys": d, "bond": D, "chi": chi}, sites)


def product_cell(nx, ny, D, chi, d=2, seed=0, state_map=None):
    torch.manual_seed(seed)
    sites = {}
    for x in range(nx):
        for y in range(ny):
            st = state_map((x, y)) if state_map else [1.0] + [0.0] * (d - 1)
            sites[(x, y)] = product_state_site(D, d, chi, st)
    return Cell(nx, ny, {"phys": d, "bond": D, "chi": chi}, sites)


# --------------------------------------------------------------------------------------
# configuration (acetn/ipeps/ipeps_config.py:16-24 defaults)
# --------------------------------------------------------------------------------------
@dataclass
class CtmrgConfig:
    steps: int = 40
    projectors: str = "half-system"
    svd_type: str = "rsvd"
    svd_cutoff: float = 1e-12
    rsvd_niter: int = 2
    rsvd_oversampling: int = 2
    disable_progressbar: bool = True


class OmegaTape:
    """Records (or replays) the Gaussian test matrices of the randomized SVD so that two
    implementations consume identical Omega (SURVEY.md §7 hard part 1).
    In record mode draws torch.randn(n, q, dtype=float64) on CPU exactly like
    fused_matmul_svd_lowrank.py:32 does on a CPU run."""

    def __init__(self, replay: Optional[List[torch.Tensor]] = None):
        self.tape: List[torch.Tensor] = list(replay) if replay is not None else []
        self.replay = replay is not None
        self.pos = 0

    def __call__(self, n, q, dtype=F64, device="cpu"):
        if self.replay:
            om = self.tape[self.pos]
            self.pos += 1
            assert tuple(om.shape) == (n, q), f"omega tape shape {tuple(om.shape)} != {(n, q)}"
            return om.to(device=device, dtype=dtype)
        om = torch.randn(n, q, dtype=dtype)
        self.tape.append(om)
        return om


def _default_omega(n, q, dtype=F64, device="cpu"):
    return torch.randn(n, q, dtype=dtype, device=device)


# --------------------------------------------------------------------------------------
# randomized SVD family (acetn/linalg)
# --------------------------------------------------------------------------------------
def svd_lowrank(A, q=6, niter=2, omega_fn: Callable = _default_omega):
    """acetn/linalg/svd_lowrank.py:26-44."""
    m, n = A.shape
    q = min(q, m, n)
    Y = A @ omega_fn(n, q, A.dtype, A.device)
    for _ in range(niter):
        Y = torch.linalg.qr(Y).Q
        Y = A @ (A.mH @ Y)
    Q = torch.linalg.qr(Y).Q
    U_B, S, Vh = torch.linalg.svd(Q.mH @ A, full_matrices=False)
    return Q @ U_B, S, Vh.mH


def fused_matmul_svd_lowrank(A, B, q=6, niter=2, omega_fn: Callable = _default_omega):
    """acetn/linalg/fused_matmul_svd_lowrank.py:26-52 : rSVD of A@B without forming it."""
    m, n = A.shape[0], B.shape[1]
    q = min(q, m, n)
    Y = A @ (B @ omega_fn(n, q, A.dtype, A.device))
    for _ in range(niter):
        Y = torch.linalg.qr(Y).Q
        Y = A @ (B @ (B.mH @ (A.mH @ Y)))
    Q = torch.linalg.qr(Y).Q
    Bt = (Q.mH @ A) @ B
    U_B, S, Vh = torch.linalg.svd(Bt, full_matrices=False)
    return Q @ U_B, S, Vh.mH


def fused_3matmul_svd_lowrank(A, B, C, D, q=6, niter=2, omega_fn: Callable = _default_omega):
    """acetn/linalg/fused_3matmul_svd_lowrank.py:28-56 : rSVD of (A@B)@(C@D); note the extra
    re-orthonormalisation between the adjoint and forward halves (:39-45)."""
    m, n = A.shape[0], D.shape[1]
    q = min(q, m, n)
    Y = A @ (B @ (C @ (D @ omega_fn(n, q, A.dtype, A.device))))
    for _ in range(niter):
        Y = torch.linalg.qr(Y).Q
        Y = D.mH @ (C.mH @ (B.mH @ (A.mH @ Y)))
        Y = torch.linalg.qr(Y).Q
        Y = A @ (B @ (C @ (D @ Y)))
    Q = torch.linalg.qr(Y).Q
    Bt = (((Q.mH @ A) @ B) @ C) @ D
    U_B, S, Vh = torch.linalg.svd(Bt, full_matrices=False)
    return Q @ U_B, S, Vh.mH


# --------------------------------------------------------------------------------------
# projectors (acetn/renormalization/projectors.py)
# --------------------------------------------------------------------------------------
def quarter_tensor(site: Site, k: int):
    """projectors.py:36-60 : Q_k[(c,r,R),(e,d,D)] = C.E.E.A*.A, divided by its max-abs."""
    ak = site.bond_permute(k)
    ck = site["C"][k % 4]
    ek1 = site["E"][(3 + k) % 4]
    ek2 = site["E"][k % 4]
    t = torch.einsum("ab,bcuU->acuU", ck, ek2)
    t = torch.einsum("acuU,ealL->cuUelL", t, ek1)
    t = torch.einsum("cuUelL,LURDP->cuelRDP", t, ak.conj())
    t = torch.einsum("lurdp,cuelRDp->crRedD", ak, t)
    shp = tuple(t.shape)
    t = t.reshape(shp[0] * shp[1] * shp[2], shp[3] * shp[4] * shp[5])
    return t / t.abs().max(), shp


def truncate_usv(U, s, V, chi, cutoff):
    """projectors.py:163-167 : s/=s[0]; chi' = min(chi, #{s>cutoff}); scale columns by s^-1/2."""
    s = s / s[0]
    keep = min(chi, int((s > cutoff).sum()))
    w = 1.0 / torch.sqrt(s[:keep])
    return U[:, :keep] * w, V[:, :keep] * w, s, keep


def half_system_projectors(cell: Cell, sites, k, cfg: CtmrgConfig, omega_fn=_default_omega,
                           record: Optional[dict] = None):
    """projectors.py:138-174 (rsvd and full-rank branches)."""
    chi = cell.dims["chi"]
    s1, s4 = sites[0], sites[3]
    Q1, d1 = quarter_tensor(cell[s1], k)
    Q4, d4 = quarter_tensor(cell[s4], k + 3)
    if cfg.svd_type == "full-rank":
        R = Q1 @ Q4
        R = R / R.abs().max()
        U, s, Vh = torch.linalg.svd(R)
        V = Vh.mH
    else:
        U, s, V = fused_matmul_svd_lowrank(Q1, Q4, q=chi + cfg.rsvd_oversampling,
                                           niter=cfg.rsvd_niter, omega_fn=omega_fn)
    U, V, sn, keep = truncate_usv(U, s, V, chi, cfg.svd_cutoff)
    proj1 = torch.einsum("xedD,xz->edDz", Q1.view(Q1.shape[0], *d1[3:]), U.conj())
    proj2 = torch.einsum("cuUy,yz->cuUz", Q4.view(*d4[:3], Q4.shape[1]), V)
    if record is not None:
        record.setdefault("spectra", []).append(sn.clone())
    return proj1, proj2


def full_system_projectors(cell: Cell, sites, k, cfg: CtmrgConfig, omega_fn=_default_omega,
                           record: Optional[dict] = None):
    """projectors.py:176-217 (rsvd branch) and :84-136 (full-rank branch)."""
    chi = cell.dims["chi"]
    s1, s2, s3, s4 = sites
    Q1, d1 = quarter_tensor(cell[s1], k)
    Q2, _ = quarter_tensor(cell[s2], k + 1)
    Q3, _ = quarter_tensor(cell[s3], k + 2)
    Q4, d4 = quarter_tensor(cell[s4], k + 3)
    if cfg.svd_type == "full-rank":
        R1 = Q2 @ Q1
        R2 = Q4 @ Q3
        R1 = R1 / R1.abs().max()
        R2 = R2 / R2.abs().max()
        Fm = R1 @ R2
        Fm = Fm / Fm.abs().max()
        U, s, Vh = torch.linalg.svd(Fm)
        V = Vh.mH
        U, V, sn, keep = truncate_usv(U, s, V, chi, cfg.svd_cutoff)
        proj1 = torch.einsum("xedD,xz->edDz", R1.view(R1.shape[0], *d1[3:]), U.conj())
        proj2 = torch.einsum("cuUy,yz->cuUz", R2.view(*d4[:3], R2.shape[1]), V)
    else:
        U, s, V = fused_3matmul_svd_lowrank(Q2, Q1, Q4, Q3, q=chi + cfg.rsvd_oversampling,
                                            niter=cfg.rsvd_niter, omega_fn=omega_fn)
        U, V, sn, keep = truncate_usv(U, s, V, chi, cfg.svd_cutoff)
        proj1 = (Q1.mH @ (Q2.mH @ U.conj())).view(*d1[3:], keep)
        proj2 = (Q4 @ (Q3 @ V)).view(*d4[:3], keep)
    if record is not None:
        record.setdefault("spectra", []).append(sn.clone())
    return proj1, proj2


# --------------------------------------------------------------------------------------
# absorption (acetn/renormalization/directional_mover.py:273-366)
# --------------------------------------------------------------------------------------
def absorb_corner1(ci, ei, proj):
    """renormalize_cj1, directional_mover.py:306-323."""
    t = torch.einsum("ablL,bc->alLc", ei, ci)
    t = torch.einsum("alLc,clLx->ax", t, proj)
    return t / t.norm()


def absorb_corner2(ci, ei, proj):
    """renormalize_cj2, directional_mover.py:325-343."""
    t = torch.einsum("ab,bcrR->acrR", ci, ei)
    t = torch.einsum("arRx,acrR->xc", proj, t)
    return t / t.norm()


def absorb_edge(ei, ai, proj2, proj1):
    """renormalize_ej, directional_mover.py:345-366."""
    t = torch.einsum("ablL,buUx->alLuUx", ei, proj1)
    t = torch.einsum("LURDP,alLuUx->RDPalux", ai.conj(), t)
    t = torch.einsum("lurdp,RDpalux->rdRDax", ai, t)
    t = torch.einsum("rdRDax,adDy->yxrR", t, proj2)
    return t / t.norm()


def renormalize_boundary(cell: Cell, proj1, proj2, s1, s2, i, j, k):
    """directional_mover.py:273-303 : writes C[(3+k)%4], C[k], E[(3+k)%4] of site s2."""
    a, b = cell[s1], cell[s2]
    b.C[(3 + k) % 4] = absorb_corner1(a.C[(3 + k) % 4], a.E[(2 + k) % 4], proj1[i])
    b.C[k] = absorb_corner2(a.C[k], a.E[k], proj2[j])
    b.E[(3 + k) % 4] = absorb_edge(a.E[(3 + k) % 4], a.bond_permute(k), proj2[i], proj1[j])


# plaquette pickers, directional_mover.py:99-181
def plaquette(cell: Cell, k: int, xi: int, yi: int):
    nx, ny = cell.nx, cell.ny
    if k == 0:
        xj, yj = (xi + 1) % nx, (yi - 1 + ny) % ny
        return [(xi, yi), (xj, yi), (xj, yj), (xi, yj)]
    if k == 2:
        xj, yj = (xi - 1 + nx) % nx, (yi + 1) % ny
        return [(xi, yi), (xj, yi), (xj, yj), (xi, yj)]
    if k == 1:
        xj, yj = (xi - 1 + nx) % nx, (yi - 1 + ny) % ny
        return [(xi, yi), (xi, yj), (xj, yj), (xj, yi)]
    if k == 3:
        xj, yj = (xi + 1) % nx, (yi + 1) % ny
        return [(xi, yi), (xi, yj), (xj, yj), (xj, yi)]
    raise ValueError(k)


def move_tasks(cell: Cell, k: int, line: int):
    """The per-site tasks of one directional move (directional_mover.py:23-97).
    Returns a list of (key, plaquette sites, s1, s2, i, j)."""
    nx, ny = cell.nx, cell.ny
    out = []
    if k == 0:      # left_move(xi=line)
        for yi in range(ny):
            out.append((yi, plaquette(cell, 0, line, yi), (line, yi), ((line + 1) % nx, yi), yi, (yi + 1) % ny))
    elif k == 2:    # right_move(xi=line)
        for yi in range(ny):
            out.append((yi, plaquette(cell, 2, line, yi), (line, yi), ((line - 1 + nx) % nx, yi), yi, (yi - 1 + ny) % ny))
    elif k == 1:    # up_move(yi=line)
        for xi in range(nx):
            out.append((xi, plaquette(cell, 1, xi, line), (xi, line), (xi, (line - 1 + ny) % ny), xi, (xi + 1) % nx))
    elif k == 3:    # down_move(yi=line)
        for xi in range(nx):
            out.append((xi, plaquette(cell, 3, xi, line), (xi, line), (xi, (line + 1) % ny), xi, (xi - 1 + nx) % nx))
    else:
        raise ValueError(k)
    return out


def directional_move(cell: Cell, k: int, line: int, cfg: CtmrgConfig, omega_fn=_default_omega,
                     record: Optional[dict] = None):
    """left/up/right/down_move: all projectors of the line first, then all absorptions."""
    calc = half_system_projectors if cfg.projectors == "half-system" else full_system_projectors
    if cfg.projectors not in ("half-system", "full-system", None):
        raise ValueError(f"Invalid ctmrg projector type: {cfg.projectors} provided.")
    tasks = move_tasks(cell, k, line)
    p1, p2 = {}, {}
    for key, plaq, *_ in tasks:
        p1[key], p2[key] = calc(cell, plaq, k, cfg, omega_fn, record)
    for key, plaq, s1, s2, i, j in tasks:
        renormalize_boundary(cell, p1, p2, s1, s2, i, j, k)


def sweep(cell: Cell, cfg: CtmrgConfig, omega_fn=_default_omega, record: Optional[dict] = None):
    """One CTMRG sweep, ctmrg.py:25-31 (non-distributed ordering)."""
    nx, ny = cell.nx, cell.ny
    for xi in range(nx):
        directional_move(cell, 0, xi, cfg, omega_fn, record)
        directional_move(cell, 2, (nx - xi + 1) % nx, cfg, omega_fn, record)
    for yi in range(ny):
        directional_move(cell, 1, (ny - yi + 1) % ny, cfg, omega_fn, record)
        directional_move(cell, 3, yi, cfg, omega_fn, record)


def ctmrg(cell: Cell, cfg: CtmrgConfig, omega_fn=_default_omega, record: Optional[dict] = None):
    for _ in range(cfg.steps):
        sweep(cell, cfg, omega_fn, record)


# --------------------------------------------------------------------------------------
# measurement (acetn/measurement/rdm.py, measure.py)
# --------------------------------------------------------------------------------------
def site_rdm(cell: Cell, site):
    """rdm.py:35-67 : rho[P,p] (bra, ket)."""
    s = cell[site]
    c1, c2, c3, c4 = s.C
    e1, e2, e3, e4 = s.E
    a1 = s.A
    t1 = torch.einsum("ab,bclL->aclL", c4, e4)
    t1 = torch.einsum("aclL,eadD->clLedD", t1, e3)
    t1 = torch.einsum("clLedD,LURDP->cledURP", t1, a1.conj())
    t2 = torch.einsum("ab,bcuU->acuU", c1, e1)
    t3 = torch.einsum("ab,carR->bcrR", c3, e2)
    t3 = torch.einsum("ec,bcrR->ebrR", c2, t3)
    t3 = torch.einsum("ebrR,aeuU->brRauU", t3, t2)
    t3 = torch.einsum("erRcuU,cledURP->ruldP", t3, t1)
    return torch.einsum("ruldP,lurdp->Pp", t3, a1)


def bond_rdm(cell: Cell, bond):
    """rdm.py:69-121 : rho[P,Q,p,q]; unblocked form of build_bond_rdm_core (the blocked
    variant :123-154 computes the same tensor d^2 blocks at a time)."""
    s1, s2, k = bond
    a, b = cell[s1], cell[s2]
    c12, e12, e11 = a.C[(k + 1) % 4], a.E[(k + 1) % 4], a.E[k % 4]
    c13, e13 = a.C[(k + 2) % 4], a.E[(k + 2) % 4]
    a1 = a.bond_permute(k)
    c21, e21, e24 = b.C[k % 4], b.E[k % 4], b.E[(k + 3) % 4]
    c24, e23 = b.C[(k + 3) % 4], b.E[(k + 2) % 4]
    a2 = b.bond_permute(k)

    t = torch.einsum("ab,bcrR->acrR", c12, e12)
    t = torch.einsum("acrR,eauU->crReuU", t, e11)
    t = torch.einsum("crReuU,LURDP->creuLDP", t, a1.conj())
    t = torch.einsum("creuLDP,lurdp->ceLDPldp", t, a1)
    r1 = torch.einsum("ab,bfdD->afdD", c13, e13)
    r1 = torch.einsum("afdD,acLDPldp->fcLPlp", r1, t)

    t = torch.einsum("ab,bcuU->acuU", c21, e21)
    t = torch.einsum("acuU,ealL->cuUelL", t, e24)
    t = torch.einsum("cuUelL,LURDQ->cuelRDQ", t, a2.conj())
    t = torch.einsum("cuelRDQ,lurdq->ceRDQrdq", t, a2)
    r2 = torch.einsum("ae,fadD->efdD", c24, e23)
    r2 = torch.einsum("efdD,ceRDQrdq->fcRQrq", r2, t)
    return torch.einsum("fcRPrp,fcRQrq->PQpq", r1, r2)


def pauli(dtype=F64):
    X = torch.tensor([[0.0, 1.0], [1.0, 0.0]], dtype=dtype)
    Z = torch.tensor([[1.0, 0.0], [0.0, -1.0]], dtype=dtype)
    iY = torch.tensor([[0.0, 1.0], [-1.0, 0.0]], dtype=dtype)    # i*sigma_y (real)
    I = torch.eye(2, dtype=dtype)
    return X, iY, Z, I


def heisenberg_bond_hamiltonian(J=1.0):
    """acetn/model/models/heisenberg.py:25-29 : 0.25 J (XX + YY + ZZ); YY = -(iY)(iY)."""
    X, iY, Z, _ = pauli()
    return 0.25 * J * (torch.kron(X, X) - torch.kron(iY, iY) + torch.kron(Z, Z))


def ising_hamiltonians(jz=1.0, hx=0.0):
    """acetn/model/models/ising.py:13-21 : site -hx X ; bond -jz ZZ."""
    X, _, Z, _ = pauli()
    return -hx * X, -jz * torch.kron(Z, Z)


def measure(cell: Cell, bond_ham, site_ham=None, site_ops: Optional[Callable] = None):
    """measure.py:5-28,114-198 : energy per site and averaged one-site observables.
    bond_ham: (d^2,d^2) matrix; site_ham: (d,d) or None; site_ops(site)->{name: (d,d)}."""
    d = cell.dims["phys"]
    out = {"Energy": torch.zeros((), dtype=F64)}
    names = list(site_ops(cell.site_list[0]).keys()) if site_ops else []
    for n in names:
        out[n] = torch.zeros((), dtype=F64)
    for site in cell.site_list:
        rho = site_rdm(cell, site)
        nrm = torch.einsum("pp->", rho).real
        if site_ham is not None:
            out["Energy"] = out["Energy"] + torch.einsum("Pp,pP->", rho, site_ham).real / nrm
        if site_ops:
            for n, op in site_ops(site).items():
                out[n] = out[n] + torch.einsum("Pp,pP->", rho, op).real / nrm
    for n in names:
        out[n] = out[n] / len(cell.site_list)
    h4 = bond_ham.reshape(d, d, d, d)
    for bond in cell.bond_list:
        rho = bond_rdm(cell, bond)
        nrm = torch.einsum("pqpq->", rho).real
        out["Energy"] = out["Energy"] + torch.einsum("PQpq,pqPQ->", rho, h4).real / nrm
    out["Energy"] = out["Energy"] / len(cell.site_list)
    return out


# --------------------------------------------------------------------------------------
# gauge-invariant comparison helpers
# --------------------------------------------------------------------------------------
def corner_spectra(cell: Cell):
    """Singular values of every corner, normalised by the largest (gauge invariant)."""
    out = {}
    for site in cell.site_list:
        for k in range(4):
            s = torch.linalg.svdvals(cell[site].C[k])
            out[(site, k)] = s / s[0]
    return out


def flops_site_move(D, chi, d=2, niter=2, p=2, chi_new=None):
    """Algorithmic flop model of SURVEY.md 8(d)."""
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


# --------------------------------------------------------------------------------------
# full-update norm tensor and ALS inner solver (acetn/evolution/full_update.py, als_solver.py)
# --------------------------------------------------------------------------------------
def norm_tensor(cell: Cell, bond, a1q, a2q):
    """full_update.py:163-227 : N12[y,x,Y,X] from the bond environment and the QR'd site tensors a1q/a2q (D,D,D,nD)."""
    s1, s2, k = bond
    a, b = cell[s1], cell[s2]
    c12, e12, e11 = a.C[(k + 1) % 4], a.E[(k + 1) % 4], a.E[k % 4]
    c13, e13 = a.C[(k + 2) % 4], a.E[(k + 2) % 4]
    c21, e21, e24 = b.C[k % 4], b.E[k % 4], b.E[(k + 3) % 4]
    c24, e23 = b.C[(k + 3) % 4], b.E[(k + 2) % 4]
    t = torch.einsum("ab,bcrR->acrR", c12, e12)
    t = torch.einsum("acrR,eauU->crReuU", t, e11)
    t = torch.einsum("crReuU,RDUY->creuDY", t, a1q.conj())
    t = torch.einsum("creuDY,rduy->ceDYdy", t, a1q)
    n1 = torch.einsum("ab,bfdD->afdD", c13, e13)
    n1 = torch.einsum("afdD,aeDYdy->feYy", n1, t)
    t = torch.einsum("ab,bcuU->acuU", c21, e21)
    t = torch.einsum("acuU,ealL->cuUelL", t, e24)
    t = torch.einsum("cuUelL,DLUX->cuelXD", t, a2q.conj())
    t = torch.einsum("cuelXD,dlux->ceXDxd", t, a2q)
    n2 = torch.einsum("ab,fadD->bfdD", c24, e23)
    n2 = torch.einsum("bfdD,cbXDxd->fcXx", n2, t)
    return torch.einsum("fcYy,fcXx->yxYX", n1, n2)


def als_cost(a1r, a2r, a12g, n12):
    """als_solver.py:246-257."""
    a12n = torch.einsum("yup,xuq->yxpq", a1r, a2r)
    d2 = torch.einsum("yxYX,yxpq->YXpq", n12, a12n)
    d2 = torch.einsum("YXpq,YXpq->", d2, a12n.conj())
    d3 = torch.einsum("yxYX,yxpq->YXpq", n12, a12g)
    d3 = torch.einsum("YXpq,YXpq->", d3, a12n.conj())
    return d2.real - 2 * d3.real


def _als_solve_ar(R, S, epsilon, method="cholesky"):
    """als_solver.py:197-229: cholesky branch (:218-225) or pinv branch (:226-228)."""
    nD, bD, pD = S.shape
    S = S.reshape(nD * bD, pD)
    R = R.reshape(nD * bD, nD * bD)
    R = 0.5 * (R + R.mH)
    if method == "pinv":
        return (torch.linalg.pinv(R, hermitian=True, rcond=epsilon) @ S).reshape(nD, bD, pD)
    R = R + epsilon * R.abs().max() * torch.eye(R.shape[0], dtype=R.dtype, device=R.device)
    L = torch.linalg.cholesky(R)
    Y = torch.linalg.solve_triangular(L, S, upper=False)
    return torch.linalg.solve_triangular(L.mH, Y, upper=True).reshape(nD, bD, pD)


def als_solve(a1r, a2r, n12g, n12, a12g, niter=100, tol=1e-15, epsilon=1e-12, method="cholesky"):
    """als_solver.py:55-82 (solve_torch) = csrc/evolution/als_solve.cpp:55-105.  Returns (a1r, a2r, iterations run)."""
    d1 = als_cost(a1r, a2r, a12g, n12).abs()
    it = 0
    for i in range(niter):
        it = i + 1
        S = torch.einsum("YXpQ,XUQ->YUp", n12g, a2r.conj())
        R = torch.einsum("yxYX,xuq->yYXuq", n12, a2r)
        R = torch.einsum("yYXuQ,XUQ->YUyu", R, a2r.conj())
        a1r = _als_solve_ar(R, S, epsilon, method)
        S = torch.einsum("YXPq,YVP->XVq", n12g, a1r.conj())
        R = torch.einsum("yxYX,yvp->xYXvp", n12, a1r)
        R = torch.einsum("xYXvP,YVP->XVxv", R, a1r.conj())
        a2r = _als_solve_ar(R, S, epsilon, method)
        d2 = als_cost(a1r, a2r, a12g, n12)
        error = abs(d2 - d1) / d1.abs()
        if error < tol and i > 1:
            break
        d1 = d2
    return a1r, a2r, it


def als_initial_guess(a12g, ar_shape):
    """als_solver.py:117-146 : truncated SVD of the gate-tensor product."""
    nD, bD, pD = ar_shape
    m = torch.einsum("yxpq->ypxq", a12g).reshape(nD * pD, nD * pD)
    U, S, Vh = torch.linalg.svd(m)
    V = Vh.mH
    S = torch.sqrt(S[:bD] / S[0])
    a1r = torch.einsum("ypu,u->yup", U[:, :bD].reshape(nD, pD, bD), S)
    a2r = torch.einsum("xqv,v->xvq", V[:, :bD].reshape(nD, pD, bD), S)
    return a1r, a2r


# --------------------------------------------------------------------------------------
# the callers around the norm tensor / ALS in one full-update bond update:
# QR split, positive approximation, gauge fix, finalisation, recomposition
# (acetn/evolution/tensor_update.py:53-75, full_update.py:34-161, 262-343)
# --------------------------------------------------------------------------------------
def decompose_site_tensors(a1, a2):
    """tensor_update.py:53-68 : a = (environment part q)(bond-local part r); a1, a2 already in the bond frame."""
    bD, pD = a1.shape[3:]
    nD = min(bD ** 3, pD * bD)
    a1q, a1r = torch.linalg.qr(torch.einsum("lurdp->rdulp", a1).reshape(bD ** 3, pD * bD))
    a2q, a2r = torch.linalg.qr(torch.einsum("lurdp->dlurp", a2).reshape(bD ** 3, pD * bD))
    return a1q.reshape(bD, bD, bD, nD), a1r.reshape(nD, bD, pD), a2q.reshape(bD, bD, bD, nD), a2r.reshape(nD, bD, pD)


def recompose_site_tensors(a1q, a1r, a2q, a2r):
    """tensor_update.py:70-75."""
    return torch.einsum("rdux,xlp->lurdp", a1q, a1r), torch.einsum("dlux,xrp->lurdp", a2q, a2r)


def positive_approx(n12, cutoff=1e-12):
    """full_update.py:262-293 : nz with N ~ nz nz^T, the spectrum shifted until its smallest eigenvalue is >= cutoff."""
    nD = n12.shape[0]
    N = n12.reshape(nD ** 2, nD ** 2).clone()
    nw, nz = torch.linalg.eigh(N)
    while nw[0] < cutoff:
        N += 2 * max(cutoff, abs(float(nw[0]))) * torch.eye(nD ** 2, dtype=N.dtype, device=N.device)
        nw, nz = torch.linalg.eigh(N)
    return nz.reshape(nD, nD, nD ** 2) * torch.sqrt(nw)


def gauge_fix(nz, a12g, atol=1e-12):
    """full_update.py:296-343."""
    nD = a12g.shape[0]
    _, nzyr = torch.linalg.qr(torch.einsum("yxz->zxy", nz).reshape(nD ** 3, nD))
    _, nzxr = torch.linalg.qr(torch.einsum("yxz->zyx", nz).reshape(nD ** 3, nD))
    nzyr_inv = torch.linalg.pinv(nzyr, atol=atol)
    nzxr_inv = torch.linalg.pinv(nzxr, atol=atol)
    nz = torch.einsum("yxz,xw->yzw", nz, nzxr_inv)
    nz = torch.einsum("yzw,yv->zvw", nz, nzyr_inv)
    n12 = torch.einsum("zvw,zVW->vwVW", nz, nz.conj())
    a12g = torch.einsum("zx,yxpq->yzpq", nzxr, a12g)
    a12g = torch.einsum("wy,yzpq->wzpq", nzyr, a12g)
    return n12, a12g, nzxr_inv, nzyr_inv


def finalize_reduced_tensors(a1r, a2r, nzxr_inv=None, nzyr_inv=None):
    """full_update.py:121-161 (Fig. 12(b) of arXiv:1405.3259): undo the gauge fix, balance the bond."""
    if nzyr_inv is not None:
        a1r = torch.einsum("yz,zup->yup", nzyr_inv, a1r)
        a2r = torch.einsum("xw,wvq->xvq", nzxr_inv, a2r)
    nD, bD, pD = a1r.shape
    q1, r1 = torch.linalg.qr(torch.einsum("yup->ypu", a1r).reshape(nD * pD, bD))
    q2, r2 = torch.linalg.qr(torch.einsum("xvq->xqv", a2r).reshape(nD * pD, bD))
    U, s, Vh = torch.linalg.svd(torch.einsum("au,bu->ab", r1, r2))
    s = torch.sqrt(s[:bD] / s.norm())
    r1 = torch.einsum("ab,b->ab", U[:, :bD], s)
    r2 = torch.einsum("ba,b->ba", Vh[:bD, :], s)
    a1r = torch.einsum("ypa,au->yup", q1.reshape(nD, pD, bD), r1)
    a2r = torch.einsum("xqb,vb->xvq", q2.reshape(nD, pD, bD), r2)
    return a1r, a2r


def full_update_bond(cell: Cell, bond, a1, a2, gate, use_gauge_fix=True, gauge_fix_atol=1e-12, positive_approx_cutoff=1e-12,
                     als_niter=100, als_tol=1e-15, als_epsilon=1e-12):
    """FullUpdater.tensor_update (full_update.py:34-96): a1, a2 in the bond frame (tensor_update.py:31-32), gate (d,d,d,d)
    -> updated, normalised a1, a2 (still in the bond frame)."""
    a1q, a1r, a2q, a2r = decompose_site_tensors(a1, a2)
    n12 = norm_tensor(cell, bond, a1q, a2q)
    a12g = torch.einsum("yup,xuq->yxpq", a1r, a2r)
    a12g = torch.einsum("yxpq,pqrs->yxrs", a12g, gate)
    nz = positive_approx(n12, cutoff=positive_approx_cutoff)
    inv = (None, None)
    if use_gauge_fix:
        n12, a12g, nzxr_inv, nzyr_inv = gauge_fix(nz, a12g, atol=gauge_fix_atol)
        inv = (nzxr_inv, nzyr_inv)
    else:
        n12 = torch.einsum("xyz,XYz->xyXY", nz, nz.conj())
    b1, b2 = als_initial_guess(a12g, a1r.shape)
    n12g = torch.einsum("yxYX,yxpq->YXpq", n12, a12g)
    b1, b2, _ = als_solve(b1, b2, n12g, n12, a12g, niter=als_niter, tol=als_tol, epsilon=als_epsilon)
    b1, b2 = finalize_reduced_tensors(b1, b2, *inv)
    a1n, a2n = recompose_site_tensors(a1q, b1, a2q, b2)
    return a1n / a1n.norm(), a2n / a2n.norm()


def bond_theta(a1, a2):
    """Gauge-invariant two-site tensor of a bond update result: a1's leg l contracted with a2's leg r (the legs the bond-frame
    recomposition attaches the reduced tensors to, tensor_update.py:72-73); sign / rotation freedom on that leg cancels."""
    return torch.einsum("burdp,LUbDq->urdpLUDq", a1, a2)
