"""Oracle parity AT THE HEADLINE SIZES.  The CPU oracle needs minutes per site-move at D=8, chi=256,
so -- as SURVEY.md 8c prescribes -- the oracle's torch restatement runs ON THE SAME GPU (torch.einsum / @ / linalg.qr /
linalg.svd -> cuBLAS + cuSOLVER; test infrastructure, never the product path) on the same inputs with the same Omega:

  * one synchronized left move of the 2x2 cell at D=8 chi=256 (BASELINE headline), D=6 chi=144 (config 3) and D=7 chi=196 d=4
    (config 5, honeycomb), with BOTH engines of the thin products (K7 = INT8 tensor cores, K1 = FP64 DMMA):
      truncated projector spectra  |ds|/s0 <= 1e-10;   Pi = P2 P1^T (gauge invariant)  rel. Frobenius <= 1e-9;
      singular values of the absorbed C, C, E  <= 1e-9 of the largest;   site RDM of an updated site <= 1e-9;
  * K7 on ill-conditioned inputs at K7-active sizes: a boundary whose chi legs are graded over 10 orders of magnitude
    (quarter tensors graded over 1e20, like the reference's converged Ising state -- DESIGN.md section 9), and factors with a
    prescribed singular spectrum 1 ... 1e-20, against K1 and against torch on the same Omega;
  * the reference's Heisenberg D=3 chi=16 state (degenerate multiplet at the cut) with the K7 engine forced.
"""
import pytest
import torch

from oracle import ctmrg_oracle as orc
from tests.util import cell_from_plain, load_golden

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from acetn_b200 import linalg, ops
    from acetn_b200.ipeps import CTMRGConfig, Ipeps
    from acetn_b200.renormalization import DirectionalMover, ProjectorCalculator

DEV = "cuda"


def cell_to(cell, dev):
    sites = {s: orc.Site(cell[s].A.to(dev), [c.to(dev) for c in cell[s].C], [e.to(dev) for e in cell[s].E]) for s in cell.site_list}
    return orc.Cell(cell.nx, cell.ny, cell.dims, sites)


class DeviceTape:
    """Omega drawn with torch.randn ON THE GPU (as the reference does on a cuda run, fused_matmul_svd_lowrank.py:32), recorded
    for replay."""

    def __init__(self):
        self.tape = []

    def __call__(self, n, q, dtype=torch.float64, device=DEV):
        om = torch.randn(n, q, dtype=dtype, device=DEV)
        self.tape.append(om)
        return om


def pi_rel_diff(p1a, p2a, p1b, p2b, block=2048):
    """|| P2a P1a^T - P2b P1b^T ||_F / || P2b P1b^T ||_F, formed in row blocks (the full matrix is 2 GiB at D=8 chi=256)."""
    m = p2a.shape[0] * p2a.shape[1] * p2a.shape[2]
    A1, A2 = p1a.reshape(-1, p1a.shape[-1]), p2a.reshape(m, -1)
    B1, B2 = p1b.reshape(-1, p1b.shape[-1]), p2b.reshape(m, -1)
    num = den = 0.0
    for r0 in range(0, m, block):
        pa = A2[r0:r0 + block] @ A1.T
        pb = B2[r0:r0 + block] @ B1.T
        num += float(((pa - pb) ** 2).sum())
        den += float((pb ** 2).sum())
    return (num / den) ** 0.5


def sv(t, rows):
    return torch.linalg.svdvals(t.reshape(rows, -1))


def oracle_left_move(cell_gpu, tape):
    """The oracle's left move of column 0 on the GPU -> (state after the move, spectra, proj1, proj2)."""
    cfg = orc.CtmrgConfig()
    rec = {}
    ref = cell_gpu.clone()
    tasks = orc.move_tasks(ref, 0, 0)
    rp1, rp2 = {}, {}
    for key, plaq, *_ in tasks:
        rp1[key], rp2[key] = orc.half_system_projectors(ref, plaq, 0, cfg, tape, rec)
    for key, plaq, s1, s2, i, j in tasks:
        orc.renormalize_boundary(ref, rp1, rp2, s1, s2, i, j, 0)
    return ref, rec["spectra"], rp1, rp2


def b200_left_move(cell_gpu, engine, tape):
    ip = Ipeps.from_plain(cell_gpu, CTMRGConfig(thin_engine=engine), device=DEV)
    mover = DirectionalMover(ip.ctmrg_config)
    mover.projector_calculator.thin_engine = engine
    mover.projector_calculator.spectra = []
    replay = orc.OmegaTape(tape)
    linalg.set_omega_source(replay)
    try:
        gp1, gp2 = mover._projectors_of_tasks(ip, mover.move_tasks(ip, 0, 0))
        spectra = list(mover.projector_calculator.spectra)
        mover.projector_calculator.spectra = None
        replay.pos = 0
        mover.left_move(ip, 0)          # the same projectors again (same Omega), then the absorptions
    finally:
        linalg.set_omega_source(None)
    torch.cuda.synchronize()
    got = orc.Cell(2, 2, ip.dims, {s: orc.Site(ip[s]['A'], list(ip[s]['C']), list(ip[s]['E'])) for s in ip.site_list})
    return got, spectra, {k[1]: v for k, v in gp1.items()}, {k[1]: v for k, v in gp2.items()}


def run_left_move(cell_gpu, engine):
    """Oracle (torch on the GPU) and the B200 path on the same state and Omega."""
    tape = DeviceTape()
    ref = oracle_left_move(cell_gpu, tape)
    got = b200_left_move(cell_gpu, engine, tape.tape)
    return ref, got, tape.tape


def deviations(a, b, chi):
    """Gauge-invariant deviations of move result `a` = (state, spectra, proj1, proj2) from `b`."""
    sa, sb = a[1], b[1]
    out = {"spectra": max(float((x.to(DEV) - y.to(DEV))[:chi].abs().max()) for x, y in zip(sa, sb)),
           "pi": max(pi_rel_diff(a[2][k], a[3][k], b[2][k], b[3][k]) for k in b[2])}
    worst = 0.0
    for y in range(2):
        for ta, tb in ((a[0][(1, y)].C[3], b[0][(1, y)].C[3]), (a[0][(1, y)].C[0], b[0][(1, y)].C[0]), (a[0][(1, y)].E[3], b[0][(1, y)].E[3])):
            assert ta.shape == tb.shape
            va, vb = sv(ta, ta.shape[0]), sv(tb, tb.shape[0])
            worst = max(worst, float((va - vb).abs().max() / vb[0]))
    out["absorbed_sv"] = worst
    r0, r1 = orc.site_rdm(b[0], (1, 0)), orc.site_rdm(a[0], (1, 0))
    out["site_rdm"] = float((r0 / r0.trace() - r1 / r1.trace()).abs().max())
    return out


@pytest.mark.parametrize("D,chi,d,engine", [(8, 256, 2, "i8"), (8, 256, 2, "dmma"), (6, 144, 2, "i8"), (6, 144, 2, "dmma"),
                                            (7, 196, 4, "i8"), (7, 196, 4, "dmma")])
def test_headline_size_left_move_vs_oracle_on_gpu(D, chi, d, engine):
    """Tolerances: truncated spectra 1e-10 of s0 (north star), unconditionally.  The other three quantities involve the projectors
    themselves, i.e. singular VECTORS scaled by s^-1/2: on synthetic random tensors the spectrum is dense at the cut (gap
    s_chi - s_chi+1 ~ 1e-5 s0, s_chi ~ 1e-3 s0), so rounding-level differences are amplified by ~ s0 / (gap sqrt(s_chi)); how much
    is MEASURED on the reference itself: the oracle re-run with a mathematically equivalent QR basis (tests/util.rotated_qr).  They
    must stay within max(1e-9, 3 x that envelope)."""
    from tests.util import rotated_qr
    cell = cell_to(orc.random_cell(2, 2, D, chi, d, seed=0), DEV)
    torch.manual_seed(17)
    ref, got, tape = run_left_move(cell, engine)
    if engine == "i8":
        assert ops.i8_supported(chi * D * D, chi * D * D, chi + 2)
    assert len(ref[1]) == len(got[1]) == 2
    for key in ref[2]:
        assert got[2][key].shape == ref[2][key].shape and got[3][key].shape == ref[3][key].shape
    dev = deviations(got, ref, chi)
    with rotated_qr(seed=2):
        ref_rot = oracle_left_move(cell, orc.OmegaTape(tape))
    env = deviations(ref_rot, ref, chi)
    print(f"D={D} chi={chi} d={d} {engine}: b200-vs-oracle {dev}   oracle-vs-oracle(rotated QR) {env}")
    assert dev["spectra"] < 1e-10
    for k in ("pi", "absorbed_sv", "site_rdm"):
        assert dev[k] <= max(1e-9, 3.0 * env[k]), (k, dev[k], env[k])


def graded_cell(D, chi, d, decades, seed):
    """Random cell whose chi legs are graded: E[k][a,b,:,:] *= g[a] g[b], C[k][a,b] *= g[a] g[b], g = 10^-(decades * i / chi).  The
    quarter tensors then span 2 * decades orders of magnitude between their largest and smallest rows / columns, the regime of
    converged physical boundaries (the matrix handed to the QR of the reference's converged Ising state has cond 7e20)."""
    cell = orc.random_cell(2, 2, D, chi, d, seed=seed)
    g = torch.logspace(0, -decades, chi, dtype=torch.float64)
    for s in cell.site_list:
        cell[s].C = [c * g[:, None] * g[None, :] for c in cell[s].C]
        cell[s].E = [e * g[:, None, None, None] * g[None, :, None, None] for e in cell[s].E]
    return cell


@pytest.mark.parametrize("D,chi", [(8, 64), (6, 144)])
def test_k7_on_graded_boundary_vs_k1_and_oracle(D, chi):
    """m = chi D^2 >= 4096 (K7 active under 'auto'), boundary graded over 10 decades: K7, K1 and torch must agree on the
    truncated spectra (1e-10 of s0), on the truncated rank, and -- within the reference's own envelope -- on Pi and the
    absorbed tensors."""
    from tests.util import rotated_qr
    cell = cell_to(graded_cell(D, chi, 2, 10.0, seed=3), DEV)
    assert chi * D * D >= ProjectorCalculator.I8_MIN_DIM
    torch.manual_seed(23)
    tape = DeviceTape()
    ref = oracle_left_move(cell, tape)
    with rotated_qr(seed=2):
        env = deviations(oracle_left_move(cell, orc.OmegaTape(tape.tape)), ref, chi)
    for engine in ("i8", "dmma"):
        got = b200_left_move(cell, engine, tape.tape)
        for key in ref[2]:
            assert got[2][key].shape == ref[2][key].shape, engine      # same truncated rank chi'
        dev = deviations(got, ref, chi)
        print(f"graded D={D} chi={chi} {engine}: b200-vs-oracle {dev}   oracle-vs-oracle(rotated QR) {env}")
        assert dev["spectra"] < 1e-10, engine
        for k in ("pi", "absorbed_sv", "site_rdm"):
            assert dev[k] <= max(1e-9, 3.0 * env[k]), (engine, k, dev[k], env[k])


@pytest.mark.parametrize("kind", ["spectrum", "rowcol"])
def test_k7_rsvd_on_ill_conditioned_factors(kind):
    """rSVD of A @ B at m = 4096, q = 130 with (spectrum) singular values of both factors graded 1 ... 1e-20 behind random
    orthogonal bases, or (rowcol) rows and columns scaled over 12 decades: K7 (encoded factors) and K1 against torch's
    cuBLAS/cuSOLVER evaluation of the same recipe with the same Omega: |ds|/s0 <= 1e-10 over the kept spectrum."""
    m, q = 4096, 130
    g = torch.Generator().manual_seed(9)

    def orth(n):
        return torch.linalg.qr(torch.randn(n, n, dtype=torch.float64, generator=g).to(DEV)).Q
    if kind == "spectrum":
        s = torch.logspace(0, -20, m, dtype=torch.float64, device=DEV)
        A = (orth(m) * s) @ orth(m).T
        B = (orth(m) * s) @ orth(m).T
    else:
        r = torch.logspace(0, -12, m, dtype=torch.float64, device=DEV)
        A = torch.randn(m, m, dtype=torch.float64, generator=g).to(DEV) * r[:, None] * r.flip(0)[None, :]
        B = torch.randn(m, m, dtype=torch.float64, generator=g).to(DEV) * r.flip(0)[:, None] * r[None, :]
    omega = torch.randn(m, q, dtype=torch.float64, generator=g).to(DEV)
    _, S_t, _ = orc.fused_matmul_svd_lowrank(A, B, q=q, niter=2, omega_fn=lambda n, qq, dt=None, dv=None: omega)
    _, S_k1, _, _ = ops.rsvd([A, B], omega, niter=2, chi=q - 2)
    encs = [ops.i8_encode(A), ops.i8_encode(B)]
    _, S_k7, _, _ = ops.rsvd([A, B], omega, niter=2, chi=q - 2, encs=encs)
    e1 = float(((S_k1 - S_t) / S_t[0]).abs().max())
    e7 = float(((S_k7 - S_t) / S_t[0]).abs().max())
    print(f"ill-conditioned rSVD ({kind}): |ds|/s0  K1 {e1:.2e}  K7 {e7:.2e}")
    assert e1 < 1e-10 and e7 < 1e-10


def test_k7_forced_on_reference_heisenberg_state(golden_dir, monkeypatch):
    """The reference's converged Heisenberg D=3 chi=16 state (quarter tensors 144 x 144; a degenerate multiplet straddles the
    cut): every move of one sweep, synchronized with the CPU oracle, with the thin products forced onto K7."""
    st = load_golden("gs_heisenberg_D3_chi16.pt")
    cell = cell_from_plain(st)
    chi = cell.dims["chi"]
    monkeypatch.setattr(ProjectorCalculator, "I8_MIN_DIM", 128)
    cfg = orc.CtmrgConfig()
    ip = Ipeps.from_plain(cell, CTMRGConfig(thin_engine="i8"))
    mover = DirectionalMover(ip.ctmrg_config)
    mover.projector_calculator.thin_engine = "i8"
    assert mover.projector_calculator._use_i8([(144, 144), (144, 144)], chi + 2)
    do = {0: mover.left_move, 1: mover.up_move, 2: mover.right_move, 3: mover.down_move}
    torch.manual_seed(31)
    from tests.test_gpu_ctmrg import energy, push_state, spectra_err, to_oracle_cell
    from tests.util import model_terms
    hb, hs, _ = model_terms(st["model"])
    for k, line in [(0, 0), (2, 1), (0, 1), (2, 0), (1, 1), (3, 0), (1, 0), (3, 1)]:
        push_state(cell, ip)
        tape, rec = orc.OmegaTape(), {}
        orc.directional_move(cell, k, line, cfg, tape, rec)
        mover.projector_calculator.spectra = []
        linalg.set_omega_source(orc.OmegaTape(tape.tape))
        try:
            do[k](ip, line)
        finally:
            linalg.set_omega_source(None)
        got = to_oracle_cell(ip)
        assert spectra_err(rec["spectra"], mover.projector_calculator.spectra, chi) < 1e-10
        e_ref, e_got = energy(cell, hb, hs), energy(got, hb, hs)
        assert abs(e_got - e_ref) <= 1e-9 * max(1.0, abs(e_ref))
