"""GPU parity tests of the CTMRG path proper (projectors, directional moves, whole sweeps) on the B200 backend
against the CPU oracle, on identical inputs and identical Omega draws, compared on gauge-invariant quantities:
truncated projector spectra |ds|/s0 <= 1e-10, energy per site <= 1e-9 (north-star tolerances).

Three regimes (DESIGN.md "Parity and the reproducibility envelope"):
  * synchronized moves  -- both implementations start every directional move from the same state: tolerances hold
    for every input, including synthetic random tensors;
  * free-running sweeps -- tolerances hold on well-conditioned inputs (product-state start, the reference's converged
    Ising state with the reference's own Omega tape);
  * free-running sweeps on inputs where CTMRG+rSVD is itself chaotic (random tensors: ungapped truncation) or cuts a
    degenerate multiplet (the reference's Heisenberg state): the reference algorithm does not reproduce ITSELF under a
    mathematically equivalent change of its QR basis (tests/util.rotated_qr); there the deviation from the oracle is
    required to stay within that envelope.
"""
import pytest
import torch

from oracle import ctmrg_oracle as orc
from tests.util import cell_from_plain, load_golden, model_terms, rotated_qr

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from acetn_b200 import linalg, ops
    from acetn_b200.ipeps import CTMRGConfig, Ipeps, SiteTensor
    from acetn_b200.renormalization import DirectionalMover, ProjectorCalculator, ctmrg

H = orc.heisenberg_bond_hamiltonian(1.0)


def neel(s):
    return [1.0, 0.0] if (s[0] + s[1]) % 2 == 0 else [0.0, 1.0]


def to_oracle_cell(ip):
    sites = {s: orc.Site(ip[s]['A'].cpu(), [c.cpu() for c in ip[s]['C']], [e.cpu() for e in ip[s]['E']]) for s in ip.site_list}
    return orc.Cell(ip.nx, ip.ny, ip.dims, sites)


def push_state(cell, ip):
    for s in cell.site_list:
        ip[s] = SiteTensor(cell[s].A.clone(), [c.clone() for c in cell[s].C], [e.clone() for e in cell[s].E]).to(ip.device)


def run_b200(cell, nsweep, tape, projectors="half-system"):
    ip = Ipeps.from_plain(cell, CTMRGConfig(steps=nsweep, projectors=projectors))
    replay = orc.OmegaTape(tape)
    linalg.set_omega_source(replay)
    try:
        mover = DirectionalMover(ip.ctmrg_config)
        mover.projector_calculator.spectra = []
        ctmrg(ip, ip.ctmrg_config, mover)
    finally:
        linalg.set_omega_source(None)
    assert replay.pos == len(tape)
    return ip, mover.projector_calculator.spectra


def run_oracle(cell, nsweep, tape=None, projectors="half-system"):
    ref = cell.clone()
    t = orc.OmegaTape(tape) if tape is not None else orc.OmegaTape()
    rec = {}
    orc.ctmrg(ref, orc.CtmrgConfig(steps=nsweep, projectors=projectors), omega_fn=t, record=rec)
    return ref, rec["spectra"], t.tape


def bond_ham_for(d):
    if d == 2:
        return H
    g = torch.Generator().manual_seed(99)
    m = torch.randn(d * d, d * d, dtype=torch.float64, generator=g)
    return 0.5 * (m + m.T)


def energy(cell, hb=None, hs=None):
    hb = bond_ham_for(cell.dims["phys"]) if hb is None else hb
    return float(orc.measure(cell, hb, hs)["Energy"])


def spectra_err(sa, sb, chi):
    return max(float((a - b)[:chi].abs().max()) for a, b in zip(sa, sb))


def assert_same_shapes(ref, got):
    for s in ref.site_list:
        for k in range(4):
            assert got[s].C[k].shape == ref[s].C[k].shape
            assert got[s].E[k].shape == ref[s].E[k].shape


# ------------------------------------------------------------------------------------------------ synchronized moves
SYNC_CASES = [("random", 2, 8, 2, 0, 2, 2), ("random", 3, 12, 2, 1, 2, 2), ("random", 2, 6, 2, 3, 3, 2), ("random", 4, 16, 2, 5, 2, 2),
              ("random", 3, 9, 3, 6, 2, 2), ("product", 3, 18, 2, 7, 2, 2),
              ("random", 4, 64, 2, 8, 2, 2),      # BASELINE config 2 shape (TFIM D=4 chi=64)
              ("random", 2, 20, 2, 9, 2, 2)]      # BASELINE config 1 shape (README quickstart D=2 chi=20)


@pytest.mark.parametrize("kind,D,chi,d,seed,nx,ny", SYNC_CASES)
def test_synchronized_moves(kind, D, chi, d, seed, nx, ny):
    """Every directional move of one sweep (ctmrg.py:26-31 order), both sides restarted from the oracle's state."""
    cell = orc.random_cell(nx, ny, D, chi, d, seed=seed) if kind == "random" else orc.product_cell(nx, ny, D, chi, d, seed=seed, state_map=neel)
    torch.manual_seed(100 + seed)
    cfg = orc.CtmrgConfig()
    ip = Ipeps.from_plain(cell, CTMRGConfig())
    mover = DirectionalMover(ip.ctmrg_config)
    moves = []
    for xi in range(nx):
        moves += [(0, xi), (2, (nx - xi + 1) % nx)]
    for yi in range(ny):
        moves += [(1, (ny - yi + 1) % ny), (3, yi)]
    do = {0: mover.left_move, 1: mover.up_move, 2: mover.right_move, 3: mover.down_move}
    for k, line in moves:
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
        assert_same_shapes(cell, got)
        assert spectra_err(rec["spectra"], mover.projector_calculator.spectra, chi) < 1e-10
        e_ref, e_got = energy(cell), energy(got)
        assert abs(e_got - e_ref) <= 1e-9 * max(1.0, abs(e_ref))
        r0, r1 = orc.site_rdm(cell, (0, 0)), orc.site_rdm(got, (0, 0))
        assert float((r0 / r0.trace() - r1 / r1.trace()).abs().max()) < 1e-9


# ------------------------------------------------------------------------------------------------ free-running sweeps
@pytest.mark.parametrize("D,chi,nsweep,seed", [(2, 10, 3, 2), (3, 18, 3, 3), (2, 20, 4, 4)])
def test_free_running_product_state(D, chi, nsweep, seed):
    """Default product-state start: boundary rank << chi, chi legs grow and differ."""
    cell = orc.product_cell(2, 2, D, chi, 2, seed=seed, state_map=neel)
    torch.manual_seed(7)
    ref, s_ref, tape = run_oracle(cell, nsweep)
    ip, s_got = run_b200(cell, nsweep, tape)
    got = to_oracle_cell(ip)
    assert_same_shapes(ref, got)
    assert spectra_err(s_ref, s_got, chi) < 1e-10
    assert abs(energy(got) - energy(ref)) < 1e-9
    for s in ref.site_list:
        for k in range(4):
            a, b = torch.linalg.svdvals(got[s].C[k]), torch.linalg.svdvals(ref[s].C[k])
            assert float((a / a[0] - b / b[0]).abs().max()) < 1e-9


def test_free_running_golden_ising():
    """The reference's converged TFIM state + the reference's own Omega tape: after two more sweeps the energy must
    equal the value the reference itself produced (tests/golden/make_golden.py) to 1e-9 and the known answer of
    tests/integration/ipeps_gs/energies.csv."""
    st = load_golden("gs_ising_D2_chi20.pt")
    cell = cell_from_plain(st)
    hb, hs, _ = model_terms(st["model"])
    ip, _ = run_b200(cell, 2, st["omega_tape_2sweeps"])
    e = energy(to_oracle_cell(ip), hb, hs)
    assert e == pytest.approx(st["after2_energy"], rel=1e-9)
    assert e == pytest.approx(st["energy_csv"], rel=1e-9)
    for key, refsv in st["after2_corner_svals"].items():
        x, y, k = (int(t) for t in key.split(","))
        s = torch.linalg.svdvals(ip[(x, y)]['C'][k].cpu())
        assert float((s / s[0] - refsv / refsv[0]).abs().max()) < 1e-8


ENVELOPE_CASES = ["random_D2_chi8", "random_D3_chi12", "random_D2_chi8_full", "golden_heisenberg"]


@pytest.mark.parametrize("case", ENVELOPE_CASES)
def test_free_running_within_reference_envelope(case):
    """Chaotic / degenerate inputs: deviation from the oracle must not exceed the oracle's own deviation under an
    equivalent QR basis (x20 margin; both are rounding-seeded chaotic quantities) -- and a loose absolute sanity bound."""
    projectors, hb, hs, tape = "half-system", H, None, None
    if case == "golden_heisenberg":
        st = load_golden("gs_heisenberg_D3_chi16.pt")
        cell = cell_from_plain(st)
        hb, hs, _ = model_terms(st["model"])
        tape = st["omega_tape_2sweeps"]
    elif case == "random_D2_chi8":
        cell = orc.random_cell(2, 2, 2, 8, 2, seed=0)
    elif case == "random_D3_chi12":
        cell = orc.random_cell(2, 2, 3, 12, 2, seed=1)
    else:
        cell, projectors = orc.random_cell(2, 2, 2, 8, 2, seed=4), "full-system"
    torch.manual_seed(5)
    ref, s_ref, tape = run_oracle(cell, 2, tape, projectors)
    env_e, env_s = 0.0, 0.0
    for rs in (1, 2):
        with rotated_qr(rs):
            alt, s_alt, _ = run_oracle(cell, 2, tape, projectors)
        env_e = max(env_e, abs(energy(alt, hb, hs) - energy(ref, hb, hs)))
        env_s = max(env_s, spectra_err(s_ref, s_alt, cell.dims["chi"]))
    ip, s_got = run_b200(cell, 2, tape, projectors)
    got = to_oracle_cell(ip)
    assert_same_shapes(ref, got)
    d_e = abs(energy(got, hb, hs) - energy(ref, hb, hs))
    d_s = spectra_err(s_ref, s_got, cell.dims["chi"])
    assert d_e <= 20 * env_e + 1e-9, (d_e, env_e)
    assert d_s <= 20 * env_s + 1e-10, (d_s, env_s)
    if case == "golden_heisenberg":
        # reference test tolerance after further evolution is rel 1e-4 (tests/integration/test_ground_states.py:33-35)
        assert energy(got, hb, hs) == pytest.approx(st["energy_csv"], rel=1e-6)
    # the first move starts from identical states and must agree to the tight tolerance
    ny = cell.ny
    assert spectra_err(s_ref[:ny], s_got[:ny], cell.dims["chi"]) < 1e-10


@pytest.mark.parametrize("D,chi,d", [(6, 36, 2), (8, 24, 2), (7, 14, 4)])
def test_single_move_larger_bond_dimension(D, chi, d):
    """One synchronized left move at the bond dimensions of BASELINE configs 3-5 (D=6, 8, 7 with d=4) at a chi the CPU
    oracle finishes in seconds: exercises the fused K2 (D=8), odd-D scalar loaders (D=7) and d>2."""
    cell = orc.random_cell(2, 2, D, chi, d, seed=11)
    torch.manual_seed(3)
    ip = Ipeps.from_plain(cell, CTMRGConfig())
    mover = DirectionalMover(ip.ctmrg_config)
    mover.projector_calculator.spectra = []
    tape, rec = orc.OmegaTape(), {}
    orc.directional_move(cell, 0, 0, orc.CtmrgConfig(), tape, rec)
    linalg.set_omega_source(orc.OmegaTape(tape.tape))
    try:
        mover.left_move(ip, 0)
    finally:
        linalg.set_omega_source(None)
    got = to_oracle_cell(ip)
    assert_same_shapes(cell, got)
    assert spectra_err(rec["spectra"], mover.projector_calculator.spectra, chi) < 1e-10
    for y in range(2):
        for k in (0, 3):
            a, b = torch.linalg.svdvals(got[(1, y)].C[k]), torch.linalg.svdvals(cell[(1, y)].C[k])
            assert float((a / a[0] - b / b[0]).abs().max()) < 1e-9
    r0, r1 = orc.site_rdm(cell, (1, 0)), orc.site_rdm(got, (1, 0))
    assert float((r0 / r0.trace() - r1 / r1.trace()).abs().max()) < 1e-9


@pytest.mark.parametrize("projectors", ["half-system", "full-system"])
def test_full_rank_svd_type(projectors):
    """svd_type='full-rank' (projectors.py:114-136): deterministic (no Omega); one sweep against the oracle."""
    cell = orc.random_cell(2, 2, 2, 8, 2, seed=12)
    ocfg = orc.CtmrgConfig(steps=1, projectors=projectors, svd_type="full-rank")
    ref = cell.clone()
    rec = {}
    orc.directional_move(ref, 0, 0, ocfg, record=rec)
    ip = Ipeps.from_plain(cell, CTMRGConfig(steps=1, projectors=projectors, svd_type="full-rank"))
    mover = DirectionalMover(ip.ctmrg_config)
    mover.projector_calculator.spectra = []
    mover.left_move(ip, 0)
    got = to_oracle_cell(ip)
    assert_same_shapes(ref, got)
    for a, b in zip(rec["spectra"], mover.projector_calculator.spectra):
        assert float((a - b)[:8].abs().max()) < 1e-10
    assert abs(energy(got) - energy(ref)) < 1e-9
    U, S, V = ProjectorCalculator.full_svd(torch.randn(60, 40, dtype=torch.float64, generator=torch.Generator().manual_seed(1)).cuda())
    assert U.shape == (60, 40) and V.shape == (40, 40)
    with pytest.raises(ValueError):
        ProjectorCalculator(CTMRGConfig(svd_type="bogus")).calculate(ip, [(0, 0), (1, 0), (1, 1), (0, 1)], 0)


def test_projector_pi_invariant():
    """Pi = proj2 proj1^T is gauge invariant; compare with the oracle on the same Omega."""
    cell = orc.random_cell(2, 2, 3, 12, 2, seed=1)
    cfg = orc.CtmrgConfig()
    tape = orc.OmegaTape()
    p1r, p2r = orc.half_system_projectors(cell, orc.plaquette(cell, 0, 0, 0), 0, cfg, omega_fn=tape)
    ip = Ipeps.from_plain(cell)
    linalg.set_omega_source(orc.OmegaTape(tape.tape))
    try:
        p1, p2 = ProjectorCalculator(ip.ctmrg_config).calculate(ip, orc.plaquette(cell, 0, 0, 0), 0)
    finally:
        linalg.set_omega_source(None)
    assert tuple(p1.shape) == tuple(p1r.shape)
    m = p1r.shape[0] * p1r.shape[1] * p1r.shape[2]
    pi_ref = p2r.reshape(m, -1) @ p1r.reshape(m, -1).T
    pi = (p2.reshape(m, -1) @ p1.reshape(m, -1).T).cpu()
    assert float((pi - pi_ref).norm() / pi_ref.norm()) < 1e-8


def test_invalid_projector_type_raises():
    with pytest.raises(ValueError):
        ProjectorCalculator(CTMRGConfig(projectors="bogus"))


def test_unknown_site_raises():
    """reference tests/integration/test_projectors.py:62-67 : ValueError on unknown sites."""
    ip = Ipeps.from_plain(orc.random_cell(2, 2, 2, 4, 2, seed=0))
    with pytest.raises(ValueError):
        ProjectorCalculator(ip.ctmrg_config).calculate(ip, [(5, 5), (1, 0), (1, 1), (0, 1)], 0)


def test_cpu_tensors_rejected():
    """No CPU fallback: host tensors must fail loudly."""
    from acetn_b200 import ops
    with pytest.raises(RuntimeError):
        ops.matmul(torch.zeros(4, 4, dtype=torch.float64), torch.zeros(4, 4, dtype=torch.float64))


@pytest.mark.parametrize("nx,ny,D,chi", [(2, 2, 4, 32), (3, 2, 3, 18)])
def test_staggered_schedule_is_bit_identical(nx, ny, D, chi):
    """The phase schedule (one low-priority bulk stream + one high-priority stream per rSVD chain, absorptions of a move
    as soon as its projectors exist) only reorders independent work: with the same Omega draws it must reproduce the
    lockstep schedule and the plain sequential moves (directional_mover.py:23-97 order) bit for bit."""
    cell = orc.random_cell(nx, ny, D, chi, 2, seed=11)
    outs = []
    for stagger, streams in ((True, 4), (False, 4), (False, 1)):
        ip = Ipeps.from_plain(cell, CTMRGConfig(steps=2))
        torch.manual_seed(5)                        # Omega: torch.randn on the device, drawn in task order by every schedule
        mover = DirectionalMover(ip.ctmrg_config, n_streams=streams)
        mover.stagger = stagger
        ctmrg(ip, ip.ctmrg_config, mover)
        torch.cuda.synchronize()
        outs.append(ip)
    for other in outs[1:]:
        for s in outs[0].site_list:
            for k in range(4):
                assert torch.equal(outs[0][s]['C'][k], other[s]['C'][k])
                assert torch.equal(outs[0][s]['E'][k], other[s]['E'][k])


@pytest.mark.parametrize("nx,ny,D,chi,paired", [(2, 2, 2, 8, True), (2, 2, 3, 12, True), (3, 2, 2, 6, False), (2, 2, 2, 20, False), (2, 2, 4, 16, True)])
def test_graph_replay_is_bit_identical(nx, ny, D, chi, paired, monkeypatch):
    """Launch-bound sizes run their phases as captured CUDA graphs once the boundary has saturated (MoveGraph: fixed-address arena,
    Omega drawn outside the graph in the reference's order, one host read per phase).  Same kernels in the same order: the tensors
    after several sweeps must equal the eager schedule's bit for bit, and graphs must actually have been replayed."""
    cell = orc.random_cell(nx, ny, D, chi, 2, seed=13)
    monkeypatch.setenv("ACETN_B200_PAIR_MOVES", "1" if paired else "0")
    out = {}
    for mode in ("0", "auto"):
        monkeypatch.setenv("ACETN_B200_GRAPHS", mode)
        torch.manual_seed(5)
        ip = Ipeps.from_plain(cell, CTMRGConfig(steps=6))
        mover = DirectionalMover(ip.ctmrg_config)
        ops.reset_launch_count()
        ctmrg(ip, ip.ctmrg_config, mover)
        torch.cuda.synchronize()
        out[mode] = (ip, mover.graph_replays, ops.launch_count())
    assert out["0"][1] == 0 and out["auto"][1] > 0
    # kernels inside replayed graphs count as launches of the library (bench.py's gpu_launches): same work, same count +- the warm-up
    assert 0.9 * out["0"][2] <= out["auto"][2] <= 1.6 * out["0"][2], (out["0"][2], out["auto"][2])
    a, b = out["0"][0], out["auto"][0]
    for s in a.site_list:
        for k in range(4):
            assert torch.equal(a[s]['C'][k], b[s]['C'][k]) and torch.equal(a[s]['E'][k], b[s]['E'][k])


def test_graph_replay_follows_replaced_site_tensors(monkeypatch):
    """`evolve` replaces the site tensor A after every bond update (fast_full_update.py:61-62): a captured phase must pick the new
    tensor up (arena copy-in) and give the eager result for it."""
    monkeypatch.setenv("ACETN_B200_PAIR_MOVES", "1")
    cell = orc.random_cell(2, 2, 2, 8, 2, seed=21)
    out = {}
    for mode in ("0", "auto"):
        monkeypatch.setenv("ACETN_B200_GRAPHS", mode)
        torch.manual_seed(9)
        ip = Ipeps.from_plain(cell, CTMRGConfig(steps=4))
        mover = DirectionalMover(ip.ctmrg_config)
        ctmrg(ip, ip.ctmrg_config, mover)
        g = torch.Generator().manual_seed(3)
        for s in ip.site_list:
            A = torch.rand(2, 2, 2, 2, 2, dtype=torch.float64, generator=g) - 0.5
            ip[s]['A'] = (A / A.norm()).cuda()
        for bond in ip.bond_list:
            mover.absorb_bond(ip, bond)
        torch.cuda.synchronize()
        out[mode] = (ip, mover.graph_replays)
    assert out["auto"][1] > out["0"][1] == 0
    for s in out["0"][0].site_list:
        for k in range(4):
            assert torch.equal(out["0"][0][s]['C'][k], out["auto"][0][s]['C'][k])
            assert torch.equal(out["0"][0][s]['E'][k], out["auto"][0][s]['E'][k])
