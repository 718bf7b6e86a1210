"""The reference's own integration / unit tests of the hot path, re-stated on the B200 path with the reference's OWN objects
(`acetn.ipeps.Ipeps`, `SiteTensor`, `Bond`, `FullUpdater`, `ALSSolver` from the unmodified package in oracle/_ref) on a CUDA device:

  tests/integration/test_projectors.py          -> acetn_b200.renormalization.ProjectorCalculator
  tests/integration/test_directional_mover.py   -> acetn_b200.renormalization.DirectionalMover
  tests/integration/test_ctmrg.py               -> acetn_b200.renormalization.ctmrg
  tests/integration/test_full_update.py         -> FullUpdater.tensor_update with backend="b200" (integration.install())
  tests/unit/test_rdm.py                        -> acetn_b200.measurement.RDM
  tests/unit/test_als.py                        -> ALSSolver.solve with backend="b200"

Same fixtures (dimensions, cells), same assertions; each test cites the lines it restates.  Value-level parity is the job of the oracle
tests (test_gpu_ctmrg.py, test_gpu_headline_parity.py, ...); these prove the mirrored classes take the reference's objects and honour
the reference's interface contracts (shapes, updated-in-place semantics, error behaviour)."""
import pytest
import torch

from oracle import vendor_ref
from tests.dropin_util import setup

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(vendor_ref.import_path() is None, reason="reference package not available (oracle/_ref)")]

SITES = [(0, 0), (1, 0), (1, 1), (0, 1)]


def _ipeps(bond, chi, projectors=None, steps=None, model=False):
    Ipeps = setup()
    cfg = {"dtype": torch.float64, "device": torch.device("cuda"),
           "TN": {"nx": 2, "ny": 2, "dims": {"phys": 2, "bond": bond, "chi": chi}},
           "ctmrg": {"disable_progressbar": True}, "evolution": {"backend": "b200", "disable_progressbar": True}}
    if projectors is not None:
        cfg["ctmrg"]["projectors"] = projectors
    if steps is not None:
        cfg["ctmrg"]["steps"] = steps
    if model:
        cfg["model"] = {"name": "heisenberg", "params": {"J": 1.0}}
    torch.manual_seed(0)
    return Ipeps(cfg)


# ------------------------------------------------------------------------------------------------ test_projectors.py
@pytest.fixture
def ipeps_and_projector_calculator():                       # test_projectors.py:6-27
    from acetn_b200.renormalization import ProjectorCalculator
    ipeps = _ipeps(6, 36, projectors="half-system")
    return ipeps, ProjectorCalculator(ipeps.config.ctmrg)


def test_projector_calculator_calculate(ipeps_and_projector_calculator):          # :37-40
    ipeps, pc = ipeps_and_projector_calculator
    for k in range(4):
        pc.calculate(ipeps, SITES, k)


def test_projector_calculator_different_projectors(ipeps_and_projector_calculator):   # :42-50
    from acetn_b200.renormalization import ProjectorCalculator
    for proj_type in ["half-system", "full-system"]:
        ipeps, _ = ipeps_and_projector_calculator
        config = ipeps.config.ctmrg
        config.projectors = proj_type
        pc = ProjectorCalculator(config)
        for k in range(4):
            pc.calculate(ipeps, SITES, k)


def test_projector_calculator_output_structure(ipeps_and_projector_calculator):   # :52-61
    ipeps, pc = ipeps_and_projector_calculator
    projectors = pc.calculate(ipeps, SITES, k=0)
    bD, cD = ipeps.dims['bond'], ipeps.dims['chi']
    assert len(projectors) == 2
    assert all(isinstance(p, torch.Tensor) and p.is_cuda for p in projectors)
    assert all(p.shape == (cD, bD, bD, cD) for p in projectors)


def test_projector_calculator_invalid_sites(ipeps_and_projector_calculator):      # :63-67
    ipeps, pc = ipeps_and_projector_calculator
    with pytest.raises(ValueError):
        pc.calculate(ipeps, [(100, 100), (200, 200), (300, 300), (400, 400)], k=0)


def test_projector_calculator_invalid_type(ipeps_and_projector_calculator):       # projectors.py:33
    from acetn_b200.renormalization import ProjectorCalculator
    ipeps, _ = ipeps_and_projector_calculator
    config = ipeps.config.ctmrg
    config.projectors = "quarter-system"
    with pytest.raises(ValueError):
        ProjectorCalculator(config)
    config.projectors = "half-system"


# ------------------------------------------------------------------------------------------------ test_directional_mover.py
@pytest.fixture
def ipeps_and_directional_mover():                          # test_directional_mover.py:7-28
    from acetn_b200.renormalization import DirectionalMover
    ipeps = _ipeps(6, 36, projectors="half-system")
    return ipeps, DirectionalMover(ipeps.config.ctmrg)


def test_directional_mover_initialization(ipeps_and_directional_mover):           # :34-37
    _, mover = ipeps_and_directional_mover
    assert callable(mover.calculate_projectors)


@pytest.mark.parametrize('move_func', ['left_move', 'up_move', 'right_move', 'down_move'])
def test_move_methods(ipeps_and_directional_mover, move_func):                     # :39-44
    ipeps, mover = ipeps_and_directional_mover
    before = {(s, k): ipeps[s]['C'][k] for s in SITES for k in range(4)}
    getattr(mover, move_func)(ipeps, 0)
    torch.cuda.synchronize()
    changed = sum(1 for (s, k), t in before.items() if ipeps[s]['C'][k] is not t)
    assert changed == 4, "a move replaces two corners on each of the two sites of the target line"


@pytest.mark.parametrize('dir_func, site_idx', [('calculate_left_projectors', 0), ('calculate_right_projectors', 1),
                                                ('calculate_up_projectors', 2), ('calculate_down_projectors', 3)])
def test_projectors(ipeps_and_directional_mover, dir_func, site_idx):              # :56-67
    ipeps, mover = ipeps_and_directional_mover
    xi, yi = SITES[site_idx]
    proj1, proj2 = getattr(mover, dir_func)(ipeps, xi, yi)
    assert isinstance(proj1, torch.Tensor) and isinstance(proj2, torch.Tensor)
    assert proj1.shape == proj2.shape == (36, 6, 6, 36)


def test_renormalize_boundary(ipeps_and_directional_mover):                        # :69-76
    ipeps, mover = ipeps_and_directional_mover
    proj1, proj2 = {}, {}
    for yi in range(2):
        proj1[yi], proj2[yi] = mover.calculate_left_projectors(ipeps, SITES[yi][0], SITES[yi][1])
    s1, s2 = (0, 0), (1, 0)
    old = ipeps[s2]['C'][0]
    mover.renormalize_boundary(ipeps, proj1, proj2, s1, s2, 0, 1, 0)
    assert isinstance(ipeps[s2]['C'][0], torch.Tensor) and ipeps[s2]['C'][0] is not old
    assert ipeps[s2]['C'][0].shape == (36, 36) and ipeps[s2]['E'][3].shape == (36, 36, 6, 6)


# ------------------------------------------------------------------------------------------------ test_ctmrg.py
def test_ctmrg():                                            # test_ctmrg.py:31-46
    from acetn_b200.renormalization import ctmrg
    ipeps = _ipeps(3, 9, steps=3)
    config = ipeps.config.ctmrg
    for projectors in ['half-system', 'full-system']:
        config.projectors = projectors
        before = {site + (k,): ipeps[site]['C'][k] for site in SITES for k in range(4)}
        ctmrg(ipeps, config)
        for site in SITES:
            for k in range(4):
                if ipeps[site]['C'][k].shape == before[site + (k,)].shape:
                    assert not (ipeps[site]['C'][k] == before[site + (k,)]).all()


# ------------------------------------------------------------------------------------------------ test_full_update.py
def test_tensor_update():                                    # test_full_update.py:39-70
    ipeps = _ipeps(5, 8, model=True)
    from acetn.evolution.full_update import FullUpdater
    from acetn.evolution.gate import Gate
    from acetn.model.model_factory import model_factory
    model = model_factory.create(ipeps.config.model)
    gate = Gate(model, 0.01, ipeps.bond_list, ipeps.site_list)

    class Config:
        backend = "b200"
        als_niter = 100
        als_tol = 1e-15
        als_method = "cholesky"
        als_epsilon = 1e-12
        use_gauge_fix = False
        gauge_fix_atol = 1e-12
        positive_approx_cutoff = 1e-12

    full_updater = FullUpdater(ipeps, gate, Config())
    torch.manual_seed(1)
    a1 = torch.rand(5, 5, 5, 5, 2, dtype=ipeps.dtype).cuda()
    a2 = torch.rand(5, 5, 5, 5, 2, dtype=ipeps.dtype).cuda()
    bond = ipeps.bond_list[0]
    u1, u2 = full_updater.tensor_update(a1, a2, bond)
    assert u1.shape == a1.shape and u2.shape == a2.shape
    one = torch.tensor(1.0, dtype=ipeps.dtype, device="cuda")
    assert torch.isclose(u1.norm(), one, atol=1e-6) and torch.isclose(u2.norm(), one, atol=1e-6)
    v1, v2 = full_updater.tensor_update(a1, a2, bond)
    assert torch.norm(u1 - v1) < 1e-6 and torch.norm(u2 - v2) < 1e-6


# ------------------------------------------------------------------------------------------------ test_rdm.py
class _MockIPEPS:                                            # test_rdm.py:7-15
    def __init__(self):
        self.site_tensors = {}

    def __getitem__(self, key):
        return self.site_tensors.get(key, {})

    def __setitem__(self, key, value):
        self.site_tensors[key] = value


@pytest.fixture
def rdm_object():                                            # :17-26
    setup()
    from acetn.ipeps.site_tensor import SiteTensor
    from acetn_b200.measurement import RDM
    ipeps = _MockIPEPS()
    dev = torch.device("cuda")
    ipeps[(0, 0)] = SiteTensor({'phys': 3, "bond": 4, "chi": 5}, device=dev)
    ipeps[(1, 0)] = SiteTensor({'phys': 3, "bond": 4, "chi": 5}, device=dev)
    return RDM(ipeps)


def test_build_site_rdm(rdm_object):                         # :29-38
    rdm = rdm_object[(0, 0)]
    assert isinstance(rdm, torch.Tensor) and rdm.shape == (3, 3)


def test_build_bond_rdm(rdm_object):                         # :41-53
    from acetn.ipeps.bond import Bond
    rdm = rdm_object[Bond((0, 0), (1, 0), 0)]
    assert isinstance(rdm, torch.Tensor) and rdm.shape == (3, 3, 3, 3)


def test_invalid_index_access(rdm_object):                   # :56-59
    with pytest.raises(KeyError):
        rdm_object[(5, 5)]


# ------------------------------------------------------------------------------------------------ test_als.py
@pytest.fixture
def setup_als_solver():                                      # test_als.py:8-35
    setup()
    from acetn.evolution.als_solver import ALSSolver
    from acetn.evolution.full_update import gauge_fix, positive_approx
    from acetn.ipeps.ipeps_config import EvolutionConfig
    dev = torch.device("cuda")
    nD, bD, pD = 8, 5, 3
    torch.manual_seed(4)
    a1r = torch.rand(nD, bD, pD, dtype=torch.float64, device=dev)
    a2r = torch.rand(nD, bD, pD, dtype=torch.float64, device=dev)
    a12g = torch.einsum("yup,xuq->yxpq", a1r, a2r)
    a12g += 1e-4 * torch.rand_like(a12g)
    n12 = torch.rand(nD, nD, nD, nD, dtype=torch.float64, device=dev)
    nz = positive_approx(n12, nD)
    n12, a12g, *_ = gauge_fix(nz, a12g, nD)
    config = EvolutionConfig(als_niter=20, als_tol=1e-15, backend="b200")
    return ALSSolver(n12, a12g, (nD, bD, pD), config)


def test_als_solver_solution(setup_als_solver):              # :43-50
    a1r, a2r = setup_als_solver.solve()
    assert a1r.shape == (8, 5, 3) and a2r.shape == (8, 5, 3) and a1r.is_cuda


def test_als_solver_convergence(setup_als_solver):           # :52-63
    s = setup_als_solver
    s.niter = 99
    a1r, a2r = s.solve()
    prev = s.calculate_cost(a1r, a2r, s.a12g, s.n12)
    s.niter = 100
    a1r, a2r = s.solve()
    nxt = s.calculate_cost(a1r, a2r, s.a12g, s.n12)
    assert abs(nxt - prev) < max(s.tol, 1e-13 * abs(float(prev))), "Convergence failed"
