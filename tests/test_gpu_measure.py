"""GPU parity tests of the measure path (site / bond RDMs, energy) against the CPU oracle and the reference's
known-answer energies (tests/integration/ipeps_gs/energies.csv, rel 1e-10)."""
import pytest
import torch

from oracle import ctmrg_oracle as orc
from tests.util import cell_from_plain, load_golden, model_terms

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from acetn_b200 import ops
    from acetn_b200.ipeps import Ipeps
    from acetn_b200.measurement import RDM, measure


def rel(a, b):
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("spec,sa,sb", [("ab,bcuU->acuU", (5, 6), (6, 7, 3, 3)), ("acuU,ealL->cuUelL", (5, 7, 3, 3), (4, 5, 3, 3)),
                                        ("cuUelL,LURDP->cuelRDP", (4, 3, 3, 5, 3, 3), (3, 3, 3, 3, 2)),
                                        ("lurdp,cuelRDp->crRedD", (3, 3, 3, 3, 2), (4, 3, 5, 3, 3, 3, 2)),
                                        ("fcRrp,fcRrq->pq", (4, 5, 3, 3, 2), (4, 5, 3, 3, 2)), ("ab,ab->", (6, 7), (6, 7))])
def test_contract_matches_einsum(spec, sa, sb):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(*sa, dtype=torch.float64, generator=g).cuda()
    B = torch.randn(*sb, dtype=torch.float64, generator=g).cuda()
    C = ops.contract(spec, A, B)
    ref = torch.einsum(spec, A, B)
    assert C.shape == ref.shape
    assert rel(C, ref) < 1e-13
    # strided views as operands
    At = A.transpose(0, 1).contiguous().transpose(0, 1)
    assert rel(ops.contract(spec, At, B), ref) < 1e-13


@pytest.mark.parametrize("D,chi,d,seed", [(2, 8, 2, 0), (3, 10, 2, 1), (3, 7, 3, 2), (4, 12, 2, 3)])
def test_rdms_match_oracle(D, chi, d, seed):
    cell = orc.random_cell(2, 2, D, chi, d, seed=seed)
    ip = Ipeps.from_plain(cell)
    rdm = RDM(ip)
    assert rel(rdm[(0, 0)].cpu(), orc.site_rdm(cell, (0, 0))) < 1e-12
    assert rel(rdm[(1, 0)].cpu(), orc.site_rdm(cell, (1, 0))) < 1e-12
    for bond in (cell.bond_list[0], cell.bond_list[3], cell.bond_list[-1]):
        got = rdm[bond].cpu()
        assert got.shape == (d, d, d, d)
        assert rel(got, orc.bond_rdm(cell, bond)) < 1e-12


@pytest.mark.parametrize("name", ["gs_ising_D2_chi20.pt", "gs_heisenberg_D3_chi16.pt"])
def test_known_answer_energy_on_gpu(name):
    """reference tests/integration/test_ground_states.py:30 through the B200 measure path."""
    st = load_golden(name)
    cell = cell_from_plain(st)
    hb, hs, ops_ = model_terms(st["model"])
    out = measure(Ipeps.from_plain(cell), hb, hs, ops_)
    assert float(out["Energy"]) == pytest.approx(st["energy_csv"], rel=1e-10)
    for k, v in st["reference_measure"].items():
        assert float(out[k]) == pytest.approx(v, rel=1e-10, abs=1e-12)


def test_rdms_with_unequal_chi_legs():
    """Every chi leg with its own extent through acetn_b200_site_rdm / acetn_b200_bond_rdm (truncation gives chi' per projector,
    SURVEY.md App. D2); the same ragged rings as the host-logic CPU test (tests/test_environment_host_cpu.py), on the real kernels."""
    from tests.test_environment_host_cpu import _ragged_sites
    D, d = 3, 2
    A, B = _ragged_sites(D, d, 7)
    cell = orc.Cell(2, 1, {}, {(0, 0): A, (1, 0): B})

    class S:
        def __init__(self, s):
            self.s = s

        def __getitem__(self, k):
            return {"A": self.s.A.cuda(), "C": [c.cuda() for c in self.s.C], "E": [e.cuda() for e in self.s.E]}[k]

        def bond_permute(self, k):
            return self.s.bond_permute(k).cuda()

    got = ops.bond_rdm(S(A), S(B), 0).cpu()
    assert rel(got, orc.bond_rdm(cell, ((0, 0), (1, 0), 0))) < 1e-12
    gg = torch.Generator().manual_seed(9)
    R = lambda *s: torch.randn(*s, dtype=torch.float64, generator=gg)     # noqa: E731
    a, b, c, e, g_, h, i, j = 2, 3, 4, 5, 6, 3, 5, 4
    St = orc.Site(R(D, D, D, D, d), [R(c, g_), R(h, i), R(j, e), R(a, b)], [R(g_, h, D, D), R(i, j, D, D), R(e, a, D, D), R(b, c, D, D)])
    ref = orc.site_rdm(orc.Cell(1, 1, {}, {(0, 0): St}), (0, 0))
    got = ops.site_rdm([t.cuda() for t in St.C], [t.cuda() for t in St.E], St.A.cuda()).cpu()
    assert rel(got, ref) < 1e-12
    with pytest.raises(RuntimeError):                 # mismatched legs are rejected with the library's message
        ops.site_rdm([t.cuda() for t in St.C][::-1], [t.cuda() for t in St.E], St.A.cuda())
