"""CPU tests of the host-side index logic of the gather-free environment contractions (acetn_b200/evolution.py env_front /
env_back, used by build_norm_tensor and the bond RDM): the two-level index descriptors handed to K1 through the C ABI are
interpreted here by a small torch emulation of acetn_b200_gemm's addressing (include/acetn_b200.h: offset(i) =
div ? (i / div) * s_hi + (i % div) * s_lo : i * s_lo), and the results are compared with the oracle's einsum chains
(full_update.py:163-227, rdm.py:69-154).  No GPU, no library call: this only checks that every leg permutation of the
reference has been folded into the descriptors correctly, including unequal chi legs."""
import pytest
import torch

from oracle import ctmrg_oracle as orc


def _off(tr, i):
    div, hi, lo = tr
    return (i // div) * hi + (i % div) * lo if div else i * lo


def _gemm_ex(M, N, K, batch, A, B, C, idx, alpha=1.0, beta=0.0, **kw):
    t = [idx[3 * j:3 * j + 3] for j in range(9)]
    assert A.is_contiguous() and B.is_contiguous() and C.is_contiguous()
    Af, Bf, Cf = A.reshape(-1), B.reshape(-1), C.view(-1)
    m, n, k, b = torch.arange(M), torch.arange(N), torch.arange(K), torch.arange(batch)
    oa = _off(t[2], b)[:, None, None] + _off(t[0], m)[None, :, None] + _off(t[1], k)[None, None, :]
    ob = _off(t[5], b)[:, None, None] + _off(t[3], k)[None, :, None] + _off(t[4], n)[None, None, :]
    oc = _off(t[8], b)[:, None, None] + _off(t[6], m)[None, :, None] + _off(t[7], n)[None, None, :]
    assert int(oa.max()) < Af.numel() and int(ob.max()) < Bf.numel() and int(oc.max()) < Cf.numel()
    assert oc.unique().numel() == oc.numel()                 # every output element written exactly once
    Cf[oc.reshape(-1)] = alpha * torch.bmm(Af[oa], Bf[ob]).reshape(-1)
    return C


@pytest.fixture
def emulated(monkeypatch):
    from acetn_b200 import evolution as evo
    from acetn_b200 import measurement as meas
    from acetn_b200 import ops
    monkeypatch.setattr(ops, "gemm_ex", _gemm_ex)
    monkeypatch.setattr(ops, "matmul", lambda A, B, transpose_a=False, **kw: (A.t() if transpose_a else A) @ B)
    monkeypatch.setattr(evo, "contract", lambda spec, A, B: torch.einsum(spec, A, B))
    monkeypatch.setattr(meas, "contract", lambda spec, A, B: torch.einsum(spec, A, B))
    return evo, meas


@pytest.mark.parametrize("D,chi,d", [(2, 3, 2), (3, 4, 2), (2, 5, 3)])
def test_norm_tensor_and_bond_rdm_descriptors(emulated, D, chi, d):
    evo, meas = emulated
    from acetn_b200.ipeps import Ipeps
    cell = orc.random_cell(2, 2, D, chi, d, seed=3)
    nD = min(D ** 3, d * D)
    g = torch.Generator().manual_seed(1)
    a1q = torch.randn(D, D, D, nD, dtype=torch.float64, generator=g)
    a2q = torch.randn(D, D, D, nD, dtype=torch.float64, generator=g)
    ip = Ipeps.from_plain(cell, device="cpu")
    rdm = meas.RDM(ip)
    for bond in cell.bond_list:
        ref = orc.norm_tensor(cell, bond, a1q, a2q)
        got = evo.build_norm_tensor(ip, bond, a1q, a2q)
        assert float((got - ref).norm() / ref.norm()) < 1e-13
        ref = orc.bond_rdm(cell, bond)
        got = rdm[bond]
        assert float((got - ref).norm() / ref.norm()) < 1e-13


def test_norm_tensor_descriptors_unequal_chi_legs(emulated):
    """Every chi leg with its own extent (truncation gives chi' = min(chi, #{s > cutoff}) per projector)."""
    evo, _ = emulated
    D, nD = 2, 4
    g = torch.Generator().manual_seed(7)
    R = lambda *s: torch.randn(*s, dtype=torch.float64, generator=g)      # noqa: E731
    p, q, r, s_, t, f, p2, q2, u2, a3 = 3, 4, 5, 6, 7, 2, 3, 5, 4, 6
    Z, Z4 = torch.zeros(1, 1, dtype=torch.float64), torch.zeros(1, 1, D, D, dtype=torch.float64)
    A_C = [Z, R(p, q), R(r, t), Z]                                      # C[1] = c12, C[2] = c13
    A_E = [R(s_, p, D, D), R(q, r, D, D), R(t, f, D, D), Z4]            # E[0] = e11, E[1] = e12, E[2] = e13
    B_C = [R(p2, q2), Z, Z, R(a3, u2)]                                  # C[0] = c21, C[3] = c24
    B_E = [R(q2, s_, D, D), Z4, R(f, a3, D, D), R(u2, p2, D, D)]        # E[0] = e21, E[2] = e23, E[3] = e24
    ip = {(0, 0): {'C': A_C, 'E': A_E}, (1, 0): {'C': B_C, 'E': B_E}}

    class S:
        def __init__(self, C, E):
            self.C, self.E = C, E

    cl = {(0, 0): S(A_C, A_E), (1, 0): S(B_C, B_E)}
    a1q, a2q = R(D, D, D, nD), R(D, D, D, nD)
    bond = ((0, 0), (1, 0), 0)
    ref = orc.norm_tensor(cl, bond, a1q, a2q)
    got = evo.build_norm_tensor(ip, bond, a1q, a2q)
    assert float((got - ref).norm() / ref.norm()) < 1e-13
