"""CPU tests of the HOST side of the C-ABI environment contractions (acetn_b200/csrc/environment.cu: acetn_b200_site_rdm,
acetn_b200_bond_rdm, acetn_b200_norm_tensor).  The file contains no kernels -- it validates chi legs, carves the workspace and
builds the two-level index descriptors / leg gathers of every step -- so it is compiled here unchanged and linked against
tests/mock_backend/mock_backend.cu, whose `gemm_launch` / `gather_nd_launch` evaluate the descriptors literally on host
memory.  Results are compared with the oracle's einsum chains (rdm.py:35-154, full_update.py:163-227), including unequal chi
legs, strided site-tensor views (bond_permute), a non-zero split-K scratch requirement (the mock scribbles
over the scratch area, so an overlap with a live buffer shows up as NaNs) and the error paths."""
import ctypes
import os
import subprocess

import pytest
import torch

from oracle import ctmrg_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock_backend")
c_i64, c_sz, c_vp, c_int = ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int
P_i64 = ctypes.POINTER(ctypes.c_int64)


@pytest.fixture(scope="module")
def lib():
    bdir = os.path.join(MOCK, "_build")
    os.makedirs(bdir, exist_ok=True)
    out = os.path.join(bdir, "libenvmock.so")
    srcs = [os.path.join(ROOT, "acetn_b200", "csrc", "environment.cu"), os.path.join(MOCK, "mock_backend.cu")]
    hdrs = [os.path.join(ROOT, "acetn_b200", "csrc", h) for h in ("common.cuh", "gemm.cuh", "kernels.cuh")] + [os.path.join(ROOT, "include", "acetn_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(f) > os.path.getmtime(out) for f in srcs + hdrs):
        cmd = [os.environ.get("NVCC", "nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               "-o", out] + srcs + ["-lcudart"]
        subprocess.check_call(cmd, cwd=MOCK)
    L = ctypes.CDLL(out)
    L.envmock_last_error.restype = ctypes.c_char_p
    L.envmock_set_fake_splitk_bytes.argtypes = [c_sz]
    for name in ("site_rdm", "bond_rdm", "norm_tensor"):
        f = getattr(L, f"acetn_b200_{name}_workspace_bytes")
        f.restype, f.argtypes = c_sz, [P_i64, c_i64, c_i64]
    L.acetn_b200_site_rdm.restype = c_int
    L.acetn_b200_site_rdm.argtypes = [c_vp] * 9 + [P_i64, P_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]
    L.acetn_b200_bond_rdm.restype = c_int
    L.acetn_b200_bond_rdm.argtypes = [c_vp] * 6 + [P_i64] + [c_vp] * 6 + [P_i64, P_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]
    L.acetn_b200_norm_tensor.restype = c_int
    L.acetn_b200_norm_tensor.argtypes = [c_vp] * 12 + [P_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]
    return L


def test_signatures_match_the_product_binding():
    """The mock build binds the same prototypes as acetn_b200/_lib.py (one source of truth: include/acetn_b200.h)."""
    from acetn_b200 import _lib
    for name in ("site_rdm", "bond_rdm", "norm_tensor"):
        assert f"acetn_b200_{name}" in _lib.SIGNATURES and f"acetn_b200_{name}_workspace_bytes" in _lib.SIGNATURES
    assert len(_lib.SIGNATURES["acetn_b200_site_rdm"][1]) == 17
    assert len(_lib.SIGNATURES["acetn_b200_bond_rdm"][1]) == 21
    assert len(_lib.SIGNATURES["acetn_b200_norm_tensor"][1]) == 19


def _arr(v):
    return (ctypes.c_int64 * len(v))(*[int(x) for x in v])


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _chi2(ts):
    out = []
    for t in ts:
        out += [t.shape[0], t.shape[1]]
    return out


def _boundary(a, b, k):
    return [a.C[(k + 1) % 4], a.E[(k + 1) % 4], a.E[k % 4], a.C[(k + 2) % 4], a.E[(k + 2) % 4],
            b.C[k % 4], b.E[k % 4], b.E[(k + 3) % 4], b.C[(k + 3) % 4], b.E[(k + 2) % 4]]


def _ws(nbytes):
    return torch.full((nbytes // 8 + 8,), float("nan"), dtype=torch.float64)


def call_norm(L, a, b, k, a1q, a2q):
    bt = [t.contiguous() for t in _boundary(a, b, k)]
    D, nD = a1q.shape[0], a1q.shape[3]
    chi = _arr(_chi2(bt))
    nb = L.acetn_b200_norm_tensor_workspace_bytes(chi, D, nD)
    assert nb > 0
    ws, out = _ws(nb), torch.full((nD, nD, nD, nD), float("nan"), dtype=torch.float64)
    st = L.acetn_b200_norm_tensor(*[_p(t) for t in bt[:5]], _p(a1q), *[_p(t) for t in bt[5:]], _p(a2q), chi, D, nD, _p(out), _p(ws), nb, None)
    assert st == 0, L.envmock_last_error()
    return out


def call_bond(L, a, b, k):
    bt = [t.contiguous() for t in _boundary(a, b, k)]
    a1, a2 = a.bond_permute(k), b.bond_permute(k)
    D, d = a1.shape[0], a1.shape[4]
    chi = _arr(_chi2(bt))
    nb = L.acetn_b200_bond_rdm_workspace_bytes(chi, D, d)
    ws, out = _ws(nb), torch.full((d, d, d, d), float("nan"), dtype=torch.float64)
    st = L.acetn_b200_bond_rdm(*[_p(t) for t in bt[:5]], _p(a1), _arr(a1.stride()), *[_p(t) for t in bt[5:]], _p(a2), _arr(a2.stride()), chi, D, d,
                               _p(out), _p(ws), nb, None)
    assert st == 0, L.envmock_last_error()
    return out


def call_site(L, s, A=None):
    A = s.A if A is None else A
    C, E = [c.contiguous() for c in s.C], [e.contiguous() for e in s.E]
    D, d = A.shape[0], A.shape[4]
    chi = _arr(_chi2(C + E))
    nb = L.acetn_b200_site_rdm_workspace_bytes(chi, D, d)
    ws, out = _ws(nb), torch.full((d, d), float("nan"), dtype=torch.float64)
    st = L.acetn_b200_site_rdm(*[_p(t) for t in C], *[_p(t) for t in E], _p(A), _arr(A.stride()), chi, D, d, _p(out), _p(ws), nb, None)
    assert st == 0, L.envmock_last_error()
    return out


@pytest.mark.parametrize("D,chi,d,scratch", [(2, 3, 2, 0), (3, 4, 2, 4096), (2, 5, 3, 256)])
def test_environment_contractions_vs_oracle(lib, D, chi, d, scratch):
    lib.envmock_set_fake_splitk_bytes(scratch)
    cell = orc.random_cell(2, 2, D, chi, d, seed=3)
    nD = min(D ** 3, d * D)
    g = torch.Generator().manual_seed(1)
    a1q = torch.randn(D, D, D, nD, dtype=torch.float64, generator=g)
    a2q = torch.randn(D, D, D, nD, dtype=torch.float64, generator=g)
    for bond in cell.bond_list:
        s1, s2, k = bond
        ref = orc.norm_tensor(cell, bond, a1q, a2q)
        got = call_norm(lib, cell[s1], cell[s2], k, a1q, a2q)
        assert float((got - ref).norm() / ref.norm()) < 1e-13
        ref = orc.bond_rdm(cell, bond)
        got = call_bond(lib, cell[s1], cell[s2], k)
        assert float((got - ref).norm() / ref.norm()) < 1e-13
    for site in cell.site_list:
        ref = orc.site_rdm(cell, site)
        got = call_site(lib, cell[site])
        assert float((got - ref).norm() / ref.norm()) < 1e-13
    lib.envmock_set_fake_splitk_bytes(0)


def _ragged_sites(D, d, seed):
    """Two neighbouring sites whose chi legs all differ where the contraction pattern allows it."""
    g = torch.Generator().manual_seed(seed)
    R = lambda *s: torch.randn(*s, dtype=torch.float64, generator=g)      # noqa: E731
    p, q, r, s_, t, f, p2, q2, u2, a3 = 3, 4, 5, 6, 7, 2, 3, 5, 4, 6
    Z, Z4 = torch.zeros(1, 1, dtype=torch.float64), torch.zeros(1, 1, D, D, dtype=torch.float64)
    A = orc.Site(R(D, D, D, D, d), [Z, R(p, q), R(r, t), Z], [R(s_, p, D, D), R(q, r, D, D), R(t, f, D, D), Z4])
    B = orc.Site(R(D, D, D, D, d), [R(p2, q2), Z, Z, R(a3, u2)], [R(q2, s_, D, D), Z4, R(f, a3, D, D), R(u2, p2, D, D)])
    return A, B


def test_unequal_chi_legs(lib):
    """Every chi leg with its own extent (truncation gives chi' = min(chi, #{s > cutoff}) per projector)."""
    D, d, nD = 2, 2, 4
    A, B = _ragged_sites(D, d, 7)
    cell = orc.Cell(2, 1, {}, {(0, 0): A, (1, 0): B})
    g = torch.Generator().manual_seed(8)
    a1q = torch.randn(D, D, D, nD, dtype=torch.float64, generator=g)
    a2q = torch.randn(D, D, D, nD, dtype=torch.float64, generator=g)
    bond = ((0, 0), (1, 0), 0)
    ref = orc.norm_tensor(cell, bond, a1q, a2q)
    got = call_norm(lib, A, B, 0, a1q, a2q)
    assert float((got - ref).norm() / ref.norm()) < 1e-13
    ref = orc.bond_rdm(cell, bond)
    got = call_bond(lib, A, B, 0)
    assert float((got - ref).norm() / ref.norm()) < 1e-13
    # site RDM with a ragged corner ring: c4 (a,b) e4 (b,c) e3 (e,a) c1 (c,g) e1 (g,h) c2 (h,i) e2 (i,j) c3 (j,e)
    gg = torch.Generator().manual_seed(9)
    R = lambda *s: torch.randn(*s, dtype=torch.float64, generator=gg)     # noqa: E731
    a, b, c, e, g_, h, i, j = 2, 3, 4, 5, 6, 3, 5, 4
    S = orc.Site(R(D, D, D, D, d), [R(c, g_), R(h, i), R(j, e), R(a, b)], [R(g_, h, D, D), R(i, j, D, D), R(e, a, D, D), R(b, c, D, D)])
    ref = orc.site_rdm(orc.Cell(1, 1, {}, {(0, 0): S}), (0, 0))
    got = call_site(lib, S)
    assert float((got - ref).norm() / ref.norm()) < 1e-13


def test_site_rdm_strided_site_tensor(lib):
    """A = a non-contiguous view (the reference hands bond_permute views around; strides go through the ABI)."""
    cell = orc.random_cell(1, 1, 2, 3, 2, seed=4)
    s = cell[(0, 0)]
    base = s.A.permute(1, 0, 3, 2, 4).contiguous()
    view = base.permute(1, 0, 3, 2, 4)
    assert not view.is_contiguous() and torch.equal(view, s.A)
    ref = orc.site_rdm(cell, (0, 0))
    got = call_site(lib, s, A=view)
    assert float((got - ref).norm() / ref.norm()) < 1e-13


def test_error_paths(lib):
    """Mismatched chi legs -> ERR_INVALID with a message; too small a workspace -> ERR_WORKSPACE; nothing is written."""
    cell = orc.random_cell(2, 1, 2, 3, 2, seed=5)
    a, b = cell[(0, 0)], cell[(1, 0)]
    bt = [t.contiguous() for t in _boundary(a, b, 2)]
    chi = _chi2(bt)
    bad = list(chi)
    bad[2] += 1                                   # e12's first chi leg no longer equals c12's second
    assert lib.acetn_b200_norm_tensor_workspace_bytes(_arr(bad), 2, 4) == 0
    a1q = torch.zeros(2, 2, 2, 4, dtype=torch.float64)
    out = torch.full((4, 4, 4, 4), 7.0, dtype=torch.float64)
    ws = _ws(1 << 20)
    st = lib.acetn_b200_norm_tensor(*[_p(t) for t in bt[:5]], _p(a1q), *[_p(t) for t in bt[5:]], _p(a1q), _arr(bad), 2, 4, _p(out), _p(ws), 1 << 20, None)
    assert st == 1 and b"chi legs" in lib.envmock_last_error()
    st = lib.acetn_b200_norm_tensor(*[_p(t) for t in bt[:5]], _p(a1q), *[_p(t) for t in bt[5:]], _p(a1q), _arr(chi), 2, 4, _p(out), _p(ws), 512, None)
    assert st == 2 and b"workspace too small" in lib.envmock_last_error()
    assert float(out.min()) == 7.0 and float(out.max()) == 7.0
