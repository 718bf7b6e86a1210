"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/acetn_b200.h declares; the ctypes binding covers all of them; host-side argument checks raise like the
reference does.  No compute entry point is called here (no GPU in this tier)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "acetn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acetn_b200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from acetn_b200 import _lib
    return _lib.load()


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ["acetn_b200_quarter_tensor", "acetn_b200_rsvd", "acetn_b200_projectors_from_usv", "acetn_b200_absorb_corner1",
                 "acetn_b200_absorb_corner2", "acetn_b200_absorb_edge", "acetn_b200_init", "acetn_b200_destroy",
                 "acetn_b200_last_error", "acetn_b200_version", "acetn_b200_gemm"]:
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    from acetn_b200 import _lib
    for name in declared_symbols():
        assert hasattr(lib, name), f"libacetn_b200.so does not export {name}"
        assert name in _lib.SIGNATURES, f"ctypes binding misses {name}"
    for name in _lib.SIGNATURES:
        assert name in declared_symbols(), f"{name} bound but not declared in include/acetn_b200.h"


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.acetn_b200_version()
    assert isinstance(lib.acetn_b200_last_error(), bytes)


def test_workspace_queries_run_without_gpu(lib):
    from acetn_b200 import _lib
    nb = lib.acetn_b200_quarter_tensor_workspace_bytes(256, 256, 256, 256, 8, 2)
    assert nb > 2 * 2**30          # T2 alone is 2 GiB at D=8 chi=256
    rows, cols = _lib.i64_array([16384, 16384]), _lib.i64_array([16384, 16384])
    assert lib.acetn_b200_rsvd_workspace_bytes(2, rows, cols, 258) > 5 * 16384 * 258 * 8
    assert lib.acetn_b200_orthonormalize_workspace_bytes(16384, 258) > 0
    # small matrices (m <= 4096, q <= 130) take the single-launch kernel: its coefficient / Gram partials come on top of the TSQR scratch
    nchunks, nblk = 4, 3
    assert lib.acetn_b200_orthonormalize_workspace_bytes(1024, 66) >= (nchunks * nblk + 2 * nchunks) * 32 * 33 * 8
    assert lib.acetn_b200_orthonormalize_workspace_bytes(1024, 66) < lib.acetn_b200_orthonormalize_workspace_bytes(16384, 258)
    assert lib.acetn_b200_jacobi_svd_workspace_bytes(258) >= 2 * 258 * 258 * 8
    # the two-stage edge absorption: stage 1 holds P1t + T (2 GiB at D=8 chi=256), stage 2 only GEMM / normalisation scratch
    full = lib.acetn_b200_absorb_edge_workspace_bytes(256, 256, 256, 256, 8, 2)
    begin = lib.acetn_b200_absorb_edge_begin_workspace_bytes(256, 256, 256, 8, 2)
    finish = lib.acetn_b200_absorb_edge_finish_workspace_bytes(256, 256, 256, 8)
    assert 2 * 2**30 < begin < full and finish < 2**30


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    """backend='b200' must fail loudly without a device (north star: no CPU fallback)."""
    from acetn_b200 import ops
    a = torch.zeros(4, 4, dtype=torch.float64)
    with pytest.raises(RuntimeError):
        ops.matmul(a, a)
    assert lib.acetn_b200_init(0) != 0
    assert len(lib.acetn_b200_last_error()) > 0


def test_host_mirror_error_behaviour():
    from acetn_b200.ipeps import CTMRGConfig, Ipeps, SiteTensor
    from acetn_b200.renormalization import ProjectorCalculator
    with pytest.raises(ValueError):
        ProjectorCalculator(CTMRGConfig(projectors="quarter-system"))
    st = SiteTensor(torch.zeros(2, 2, 2, 2, 2), [torch.zeros(1, 1)] * 4, [torch.zeros(1, 1, 2, 2)] * 4)
    with pytest.raises(ValueError):
        st['X']
    assert st.bond_permute(1).shape == (2, 2, 2, 2, 2)
    ip = Ipeps(1, 1, {"phys": 2, "bond": 2, "chi": 1}, {(0, 0): st}, device="cpu")
    with pytest.raises(ValueError):
        ip[(3, 3)]
    assert len(ip.bond_list) == 2 and ip.bond_list[0][2] == 2 and ip.bond_list[1][2] == 1


def test_synthetic_inputs_match_oracle_generator():
    """bench.py's product arm builds its inputs with acetn_b200.synthetic (no oracle import on the product path); the
    generator must produce exactly the tensors the oracle-side tests use and the same flop model."""
    from acetn_b200.synthetic import flops_sweep, random_ipeps
    from oracle import ctmrg_oracle as orc
    ip = random_ipeps(2, 3, 3, 5, 2, seed=7, device="cpu")
    cell = orc.random_cell(2, 3, 3, 5, 2, seed=7)
    for s in cell.site_list:
        assert torch.equal(ip[s]['A'], cell[s].A)
        for k in range(4):
            assert torch.equal(ip[s]['C'][k], cell[s].C[k]) and torch.equal(ip[s]['E'][k], cell[s].E[k])
    assert flops_sweep(2, 2, 8, 256) == orc.flops_sweep(2, 2, 8, 256)
