"""GPU tests of K7 (acetn_b200/csrc/i8crt.cu): the "big x thin" products of the rSVD chain and projector formation evaluated
exactly in integer arithmetic on the INT8 tensor cores (tcgen05.mma kind::i8, residue number system + CRT).

  * the product itself against FP64 references (tolerance: 1e-13 of the row scale; integer inputs must come out EXACT);
  * the CTMRG path with thin_engine='i8' against the CPU oracle on the north-star tolerances (spectra 1e-10, energy 1e-9);
  * K7 against K1 (FP64 DMMA) on the same inputs at larger sizes."""
import pytest
import torch

from oracle import ctmrg_oracle as orc

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from acetn_b200 import linalg, ops
    from acetn_b200.ipeps import CTMRGConfig, Ipeps
    from acetn_b200.renormalization import DirectionalMover, ProjectorCalculator
    from tests.test_gpu_ctmrg import assert_same_shapes, energy, neel, push_state, run_b200, run_oracle, spectra_err, to_oracle_cell


def _rand(shape, seed, dev="cuda"):
    return torch.randn(*shape, dtype=torch.float64, generator=torch.Generator().manual_seed(seed)).to(dev)


@pytest.mark.parametrize("rows,cols,q", [(128, 128, 1), (256, 384, 40), (300, 200, 37), (1000, 777, 258), (2048, 1024, 272), (640, 4224, 130)])
@pytest.mark.parametrize("adjoint", [False, True])
def test_i8_matmul_vs_fp64(rows, cols, q, adjoint):
    Q = _rand((rows, cols), rows + cols)
    Y = _rand((rows if adjoint else cols, q), q)
    enc = ops.i8_encode(Q)
    out = ops.i8_matmul(enc, Y, adjoint=adjoint)
    ref = ((Q.T if adjoint else Q).cpu() @ Y.cpu())
    scale = ref.abs().amax(dim=1, keepdim=True)
    assert float(((out.cpu() - ref).abs() / scale).max()) < 1e-13


@pytest.mark.parametrize("adjoint", [False, True])
def test_i8_matmul_integer_inputs_exact(adjoint):
    """Small-integer operands: every intermediate is exactly representable, so the result must equal the integer product
    bit for bit (this is what 'exact arithmetic' means here; FP64 DGEMM only achieves it for products below 2^53)."""
    g = torch.Generator().manual_seed(5)
    Q = torch.randint(-1000, 1000, (512, 640), generator=g)
    Y = torch.randint(-1000, 1000, (512 if adjoint else 640, 66), generator=g)
    ref = (Q.T if adjoint else Q) @ Y
    out = ops.i8_matmul(ops.i8_encode(Q.double().cuda()), Y.double().cuda(), adjoint=adjoint)
    assert torch.equal(out.cpu(), ref.double())


@pytest.mark.parametrize("adjoint", [False, True])
def test_i8_matmul_graded_and_zero_lines(adjoint):
    """Rows/columns spanning many orders of magnitude (the chi legs of converged boundaries do) and exactly-zero rows and
    columns (product-state boundaries): the two-sided scaling keeps the error at FP64 level relative to each output row."""
    rows, cols, q = 384, 512, 50
    Q = _rand((rows, cols), 1) * torch.logspace(0, -11, rows, dtype=torch.float64, device="cuda")[:, None] \
        * torch.logspace(0, -9, cols, dtype=torch.float64, device="cuda")[None, :]
    Q[7, :] = 0.0
    Q[:, 11] = 0.0
    Y = _rand((rows if adjoint else cols, q), 2)
    Y[:, 4] = 0.0
    out = ops.i8_matmul(ops.i8_encode(Q), Y, adjoint=adjoint).cpu()
    A = (Q.T if adjoint else Q).cpu()
    ref = A @ Y.cpu()
    bound = (A.abs() @ Y.cpu().abs())                       # componentwise DGEMM-style scale
    assert float(((out - ref).abs() / bound.clamp_min(1e-300)).max()) < 1e-11
    assert torch.all(out[:, 4] == 0)
    assert torch.all(out[11 if adjoint else 7, :] == 0)


def test_i8_unsupported_shapes_raise():
    assert not ops.i8_supported(64, 4096, 10)
    assert not ops.i8_supported(4096, 4096, 300)
    with pytest.raises(RuntimeError):
        ops.i8_encode(_rand((64, 64), 0))
    cell = orc.random_cell(2, 2, 2, 8, 2, seed=0)
    ip = Ipeps.from_plain(cell, CTMRGConfig(thin_engine="i8"))
    with pytest.raises(RuntimeError):
        ProjectorCalculator(ip.ctmrg_config).calculate(ip, orc.plaquette(cell, 0, 0, 0), 0)
    with pytest.raises(ValueError):
        ProjectorCalculator(CTMRGConfig(thin_engine="bogus"))


def test_rsvd_low_rank_recovery_i8():
    """reference tests/unit/test_linalg.py:45-55 (exactly low-rank product is recovered to 1e-10), through K7."""
    m, k, n, r = 512, 384, 640, 20
    A = _rand((m, r), 1) @ _rand((r, k), 2)
    B = _rand((k, n), 3)
    omega = _rand((n, r + 4), 4)
    encs = [ops.i8_encode(A), ops.i8_encode(B)]
    U, S, V, _ = ops.rsvd([None, None], omega, niter=2, encs=encs)
    rec = (U * S) @ V.T
    full = A @ B
    assert float((rec - full).norm() / full.norm()) < 1e-10
    assert float((U.T @ U - torch.eye(r + 4, dtype=torch.float64, device="cuda")).abs().max()) < 1e-12


SYNC_I8 = [("random", 4, 16, 2, 5), ("random", 3, 16, 3, 6), ("product", 3, 18, 2, 7), ("random", 4, 64, 2, 8)]


@pytest.mark.parametrize("kind,D,chi,d,seed", SYNC_I8)
def test_synchronized_moves_i8(kind, D, chi, d, seed):
    """tests/test_gpu_ctmrg.py::test_synchronized_moves with every thin product on the INT8 tensor cores."""
    nx = ny = 2
    cell = orc.random_cell(nx, ny, D, chi, d, seed=seed) if kind == "random" else orc.product_cell(nx, ny, D, chi, d, seed=seed, state_map=neel)
    torch.manual_seed(100 + seed)
    cfg = orc.CtmrgConfig()
    # product-state boundaries start with chi' << chi (quarter tensors below K7's 128-row minimum): 'auto' with the size
    # threshold lowered sends every product K7 supports through it; random cells are saturated, so 'i8' can be forced
    ip = Ipeps.from_plain(cell, CTMRGConfig(thin_engine="i8" if kind == "random" else "auto"))
    mover = DirectionalMover(ip.ctmrg_config)
    mover.projector_calculator.I8_MIN_DIM = 128
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


@pytest.mark.parametrize("projectors", ["half-system", "full-system"])
def test_free_running_product_state_i8(projectors):
    """Product-state start (rank-deficient quarter tensors with exactly-zero rows/columns), 3 free-running sweeps."""
    D, chi = 3, 18
    cell = orc.product_cell(2, 2, D, chi, 2, seed=3, state_map=neel)
    torch.manual_seed(7)
    ref, s_ref, tape = run_oracle(cell, 3, projectors=projectors)
    ip = Ipeps.from_plain(cell, CTMRGConfig(steps=3, projectors=projectors, thin_engine="auto"))
    mover = DirectionalMover(ip.ctmrg_config)
    mover.projector_calculator.thin_engine = "auto"
    mover.projector_calculator.I8_MIN_DIM = 128          # route every supported product through K7, fall back below 128
    mover.projector_calculator.spectra = []
    from acetn_b200.renormalization import ctmrg
    linalg.set_omega_source(orc.OmegaTape(tape))
    try:
        ctmrg(ip, ip.ctmrg_config, mover)
    finally:
        linalg.set_omega_source(None)
    got = to_oracle_cell(ip)
    assert_same_shapes(ref, got)
    assert spectra_err(s_ref, mover.projector_calculator.spectra, chi) < 1e-10
    assert abs(energy(got) - energy(ref)) < 1e-9


@pytest.mark.parametrize("D,chi", [(8, 64), (6, 144)])
def test_i8_engine_matches_dmma_engine(D, chi):
    """One left move from the same state and Omega with both engines: spectra to 1e-10, absorbed tensors' gauge invariants."""
    cell = orc.random_cell(2, 2, D, chi, 2, seed=2)
    out = {}
    for eng in ("dmma", "i8"):
        ip = Ipeps.from_plain(cell, CTMRGConfig(thin_engine=eng))
        mover = DirectionalMover(ip.ctmrg_config)
        mover.projector_calculator.spectra = []
        torch.manual_seed(11)
        mover.left_move(ip, 0)
        out[eng] = (mover.projector_calculator.spectra, ip)
    for a, b in zip(out["dmma"][0], out["i8"][0]):
        assert float((a - b)[:chi].abs().max()) < 1e-10
    for y in range(2):
        for k in (0, 3):
            sa = torch.linalg.svdvals(out["dmma"][1][(1, y)]['C'][k].cpu())
            sb = torch.linalg.svdvals(out["i8"][1][(1, y)]['C'][k].cpu())
            assert float((sa / sa[0] - sb / sb[0]).abs().max()) < 1e-9


@pytest.mark.parametrize("D,chi,d", [(8, 32, 2), (8, 20, 2), (6, 16, 2), (4, 12, 2)])
def test_quarter_tensor_enc_equals_separate_encoding(D, chi, d):
    """acetn_b200_quarter_tensor_enc (quarter tensor + K7 encoding in one call; for D = 8, d = 2 the column exponents come from the
    producing kernel's epilogue) must give bit for bit the storage of acetn_b200_i8_encode applied to the same tensor: residue
    planes, row and column exponents.  Ragged chi legs and an exactly-zero edge slice included."""
    cell = orc.random_cell(2, 2, D, chi, d, seed=11)
    st = cell[(0, 0)]
    C, E2, E1 = st.C[0].cuda(), st.E[0][:, :chi - 1].contiguous().cuda(), st.E[3][:chi - 2].contiguous().cuda()
    E2[:, 3] = 0.0                                   # a block of exactly-zero rows of Q
    E1[5] = 0.0                                      # a block of exactly-zero columns of Q
    A = st.bond_permute(0).cuda()
    rows, cols = (chi - 1) * D * D, (chi - 2) * D * D
    nb = ops.i8_encoded_bytes(rows, cols)
    s_fused = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    Q, _, enc = ops.quarter_tensor(C, E2, E1, A, normalize=False, enc_storage=s_fused)
    assert (enc.rows, enc.cols) == (rows, cols) and tuple(Q.shape) == (rows, cols)
    Qref, _ = ops.quarter_tensor(C, E2, E1, A, normalize=False)
    assert torch.equal(Q, Qref)
    s_sep = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    ops.i8_encode(Qref, storage=s_sep)
    ld = (cols + 127) // 128 * 128
    planes = 16 * rows * ld
    a = s_fused[:planes].view(16, rows, ld)[:, :, :cols]
    b = s_sep[:planes].view(16, rows, ld)[:, :, :cols]
    assert torch.equal(a, b)
    off = (planes + 255) // 256 * 256
    rexp = lambda s: s[off:off + 4 * rows].view(torch.int32)                                  # noqa: E731
    cexp = lambda s: s[off + (4 * rows + 255) // 256 * 256:][:4 * cols].view(torch.int32)      # noqa: E731
    assert torch.equal(rexp(s_fused), rexp(s_sep)) and torch.equal(cexp(s_fused), cexp(s_sep))
    # and the products agree with FP64
    Y = _rand((cols, 10), 3)
    ref = Qref @ Y
    out = ops.i8_matmul(enc, Y)
    scale = ref.abs().amax(dim=1, keepdim=True).clamp_min(1e-300)
    assert float(((out - ref).abs() / scale).max()) < 1e-13
