"""GPU parity tests of the individual kernels behind the C ABI (K1 GEMM, K4 TSQR, K5 Jacobi, rSVD, quarter tensor,
projectors, absorption) against torch fp64 / the CPU oracle on identical seeded inputs."""
import pytest
import torch

from oracle import ctmrg_oracle as orc

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from acetn_b200 import ops

DEV = "cuda"


def rel(a, b):
    return float((a - b).norm() / b.norm())


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, dtype=torch.float64, generator=g).to(DEV)


# ---------------------------------------------------------------------------------------------- K1 GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (129, 131, 67), (256, 258, 1000), (1000, 88, 515), (64, 64, 16),
                                   (7, 5, 3), (300, 33, 2048), (512, 264, 4096)])
@pytest.mark.parametrize("tile", [0, 1, 2, 3])
def test_gemm_nn(M, N, K, tile):
    A, B = rnd(M, K, seed=1), rnd(K, N, seed=2)
    C = ops.matmul(A, B, force_tile=tile)
    assert rel(C, A @ B) < 1e-13


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (130, 90, 77), (258, 258, 4096), (1024, 32, 256)])
@pytest.mark.parametrize("tile", [1, 2, 3])
def test_gemm_tn(M, N, K, tile):
    A, B = rnd(K, M, seed=3), rnd(K, N, seed=4)
    C = ops.matmul(A, B, transpose_a=True, force_tile=tile)
    assert rel(C, A.T @ B) < 1e-13


@pytest.mark.parametrize("splitk", [2, 3, 7])
def test_gemm_splitk(splitk):
    A, B = rnd(200, 3000, seed=5), rnd(3000, 150, seed=6)
    C = ops.matmul(A, B, force_splitk=splitk)
    assert rel(C, A @ B) < 1e-13


def test_gemm_b_kcontig_and_beta():
    # C = alpha * A @ Bt^T + beta * C with Bt stored (N, K)
    M, N, K = 150, 70, 333
    A, Bt, C0 = rnd(M, K, seed=7), rnd(N, K, seed=8), rnd(M, N, seed=9)
    C = C0.clone()
    idx = [0, 0, K, 0, 0, 1, 0, 0, 0] + [0, 0, 1, 0, 0, K, 0, 0, 0] + [0, 0, N, 0, 0, 1, 0, 0, 0]
    ops.gemm_ex(M, N, K, 1, A, Bt, C, idx, alpha=-0.5, beta=2.0)
    assert rel(C, -0.5 * A @ Bt.T + 2.0 * C0) < 1e-13


@pytest.mark.parametrize("D", [2, 3, 4])
def test_gemm_two_level_and_batched(D):
    # T2[(c,uU),(e,lL)] = sum_a T1[a,(c,uU)] E1[e,a,lL]  (projectors.py:53), two-level n index on B
    xa, xc, xe = 11, 9, 10
    D2 = D * D
    T1, E1 = rnd(xa, xc * D2, seed=10), rnd(xe, xa, D, D, seed=11)
    C = torch.empty(xc * D2, xe * D2, dtype=torch.float64, device=DEV)
    idx = [0, 0, 1, 0, 0, xc * D2, 0, 0, 0] + [0, 0, D2, D2, xa * D2, 1, 0, 0, 0] + [0, 0, xe * D2, 0, 0, 1, 0, 0, 0]
    ops.gemm_ex(xc * D2, xe * D2, xa, 1, T1, E1, C, idx)
    ref = torch.einsum("am,eal->mel", T1, E1.reshape(xe, xa, D2)).reshape(xc * D2, xe * D2)
    assert rel(C, ref) < 1e-13
    # batched: C[b] = A[b] @ B[b]
    nb, M, N, K = 13, 20, 24, 17
    A, B = rnd(nb, M, K, seed=12), rnd(nb, K, N, seed=13)
    Cb = torch.empty(nb, M, N, dtype=torch.float64, device=DEV)
    idx = [0, 0, K, 0, 0, 1, 0, 0, M * K] + [0, 0, N, 0, 0, 1, 0, 0, K * N] + [0, 0, N, 0, 0, 1, 0, 0, M * N]
    ops.gemm_ex(M, N, K, nb, A, B, Cb, idx)
    assert rel(Cb, A @ B) < 1e-13


def test_gemm_epilogue_paired_and_scalar_stores():
    """K1's epilogue writes adjacent column pairs as one 16-byte store when the n index of C is contiguous and aligned
    (gemm.cu store_accumulators) and falls back to scalar stores otherwise; alpha = 1 / beta = 0 skips the FP64 scaling.
    Ragged edges, odd leading dimension (scalar path), batches with beta != 0, the two-level contraction class."""
    A, B = rnd(2763, 48, seed=21), rnd(48, 2901, seed=22)       # N odd: rows of C start at odd offsets -> scalar stores
    assert rel(ops.matmul(A, B, force_tile=3), A @ B) < 1e-13
    At, Bn = rnd(64, 2890, seed=23), rnd(64, 2800, seed=24)     # paired stores, ragged M
    assert rel(ops.matmul(At, Bn, transpose_a=True, force_tile=3), At.T @ Bn) < 1e-13
    for tile in (1, 2, 3):
        A2, B2 = rnd(333, 40, seed=30 + tile), rnd(40, 301, seed=40 + tile)   # last column unpaired
        assert rel(ops.matmul(A2, B2, force_tile=tile), A2 @ B2) < 1e-13
    nb, M, N, K = 3, 1600, 1700, 32
    Ab, Bb, C0 = rnd(nb, M, K, seed=25), rnd(nb, K, N, seed=26), rnd(nb, M, N, seed=27)
    idx = [0, 0, K, 0, 0, 1, 0, 0, M * K] + [0, 0, N, 0, 0, 1, 0, 0, K * N] + [0, 0, N, 0, 0, 1, 0, 0, M * N]
    Cb = C0.clone()
    ops.gemm_ex(M, N, K, nb, Ab, Bb, Cb, idx, alpha=0.5, beta=-1.5, force_tile=3)
    assert rel(Cb, 0.5 * Ab @ Bb - 1.5 * C0) < 1e-13
    # the 2 chi^3 D^4 contraction class of projectors.py:53 (two-level n index on B), D = 4
    D, xa, xc, xe = 4, 64, 170, 172
    D2 = D * D
    T1, E1 = rnd(xa, xc * D2, seed=28), rnd(xe, xa, D, D, seed=29)
    idx = [0, 0, 1, 0, 0, xc * D2, 0, 0, 0] + [0, 0, D2, D2, xa * D2, 1, 0, 0, 0] + [0, 0, xe * D2, 0, 0, 1, 0, 0, 0]
    Cc = torch.empty(xc * D2, xe * D2, dtype=torch.float64, device=DEV)
    ops.gemm_ex(xc * D2, xe * D2, xa, 1, T1, E1, Cc, idx, force_tile=3)
    ref = torch.einsum("am,eal->mel", T1, E1.reshape(xe, xa, D2)).reshape(xc * D2, xe * D2)
    assert rel(Cc, ref) < 1e-13
    # C written through a transposed descriptor (n index strided): scalar path
    M, N, K = 130, 70, 48
    A3, B3 = rnd(M, K, seed=50), rnd(K, N, seed=51)
    Ct = torch.empty(N, M, dtype=torch.float64, device=DEV)
    idx = [0, 0, K, 0, 0, 1, 0, 0, 0] + [0, 0, N, 0, 0, 1, 0, 0, 0] + [0, 0, 1, 0, 0, M, 0, 0, 0]
    ops.gemm_ex(M, N, K, 1, A3, B3, Ct, idx)
    assert rel(Ct, (A3 @ B3).T) < 1e-13


# ---------------------------------------------------------------------------------------------- K4 TSQR
@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("m,q", [(300, 20), (1000, 64), (4097, 130), (16384, 258), (80, 22), (40, 40), (257, 33), (1024, 66), (4096, 130),
                                 (1296, 38), (513, 97)])
def test_orthonormalize_random(m, q, fused, monkeypatch):
    """K4 on both of its launch forms: the single cooperative kernel for small matrices (m <= 4096, q <= 130) and the multi-launch path."""
    if fused == "0" and (m > 4096 or q > 130):
        pytest.skip("only the multi-launch path exists at this size (covered by fused='1')")
    monkeypatch.setenv("ACETN_B200_ORTHO_FUSED", fused)
    Y0 = rnd(m, q, seed=20)
    Q = ops.orthonormalize(Y0.clone())
    eye = torch.eye(q, dtype=torch.float64, device=DEV)
    assert float((Q.T @ Q - eye).abs().max()) < 5e-14
    assert rel(Q @ (Q.T @ Y0), Y0) < 1e-13


@pytest.mark.parametrize("fused", ["1", "0"])
@pytest.mark.parametrize("householder", ["0", "1"])
@pytest.mark.parametrize("kind", ["graded", "rank_deficient", "graded_columns"])
def test_orthonormalize_ill_conditioned(kind, householder, fused, monkeypatch):
    """Graded spectrum over 20 decades, exact rank deficiency, columns scaled over 30 decades: the pivot test of the CholeskyQR2 fast
    path must hand such panels to the Householder TSQR (fused kernel and multi-launch path alike); `householder=1` forces it."""
    monkeypatch.setenv("ACETN_B200_ORTHO_FUSED", fused)
    monkeypatch.setenv("ACETN_B200_TSQR_HOUSEHOLDER", householder)
    m, q = 2048, 96
    U = torch.linalg.qr(rnd(m, q, seed=21)).Q
    V = torch.linalg.qr(rnd(q, q, seed=22)).Q
    if kind == "graded":
        s = torch.logspace(0, -20, q, dtype=torch.float64, device=DEV)
        Y0 = (U * s) @ V.T
    elif kind == "rank_deficient":
        s = torch.cat([torch.ones(10, dtype=torch.float64, device=DEV), torch.zeros(q - 10, dtype=torch.float64, device=DEV)])
        Y0 = (U * s) @ V.T
    else:
        Y0 = rnd(m, q, seed=23) * torch.logspace(0, -30, q, dtype=torch.float64, device=DEV)
    Q = ops.orthonormalize(Y0.clone())
    eye = torch.eye(q, dtype=torch.float64, device=DEV)
    assert float((Q.T @ Q - eye).abs().max()) < 1e-13
    if kind == "graded_columns":
        # every column of Y0 must lie in span(Q) relative to ITS OWN norm (a scaled column is as good a direction as any other)
        R = Y0 - Q @ (Q.T @ Y0)
        assert float((R.norm(dim=0) / Y0.norm(dim=0)).max()) < 1e-12
    else:
        assert float((Q @ (Q.T @ Y0) - Y0).norm() / Y0.norm()) < 1e-13


@pytest.mark.parametrize("m,q", [(1024, 66), (80, 22), (2500, 128)])
def test_orthonormalize_fused_matches_multi_launch(m, q, monkeypatch):
    """Same algorithm, same per-chunk arithmetic: the fused kernel's basis equals the multi-launch path's up to the rounding of the
    inter-panel projections (FMA chunk code vs K1), far inside the rSVD's tolerance."""
    Y0 = rnd(m, q, seed=31)
    out = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("ACETN_B200_ORTHO_FUSED", fused)
        out[fused] = ops.orthonormalize(Y0.clone())
    assert float((out["0"] - out["1"]).abs().max()) < 1e-12


# ---------------------------------------------------------------------------------------------- K5 Jacobi
@pytest.mark.parametrize("small", ["1", "0"])
@pytest.mark.parametrize("q", [1, 2, 5, 22, 37, 64, 66, 111, 112, 113, 129, 258])
def test_jacobi_svd(q, small, monkeypatch):
    # the core handed to K5 is the triangular factor of a QR (acetn_b200 rsvd pipeline); condition number 1e9.  Cores up to q = 112 run
    # in the single-CTA kernel (jacobi_small_kernel), larger ones in the multi-CTA block kernel; small="0" forces the latter.
    if small == "0" and q > 112:
        pytest.skip("covered by small='1' (the block kernel is the only one at this size)")
    monkeypatch.setenv("ACETN_B200_JACOBI_SMALL", small)
    G = rnd(q, q, seed=30) * torch.logspace(0, -9, q, dtype=torch.float64, device=DEV)[None, :]
    R = torch.linalg.qr(G.cpu()).R.to(DEV).contiguous()
    S, Wt, Jt, info = ops.jacobi_svd(R)
    ref = torch.linalg.svdvals(R.cpu()).to(DEV)
    assert float((S - ref).abs().max() / ref[0]) < 5e-14
    assert torch.all(S[:-1] >= S[1:])
    eye = torch.eye(q, dtype=torch.float64, device=DEV)
    assert float((Wt @ Wt.T - eye).abs().max()) < 1e-13
    assert float((Jt @ Jt.T - eye).abs().max()) < 1e-13
    assert rel(Jt.T @ torch.diag(S) @ Wt, R) < 1e-13
    assert 0 < int(info[1]) <= 15


def test_jacobi_svd_rank_deficient_and_truncation_count(monkeypatch):
    """Zero singular values (exactly rank-deficient core, as on product-state boundaries) and the truncation count s/s0 > cutoff on
    both kernels."""
    q = 48
    g = torch.Generator().manual_seed(33)
    A = torch.randn(q, 10, dtype=torch.float64, generator=g) @ torch.randn(10, q, dtype=torch.float64, generator=g)
    R = torch.linalg.qr(A).R.to(DEV).contiguous()
    ref = torch.linalg.svdvals(R.cpu()).to(DEV)
    for small in ("1", "0"):
        monkeypatch.setenv("ACETN_B200_JACOBI_SMALL", small)
        S, Wt, Jt, info = ops.jacobi_svd(R, chi=40, cutoff=1e-10)
        assert float((S - ref).abs().max() / ref[0]) < 5e-14
        assert int(info[0]) == 10
        assert rel(Jt.T @ torch.diag(S) @ Wt, R) < 1e-13


@pytest.mark.parametrize("small", ["1", "0"])
def test_jacobi_svd_dense_unpreconditioned(small, monkeypatch):
    """A dense core that is not QR-preconditioned needs more sweeps but must still converge."""
    monkeypatch.setenv("ACETN_B200_JACOBI_SMALL", small)
    q = 64
    R = rnd(q, q, seed=31) * torch.logspace(0, -6, q, dtype=torch.float64, device=DEV)[None, :]
    S, Wt, Jt, info = ops.jacobi_svd(R)
    ref = torch.linalg.svdvals(R.cpu()).to(DEV)
    assert float((S - ref).abs().max() / ref[0]) < 5e-14
    assert rel(Jt.T @ torch.diag(S) @ Wt, R) < 1e-13


# ---------------------------------------------------------------------------------------------- rSVD
def test_rsvd_against_oracle_same_omega():
    g = torch.Generator().manual_seed(40)
    A = torch.randn(300, 200, dtype=torch.float64, generator=g)
    B = torch.randn(200, 260, dtype=torch.float64, generator=g)
    tape = orc.OmegaTape()
    U0, S0, V0 = orc.fused_matmul_svd_lowrank(A, B, q=34, niter=2, omega_fn=tape)
    U, S, V, info = ops.rsvd([A.to(DEV), B.to(DEV)], tape.tape[0].to(DEV), niter=2)
    assert float((S.cpu() - S0).abs().max() / S0[0]) < 1e-10
    eye = torch.eye(34, dtype=torch.float64, device=DEV)
    assert float((U.T @ U - eye).abs().max()) < 1e-10
    assert float((V.T @ V - eye).abs().max()) < 1e-10
    # same subspaces / same reconstruction (gauge invariant)
    assert rel((U * S) @ V.T, ((U0 * S0) @ V0.T).to(DEV)) < 1e-8


@pytest.mark.parametrize("nmat", [1, 2, 4])
def test_rsvd_low_rank_recovery(nmat):
    """reference tests/unit/test_linalg.py:45-55,151-161,231-241 : exact recovery of low-rank inputs to 1e-10."""
    g = torch.Generator().manual_seed(41)
    low = (torch.randn(120, 6, dtype=torch.float64, generator=g) @ torch.randn(6, 90, dtype=torch.float64, generator=g)).to(DEV)
    mats = [low]
    if nmat >= 2:
        mats.append(rnd(90, 100, seed=42))
    if nmat == 4:
        mats += [rnd(100, 80, seed=43), rnd(80, 70, seed=44)]
    full = mats[0]
    for m_ in mats[1:]:
        full = full @ m_
    omega = rnd(mats[-1].shape[1], 12, seed=45)
    U, S, V, _ = ops.rsvd(mats, omega, niter=2, reorth_adjoint=(nmat == 4))
    assert rel((U * S) @ V.T, full) < 1e-10
    assert torch.all(S[:-1] >= S[1:])


# ---------------------------------------------------------------------------------------------- quarter / absorb
CASES = [(2, 8, 2, 0), (3, 12, 2, 1), (4, 16, 2, 2), (3, 10, 3, 3), (8, 16, 2, 4)]


@pytest.mark.parametrize("D,chi,d,seed", CASES)
def test_quarter_tensor(D, chi, d, seed):
    cell = orc.random_cell(2, 2, D, chi, d, seed=seed)
    st = cell[(0, 0)]
    A = st.A.to(DEV)
    for k in range(4):
        ref, shp = orc.quarter_tensor(st, k)
        ak = A.permute([(i + k) % 4 for i in range(4)] + [4])
        Q, qD = ops.quarter_tensor(st.C[k % 4].to(DEV), st.E[k % 4].to(DEV), st.E[(3 + k) % 4].to(DEV), ak)
        assert tuple(qD) == tuple(shp)
        assert rel(Q.cpu(), ref) < 1e-13


def test_quarter_tensor_ragged_chi():
    """chi legs need not be equal."""
    D, d = 3, 2
    g = torch.Generator().manual_seed(5)
    A = torch.rand(D, D, D, D, d, dtype=torch.float64, generator=g) - 0.5
    C = torch.rand(7, 5, dtype=torch.float64, generator=g)
    E2 = torch.rand(5, 6, D, D, dtype=torch.float64, generator=g)
    E1 = torch.rand(4, 7, D, D, dtype=torch.float64, generator=g)
    st = orc.Site(A, [C] * 4, [E2, E2, E2, E1])
    ref, shp = orc.quarter_tensor(st, 0)
    Q, qD = ops.quarter_tensor(C.to(DEV), E2.to(DEV), E1.to(DEV), A.to(DEV))
    assert tuple(qD) == tuple(shp) == (6, D, D, 4, D, D)
    assert rel(Q.cpu(), ref) < 1e-13


@pytest.mark.parametrize("D,chi,d,seed", CASES)
def test_absorption(D, chi, d, seed):
    cell = orc.random_cell(2, 2, D, chi, d, seed=seed)
    st = cell[(0, 0)]
    A = st.A.to(DEV)
    g = torch.Generator().manual_seed(seed)
    for k in range(4):
        pj1 = torch.rand(chi, D, D, chi - 1, dtype=torch.float64, generator=g)
        pj2 = torch.rand(chi, D, D, chi - 2, dtype=torch.float64, generator=g)
        ak = A.permute([(i + k) % 4 for i in range(4)] + [4])
        c1 = ops.absorb_corner1(st.C[(3 + k) % 4].to(DEV), st.E[(2 + k) % 4].to(DEV), pj1.to(DEV))
        c2 = ops.absorb_corner2(st.C[k].to(DEV), st.E[k].to(DEV), pj2.to(DEV))
        e = ops.absorb_edge(st.E[(3 + k) % 4].to(DEV), ak, pj2.to(DEV), pj1.to(DEV))
        assert rel(c1.cpu(), orc.absorb_corner1(st.C[(3 + k) % 4], st.E[(2 + k) % 4], pj1)) < 1e-13
        assert rel(c2.cpu(), orc.absorb_corner2(st.C[k], st.E[k], pj2)) < 1e-13
        ref_e = orc.absorb_edge(st.E[(3 + k) % 4], st.bond_permute(k), pj2, pj1)
        assert e.shape == ref_e.shape
        assert rel(e.cpu(), ref_e) < 1e-13


# ---------------------------------------------------------------------------------------------- K1 fuzz
def _rand_operand(rng, rows, cols, nb):
    """A (rows x cols) logical matrix per batch stored inside a random larger tensor, with a random two-level split of
    each index group.  Returns (tensor, ptr_tensor, idx9, dense) where idx9 = {div,s_hi,s_lo} for (row, col, batch)."""
    def split(n):
        divs = [d for d in range(2, n) if n % d == 0]
        if divs and rng.random() < 0.6:
            d = divs[int(rng.integers(len(divs)))]
            return n // d, d           # (hi extent, lo extent)
        return 1, n
    rh, rl = split(rows)
    ch, cl = split(cols)
    # storage order: a random permutation of the legs (b, rh, rl, ch, cl), contiguous
    legs = {"b": nb, "rh": rh, "rl": rl, "ch": ch, "cl": cl}
    order = list(legs)
    rng.shuffle(order)
    shape = [legs[k] for k in order]
    t = torch.from_numpy(rng.standard_normal(shape)).to(DEV)
    strides = dict(zip(order, t.stride()))
    dense = lambda: t.permute([order.index(k) for k in ("b", "rh", "rl", "ch", "cl")]).reshape(nb, rows, cols)   # noqa: E731
    idx = [rl if rh > 1 else 0, strides["rh"] if rh > 1 else 0, strides["rl"],
           cl if ch > 1 else 0, strides["ch"] if ch > 1 else 0, strides["cl"],
           0, 0, strides["b"] if nb > 1 else 0]
    return t, idx, dense


@pytest.mark.parametrize("seed", range(24))
def test_gemm_fuzz_two_level_descriptors(seed):
    """Random shapes, leg orders (all four smem orientations + scalar/vector loaders), two-level splits, batches,
    alpha/beta, tiles and split-K against torch.einsum."""
    import numpy as np
    rng = np.random.default_rng(seed)
    M, N, K = (int(rng.integers(1, 200)) for _ in range(3))
    nb = int(rng.integers(1, 4))
    A_t, a_idx, A_d = _rand_operand(rng, M, K, nb)
    B_t, b_idx, B_d = _rand_operand(rng, K, N, nb)
    C_t, c_idx, C_d = _rand_operand(rng, M, N, nb)
    alpha, beta = float(rng.standard_normal()), float(rng.choice([0.0, 1.0, -0.5]))
    ref = alpha * torch.einsum("bmk,bkn->bmn", A_d(), B_d()) + beta * C_d()
    tile = int(rng.choice([0, 1, 2, 3, 5, 6, 7]))
    splitk = int(rng.choice([0, 0, 2, 3])) if K >= 128 else 0
    ops.gemm_ex(M, N, K, nb, A_t, B_t, C_t, a_idx + b_idx + c_idx, alpha=alpha, beta=beta, force_tile=tile, force_splitk=splitk)
    assert rel(C_d(), ref) < 1e-12


@pytest.mark.parametrize("cond", [1e1, 1e3, 1e5, 1e7, 1e12])
@pytest.mark.parametrize("m,q", [(8192, 96), (2048, 96)])
def test_orthonormalize_fast_path_matches_householder_quality(cond, m, q, monkeypatch):
    """K4's CholeskyQR2 fast path (taken per 32-column panel when every Cholesky pivot keeps > 1e-11 of its diagonal) against the
    Householder TSQR path on tall matrices of prescribed condition number: orthogonality at machine precision on both, and the
    basis must capture every left singular direction u_k of Y as well as the Householder basis does (the residual of u_k outside
    span(Q), weighted by s_k / s_0 -- what the rSVD's spectrum sees -- stays at 1e-14)."""
    g = torch.Generator().manual_seed(17)
    U = torch.linalg.qr(torch.randn(m, q, dtype=torch.float64, generator=g)).Q.cuda()
    V = torch.linalg.qr(torch.randn(q, q, dtype=torch.float64, generator=g)).Q.cuda()
    s = torch.logspace(0, -torch.log10(torch.tensor(cond)).item(), q, dtype=torch.float64, device="cuda")
    Y = (U * s) @ V.T
    res = {}
    for name, flag in (("fast", "0"), ("householder", "1")):
        monkeypatch.setenv("ACETN_B200_TSQR_HOUSEHOLDER", flag)
        Q = ops.orthonormalize(Y.clone())
        orth = float((Q.T @ Q - torch.eye(q, dtype=torch.float64, device="cuda")).abs().max())
        resid = (U - Q @ (Q.T @ U)).norm(dim=0) * s / s[0]            # weighted residual per singular direction
        res[name] = (orth, float(resid.max()))
    print(f"cond {cond:.0e}: fast (orth, weighted residual) = {res['fast']}, householder = {res['householder']}")
    for name in res:
        assert res[name][0] < 5e-14, (name, res[name])
        assert res[name][1] < 1e-13, (name, res[name])
