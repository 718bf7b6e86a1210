"""The reference's own unit tests of the rSVD family (tests/unit/test_linalg.py of ace-tn, cited per test) re-stated against
`acetn_b200.linalg` on the GPU -- same cases, same shapes, same seeds, same tolerances -- plus a direct comparison with the UNMODIFIED
reference functions (oracle/_ref, same GPU, same seed => the same Gaussian test matrix): singular values to 1e-12, the rank-q
approximation U diag(S) V^T to 1e-10.  complex128 (test_linalg.py:264-307) is outside the FP64-real metric: the backend must refuse it
loudly instead of computing something else."""
import pytest
import torch

from oracle import vendor_ref

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from acetn_b200 import linalg

DEV = "cuda"


def _rand(*shape, seed=42):
    torch.manual_seed(seed)
    return torch.randn(*shape, dtype=torch.float64).to(DEV)


def _approx(U, S, V):
    return U @ torch.diag(S) @ V.T


def _rel(A, B):
    return float(torch.norm(A - B) / torch.norm(A))


# ------------------------------------------------------------------------------------------------ TestSvdLowrank (test_linalg.py:6-98)
class TestSvdLowrank:
    def test_output_shapes(self):                                   # :23-31
        A = _rand(100, 80)
        U, S, V = linalg.svd_lowrank(A, q=20, niter=2)
        assert U.shape == (100, 20) and S.shape == (20,) and V.shape == (80, 20)

    def test_reconstruction_error(self):                            # :33-43
        A = _rand(100, 80)
        assert _rel(A, _approx(*linalg.svd_lowrank(A, q=40, niter=2))) < 0.5

    def test_low_rank_exact_recovery(self):                         # :45-55
        torch.manual_seed(42)
        A = (torch.randn(100, 10, dtype=torch.float64) @ torch.randn(80, 10, dtype=torch.float64).T).to(DEV)
        assert _rel(A, _approx(*linalg.svd_lowrank(A, q=15, niter=3))) < 1e-10

    def test_singular_values_ordering(self):                        # :57-64
        _, S, _ = linalg.svd_lowrank(_rand(100, 80), q=20, niter=2)
        assert torch.all(S >= 0) and torch.all(S[:-1] >= S[1:])

    def test_orthogonality(self):                                   # :66-77
        U, _, V = linalg.svd_lowrank(_rand(100, 80), q=20, niter=2)
        eye = torch.eye(20, dtype=torch.float64, device=DEV)
        assert torch.allclose(U.T @ U, eye, atol=1e-10) and torch.allclose(V.T @ V, eye, atol=1e-10)

    def test_comparison_with_torch(self):                           # :79-98
        A = _rand(100, 80)
        torch.manual_seed(456)
        err1 = _rel(A, _approx(*linalg.svd_lowrank(A, q=20, niter=2)))
        torch.manual_seed(456)
        err2 = _rel(A, _approx(*torch.svd_lowrank(A, q=20, niter=2)))
        assert abs(err1 - err2) < 0.1


# ------------------------------------------------------------------------------------------------ TestFusedMatmulSvdLowrank (:101-171)
class TestFusedMatmulSvdLowrank:
    def test_output_shapes(self):                                   # :119-127
        A, B = _rand(60, 50), _rand(50, 70, seed=43)
        U, S, V = linalg.fused_matmul_svd_lowrank(A, B, q=20, niter=2)
        assert U.shape == (60, 20) and S.shape == (20,) and V.shape == (70, 20)

    def test_equivalence_to_explicit(self):                         # :129-149
        A, B = _rand(60, 50), _rand(50, 70, seed=43)
        C = A @ B
        torch.manual_seed(789)
        err1 = _rel(C, _approx(*linalg.svd_lowrank(C, q=25, niter=2)))
        torch.manual_seed(789)
        err2 = _rel(C, _approx(*linalg.fused_matmul_svd_lowrank(A, B, q=25, niter=2)))
        assert abs(err1 - err2) < 0.05

    def test_low_rank_exact_recovery(self):                         # :151-161
        A, B = _rand(60, 8), _rand(8, 70, seed=43)
        C = A @ B
        assert _rel(C, _approx(*linalg.fused_matmul_svd_lowrank(A, B, q=12, niter=3))) < 1e-10

    def test_reconstruction_quality(self):                          # :163-171
        A, B = _rand(60, 50), _rand(50, 70, seed=43)
        C = A @ B
        assert _rel(C, _approx(*linalg.fused_matmul_svd_lowrank(A, B, q=30, niter=2))) < 0.5


# ------------------------------------------------------------------------------------------------ TestFused3MatmulSvdLowrank (:174-261)
class TestFused3MatmulSvdLowrank:
    @staticmethod
    def _quad():
        return _rand(40, 35), _rand(35, 30, seed=43), _rand(30, 35, seed=44), _rand(35, 45, seed=45)

    def test_output_shapes(self):                                   # :199-207
        A, B, C, D = self._quad()
        U, S, V = linalg.fused_3matmul_svd_lowrank(A, B, C, D, q=20, niter=2)
        assert U.shape == (40, 20) and S.shape == (20,) and V.shape == (45, 20)

    def test_equivalence_to_explicit(self):                         # :209-229
        A, B, C, D = self._quad()
        F = A @ B @ C @ D
        torch.manual_seed(101)
        err1 = _rel(F, _approx(*linalg.svd_lowrank(F, q=20, niter=2)))
        torch.manual_seed(101)
        err2 = _rel(F, _approx(*linalg.fused_3matmul_svd_lowrank(A, B, C, D, q=20, niter=2)))
        assert abs(err1 - err2) < 0.05

    def test_low_rank_exact_recovery(self):                         # :231-241
        A, D = _rand(40, 6), _rand(6, 45, seed=43)
        I6 = torch.eye(6, dtype=torch.float64, device=DEV)
        F = A @ D
        assert _rel(F, _approx(*linalg.fused_3matmul_svd_lowrank(A, I6, I6.clone(), D, q=10, niter=3))) < 1e-10

    def test_reconstruction_quality(self):                          # :243-253
        A, B, C, D = self._quad()
        F = A @ B @ C @ D
        assert _rel(F, _approx(*linalg.fused_3matmul_svd_lowrank(A, B, C, D, q=25, niter=2))) < 0.5

    def test_singular_values_ordering(self):                        # :255-261
        A, B, C, D = self._quad()
        _, S, _ = linalg.fused_3matmul_svd_lowrank(A, B, C, D, q=20, niter=2)
        assert torch.all(S >= 0) and torch.all(S[:-1] >= S[1:])


# ------------------------------------------------------------------------------------------------ TestEdgeCases (:310-367)
class TestEdgeCases:
    def test_q_larger_than_dimensions(self):                        # :313-324
        U, S, V = linalg.svd_lowrank(_rand(20, 15), q=50, niter=2)
        assert U.shape[1] <= 15 and S.shape[0] <= 15 and V.shape[1] <= 15

    def test_square_matrix(self):                                   # :326-335
        A = _rand(30, 30)
        assert _rel(A, _approx(*linalg.svd_lowrank(A, q=15, niter=2))) < 0.6

    def test_tall_matrix(self):                                     # :337-345
        U, _, V = linalg.svd_lowrank(_rand(100, 20), q=15, niter=2)
        assert U.shape == (100, 15) and V.shape == (20, 15)

    def test_wide_matrix(self):                                     # :347-355
        U, _, V = linalg.svd_lowrank(_rand(20, 100), q=15, niter=2)
        assert U.shape == (20, 15) and V.shape == (100, 15)

    def test_niter_zero(self):                                      # :357-367
        A = _rand(50, 40)
        assert _rel(A, _approx(*linalg.svd_lowrank(A, q=20, niter=0))) < 1.0


# ------------------------------------------------------------------------------------------------ TestComplexMatrices (:264-307)
def test_complex_input_is_refused():
    """complex128 is SURVEY 8f-4 (not built): the backend says so instead of silently dropping the imaginary part."""
    torch.manual_seed(42)
    A = torch.randn(50, 40, dtype=torch.complex128).to(DEV)
    with pytest.raises(RuntimeError, match="float64"):
        linalg.svd_lowrank(A, q=20, niter=2)


# ------------------------------------------------------------------------------------------------ against the reference itself
@pytest.mark.skipif(vendor_ref.import_path() is None, reason="reference package not available (oracle/_ref)")
@pytest.mark.parametrize("case", ["svd", "fused2", "fused4"])
@pytest.mark.parametrize("graded", [False, True])
def test_same_seed_matches_the_reference_functions(case, graded):
    """acetn.linalg.* (unmodified, cuBLAS / cuSOLVER through torch) and acetn_b200.linalg.* on the same GPU with the same seed draw the
    same Omega (fused_matmul_svd_lowrank.py:32); the truncated factorisations must then agree as far as the problem is conditioned:
    singular values to 1e-12 of s0, the rank-q approximation to 1e-10 (the individual vectors only up to sign / rotations inside
    clusters, so they are compared through U diag(S) V^T)."""
    vendor_ref.enable()
    from acetn.linalg import fused_3matmul_svd_lowrank, fused_matmul_svd_lowrank, svd_lowrank
    scale = torch.logspace(0, -8, 90, dtype=torch.float64, device=DEV) if graded else torch.ones(90, dtype=torch.float64, device=DEV)
    if case == "svd":
        mats = [_rand(120, 90, seed=7) * scale]
        ref_fn, our_fn = svd_lowrank, linalg.svd_lowrank
    elif case == "fused2":
        mats = [_rand(120, 100, seed=7), _rand(100, 90, seed=8) * scale]
        ref_fn, our_fn = fused_matmul_svd_lowrank, linalg.fused_matmul_svd_lowrank
    else:
        mats = [_rand(120, 100, seed=7), _rand(100, 80, seed=8), _rand(80, 100, seed=9), _rand(100, 90, seed=10) * scale]
        ref_fn, our_fn = fused_3matmul_svd_lowrank, linalg.fused_3matmul_svd_lowrank
    q = 30
    torch.manual_seed(2024)
    Ur, Sr, Vr = ref_fn(*mats, q=q, niter=2)
    torch.manual_seed(2024)
    U, S, V = our_fn(*mats, q=q, niter=2)
    assert float((S - Sr).abs().max() / Sr[0]) < 1e-12
    F = _approx(Ur, Sr, Vr)
    assert float(torch.norm(_approx(U, S, V) - F) / torch.norm(F)) < 1e-10
