// Declarations of the non-GEMM kernels' host launchers (K3 helpers, K4 TSQR, K5 Jacobi).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace ab200 {

// ---- K3: reductions / scaling / gathers (reduce.cu) -----------------------------------------------------
// *dst = max(*dst, max|x|) ; dst must be zeroed by the caller (or hold a previous non-negative value)
int absmax_launch(const double* x, size_t n, double* dst, cudaStream_t s);
// x[i] *= 1 / *scalar   (no-op when *scalar == 0)
int scale_inv_launch(double* x, size_t n, const double* scalar, cudaStream_t s);
// x /= ||x||_F ; deterministic two-stage reduction; scratch needs frob_scratch_doubles() doubles
size_t frob_scratch_doubles();
int frob_normalize_launch(double* x, size_t n, double* scratch, cudaStream_t s);
// dst (contiguous, dims d0..d4 row-major) = src[i0*s0 + ... + i4*s4]
int gather5_launch(double* dst, const double* src, const int64_t dims[5], const int64_t strides[5], cudaStream_t s);
// dst (contiguous over dims) = src gathered with arbitrary strides, up to 8 dims (the 'transpose' T of TTGT)
int gather_nd_launch(double* dst, const double* src, int nd, const int64_t* dims, const int64_t* strides, cudaStream_t s);
// dst[r*ldd + c] = src[r*lds + c] * w[c] / *div   for c < ncols   (w, div optional)
int scale_cols_launch(double* dst, int64_t ldd, const double* src, int64_t lds, const double* w, const double* div,
                      int64_t nrows, int ncols, cudaStream_t s);
int copy2d_launch(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t nrows, int ncols, cudaStream_t s);
// sum_i a[i]*b[i] (deterministic) -> *dst ; scratch as frob
int dot_launch(const double* a, const double* b, size_t n, double* dst, double* scratch, cudaStream_t s);

// ---- K4: tall-skinny orthonormalisation (tsqr.cu) -----------------------------------------------------------
// Y (m x q, row-major, leading dimension ld) is overwritten by an orthonormal basis of its range
// (block classical Gram-Schmidt with re-orthogonalisation; panels factored by Householder TSQR).
size_t orthonormalize_workspace_bytes(int64_t m, int q);
int orthonormalize_launch(double* Y, int64_t m, int q, int64_t ld, void* ws, size_t ws_bytes, cudaStream_t s);

// ---- K5: one-sided Jacobi SVD of a small square core (jacobi.cu) -------------------------------------------
// X (n x n row-major, ld = n) = R.  Finds orthogonal J with J*X = diag(S)*W (rows of W orthonormal).
// Outputs, sorted by descending S: S[n]; Wt[i][:] = i-th row of W ; Jt[i][:] = matching row of J.
//   => R = Jt^T diag(S) Wt.
// count (device int32[2]): [0] = min(chi, #{S/S[0] > cutoff}), [1] = sweeps used.
size_t jacobi_workspace_bytes(int n);
int jacobi_svd_launch(const double* R, int n, double* S, double* Wt, double* Jt, int chi, double cutoff, int* count,
                      void* ws, size_t ws_bytes, cudaStream_t s);
// w[z] = 1/sqrt(S[z]/S[0]) for z < ncols
int inv_sqrt_weights_launch(const double* S, double* w, int ncols, cudaStream_t s);

}  // namespace ab200
