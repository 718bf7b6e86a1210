/*
 * acetn_b200.h -- C ABI of libacetn_b200.so: the B200 (sm_100a) implementation of Ace-TN's CTMRG hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one reference call site
 * (cited per function; paths relative to the ace-tn repository root) and is what a ctypes / pybind / cgo stub
 * on the reference side binds.
 *
 * Conventions
 *   - all tensors are FP64, row-major contiguous device memory unless a stride array is taken;
 *   - one int64 extent per tensor leg (chi legs may differ between tensors);
 *   - the caller owns all memory, including the scratch workspace: query *_workspace_bytes(), allocate
 *     (torch.empty on the host side), pass pointer + size;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point synchronises the device
 *     or allocates device memory;
 *   - every function returns 0 on success; on failure a non-zero ACETN_B200_ERR_* code and
 *     acetn_b200_last_error() holds a message (thread local).  Nothing throws across the boundary.
 *   - there is NO CPU fallback: without a CUDA device the compute entry points fail with ACETN_B200_ERR_CUDA.
 *
 * Index names follow the reference's einsum strings so the two can be read side by side.
 * Bond direction k: 0=left 1=up 2=right 3=down; A[l,u,r,d,p]; C[k] (chi,chi); E[k] (chi,chi,D,D).
 */
#ifndef ACETN_B200_H
#define ACETN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACETN_B200_OK 0
#define ACETN_B200_ERR_INVALID 1
#define ACETN_B200_ERR_WORKSPACE 2
#define ACETN_B200_ERR_CUDA 3
#define ACETN_B200_ERR_UNSUPPORTED 4

/* ---- library lifetime (replaces the function-local static cuTENSOR/cuSOLVER handle singletons,
 *      csrc/linalg/contraction.h:76-94, csrc/linalg/cholesky_solve.h:12-30) ------------------------------ */
int acetn_b200_init(int device);
int acetn_b200_destroy(void);
const char* acetn_b200_last_error(void);
const char* acetn_b200_version(void);
/* number of kernel launches issued by this library in the calling process since the last reset */
int64_t acetn_b200_launch_count(void);
void acetn_b200_reset_launch_count(void);

/* ---- K1: general two-level-index FP64 DMMA GEMM (the primitive behind every torch.einsum / @ of the path;
 *      replaces cuTENSOR's TTGT contraction, csrc/linalg/contraction.h:96-309) ---------------------------------
 * C(m,n)[b] = alpha * sum_k A(m,k)[b] B(k,n)[b] + beta * C(m,n)[b]
 * Each index group g of each operand is described by 3 int64: {div, s_hi, s_lo}:
 *      offset(i) = div ? (i / div) * s_hi + (i % div) * s_lo : i * s_lo.
 * idx = 27 int64: A{row=m, col=k, batch}, B{row=k, col=n, batch}, C{m, n, batch}. */
size_t acetn_b200_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch, const int64_t* idx,
                                       int force_tile, int force_splitk);
int acetn_b200_gemm(int64_t M, int64_t N, int64_t K, int64_t batch, const double* A, const double* B, double* C,
                    const int64_t* idx, double alpha, double beta, int force_tile, int force_splitk, void* ws,
                    size_t ws_bytes, void* stream);

/* ---- quarter tensor: ProjectorCalculator.make_quarter_tensor, acetn/renormalization/projectors.py:36-60 -----
 *   Q[(c,r,R),(e,d,D)] = sum C[a,b] E2[b,c,u,U] E1[e,a,l,L] conj(A)[L,U,R,D,P] A[l,u,r,d,P],  then Q /= max|Q|
 *   C  (chi_a, chi_b) = site.C[k];  E2 (chi_b, chi_c, D, D) = site.E[k];  E1 (chi_e, chi_a, D, D) = site.E[(3+k)%4]
 *   A  = site.bond_permute(k), a strided view: a_strides[5] in elements for legs (l,u,r,d,p) of the view.
 *   Q out: (chi_c*D*D) x (chi_e*D*D) row-major.  normalize != 0 applies the max-abs division (projectors.py:59).
 *   absmax_out (device double, may be NULL): receives max|Q| of the un-normalised tensor. */
size_t acetn_b200_quarter_tensor_workspace_bytes(int64_t chi_a, int64_t chi_b, int64_t chi_c, int64_t chi_e,
                                                 int64_t D, int64_t d);
int acetn_b200_quarter_tensor(const double* C, const double* E2, const double* E1, const double* A,
                              const int64_t* a_strides, int64_t chi_a, int64_t chi_b, int64_t chi_c, int64_t chi_e,
                              int64_t D, int64_t d, int normalize, double* Q, double* absmax_out, void* ws, size_t ws_bytes,
                              void* stream);

/* Quarter tensor + its K7 residue encoding in one call (see "K7" below): Q is left un-normalised (absmax_out carries max|Q|, which
 * acetn_b200_projectors_from_usv applies to the projectors), enc_storage (acetn_b200_i8_encoded_bytes(chi_c D^2, chi_e D^2) bytes)
 * receives the same encoding acetn_b200_i8_encode(Q) would produce.  For D = 8, d = 2 the producing kernel's epilogue delivers the
 * column exponents, so the encoding costs ONE pass over the 2 GiB tensor (8 B/element read + 16 B/element written). */
int acetn_b200_quarter_tensor_enc(const double* C, const double* E2, const double* E1, const double* A,
                                  const int64_t* a_strides, int64_t chi_a, int64_t chi_b, int64_t chi_c, int64_t chi_e,
                                  int64_t D, int64_t d, double* Q, double* absmax_out, void* enc_storage, size_t enc_bytes,
                                  void* ws, size_t ws_bytes, void* stream);

/* ---- randomized SVD family: acetn/linalg/{svd_lowrank,fused_matmul_svd_lowrank,fused_3matmul_svd_lowrank}.py
 *   rSVD of the product M_0 M_1 ... M_{nmat-1} (nmat = 1, 2 or 4) without forming it.  mats[i] is rows[i] x cols[i]
 *   row-major contiguous; Omega is (cols[nmat-1] x q), drawn by the caller with torch.randn exactly as the reference
 *   does (fused_matmul_svd_lowrank.py:32) so both consume identical test matrices.
 *   reorth_adjoint != 0 re-orthonormalises between the adjoint and forward halves of each power step
 *   (fused_3matmul_svd_lowrank.py:39-45).
 *   Outputs: U (rows[0] x q), S (q, descending), V (cols[nmat-1] x q)  [V not transposed, as the reference],
 *   info (device int32[2]): info[0] = min(chi, #{S/S[0] > cutoff})  (projectors.py:163-164), info[1] = Jacobi sweeps.
 *   Optional (may be NULL): U may be NULL when only AtQ/Wt are wanted; AtQ (cols[0] x q) receives M_0^T Q, the first
 *   product of the final adjoint pass, and Wt (q x q) the left singular vectors of the core as rows (U = Q Wt^T), so that
 *   proj1 = M_0^T conj(U) = AtQ Wt^T can be formed without another pass over M_0 (see projectors_from_usv). */
size_t acetn_b200_rsvd_workspace_bytes(int nmat, const int64_t* rows, const int64_t* cols, int64_t q);
int acetn_b200_rsvd(int nmat, const double* const* mats, const int64_t* rows, const int64_t* cols, const double* Omega,
                    int64_t q, int niter, int reorth_adjoint, int64_t chi, double cutoff, double* U, double* S,
                    double* V, int32_t* info, double* AtQ, double* Wt, void* ws, size_t ws_bytes, void* stream);

/* ---- orthonormal basis of a tall matrix (torch.linalg.qr(Y).Q, fused_matmul_svd_lowrank.py:38,43), in place */
size_t acetn_b200_orthonormalize_workspace_bytes(int64_t m, int64_t q);
int acetn_b200_orthonormalize(double* Y, int64_t m, int64_t q, int64_t ld, void* ws, size_t ws_bytes, void* stream);

/* ---- SVD of a small square matrix R (q x q): R = Jt^T diag(S) Wt  (torch.linalg.svd core) -------------------- */
size_t acetn_b200_jacobi_svd_workspace_bytes(int64_t q);
int acetn_b200_jacobi_svd(const double* R, int64_t q, double* S, double* Wt, double* Jt, int64_t chi, double cutoff,
                          int32_t* info, void* ws, size_t ws_bytes, void* stream);

/* ---- projector formation: projectors.py:166-173 (half-system) ---------------------------------------------------
 *   proj1[(e,d,D), z] = sum_x Q1[x,(e,d,D)] U[x,z] w[z],  proj2[(c,u,U), z] = sum_y Q4[(c,u,U),y] V[y,z] w[z],
 *   w[z] = (S[z]/S[0])^-1/2, z < keep.   Q1: m1 x n1, Q4: m4 x n4, U: m1 x ldu, V: n4 x ldv.
 *   proj1 out: n1 x keep, proj2 out: m4 x keep.
 *   qmax1 / qmax4 (device doubles, may be NULL): when Q1 / Q4 were produced with normalize = 0, passing their max|Q|
 *   here divides proj1 / proj2 by it, which reproduces the reference's normalised-Q projectors exactly while saving the
 *   two HBM passes of the division over the 2 GiB tensors.
 *   AtQ (n1 x q) and Wt (q x q) from acetn_b200_rsvd (may both be NULL): when given, proj1 = AtQ (Wt^T diag(w)) is formed
 *   by one small GEMM and Q1 / U are not read (saves 2 m1 n1 keep flop = one of the 14 large GEMMs of a site-move). */
size_t acetn_b200_projectors_workspace_bytes(int64_t m1, int64_t n1, int64_t m4, int64_t n4, int64_t keep);
int acetn_b200_projectors_from_usv(const double* Q1, int64_t m1, int64_t n1, const double* Q4, int64_t m4, int64_t n4,
                                   const double* U, int64_t ldu, const double* V, int64_t ldv, const double* S,
                                   int64_t keep, const double* qmax1, const double* qmax4, const double* AtQ,
                                   const double* Wt, int64_t q, double* proj1, double* proj2, void* ws, size_t ws_bytes,
                                   void* stream);

/* ---- absorption: DirectionalMover.renormalize_cj1/cj2/ej, acetn/renormalization/directional_mover.py:306-366 ---
 *   corner1: out[a,x] = sum ei[a,b,l,L] ci[b,c] proj[c,l,L,x] / ||.||     ci (chi_b,chi_c) ei (chi_a,chi_b,D,D) proj (chi_c,D,D,chi_x)
 *   corner2: out[x,c] = sum proj[a,r,R,x] ci[a,b] ei[b,c,r,R] / ||.||     ci (chi_a,chi_b) ei (chi_b,chi_c,D,D) proj (chi_a,D,D,chi_x)
 *   edge   : out[y,x,r,R] = sum ei[a,b,l,L] proj1[b,u,U,x] conj(A)[L,U,R,D,P] A[l,u,r,d,P] proj2[a,d,D,y] / ||.||
 *            ei (chi_a,chi_b,D,D), proj1 (chi_b,D,D,chi_x), proj2 (chi_a,D,D,chi_y), A = bond_permute(k) view.
 *            normalize = 0 returns the un-normalised sum: the contraction is linear in the `a` leg shared by ei and
 *            proj2, so ranks holding a-blocks of both can all-reduce their partial results and normalise afterwards. */
size_t acetn_b200_absorb_corner_workspace_bytes(int64_t chi_a, int64_t chi_b, int64_t chi_c, int64_t chi_x, int64_t D);
int acetn_b200_absorb_corner1(const double* ci, const double* ei, const double* proj, int64_t chi_a, int64_t chi_b,
                              int64_t chi_c, int64_t chi_x, int64_t D, double* out, void* ws, size_t ws_bytes,
                              void* stream);
int acetn_b200_absorb_corner2(const double* ci, const double* ei, const double* proj, int64_t chi_a, int64_t chi_b,
                              int64_t chi_c, int64_t chi_x, int64_t D, double* out, void* ws, size_t ws_bytes,
                              void* stream);
size_t acetn_b200_absorb_edge_workspace_bytes(int64_t chi_a, int64_t chi_b, int64_t chi_x, int64_t chi_y, int64_t D,
                                              int64_t d);
int acetn_b200_absorb_edge(const double* ei, const double* A, const int64_t* a_strides, const double* proj2,
                           const double* proj1, int64_t chi_a, int64_t chi_b, int64_t chi_x, int64_t chi_y, int64_t D,
                           int64_t d, int normalize, double* out, void* ws, size_t ws_bytes, void* stream);

/* The edge absorption in two stages (same kernels and launch parameters as acetn_b200_absorb_edge => bit-identical): `begin` needs
 * only proj1 of the neighbouring task and produces T3 (chi_a * chi_x * D^4 doubles, caller-owned); `finish` contracts it with
 * this task's proj2 (directional_mover.py:362-366).  Lets a scheduler start most of the absorption before the last projector pair
 * of a move exists. */
size_t acetn_b200_absorb_edge_begin_workspace_bytes(int64_t chi_a, int64_t chi_b, int64_t chi_x, int64_t D, int64_t d);
int acetn_b200_absorb_edge_begin(const double* ei, const double* A, const int64_t* a_strides, const double* proj1, int64_t chi_a,
                                 int64_t chi_b, int64_t chi_x, int64_t D, int64_t d, double* T3, void* ws, size_t ws_bytes, void* stream);
size_t acetn_b200_absorb_edge_finish_workspace_bytes(int64_t chi_a, int64_t chi_x, int64_t chi_y, int64_t D);
int acetn_b200_absorb_edge_finish(const double* proj2, const double* T3, int64_t chi_a, int64_t chi_x, int64_t chi_y, int64_t D,
                                  int normalize, double* out, void* ws, size_t ws_bytes, void* stream);

/* ---- double-layer site absorption on its own (K2; the `cuUelL,LURDP->cuelRDP` + `lurdp,cuelRDp->crRedD` pair of
 *      projectors.py:54-55 and the matching pair of directional_mover.py:363-364), exposed for tests/benchmarks.
 *   in : X[b0, (i0,I0), b1, (i1,I1)] with block (b0,b1) at  b0*in_s0 + b1*in_s1 and element strides in_es[4] for
 *        (i0,I0,i1,I1) = the ket/bra legs contracted with the first two legs (l,u) / (L,U) of A in `order`:
 *        order 0: (i0,I0)=(u,U) , (i1,I1)=(l,L)   (quarter tensor);  order 1: (i0,I0)=(l,L), (i1,I1)=(u,U)   (edge)
 *   out: Y[block][(r,R) , (d,D)] with element strides out_es[4] for (r,R,d,D) and block strides out_s0,out_s1. */
size_t acetn_b200_double_layer_workspace_bytes(int64_t n0, int64_t n1, int64_t D, int64_t d);
int acetn_b200_double_layer(const double* X, int64_t n0, int64_t n1, int64_t in_s0, int64_t in_s1,
                            const int64_t* in_es, int order, const double* A, const int64_t* a_strides, int64_t D,
                            int64_t d, double* Y, int64_t out_s0, int64_t out_s1, const int64_t* out_es, void* ws,
                            size_t ws_bytes, void* stream);

/* ---- K7: FP64-accurate (normwise, P-bit fixed point per row/column scale) "big x thin" products on the INT8 tensor cores (tcgen05.mma kind::i8 + TMEM + TMA) --------------
 *   The thin products of the rSVD chain and of the projector formation (fused_matmul_svd_lowrank.py:33-46,
 *   projectors.py:166-172; `A @ (B @ omega)` etc. in the reference) multiply the same quarter tensors 13 times per
 *   site-move.  i8_encode turns a matrix Q (rows x cols) once into 16 planes of int8 residues (two-sided power-of-two
 *   scaling to P-bit integers, P = 54 up to a contraction length of 16384, then mod 16 coprime moduli <= 256); i8_matmul evaluates
 *   out = Q Y  (adjoint = 0, Y: cols x q)  or  out = Q^T Y  (adjoint = 1, Y: rows x q)  in integer arithmetic (one INT8
 *   tensor-core GEMM per modulus, Chinese-remainder reconstruction in 128-bit integers): the product of the ROUNDED operands is
 *   exact; the only rounding is the fixed-point conversion of the operands, 2^-P relative to each entry's row x column scale.  The
 *   result is therefore NORMWISE FP64-accurate (per row / column scale), not componentwise: entries more than 2^-P below their
 *   row-and-column maximum are flushed (acetn_b200/csrc/i8crt.cu).  q <= 272; 128 <= rows, cols <= 65535.
 *   storage (device, acetn_b200_i8_encoded_bytes) is caller-owned and opaque. */
int acetn_b200_i8_supported(int64_t rows, int64_t cols, int64_t q);
size_t acetn_b200_i8_encoded_bytes(int64_t rows, int64_t cols);
int acetn_b200_i8_encode(const double* Q, int64_t rows, int64_t cols, int64_t ldq, void* storage, size_t storage_bytes,
                         void* stream);
size_t acetn_b200_i8_matmul_workspace_bytes(int64_t rows, int64_t cols, int64_t q);
int acetn_b200_i8_matmul(const void* storage, int64_t rows, int64_t cols, int adjoint, const double* Y, int64_t q,
                         int64_t ldy, double* out, int64_t ldo, void* ws, size_t ws_bytes, void* stream);

/* rSVD / projector formation with K7: encs[i] (or enc1 / enc4) is the acetn_b200_i8_encode storage of factor i, or NULL to
 * keep that factor on the FP64 DMMA path; mats[i] (Q1 / Q4) may be NULL where an encoding is given.  use_enc[i] != 0 in the
 * workspace queries marks the encoded factors.  Everything else as acetn_b200_rsvd / acetn_b200_projectors_from_usv. */
size_t acetn_b200_rsvd_enc_workspace_bytes(int nmat, const int64_t* rows, const int64_t* cols, int64_t q,
                                           const int32_t* use_enc);
int acetn_b200_rsvd_enc(int nmat, const double* const* mats, const void* const* encs, const int64_t* rows,
                        const int64_t* cols, const double* Omega, int64_t q, int niter, int reorth_adjoint, int64_t chi,
                        double cutoff, double* U, double* S, double* V, int32_t* info, double* AtQ, double* Wt, void* ws,
                        size_t ws_bytes, void* stream);
size_t acetn_b200_projectors_enc_workspace_bytes(int64_t m1, int64_t n1, int64_t m4, int64_t n4, int64_t keep,
                                                 int use_enc);
int acetn_b200_projectors_from_usv_enc(const double* Q1, const void* enc1, int64_t m1, int64_t n1, const double* Q4,
                                       const void* enc4, int64_t m4, int64_t n4, const double* U, int64_t ldu,
                                       const double* V, int64_t ldv, const double* S, int64_t keep, const double* qmax1,
                                       const double* qmax4, const double* AtQ, const double* Wt, int64_t q,
                                       double* proj1, double* proj2, void* ws, size_t ws_bytes, void* stream);

/* ---- bench-only: launches a register-resident DMMA.8x8x4 loop on every SM and returns its flop count; timed by the
 *      caller with CUDA events it gives the live FP64 tensor-pipe roof used as the roofline denominator ---------- */
double acetn_b200_fp64_peak_probe(void* scratch, int iters, void* stream);

/* ---- ALS inner solver of the full update: ALSSolver.solve_torch (acetn/evolution/als_solver.py:55-82) = als_solve
 *      of the reference's cuTENSOR extension (csrc/evolution/als_solve.cpp:107-137).
 *   method 0 = "cholesky" (als_solver.py:218-225, als_solve.cpp:25-46): (R + R^T)/2 + epsilon max|R| I solved by Cholesky;
 *   method 1 = "pinv" (als_solver.py:226-228, als_solve.cpp:47-50): pinv((R + R^T)/2, hermitian, rcond = epsilon) S, the symmetric
 *   eigen-decomposition by two-sided Jacobi inside the kernel (needs nD*bD <= 159).
 *   a1r, a2r (nD,bD,pD): in = initial guess (als_solver.py:117-146), out = result.  n12 (nD^4) [y,x,Y,X],
 *   n12g (nD,nD,pD,pD) [Y,X,p,q], a12g (nD,nD,pD,pD) [y,x,p,q].  The whole loop (<= niter iterations, stop when the
 *   relative cost change < tol and i > 1) runs in one cooperative kernel; info (device int32[2]) = {iterations run,
 *   number of non-positive Cholesky pivots met (0 = clean)}. */
size_t acetn_b200_als_workspace_bytes(int64_t nD, int64_t bD, int64_t pD);
int acetn_b200_als_solve(double* a1r, double* a2r, const double* n12g, const double* n12, const double* a12g, int64_t nD,
                         int64_t bD, int64_t pD, int64_t niter, double tol, double epsilon, int64_t method, int32_t* info,
                         void* ws, size_t ws_bytes, void* stream);

/* ---- environment contractions of `measure` and of the full update (SURVEY.md 8b minimum set; acetn_b200/csrc/environment.cu) --------
 *   Every step is a K1 launch whose index descriptors absorb the leg permutations of the reference's einsum chain.
 *   chi: the chi extents of the boundary tensors, 2 per tensor in ARGUMENT ORDER (chi legs may differ); the
 *   entry points check that contracted legs agree.  Site-tensor arguments are strided views (5 element strides, legs l,u,r,d,p).
 *
 *   site_rdm : RDM.build_site_rdm, acetn/measurement/rdm.py:35-67.   c1..c4 = site.C[0..3], e1..e4 = site.E[0..3], A = site['A'];
 *              chi = 16 extents (c1,c2,c3,c4,e1,e2,e3,e4);  rho out: (d,d) [bra P, ket p].
 *   bond_rdm : RDM.build_bond_rdm + build_bond_rdm_core_blocked, rdm.py:69-154, for the bond (s1, s2, k):
 *              c12 = s1.C[(k+1)%4], e12 = s1.E[(k+1)%4], e11 = s1.E[k], c13 = s1.C[(k+2)%4], e13 = s1.E[(k+2)%4], a1 = s1.bond_permute(k),
 *              c21 = s2.C[k], e21 = s2.E[k], e24 = s2.E[(k+3)%4], c24 = s2.C[(k+3)%4], e23 = s2.E[(k+2)%4], a2 = s2.bond_permute(k);
 *              chi = 20 extents (c12,e12,e11,c13,e13,c21,e21,e24,c24,e23);  rho out: (d,d,d,d) [P,Q,p,q].
 *   norm_tensor : build_norm_tensor, acetn/evolution/full_update.py:163-227: the same ten boundary tensors with the QR-reduced site
 *              factors a1q, a2q (D,D,D,nD) contiguous;  n12 out: (nD,nD,nD,nD) [y,x,Y,X]. */
size_t acetn_b200_site_rdm_workspace_bytes(const int64_t* chi, int64_t D, int64_t d);
int acetn_b200_site_rdm(const double* c1, const double* c2, const double* c3, const double* c4, const double* e1, const double* e2,
                        const double* e3, const double* e4, const double* A, const int64_t* a_strides, const int64_t* chi, int64_t D,
                        int64_t d, double* rho, void* ws, size_t ws_bytes, void* stream);
size_t acetn_b200_bond_rdm_workspace_bytes(const int64_t* chi, int64_t D, int64_t d);
int acetn_b200_bond_rdm(const double* c12, const double* e12, const double* e11, const double* c13, const double* e13, const double* a1,
                        const int64_t* a1_strides, const double* c21, const double* e21, const double* e24, const double* c24,
                        const double* e23, const double* a2, const int64_t* a2_strides, const int64_t* chi, int64_t D, int64_t d,
                        double* rho, void* ws, size_t ws_bytes, void* stream);
size_t acetn_b200_norm_tensor_workspace_bytes(const int64_t* chi, int64_t D, int64_t nD);
int acetn_b200_norm_tensor(const double* c12, const double* e12, const double* e11, const double* c13, const double* e13, const double* a1q,
                           const double* c21, const double* e21, const double* e24, const double* c24, const double* e23, const double* a2q,
                           const int64_t* chi, int64_t D, int64_t nD, double* n12, void* ws, size_t ws_bytes, void* stream);

/* ---- generic pairwise contraction support (measure / norm-tensor paths: rdm.py:35-154, full_update.py:209-227): the
 *      transpose step of a transpose-transpose-GEMM-transpose contraction (what cuTENSOR's TTGT plan does in the
 *      reference's extension, csrc/linalg/contraction.h:263).  dst is contiguous row-major over dims[0..nd);
 *      src is read at sum_i idx_i*strides[i] (elements).  nd <= 8. */
int acetn_b200_permute(double* dst, const double* src, int nd, const int64_t* dims, const int64_t* strides, void* stream);

/* ---- small helpers used by the host shim ------------------------------------------------------------------------ */
int acetn_b200_absmax(const double* x, int64_t n, double* out_scalar_zeroed, void* stream);
int acetn_b200_frob_normalize(double* x, int64_t n, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACETN_B200_H */
