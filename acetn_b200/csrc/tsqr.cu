// K4: orthonormal basis of a tall-skinny matrix (the QR steps of the randomized SVD, reference
// acetn/linalg/fused_matmul_svd_lowrank.py:38,43; only the Q factor is ever used there).
//
// Block classical Gram-Schmidt with re-orthogonalisation (BCGS2) over 32-column panels; the inter-panel projections are two K1
// DGEMMs each.  A panel P (m x b) is orthonormalised in one of two ways, decided ON THE DEVICE (no host read):
//
//   fast path, CholeskyQR2:  twice { G = P^T P (partial Gram per 256-row chunk, fixed-order reduction) ; G = R^T R ; P <- P R^-1 }
//     5 short launches.  Taken only when it is safe: every Cholesky pivot of the first pass must keep more than 1e-11 of its
//     diagonal entry (cond(P) below ~3e5, so that the second pass restores orthogonality to machine precision), no zero columns.
//
//   fallback, Householder TSQR with explicit chunk factors (the only path of round 1, 4x the latency): every 256-row chunk is
//     factored X = Q0 R0 by Householder reflections and Q0 formed from the reflectors, the stack of R factors is factored the
//     same way until one chunk is left, then chunk c of level l <- Q0_c * (rows [c b, c b + b) of the final Q of level l + 1).
//     Keeps ||Q^T Q - I|| at machine precision for ANY conditioning of Y -- the power-iterated Y of converged physical states reaches
//     cond 1e20 (SURVEY.md section 7 hard part 2), where CholeskyQR would lose the trailing directions; such panels fail the pivot
//     test and take this path, while well-conditioned panels (the first and last orthonormalisation of a projector, random benchmark
//     tensors) take the short one.
//
// Two launch forms of the same algorithm: one kernel per stage (~25 launches per panel; any size), and -- for m <= 4096, q <= 130,
// where those launches are pure latency -- ortho_fused_kernel: everything in ONE launch, one CTA per 256-row chunk, the stage
// boundaries replaced by __syncthreads / a thread-block-cluster barrier / a cooperative grid barrier.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace ab200 {

constexpr int TS_CH = 256;   // rows per chunk (one thread per row)
constexpr int TS_PB = 32;    // panel width
constexpr int TS_NW = TS_CH / 32;

// Thread layout of both kernels: lane = column of the panel (b <= 32), warp w owns rows [32w, 32w+32) of the chunk;
// every thread keeps its 32 rows of its column in registers, so a Householder step is 32 FMAs for the dot, an 8-way
// partial sum through shared memory and 32 FMAs for the update -- no warp shuffles on the critical path.

// Factor one chunk: P[r0:r0+rows, 0:b] = H_0 ... H_{b-1} [R; 0].  Reflectors overwrite the strictly lower part of P.
//
// Step j works on the RAW pivot column (as it stands after the previous updates), which its owner lanes (lane j of every warp)
// dumped to shared memory at the end of step j-1 together with their partial sums of squares:
//   v = [0 .. 0, 1, x_{j+1..} / (alpha - beta)],  H_j = I - tau v v^T   (LAPACK dlarfg)
//   x_c <- x_c - vraw (g * vraw^T x_c),  vraw = (alpha - beta) v,  g = tau / (alpha - beta)^2 = 1 / (nrm (nrm + |alpha|))
// so nobody has to scale the pivot column inside the loop (the 1/(alpha - beta) factors and the diagonal beta are applied when
// the reflectors are written out), every warp executes the same ~200 instructions per step, and a step needs two barriers.
struct FactorSmem {
    double col_s[TS_CH];            // raw pivot column of the current step
    double part[TS_NW][TS_PB];      // per-warp partial dots
    double partn[TS_NW];            // per-warp partial sums of squares of the pivot column below the diagonal
    double tau_s[TS_PB], beta_s[TS_PB], scale_s[TS_PB];
};

__device__ __forceinline__ void tsqr_factor_body(double* __restrict__ P, int64_t ld, int64_t nrows, int b, double* __restrict__ Rstack,
                                                 double* __restrict__ tau_out, FactorSmem& fs, int chunk) {
    double (&col_s)[TS_CH] = fs.col_s;
    double (&part)[TS_NW][TS_PB] = fs.part;
    double (&partn)[TS_NW] = fs.partn;
    double (&tau_s)[TS_PB] = fs.tau_s;
    double (&beta_s)[TS_PB] = fs.beta_s;
    double (&scale_s)[TS_PB] = fs.scale_s;
    const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
    const int64_t r0 = (int64_t)chunk * TS_CH;
    const int rows = (int)((nrows - r0) < TS_CH ? (nrows - r0) : TS_CH);
    const int nref = b < rows ? b : rows;
    const int rbase = 32 * w;

    double x[32];
#pragma unroll
    for (int i = 0; i < 32; i++) {
        int r = rbase + i;
        x[i] = (r < rows && c < b) ? P[(r0 + r) * ld + c] : 0.0;
    }
    if (tid < TS_PB) { tau_s[tid] = 0.0; beta_s[tid] = 0.0; scale_s[tid] = 1.0; }

    // sum of 32 products with 8 independent chains
#define TS_DOT32(res, EXPR)                                                     \
    {                                                                           \
        double _p[8];                                                           \
        _Pragma("unroll") for (int _k = 0; _k < 8; _k++) { int i = _k; _p[_k] = (EXPR); } \
        _Pragma("unroll") for (int _q = 1; _q < 4; _q++)                        \
            _Pragma("unroll") for (int _k = 0; _k < 8; _k++) { int i = _q * 8 + _k; _p[_k] += (EXPR); } \
        res = ((_p[0] + _p[1]) + (_p[2] + _p[3])) + ((_p[4] + _p[5]) + (_p[6] + _p[7])); \
    }
    // owner lanes of column jj: dump the raw column and the partial sum of squares of its rows below the diagonal
#define TS_PUBLISH(jj)                                                          \
    {                                                                           \
        double2* _d = reinterpret_cast<double2*>(col_s + rbase);                \
        _Pragma("unroll") for (int _i = 0; _i < 16; _i++) _d[_i] = make_double2(x[2 * _i], x[2 * _i + 1]); \
        double pn;                                                              \
        if (rbase > (jj)) { TS_DOT32(pn, x[i] * x[i]); }                        \
        else if (rbase + 31 <= (jj)) pn = 0.0;                                  \
        else { TS_DOT32(pn, (rbase + i > (jj)) ? x[i] * x[i] : 0.0); }          \
        partn[w] = pn;                                                          \
    }

    if (c == 0) TS_PUBLISH(0);
    __syncthreads();
    for (int j = 0; j < nref; j++) {
        // ---- reflector j (every thread computes the same scalars)
        double ss = ((partn[0] + partn[1]) + (partn[2] + partn[3])) + ((partn[4] + partn[5]) + (partn[6] + partn[7]));
        const double alpha = col_s[j];
        double g = 0.0, amb = 0.0;                 // g = tau / (alpha - beta)^2,  amb = alpha - beta
        if (ss != 0.0) {
            const double nrm = sqrt(alpha * alpha + ss);
            const double beta = alpha >= 0.0 ? -nrm : nrm;
            amb = alpha - beta;
            g = 1.0 / (nrm * (nrm + fabs(alpha)));
            // tau = (beta - alpha) / beta = g amb^2 and 1 / amb = -g beta: no further divisions (they would make warp 0 late at the barrier)
            if (tid == 0) { tau_s[j] = g * amb * amb; beta_s[j] = beta; scale_s[j] = -g * beta; }
        } else if (tid == 0) { tau_s[j] = 0.0; beta_s[j] = alpha; scale_s[j] = 0.0; }
        // ---- vraw . x_c over this warp's 32 rows (rows above the pivot do not take part; the pivot row carries alpha - beta)
        // (j < b <= 32: the pivot row always lies in warp 0, every other warp is entirely below it)
        double vv[32];
        double dsum;
        {
            const double2* v2 = reinterpret_cast<const double2*>(col_s + rbase);
#pragma unroll
            for (int i = 0; i < 16; i++) { double2 t = v2[i]; vv[2 * i] = t.x; vv[2 * i + 1] = t.y; }
            if (w == 0) {
#pragma unroll
                for (int i = 0; i < 32; i++) vv[i] = i > j ? vv[i] : (i == j ? amb : 0.0);
            }
            TS_DOT32(dsum, vv[i] * x[i]);
        }
        part[w][c] = dsum;
        __syncthreads();
        {
            double wc = ((part[0][c] + part[1][c]) + (part[2][c] + part[3][c])) + ((part[4][c] + part[5][c]) + (part[6][c] + part[7][c]));
            wc = (c > j && c < b) ? wc * g : 0.0;      // finished columns (c <= j) are left alone
            // the owner lanes of column j+1 publish it for the next step while the update is still being issued
            const bool pub = (c == j + 1) && (j + 1 < nref);
            double2* dst = reinterpret_cast<double2*>(col_s + rbase);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                x[2 * i] -= vv[2 * i] * wc;
                x[2 * i + 1] -= vv[2 * i + 1] * wc;
                if (pub) dst[i] = make_double2(x[2 * i], x[2 * i + 1]);
            }
            if (pub) {
                double pn;
                if (w > 0) { TS_DOT32(pn, x[i] * x[i]); }
                else { TS_DOT32(pn, (i > j + 1) ? x[i] * x[i] : 0.0); }
                partn[w] = pn;
            }
        }
        __syncthreads();
    }
    // reflectors (scaled to unit diagonal) and R back to P; R block (b x b, zero below the diagonal / beyond rows) to the stack
    {
        const double sc = c < TS_PB ? scale_s[c] : 1.0, be = beta_s[c];
        const bool has_ref = c < nref;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            int r = rbase + i;
            double v = x[i];
            if (has_ref) v = r > c ? v * sc : (r == c ? be : v);
            if (r < rows && c < b) P[(r0 + r) * ld + c] = v;
            if (r < b && c < b) Rstack[((int64_t)chunk * b + r) * b + c] = (r <= c && r < nref) ? v : 0.0;
        }
    }
    if (tid < b) tau_out[(int64_t)chunk * b + tid] = tau_s[tid];
}

// Form the explicit Q rows of one chunk: Q_chunk = H_0 ... H_{b-1} [M; 0], M = Min rows [chunk*b, chunk*b + b) (identity if null).
__device__ __forceinline__ void tsqr_apply_body(double* __restrict__ P, int64_t ld, int64_t nrows, int b, const double* __restrict__ tau_in,
                                                const double* __restrict__ Min, double* sm, int chunk) {
    constexpr int VP = TS_CH + 2;                  // even pitch: 16-byte aligned rows for LDS.128 reads of the reflectors
    double* Vs = sm;                               // [TS_PB][VP]: reflector j with unit diagonal, zeros above
    double* part = sm + TS_PB * VP;             // [2][TS_NW][TS_PB]
    double* tau_s = part + 2 * TS_NW * TS_PB;      // [TS_PB]
    const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
    const int64_t r0 = (int64_t)chunk * TS_CH;
    const int rows = (int)((nrows - r0) < TS_CH ? (nrows - r0) : TS_CH);
    const int nref = b < rows ? b : rows;
    const int rbase = 32 * w;

    double z[32];
#pragma unroll
    for (int i = 0; i < 32; i++) {
        int r = rbase + i;
        double v = (r < rows && c < b) ? P[(r0 + r) * ld + c] : 0.0;
        Vs[c * VP + r] = (r > c) ? v : (r == c ? 1.0 : 0.0);
        double zz = 0.0;
        if (r < nref && c < b) zz = Min ? Min[((int64_t)chunk * b + r) * b + c] : (r == c ? 1.0 : 0.0);
        z[i] = zz;
    }
    if (tid < TS_PB) tau_s[tid] = tid < b ? tau_in[(int64_t)chunk * b + tid] : 0.0;
    __syncthreads();

    int buf = 0;
    for (int j = nref - 1; j >= 0; j--) {
        const double2* v2 = reinterpret_cast<const double2*>(Vs + j * VP + rbase);
        double v[32];
#pragma unroll
        for (int i = 0; i < 16; i++) { double2 t = v2[i]; v[2 * i] = t.x; v[2 * i + 1] = t.y; }
        double dsum;
        TS_DOT32(dsum, v[i] * z[i]);
        double* pb = part + buf * TS_NW * TS_PB;
        pb[w * TS_PB + c] = dsum;
        __syncthreads();
        double wc = 0.0;
#pragma unroll
        for (int k = 0; k < TS_NW; k++) wc += pb[k * TS_PB + c];
        wc *= tau_s[j];
#pragma unroll
        for (int i = 0; i < 32; i++) z[i] -= v[i] * wc;
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 32; i++) {
        int r = rbase + i;
        if (r < rows && c < b) P[(r0 + r) * ld + c] = z[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Shared pieces.  Thread (warp w, lane c) keeps rows 32w..32w+31 of column c of its 256-row chunk in registers (as in the
// Householder kernels); the chunk is staged transposed in shared memory, Xt[c][r], so a column's 32 rows are contiguous.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int CQ_XP = TS_CH + 2;            // pitch of Xt (even: 16-byte aligned rows for 128-bit broadcast reads)
constexpr int CQ_GP = TS_PB + 1;            // pitch of the small matrices
constexpr double CQ_PIVOT_TOL = 1e-11;

__device__ __forceinline__ void chunk_load(double (&x)[32], const double* __restrict__ P, int64_t ld, int64_t r0, int rows, int b, int w, int c) {
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const int r = 32 * w + i;
        x[i] = (r < rows && c < b) ? P[(r0 + r) * ld + c] : 0.0;
    }
}
__device__ __forceinline__ void chunk_store(const double (&x)[32], double* __restrict__ P, int64_t ld, int64_t r0, int rows, int b, int w, int c) {
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const int r = 32 * w + i;
        if (r < rows && c < b) P[(r0 + r) * ld + c] = x[i];
    }
}
__device__ __forceinline__ void chunk_stage(const double (&x)[32], double* __restrict__ Xt, int w, int c) {
    double2* dst = reinterpret_cast<double2*>(Xt + c * CQ_XP + 32 * w);
#pragma unroll
    for (int i = 0; i < 16; i++) dst[i] = make_double2(x[2 * i], x[2 * i + 1]);
}
// x[i] (rows 32w+i of column c)  <-  sum_{c' < b} X[row][c'] * M[c'][c]   with X staged in Xt and M (pitch mp) in shared memory
__device__ __forceinline__ void chunk_times_matrix(double (&x)[32], const double* __restrict__ Xt, const double* __restrict__ M, int mp, int b,
                                                   int w, int c, bool upper) {
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0.0;
    const int kmax = upper ? (c < b ? c + 1 : 0) : b;        // upper triangular M: only c' <= c contribute
    for (int k = 0; k < kmax; k++) {
        const double m = M[k * mp + c];
        const double2* src = reinterpret_cast<const double2*>(Xt + k * CQ_XP + 32 * w);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double2 t = src[i];
            acc[2 * i] = fma(t.x, m, acc[2 * i]);
            acc[2 * i + 1] = fma(t.y, m, acc[2 * i + 1]);
        }
    }
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = acc[i];
}
// x[i] (rows 32w+i of column c)  -=  sum_{k < 32} X[row][k] * M[k][c]   (block classical Gram-Schmidt update inside the fused kernel)
__device__ __forceinline__ void chunk_minus_times_matrix(double (&x)[32], const double* __restrict__ Xt, const double* __restrict__ M, int mp, int w,
                                                         int c) {
    for (int k = 0; k < TS_PB; k++) {
        const double m = -M[k * mp + c];
        const double2* src = reinterpret_cast<const double2*>(Xt + k * CQ_XP + 32 * w);
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const double2 t = src[i];
            x[2 * i] = fma(t.x, m, x[2 * i]);
            x[2 * i + 1] = fma(t.y, m, x[2 * i + 1]);
        }
    }
}
// partial Gram matrix of this CTA's chunk -> Gout[32][CQ_GP] (global): G[c][k] = sum_rows X[r][c] X[r][k]
// (every thread accumulates its column against all b columns over its warp's 32 rows; 8-way reduction through shared memory)
// (bi = number of valid columns on the x side; the fused kernel multiplies a full 32-column block of Q by a narrower panel)
__device__ __forceinline__ void chunk_gram(const double (&x)[32], const double* __restrict__ Xt, double* __restrict__ part, int b, int w, int c,
                                           double* __restrict__ Gout, int bi = -1) {
    if (bi < 0) bi = b;
    for (int k = 0; k < b; k += 2) {
        const double2* s0 = reinterpret_cast<const double2*>(Xt + k * CQ_XP + 32 * w);
        const double2* s1 = reinterpret_cast<const double2*>(Xt + (k + 1) * CQ_XP + 32 * w);      // k + 1 <= 31: inside Xt (zero column if >= b)
        double p[8];
#pragma unroll
        for (int u = 0; u < 8; u++) p[u] = 0.0;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const double2 t = s0[i], u = s0[i + 1], v = s1[i], z = s1[i + 1];
            p[0] = fma(t.x, x[2 * i], p[0]); p[1] = fma(t.y, x[2 * i + 1], p[1]);
            p[2] = fma(u.x, x[2 * i + 2], p[2]); p[3] = fma(u.y, x[2 * i + 3], p[3]);
            p[4] = fma(v.x, x[2 * i], p[4]); p[5] = fma(v.y, x[2 * i + 1], p[5]);
            p[6] = fma(z.x, x[2 * i + 2], p[6]); p[7] = fma(z.y, x[2 * i + 3], p[7]);
        }
        part[(w * TS_PB + c) * CQ_GP + k] = (p[0] + p[1]) + (p[2] + p[3]);
        part[(w * TS_PB + c) * CQ_GP + k + 1] = (p[4] + p[5]) + (p[6] + p[7]);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TS_PB * TS_PB; e += TS_CH) {
        const int i = e >> 5, k = e & 31;
        double g = 0.0;
#pragma unroll
        for (int ww = 0; ww < TS_NW; ww++) g += part[(ww * TS_PB + i) * CQ_GP + k];
        Gout[i * CQ_GP + k] = (i < bi && k < b) ? g : 0.0;
    }
}
constexpr size_t CQ_SMEM = (size_t)(TS_PB * CQ_XP + TS_NW * TS_PB * CQ_GP + TS_PB * CQ_GP) * sizeof(double);

// ---------------------------------------------------------------------------------------------------------------------
// Fast path: CholeskyQR2 of the whole panel.  state[0] = 1 while the fast path is valid, state[1] = 1 once the panel is done.
// ---------------------------------------------------------------------------------------------------------------------
// pass 0: partial Gram of every chunk.   pass 1: P <- P Rinv (if the first factorisation was accepted), then the partial Gram again.
// pass 2: P <- P Rinv (if the second one was accepted).
__global__ void __launch_bounds__(TS_CH, 1)
cholqr_chunk_kernel(double* __restrict__ P, int64_t ld, int64_t nrows, int b, const double* __restrict__ Rinv, double* __restrict__ Gpart,
                    const int* __restrict__ state, const int* __restrict__ run_flag, int pass) {
    if (run_flag != nullptr && *run_flag == 0) return;    // second BCGS pass found the panel already orthogonal to eps
    if (pass > 0 && state[0] == 0) return;                // the panel went to the Householder path
    extern __shared__ __align__(16) double sm[];
    double* Xt = sm;
    double* part = Xt + TS_PB * CQ_XP;
    double* Ms = part + TS_NW * TS_PB * CQ_GP;
    const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * TS_CH;
    const int rows = (int)((nrows - r0) < TS_CH ? (nrows - r0) : TS_CH);
    double x[32];
    chunk_load(x, P, ld, r0, rows, b, w, c);
    if (pass > 0) {
        for (int e = tid; e < TS_PB * TS_PB; e += TS_CH) Ms[(e >> 5) * CQ_GP + (e & 31)] = Rinv[(e >> 5) * CQ_GP + (e & 31)];
        chunk_stage(x, Xt, w, c);
        __syncthreads();
        chunk_times_matrix(x, Xt, Ms, CQ_GP, b, w, c, true);
        chunk_store(x, P, ld, r0, rows, b, w, c);
        if (pass == 2) return;
        __syncthreads();
    }
    chunk_stage(x, Xt, w, c);
    __syncthreads();
    chunk_gram(x, Xt, part, b, w, c, Gpart + (int64_t)blockIdx.x * TS_PB * CQ_GP);
}

// One CTA: G = sum of the partial Gram matrices (fixed order), Cholesky G = R^T R in the registers of warp 0 (lane k = column k; 32
// fully unrolled steps of shuffles + FMAs), R^-1 by back substitution.  pass 1: the factorisation is accepted only if every pivot
// keeps more than CQ_PIVOT_TOL of its diagonal entry, otherwise state[0] <- 0 (Householder path).  pass 2: G is I + O(eps cond^2);
// a failed pivot there also hands the (already better conditioned) panel to the Householder path; success sets state[1].
// Body shared by cholqr_factor_kernel and the fused kernel.  All 256 threads take part in the reduction; warp 0 factors.  Returns (in
// every thread of warp 0; other warps get `true`) whether every pivot passed the test.  G: shared scratch [32][CQ_GP]; Rinv: [32][CQ_GP],
// global or shared.  Ends with the warp-0 writes outstanding: the caller synchronises before anybody reads Rinv.
__device__ __forceinline__ bool cholqr_factor_body(const double* __restrict__ Gpart, int nchunks, int b, double* __restrict__ Rinv, double* __restrict__ G) {
    const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
    // fixed-order sum of the partial Gram matrices; all four entries of a thread and eight chunks at a time in flight (the loads are
    // L2 hits of ~0.5 us latency: issued one by one they would cost more than the factorisation itself)
    {
        double g[4] = {0.0, 0.0, 0.0, 0.0};
        const double* base = Gpart + (tid >> 5) * CQ_GP + (tid & 31);            // entry e = tid + 256 u  ->  row (tid >> 5) + 8 u, column tid & 31
        int ch = 0;
        for (; ch + 8 <= nchunks; ch += 8) {
            double v[4][8];
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int t = 0; t < 8; t++) v[u][t] = __ldcg(base + (int64_t)(ch + t) * TS_PB * CQ_GP + u * 8 * CQ_GP);
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int t = 0; t < 8; t++) g[u] += v[u][t];
        }
        for (; ch < nchunks; ch++)
#pragma unroll
            for (int u = 0; u < 4; u++) g[u] += __ldcg(base + (int64_t)ch * TS_PB * CQ_GP + u * 8 * CQ_GP);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = (tid >> 5) + 8 * u, k = tid & 31;
            G[i * CQ_GP + k] = (i < b && k < b) ? g[u] : (i == k ? 1.0 : 0.0);   // identity padding keeps the unrolled steps regular
        }
    }
    __syncthreads();
    if (w != 0) return true;
    double g[TS_PB], dinv[TS_PB];
#pragma unroll
    for (int i = 0; i < TS_PB; i++) g[i] = G[i * CQ_GP + c];
    const double gdiag = G[c * CQ_GP + c];
    bool good = true;
#pragma unroll
    for (int j = 0; j < TS_PB; j++) {
        const double d = __shfl_sync(0xffffffffu, g[j], j);                       // current pivot G[j][j]
        const double d0 = __shfl_sync(0xffffffffu, gdiag, j);                     // its original diagonal entry
        good = good && (d > CQ_PIVOT_TOL * d0) && (d0 < 1e300);
        const double inv = d > 0.0 ? rsqrt(d) : 0.0;
        dinv[j] = inv;                                                            // 1 / R[j][j] (every lane)
        const double rjk = c >= j ? g[j] * inv : 0.0;                             // R[j][c]
        g[j] = rjk;
#pragma unroll
        for (int i = j + 1; i < TS_PB; i++) g[i] = fma(-__shfl_sync(0xffffffffu, rjk, i), rjk, g[i]);
    }
#pragma unroll
    for (int i = 0; i < TS_PB; i++) G[i * CQ_GP + c] = i <= c ? g[i] : 0.0;       // R (upper)
    __syncwarp();
    // R^-1 (upper), column c by back substitution in axpy form (one FMA + one multiply on the critical path per step)
    double sv[TS_PB];
#pragma unroll
    for (int i = 0; i < TS_PB; i++) sv[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int i = TS_PB - 1; i >= 0; i--) {
        const double vi = i <= c ? sv[i] * dinv[i] : 0.0;
        sv[i] = vi;
#pragma unroll
        for (int k = 0; k < i; k++) sv[k] = fma(-G[k * CQ_GP + i], vi, sv[k]);
    }
#pragma unroll
    for (int i = 0; i < TS_PB; i++) Rinv[i * CQ_GP + c] = sv[i];
    return good;
}

__global__ void __launch_bounds__(TS_CH, 1)
cholqr_factor_kernel(const double* __restrict__ Gpart, int nchunks, int b, double* __restrict__ Rinv, int* __restrict__ state,
                     const int* __restrict__ run_flag, int pass) {
    if (run_flag != nullptr && *run_flag == 0) return;
    if (pass == 1) { if (threadIdx.x == 0) { state[0] = 1; state[1] = 0; } }
    else if (state[0] == 0) return;
    __shared__ double G[TS_PB * CQ_GP];
    const bool good = cholqr_factor_body(Gpart, nchunks, b, Rinv, G);
    if (threadIdx.x == 0) {
        if (!good) state[0] = 0;
        else if (pass == 2) state[1] = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fallback: Householder factorisation of one chunk + the explicit Q from its reflectors (the two bodies above), one launch per level.
// Runs only when the fast path gave up (state[0] == 0).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TS_CH, 1)
tsqr_chunk_kernel(double* __restrict__ P, int64_t ld, int64_t nrows, int b, double* __restrict__ Rstack, double* __restrict__ tau_scratch,
                  const int* __restrict__ state, const int* __restrict__ run_flag) {
    if (run_flag != nullptr && *run_flag == 0) return;
    if (state != nullptr && state[1] != 0) return;        // CholeskyQR2 finished the panel
    extern __shared__ __align__(16) double sm[];
    __shared__ FactorSmem fs;
    tsqr_factor_body(P, ld, nrows, b, Rstack, tau_scratch, fs, blockIdx.x);
    __syncthreads();
    tsqr_apply_body(P, ld, nrows, b, tau_scratch, nullptr, sm, blockIdx.x);
}

// Q chunk c of a level  <-  Q0_c * M_c,  M_c = rows [c b, c b + b) of the final explicit Q of the next level (ld = b)
__global__ void __launch_bounds__(TS_CH, 1)
tsqr_mul_kernel(double* __restrict__ P, int64_t ld, int64_t nrows, int b, const double* __restrict__ Mnext, const int* __restrict__ state,
                const int* __restrict__ run_flag) {
    if (run_flag != nullptr && *run_flag == 0) return;
    if (state != nullptr && state[1] != 0) return;
    extern __shared__ __align__(16) double sm[];
    double* Xt = sm;
    double* Ms = sm + TS_PB * CQ_XP;                  // [TS_PB][CQ_GP]
    const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * TS_CH;
    const int rows = (int)((nrows - r0) < TS_CH ? (nrows - r0) : TS_CH);
    double x[32];
    chunk_load(x, P, ld, r0, rows, b, w, c);
    chunk_stage(x, Xt, w, c);
    for (int e = tid; e < TS_PB * TS_PB; e += TS_CH) {
        const int i = e >> 5, k = e & 31;
        Ms[i * CQ_GP + k] = (i < b && k < b) ? Mnext[((int64_t)blockIdx.x * b + i) * b + k] : 0.0;
    }
    __syncthreads();
    chunk_times_matrix(x, Xt, Ms, CQ_GP, b, w, c, false);
    chunk_store(x, P, ld, r0, rows, b, w, c);
}
constexpr size_t MUL_SMEM = (size_t)(TS_PB * CQ_XP + TS_PB * CQ_GP) * sizeof(double);

// run_flag = 1 iff max|c| > thresh: the re-orthogonalisation coefficients of the second BCGS pass are not negligible
__global__ void bcgs_flag_kernel(const double* __restrict__ c, int n, double thresh, int* __restrict__ run_flag) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine |= (fabs(c[i]) > thresh) ? 1 : 0;
    if (mine) any = 1;
    __syncthreads();
    if (threadIdx.x == 0) *run_flag = any;
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused orthonormalisation for small tall matrices (m <= FUSED_MAX_ROWS, q <= FUSED_MAX_COLS): the WHOLE algorithm above --
// every panel, both BCGS passes, CholeskyQR2 with its pivot test and the Householder TSQR fallback -- in ONE cooperative
// kernel, one CTA per 256-row chunk, grid barriers where the multi-launch path has kernel boundaries.  At these sizes the
// multi-launch path is ~25 dependent launches of 1-8 us per panel (launch-latency bound: 58 % of the launches and half of the
// kernel time of a D=4, chi=64 sweep); the matrix (<= 4 MB) stays in L2 between the stages.  Same arithmetic in the same
// order per chunk as the multi-launch path except for the inter-panel projections, which use the chunk kernels' FMA code
// instead of K1 (results agree to rounding, not bit for bit).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FUSED_MAX_ROWS = 4096;     // <= 16 CTAs: a cooperative grid this small is always co-resident next to the bulk kernels
constexpr int FUSED_MAX_COLS = 130;      // coefficient block of the projections lives in shared memory

struct OrthoFusedParams {
    double* Y;
    int64_t ld;
    int m, q, nchunks, force_householder;
    double* Spart;       // [nchunks][q/32 blocks][32][CQ_GP]  partial projection coefficients
    double* Gpart[2];    // [nchunks][32][CQ_GP]  partial Gram matrices of the two CholeskyQR passes
    // Householder levels (as in tsqr_panel): the R stack rst[l] of level l is the matrix of level l + 1 (sized for 32-column panels)
    double* rst[4];
    double* tau[4];
};

// SYNC: how the CTAs (one per chunk) meet between stages.  0: a single CTA, __syncthreads.  1: one thread-block cluster (<= 8 CTAs),
// hardware cluster barrier with release/acquire -- an ordinary launch, which also replays inside CUDA graphs next to other branches.
// 2: cooperative grid barrier (9..16 CTAs).
template <int SYNC>
__device__ __forceinline__ void fused_sync() {
    if (SYNC == 0) __syncthreads();
    else if (SYNC == 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else cooperative_groups::this_grid().sync();
}

template <int SYNC>
__global__ void __launch_bounds__(TS_CH, 1) ortho_fused_kernel(const OrthoFusedParams p) {
    extern __shared__ __align__(16) double sm[];
    __shared__ FactorSmem fs;
    __shared__ int s_flag;
    double* Xt = sm;                                   // [32][CQ_XP]   staged panel / block, transposed
    double* part = Xt + TS_PB * CQ_XP;                 // [8 * 32][CQ_GP]
    double* Ms = part + TS_NW * TS_PB * CQ_GP;         // [32][CQ_GP]   R^-1 / M block
    double* Gs = Ms + TS_PB * CQ_GP;                   // [32][CQ_GP]   Gram scratch
    double* Ss = Gs + TS_PB * CQ_GP;                   // [q32][CQ_GP]  projection coefficients
    const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
    const int chunk = blockIdx.x;
    const int64_t r0 = (int64_t)chunk * TS_CH;
    const int rows = (int)((p.m - r0) < TS_CH ? (p.m - r0) : TS_CH);
    const int nblk_max = (p.q + TS_PB - 1) / TS_PB;
    double xp[32], xq[32];

    for (int j0 = 0; j0 < p.q; j0 += TS_PB) {
        const int b = (p.q - j0) < TS_PB ? (p.q - j0) : TS_PB;
        double* P = p.Y + j0;
        const int nblk = j0 / TS_PB;
        for (int pass = 0; pass < 2; pass++) {
            if (j0 == 0 && pass == 1) break;
            chunk_load(xp, P, p.ld, r0, rows, b, w, c);
            if (j0 > 0) {
                // ---- coefficients of the panel against the finished columns: S = Q_prev^T P, chunk partials then a fixed-order sum
                __syncthreads();
                chunk_stage(xp, Xt, w, c);
                __syncthreads();
                for (int ib = 0; ib < nblk; ib++) {
                    chunk_load(xq, p.Y + ib * TS_PB, p.ld, r0, rows, TS_PB, w, c);
                    chunk_gram(xq, Xt, part, b, w, c, p.Spart + ((int64_t)(chunk * nblk_max + ib) * TS_PB) * CQ_GP, TS_PB);
                    __syncthreads();
                }
                fused_sync<SYNC>();
                double mx = 0.0;
                for (int e = tid; e < nblk * TS_PB * TS_PB; e += TS_CH) {
                    const int ib = e / (TS_PB * TS_PB), r = e - ib * TS_PB * TS_PB, i = r >> 5, k = r & 31;
                    double sacc = 0.0;
                    for (int ch = 0; ch < p.nchunks; ch++) sacc += __ldcg(p.Spart + ((int64_t)(ch * nblk_max + ib) * TS_PB + i) * CQ_GP + k);
                    Ss[(ib * TS_PB + i) * CQ_GP + k] = sacc;
                    mx = fmax(mx, fabs(sacc));
                }
                if (tid == 0) s_flag = 0;
                __syncthreads();
                // "twice is enough": with second-pass coefficients below 1e-10 the update below leaves the panel orthonormal to O(1e-20),
                // so its re-factorisation is skipped (uniform decision: every CTA sums the same numbers in the same order)
                bool refactor = true;
                if (pass == 1) {
                    if (mx > 1e-10) s_flag = 1;
                    __syncthreads();
                    refactor = s_flag != 0;
                }
                // ---- P -= Q_prev S on this chunk
                for (int ib = 0; ib < nblk; ib++) {
                    chunk_load(xq, p.Y + ib * TS_PB, p.ld, r0, rows, TS_PB, w, c);
                    __syncthreads();
                    chunk_stage(xq, Xt, w, c);
                    __syncthreads();
                    chunk_minus_times_matrix(xp, Xt, Ss + ib * TS_PB * CQ_GP, CQ_GP, w, c);
                }
                __syncthreads();
                if (!refactor) {
                    chunk_store(xp, P, p.ld, r0, rows, b, w, c);
                    break;
                }
            }
            // ---- the panel itself: CholeskyQR2, accepted pass by pass on the pivot test
            bool done = false;
            if (!p.force_householder && p.m >= b) {
                bool ok = true;
                for (int cp = 0; cp < 2 && ok; cp++) {
                    chunk_stage(xp, Xt, w, c);
                    __syncthreads();
                    chunk_gram(xp, Xt, part, b, w, c, p.Gpart[cp] + (int64_t)chunk * TS_PB * CQ_GP);
                    fused_sync<SYNC>();
                    const bool good = cholqr_factor_body(p.Gpart[cp], p.nchunks, b, Ms, Gs);
                    if (tid == 0) s_flag = good ? 1 : 0;
                    __syncthreads();
                    ok = s_flag != 0;
                    if (ok) chunk_times_matrix(xp, Xt, Ms, CQ_GP, b, w, c, true);      // P <- P R^-1 (Xt still holds the panel)
                    __syncthreads();
                }
                done = ok;
            }
            chunk_store(xp, P, p.ld, r0, rows, b, w, c);
            if (done) continue;
            // ---- Householder TSQR with explicit chunk factors, level by level (uniform branch: `done` is the same in every CTA)
            __syncthreads();
            int lrows[4], lchunks[4], nlevels = 0;
            for (int rws = p.m; nlevels < 4;) {
                lrows[nlevels] = rws; lchunks[nlevels] = (rws + TS_CH - 1) / TS_CH; nlevels++;
                if (lchunks[nlevels - 1] == 1) break;
                rws = lchunks[nlevels - 1] * b;
            }
            for (int l = 0; l < nlevels; l++) {
                if (chunk < lchunks[l]) {
                    double* M = l == 0 ? P : p.rst[l - 1];
                    const int64_t mld = l == 0 ? p.ld : b;
                    tsqr_factor_body(M, mld, lrows[l], b, p.rst[l], p.tau[l], fs, chunk);
                    __syncthreads();
                    tsqr_apply_body(M, mld, lrows[l], b, p.tau[l], nullptr, sm, chunk);
                    __syncthreads();
                }
                fused_sync<SYNC>();
            }
            for (int l = nlevels - 2; l >= 0; l--) {
                if (chunk < lchunks[l]) {
                    double* M = l == 0 ? P : p.rst[l - 1];
                    const int64_t mld = l == 0 ? p.ld : b;
                    const int64_t rr0 = (int64_t)chunk * TS_CH;
                    const int rrows = (int)((lrows[l] - rr0) < TS_CH ? (lrows[l] - rr0) : TS_CH);
                    chunk_load(xq, M, mld, rr0, rrows, b, w, c);
                    chunk_stage(xq, Xt, w, c);
                    for (int e = tid; e < TS_PB * TS_PB; e += TS_CH) {
                        const int i = e >> 5, k = e & 31;
                        Ms[i * CQ_GP + k] = (i < b && k < b) ? p.rst[l][((int64_t)chunk * b + i) * b + k] : 0.0;
                    }
                    __syncthreads();
                    chunk_times_matrix(xq, Xt, Ms, CQ_GP, b, w, c, false);
                    chunk_store(xq, M, mld, rr0, rrows, b, w, c);
                    __syncthreads();
                }
                fused_sync<SYNC>();
            }
        }
    }
}
constexpr size_t fused_smem_bytes(int q) {
    return (size_t)(TS_PB * CQ_XP + TS_NW * TS_PB * CQ_GP + 2 * TS_PB * CQ_GP + ((q + TS_PB - 1) / TS_PB) * TS_PB * CQ_GP) * sizeof(double);
}

namespace {
constexpr size_t APPLY_SMEM = (size_t)(TS_PB * (TS_CH + 2) + 2 * TS_NW * TS_PB + TS_PB) * sizeof(double);

struct Levels {
    int n;
    int64_t rows[8];
    int64_t chunks[8];
};
Levels plan_levels(int64_t m, int b) {
    Levels L; L.n = 0;
    int64_t rows = m;
    while (true) {
        int64_t ch = (rows + TS_CH - 1) / TS_CH;
        L.rows[L.n] = rows; L.chunks[L.n] = ch; L.n++;
        if (ch == 1 || L.n == 8) break;
        rows = ch * b;
    }
    return L;
}
size_t tsqr_scratch_doubles(int64_t m) {
    Levels L = plan_levels(m, TS_PB);
    size_t tot = 0;
    for (int l = 0; l < L.n; l++) tot += (size_t)L.chunks[l] * TS_PB * TS_PB + (size_t)L.chunks[l] * TS_PB + 64;
    // CholeskyQR2: partial Gram matrices of the level-0 chunks, R^-1, state words
    tot += (size_t)L.chunks[0] * TS_PB * CQ_GP + TS_PB * CQ_GP + 64;
    return tot;
}

// orthonormalise one panel P (m x b, leading dimension ld) in place
int tsqr_panel(double* P, int64_t m, int b, int64_t ld, double* scratch, const int* run_flag, cudaStream_t s) {
    static_assert(CQ_SMEM >= APPLY_SMEM, "the Householder fallback needs less dynamic shared memory than the CholeskyQR kernels");
    AB_ENSURE_SMEM(cholqr_chunk_kernel, CQ_SMEM);
    AB_ENSURE_SMEM(tsqr_chunk_kernel, APPLY_SMEM);
    AB_ENSURE_SMEM(tsqr_mul_kernel, MUL_SMEM);
    // dev / test knob, read per call: ACETN_B200_TSQR_HOUSEHOLDER=1 disables the CholeskyQR2 fast path
    const char* fh = getenv("ACETN_B200_TSQR_HOUSEHOLDER");
    const int force_householder = fh != nullptr && fh[0] == '1';
    Levels L = plan_levels(m, b);
    double* mat[9]; int64_t lds[9]; double* rst[8]; double* tau[8];
    mat[0] = P; lds[0] = ld;
    double* cur = scratch;
    for (int l = 0; l < L.n; l++) {
        rst[l] = cur; cur += (size_t)L.chunks[l] * b * b;
        tau[l] = cur; cur += (size_t)L.chunks[l] * b + 64 - ((size_t)L.chunks[l] * b) % 2;
        mat[l + 1] = rst[l]; lds[l + 1] = b;
    }
    double* Gpart = scratch + tsqr_scratch_doubles(m) - ((size_t)L.chunks[0] * TS_PB * CQ_GP + TS_PB * CQ_GP + 64);
    double* Rinv = Gpart + (size_t)L.chunks[0] * TS_PB * CQ_GP;
    int* state = reinterpret_cast<int*>(Rinv + TS_PB * CQ_GP);
    const unsigned nch = (unsigned)L.chunks[0];
    const int* fb_state = state;
    if (!force_householder && m >= b) {
        // ---- fast path: CholeskyQR2 (5 launches); every kernel after the first factorisation is a no-op once state[0] drops to 0
        cholqr_chunk_kernel<<<nch, TS_CH, CQ_SMEM, s>>>(P, ld, m, b, Rinv, Gpart, state, run_flag, 0);
        AB_LAUNCHED();
        cholqr_factor_kernel<<<1, TS_CH, 0, s>>>(Gpart, (int)nch, b, Rinv, state, run_flag, 1);
        AB_LAUNCHED();
        cholqr_chunk_kernel<<<nch, TS_CH, CQ_SMEM, s>>>(P, ld, m, b, Rinv, Gpart, state, run_flag, 1);
        AB_LAUNCHED();
        cholqr_factor_kernel<<<1, TS_CH, 0, s>>>(Gpart, (int)nch, b, Rinv, state, run_flag, 2);
        AB_LAUNCHED();
        cholqr_chunk_kernel<<<nch, TS_CH, CQ_SMEM, s>>>(P, ld, m, b, Rinv, Gpart, state, run_flag, 2);
        AB_LAUNCHED();
    } else {
        fb_state = nullptr;
    }
    // ---- fallback (skipped on the device when the fast path finished): Householder TSQR with explicit chunk factors.
    // bottom-up: the R stack of level l is the matrix of level l + 1
    for (int l = 0; l < L.n; l++) {
        tsqr_chunk_kernel<<<(unsigned)L.chunks[l], TS_CH, APPLY_SMEM, s>>>(mat[l], lds[l], L.rows[l], b, rst[l], tau[l], fb_state, run_flag);
        AB_LAUNCHED();
    }
    // top-down: chunk c of level l <- Q0_c * (rows of the final Q of level l + 1); the top level is final as it stands
    for (int l = L.n - 2; l >= 0; l--) {
        tsqr_mul_kernel<<<(unsigned)L.chunks[l], TS_CH, MUL_SMEM, s>>>(mat[l], lds[l], L.rows[l], b, mat[l + 1], fb_state, run_flag);
        AB_LAUNCHED();
    }
    return OK;
}

GemmDesc proj_coeff_desc(const double* Y, int64_t ld, int64_t m, int j0, int b, double* Sc) {
    // Sc[i][c] = sum_r Y[r][i] * Y[r][j0 + c]
    return gemm_desc(j0, b, (int)m, operand(Y, idx1(1), idx1(ld)), operand(Y + j0, idx1(ld), idx1(1)), Sc, idx1(b), idx1(1));
}
GemmDesc proj_update_desc(double* Y, int64_t ld, int64_t m, int j0, int b, const double* Sc) {
    // P[r][c] -= sum_i Y[r][i] * Sc[i][c]
    return gemm_desc((int)m, b, j0, operand(Y, idx1(ld), idx1(1)), operand(Sc, idx1(b), idx1(1)), Y + j0, idx1(ld), idx1(1), -1.0, 1.0);
}
}  // namespace

namespace {
bool fused_eligible(int64_t m, int q) {
    if (m > FUSED_MAX_ROWS || q > FUSED_MAX_COLS || m < q) return false;
    const char* e = getenv("ACETN_B200_ORTHO_FUSED");     // dev / test knob, read per call
    return !(e != nullptr && e[0] == '0');
}
size_t fused_extra_doubles(int64_t m, int q) {
    const size_t nch = (size_t)((m + TS_CH - 1) / TS_CH), nblk = (size_t)((q + TS_PB - 1) / TS_PB);
    return nch * nblk * TS_PB * CQ_GP + 2 * nch * TS_PB * CQ_GP + 64;
}

int ortho_fused_launch(double* Y, int64_t m, int q, int64_t ld, double* scratch, double* extra, cudaStream_t s) {
    OrthoFusedParams p;
    memset(&p, 0, sizeof(p));
    p.Y = Y; p.ld = ld; p.m = (int)m; p.q = q;
    p.nchunks = (int)((m + TS_CH - 1) / TS_CH);
    const char* fh = getenv("ACETN_B200_TSQR_HOUSEHOLDER");
    p.force_householder = fh != nullptr && fh[0] == '1';
    const size_t nblk = (size_t)((q + TS_PB - 1) / TS_PB);
    p.Spart = extra;
    p.Gpart[0] = extra + (size_t)p.nchunks * nblk * TS_PB * CQ_GP;
    p.Gpart[1] = p.Gpart[0] + (size_t)p.nchunks * TS_PB * CQ_GP;
    Levels L = plan_levels(m, TS_PB);                       // pointer layout of tsqr_panel (sized for full panels)
    AB_REQUIRE(L.n <= 4, "orthonormalize (fused): too many TSQR levels");
    double* cur = scratch;
    for (int l = 0; l < L.n; l++) {
        p.rst[l] = cur; cur += (size_t)L.chunks[l] * TS_PB * TS_PB;
        p.tau[l] = cur; cur += (size_t)L.chunks[l] * TS_PB + 64 - ((size_t)L.chunks[l] * TS_PB) % 2;
    }
    const size_t smem = fused_smem_bytes(q);
    if (p.nchunks == 1) {
        AB_ENSURE_SMEM(ortho_fused_kernel<0>, smem);
        ortho_fused_kernel<0><<<1, TS_CH, smem, s>>>(p);
        AB_LAUNCHED();
    } else if (p.nchunks <= 8) {
        AB_ENSURE_SMEM(ortho_fused_kernel<1>, smem);
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)p.nchunks); cfg.blockDim = dim3(TS_CH); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.nchunks; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        AB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ortho_fused_kernel<1>, p));
        note_launch(1);
    } else {
        AB_ENSURE_SMEM(ortho_fused_kernel<2>, smem);
        void* args[] = {&p};
        AB_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)ortho_fused_kernel<2>, dim3((unsigned)p.nchunks), dim3(TS_CH), args, smem, s));
        note_launch(1);
    }
    return OK;
}
}  // namespace

size_t orthonormalize_workspace_bytes(int64_t m, int q) {
    size_t bytes = ws_round(tsqr_scratch_doubles(m) * sizeof(double));
    if (m <= FUSED_MAX_ROWS && q <= FUSED_MAX_COLS) bytes += ws_round(fused_extra_doubles(m, q) * sizeof(double));
    bytes += ws_round((size_t)q * TS_PB * sizeof(double)) + ws_round(16 * sizeof(int));
    size_t g = 0;
    for (int j0 = TS_PB; j0 < q; j0 += TS_PB) {
        int b = (q - j0) < TS_PB ? (q - j0) : TS_PB;
        size_t a = gemm_workspace_bytes(proj_coeff_desc(nullptr, q, m, j0, b, nullptr));
        size_t c = gemm_workspace_bytes(proj_update_desc(nullptr, q, m, j0, b, nullptr));
        if (a > g) g = a;
        if (c > g) g = c;
    }
    return bytes + g + 1024;
}

int orthonormalize_launch(double* Y, int64_t m, int q, int64_t ld, void* wsp, size_t ws_bytes, cudaStream_t s) {
    AB_REQUIRE(q >= 1 && m >= q, "orthonormalize: need m >= q >= 1 (m=%lld q=%d)", (long long)m, q);
    AB_REQUIRE(m < 2147483647LL, "orthonormalize: m too large");
    Workspace ws(wsp, ws_bytes);
    double* scratch = ws.take<double>(tsqr_scratch_doubles(m));
    double* extra = (m <= FUSED_MAX_ROWS && q <= FUSED_MAX_COLS) ? ws.take<double>(fused_extra_doubles(m, q)) : nullptr;
    double* Sc = ws.take<double>((size_t)q * TS_PB);
    int* flag = ws.take<int>(16);
    if (ws.overflow) { set_error("orthonormalize: workspace too small"); return ERR_WORKSPACE; }
    if (extra != nullptr && fused_eligible(m, q)) return ortho_fused_launch(Y, m, q, ld, scratch, extra, s);
    void* gws = ws.base + ws.used;
    size_t gws_bytes = ws.bytes - ws.used;
    for (int j0 = 0; j0 < q; j0 += TS_PB) {
        int b = (q - j0) < TS_PB ? (q - j0) : TS_PB;
        for (int pass = 0; pass < 2; pass++) {
            if (j0 == 0 && pass == 1) break;   // the first panel has nothing to be re-orthogonalised against
            const int* run_flag = nullptr;
            if (j0 > 0) {
                AB_TRY(gemm_launch(proj_coeff_desc(Y, ld, m, j0, b, Sc), gws, gws_bytes, s));
                if (pass == 1) {
                    // "twice is enough": if the second-pass coefficients are below 1e-10 the panel is orthonormal to
                    // O(1e-20 * j0*b) after the update and the second TSQR is skipped on the device (no host sync)
                    bcgs_flag_kernel<<<1, 256, 0, s>>>(Sc, j0 * b, 1e-10, flag);
                    AB_LAUNCHED();
                    run_flag = flag;
                }
                AB_TRY(gemm_launch(proj_update_desc(Y, ld, m, j0, b, Sc), gws, gws_bytes, s));
            }
            AB_TRY(tsqr_panel(Y + j0, m, b, ld, scratch, run_flag, s));
        }
    }
    return OK;
}

}  // namespace ab200
