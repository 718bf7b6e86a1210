// K4: orthonormal basis of a tall-skinny matrix (the QR steps of the randomized SVD, reference
// acetn/linalg/fused_matmul_svd_lowrank.py:38,43; only the Q factor is ever used there).
//
// Block classical Gram-Schmidt with re-orthogonalisation (BCGS2) over 32-column panels; the inter-panel
// projections are two K1 DGEMMs each, the intra-panel factorisation is a Householder TSQR:
//   level 0: every 256-row chunk is factored in shared memory (reflectors stay in place, R goes to a stack),
//   level l: the stack of R factors is factored the same way until one chunk is left,
//   then the explicit Q is formed top-down by applying the stored reflectors to [M;0] blocks.
// Householder keeps ||Q^T Q - I|| at machine precision for any conditioning of Y (CholeskyQR would lose the
// trailing directions of the power-iterated Y, SURVEY.md section 7 hard part 2).
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
__global__ void __launch_bounds__(TS_CH, 1)
tsqr_factor_kernel(double* __restrict__ P, int64_t ld, int64_t nrows, int b, double* __restrict__ Rstack,
                   double* __restrict__ tau_out, const int* __restrict__ run_flag) {
    if (run_flag != nullptr && *run_flag == 0) return;    // second BCGS pass found the panel already orthogonal to eps
    __shared__ __align__(16) double col_s[TS_CH];   // raw pivot column of the current step
    __shared__ double part[TS_NW][TS_PB];           // per-warp partial dots
    __shared__ double partn[TS_NW];                 // per-warp partial sums of squares of the pivot column below the diagonal
    __shared__ double tau_s[TS_PB], beta_s[TS_PB], scale_s[TS_PB];
    const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * TS_CH;
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
            if (r < b && c < b) Rstack[((int64_t)blockIdx.x * b + r) * b + c] = (r <= c && r < nref) ? v : 0.0;
        }
    }
    if (tid < b) tau_out[(int64_t)blockIdx.x * b + tid] = tau_s[tid];
}

// Form the explicit Q rows of one chunk: Q_chunk = H_0 ... H_{b-1} [M; 0], M = Min rows [chunk*b, chunk*b + b) (identity if null).
__global__ void __launch_bounds__(TS_CH, 1)
tsqr_apply_kernel(double* __restrict__ P, int64_t ld, int64_t nrows, int b, const double* __restrict__ tau_in,
                  const double* __restrict__ Min, const int* __restrict__ run_flag) {
    if (run_flag != nullptr && *run_flag == 0) return;
    extern __shared__ double sm[];
    constexpr int VP = TS_CH + 2;                  // even pitch: 16-byte aligned rows for LDS.128 reads of the reflectors
    double* Vs = sm;                               // [TS_PB][VP]: reflector j with unit diagonal, zeros above
    double* part = sm + TS_PB * VP;             // [2][TS_NW][TS_PB]
    double* tau_s = part + 2 * TS_NW * TS_PB;      // [TS_PB]
    const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * TS_CH;
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
        if (r < nref && c < b) zz = Min ? Min[((int64_t)blockIdx.x * b + r) * b + c] : (r == c ? 1.0 : 0.0);
        z[i] = zz;
    }
    if (tid < TS_PB) tau_s[tid] = tid < b ? tau_in[(int64_t)blockIdx.x * b + tid] : 0.0;
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

namespace {
constexpr size_t FACTOR_SMEM = 0;
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
    return tot;
}

// orthonormalise one panel P (m x b, leading dimension ld) in place
int tsqr_panel(double* P, int64_t m, int b, int64_t ld, double* scratch, const int* run_flag, cudaStream_t s) {
    AB_ENSURE_SMEM(tsqr_apply_kernel, APPLY_SMEM);
    Levels L = plan_levels(m, b);
    double* mat[9]; int64_t lds[9]; double* rst[8]; double* tau[8];
    mat[0] = P; lds[0] = ld;
    double* cur = scratch;
    for (int l = 0; l < L.n; l++) {
        rst[l] = cur; cur += (size_t)L.chunks[l] * b * b;
        tau[l] = cur; cur += (size_t)L.chunks[l] * b + 64 - ((size_t)L.chunks[l] * b) % 2;
        mat[l + 1] = rst[l]; lds[l + 1] = b;
    }
    for (int l = 0; l < L.n; l++) {
        tsqr_factor_kernel<<<(unsigned)L.chunks[l], TS_CH, FACTOR_SMEM, s>>>(mat[l], lds[l], L.rows[l], b, rst[l], tau[l], run_flag);
        AB_LAUNCHED();
    }
    for (int l = L.n - 1; l >= 0; l--) {
        const double* Min = (l == L.n - 1) ? nullptr : mat[l + 1];
        tsqr_apply_kernel<<<(unsigned)L.chunks[l], TS_CH, APPLY_SMEM, s>>>(mat[l], lds[l], L.rows[l], b, tau[l], Min, run_flag);
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

size_t orthonormalize_workspace_bytes(int64_t m, int q) {
    size_t bytes = ws_round(tsqr_scratch_doubles(m) * sizeof(double));
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
    double* Sc = ws.take<double>((size_t)q * TS_PB);
    int* flag = ws.take<int>(16);
    if (ws.overflow) { set_error("orthonormalize: workspace too small"); return ERR_WORKSPACE; }
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
