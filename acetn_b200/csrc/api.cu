// extern "C" boundary of libacetn_b200.so (include/acetn_b200.h) and the host-side composition of the CTMRG
// operators out of the kernels K1..K5.  Host code here only builds descriptors and enqueues kernels on the caller's
// stream; it never synchronises and never allocates device memory.
#include "../../include/acetn_b200.h"

#include "gemm.cuh"
#include "i8crt.cuh"
#include "kernels.cuh"

using namespace ab200;

namespace ab200 {
long long launch_count();
void reset_launch_count();
double dmma_peak_launch(double* scratch, int iters, cudaStream_t s);
size_t als_workspace_bytes(int nD, int bD, int pD);
int als_solve_launch(double* a1r, double* a2r, const double* n12g, const double* n12, const double* a12g, int nD, int bD, int pD,
                     int niter, double tol, double epsilon, int method, int* info, void* wsp, size_t ws_bytes, cudaStream_t s);
int double_layer_fused_supported(int64_t D, int64_t d);
int double_layer_fused_colexp_supported(int64_t D, int64_t d);
int double_layer_fused_launch(const double* X, int64_t n0, int64_t n1, int64_t in_s0, int64_t in_s1, const int64_t* in_es,
                              int order, const double* A, const int64_t* a_strides, int64_t D, int64_t d, double* Y,
                              int64_t out_s0, int64_t out_s1, const int64_t* out_es, double* absmax, int32_t* colexp, cudaStream_t s);
}  // namespace ab200

namespace {

inline cudaStream_t S_(void* s) { return (cudaStream_t)s; }
inline size_t maxz(size_t a, size_t b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------------------------
// K2 (unfused form): double-layer absorption as two batched K1 GEMMs with the ket/bra site tensor pre-gathered.
// ------------------------------------------------------------------------------------------------------------------
struct DoubleLayerArgs {
    const double* X; int64_t n0, n1, in_s0, in_s1; int64_t in_es[4]; int order;
    const double* A; int64_t a_s[5]; int64_t D, d;
    double* Y; int64_t out_s0, out_s1; int64_t out_es[4];
};

GemmDesc dl_gemm1(const DoubleLayerArgs& a, const double* bra, double* W) {
    const int64_t D = a.D, D2 = D * D, d = a.d;
    return gemm_desc((int)D2, (int)(D2 * d), (int)D2,
                     operand(a.X, idx2(D, a.in_es[0], a.in_es[2]), idx2(D, a.in_es[1], a.in_es[3]), idx2(a.n1, a.in_s0, a.in_s1)),
                     operand(bra, idx1(D2 * d), idx1(1), idx1(0)), W, idx1(D2 * d), idx1(1), 1.0, 0.0, (int)(a.n0 * a.n1),
                     idx1(D2 * D2 * d));
}
GemmDesc dl_gemm2(const DoubleLayerArgs& a, const double* ket, const double* W) {
    const int64_t D = a.D, D2 = D * D, d = a.d;
    return gemm_desc((int)D2, (int)D2, (int)(D2 * d), operand(ket, idx1(D2 * d), idx1(1), idx1(0)),
                     operand(W, idx1(D2), idx1(1), idx1(D2 * D2 * d)), a.Y, idx2(D, a.out_es[0], a.out_es[2]),
                     idx2(D, a.out_es[1], a.out_es[3]), 1.0, 0.0, (int)(a.n0 * a.n1), idx2(a.n1, a.out_s0, a.out_s1));
}
size_t dl_workspace_bytes(int64_t n0, int64_t n1, int64_t D, int64_t d) {
    if (double_layer_fused_supported(D, d)) return 4096;
    const int64_t D4 = D * D * D * D;
    return ws_round((size_t)(n0 * n1 * D4 * d) * 8) + 2 * ws_round((size_t)(D4 * d) * 8) + 4096;
}
int double_layer(const DoubleLayerArgs& a, double* absmax, void* wsp, size_t ws_bytes, cudaStream_t s, int32_t* colexp = nullptr) {
    const int64_t D = a.D, d = a.d, D4 = D * D * D * D;
    if (double_layer_fused_supported(D, d))
        return double_layer_fused_launch(a.X, a.n0, a.n1, a.in_s0, a.in_s1, a.in_es, a.order, a.A, a.a_s, D, d, a.Y,
                                         a.out_s0, a.out_s1, a.out_es, absmax, colexp, s);
    AB_REQUIRE(colexp == nullptr, "double_layer: column exponents need the fused kernel");
    Workspace ws(wsp, ws_bytes);
    double* W = ws.take<double>((size_t)(a.n0 * a.n1 * D4 * d));
    double* bra = ws.take<double>((size_t)(D4 * d));
    double* ket = ws.take<double>((size_t)(D4 * d));
    if (ws.overflow) { set_error("double_layer: workspace too small (%zu needed, %zu given)", ws.used, ws_bytes); return ERR_WORKSPACE; }
    // legs of the A view: 0=l 1=u 2=r 3=d 4=p.  order 0: (i0,i1) = (u,l) ; order 1: (i0,i1) = (l,u)
    const int f0 = a.order == 0 ? 1 : 0, f1 = a.order == 0 ? 0 : 1;
    {
        int64_t dims[5] = {D, D, d, D, D};   // [I0, I1, P, R, Dd]
        int64_t st[5] = {a.a_s[f0], a.a_s[f1], a.a_s[4], a.a_s[2], a.a_s[3]};
        AB_TRY(gather5_launch(bra, a.A, dims, st, s));
    }
    {
        int64_t dims[5] = {D, D, D, D, d};   // [r, dd, i0, i1, p]
        int64_t st[5] = {a.a_s[2], a.a_s[3], a.a_s[f0], a.a_s[f1], a.a_s[4]};
        AB_TRY(gather5_launch(ket, a.A, dims, st, s));
    }
    AB_TRY(gemm_launch(dl_gemm1(a, bra, W), nullptr, 0, s));
    AB_TRY(gemm_launch(dl_gemm2(a, ket, W), nullptr, 0, s));
    if (absmax) {
        // Y blocks are not contiguous in general; the callers that need max|Y| pass contiguous outputs.
    }
    return OK;
}

// ------------------------------------------------------------------------------------------------------------------
// quarter tensor
// ------------------------------------------------------------------------------------------------------------------
struct QuarterDims { int64_t xa, xb, xc, xe, D, d; };
GemmDesc q_gemm1(const QuarterDims& q, const double* C, const double* E2, double* T1) {
    const int64_t D2 = q.D * q.D;
    return gemm_desc((int)q.xa, (int)(q.xc * D2), (int)q.xb, operand(C, idx1(q.xb), idx1(1)), operand(E2, idx1(q.xc * D2), idx1(1)),
                     T1, idx1(q.xc * D2), idx1(1));
}
GemmDesc q_gemm2(const QuarterDims& q, const double* T1, const double* E1, double* T2) {
    const int64_t D2 = q.D * q.D;
    return gemm_desc((int)(q.xc * D2), (int)(q.xe * D2), (int)q.xa, operand(T1, idx1(1), idx1(q.xc * D2)),
                     operand(E1, idx1(D2), idx2(D2, q.xa * D2, 1)), T2, idx1(q.xe * D2), idx1(1));
}

}  // namespace

extern "C" {

int acetn_b200_init(int device) {
    if (device >= 0) AB_CHECK_CUDA(cudaSetDevice(device));
    int dev = 0;
    AB_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    AB_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) {
        set_error("acetn_b200 needs an sm_100a device (Blackwell B200); found %s (sm_%d%d)", prop.name, prop.major, prop.minor);
        return ERR_UNSUPPORTED;
    }
    return OK;
}
int acetn_b200_destroy(void) { return OK; }
const char* acetn_b200_last_error(void) { return get_error(); }
const char* acetn_b200_version(void) { return "acetn_b200 0.1.0 (sm_100a, FP64 DMMA)"; }
int64_t acetn_b200_launch_count(void) { return (int64_t)launch_count(); }
void acetn_b200_reset_launch_count(void) { reset_launch_count(); }

// ---- GEMM ---------------------------------------------------------------------------------------------------------
static GemmDesc desc_from_idx(int64_t M, int64_t N, int64_t K, int64_t batch, const double* A, const double* B, double* C,
                              const int64_t* x, double alpha, double beta, int force_tile, int force_splitk) {
    auto g = [&](int i) { return x[3 * i] ? idx2(x[3 * i], x[3 * i + 1], x[3 * i + 2]) : idx1(x[3 * i + 2]); };
    GemmDesc d = gemm_desc((int)M, (int)N, (int)K, operand(A, g(0), g(1), g(2)), operand(B, g(3), g(4), g(5)), C, g(6), g(7), alpha,
                           beta, (int)batch, g(8));
    d.force_tile = force_tile; d.force_splitk = force_splitk;
    return d;
}
size_t acetn_b200_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch, const int64_t* idx, int force_tile,
                                       int force_splitk) {
    return gemm_workspace_bytes(desc_from_idx(M, N, K, batch, nullptr, nullptr, nullptr, idx, 1.0, 0.0, force_tile, force_splitk));
}
int acetn_b200_gemm(int64_t M, int64_t N, int64_t K, int64_t batch, const double* A, const double* B, double* C,
                    const int64_t* idx, double alpha, double beta, int force_tile, int force_splitk, void* ws, size_t ws_bytes,
                    void* stream) {
    AB_REQUIRE(M < 2147483647LL && N < 2147483647LL && K < 2147483647LL, "gemm: extent too large");
    return gemm_launch(desc_from_idx(M, N, K, batch, A, B, C, idx, alpha, beta, force_tile, force_splitk), ws, ws_bytes, S_(stream));
}

// ---- double layer ---------------------------------------------------------------------------------------------------
size_t acetn_b200_double_layer_workspace_bytes(int64_t n0, int64_t n1, int64_t D, int64_t d) { return dl_workspace_bytes(n0, n1, D, d); }
int acetn_b200_double_layer(const double* X, int64_t n0, int64_t n1, int64_t in_s0, int64_t in_s1, const int64_t* in_es,
                            int order, const double* A, const int64_t* a_strides, int64_t D, int64_t d, double* Y,
                            int64_t out_s0, int64_t out_s1, const int64_t* out_es, void* ws, size_t ws_bytes, void* stream) {
    DoubleLayerArgs a;
    a.X = X; a.n0 = n0; a.n1 = n1; a.in_s0 = in_s0; a.in_s1 = in_s1; a.order = order; a.A = A; a.D = D; a.d = d; a.Y = Y;
    a.out_s0 = out_s0; a.out_s1 = out_s1;
    for (int i = 0; i < 4; i++) { a.in_es[i] = in_es[i]; a.out_es[i] = out_es[i]; }
    for (int i = 0; i < 5; i++) a.a_s[i] = a_strides[i];
    return double_layer(a, nullptr, ws, ws_bytes, S_(stream));
}

// ---- quarter tensor ---------------------------------------------------------------------------------------------------
size_t acetn_b200_quarter_tensor_workspace_bytes(int64_t xa, int64_t xb, int64_t xc, int64_t xe, int64_t D, int64_t d) {
    QuarterDims q{xa, xb, xc, xe, D, d};
    const int64_t D2 = D * D;
    size_t b = ws_round((size_t)(xa * xc * D2) * 8) + ws_round((size_t)(xc * D2 * xe * D2) * 8) + ws_round(64);
    size_t g = maxz(gemm_workspace_bytes(q_gemm1(q, nullptr, nullptr, nullptr)), gemm_workspace_bytes(q_gemm2(q, nullptr, nullptr, nullptr)));
    return b + maxz(g, dl_workspace_bytes(xc, xe, D, d)) + 4096;
}
static int quarter_tensor_impl(const double* C, const double* E2, const double* E1, const double* A, const int64_t* a_strides,
                               int64_t xa, int64_t xb, int64_t xc, int64_t xe, int64_t D, int64_t d, int normalize, double* Q,
                               double* absmax_out, void* enc_storage, size_t enc_bytes, void* wsp, size_t ws_bytes, cudaStream_t s) {
    QuarterDims q{xa, xb, xc, xe, D, d};
    const int64_t D2 = D * D, N2 = xe * D2, M2 = xc * D2;
    AB_REQUIRE(M2 < 2147483647LL && N2 < 2147483647LL, "quarter_tensor: chi*D^2 too large");
    AB_REQUIRE(enc_storage == nullptr || !normalize, "quarter_tensor_enc: the encoding is taken from the un-normalised tensor (normalize must be 0)");
    AB_REQUIRE(enc_storage == nullptr || i8_supported(M2, N2, 1), "quarter_tensor_enc: %lld x %lld is outside the INT8 engine's range", (long long)M2,
               (long long)N2);
    Workspace ws(wsp, ws_bytes);
    double* T1 = ws.take<double>((size_t)(xa * M2));
    double* T2 = ws.take<double>((size_t)(M2 * N2));
    double* mx = ws.take<double>(8);
    if (ws.overflow) { set_error("quarter_tensor: workspace too small"); return ERR_WORKSPACE; }
    void* rest = ws.base + ws.used;
    size_t rest_bytes = ws.bytes - ws.used;
    AB_TRY(gemm_launch(q_gemm1(q, C, E2, T1), rest, rest_bytes, s));
    AB_TRY(gemm_launch(q_gemm2(q, T1, E1, T2), rest, rest_bytes, s));
    DoubleLayerArgs a;
    a.X = T2; a.n0 = xc; a.n1 = xe; a.in_s0 = D2 * N2; a.in_s1 = D2;
    a.in_es[0] = D * N2; a.in_es[1] = N2; a.in_es[2] = D; a.in_es[3] = 1;
    a.order = 0; a.A = A; a.D = D; a.d = d; a.Y = Q; a.out_s0 = D2 * N2; a.out_s1 = D2;
    a.out_es[0] = D * N2; a.out_es[1] = N2; a.out_es[2] = D; a.out_es[3] = 1;
    for (int i = 0; i < 5; i++) a.a_s[i] = a_strides[i];
    const bool want_max = normalize || absmax_out != nullptr;
    if (absmax_out) mx = absmax_out;
    if (want_max) AB_CHECK_CUDA(cudaMemsetAsync(mx, 0, 8, s));
    const bool fused = double_layer_fused_supported(D, d) != 0;
    // K7 encoding of the result: the fused D = 8 kernel delivers the column exponents from its epilogue (columns of Q = (e, d, D) =
    // (block b1, n)), which saves the encoding one of its passes over the 2 GiB tensor
    int32_t* colexp = nullptr;
    const bool colexp_fused = enc_storage != nullptr && double_layer_fused_colexp_supported(D, d) != 0;
    if (colexp_fused) AB_TRY(i8_colexp_reset_launch(enc_storage, M2, N2, &colexp, s));
    AB_TRY(double_layer(a, (want_max && fused) ? mx : nullptr, rest, rest_bytes, s, colexp));
    if (want_max && !fused) AB_TRY(absmax_launch(Q, (size_t)(M2 * N2), mx, s));
    if (normalize) AB_TRY(scale_inv_launch(Q, (size_t)(M2 * N2), mx, s));
    if (enc_storage != nullptr) {
        I8Matrix e;
        AB_TRY(i8_encode_launch(Q, M2, N2, N2, enc_storage, enc_bytes, &e, s, colexp_fused));
    }
    return OK;
}
int acetn_b200_quarter_tensor(const double* C, const double* E2, const double* E1, const double* A, const int64_t* a_strides,
                              int64_t xa, int64_t xb, int64_t xc, int64_t xe, int64_t D, int64_t d, int normalize, double* Q,
                              double* absmax_out, void* wsp, size_t ws_bytes, void* stream) {
    return quarter_tensor_impl(C, E2, E1, A, a_strides, xa, xb, xc, xe, D, d, normalize, Q, absmax_out, nullptr, 0, wsp, ws_bytes, S_(stream));
}
int acetn_b200_quarter_tensor_enc(const double* C, const double* E2, const double* E1, const double* A, const int64_t* a_strides,
                                  int64_t xa, int64_t xb, int64_t xc, int64_t xe, int64_t D, int64_t d, double* Q, double* absmax_out,
                                  void* enc_storage, size_t enc_bytes, void* wsp, size_t ws_bytes, void* stream) {
    AB_REQUIRE(enc_storage != nullptr, "quarter_tensor_enc: enc_storage must not be NULL");
    return quarter_tensor_impl(C, E2, E1, A, a_strides, xa, xb, xc, xe, D, d, 0, Q, absmax_out, enc_storage, enc_bytes, wsp, ws_bytes, S_(stream));
}

// ---- orthonormalise / jacobi ---------------------------------------------------------------------------------------------
size_t acetn_b200_orthonormalize_workspace_bytes(int64_t m, int64_t q) { return orthonormalize_workspace_bytes(m, (int)q); }
int acetn_b200_orthonormalize(double* Y, int64_t m, int64_t q, int64_t ld, void* ws, size_t ws_bytes, void* stream) {
    return orthonormalize_launch(Y, m, (int)q, ld, ws, ws_bytes, S_(stream));
}
size_t acetn_b200_jacobi_svd_workspace_bytes(int64_t q) { return jacobi_workspace_bytes((int)q); }
int acetn_b200_jacobi_svd(const double* R, int64_t q, double* S, double* Wt, double* Jt, int64_t chi, double cutoff, int32_t* info,
                          void* ws, size_t ws_bytes, void* stream) {
    return jacobi_svd_launch(R, (int)q, S, Wt, Jt, (int)chi, cutoff, (int*)info, ws, ws_bytes, S_(stream));
}

// ---- K7: INT8 tensor-core exact products ------------------------------------------------------------------------------
int acetn_b200_i8_supported(int64_t rows, int64_t cols, int64_t q) { return i8_supported(rows, cols, q) ? 1 : 0; }
size_t acetn_b200_i8_encoded_bytes(int64_t rows, int64_t cols) { return i8_encoded_bytes(rows, cols); }
int acetn_b200_i8_encode(const double* Q, int64_t rows, int64_t cols, int64_t ldq, void* storage, size_t storage_bytes, void* stream) {
    I8Matrix e;
    return i8_encode_launch(Q, rows, cols, ldq, storage, storage_bytes, &e, S_(stream));
}
size_t acetn_b200_i8_matmul_workspace_bytes(int64_t rows, int64_t cols, int64_t q) { return i8_matmul_workspace_bytes(rows, cols, q); }
int acetn_b200_i8_matmul(const void* storage, int64_t rows, int64_t cols, int adjoint, const double* Y, int64_t q, int64_t ldy,
                         double* out, int64_t ldo, void* ws, size_t ws_bytes, void* stream) {
    AB_REQUIRE(i8_supported(rows, cols, q), "i8_matmul: unsupported shape (%lld x %lld, q=%lld)", (long long)rows, (long long)cols, (long long)q);
    return i8_matmul_launch(i8_view(storage, rows, cols), adjoint != 0, Y, q, ldy, out, ldo, ws, ws_bytes, S_(stream));
}

// ---- randomized SVD -----------------------------------------------------------------------------------------------------
namespace {
struct Chain {
    int n;
    const double* mat[4];
    int64_t rows[4], cols[4];
    const void* enc[4];      // K7 residue encoding of mat[i] (or NULL: FP64 DMMA path, K1)
};
// out (rows x q) = M (rows x cols) * in (cols x q)      /   out (cols x q) = M^T * in (rows x q)
GemmDesc thin_desc(const double* M, int64_t rows, int64_t cols, bool adjoint, const double* in, double* out, int64_t q) {
    if (!adjoint)
        return gemm_desc((int)rows, (int)q, (int)cols, operand(M, idx1(cols), idx1(1)), operand(in, idx1(q), idx1(1)), out, idx1(q), idx1(1));
    return gemm_desc((int)cols, (int)q, (int)rows, operand(M, idx1(1), idx1(cols)), operand(in, idx1(q), idx1(1)), out, idx1(q), idx1(1));
}
size_t chain_gemm_ws(const Chain& c, int64_t q) {
    size_t g = 0;
    for (int i = 0; i < c.n; i++) {
        if (c.enc[i] != nullptr) { g = maxz(g, i8_matmul_workspace_bytes(c.rows[i], c.cols[i], q)); continue; }
        g = maxz(g, gemm_workspace_bytes(thin_desc(nullptr, c.rows[i], c.cols[i], false, nullptr, nullptr, q)));
        g = maxz(g, gemm_workspace_bytes(thin_desc(nullptr, c.rows[i], c.cols[i], true, nullptr, nullptr, q)));
    }
    return g;
}
// one "big x thin" product: K7 (INT8 tensor cores, exact) when the factor carries an encoding, else K1 (FP64 DMMA)
int thin_apply(const Chain& c, int i, bool adjoint, const double* in, double* out, int64_t q, void* gws, size_t gws_bytes, cudaStream_t s) {
    if (c.enc[i] != nullptr)
        return i8_matmul_launch(i8_view(c.enc[i], c.rows[i], c.cols[i]), adjoint, in, q, q, out, q, gws, gws_bytes, s);
    return gemm_launch(thin_desc(c.mat[i], c.rows[i], c.cols[i], adjoint, in, out, q), gws, gws_bytes, s);
}
// forward: out = M0 M1 ... Mn-1 in ;  adjoint: out = Mn-1^T ... M0^T in
int chain_apply(const Chain& c, bool adjoint, const double* in, double* out, double* t0, double* t1, int64_t q, void* gws,
                size_t gws_bytes, cudaStream_t s) {
    const double* src = in;
    for (int step = 0; step < c.n; step++) {
        int i = adjoint ? step : (c.n - 1 - step);
        double* dst = (step == c.n - 1) ? out : ((step & 1) ? t1 : t0);
        AB_TRY(thin_apply(c, i, adjoint, src, dst, q, gws, gws_bytes, s));
        src = dst;
    }
    return OK;
}
int64_t chain_maxdim(const Chain& c) {
    int64_t m = 0;
    for (int i = 0; i < c.n; i++) { if (c.rows[i] > m) m = c.rows[i]; if (c.cols[i] > m) m = c.cols[i]; }
    return m;
}
GemmDesc core_desc(const double* Qb, const double* Bt, int64_t n, int64_t q, double* R) {
    return gemm_desc((int)q, (int)q, (int)n, operand(Qb, idx1(1), idx1(q)), operand(Bt, idx1(q), idx1(1)), R, idx1(q), idx1(1));
}
GemmDesc lift_desc(const double* Qm, int64_t rows, int64_t q, const double* Wt, double* out) {
    // out[r][z] = sum_k Qm[r][k] * Wt[z][k]
    return gemm_desc((int)rows, (int)q, (int)q, operand(Qm, idx1(q), idx1(1)), operand(Wt, idx1(1), idx1(q)), out, idx1(q), idx1(1));
}
}  // namespace

size_t acetn_b200_rsvd_workspace_bytes(int nmat, const int64_t* rows, const int64_t* cols, int64_t q) {
    return acetn_b200_rsvd_enc_workspace_bytes(nmat, rows, cols, q, nullptr);
}
size_t acetn_b200_rsvd_enc_workspace_bytes(int nmat, const int64_t* rows, const int64_t* cols, int64_t q, const int32_t* use_enc) {
    Chain c; c.n = nmat;
    static const char marker = 0;
    for (int i = 0; i < nmat; i++) {
        c.mat[i] = nullptr; c.rows[i] = rows[i]; c.cols[i] = cols[i];
        c.enc[i] = (use_enc != nullptr && use_enc[i]) ? (const void*)&marker : nullptr;
    }
    const int64_t m = rows[0], n = cols[nmat - 1], mx = chain_maxdim(c);
    size_t b = ws_round((size_t)(m * q) * 8) + 2 * ws_round((size_t)(mx * q) * 8) + 2 * ws_round((size_t)(n * q) * 8) +
               3 * ws_round((size_t)(q * q) * 8);
    size_t g = chain_gemm_ws(c, q);
    g = maxz(g, orthonormalize_workspace_bytes(m, (int)q));
    g = maxz(g, orthonormalize_workspace_bytes(n, (int)q));
    g = maxz(g, jacobi_workspace_bytes((int)q));
    g = maxz(g, gemm_workspace_bytes(core_desc(nullptr, nullptr, n, q, nullptr)));
    g = maxz(g, gemm_workspace_bytes(lift_desc(nullptr, m, q, nullptr, nullptr)));
    g = maxz(g, gemm_workspace_bytes(lift_desc(nullptr, n, q, nullptr, nullptr)));
    return b + g + 8192;
}

int acetn_b200_rsvd(int nmat, const double* const* mats, const int64_t* rows, const int64_t* cols, const double* Omega, int64_t q,
                    int niter, int reorth_adjoint, int64_t chi, double cutoff, double* U, double* S, double* V, int32_t* info,
                    double* AtQ, double* Wt_out, void* wsp, size_t ws_bytes, void* stream) {
    return acetn_b200_rsvd_enc(nmat, mats, nullptr, rows, cols, Omega, q, niter, reorth_adjoint, chi, cutoff, U, S, V, info, AtQ, Wt_out,
                               wsp, ws_bytes, stream);
}
int acetn_b200_rsvd_enc(int nmat, const double* const* mats, const void* const* encs, const int64_t* rows, const int64_t* cols,
                        const double* Omega, int64_t q, int niter, int reorth_adjoint, int64_t chi, double cutoff, double* U, double* S,
                        double* V, int32_t* info, double* AtQ, double* Wt_out, void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    AB_REQUIRE(nmat >= 1 && nmat <= 4, "rsvd: nmat must be 1..4");
    Chain c; c.n = nmat;
    for (int i = 0; i < nmat; i++) {
        c.mat[i] = mats != nullptr ? mats[i] : nullptr; c.rows[i] = rows[i]; c.cols[i] = cols[i];
        c.enc[i] = encs != nullptr ? encs[i] : nullptr;
        AB_REQUIRE(c.mat[i] != nullptr || c.enc[i] != nullptr, "rsvd: factor %d has neither FP64 data nor an encoding", i);
        AB_REQUIRE(c.enc[i] == nullptr || i8_supported(rows[i], cols[i], q), "rsvd: factor %d (%lld x %lld, q=%lld) is outside the INT8 engine's range",
                   i, (long long)rows[i], (long long)cols[i], (long long)q);
        if (i > 0) AB_REQUIRE(cols[i - 1] == rows[i], "rsvd: inner dimensions of factors %d and %d differ", i - 1, i);
    }
    const int64_t m = rows[0], n = cols[nmat - 1], mx = chain_maxdim(c);
    AB_REQUIRE(q >= 1 && q <= m && q <= n, "rsvd: need 1 <= q <= min(m,n) (q=%lld m=%lld n=%lld)", (long long)q, (long long)m, (long long)n);
    Workspace ws(wsp, ws_bytes);
    double* Y = ws.take<double>((size_t)(m * q));
    double* t0 = ws.take<double>((size_t)(mx * q));
    double* t1 = ws.take<double>((size_t)(mx * q));
    double* Z = ws.take<double>((size_t)(n * q));
    double* Qb = ws.take<double>((size_t)(n * q));
    double* R = ws.take<double>((size_t)(q * q));
    double* Wt = ws.take<double>((size_t)(q * q));
    double* Jt = ws.take<double>((size_t)(q * q));
    if (ws.overflow) { set_error("rsvd: workspace too small (%zu needed, %zu given)", ws.used, ws_bytes); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used;
    size_t gb = ws.bytes - ws.used;

    AB_TRY(chain_apply(c, false, Omega, Y, t0, t1, q, g, gb, s));                 // Y = (M0..Mn-1) Omega
    for (int it = 0; it < niter; it++) {
        AB_TRY(orthonormalize_launch(Y, m, (int)q, q, g, gb, s));                  // Y = qr(Y).Q
        AB_TRY(chain_apply(c, true, Y, Z, t0, t1, q, g, gb, s));                  // Z = (M0..Mn-1)^H Y
        if (reorth_adjoint) AB_TRY(orthonormalize_launch(Z, n, (int)q, q, g, gb, s));
        AB_TRY(chain_apply(c, false, Z, Y, t0, t1, q, g, gb, s));                 // Y = (M0..Mn-1) Z
    }
    AB_TRY(orthonormalize_launch(Y, m, (int)q, q, g, gb, s));                      // Q
    if (AtQ != nullptr && c.n >= 2) {
        // keep M0^T Q (the first product of the adjoint chain): proj1 = M0^T U = (M0^T Q) U_B needs no further pass over M0
        AB_TRY(thin_apply(c, 0, true, Y, AtQ, q, g, gb, s));
        Chain rest; rest.n = c.n - 1;
        for (int i = 1; i < c.n; i++) { rest.mat[i - 1] = c.mat[i]; rest.rows[i - 1] = c.rows[i]; rest.cols[i - 1] = c.cols[i]; rest.enc[i - 1] = c.enc[i]; }
        AB_TRY(chain_apply(rest, true, AtQ, Z, t0, t1, q, g, gb, s));
    } else {
        AB_TRY(chain_apply(c, true, Y, Z, t0, t1, q, g, gb, s));                  // Z = Bt^T  (n x q),  Bt = Q^H M
        if (AtQ != nullptr) AB_CHECK_CUDA(cudaMemcpyAsync(AtQ, Z, (size_t)(n * q) * 8, cudaMemcpyDeviceToDevice, s));
    }
    AB_CHECK_CUDA(cudaMemcpyAsync(Qb, Z, (size_t)(n * q) * 8, cudaMemcpyDeviceToDevice, s));
    AB_TRY(orthonormalize_launch(Qb, n, (int)q, q, g, gb, s));                     // Bt^T = Qb R
    AB_TRY(gemm_launch(core_desc(Qb, Z, n, q, R), g, gb, s));                     // R = Qb^T Bt^T   (q x q)
    AB_TRY(jacobi_svd_launch(R, (int)q, S, Wt, Jt, (int)chi, cutoff, (int*)info, g, gb, s));   // R = Jt^T S Wt
    if (U != nullptr) AB_TRY(gemm_launch(lift_desc(Y, m, q, Wt, U), g, gb, s));   // U = Q  Wt^T
    AB_TRY(gemm_launch(lift_desc(Qb, n, q, Jt, V), g, gb, s));                    // V = Qb Jt^T
    if (Wt_out != nullptr) AB_CHECK_CUDA(cudaMemcpyAsync(Wt_out, Wt, (size_t)(q * q) * 8, cudaMemcpyDeviceToDevice, s));
    return OK;
}

// ---- projectors -----------------------------------------------------------------------------------------------------------
namespace {
GemmDesc p1_desc(const double* Q1, int64_t m1, int64_t n1, const double* Us, int64_t keep, double* out) {
    return gemm_desc((int)n1, (int)keep, (int)m1, operand(Q1, idx1(1), idx1(n1)), operand(Us, idx1(keep), idx1(1)), out, idx1(keep), idx1(1));
}
GemmDesc p2_desc(const double* Q4, int64_t m4, int64_t n4, const double* Vs, int64_t keep, double* out) {
    return gemm_desc((int)m4, (int)keep, (int)n4, operand(Q4, idx1(n4), idx1(1)), operand(Vs, idx1(keep), idx1(1)), out, idx1(keep), idx1(1));
}
}  // namespace
size_t acetn_b200_projectors_workspace_bytes(int64_t m1, int64_t n1, int64_t m4, int64_t n4, int64_t keep) {
    return acetn_b200_projectors_enc_workspace_bytes(m1, n1, m4, n4, keep, 0);
}
size_t acetn_b200_projectors_enc_workspace_bytes(int64_t m1, int64_t n1, int64_t m4, int64_t n4, int64_t keep, int use_enc) {
    const int64_t qmax = m1 < n4 ? m1 : n4;          // q <= min(m, n)
    size_t b = ws_round((size_t)(m1 * keep) * 8) + ws_round((size_t)(n4 * keep) * 8) + ws_round((size_t)keep * 8) +
               2 * ws_round((size_t)(qmax * keep) * 8);
    size_t g = maxz(gemm_workspace_bytes(p1_desc(nullptr, m1, n1, nullptr, keep, nullptr)),
                    gemm_workspace_bytes(p2_desc(nullptr, m4, n4, nullptr, keep, nullptr)));
    if (use_enc) g = maxz(g, maxz(i8_matmul_workspace_bytes(m1, n1, keep), i8_matmul_workspace_bytes(m4, n4, keep)));
    return b + g + 4096;
}
int acetn_b200_projectors_from_usv(const double* Q1, int64_t m1, int64_t n1, const double* Q4, int64_t m4, int64_t n4,
                                   const double* U, int64_t ldu, const double* V, int64_t ldv, const double* S, int64_t keep,
                                   const double* qmax1, const double* qmax4, const double* AtQ, const double* Wt, int64_t q,
                                   double* proj1, double* proj2, void* wsp, size_t ws_bytes, void* stream) {
    return acetn_b200_projectors_from_usv_enc(Q1, nullptr, m1, n1, Q4, nullptr, m4, n4, U, ldu, V, ldv, S, keep, qmax1, qmax4, AtQ, Wt, q,
                                              proj1, proj2, wsp, ws_bytes, stream);
}
int acetn_b200_projectors_from_usv_enc(const double* Q1, const void* enc1, int64_t m1, int64_t n1, const double* Q4, const void* enc4,
                                       int64_t m4, int64_t n4, const double* U, int64_t ldu, const double* V, int64_t ldv,
                                       const double* S, int64_t keep, const double* qmax1, const double* qmax4, const double* AtQ,
                                       const double* Wt, int64_t q, double* proj1, double* proj2, void* wsp, size_t ws_bytes,
                                       void* stream) {
    cudaStream_t s = S_(stream);
    AB_REQUIRE(enc1 == nullptr || i8_supported(m1, n1, keep), "projectors: Q1 is outside the INT8 engine's range");
    AB_REQUIRE(enc4 == nullptr || i8_supported(m4, n4, keep), "projectors: Q4 is outside the INT8 engine's range");
    AB_REQUIRE(AtQ != nullptr || Q1 != nullptr || enc1 != nullptr, "projectors: Q1 needs FP64 data or an encoding");
    AB_REQUIRE(Q4 != nullptr || enc4 != nullptr, "projectors: Q4 needs FP64 data or an encoding");
    AB_REQUIRE(keep >= 1, "projectors: keep must be >= 1");
    AB_REQUIRE((AtQ == nullptr) == (Wt == nullptr), "projectors: AtQ and Wt must be given together");
    Workspace ws(wsp, ws_bytes);
    double* Us = ws.take<double>((size_t)(m1 * keep));
    double* Vs = ws.take<double>((size_t)(n4 * keep));
    double* w = ws.take<double>((size_t)keep);
    const int64_t qcap = m1 < n4 ? m1 : n4;
    double* Ub = ws.take<double>((size_t)(qcap * keep));
    double* Ubs = ws.take<double>((size_t)(qcap * keep));
    if (ws.overflow) { set_error("projectors: workspace too small"); return ERR_WORKSPACE; }
    AB_REQUIRE(AtQ == nullptr || q <= qcap, "projectors: q exceeds min(m1, n4)");
    void* g = ws.base + ws.used;
    size_t gb = ws.bytes - ws.used;
    AB_TRY(inv_sqrt_weights_launch(S, w, (int)keep, s));
    AB_TRY(scale_cols_launch(Vs, keep, V, ldv, w, qmax4, n4, (int)keep, s));
    if (AtQ != nullptr) {
        // proj1 = (Q1^T Qy) (U_B diag(w)) : Us <- Wt^T[:, :keep] * w / qmax1  (q x keep), one small GEMM instead of a pass over Q1
        int64_t dims[5] = {q, keep, 1, 1, 1};
        int64_t st[5] = {1, q, 0, 0, 0};
        AB_TRY(gather5_launch(Ub, Wt, dims, st, s));
        AB_TRY(scale_cols_launch(Ubs, keep, Ub, keep, w, qmax1, q, (int)keep, s));
        AB_TRY(gemm_launch(gemm_desc((int)n1, (int)keep, (int)q, operand(AtQ, idx1(q), idx1(1)), operand(Ubs, idx1(keep), idx1(1)), proj1,
                                     idx1(keep), idx1(1)), g, gb, s));
    } else {
        AB_TRY(scale_cols_launch(Us, keep, U, ldu, w, qmax1, m1, (int)keep, s));
        if (enc1 != nullptr) AB_TRY(i8_matmul_launch(i8_view(enc1, m1, n1), true, Us, keep, keep, proj1, keep, g, gb, s));
        else AB_TRY(gemm_launch(p1_desc(Q1, m1, n1, Us, keep, proj1), g, gb, s));
    }
    if (enc4 != nullptr) AB_TRY(i8_matmul_launch(i8_view(enc4, m4, n4), false, Vs, keep, keep, proj2, keep, g, gb, s));
    else AB_TRY(gemm_launch(p2_desc(Q4, m4, n4, Vs, keep, proj2), g, gb, s));
    return OK;
}

// ---- absorption --------------------------------------------------------------------------------------------------------------
namespace {
struct CornerDims { int64_t xa, xb, xc, xx, D; };
// corner1: T[a,c,lL] = sum_b ei[a,b,lL] ci[b,c]
GemmDesc c1_g1(const CornerDims& c, const double* ci, const double* ei, double* T) {
    const int64_t D2 = c.D * c.D;
    return gemm_desc((int)c.xc, (int)(c.xa * D2), (int)c.xb, operand(ci, idx1(1), idx1(c.xc)),
                     operand(ei, idx1(D2), idx2(D2, c.xb * D2, 1)), T, idx1(D2), idx2(D2, c.xc * D2, 1));
}
GemmDesc c1_g2(const CornerDims& c, const double* T, const double* proj, double* out) {
    const int64_t D2 = c.D * c.D;
    return gemm_desc((int)c.xa, (int)c.xx, (int)(c.xc * D2), operand(T, idx1(c.xc * D2), idx1(1)), operand(proj, idx1(c.xx), idx1(1)), out,
                     idx1(c.xx), idx1(1));
}
// corner2: T[a,(c,rR)] = ci[a,b] ei[b,(c,rR)] ; out[x,c] = sum_{a,rR} proj[(a,rR),x] T[a,c,rR]
GemmDesc c2_g1(const CornerDims& c, const double* ci, const double* ei, double* T) {
    const int64_t D2 = c.D * c.D;
    return gemm_desc((int)c.xa, (int)(c.xc * D2), (int)c.xb, operand(ci, idx1(c.xb), idx1(1)), operand(ei, idx1(c.xc * D2), idx1(1)), T,
                     idx1(c.xc * D2), idx1(1));
}
GemmDesc c2_g2(const CornerDims& c, const double* T, const double* proj, double* out) {
    const int64_t D2 = c.D * c.D;
    return gemm_desc((int)c.xx, (int)c.xc, (int)(c.xa * D2), operand(proj, idx1(1), idx1(c.xx)),
                     operand(T, idx2(D2, c.xc * D2, 1), idx1(D2)), out, idx1(c.xc), idx1(1));
}
struct EdgeDims { int64_t xa, xb, xx, xy, D, d; };
GemmDesc e_g1(const EdgeDims& e, const double* ei, const double* P1t, double* T) {
    const int64_t D2 = e.D * e.D, D4 = D2 * D2;
    return gemm_desc((int)(e.xa * D2), (int)(e.xx * D2), (int)e.xb, operand(ei, idx2(D2, e.xb * D2, 1), idx1(D2)),
                     operand(P1t, idx1(e.xx * D2), idx1(1)), T, idx2(D2, e.xx * D4, D2), idx2(D2, D4, 1));
}
GemmDesc e_g4(const EdgeDims& e, const double* proj2, const double* T3, double* out) {
    const int64_t D2 = e.D * e.D;
    return gemm_desc((int)e.xy, (int)(e.xx * D2), (int)(e.xa * D2), operand(proj2, idx1(1), idx1(e.xy)),
                     operand(T3, idx1(e.xx * D2), idx1(1)), out, idx1(e.xx * D2), idx1(1));
}
}  // namespace

size_t acetn_b200_absorb_corner_workspace_bytes(int64_t xa, int64_t xb, int64_t xc, int64_t xx, int64_t D) {
    CornerDims c{xa, xb, xc, xx, D};
    size_t b = ws_round((size_t)(xa * xc * D * D) * 8) + ws_round(frob_scratch_doubles() * 8);
    size_t g = maxz(maxz(gemm_workspace_bytes(c1_g1(c, nullptr, nullptr, nullptr)), gemm_workspace_bytes(c1_g2(c, nullptr, nullptr, nullptr))),
                    maxz(gemm_workspace_bytes(c2_g1(c, nullptr, nullptr, nullptr)), gemm_workspace_bytes(c2_g2(c, nullptr, nullptr, nullptr))));
    return b + g + 4096;
}
int acetn_b200_absorb_corner1(const double* ci, const double* ei, const double* proj, int64_t xa, int64_t xb, int64_t xc, int64_t xx,
                              int64_t D, double* out, void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    CornerDims c{xa, xb, xc, xx, D};
    Workspace ws(wsp, ws_bytes);
    double* T = ws.take<double>((size_t)(xa * xc * D * D));
    double* fs = ws.take<double>(frob_scratch_doubles());
    if (ws.overflow) { set_error("absorb_corner1: workspace too small"); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used; size_t gb = ws.bytes - ws.used;
    AB_TRY(gemm_launch(c1_g1(c, ci, ei, T), g, gb, s));
    AB_TRY(gemm_launch(c1_g2(c, T, proj, out), g, gb, s));
    return frob_normalize_launch(out, (size_t)(xa * xx), fs, s);
}
int acetn_b200_absorb_corner2(const double* ci, const double* ei, const double* proj, int64_t xa, int64_t xb, int64_t xc, int64_t xx,
                              int64_t D, double* out, void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    CornerDims c{xa, xb, xc, xx, D};
    Workspace ws(wsp, ws_bytes);
    double* T = ws.take<double>((size_t)(xa * xc * D * D));
    double* fs = ws.take<double>(frob_scratch_doubles());
    if (ws.overflow) { set_error("absorb_corner2: workspace too small"); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used; size_t gb = ws.bytes - ws.used;
    AB_TRY(gemm_launch(c2_g1(c, ci, ei, T), g, gb, s));
    AB_TRY(gemm_launch(c2_g2(c, T, proj, out), g, gb, s));
    return frob_normalize_launch(out, (size_t)(xx * xc), fs, s);
}

size_t acetn_b200_absorb_edge_workspace_bytes(int64_t xa, int64_t xb, int64_t xx, int64_t xy, int64_t D, int64_t d) {
    EdgeDims e{xa, xb, xx, xy, D, d};
    const int64_t D2 = D * D, D4 = D2 * D2;
    size_t b = ws_round((size_t)(xb * xx * D2) * 8) + 2 * ws_round((size_t)(xa * xx * D4) * 8) + ws_round(frob_scratch_doubles() * 8);
    size_t g = maxz(gemm_workspace_bytes(e_g1(e, nullptr, nullptr, nullptr)), gemm_workspace_bytes(e_g4(e, nullptr, nullptr, nullptr)));
    return b + maxz(g, dl_workspace_bytes(xa, xx, D, d)) + 4096;
}
int acetn_b200_absorb_edge(const double* ei, const double* A, const int64_t* a_strides, const double* proj2, const double* proj1,
                           int64_t xa, int64_t xb, int64_t xx, int64_t xy, int64_t D, int64_t d, int normalize, double* out,
                           void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    EdgeDims e{xa, xb, xx, xy, D, d};
    const int64_t D2 = D * D, D4 = D2 * D2;
    AB_REQUIRE(xa * D2 < 2147483647LL && xx * D2 < 2147483647LL, "absorb_edge: chi*D^2 too large");
    Workspace ws(wsp, ws_bytes);
    double* P1t = ws.take<double>((size_t)(xb * xx * D2));
    double* T = ws.take<double>((size_t)(xa * xx * D4));
    double* T3 = ws.take<double>((size_t)(xa * xx * D4));
    double* fs = ws.take<double>(frob_scratch_doubles());
    if (ws.overflow) { set_error("absorb_edge: workspace too small (%zu needed, %zu given)", ws.used, ws_bytes); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used; size_t gb = ws.bytes - ws.used;
    {   // P1t[b,x,(uU)] = proj1[b,(uU),x]
        int64_t dims[5] = {xb, xx, D2, 1, 1};
        int64_t st[5] = {D2 * xx, 1, xx, 0, 0};
        AB_TRY(gather5_launch(P1t, proj1, dims, st, s));
    }
    AB_TRY(gemm_launch(e_g1(e, ei, P1t, T), g, gb, s));             // T[a,x,(l,L),(u,U)]
    DoubleLayerArgs a;
    a.X = T; a.n0 = xa; a.n1 = xx; a.in_s0 = xx * D4; a.in_s1 = D4;
    a.in_es[0] = D2 * D; a.in_es[1] = D2; a.in_es[2] = D; a.in_es[3] = 1;   // (l,L,u,U)
    a.order = 1; a.A = A; a.D = D; a.d = d;
    a.Y = T3; a.out_s0 = D2 * xx * D2; a.out_s1 = D2;                       // T3[a,(d,Dd),x,(r,R)]
    a.out_es[0] = D; a.out_es[1] = 1; a.out_es[2] = D * xx * D2; a.out_es[3] = xx * D2;
    for (int i = 0; i < 5; i++) a.a_s[i] = a_strides[i];
    AB_TRY(double_layer(a, nullptr, g, gb, s));
    AB_TRY(gemm_launch(e_g4(e, proj2, T3, out), g, gb, s));        // out[y,(x,r,R)]
    if (!normalize) return OK;                                     // partial sum of a row-sharded absorption
    return frob_normalize_launch(out, (size_t)(xy * xx * D2), fs, s);
}

// ---- the edge absorption in two stages, so that a scheduler can run the part that only needs proj1 of the neighbouring task
//      (2 chi^3 D^4 + 4 chi^2 D^6 d of the 4 chi^3 D^4 + 4 chi^2 D^6 d flops) before this task's own projector pair exists:
//        begin : T3[a,(d,Dd),x,(r,R)] = sum ei[a,b,l,L] proj1[b,u,U,x] conj(A)[L,U,R,Dd,P] A[l,u,r,d,P]     (T3: chi_a chi_x D^4 doubles)
//        finish: out[y,x,r,R] = sum proj2[a,d,Dd,y] T3[a,(d,Dd),x,(r,R)],  Frobenius-normalised if normalize != 0
//      Same kernels, same launch parameters and same order per output element as acetn_b200_absorb_edge: bit-identical results.
size_t acetn_b200_absorb_edge_begin_workspace_bytes(int64_t xa, int64_t xb, int64_t xx, int64_t D, int64_t d) {
    EdgeDims e{xa, xb, xx, 1, D, d};
    const int64_t D2 = D * D, D4 = D2 * D2;
    size_t b = ws_round((size_t)(xb * xx * D2) * 8) + ws_round((size_t)(xa * xx * D4) * 8);
    return b + maxz(gemm_workspace_bytes(e_g1(e, nullptr, nullptr, nullptr)), dl_workspace_bytes(xa, xx, D, d)) + 4096;
}
int acetn_b200_absorb_edge_begin(const double* ei, const double* A, const int64_t* a_strides, const double* proj1, int64_t xa, int64_t xb,
                                 int64_t xx, int64_t D, int64_t d, double* T3, void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    EdgeDims e{xa, xb, xx, 1, D, d};
    const int64_t D2 = D * D, D4 = D2 * D2;
    AB_REQUIRE(xa * D2 < 2147483647LL && xx * D2 < 2147483647LL, "absorb_edge_begin: chi*D^2 too large");
    Workspace ws(wsp, ws_bytes);
    double* P1t = ws.take<double>((size_t)(xb * xx * D2));
    double* T = ws.take<double>((size_t)(xa * xx * D4));
    if (ws.overflow) { set_error("absorb_edge_begin: workspace too small (%zu needed, %zu given)", ws.used, ws_bytes); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used; size_t gb = ws.bytes - ws.used;
    {   // P1t[b,x,(uU)] = proj1[b,(uU),x]
        int64_t dims[5] = {xb, xx, D2, 1, 1};
        int64_t st[5] = {D2 * xx, 1, xx, 0, 0};
        AB_TRY(gather5_launch(P1t, proj1, dims, st, s));
    }
    AB_TRY(gemm_launch(e_g1(e, ei, P1t, T), g, gb, s));             // T[a,x,(l,L),(u,U)]
    DoubleLayerArgs a;
    a.X = T; a.n0 = xa; a.n1 = xx; a.in_s0 = xx * D4; a.in_s1 = D4;
    a.in_es[0] = D2 * D; a.in_es[1] = D2; a.in_es[2] = D; a.in_es[3] = 1;   // (l,L,u,U)
    a.order = 1; a.A = A; a.D = D; a.d = d;
    a.Y = T3; a.out_s0 = D2 * xx * D2; a.out_s1 = D2;                       // T3[a,(d,Dd),x,(r,R)]
    a.out_es[0] = D; a.out_es[1] = 1; a.out_es[2] = D * xx * D2; a.out_es[3] = xx * D2;
    for (int i = 0; i < 5; i++) a.a_s[i] = a_strides[i];
    return double_layer(a, nullptr, g, gb, s);
}
size_t acetn_b200_absorb_edge_finish_workspace_bytes(int64_t xa, int64_t xx, int64_t xy, int64_t D) {
    EdgeDims e{xa, 1, xx, xy, D, 1};
    return ws_round(frob_scratch_doubles() * 8) + gemm_workspace_bytes(e_g4(e, nullptr, nullptr, nullptr)) + 4096;
}
int acetn_b200_absorb_edge_finish(const double* proj2, const double* T3, int64_t xa, int64_t xx, int64_t xy, int64_t D, int normalize,
                                  double* out, void* wsp, size_t ws_bytes, void* stream) {
    cudaStream_t s = S_(stream);
    EdgeDims e{xa, 1, xx, xy, D, 1};
    Workspace ws(wsp, ws_bytes);
    double* fs = ws.take<double>(frob_scratch_doubles());
    if (ws.overflow) { set_error("absorb_edge_finish: workspace too small"); return ERR_WORKSPACE; }
    void* g = ws.base + ws.used; size_t gb = ws.bytes - ws.used;
    AB_TRY(gemm_launch(e_g4(e, proj2, T3, out), g, gb, s));        // out[y,(x,r,R)]
    if (!normalize) return OK;
    return frob_normalize_launch(out, (size_t)(xy * xx * D * D), fs, s);
}

double acetn_b200_fp64_peak_probe(void* scratch, int iters, void* stream) { return dmma_peak_launch((double*)scratch, iters, S_(stream)); }

size_t acetn_b200_als_workspace_bytes(int64_t nD, int64_t bD, int64_t pD) { return als_workspace_bytes((int)nD, (int)bD, (int)pD); }
int acetn_b200_als_solve(double* a1r, double* a2r, const double* n12g, const double* n12, const double* a12g, int64_t nD, int64_t bD,
                         int64_t pD, int64_t niter, double tol, double epsilon, int64_t method, int32_t* info, void* ws, size_t ws_bytes,
                         void* stream) {
    return als_solve_launch(a1r, a2r, n12g, n12, a12g, (int)nD, (int)bD, (int)pD, (int)niter, tol, epsilon, (int)method, (int*)info, ws,
                            ws_bytes, S_(stream));
}

int acetn_b200_permute(double* dst, const double* src, int nd, const int64_t* dims, const int64_t* strides, void* stream) {
    return gather_nd_launch(dst, src, nd, dims, strides, S_(stream));
}

int acetn_b200_absmax(const double* x, int64_t n, double* out, void* stream) { return absmax_launch(x, (size_t)n, out, S_(stream)); }
int acetn_b200_frob_normalize(double* x, int64_t n, void* ws, size_t ws_bytes, void* stream) {
    if (ws_bytes < frob_scratch_doubles() * 8) { set_error("frob_normalize: workspace too small"); return ERR_WORKSPACE; }
    return frob_normalize_launch(x, (size_t)n, (double*)ws, S_(stream));
}

}  // extern "C"
