// K1: FP64 tensor-core (DMMA.8x8x4) GEMM with two-level operand addressing.
// C(m,n)[b] = alpha * sum_k A(m,k)[b] * B(k,n)[b] + beta * C(m,n)[b]
// Every index group (m, n, k, batch) of every operand is a two-level index (Idx2), so tensor legs that
// the reference permutes with torch.einsum copies (Q.s2, E.1, C1/C2, P.*) are consumed and
// produced in place.
#pragma once
#include "common.cuh"

namespace ab200 {

struct GemmOperand {
    const double* ptr;
    Idx2 row;    // first logical index  (A: m, B: k)
    Idx2 col;    // second logical index (A: k, B: n)
    Idx2 batch;
};

struct GemmDesc {
    int M, N, K, batch;
    GemmOperand A;      // (m, k)
    GemmOperand B;      // (k, n)
    double* C;
    Idx2 cm, cn, cb;
    double alpha, beta;
    int force_tile;     // 0 = heuristic; 1 = 128x128, 2 = 128x88, 3 = 64x64
    int force_splitk;   // 0 = heuristic
};

static inline GemmOperand operand(const double* p, Idx2 row, Idx2 col, Idx2 batch = idx1(0)) {
    GemmOperand o; o.ptr = p; o.row = row; o.col = col; o.batch = batch; return o;
}

// plain row-major helpers
static inline GemmDesc gemm_desc(int M, int N, int K, GemmOperand A, GemmOperand B, double* C, Idx2 cm, Idx2 cn,
                                 double alpha = 1.0, double beta = 0.0, int batch = 1, Idx2 cb = idx1(0)) {
    GemmDesc d; memset(&d, 0, sizeof(d));
    d.M = M; d.N = N; d.K = K; d.batch = batch; d.A = A; d.B = B; d.C = C; d.cm = cm; d.cn = cn; d.cb = cb;
    d.alpha = alpha; d.beta = beta; d.force_tile = 0; d.force_splitk = 0;
    return d;
}

// bytes of split-K scratch the heuristic may use for this problem (0 if it will not split)
size_t gemm_workspace_bytes(const GemmDesc& d);
// enqueue on stream; ws may be nullptr if gemm_workspace_bytes()==0. Never synchronises.
int gemm_launch(const GemmDesc& d, void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace ab200
