// K7: FP64-exact "big x thin" matrix products on the INT8 tensor cores of sm_100a (tcgen05.mma kind::i8, TMEM
// accumulators, TMA-staged operands) through a residue number system.  Host launchers; kernels in i8crt.cu.
#pragma once
#include "common.cuh"

namespace ab200 {

constexpr int I8_NMOD = 16;        // moduli (crt_tables.h); product M ~ 2^125.4
constexpr int I8_MAX_THIN = 272;   // widest thin operand (columns) one accumulator tile holds

// Encoded big matrix Q (rows x cols): residues[l][i][j] = round(Q[i][j] * 2^(P - rowexp[i] - colexp[j])) mod m_l
// (balanced, int8), plane pitch rows*ld bytes, row pitch ld bytes (ld = cols rounded up to 128).
struct I8Matrix {
    const int8_t* res;
    const int32_t* rowexp;
    const int32_t* colexp;
    int64_t rows, cols, ld;
    int P;
};

static inline int64_t i8_ld(int64_t cols) { return (cols + 127) / 128 * 128; }
// bits kept per operand for a contraction length up to kmax (see i8crt.cu header for the bound)
int i8_operand_bits(int64_t kmax);
bool i8_supported(int64_t rows, int64_t cols, int64_t q);
size_t i8_encoded_bytes(int64_t rows, int64_t cols);
I8Matrix i8_view(const void* storage, int64_t rows, int64_t cols);   // the carve i8_encode_launch used            // residues + exponents, as carved by i8_encode_launch
// Q (rows x cols FP64, row pitch ldq) -> enc (I8Matrix pointing into `storage`, which must hold i8_encoded_bytes())
// colexp_ready: the column exponents (max_i exponent(Q_ij), EXP_NONE = -1000000 for an all-zero column) have already been written
// into the storage by the producer of Q (i8_colexp_reset_launch + atomic max from the producing kernel's epilogue)
int i8_encode_launch(const double* Q, int64_t rows, int64_t cols, int64_t ldq, void* storage, size_t storage_bytes, I8Matrix* enc,
                     cudaStream_t s, bool colexp_ready = false);
// colexp array inside `storage`, reset to "all zero"; a producer then raises entry j with atomicMax(exponent - 1022)
int i8_colexp_reset_launch(void* storage, int64_t rows, int64_t cols, int32_t** colexp_out, cudaStream_t s);
size_t i8_matmul_workspace_bytes(int64_t rows, int64_t cols, int64_t q);
// out (rows x q) = Q * Y (Y: cols x q)            adjoint = false
// out (cols x q) = Q^T * Y (Y: rows x q)          adjoint = true
int i8_matmul_launch(const I8Matrix& A, bool adjoint, const double* Y, int64_t q, int64_t ldy, double* out, int64_t ldo, void* ws,
                     size_t ws_bytes, cudaStream_t s);

// stage-level entry points (tests / tools)
int i8_thin_encode_launch(const double* Y, int64_t k, int64_t q, int64_t ldy, const int32_t* rowshift, int P, int8_t* res, int64_t ldk,
                          int npad, int32_t* colexp, cudaStream_t s);
int i8_gemm_launch(const int8_t* Ares, int64_t rows, int64_t cols, int64_t ld, bool adjoint, const int8_t* Bres, int64_t ldk, int npad,
                   int8_t* Cres, cudaStream_t s);
int i8_crt_launch(const int8_t* Cres, int64_t m, int npad, int64_t q, const int32_t* erow, const int32_t* ecol, int P2, double* out,
                  int64_t ldo, cudaStream_t s);

}  // namespace ab200
