// K7 -- FP64-accurate (normwise; exact products of P-bit fixed-point operands) "big x thin" products on the INT8 tensor cores
// (sm_100a tcgen05.mma kind::i8).
//
// The 12+1 thin products of a half-system projector (fused_matmul_svd_lowrank.py:33-46, projectors.py:172) multiply the
// same two 16384 x 16384 quarter tensors by 258-column matrices.  The FP64 DMMA pipe tops out at 37 TFLOP/s; the INT8
// tensor pipe of the same SM is ~120x wider.  This file evaluates  D = op(Q) Y  *to FP64 accuracy* with integer
// arithmetic only (residue number system, the "Ozaki scheme II" idea):
//
//   1. two-sided power-of-two scaling turns both operands into integer matrices of P <= 54 bits
//        a'_ij = rint(Q_ij 2^(P - r_i - c_j)),   b'_jz = rint(Y_jz 2^(c_j) 2^(P - e_z))       (|a'|,|b'| <= 2^P)
//      c_j / r_i = column / row binary exponents of Q (c_j = exponent of the largest entry of column j; r_i taken after the
//      column scaling, so r_i <= 0), e_z = column exponent of the row-shifted Y.  The only rounding of the whole product
//      happens here: 2^-P relative to the row x column scale, i.e. the result is NORMWISE FP64-accurate per row / column
//      scale, not componentwise (entries more than 2^-P below their row-and-column scale are flushed).
//   2. a', b' are reduced modulo 16 pairwise coprime moduli m_l <= 256 (a': residues in [0,m) as uint8, b': balanced
//      residues as int8): one pass over Q per site-move (i8_encode), one cheap pass over each thin operand.
//   3. per modulus an exact UINT8 x INT8 -> INT32 GEMM on the tensor cores (k * 255 * 128 < 2^31 for k <= 65535), epilogue
//      reduces the accumulator mod m_l back to int8.
//   4. Chinese remainder reconstruction in 128-bit integer arithmetic: x = sum_l t_l W_l - round(sum_l t_l/m_l) M is the
//      exact integer sum_j a'_ij b'_jz because |x| <= k 2^(2P) < M/4 (M = prod m_l ~ 2^125.4);  D_iz = x 2^(r_i + e_z - 2P).
//
// GEMM kernel: persistent, one CTA per SM, 6 warps: warp 0 = TMA producer (cp.async.bulk.tensor 3-D, SWIZZLE_128B),
// warp 1 = MMA issuer (single thread, tcgen05.mma.cta_group::1.kind::i8, M=128, N=128+144 for the 258(+pad) thin columns,
// accumulators in TMEM), warps 2-5 = epilogue (tcgen05.ld -> mod m_l -> int8 store).  The adjoint product reads the SAME
// residue planes through an MN-major shared-memory descriptor, so Q is encoded once for Q Y and Q^T Y.
#include <cuda.h>
#include <stdlib.h>

#include "crt_tables.h"
#include "i8crt.cuh"

namespace ab200 {

namespace {

__constant__ int c_mod[I8_NMOD] = CRT_MOD_INIT;
__constant__ int c_lo[I8_NMOD] = CRT_LO_INIT;
__constant__ float c_inv[I8_NMOD] = CRT_INV_INIT;
__constant__ int c_qinv[I8_NMOD] = CRT_QINV_INIT;
__constant__ uint32_t c_p256lo[I8_NMOD] = CRT_P256_LO_INIT;
__constant__ uint32_t c_p256hi[I8_NMOD] = CRT_P256_HI_INIT;
__constant__ int c_off0[I8_NMOD] = CRT_OFF0_INIT;
__constant__ int c_off1[I8_NMOD] = CRT_OFF1_INIT;
__constant__ int c_off55[I8_NMOD] = CRT_OFF55_INIT;
__constant__ int c_negmod[I8_NMOD] = CRT_NEGMOD_INIT;
__constant__ uint32_t c_magic[I8_NMOD] = CRT_MAGIC_INIT;
__constant__ unsigned long long c_wlo[I8_NMOD] = CRT_W_LO_INIT;
__constant__ unsigned long long c_whi[I8_NMOD] = CRT_W_HI_INIT;

constexpr int EXP_NONE = -1000000;     // exponent of an all-zero row / column

// ---------------------------------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
// balanced residue of v modulo m (|v| < 2^29): result in [lo, lo + m - 1]
__device__ __forceinline__ int bal_mod(int v, int m, int lo, float inv) {
    int q = __float2int_rn(__int2float_rn(v) * inv);
    int r = v - q * m;
    if (r < lo) r += m;
    else if (r > lo + m - 1) r -= m;
    return r;
}
// biased exponent field of a double (0 for zero / subnormal)
__device__ __forceinline__ int exp_field(double x) { return (int)((__double_as_longlong(x) >> 52) & 0x7ff); }
// rint(x * 2^shift) as a 64-bit integer; |result| <= 2^55 guaranteed by the caller's choice of shift
__device__ __forceinline__ long long scaled_int(double x, int shift) {
    long long bits = __double_as_longlong(x);
    int ef = (int)((bits >> 52) & 0x7ff);
    int nf = ef + shift;
    if (ef == 0 || nf < 1022) return 0;                  // |x 2^shift| < 1/2
    long long mag = (bits & 0x000fffffffffffffLL) | ((long long)nf << 52);
    long long v = __double2ll_rn(__longlong_as_double(mag));
    return bits < 0 ? -v : v;
}
// 16 residues of v (|v| <= 2^55) in [0, m_l).  The eight bytes of the two's-complement representation are reduced with two
// dp4a against 256^b mod m (signed bytes); the accumulator starts at a multiple of m that keeps the sum positive (and, for
// negative v, removes 2^64 mod m); floor division by the 32-bit reciprocal is exact for sums < 2^20.
__device__ __forceinline__ void residues16u(long long v, int r[I8_NMOD]) {
    const unsigned long long u = (unsigned long long)v;
    const uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
    const bool neg = v < 0;
#pragma unroll
    for (int l = 0; l < I8_NMOD; l++) {
        int s = dp4a_us(hi, c_p256hi[l], neg ? c_off1[l] : c_off0[l]);
        s = dp4a_us(lo, c_p256lo[l], s);
        r[l] = s - (int)__umulhi((uint32_t)s, c_magic[l]) * c_mod[l];
    }
}

// The same through v + 2^55 >= 0 (|v| <= 2^55): no sign case (the -2^55 mod m is folded into the accumulator start), and the
// residue is packed straight into byte `pos` of the plane word -- 5 instructions per residue (2 dp4a, mulhi, mul-sub, insert).
template <int POS>
__device__ __forceinline__ void residues16_pack(long long v, uint32_t w[I8_NMOD]) {
    const unsigned long long u = (unsigned long long)(v + (1LL << 55));
    const uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
#pragma unroll
    for (int l = 0; l < I8_NMOD; l++) {
        int s = dp4a_us(hi, c_p256hi[l], c_off55[l]);
        s = dp4a_us(lo, c_p256lo[l], s);
        const uint32_t r = __umulhi((uint32_t)s, c_magic[l]) * (uint32_t)c_negmod[l] + (uint32_t)s;      // s - floor(s / m) m: one IMAD
        if (POS == 0) w[l] = r;                                                        // r < 256: the upper bytes start as zero
        else asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(w[l]) : "r"(r), "r"(1u << (8 * POS)));   // disjoint bytes: add == insert
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// encode: exponents
// ---------------------------------------------------------------------------------------------------------------------
// colexp[j] = max_i exponent(Q_ij)  (|Q_ij| < 2^colexp[j]; EXP_NONE for an all-zero column); colexp pre-set to EXP_NONE.
// Only used when the producer of Q did not deliver the column exponents itself (the fused double-layer kernel does, from its
// epilogue: absorb_fused.cu).
__global__ void __launch_bounds__(256) col_max_kernel(const double* __restrict__ Q, int64_t rows, int64_t cols, int64_t ldq,
                                                      int32_t* __restrict__ colexp) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.y * 64;
    if (j >= cols) return;
    int mx = 0;
    for (int64_t i = i0; i < i0 + 64 && i < rows; i++) mx = max(mx, exp_field(Q[i * ldq + j]));
    if (mx > 0) atomicMax(colexp + j, mx - 1022);
}
__global__ void fill_i32_kernel(int32_t* p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------------
// encode: residues of the big matrix in two passes over Q (three before the column exponents came from the producer):
//   row_exp_kernel   : r_i = max_j (exponent(Q_ij) - c_j) <= 0, one CTA per row (HBM bound, 8 B/element)
//   encode_big_kernel: thread = 8 consecutive columns of one row -> 16 eight-byte stores, one per plane (issue bound)
// A fused single-pass variant (persistent CTA per row, second read of the row from L2) was measured slower (2.25 - 2.34 ms against
// 0.4 + 1.6 ms at 16384^2): with one row per CTA the exponent pass and the residue pass serialise inside the CTA.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_exp_kernel(const double* __restrict__ Q, int64_t cols, int64_t ldq, const int32_t* __restrict__ colexp,
                                                      int32_t* __restrict__ rowexp, int vec_ok) {
    const double* row = Q + (int64_t)blockIdx.x * ldq;
    int mx = EXP_NONE;
    if (vec_ok) {
        for (int64_t j0 = (int64_t)threadIdx.x * 2; j0 < cols; j0 += 2048) {
            double2 t[4];
            int2 c[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t j = j0 + 512 * u;
                if (j + 2 <= cols) {
                    t[u] = __ldg(reinterpret_cast<const double2*>(row + j));
                    c[u] = __ldg(reinterpret_cast<const int2*>(colexp + j));
                } else {
                    t[u].x = j < cols ? row[j] : 0.0; t[u].y = 0.0;
                    c[u].x = j < cols ? colexp[j] : 0; c[u].y = 0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e0 = exp_field(t[u].x), e1 = exp_field(t[u].y);
                if (e0) mx = max(mx, e0 - 1022 - c[u].x);
                if (e1) mx = max(mx, e1 - 1022 - c[u].y);
            }
        }
    } else {
        for (int64_t j = threadIdx.x; j < cols; j += 256) { int e = exp_field(row[j]); if (e) mx = max(mx, e - 1022 - colexp[j]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __shared__ int sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) mx = max(mx, sm[w]);
        rowexp[blockIdx.x] = mx;                              // EXP_NONE for an all-zero row
    }
}

__global__ void __launch_bounds__(256) encode_big_kernel(const double* __restrict__ Q, int64_t rows, int64_t cols, int64_t ldq,
                                                         const int32_t* __restrict__ rowexp, const int32_t* __restrict__ colexp, int P,
                                                         int8_t* __restrict__ res, int64_t ld, int vec_ok) {
    const int64_t i = blockIdx.y;
    const int64_t j0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
    if (j0 >= ld) return;
    const int re = rowexp[i];
    const double* row = Q + i * ldq;
    double x[8];
    int ce[8];
    if (vec_ok && j0 + 8 <= cols) {
        const double2* src = reinterpret_cast<const double2*>(row + j0);
        const int4* csrc = reinterpret_cast<const int4*>(colexp + j0);
#pragma unroll
        for (int c = 0; c < 4; c++) { double2 t = __ldcs(src + c); x[2 * c] = t.x; x[2 * c + 1] = t.y; }
        int4 c0 = __ldg(csrc), c1 = __ldg(csrc + 1);
        ce[0] = c0.x; ce[1] = c0.y; ce[2] = c0.z; ce[3] = c0.w; ce[4] = c1.x; ce[5] = c1.y; ce[6] = c1.z; ce[7] = c1.w;
    } else {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int64_t j = j0 + c;
            x[c] = j < cols ? row[j] : 0.0;
            ce[c] = j < cols ? colexp[j] : EXP_NONE;
        }
    }
    uint32_t w0[I8_NMOD], w1[I8_NMOD];
    long long v[8];
#pragma unroll
    for (int c = 0; c < 8; c++) v[c] = (re > EXP_NONE && ce[c] > EXP_NONE) ? scaled_int(x[c], P - re - ce[c]) : 0;
    residues16_pack<0>(v[0], w0); residues16_pack<1>(v[1], w0); residues16_pack<2>(v[2], w0); residues16_pack<3>(v[3], w0);
    residues16_pack<0>(v[4], w1); residues16_pack<1>(v[5], w1); residues16_pack<2>(v[6], w1); residues16_pack<3>(v[7], w1);
    const int64_t plane = rows * ld;
#pragma unroll
    for (int l = 0; l < I8_NMOD; l++)
        __stcs(reinterpret_cast<uint2*>(res + l * plane + i * ld + j0), make_uint2(w0[l], w1[l]));
}

// ---------------------------------------------------------------------------------------------------------------------
// thin operand Y (k x q): column exponents of the row-shifted matrix, then residues written K-major: res[l][z][j]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288) thin_colexp_kernel(const double* __restrict__ Y, int64_t k, int q, int64_t ldy,
                                                          const int32_t* __restrict__ rowshift, int32_t* __restrict__ colexp) {
    const int z = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * 128;
    if (z >= q) return;
    int mx = EXP_NONE;
    for (int64_t j = j0; j < j0 + 128 && j < k; j++) {
        int rs = rowshift[j];
        int ef = exp_field(Y[j * ldy + z]);
        if (ef != 0 && rs > EXP_NONE) mx = max(mx, ef - 1022 + rs);
    }
    if (mx > EXP_NONE) atomicMax(colexp + z, mx);
}
// tile = 16 columns (z) x 128 rows (j); 256 threads
__global__ void __launch_bounds__(256) thin_encode_kernel(const double* __restrict__ Y, int64_t k, int q, int64_t ldy,
                                                          const int32_t* __restrict__ rowshift, const int32_t* __restrict__ colexp, int P,
                                                          int8_t* __restrict__ res, int64_t ldk, int npad) {
    __shared__ long long sv[16][129];
    const int z0 = blockIdx.y * 16;
    const int64_t j0 = (int64_t)blockIdx.x * 128;
    {
        const int tz = threadIdx.x & 15, tj = threadIdx.x >> 4;
        const int z = z0 + tz;
        const int ce = z < q ? colexp[z] : EXP_NONE;
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int jj = tj + 16 * it;
            const int64_t j = j0 + jj;
            long long v = 0;
            if (j < k && ce > EXP_NONE) {
                int rs = rowshift[j];
                if (rs > EXP_NONE) v = scaled_int(Y[j * ldy + z], P - ce + rs);
            }
            sv[tz][jj] = v;
        }
    }
    __syncthreads();
    {
        const int tz = threadIdx.x >> 4, jb = (threadIdx.x & 15) * 8;
        uint32_t w0[I8_NMOD], w1[I8_NMOD];
#pragma unroll
        for (int l = 0; l < I8_NMOD; l++) { w0[l] = 0; w1[l] = 0; }
#pragma unroll
        for (int c = 0; c < 8; c++) {
            int r[I8_NMOD];
            residues16u(sv[tz][jb + c], r);
#pragma unroll
            for (int l = 0; l < I8_NMOD; l++) {
                const int rb = r[l] > c_lo[l] + c_mod[l] - 1 ? r[l] - c_mod[l] : r[l];      // balanced: the B operand is signed
                uint32_t b = (uint32_t)(rb & 0xff) << (8 * (c & 3));
                if (c < 4) w0[l] |= b; else w1[l] |= b;
            }
        }
        const int64_t plane = (int64_t)npad * ldk;
        if (j0 + jb < ldk) {
#pragma unroll
            for (int l = 0; l < I8_NMOD; l++)
                *reinterpret_cast<uint2*>(res + l * plane + (int64_t)(z0 + tz) * ldk + j0 + jb) = make_uint2(w0[l], w1[l]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier, TMA, tcgen05)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((unsigned long long)map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], u8 x s8 -> s32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t v[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;                         // descriptor version (sm_100)
    d |= 2ull << 61;                         // LayoutType::SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): s32 accumulate, A = u8 (residues in [0,m)), B = s8 (balanced), M = 128
__host__ __device__ constexpr uint32_t i8_idesc(int n, int a_mn_major) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct GemmParams {
    int8_t* out;          // [NMOD][m_out][npad]
    int64_t m_out;
    int mtiles, nk, trans;
    int npad, n1, n2;     // n1 + n2 = npad, each a multiple of 16 (n2 may be 0)
    int stages, stage_bytes;
    int b_box_rows, b_loads;
};

constexpr int GEMM_THREADS = 192;
constexpr int A_TILE_BYTES = 128 * 128;
constexpr int TMEM_COLS = 512;

__global__ void __launch_bounds__(GEMM_THREADS, 1)
i8_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + (uint32_t)p.stages * (uint32_t)p.stage_bytes;
    // barriers: full[s] at +16 s, empty[s] at +16 s + 8, tmem_full, tmem_empty, tmem base holder
    const uint32_t bar_tmem_full = bar_base + 16u * (uint32_t)p.stages;
    const uint32_t bar_tmem_empty = bar_tmem_full + 8;
    const uint32_t tmem_holder = bar_tmem_empty + 8;
    uint32_t* tmem_holder_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_holder - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) {
            mbar_init(bar_base + 16u * s, 1);
            mbar_init(bar_base + 16u * s + 8, 1);
        }
        mbar_init(bar_tmem_full, 1);
        mbar_init(bar_tmem_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_holder), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder_ptr;
    const int total = I8_NMOD * p.mtiles;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            int s = 0;
            uint32_t ph = 0;
            const uint32_t tx = (uint32_t)(A_TILE_BYTES + p.npad * 128);
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                const int l = item / p.mtiles, mt = item - l * p.mtiles;
                for (int kc = 0; kc < p.nk; kc++) {
                    const uint32_t full = bar_base + 16u * s, empty = full + 8;
                    mbar_wait(empty, ph ^ 1);
                    mbar_arrive_expect_tx(full, tx);
                    const uint32_t sa = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                    if (!p.trans) tma_load_3d(sa, &tmA, full, kc * 128, mt * 128, l);
                    else tma_load_3d(sa, &tmA, full, mt * 128, kc * 128, l);
                    for (int b = 0; b < p.b_loads; b++)
                        tma_load_3d(sa + A_TILE_BYTES + (uint32_t)(b * p.b_box_rows * 128), &tmB, full, kc * 128, b * p.b_box_rows, l);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc1 = i8_idesc(p.n1, p.trans);
            const uint32_t idesc2 = i8_idesc(p.n2 > 0 ? p.n2 : 16, p.trans);
            const uint32_t a_kstep = p.trans ? (32u * 128u) >> 4 : 32u >> 4;      // descriptor units (16 B) per 32-deep MMA
            int s = 0;
            uint32_t ph = 0, tph = 0;
            for (int item = blockIdx.x; item < total; item += gridDim.x) {
                mbar_wait(bar_tmem_empty, tph ^ 1);
                tc_fence_after();
                for (int kc = 0; kc < p.nk; kc++) {
                    const uint32_t full = bar_base + 16u * s, empty = full + 8;
                    mbar_wait(full, ph);
                    tc_fence_after();
                    const uint32_t sa = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                    const uint32_t sb = sa + A_TILE_BYTES;
                    // A: K-major rows of 128 B (8-row groups 1024 B apart)  |  MN-major: 128 B of M per K row
                    const uint64_t da = smem_desc(sa, p.trans ? 16384u : 16u, 1024u);
                    const uint64_t db1 = smem_desc(sb, 16u, 1024u);
                    const uint64_t db2 = smem_desc(sb + (uint32_t)p.n1 * 128u, 16u, 1024u);
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) {
                        const uint32_t acc = (kc | k4) != 0;
                        tc_mma_i8(tmem_base, da + (uint64_t)(a_kstep * k4), db1 + (uint64_t)(2 * k4), idesc1, acc);
                        if (p.n2 > 0)
                            tc_mma_i8(tmem_base + (uint32_t)p.n1, da + (uint64_t)(a_kstep * k4), db2 + (uint64_t)(2 * k4), idesc2, acc);
                    }
                    tc_commit(empty);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                tc_commit(bar_tmem_full);
                tph ^= 1;
            }
        }
    } else {
        // ===== epilogue: TMEM -> mod m_l -> int8 =====
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        uint32_t tph = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x) {
            const int l = item / p.mtiles, mt = item - l * p.mtiles;
            const int m = c_mod[l], lo = c_lo[l];
            const float inv = c_inv[l];
            mbar_wait(bar_tmem_full, tph);
            tc_fence_after();
            const int64_t row = (int64_t)mt * 128 + quarter * 32 + lane;
            int8_t* orow = p.out + ((int64_t)l * p.m_out + row) * p.npad;
            for (int c0 = 0; c0 < p.npad; c0 += 16) {
                uint32_t v[16];
                tc_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    int r = bal_mod((int)v[c], m, lo, inv);
                    w[c >> 2] |= (uint32_t)(r & 0xff) << (8 * (c & 3));
                }
                if (row < p.m_out) *reinterpret_cast<uint4*>(orow + c0) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tmem_empty);
            tph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// CRT reconstruction.  Thread = 4 consecutive columns of one output row.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) crt_kernel(const int8_t* __restrict__ res, int64_t m, int npad, int q, const int32_t* __restrict__ erow,
                                                  const int32_t* __restrict__ ecol, int P2, double* __restrict__ out, int64_t ldo) {
    const int groups = npad >> 2;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i = idx / groups;
    const int zg = (int)(idx - i * groups);
    if (i >= m || zg * 4 >= q) return;
    uint32_t w[I8_NMOD];
    const int64_t plane = m * (int64_t)npad;
#pragma unroll
    for (int l = 0; l < I8_NMOD; l++) w[l] = *reinterpret_cast<const uint32_t*>(res + l * plane + i * npad + zg * 4);
    const int er = erow[i];
    const unsigned __int128 Mfull = ((unsigned __int128)CRT_M_HI << 64) | (unsigned __int128)CRT_M_LO;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int z = zg * 4 + c;
        if (z >= q) break;
        unsigned __int128 acc = 0;
        float f = 0.f;
#pragma unroll
        for (int l = 0; l < I8_NMOD; l++) {
            const int r = (int)(int8_t)((w[l] >> (8 * c)) & 0xff);
            const int mm = c_mod[l];
            int u = r * c_qinv[l];
            int t = u - __float2int_rd(__int2float_rn(u) * c_inv[l]) * mm;
            if (t < 0) t += mm;
            else if (t >= mm) t -= mm;
            const unsigned __int128 W = ((unsigned __int128)c_whi[l] << 64) | (unsigned __int128)c_wlo[l];
            acc += W * (unsigned __int128)(unsigned)t;
            f += (float)t * c_inv[l];
        }
        const int kq = __float2int_rn(f);
        acc -= Mfull * (unsigned __int128)(unsigned)kq;
        const __int128 x = (__int128)acc;
        const long long hi = (long long)(x >> 64);
        const unsigned long long lo = (unsigned long long)x;
        double d = __ll2double_rn(hi) * 18446744073709551616.0 + __ull2double_rn(lo);
        const int ec = ecol[z];
        double val = 0.0;
        if (er > EXP_NONE && ec > EXP_NONE) val = scalbn(d, er + ec - P2);
        out[i * ldo + z] = val;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// 3-D uint8 map over [planes][rows][ld]: dims (inner -> outer) = {cols, rows, planes}; box = {128, box_rows, 1}
int make_map(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t ld, int planes, int box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) { set_error("i8: cuTensorMapEncodeTiled is not available from the driver"); return ERR_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)ld, (cuuint64_t)(ld * rows)};
    cuuint32_t box[3] = {128u, (cuuint32_t)box_rows, 1u};
    cuuint32_t es[3] = {1u, 1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("i8: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return ERR_CUDA; }
    return OK;
}
inline int npad_of(int64_t q) { return (int)((q + 15) / 16 * 16); }
inline int env_int(const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; }

}  // namespace

I8Matrix i8_view(const void* storage, int64_t rows, int64_t cols) {
    Workspace ws(const_cast<void*>(storage), (size_t)-1);
    I8Matrix e;
    e.ld = i8_ld(cols);
    e.res = ws.take<int8_t>((size_t)I8_NMOD * rows * e.ld);
    e.rowexp = ws.take<int32_t>((size_t)rows);
    e.colexp = ws.take<int32_t>((size_t)cols);
    e.rows = rows; e.cols = cols;
    e.P = i8_operand_bits(rows > cols ? rows : cols);
    return e;
}

int i8_operand_bits(int64_t kmax) {
    int lg = 0;
    while (((int64_t)1 << lg) < kmax) lg++;
    int P = (int)((CRT_LOG2_M - 2.0 - lg) / 2.0);      // k 2^(2P) <= M / 4
    return P > 54 ? 54 : P;
}
bool i8_supported(int64_t rows, int64_t cols, int64_t q) {
    return q >= 1 && q <= I8_MAX_THIN && rows >= 128 && cols >= 128 && rows <= 65535 && cols <= 65535;
}
size_t i8_encoded_bytes(int64_t rows, int64_t cols) {
    return ws_round((size_t)I8_NMOD * (size_t)rows * (size_t)i8_ld(cols)) + ws_round((size_t)rows * 4) + ws_round((size_t)cols * 4) + 1024;
}

void i8_carve(void* storage, int64_t rows, int64_t cols, int8_t** res, int32_t** rowexp, int32_t** colexp) {
    I8Matrix e = i8_view(storage, rows, cols);
    *res = const_cast<int8_t*>(e.res); *rowexp = const_cast<int32_t*>(e.rowexp); *colexp = const_cast<int32_t*>(e.colexp);
}

int i8_colexp_reset_launch(void* storage, int64_t rows, int64_t cols, int32_t** colexp_out, cudaStream_t s) {
    int8_t* res; int32_t *rowexp, *colexp;
    i8_carve(storage, rows, cols, &res, &rowexp, &colexp);
    fill_i32_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, s>>>(colexp, (int)cols, EXP_NONE);
    AB_LAUNCHED();
    *colexp_out = colexp;
    return OK;
}

int i8_encode_launch(const double* Q, int64_t rows, int64_t cols, int64_t ldq, void* storage, size_t storage_bytes, I8Matrix* enc,
                     cudaStream_t s, bool colexp_ready) {
    AB_REQUIRE(i8_supported(rows, cols, 1), "i8_encode: unsupported shape %lld x %lld", (long long)rows, (long long)cols);
    AB_REQUIRE(storage != nullptr && storage_bytes >= i8_encoded_bytes(rows, cols), "i8_encode: storage too small (%zu needed, %zu given)",
               i8_encoded_bytes(rows, cols), storage_bytes);
    int8_t* res; int32_t *rowexp, *colexp;
    i8_carve(storage, rows, cols, &res, &rowexp, &colexp);
    const int64_t ld = i8_ld(cols);
    const int P = i8_operand_bits(rows > cols ? rows : cols);
    if (!colexp_ready) {
        // no producer-side column exponents: one extra pass over Q
        fill_i32_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, s>>>(colexp, (int)cols, EXP_NONE);
        AB_LAUNCHED();
        col_max_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)((rows + 63) / 64)), 256, 0, s>>>(Q, rows, cols, ldq, colexp);
        AB_LAUNCHED();
    }
    const int vec_ok = (ldq % 2 == 0) && (((uintptr_t)Q & 15) == 0) && (((uintptr_t)colexp & 15) == 0);
    row_exp_kernel<<<(unsigned)rows, 256, 0, s>>>(Q, cols, ldq, colexp, rowexp, vec_ok);
    AB_LAUNCHED();
    encode_big_kernel<<<dim3((unsigned)((ld / 8 + 255) / 256), (unsigned)rows), 256, 0, s>>>(Q, rows, cols, ldq, rowexp, colexp, P, res, ld, vec_ok);
    AB_LAUNCHED();
    enc->res = res; enc->rowexp = rowexp; enc->colexp = colexp; enc->rows = rows; enc->cols = cols; enc->ld = ld; enc->P = P;
    return OK;
}

int i8_thin_encode_launch(const double* Y, int64_t k, int64_t q, int64_t ldy, const int32_t* rowshift, int P, int8_t* res, int64_t ldk,
                          int npad, int32_t* colexp, cudaStream_t s) {
    AB_REQUIRE(q <= I8_MAX_THIN && npad % 16 == 0 && npad >= q && ldk % 128 == 0 && ldk >= k, "i8_thin_encode: bad shape");
    fill_i32_kernel<<<(npad + 255) / 256, 256, 0, s>>>(colexp, npad, EXP_NONE);
    AB_LAUNCHED();
    thin_colexp_kernel<<<(unsigned)((k + 127) / 128), 288, 0, s>>>(Y, k, (int)q, ldy, rowshift, colexp);
    AB_LAUNCHED();
    thin_encode_kernel<<<dim3((unsigned)(ldk / 128), (unsigned)(npad / 16)), 256, 0, s>>>(Y, k, (int)q, ldy, rowshift, colexp, P, res, ldk, npad);
    AB_LAUNCHED();
    return OK;
}

int i8_gemm_launch(const int8_t* Ares, int64_t rows, int64_t cols, int64_t ld, bool adjoint, const int8_t* Bres, int64_t ldk, int npad,
                   int8_t* Cres, cudaStream_t s) {
    AB_REQUIRE(npad % 16 == 0 && npad >= 16 && npad <= I8_MAX_THIN, "i8_gemm: npad must be a multiple of 16 in [16, %d]", I8_MAX_THIN);
    const int64_t m_out = adjoint ? cols : rows, kdim = adjoint ? rows : cols;
    AB_REQUIRE(kdim <= 65535, "i8_gemm: contraction length %lld exceeds the INT32 accumulator bound", (long long)kdim);
    GemmParams p;
    p.out = Cres; p.m_out = m_out; p.nk = (int)((kdim + 127) / 128); p.trans = adjoint ? 1 : 0;
    p.npad = npad;
    const int sms = device_sm_count();
    const int smem_budget = 227 * 1024 - 1024 - 256;
    CUtensorMap tmA, tmB;
    AB_TRY(make_map(&tmA, Ares, cols, rows, ld, I8_NMOD, 128));
    p.mtiles = (int)((m_out + 127) / 128);
    if (npad <= 256) { p.n1 = npad; p.n2 = 0; p.b_box_rows = npad; p.b_loads = 1; }
    else { p.n1 = 128; p.n2 = npad - 128; p.b_box_rows = npad / 2; p.b_loads = 2; }
    // tuning knobs (dev tools only): ACETN_B200_I8_N1 = columns of the first MMA when the tile needs two, ACETN_B200_I8_STAGES.
    // Measured (profiles/r02_k7_stage_probe.txt): 2 / 3 / 4 stages = 1.67 / 1.24 / 1.18 ms per product -- four stages (all that fit)
    // cover the load latency.  A CTA-pair variant (tcgen05.mma.cta_group::2, half of the thin-operand tile per CTA, 6 stages) measured
    // the same 1.19 ms: the kernel is bound by the INT8 pipe under the power cap, not by shared-memory fill; it was removed.
    static const int dbg_n1 = env_int("ACETN_B200_I8_N1"), dbg_stages = env_int("ACETN_B200_I8_STAGES");
    if (p.n2 > 0 && dbg_n1 >= 16 && dbg_n1 <= 256 && dbg_n1 % 16 == 0 && npad - dbg_n1 >= 16) { p.n1 = dbg_n1; p.n2 = npad - dbg_n1; }
    p.stage_bytes = (A_TILE_BYTES + npad * 128 + 1023) / 1024 * 1024;
    p.stages = smem_budget / p.stage_bytes;
    if (p.stages > 8) p.stages = 8;
    if (dbg_stages >= 2 && dbg_stages < p.stages) p.stages = dbg_stages;
    AB_REQUIRE(p.stages >= 2, "i8_gemm: tile does not fit shared memory");
    const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + 1024 + 256;
    AB_TRY(make_map(&tmB, Bres, kdim, npad, ldk, I8_NMOD, p.b_box_rows));
    AB_ENSURE_SMEM(i8_gemm_kernel, smem_bytes);
    int grid = I8_NMOD * p.mtiles;
    if (grid > sms) grid = sms;
    i8_gemm_kernel<<<grid, GEMM_THREADS, smem_bytes, s>>>(tmA, tmB, p);
    AB_LAUNCHED();
    return OK;
}

int i8_crt_launch(const int8_t* Cres, int64_t m, int npad, int64_t q, const int32_t* erow, const int32_t* ecol, int P2, double* out,
                  int64_t ldo, cudaStream_t s) {
    const int64_t threads = m * (npad / 4);
    crt_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(Cres, m, npad, (int)q, erow, ecol, P2, out, ldo);
    AB_LAUNCHED();
    return OK;
}

size_t i8_matmul_workspace_bytes(int64_t rows, int64_t cols, int64_t q) {
    const int npad = npad_of(q);
    const int64_t mx = rows > cols ? rows : cols;
    return ws_round((size_t)I8_NMOD * npad * (size_t)i8_ld(mx)) + ws_round((size_t)I8_NMOD * (size_t)mx * npad) + ws_round((size_t)npad * 4) + 1024;
}

int i8_matmul_launch(const I8Matrix& A, bool adjoint, const double* Y, int64_t q, int64_t ldy, double* out, int64_t ldo, void* wsp,
                     size_t ws_bytes, cudaStream_t s) {
    AB_REQUIRE(i8_supported(A.rows, A.cols, q), "i8_matmul: unsupported shape");
    const int npad = npad_of(q);
    const int64_t kdim = adjoint ? A.rows : A.cols, m_out = adjoint ? A.cols : A.rows;
    const int64_t ldk = i8_ld(kdim);
    Workspace ws(wsp, ws_bytes);
    int8_t* Bres = ws.take<int8_t>((size_t)I8_NMOD * npad * ldk);
    int8_t* Cres = ws.take<int8_t>((size_t)I8_NMOD * m_out * npad);
    int32_t* ez = ws.take<int32_t>((size_t)npad);
    if (ws.overflow) { set_error("i8_matmul: workspace too small (%zu needed, %zu given)", ws.used, ws_bytes); return ERR_WORKSPACE; }
    // Q Y: rows of Y carry the column exponents of Q;  Q^T Y: rows of Y carry the row exponents of Q
    AB_TRY(i8_thin_encode_launch(Y, kdim, q, ldy, adjoint ? A.rowexp : A.colexp, A.P, Bres, ldk, npad, ez, s));
    AB_TRY(i8_gemm_launch(A.res, A.rows, A.cols, A.ld, adjoint, Bres, ldk, npad, Cres, s));
    AB_TRY(i8_crt_launch(Cres, m_out, npad, q, adjoint ? A.colexp : A.rowexp, ez, 2 * A.P, out, ldo, s));
    return OK;
}

}  // namespace ab200
