// K1: FP64 DMMA GEMM for sm_100a.  See gemm.cuh.
//
// Design (B200): FP64 has no tcgen05/UMMA kind; the native FP64 tensor instruction on sm_100a is the warp-level
// DMMA.8x8x4 (measured 37.2 TFLOP/s chip peak, 16 issue cycles per SM sub-partition, profiles/r01_fp64_peak_microbench.txt).
// Accumulators live in registers; operand tiles are staged through shared memory by a multi-stage cp.async (LDGSTS)
// pipeline in the orientation they have in HBM (k-contiguous or m/n-contiguous), with +4-double row padding that makes
// every 8-byte fragment load conflict free.  The index permutations of the CTMRG contractions are folded into the
// loader/epilogue address computation (Idx2), so no operand is ever transposed through HBM.
//
// Main loop (per 16- or 32-wide k tile): wait for the tile, one __syncthreads, then 4 (8) k-steps of
// [prefetch next k-step's fragments] + MT*NTL DMMAs; the 16-byte copies of the tile STAGES-1 ahead are issued one per
// n-tile slot *between* the DMMAs (pointer-increment addressing, ~4 instructions per copy), so there is no loader phase
// during which the tensor pipe idles.  Measured on the thin GEMM 16384x258x16384 (split-K 5): 87 % tensor-pipe active,
// 32.3 TFLOP/s (cuBLAS DGEMM through torch: 31.2), DRAM traffic 1.1x algorithmic (profiles/).
#include "gemm.cuh"

namespace ab200 {

constexpr int BK_DEFAULT = 16;
// BK = 16 -> 4 stages, BK = 32 -> 3 stages (shared-memory budget); k-contiguous smem rows are padded by 4 doubles

struct GemmKernelParams {
    int M, N, K, batch, splitk, tiles_n, tiles_mn, kt_total;
    GemmOperand A, B;
    double* C;
    Idx2 cm, cn, cb;
    double alpha, beta;
    double* ws;          // split-K partials [batch*splitk][M][N] (dense) when splitk > 1
    int a_vec, b_vec;    // 16-byte loads allowed along the contiguous index
    int fast;            // both k indices single-level and both operands vectorisable: pointer-increment loader
    int c_vec;           // adjacent output columns are adjacent in memory and 16-byte aligned: paired stores
};

template <int BM, int BN, bool A_MC, bool B_KC, int BK>
struct SmemLayout {
    static constexpr int STAGES = BK == 16 ? 4 : 3;
    static constexpr int PADK = BK + 4;
    static constexpr int A_LD = A_MC ? (BM + 4) : PADK;
    static constexpr int A_ROWS = A_MC ? BK : BM;
    static constexpr int B_LD = B_KC ? PADK : (BN + 4);
    static constexpr int B_ROWS = B_KC ? BN : BK;
    static constexpr int A_ELEMS = A_ROWS * A_LD;
    static constexpr int B_ELEMS = B_ROWS * B_LD;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double);
};

// Epilogue store of one warp's accumulators: C = alpha * acc + beta * C through the two-level output descriptors.
// DMUL / DSETP run on the same FP64 datapath as the DMMAs of the co-resident warps and queue behind them (the epilogue was
// 15-20 % of a warp's life in the K = chi class, profiles/r02_ncu_k1_epilogue.txt), so the common alpha = 1, beta = 0 case
// issues no FP64 instruction at all, and adjacent column pairs go out as one 16-byte store when the n index is contiguous.
template <int MT, int NTL>
__device__ __forceinline__ void store_accumulators(const GemmKernelParams& p, double (&acc)[MT][NTL][2], int b, int mw, int nw,
                                                   int lr, int lc) {
    double* Cb = p.C + p.cb.off(b);
    const bool has_beta = p.beta != 0.0;
    const bool scale = p.alpha != 1.0;
    if (p.c_vec) {
        int64_t noff[NTL];
        bool pair[NTL];
#pragma unroll
        for (int j = 0; j < NTL; j++) {
            int n = nw + j * 8 + 2 * lc;
            noff[j] = n < p.N ? p.cn.off(n) : -1;
            pair[j] = n + 1 < p.N;
        }
#pragma unroll
        for (int i = 0; i < MT; i++) {
            int m = mw + i * 8 + lr;
            if (m >= p.M) continue;
            double* row = Cb + p.cm.off(m);
#pragma unroll
            for (int j = 0; j < NTL; j++) {
                if (noff[j] < 0) continue;
                double* dst = row + noff[j];
                double v0 = acc[i][j][0], v1 = acc[i][j][1];
                if (scale) { v0 *= p.alpha; v1 *= p.alpha; }
                if (pair[j]) {
                    if (has_beta) { double2 o = *reinterpret_cast<const double2*>(dst); v0 += p.beta * o.x; v1 += p.beta * o.y; }
                    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                } else {
                    if (has_beta) v0 += p.beta * (*dst);
                    *dst = v0;
                }
            }
        }
        return;
    }
    int64_t noff[NTL][2];
#pragma unroll
    for (int j = 0; j < NTL; j++) {
        int n = nw + j * 8 + 2 * lc;
        noff[j][0] = n < p.N ? p.cn.off(n) : -1;
        noff[j][1] = n + 1 < p.N ? p.cn.off(n + 1) : -1;
    }
#pragma unroll
    for (int i = 0; i < MT; i++) {
        int m = mw + i * 8 + lr;
        if (m >= p.M) continue;
        double* row = Cb + p.cm.off(m);
#pragma unroll
        for (int j = 0; j < NTL; j++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if (noff[j][e] < 0) continue;
                double* dst = row + noff[j][e];
                double v = acc[i][j][e];
                if (scale) v *= p.alpha;
                if (has_beta) v += p.beta * (*dst);
                *dst = v;
            }
        }
    }
}

template <int BM, int BN, int WM, int WN, bool A_MC, bool B_KC, int BK>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, 1)
dgemm_dmma_kernel(const GemmKernelParams p) {
    using L = SmemLayout<BM, BN, A_MC, B_KC, BK>;
    constexpr int STAGES = L::STAGES;
    constexpr int NWARP_N = BN / WN;
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int MT = WM / 8, NTL = WN / 8;

    extern __shared__ __align__(16) double smem[];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int wm0 = (warp / NWARP_N) * WM;
    const int wn0 = (warp % NWARP_N) * WN;

    const int bz = blockIdx.x / p.tiles_mn;
    const int tile = blockIdx.x - bz * p.tiles_mn;
    const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
    const int m0 = tm * BM, n0 = tn * BN;
    const int b = bz / p.splitk, z = bz - b * p.splitk;
    const int kt_begin = (int)(((long long)p.kt_total * z) / p.splitk);
    const int kt_end = (int)(((long long)p.kt_total * (z + 1)) / p.splitk);
    const int nkt = kt_end - kt_begin;

    const double* __restrict__ Ab = p.A.ptr + p.A.batch.off(b);
    const double* __restrict__ Bb = p.B.ptr + p.B.batch.off(b);

    // ---- loader bookkeeping -------------------------------------------------------------------------------
    // A tile: !A_MC: rows = m (BM), cols = k (BK) ; A_MC: rows = k (BK), cols = m (BM)
    constexpr int A_ROWS = L::A_ROWS, A_COLS = A_MC ? BM : BK, A_CPR = A_COLS / 2;
    constexpr int A_CH = (A_ROWS * A_CPR + NT - 1) / NT;
    constexpr int B_ROWS = L::B_ROWS, B_COLS = B_KC ? BK : BN, B_CPR = B_COLS / 2;
    constexpr int B_CH = (B_ROWS * B_CPR + NT - 1) / NT;

    // general (slow) loader: every address is recomputed from the two-level descriptors; used for odd shapes,
    // non-vectorisable operands, two-level k indices and the k-remainder tile
    auto load_tile = [&](int stage, int kt) {
        double* As = smem + (size_t)stage * L::STAGE_ELEMS;
        double* Bs = As + L::A_ELEMS;
        const int k0 = kt * BK;
#pragma unroll 1
        for (int i = 0; i < A_CH; i++) {
            int c = tid + i * NT;
            if (c >= A_ROWS * A_CPR) continue;
            int r = c / A_CPR, cp = c - r * A_CPR;
            double* dst = As + r * L::A_LD + 2 * cp;
            if (A_MC) {
                int m = m0 + 2 * cp, k = k0 + r;
                bool kok = k < p.K;
                int nv = kok ? (m + 1 < p.M ? 2 : (m < p.M ? 1 : 0)) : 0;
                int64_t ko = kok ? p.A.col.off(k) : 0;
                int64_t f0 = nv >= 1 ? p.A.row.off(m) : 0;
                if (p.a_vec) {
                    cp_async16(dst, Ab + f0 + ko, nv * 8);
                } else {
                    int64_t f1 = nv >= 2 ? p.A.row.off(m + 1) : 0;
                    cp_async8(dst, Ab + f0 + ko, nv >= 1 ? 8 : 0);
                    cp_async8(dst + 1, Ab + f1 + ko, nv >= 2 ? 8 : 0);
                }
            } else {
                int m = m0 + r, k = k0 + 2 * cp;
                int nv = (m < p.M) ? (k + 1 < p.K ? 2 : (k < p.K ? 1 : 0)) : 0;
                int64_t f0 = nv ? p.A.row.off(m) : 0;
                if (p.a_vec) {
                    int64_t ko = nv ? p.A.col.off(k) : 0;
                    cp_async16(dst, Ab + f0 + ko, nv * 8);
                } else {
                    int64_t ko0 = nv >= 1 ? p.A.col.off(k) : 0;
                    int64_t ko1 = nv >= 2 ? p.A.col.off(k + 1) : 0;
                    cp_async8(dst, Ab + f0 + ko0, nv >= 1 ? 8 : 0);
                    cp_async8(dst + 1, Ab + f0 + ko1, nv >= 2 ? 8 : 0);
                }
            }
        }
#pragma unroll 1
        for (int i = 0; i < B_CH; i++) {
            int c = tid + i * NT;
            if (c >= B_ROWS * B_CPR) continue;
            int r = c / B_CPR, cp = c - r * B_CPR;
            double* dst = Bs + r * L::B_LD + 2 * cp;
            if (B_KC) {
                int n = n0 + r, k = k0 + 2 * cp;
                int nv = (n < p.N) ? (k + 1 < p.K ? 2 : (k < p.K ? 1 : 0)) : 0;
                int64_t f0 = nv ? p.B.col.off(n) : 0;
                if (p.b_vec) {
                    int64_t ko = nv ? p.B.row.off(k) : 0;
                    cp_async16(dst, Bb + f0 + ko, nv * 8);
                } else {
                    int64_t ko0 = nv >= 1 ? p.B.row.off(k) : 0;
                    int64_t ko1 = nv >= 2 ? p.B.row.off(k + 1) : 0;
                    cp_async8(dst, Bb + f0 + ko0, nv >= 1 ? 8 : 0);
                    cp_async8(dst + 1, Bb + f0 + ko1, nv >= 2 ? 8 : 0);
                }
            } else {
                int n = n0 + 2 * cp, k = k0 + r;
                bool kok = k < p.K;
                int nv = kok ? (n + 1 < p.N ? 2 : (n < p.N ? 1 : 0)) : 0;
                int64_t ko = kok ? p.B.row.off(k) : 0;
                int64_t f0 = nv >= 1 ? p.B.col.off(n) : 0;
                if (p.b_vec) {
                    cp_async16(dst, Bb + f0 + ko, nv * 8);
                } else {
                    int64_t f1 = nv >= 2 ? p.B.col.off(n + 1) : 0;
                    cp_async8(dst, Bb + f0 + ko, nv >= 1 ? 8 : 0);
                    cp_async8(dst + 1, Bb + f1 + ko, nv >= 2 ? 8 : 0);
                }
            }
        }
    };

    // ---- fast loader state (p.fast): one global pointer + one smem byte offset + byte count per 16-byte chunk; the k
    //      advance is a single add per tile, so the steady-state loop issues ~4 instructions per chunk ------------------
    const double* a_gp[A_CH];
    uint32_t a_so[A_CH];
    int a_nb[A_CH];
    const double* b_gp[B_CH];
    uint32_t b_so[B_CH];
    int b_nb[B_CH];
    // k advance per tile: single-level k: BK * stride; two-level k whose div divides BK (and hence the split-K base): the low
    // part of a thread's k never changes and the high part advances by BK / div per tile
    const int64_t a_kstep = p.A.col.div ? (int64_t)(BK / (int)p.A.col.div) * p.A.col.s_hi : (int64_t)BK * p.A.col.s_lo;
    const int64_t b_kstep = p.B.row.div ? (int64_t)(BK / (int)p.B.row.div) * p.B.row.s_hi : (int64_t)BK * p.B.row.s_lo;
    if (p.fast) {
        const int kbase = kt_begin * BK;
#pragma unroll
        for (int i = 0; i < A_CH; i++) {
            int c = tid + i * NT;
            int r = c / A_CPR, cp = c - r * A_CPR;
            a_so[i] = (uint32_t)((r * L::A_LD + 2 * cp) * 8);
            int klocal = A_MC ? r : 2 * cp;
            int m = A_MC ? m0 + 2 * cp : m0 + r;
            int okc = A_MC ? (m + 1 < p.M ? 2 : (m < p.M ? 1 : 0)) : (m < p.M ? 1 : 0);
            int64_t fix = okc ? p.A.row.off(m) : 0;
            a_gp[i] = Ab + fix + p.A.col.off(kbase + klocal);
            a_nb[i] = (c < A_ROWS * A_CPR) ? (A_MC ? okc * 8 : (okc ? 16 : 0)) : -1;
        }
#pragma unroll
        for (int i = 0; i < B_CH; i++) {
            int c = tid + i * NT;
            int r = c / B_CPR, cp = c - r * B_CPR;
            b_so[i] = (uint32_t)((r * L::B_LD + 2 * cp) * 8);
            int klocal = B_KC ? 2 * cp : r;
            int n = B_KC ? n0 + r : n0 + 2 * cp;
            int okc = B_KC ? (n < p.N ? 1 : 0) : (n + 1 < p.N ? 2 : (n < p.N ? 1 : 0));
            int64_t fix = okc ? p.B.col.off(n) : 0;
            b_gp[i] = Bb + fix + p.B.row.off(kbase + klocal);
            b_nb[i] = (c < B_ROWS * B_CPR) ? (B_KC ? (okc ? 16 : 0) : okc * 8) : -1;
        }
    }
    const uint32_t smem_base = smem_u32(smem);
    auto load_tile_fast = [&](int stage, int t_rel) {
        const uint32_t sA = smem_base + (uint32_t)(stage * L::STAGE_ELEMS * 8);
        const uint32_t sB = sA + (uint32_t)(L::A_ELEMS * 8);
        const int64_t ao = (int64_t)t_rel * a_kstep, bo = (int64_t)t_rel * b_kstep;
#pragma unroll
        for (int i = 0; i < A_CH; i++) {
            if (a_nb[i] < 0) continue;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sA + a_so[i]), "l"(a_gp[i] + ao), "r"(a_nb[i]));
        }
#pragma unroll
        for (int i = 0; i < B_CH; i++) {
            if (b_nb[i] < 0) continue;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sB + b_so[i]), "l"(b_gp[i] + bo), "r"(b_nb[i]));
        }
    };
    // a tile is "full" when it lies entirely inside [0, K): no k predication needed
    const int full_tiles_end = p.K / BK;   // tiles with index < full_tiles_end are full
    auto load_any = [&](int stage, int t_rel) {
        if (p.fast && (kt_begin + t_rel) < full_tiles_end) load_tile_fast(stage, t_rel);
        else load_tile(stage, kt_begin + t_rel);
    };

    double acc[MT][NTL][2];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NTL; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

    // ---- pipeline prologue -----------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nkt) load_any(s, s);
        cp_async_commit();
    }

    const int lr = lane >> 2, lc = lane & 3;
    constexpr int KSTEPS = BK / 4;
    // per-fragment smem element offsets of this thread (stage- and k-step-independent part)
    int a_off[MT], b_off[NTL];
#pragma unroll
    for (int i = 0; i < MT; i++) {
        int m = wm0 + i * 8 + lr;
        a_off[i] = A_MC ? (lc * L::A_LD + m) : (m * L::A_LD + lc);
    }
#pragma unroll
    for (int j = 0; j < NTL; j++) {
        int n = wn0 + j * 8 + lr;
        b_off[j] = B_KC ? (n * L::B_LD + lc) : (lc * L::B_LD + n);
    }
    constexpr int A_KSTRIDE = A_MC ? 4 * L::A_LD : 4;   // smem element stride of one k-step
    constexpr int B_KSTRIDE = B_KC ? 4 : 4 * L::B_LD;

    for (int it = 0; it < nkt; it++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = it + STAGES - 1;
        const bool do_load = nxt < nkt;
        const bool fast_tile = do_load && p.fast && (kt_begin + nxt) < full_tiles_end;
        if (do_load && !fast_tile) load_tile(nxt % STAGES, kt_begin + nxt);
        // fast path: the 16-byte copies of the next tile are issued one by one between the DMMAs below, so the
        // address arithmetic overlaps the tensor pipe instead of forming a loader phase after the barrier
        const uint32_t sA_n = smem_base + (uint32_t)((nxt % STAGES) * L::STAGE_ELEMS * 8);
        const uint32_t sB_n = sA_n + (uint32_t)(L::A_ELEMS * 8);
        const int64_t ao = (int64_t)nxt * a_kstep, bo = (int64_t)nxt * b_kstep;

        // fragments are fetched with volatile ld.shared one k-step ahead of the DMMAs that consume them (explicit
        // register double buffering: ptxas would otherwise sink the loads next to their use and expose LDS latency)
        const uint32_t sA_c = smem_base + (uint32_t)((it % STAGES) * L::STAGE_ELEMS * 8);
        const uint32_t sB_c = sA_c + (uint32_t)(L::A_ELEMS * 8);
        double af[2][MT], bf[2][NTL];
#pragma unroll
        for (int i = 0; i < MT; i++) af[0][i] = lds_f64(sA_c + (uint32_t)(a_off[i] * 8));
#pragma unroll
        for (int j = 0; j < NTL; j++) bf[0][j] = lds_f64(sB_c + (uint32_t)(b_off[j] * 8));
#pragma unroll
        for (int kk = 0; kk < KSTEPS; kk++) {
            const int cur = kk & 1, nx = cur ^ 1;
            if (kk + 1 < KSTEPS) {
#pragma unroll
                for (int i = 0; i < MT; i++) af[nx][i] = lds_f64(sA_c + (uint32_t)((a_off[i] + (kk + 1) * A_KSTRIDE) * 8));
#pragma unroll
                for (int j = 0; j < NTL; j++) bf[nx][j] = lds_f64(sB_c + (uint32_t)((b_off[j] + (kk + 1) * B_KSTRIDE) * 8));
            }
#pragma unroll
            for (int j = 0; j < NTL; j++) {
#pragma unroll
                for (int i = 0; i < MT; i++) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
                // one copy chunk per n-tile slot, starting with the second k-step
                if (kk >= 1) {
                    const int slot = (kk - 1) * NTL + j;
                    if (slot < A_CH) {
                        if (fast_tile && a_nb[slot] >= 0)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sA_n + a_so[slot]), "l"(a_gp[slot] + ao), "r"(a_nb[slot]));
                    } else if (slot < A_CH + B_CH) {
                        if (fast_tile && b_nb[slot - A_CH] >= 0)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sB_n + b_so[slot - A_CH]), "l"(b_gp[slot - A_CH] + bo), "r"(b_nb[slot - A_CH]));
                    }
                }
            }
        }
        static_assert(A_CH + B_CH <= (KSTEPS - 1) * NTL, "not enough DMMA slots to interleave the tile copies");
        cp_async_commit();
    }
    cp_async_wait<0>();

    // ---- epilogue ---------------------------------------------------------------------------------------------
    if (p.splitk > 1) {
        double* W = p.ws + (size_t)bz * (size_t)p.M * (size_t)p.N;
#pragma unroll
        for (int i = 0; i < MT; i++) {
            int m = m0 + wm0 + i * 8 + lr;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < NTL; j++) {
                int n = n0 + wn0 + j * 8 + 2 * lc;
                if (n < p.N) W[(size_t)m * p.N + n] = acc[i][j][0];
                if (n + 1 < p.N) W[(size_t)m * p.N + n + 1] = acc[i][j][1];
            }
        }
        return;
    }
    store_accumulators<MT, NTL>(p, acc, b, m0 + wm0, n0 + wn0, lr, lc);
}

// split-K reduction + general store:  C = alpha * sum_z ws[b][z] + beta * C
__global__ void splitk_reduce_kernel(const double* __restrict__ ws, int M, int N, int batch, int splitk, double* C,
                                     Idx2 cm, Idx2 cn, Idx2 cb, double alpha, double beta) {
    size_t total = (size_t)M * N * batch;
    size_t mn = (size_t)M * N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int b = (int)(idx / mn);
        size_t r = idx - (size_t)b * mn;
        int m = (int)(r / N), n = (int)(r - (size_t)m * N);
        const double* w = ws + (size_t)b * splitk * mn + r;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int z = 0;
        for (; z + 3 < splitk; z += 4) {   // fixed summation order => deterministic; 4 loads in flight
            s0 += w[(size_t)z * mn]; s1 += w[(size_t)(z + 1) * mn]; s2 += w[(size_t)(z + 2) * mn]; s3 += w[(size_t)(z + 3) * mn];
        }
        for (; z < splitk; z++) s0 += w[(size_t)z * mn];
        double s = (s0 + s1) + (s2 + s3);
        double* dst = C + cb.off(b) + cm.off(m) + cn.off(n);
        double v = alpha * s;
        if (beta != 0.0) v += beta * (*dst);
        *dst = v;
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------
namespace {

struct Plan {
    int bk;
    int tile;        // 1: 128x128 (16 warps), 2: 128x88, 3: 64x64, 4: 128x128 (8 warps, 32x64 warp tiles), 5: 64x88 (4 warps), 6: 128x88 BK=32, 7: 64x64 BK=32
    int bm, bn;
    int splitk;
    bool a_mc, b_kc;
    int a_vec, b_vec;
};

inline bool even64(int64_t v) { return (v & 1) == 0; }

// vectorisable along a contiguous two-level index: unit inner stride, pairs never straddle the div boundary
inline bool contig_ok(const Idx2& x) { return x.s_lo == 1 && (x.div == 0 || ((x.div & 1) == 0 && even64(x.s_hi))); }
inline bool strides_even(const Idx2& x) { return x.div == 0 ? even64(x.s_lo) : (even64(x.s_lo) && even64(x.s_hi)); }

Plan make_plan(const GemmDesc& d) {
    Plan pl;
    // orientation: prefer the index with unit stride as the smem-contiguous one
    bool a_k_unit = d.A.col.s_lo == 1, a_m_unit = d.A.row.s_lo == 1;
    pl.a_mc = (!a_k_unit && a_m_unit) || (a_k_unit && a_m_unit && d.K == 1);
    bool b_n_unit = d.B.col.s_lo == 1, b_k_unit = d.B.row.s_lo == 1;
    pl.b_kc = (!b_n_unit && b_k_unit);
    const bool a16 = (((uintptr_t)d.A.ptr) & 15) == 0, b16 = (((uintptr_t)d.B.ptr) & 15) == 0;
    if (pl.a_mc) pl.a_vec = a16 && contig_ok(d.A.row) && strides_even(d.A.col) && strides_even(d.A.batch);
    else         pl.a_vec = a16 && contig_ok(d.A.col) && strides_even(d.A.row) && strides_even(d.A.batch);
    if (pl.b_kc) pl.b_vec = b16 && contig_ok(d.B.row) && strides_even(d.B.col) && strides_even(d.B.batch);
    else         pl.b_vec = b16 && contig_ok(d.B.col) && strides_even(d.B.row) && strides_even(d.B.batch);

    const int sms = device_sm_count();
    auto waves_cost = [&](int bm, int bn, int splitk) {
        long long tiles = (long long)((d.M + bm - 1) / bm) * ((d.N + bn - 1) / bn) * d.batch * splitk;
        long long waves = (tiles + sms - 1) / sms;
        // time ~ waves * tile_area * K/splitk  (+ fixed per-CTA overhead of ~6 k-tiles)
        double kt = (double)((d.K + 15) / 16) / splitk + 6.0;
        double t = (double)waves * bm * bn * kt;
        if (splitk > 1) t += 3.0 * (double)d.M * d.N * d.batch * splitk / sms * 2.0;   // partial write + reduce traffic
        return t;
    };
    int tiles_opt[7][2] = {{128, 128}, {128, 88}, {64, 64}, {128, 128}, {64, 88}, {128, 88}, {64, 64}};
    int bk_opt[7] = {16, 16, 16, 16, 16, 32, 32};
    double best = 1e300;
    pl.tile = 1; pl.bm = 128; pl.bn = 128; pl.splitk = 1; pl.bk = 16;
    const int kt_total = (d.K + 15) / 16;
    for (int t = 0; t < 7; t++) {
        if (d.force_tile && d.force_tile != t + 1) continue;
        if (!d.force_tile && t >= 3) continue;   // tiles 4,5 only on request until measured
        int bm = tiles_opt[t][0], bn = tiles_opt[t][1];
        for (int s = 1; s <= 64; s++) {
            if (d.force_splitk && s != d.force_splitk) continue;
            if (s > 1 && kt_total / s < 8) break;
            double c = waves_cost(bm, bn, s);
            if (t == 0) c *= 1.15;   // measured: 128x128 (16 warps, 128 regs) 28 TF vs 64x64 (3 CTAs/SM) 32 TF on square shapes
            if (c < best) { best = c; pl.tile = t + 1; pl.bm = bm; pl.bn = bn; pl.splitk = s; pl.bk = bk_opt[t]; }
        }
    }
    if (d.force_splitk) pl.splitk = d.force_splitk;
    return pl;
}

template <int BM, int BN, int WM, int WN, bool A_MC, bool B_KC, int BK>
int launch_cfg(const GemmKernelParams& kp, cudaStream_t stream) {
    using L = SmemLayout<BM, BN, A_MC, B_KC, BK>;
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    auto kern = dgemm_dmma_kernel<BM, BN, WM, WN, A_MC, B_KC, BK>;
    AB_ENSURE_SMEM(kern, L::BYTES);
    GemmKernelParams k2 = kp;
    int tiles_m = (kp.M + BM - 1) / BM;
    k2.tiles_mn = tiles_m * kp.tiles_n;
    long long nblk = (long long)k2.tiles_mn * kp.batch * kp.splitk;
    if (nblk > 2147483647LL) { set_error("gemm: grid too large (%lld CTAs)", nblk); return ERR_INVALID; }
    kern<<<(unsigned)nblk, NT, L::BYTES, stream>>>(k2);
    AB_LAUNCHED();
    return OK;
}

// pointer-increment loader: both operands vectorisable and every k index either single level or two-level with div | BK
bool fast_loader_ok(const GemmDesc& d, const Plan& pl) {
    auto k_fast = [&](const Idx2& x) { return x.div == 0 || (x.div > 0 && x.div <= (uint32_t)pl.bk && pl.bk % x.div == 0); };
    return pl.a_vec && pl.b_vec && k_fast(d.A.col) && k_fast(d.B.row);
}

template <int BM, int BN, int WM, int WN, int BK = BK_DEFAULT>
int launch_orient(const GemmKernelParams& kp, bool a_mc, bool b_kc, cudaStream_t stream) {
    if (!a_mc && !b_kc) return launch_cfg<BM, BN, WM, WN, false, false, BK>(kp, stream);
    if (a_mc && !b_kc) return launch_cfg<BM, BN, WM, WN, true, false, BK>(kp, stream);
    if (!a_mc && b_kc) return launch_cfg<BM, BN, WM, WN, false, true, BK>(kp, stream);
    return launch_cfg<BM, BN, WM, WN, true, true, BK>(kp, stream);
}

}  // namespace

size_t gemm_workspace_bytes(const GemmDesc& d) {
    if (d.M <= 0 || d.N <= 0 || d.batch <= 0) return 0;
    Plan pl = make_plan(d);
    if (pl.splitk <= 1) return 0;
    return ws_round((size_t)d.M * d.N * d.batch * pl.splitk * sizeof(double));
}

int gemm_launch(const GemmDesc& d, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (d.M <= 0 || d.N <= 0 || d.batch <= 0) return OK;
    AB_REQUIRE(d.K >= 0, "gemm: negative K");
    Plan pl = make_plan(d);
    GemmKernelParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.M = d.M; kp.N = d.N; kp.K = d.K; kp.batch = d.batch; kp.splitk = pl.splitk;
    kp.tiles_n = (d.N + pl.bn - 1) / pl.bn;
    kp.kt_total = (d.K + pl.bk - 1) / pl.bk;
    kp.A = d.A; kp.B = d.B; kp.C = d.C; kp.cm = d.cm; kp.cn = d.cn; kp.cb = d.cb;
    kp.alpha = d.alpha; kp.beta = d.beta;
    kp.a_vec = pl.a_vec; kp.b_vec = pl.b_vec;
    kp.fast = fast_loader_ok(d, pl) ? 1 : 0;
    kp.c_vec = ((((uintptr_t)d.C) & 15) == 0 && contig_ok(d.cn) && strides_even(d.cm) && strides_even(d.cb)) ? 1 : 0;
    kp.ws = nullptr;
    if (pl.splitk > 1) {
        size_t need = (size_t)d.M * d.N * d.batch * pl.splitk * sizeof(double);
        if (ws == nullptr || ws_bytes < need) {
            set_error("gemm: split-K workspace too small (%zu needed, %zu given)", need, ws_bytes);
            return ERR_WORKSPACE;
        }
        kp.ws = (double*)ws;
    }
    int st;
    if (pl.tile == 1) st = launch_orient<128, 128, 32, 32>(kp, pl.a_mc, pl.b_kc, stream);
    else if (pl.tile == 2) st = launch_orient<128, 88, 16, 88>(kp, pl.a_mc, pl.b_kc, stream);
    else if (pl.tile == 4) st = launch_orient<128, 128, 32, 64>(kp, pl.a_mc, pl.b_kc, stream);
    else if (pl.tile == 5) st = launch_orient<64, 88, 16, 88>(kp, pl.a_mc, pl.b_kc, stream);
    else if (pl.tile == 6) st = launch_orient<128, 88, 16, 88, 32>(kp, pl.a_mc, pl.b_kc, stream);
    else if (pl.tile == 7) st = launch_orient<64, 64, 32, 32, 32>(kp, pl.a_mc, pl.b_kc, stream);
    else st = launch_orient<64, 64, 32, 32>(kp, pl.a_mc, pl.b_kc, stream);
    if (st) return st;
    if (pl.splitk > 1) {
        size_t total = (size_t)d.M * d.N * d.batch;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(kp.ws, d.M, d.N, d.batch, pl.splitk, d.C, d.cm, d.cn, d.cb,
                                                        d.alpha, d.beta);
        AB_LAUNCHED();
    }
    return OK;
}

}  // namespace ab200
