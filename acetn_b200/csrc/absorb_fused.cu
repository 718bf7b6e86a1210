// K2 (fused form): double-layer site absorption with the intermediate kept on chip (Q.s3+Q.s4,
// E.2+E.3 -- reference acetn/renormalization/projectors.py:54-55 and directional_mover.py:363-364).
//
// Per (chi,chi) block:  Y[(r,R),(d,D)] = sum_{i0,I0,i1,I1,p} X[i0,I0,i1,I1] * conj(A)[I.,I.,R,D,p] * A[i.,i.,r,d,p]
// done as two DMMA GEMMs whose intermediate W never leaves the register file:
//   step 1  W^T[(p,y),(i0,i1)] = sum_{(I0,I1)} S[(I0,I1)][p][y] * X[(i0,I0),(i1,I1)]          y = (R,D)
//   step 2  Y[y, n]            = sum_{(i0,i1),p} W^T[(p,y),(i0,i1)] * S[(i0,i1)][p][n]         n = (r,d)
// The accumulator fragment of step 1 (row lane/4, columns 2*(lane%4)+{0,1}) is turned into the A fragment of step 2
// with two warp shuffles per k-step; S (the site tensor, 64 KiB at D=8,d=2, one copy serves bra and ket because the
// path is real FP64) stays in shared memory for the lifetime of a persistent CTA, X blocks stream through a
// double-buffered cp.async stage.  Algorithmic HBM traffic per block: D^4*8 B in + D^4*8 B out (32 flop/B at D=8).
//
// Specialised for D = 8, d = 2 (the headline configuration); other shapes use the unfused two-GEMM form in api.cu.
#include "kernels.cuh"

namespace ab200 {

namespace {
constexpr int FD = 8, FD2 = 64, Fd = 2;
constexpr int S_PX = Fd * FD2 + 4;        // 132: pitch of the pair index x  (== 4 mod 16 -> conflict-free fragments)
constexpr int S_PP = FD2;                 // pitch of the physical index
constexpr int X_ROW = 96;                 // X stage row pitch: 8 groups (i1) of 8 (+4 pad) doubles
constexpr int X_I1 = 12;
constexpr int S_ELEMS = FD2 * S_PX;       // 8448
constexpr int X_ELEMS = FD2 * X_ROW;      // 6144
constexpr size_t FUSED_SMEM = (size_t)(S_ELEMS + 2 * X_ELEMS) * sizeof(double);
constexpr int F_THREADS = 256;

struct FusedParams {
    const double* X; int64_t n0, n1, in_s0, in_s1, es0, es1, es2;
    const double* A; int64_t ax0, ax1, ap, ay0, ay1;       // strides of A for x=(f0,f1), p, y=(r,d)
    double* Y; int64_t out_s0, out_s1, oes0, oes1, oes2, oes3;
    double* absmax;
    int32_t* colexp;       // optional (D = 8 kernel): colexp[b1 * D^2 + n] = max over (b0, y) of the binary exponent of Y, for the K7 encoding
};

__global__ void __launch_bounds__(F_THREADS, 1) double_layer_fused_d8_kernel(const FusedParams p) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* Xs = sm + S_ELEMS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const int64_t nblk = p.n0 * p.n1;

    // ---- site tensor -> S[x][p][y]
    for (int idx = tid; idx < FD2 * Fd * FD2; idx += F_THREADS) {
        int y = idx & 63, pp = (idx >> 6) & 1, x = idx >> 7;
        S[x * S_PX + pp * S_PP + y] =
            p.A[(x >> 3) * p.ax0 + (x & 7) * p.ax1 + pp * p.ap + (y >> 3) * p.ay0 + (y & 7) * p.ay1];
    }

    // ---- X stage loader: thread -> (I0 = w, i1 = (lane>>2), pair = lane&3), 8 chunks over i0
    const int64_t x_thread = (int64_t)w * p.es1 + (int64_t)(lane >> 2) * p.es2 + 2 * (lane & 3);
    const uint32_t xs_base = smem_u32(Xs);
    const uint32_t xs_thread = (uint32_t)((w * X_ROW + (lane >> 2) * X_I1 + 2 * (lane & 3)) * 8);
    auto load_block = [&](int64_t blk, int buf) {
        const int64_t b0 = blk / p.n1, b1 = blk - b0 * p.n1;
        const double* src = p.X + b0 * p.in_s0 + b1 * p.in_s1 + x_thread;
        const uint32_t dst = xs_base + (uint32_t)(buf * X_ELEMS * 8) + xs_thread;
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(i * 8 * X_ROW * 8)), "l"(src + i * p.es0));
    };

    // column exponents of the output for the K7 residue encoding (i8crt.cu): the 8 warps of the CTA hold the 64 rows y of a block,
    // so the per-block column maximum is an 8-way shared-memory max, flushed with one global atomic max per column and block
    __shared__ int cmx[FD2];
    if (tid < FD2) cmx[tid] = 0;

    double vmax = 0.0;
    int64_t blk = blockIdx.x;
    if (blk < nblk) load_block(blk, 0);
    cp_async_commit();
    int buf = 0;
    for (; blk < nblk; blk += gridDim.x, buf ^= 1) {
        const int64_t nxt = blk + gridDim.x;
        if (nxt < nblk) load_block(nxt, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double* Xb = Xs + buf * X_ELEMS;

        // ---- step 1: acc1[P][j] = W^T[(P, y = 8w+lr), x = 8j + 2lc + {0,1}]
        double acc1[2][8][2];
#pragma unroll
        for (int P = 0; P < 2; P++)
#pragma unroll
            for (int j = 0; j < 8; j++) { acc1[P][j][0] = 0.0; acc1[P][j][1] = 0.0; }
        const double* Sa = S + lc * S_PX + 8 * w + lr;           // + 4*ks*S_PX + P*S_PP
        const double* Xf = Xb + lr * X_I1 + lc;                  // + (8j + (ks>>1))*X_ROW + (ks&1)*4
#pragma unroll
        for (int ks = 0; ks < 16; ks++) {
            double a0 = Sa[4 * ks * S_PX], a1 = Sa[4 * ks * S_PX + S_PP];
            double bfr[8];
#pragma unroll
            for (int j = 0; j < 8; j++) bfr[j] = Xf[(8 * j + (ks >> 1)) * X_ROW + (ks & 1) * 4];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                dmma884(acc1[0][j][0], acc1[0][j][1], a0, bfr[j]);
                dmma884(acc1[1][j][0], acc1[1][j][1], a1, bfr[j]);
            }
        }

        // ---- step 2: acc2[jn] = Y[y = 8w+lr, n = 8jn + 2lc + {0,1}]
        double acc2[8][2];
#pragma unroll
        for (int j = 0; j < 8; j++) { acc2[j][0] = 0.0; acc2[j][1] = 0.0; }
        const double* Sb = S + lc * S_PX + lr;                   // + x0*S_PX + P*S_PP + 8*jn
#pragma unroll
        for (int P = 0; P < 2; P++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    // A fragment of step 2: W^T[(P,y), x0 + lc], x0 = 8j + 4h, lives in lane (lr, 2h + lc/2), register lc&1
                    const int srcl = (lane & ~3) | (2 * h + (lc >> 1));
                    double v0 = __shfl_sync(0xffffffffu, acc1[P][j][0], srcl);
                    double v1 = __shfl_sync(0xffffffffu, acc1[P][j][1], srcl);
                    double a = (lc & 1) ? v1 : v0;
                    const double* sb = Sb + (8 * j + 4 * h) * S_PX + P * S_PP;
                    double bfr[8];
#pragma unroll
                    for (int jn = 0; jn < 8; jn++) bfr[jn] = sb[8 * jn];
#pragma unroll
                    for (int jn = 0; jn < 8; jn++) dmma884(acc2[jn][0], acc2[jn][1], a, bfr[jn]);
                }
            }
        }

        // ---- store: y = (R = w, Dd = lr), n = (r = jn, dd = 2lc + e)
        {
            const int64_t b0 = blk / p.n1, b1 = blk - b0 * p.n1;
            double* Yb = p.Y + b0 * p.out_s0 + b1 * p.out_s1 + (int64_t)w * p.oes1 + (int64_t)lr * p.oes3;
            int ef[2] = {0, 0};
#pragma unroll
            for (int jn = 0; jn < 8; jn++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    double v = acc2[jn][e];
                    Yb[(int64_t)jn * p.oes0 + (int64_t)(2 * lc + e) * p.oes2] = v;
                    vmax = fmax(vmax, fabs(v));
                    ef[e] = max(ef[e], (__double2hiint(v) >> 20) & 0x7ff);
                }
            }
            if (p.colexp != nullptr) {
                // this thread's columns: n = (dd = 2 lc + e, Dd = lr)
                atomicMax(&cmx[(2 * lc) * FD + lr], ef[0]);
                atomicMax(&cmx[(2 * lc + 1) * FD + lr], ef[1]);
                __syncthreads();
                if (tid < FD2) {
                    const int m = cmx[tid];
                    if (m > 0) atomicMax(p.colexp + b1 * FD2 + tid, m - 1022);
                    cmx[tid] = 0;
                }
            }
        }
        __syncthreads();     // all warps are done with Xs[buf] (and cmx) before the next iteration refills it
    }
    cp_async_wait<0>();
    if (p.absmax) {
        vmax = warp_max(vmax);
        if (lane == 0) atomic_max_nonneg(p.absmax, vmax);
    }
}
}  // namespace


// ------------------------------------------------------------------------------------------------------------------
// General (D, d) version of the same algorithm.  D^2 is padded to multiples of 8 (rows y / columns n / ket pairs x_k)
// and 4 (bra pairs x_b); the padding lives only in shared memory (zero-filled once).  NB blocks are processed
// concurrently by one CTA so that small D still fills the SM: warp w -> (block w / NY, y-tile w % NY).
// X is staged pair-major, Xk[x_k = (i0,i1)][x_b = (I0,I1)] with a pitch == 4 (mod 16) doubles, which makes the B
// fragments of step 1 conflict free for every D.
// ------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int ceil_to(int v, int m) { return (v + m - 1) / m * m; }
constexpr int pitch4mod16(int v) { return v + ((20 - v % 16) % 16); }

template <int D, int d>
struct FusedCfg {
    static constexpr int D2 = D * D;
    static constexpr int Y8 = ceil_to(D2, 8);            // padded y, n, x_k extent
    static constexpr int X4 = ceil_to(D2, 4);            // padded x_b extent (K of step 1)
    static constexpr int NY = Y8 / 8;                    // y-tiles = warps per block; also x_k tiles and n tiles
    static constexpr int NB = (NY >= 6) ? 1 : ((NY >= 3) ? 2 : ((NY == 2) ? 4 : 8));
    static constexpr int WARPS = NB * NY;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int PP = Y8;
    static constexpr int PX = pitch4mod16(d * Y8);
    static constexpr int XP = pitch4mod16(X4);
    static constexpr int S_ELEMS = Y8 * PX;
    static constexpr int XB_ELEMS = Y8 * XP;             // one staged block
    static constexpr size_t SMEM = (size_t)(S_ELEMS + 2 * NB * XB_ELEMS) * sizeof(double);
};

template <int D, int d>
__global__ void __launch_bounds__(FusedCfg<D, d>::THREADS, 1) double_layer_fused_generic_kernel(const FusedParams p) {
    using C = FusedCfg<D, d>;
    constexpr int D2 = C::D2, Y8 = C::Y8, X4 = C::X4, NY = C::NY, NB = C::NB, PX = C::PX, PP = C::PP, XP = C::XP;
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* Xs = sm + C::S_ELEMS;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int lr = lane >> 2, lc = lane & 3;
    const int wb = w / NY, yt = w - wb * NY;             // block slot and y-tile of this warp
    const int64_t nblk = p.n0 * p.n1;
    const int64_t ngroups = (nblk + NB - 1) / NB;

    // ---- zero everything once (padding must read as 0), then the site tensor -> S[x][p][y]
    for (int i = tid; i < C::S_ELEMS + 2 * NB * C::XB_ELEMS; i += C::THREADS) sm[i] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < D2 * d * D2; idx += C::THREADS) {
        int y = idx % D2, pp = (idx / D2) % d, x = idx / (D2 * d);
        S[x * PX + pp * PP + y] = p.A[(x / D) * p.ax0 + (x % D) * p.ax1 + pp * p.ap + (y / D) * p.ay0 + (y % D) * p.ay1];
    }

    // ---- X stage loader: element (i0,I0,i1,I1) -> Xk[(i0*D+i1)*XP + I0*D + I1]; 8-byte async copies (any D, any alignment)
    const uint32_t xs_base = smem_u32(Xs);
    auto load_group = [&](int64_t grp, int buf) {
        for (int b = 0; b < NB; b++) {
            const int64_t blk = grp * NB + b;
            if (blk >= nblk) break;
            const int64_t b0 = blk / p.n1, b1 = blk - b0 * p.n1;
            const double* src = p.X + b0 * p.in_s0 + b1 * p.in_s1;
            const uint32_t dst = xs_base + (uint32_t)(((buf * NB + b) * C::XB_ELEMS) * 8);
            for (int e = tid; e < D2 * D2; e += C::THREADS) {
                int I1 = e % D, i1 = (e / D) % D, I0 = (e / D2) % D, i0 = e / (D2 * D);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (uint32_t)(((i0 * D + i1) * XP + I0 * D + I1) * 8)),
                             "l"(src + i0 * p.es0 + I0 * p.es1 + i1 * p.es2 + I1));
            }
        }
    };

    double vmax = 0.0;
    int64_t grp = blockIdx.x;
    if (grp < ngroups) load_group(grp, 0);
    cp_async_commit();
    int buf = 0;
    for (; grp < ngroups; grp += gridDim.x, buf ^= 1) {
        const int64_t nxt = grp + gridDim.x;
        if (nxt < ngroups) load_group(nxt, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int64_t blk = grp * NB + wb;
        if (blk < nblk) {
            const double* Xb = Xs + (size_t)(buf * NB + wb) * C::XB_ELEMS;
            // ---- step 1: acc1[P][j] = W^T[(P, y = 8yt+lr), x_k = 8j + 2lc + {0,1}]
            double acc1[d][NY][2];
#pragma unroll
            for (int P = 0; P < d; P++)
#pragma unroll
                for (int j = 0; j < NY; j++) { acc1[P][j][0] = 0.0; acc1[P][j][1] = 0.0; }
            const double* Sa = S + lc * PX + 8 * yt + lr;
            const double* Xf = Xb + lr * XP + lc;
#pragma unroll
            for (int ks = 0; ks < X4 / 4; ks++) {
                double af[d], bfr[NY];
#pragma unroll
                for (int P = 0; P < d; P++) af[P] = Sa[4 * ks * PX + P * PP];
#pragma unroll
                for (int j = 0; j < NY; j++) bfr[j] = Xf[8 * j * XP + 4 * ks];
#pragma unroll
                for (int j = 0; j < NY; j++)
#pragma unroll
                    for (int P = 0; P < d; P++) dmma884(acc1[P][j][0], acc1[P][j][1], af[P], bfr[j]);
            }
            // ---- step 2: acc2[jn] = Y[y = 8yt+lr, n = 8jn + 2lc + {0,1}]
            double acc2[NY][2];
#pragma unroll
            for (int j = 0; j < NY; j++) { acc2[j][0] = 0.0; acc2[j][1] = 0.0; }
            const double* Sb = S + lc * PX + lr;
#pragma unroll
            for (int P = 0; P < d; P++) {
#pragma unroll
                for (int j = 0; j < NY; j++) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int srcl = (lane & ~3) | (2 * h + (lc >> 1));
                        double v0 = __shfl_sync(0xffffffffu, acc1[P][j][0], srcl);
                        double v1 = __shfl_sync(0xffffffffu, acc1[P][j][1], srcl);
                        double a = (lc & 1) ? v1 : v0;
                        const double* sb = Sb + (8 * j + 4 * h) * PX + P * PP;
                        double bfr[NY];
#pragma unroll
                        for (int jn = 0; jn < NY; jn++) bfr[jn] = sb[8 * jn];
#pragma unroll
                        for (int jn = 0; jn < NY; jn++) dmma884(acc2[jn][0], acc2[jn][1], a, bfr[jn]);
                    }
                }
            }
            // ---- store: y = (R, Dd), n = (r, dd)
            const int64_t b0 = blk / p.n1, b1 = blk - b0 * p.n1;
            const int y = 8 * yt + lr;
            if (y < D2) {
                double* Yb = p.Y + b0 * p.out_s0 + b1 * p.out_s1 + (int64_t)(y / D) * p.oes1 + (int64_t)(y % D) * p.oes3;
#pragma unroll
                for (int jn = 0; jn < NY; jn++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int n = 8 * jn + 2 * lc + e;
                        if (n < D2) {
                            double v = acc2[jn][e];
                            Yb[(int64_t)(n / D) * p.oes0 + (int64_t)(n % D) * p.oes2] = v;
                            vmax = fmax(vmax, fabs(v));
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    if (p.absmax) {
        vmax = warp_max(vmax);
        if (lane == 0) atomic_max_nonneg(p.absmax, vmax);
    }
}

template <int D, int d>
int launch_generic(const FusedParams& p, cudaStream_t s) {
    using C = FusedCfg<D, d>;
    static_assert(C::SMEM <= 227 * 1024, "fused double-layer: shared memory budget exceeded");
    AB_ENSURE_SMEM((double_layer_fused_generic_kernel<D, d>), C::SMEM);
    const int64_t ngroups = (p.n0 * p.n1 + C::NB - 1) / C::NB;
    int grid = device_sm_count();
    if (ngroups < grid) grid = (int)ngroups;
    if (grid < 1) return OK;
    double_layer_fused_generic_kernel<D, d><<<grid, C::THREADS, C::SMEM, s>>>(p);
    AB_LAUNCHED();
    return OK;
}

// (D, d) pairs with a compiled generic kernel
#define AB_FUSED_LIST(X) X(2, 2) X(3, 2) X(3, 3) X(4, 2) X(5, 2) X(6, 2) X(7, 2) X(7, 4)
}  // namespace

// 1 when the fused kernel for (D, d) can deliver the column exponents of its output (see FusedParams::colexp)
int double_layer_fused_colexp_supported(int64_t D, int64_t d) { return (D == 8 && d == 2) ? 1 : 0; }

int double_layer_fused_supported(int64_t D, int64_t d) {
    if (D == 8 && d == 2) return 1;
#define AB_CASE(DD, dd) if (D == DD && d == dd) return 1;
    AB_FUSED_LIST(AB_CASE)
#undef AB_CASE
    return 0;
}

int double_layer_fused_launch(const double* X, int64_t n0, int64_t n1, int64_t in_s0, int64_t in_s1, const int64_t* in_es,
                              int order, const double* A, const int64_t* a_strides, int64_t D, int64_t d, double* Y,
                              int64_t out_s0, int64_t out_s1, const int64_t* out_es, double* absmax, int32_t* colexp, cudaStream_t s) {
    if (!double_layer_fused_supported(D, d)) {
        set_error("double_layer_fused: no specialisation for D=%lld d=%lld", (long long)D, (long long)d);
        return ERR_UNSUPPORTED;
    }
    AB_REQUIRE(in_es[3] == 1, "double_layer_fused: innermost input leg must have unit stride");
    FusedParams p;
    p.X = X; p.n0 = n0; p.n1 = n1; p.in_s0 = in_s0; p.in_s1 = in_s1; p.es0 = in_es[0]; p.es1 = in_es[1]; p.es2 = in_es[2];
    // legs of the A view: 0=l 1=u 2=r 3=d 4=p ; order 0: (i0,i1) = (u,l) ; order 1: (i0,i1) = (l,u)
    const int f0 = order == 0 ? 1 : 0, f1 = order == 0 ? 0 : 1;
    p.A = A; p.ax0 = a_strides[f0]; p.ax1 = a_strides[f1]; p.ap = a_strides[4]; p.ay0 = a_strides[2]; p.ay1 = a_strides[3];
    p.Y = Y; p.out_s0 = out_s0; p.out_s1 = out_s1; p.oes0 = out_es[0]; p.oes1 = out_es[1]; p.oes2 = out_es[2]; p.oes3 = out_es[3];
    p.absmax = absmax;
    p.colexp = colexp;
    if (!(D == 8 && d == 2)) {
        AB_REQUIRE(colexp == nullptr, "double_layer_fused: column exponents are produced by the D = 8, d = 2 kernel only");
#define AB_CASE(DD, dd) if (D == DD && d == dd) return launch_generic<DD, dd>(p, s);
        AB_FUSED_LIST(AB_CASE)
#undef AB_CASE
    }
    AB_REQUIRE((((uintptr_t)X) & 15) == 0 && (in_s0 % 2) == 0 && (in_s1 % 2) == 0 && (in_es[0] % 2) == 0 && (in_es[1] % 2) == 0 &&
                   (in_es[2] % 2) == 0,
               "double_layer_fused: input block addressing must be 16-byte aligned");
    AB_ENSURE_SMEM(double_layer_fused_d8_kernel, FUSED_SMEM);
    int64_t nblk = n0 * n1;
    // one CTA per SM is resident at a time; `waves` CTAs per SM in the grid (each stages the site tensor once for >= 64 blocks)
    // bound the time a higher-priority stream waits for an SM (the rSVD chains of other site tasks, renormalization.py)
    static const int waves_env = getenv("ACETN_B200_K2_WAVES") ? atoi(getenv("ACETN_B200_K2_WAVES")) : 0;
    int waves = waves_env > 0 ? waves_env : 4;
    while (waves > 1 && nblk < (int64_t)device_sm_count() * waves * 64) waves--;
    int64_t grid = (int64_t)device_sm_count() * waves;
    if (nblk < grid) grid = nblk;
    if (grid < 1) return OK;
    double_layer_fused_d8_kernel<<<(unsigned)grid, F_THREADS, FUSED_SMEM, s>>>(p);
    AB_LAUNCHED();
    return OK;
}

}  // namespace ab200
