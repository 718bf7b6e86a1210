// K5: one-sided (Hestenes) Jacobi SVD of the small square core of the randomized SVD
// (reference: torch.linalg.svd of the projected matrix, acetn/linalg/fused_matmul_svd_lowrank.py:48, after the
// tall factor has been reduced to a (chi+p) x (chi+p) core by K4 + one DGEMM).
//
// Rows of X are rotated pairwise until mutually orthogonal:  J X = diag(S) W.  Rotations use the relative criterion
// |x_p.x_q| <= tol ||x_p|| ||x_q||, which gives high relative accuracy of small singular values.  Two kernels:
//   jacobi_small_kernel (q <= 112): the whole SVD on one CTA, X and J in its shared memory, one __syncthreads per step;
//   jacobi_block_kernel (larger cores): row blocks of 8, every CTA of a cooperative grid stages a block pair from L2 per phase.
#include <stdlib.h>

#include "kernels.cuh"

namespace ab200 {

constexpr int JB_THREADS = 256;   // 8 warps: one warp per row pair of a local step (br <= 8)
constexpr int JB_NIT = 5;         // double2 per lane and row held in registers by the cached rotation (rows up to 320 columns)

struct JacobiParams {
    double* X;      // n x ld, row major (n = nb * br rows, zero padded)
    double* J;      // n x ld
    int n, ld, ncols;
    int br;         // rows per block (16, or 8 for wide cores)
    int nb;         // number of row blocks (even)
    int max_sweeps;
    double tol;
    double* conv;   // [max_sweeps] max relative off-diagonal seen in each sweep
    int* info;      // [0] sweeps used
};

// Rotation that orthogonalises two rows from their Gram entries, without sqrt or division: two dependent rsqrt instead of
// sqrt -> division -> rsqrt (this scalar chain is the critical path of every Jacobi step).
//   cos^2 = (1 + |d| / h) / 2,  sin = sgn(d) s_ab / (h cos),  d = s_bb - s_aa,  h = sqrt(d^2 + 4 s_ab^2)
// (the same rotation as t = sgn(d) 2 s_ab / (|d| + h), cos = 1 / sqrt(1 + t^2), sin = cos t.)
__device__ __forceinline__ void rotation_from_gram(double saa, double sbb, double sab, double& cs, double& sn) {
    const double d = sbb - saa;
    const double rh = rsqrt(fma(d, d, 4.0 * sab * sab));
    const double c2 = fma(0.5 * fabs(d), rh, 0.5);
    const double rc = rsqrt(c2);
    cs = c2 * rc;
    sn = (d >= 0.0 ? sab : -sab) * rh * rc;
}

// One row pair, owned by a group of G lanes (G = 32: a warp; smaller groups share a warp and own different pairs).  The X rows are
// read ONCE into registers (NIT double2 per lane and row), the J rows are requested before the reduction + scalar chain and consumed
// after it (CACHE_J), so a step is one shared-memory round trip, the dot products, log2(G) shuffle rounds, two rsqrt and the
// rotation FMAs -- instead of two passes over shared memory around a sqrt / division / rsqrt chain.  Needs ncols2 <= G * NIT.
template <int G, int NIT, bool CACHE_J>
__device__ __forceinline__ int rotate_pair_cached(double* xa, double* xb, double* ja, double* jb, int ncols2, double tol, int gl) {
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (threadIdx.x & ((32 - G) & 31)));
    double2* xa2 = reinterpret_cast<double2*>(xa);
    double2* xb2 = reinterpret_cast<double2*>(xb);
    double2* ja2 = reinterpret_cast<double2*>(ja);
    double2* jb2 = reinterpret_cast<double2*>(jb);
    double2 u[NIT], v[NIT], ju[CACHE_J ? NIT : 1], jv[CACHE_J ? NIT : 1];
#pragma unroll
    for (int i = 0; i < NIT; i++) {
        const int c = gl + i * G;
        const bool ok = c < ncols2;
        u[i] = ok ? xa2[c] : make_double2(0.0, 0.0);
        v[i] = ok ? xb2[c] : make_double2(0.0, 0.0);
    }
    if (CACHE_J) {
#pragma unroll
        for (int i = 0; i < NIT; i++) {
            const int c = gl + i * G;
            const bool ok = c < ncols2;
            ju[i] = ok ? ja2[c] : make_double2(0.0, 0.0);
            jv[i] = ok ? jb2[c] : make_double2(0.0, 0.0);
        }
    }
    double saa = 0.0, sbb = 0.0, sab = 0.0;
#pragma unroll
    for (int i = 0; i < NIT; i++) {
        saa = fma(u[i].x, u[i].x, fma(u[i].y, u[i].y, saa));
        sbb = fma(v[i].x, v[i].x, fma(v[i].y, v[i].y, sbb));
        sab = fma(u[i].x, v[i].x, fma(u[i].y, v[i].y, sab));
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
        saa += __shfl_xor_sync(gmask, saa, o);
        sbb += __shfl_xor_sync(gmask, sbb, o);
        sab += __shfl_xor_sync(gmask, sab, o);
    }
    if (saa == 0.0 || sbb == 0.0) return 0;
    // relative criterion |xa.xb| <= tol |xa||xb| tested without sqrt/div
    if (sab * sab <= tol * tol * (saa * sbb)) return 0;
    double cs, sn;
    rotation_from_gram(saa, sbb, sab, cs, sn);
#pragma unroll
    for (int i = 0; i < NIT; i++) {
        const int c = gl + i * G;
        if (c < ncols2) {
            double2 r;
            r.x = cs * u[i].x - sn * v[i].x; r.y = cs * u[i].y - sn * v[i].y; xa2[c] = r;
            r.x = sn * u[i].x + cs * v[i].x; r.y = sn * u[i].y + cs * v[i].y; xb2[c] = r;
            const double2 a = CACHE_J ? ju[i] : ja2[c], b = CACHE_J ? jv[i] : jb2[c];
            r.x = cs * a.x - sn * b.x; r.y = cs * a.y - sn * b.y; ja2[c] = r;
            r.x = sn * a.x + cs * b.x; r.y = sn * a.y + cs * b.y; jb2[c] = r;
        }
    }
    return 1;
}

// Rotate rows (xa, xb) of X (and ja, jb of J) held in shared memory so that xa . xb = 0.  One warp; rows are read
// as double2 (the row pitch is even and the pad column, if any, is zero in X and J).
__device__ __forceinline__ double rotate_pair(double* xa, double* xb, double* ja, double* jb, int ncols2, double tol, int lane) {
    double2* xa2 = reinterpret_cast<double2*>(xa);
    double2* xb2 = reinterpret_cast<double2*>(xb);
    double saa = 0.0, sbb = 0.0, sab = 0.0;
    for (int c = lane; c < ncols2; c += 32) {
        double2 u = xa2[c], v = xb2[c];
        saa += u.x * u.x + u.y * u.y; sbb += v.x * v.x + v.y * v.y; sab += u.x * v.x + u.y * v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        saa += __shfl_xor_sync(0xffffffffu, saa, o);
        sbb += __shfl_xor_sync(0xffffffffu, sbb, o);
        sab += __shfl_xor_sync(0xffffffffu, sab, o);
    }
    if (saa == 0.0 || sbb == 0.0) return 0.0;
    // relative criterion |xa.xb| <= tol |xa||xb| tested without sqrt/div; the ratio is only formed for pairs that rotate
    const double r2 = sab * sab, den = saa * sbb;
    if (r2 <= tol * tol * den) return 0.0;
    // t = sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (sbb - saa) / (2 sab), written with one sqrt and one division (this
    // scalar chain is the critical path of a local step); the return value only says "rotated" (1.0): a sweep without any
    // rotation is the convergence criterion
    const double d = sbb - saa;
    const double h = sqrt(d * d + 4.0 * r2);
    double t = (2.0 * sab) / (fabs(d) + h);
    t = d >= 0.0 ? t : -t;
    const double rel = 1.0;
    double cs = rsqrt(1.0 + t * t), sn = cs * t;
    double2* ja2 = reinterpret_cast<double2*>(ja);
    double2* jb2 = reinterpret_cast<double2*>(jb);
    for (int c = lane; c < ncols2; c += 32) {
        double2 u = xa2[c], v = xb2[c], r;
        r.x = cs * u.x - sn * v.x; r.y = cs * u.y - sn * v.y; xa2[c] = r;
        r.x = sn * u.x + cs * v.x; r.y = sn * u.y + cs * v.y; xb2[c] = r;
        double2 ju = ja2[c], jv = jb2[c];
        r.x = cs * ju.x - sn * jv.x; r.y = cs * ju.y - sn * jv.y; ja2[c] = r;
        r.x = sn * ju.x + cs * jv.x; r.y = sn * ju.y + cs * jv.y; jb2[c] = r;
    }
    return rel;
}

// Block one-sided Jacobi.  Rows are grouped in nb blocks of br rows.  A sweep = one "diagonal" phase (all pairs inside
// each block) + nb-1 round-robin phases in which every CTA owns one block pair (A,B), stages its 2*br rows of X and J
// in shared memory and orthogonalises all br*br cross pairs in br conflict-free local steps.  Phases are separated by
// a cooperative grid barrier (nb per sweep instead of n-1 for the plain cyclic ordering).
// Grid barrier of the (cooperatively launched, hence co-resident) block kernel: one arrival counter that only grows, zeroed by the host
// before the launch; thread 0 of every CTA arrives and polls.  Lighter than cooperative_groups' grid.sync() for ~17 CTAs.
__device__ __forceinline__ void jb_grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nblocks) {
    __syncthreads();
    target += nblocks;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(JB_THREADS, 1) jacobi_block_kernel(JacobiParams p) {
    unsigned int* bar = reinterpret_cast<unsigned int*>(p.info + 8);
    unsigned int bar_target = 0;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int br = p.br, nb = p.nb, ld = p.ld, ncols2 = p.ld / 2;
    const int npairs = nb / 2, nbm1 = nb - 1;
    double* Xs = sm;                       // [2*br][ld]
    double* Js = sm + (size_t)2 * br * ld;
    // rows of up to 320 columns (the rSVD cores of every BASELINE config) are rotated out of registers
    const bool cached = ncols2 <= 32 * JB_NIT;
#define JB_ROTATE(xa, xb, ja, jb) \
    (cached ? (double)rotate_pair_cached<32, JB_NIT, true>(xa, xb, ja, jb, ncols2, p.tol, lane) : rotate_pair(xa, xb, ja, jb, ncols2, p.tol, lane))

    int sweep = 0;
    for (; sweep < p.max_sweeps; sweep++) {
        double worst = 0.0;
        for (int phase = 0; phase < nb; phase++) {
            for (int pi = blockIdx.x; pi < npairs; pi += gridDim.x) {
                int A, B;
                if (phase == 0) { A = 2 * pi; B = 2 * pi + 1; }
                else {
                    int s = phase - 1;
                    if (pi == 0) { A = nbm1; B = s; }
                    else { A = (s + pi) % nbm1; B = (s - pi + nbm1) % nbm1; }
                }
                // stage rows of blocks A and B
                const int rowlen2 = ld / 2;
                for (int idx = threadIdx.x; idx < 2 * br * rowlen2; idx += JB_THREADS) {
                    int r = idx / rowlen2, c2 = idx - r * rowlen2;
                    int g = (r < br ? A * br + r : B * br + (r - br));
                    reinterpret_cast<double2*>(Xs + (size_t)r * ld)[c2] = __ldcg(reinterpret_cast<const double2*>(p.X + (size_t)g * ld) + c2);
                    reinterpret_cast<double2*>(Js + (size_t)r * ld)[c2] = __ldcg(reinterpret_cast<const double2*>(p.J + (size_t)g * ld) + c2);
                }
                __syncthreads();
                if (phase == 0) {
                    // all pairs inside A (warps [0, br/2)) and inside B (warps [br/2, br)): cyclic ordering on br players
                    const int hb = br / 2, brm1 = br - 1;
                    for (int t = 0; t < brm1; t++) {
                        if (warp < br) {
                            int blk = warp / hb, i = warp - blk * hb;
                            int a, b;
                            if (i == 0) { a = brm1; b = t; }
                            else { a = (t + i) % brm1; b = (t - i + brm1) % brm1; }
                            a += blk * br; b += blk * br;
                            double rel = JB_ROTATE(Xs + (size_t)a * ld, Xs + (size_t)b * ld, Js + (size_t)a * ld, Js + (size_t)b * ld);
                            worst = fmax(worst, rel);
                        }
                        __syncthreads();
                    }
                } else {
                    for (int t = 0; t < br; t++) {
                        if (warp < br) {
                            int a = warp, b = br + (warp + t) % br;
                            double rel = JB_ROTATE(Xs + (size_t)a * ld, Xs + (size_t)b * ld, Js + (size_t)a * ld, Js + (size_t)b * ld);
                            worst = fmax(worst, rel);
                        }
                        __syncthreads();
                    }
                }
                for (int idx = threadIdx.x; idx < 2 * br * rowlen2; idx += JB_THREADS) {
                    int r = idx / rowlen2, c2 = idx - r * rowlen2;
                    int g = (r < br ? A * br + r : B * br + (r - br));
                    reinterpret_cast<double2*>(p.X + (size_t)g * ld)[c2] = reinterpret_cast<const double2*>(Xs + (size_t)r * ld)[c2];
                    reinterpret_cast<double2*>(p.J + (size_t)g * ld)[c2] = reinterpret_cast<const double2*>(Js + (size_t)r * ld)[c2];
                }
                __syncthreads();
            }
            jb_grid_barrier(bar, bar_target, gridDim.x);
        }
        if (lane == 0 && worst > 0.0) atomic_max_nonneg(p.conv + sweep, worst);
        jb_grid_barrier(bar, bar_target, gridDim.x);
        double wv = *((volatile double*)(p.conv + sweep));
        if (wv <= p.tol) { sweep++; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) p.info[0] = sweep;
#undef JB_ROTATE
}

// ---------------------------------------------------------------------------------------------------------------------
// Small cores (q <= JS_MAX_Q: X and J fit the shared memory of one SM): the whole SVD -- initialisation from R, the cyclic
// one-sided Jacobi sweeps, singular values, sort, truncation count and the emission of W^T and J^T -- in ONE kernel on ONE
// CTA.  No grid barrier and no L2 round trip per phase: a rotation step costs one __syncthreads.  A row pair is owned by a
// group of JS_GROUP lanes, so the n/2 <= 56 independent pairs of a step all run at once on the CTA's 64 groups.
// At q = 66 (D = 4, chi = 64) the multi-CTA kernel above took 0.73 ms = 27 % of a sweep (profiles/r02_small_configs.txt).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int JS_THREADS = 512;
constexpr int JS_GROUP = 8;      // lanes per row pair.  FP64 warp instructions cost the same pipe time however many lanes are active, and the
                                 // scalar rotation chain (sqrt, division, rsqrt: ~60 of them) is per pair: four pairs per warp instead of two
                                 // halve the instruction stream of a step; with 16 lanes per pair the step was bound by the SM's FP64 issue rate
constexpr int JS_MAX_Q = 112;
constexpr int JS_NIT = (JS_MAX_Q / 2 + JS_GROUP - 1) / JS_GROUP;   // double2 per lane and row (X rows rotate out of registers)

struct JacobiSmallParams {
    const double* R;     // q x q, row major
    int q, n, ld;        // n = q rounded up to even rows (a zero row never rotates), ld = even row pitch
    int max_sweeps, chi;
    double tol, cutoff;
    double* S;           // [q] descending
    double* Wt;          // [q][q]
    double* Jt;          // [q][q]
    int* count;          // [0] kept rank, [1] sweeps used
};

__global__ void __launch_bounds__(JS_THREADS, 1) jacobi_small_kernel(const JacobiSmallParams p) {
    extern __shared__ __align__(16) double sm[];
    const int n = p.n, ld = p.ld, q = p.q;
    double* Xs = sm;                         // [n][ld]
    double* Js = Xs + (size_t)n * ld;        // [n][ld]
    double* sig = Js + (size_t)n * ld;       // [n]
    double* Ssorted = sig + n;               // [n]
    int* rank = reinterpret_cast<int*>(Ssorted + n);   // [n]
    const int tid = threadIdx.x, grp = tid / JS_GROUP, gl = tid % JS_GROUP;
    for (int idx = tid; idx < n * ld; idx += JS_THREADS) {
        const int r = idx / ld, c = idx - r * ld;
        Xs[idx] = (r < q && c < q) ? p.R[(size_t)r * q + c] : 0.0;
        Js[idx] = (r == c && r < q) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int npairs = n / 2, nm1 = n - 1, ncols2 = ld / 2;
    int sweep = 0;
    for (; sweep < p.max_sweeps; sweep++) {
        int rotated = 0;
        for (int t = 0; t < nm1; t++) {
            if (grp < npairs) {
                int a, b;
                if (grp == 0) { a = nm1; b = t; }
                else { a = (t + grp) % nm1; b = (t - grp + nm1) % nm1; }
                rotated |= rotate_pair_cached<JS_GROUP, JS_NIT, false>(Xs + (size_t)a * ld, Xs + (size_t)b * ld, Js + (size_t)a * ld, Js + (size_t)b * ld,
                                                                        ncols2, p.tol, gl);
            }
            __syncthreads();
        }
        if (!__syncthreads_or(rotated)) { sweep++; break; }    // a sweep without any rotation: converged
    }
    // singular values = row norms; sort descending (rank by counting, stable); emit normalised rows
    for (int r = grp; r < n; r += JS_THREADS / JS_GROUP) {
        double s2 = 0.0;
        for (int c = gl; c < q; c += JS_GROUP) { const double v = Xs[(size_t)r * ld + c]; s2 += v * v; }
#pragma unroll
        for (int o = JS_GROUP / 2; o > 0; o >>= 1) s2 += __shfl_xor_sync(((1u << JS_GROUP) - 1u) << (threadIdx.x & (32 - JS_GROUP)), s2, o);
        if (gl == 0) sig[r] = sqrt(s2);
    }
    __syncthreads();
    for (int r = tid; r < n; r += JS_THREADS) {
        const double sr = sig[r];
        int k = 0;
        for (int o = 0; o < n; o++) { const double so = sig[o]; k += (so > sr) || (so == sr && o < r); }
        rank[r] = k;
        Ssorted[k] = sr;
        if (k < q) p.S[k] = sr;
    }
    __syncthreads();
    for (int r = grp; r < n; r += JS_THREADS / JS_GROUP) {
        const int k = rank[r];
        if (k >= q) continue;
        const double sr = sig[r];
        const double inv = sr > 0.0 ? 1.0 / sr : 0.0;
        for (int c = gl; c < q; c += JS_GROUP) {
            p.Wt[(size_t)k * q + c] = Xs[(size_t)r * ld + c] * inv;
            p.Jt[(size_t)k * q + c] = Js[(size_t)r * ld + c];
        }
    }
    if (tid == 0) {
        const double s0 = Ssorted[0];
        int k = 0;
        for (int i = 0; i < q; i++) k += (Ssorted[i] / s0 > p.cutoff) ? 1 : 0;
        p.count[0] = k < p.chi ? k : p.chi;
        p.count[1] = sweep;
    }
}

__global__ void jacobi_init_kernel(double* X, double* J, const double* R, int q, int n, int ld) {
    size_t total = (size_t)n * ld;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / ld), c = (int)(i - (size_t)r * ld);
        X[i] = (r < q && c < q) ? R[(size_t)r * q + c] : 0.0;
        J[i] = (r == c && r < q) ? 1.0 : 0.0;
    }
}

// singular values = row norms; sort descending (rank by counting, stable); emit normalised rows.
__global__ void jacobi_finalize_kernel(const double* __restrict__ X, const double* __restrict__ J, int q, int n, int ld,
                                       double* __restrict__ S, double* __restrict__ Wt, double* __restrict__ Jt, int chi,
                                       double cutoff, int* __restrict__ count, const int* __restrict__ info,
                                       double* __restrict__ sig /*[n] scratch*/, int* __restrict__ rank /*[n] scratch*/,
                                       int phase) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    if (phase == 0) {
        for (int r = gw; r < n; r += nw) {
            double s = 0.0;
            for (int c = lane; c < q; c += 32) { double v = X[(size_t)r * ld + c]; s += v * v; }
            s = warp_sum(s);
            if (lane == 0) sig[r] = sqrt(s);
        }
    } else if (phase == 1) {
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
            double sr = sig[r];
            int k = 0;
            for (int o = 0; o < n; o++) { double so = sig[o]; k += (so > sr) || (so == sr && o < r); }
            rank[r] = k;
            if (k < q) S[k] = sr;
        }
    } else if (phase == 2) {
        for (int r = gw; r < n; r += nw) {
            int k = rank[r];
            if (k >= q) continue;
            double sr = sig[r];
            double inv = sr > 0.0 ? 1.0 / sr : 0.0;
            for (int c = lane; c < q; c += 32) {
                Wt[(size_t)k * q + c] = X[(size_t)r * ld + c] * inv;
                Jt[(size_t)k * q + c] = J[(size_t)r * ld + c];
            }
        }
    } else {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            double s0 = S[0];
            int k = 0;
            for (int i = 0; i < q; i++) k += (S[i] / s0 > cutoff) ? 1 : 0;
            count[0] = k < chi ? k : chi;
            count[1] = info[0];
        }
    }
}

namespace {
struct JacobiGeom { int br, nb, n, ld; size_t smem; };
JacobiGeom jacobi_geom(int q) {
    JacobiGeom g;
    g.ld = q + (q & 1);                                  // even row pitch: 16-byte staging copies
    g.br = 8;                                            // 8-row blocks: more CTAs (FP64 throughput of the active SMs is the limiter), 2 warps/SMSP
    while (g.br > 2 && (size_t)4 * g.br * g.ld * sizeof(double) > 220 * 1024) g.br /= 2;
    g.nb = (q + g.br - 1) / g.br;
    if (g.nb & 1) g.nb++;
    if (g.nb < 2) g.nb = 2;
    g.n = g.nb * g.br;
    g.smem = (size_t)4 * g.br * g.ld * sizeof(double);
    return g;
}
}  // namespace

size_t jacobi_workspace_bytes(int q) {
    JacobiGeom g = jacobi_geom(q);
    return ws_round((size_t)g.n * g.ld * sizeof(double)) * 2 + ws_round(64 * sizeof(double)) + ws_round((size_t)g.n * sizeof(double)) +
           ws_round((size_t)g.n * sizeof(int)) + ws_round(16 * sizeof(int)) + 1024;
}

int jacobi_svd_launch(const double* R, int q, double* S, double* Wt, double* Jt, int chi, double cutoff, int* count,
                      void* wsp, size_t ws_bytes, cudaStream_t s) {
    AB_REQUIRE(q >= 1 && q <= 3400, "jacobi_svd: q=%d out of range (1..3400)", q);
    const int max_sweeps = 60;
    {
        // small cores: everything in one kernel on one CTA (ACETN_B200_JACOBI_SMALL=0 keeps the multi-CTA kernel: dev / test knob)
        const char* e = getenv("ACETN_B200_JACOBI_SMALL");
        if (q <= JS_MAX_Q && !(e != nullptr && e[0] == '0')) {
            JacobiSmallParams sp;
            sp.R = R; sp.q = q; sp.n = q + (q & 1); sp.ld = q + (q & 1); sp.max_sweeps = max_sweeps; sp.chi = chi;
            sp.tol = 2.3e-16 * sqrt((double)(q > 64 ? q : 64)); sp.cutoff = cutoff; sp.S = S; sp.Wt = Wt; sp.Jt = Jt; sp.count = count;
            const size_t smem = ((size_t)2 * sp.n * sp.ld + 2 * sp.n) * sizeof(double) + (size_t)sp.n * sizeof(int) + 16;
            AB_ENSURE_SMEM(jacobi_small_kernel, smem);
            jacobi_small_kernel<<<1, JS_THREADS, smem, s>>>(sp);
            AB_LAUNCHED();
            return OK;
        }
    }
    const JacobiGeom g = jacobi_geom(q);
    Workspace ws(wsp, ws_bytes);
    double* X = ws.take<double>((size_t)g.n * g.ld);
    double* J = ws.take<double>((size_t)g.n * g.ld);
    double* conv = ws.take<double>(64);
    double* sig = ws.take<double>(g.n);
    int* rank = ws.take<int>(g.n);
    int* info = ws.take<int>(16);
    if (ws.overflow) { set_error("jacobi_svd: workspace too small"); return ERR_WORKSPACE; }
    AB_CHECK_CUDA(cudaMemsetAsync(conv, 0, 64 * sizeof(double), s));
    AB_CHECK_CUDA(cudaMemsetAsync(info, 0, 16 * sizeof(int), s));
    jacobi_init_kernel<<<148, 256, 0, s>>>(X, J, R, q, g.n, g.ld);
    AB_LAUNCHED();
    {
        AB_ENSURE_SMEM(jacobi_block_kernel, g.smem);
        JacobiParams p;
        p.X = X; p.J = J; p.n = g.n; p.ld = g.ld; p.ncols = q; p.br = g.br; p.nb = g.nb; p.max_sweeps = max_sweeps;
        p.tol = 2.3e-16 * sqrt((double)(q > 64 ? q : 64)); p.conv = conv; p.info = info;
        int blocks = g.nb / 2;
        int cap = device_sm_count();
        if (blocks > cap) blocks = cap;
        void* args[] = {&p};
        AB_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_block_kernel, dim3(blocks), dim3(JB_THREADS), args, g.smem, s));
        note_launch(1);
    }
    for (int phase = 0; phase < 4; phase++) {
        jacobi_finalize_kernel<<<phase == 3 ? 1 : 64, 256, 0, s>>>(X, J, q, g.n, g.ld, S, Wt, Jt, chi, cutoff, count, info, sig, rank, phase);
        AB_LAUNCHED();
    }
    return OK;
}

}  // namespace ab200
