// K5: one-sided (Hestenes) Jacobi SVD of the small square core of the randomized SVD
// (reference: torch.linalg.svd of the projected matrix, acetn/linalg/fused_matmul_svd_lowrank.py:48, after the
// tall factor has been reduced to a (chi+p) x (chi+p) core by K4 + one DGEMM).
//
// Rows of X are rotated pairwise until mutually orthogonal:  J X = diag(S) W.  One warp owns one row pair per
// round-robin step (n/2 independent pairs per step, n-1 steps per sweep); the matrix (n <= ~1k, <= 8 MB with J)
// stays L2 resident and the steps are separated by a cooperative grid barrier.  Rotations use the relative
// criterion |x_p.x_q| <= tol ||x_p|| ||x_q||, which gives high relative accuracy of small singular values.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace ab200 {

constexpr int JC_WARPS = 4;   // warps per CTA

struct JacobiParams {
    double* X;      // n x n (padded even), row major, ld = n
    double* J;      // n x n
    int n;
    int max_sweeps;
    double tol;
    double* conv;   // [max_sweeps] max relative off-diagonal seen in each sweep
    int* info;      // [0] sweeps used
};

__global__ void __launch_bounds__(JC_WARPS * 32) jacobi_rows_kernel(JacobiParams p) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * JC_WARPS + (threadIdx.x >> 5);
    const int nw = gridDim.x * JC_WARPS;
    const int n = p.n, half = n / 2, nm1 = n - 1;

    int sweep = 0;
    for (; sweep < p.max_sweeps; sweep++) {
        double worst = 0.0;
        for (int step = 0; step < nm1; step++) {
            for (int i = gw; i < half; i += nw) {
                int a, b;
                if (i == 0) { a = nm1; b = step; }
                else { a = (step + i) % nm1; b = (step - i + nm1) % nm1; }
                double* xp = p.X + (size_t)a * n;
                double* xq = p.X + (size_t)b * n;
                double saa = 0.0, sbb = 0.0, sab = 0.0;
                for (int c = lane; c < n; c += 32) {
                    double u = __ldcg(xp + c), v = __ldcg(xq + c);
                    saa += u * u; sbb += v * v; sab += u * v;
                }
                saa = warp_sum(saa); sbb = warp_sum(sbb); sab = warp_sum(sab);
                if (saa == 0.0 || sbb == 0.0) continue;
                double rel = fabs(sab) / sqrt(saa * sbb);
                worst = fmax(worst, rel);
                if (rel <= p.tol) continue;
                double zeta = (sbb - saa) / (2.0 * sab);
                double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                double* jp = p.J + (size_t)a * n;
                double* jq = p.J + (size_t)b * n;
                for (int c = lane; c < n; c += 32) {
                    double u = __ldcg(xp + c), v = __ldcg(xq + c);
                    xp[c] = cs * u - sn * v;
                    xq[c] = sn * u + cs * v;
                    double ju = __ldcg(jp + c), jv = __ldcg(jq + c);
                    jp[c] = cs * ju - sn * jv;
                    jq[c] = sn * ju + cs * jv;
                }
            }
            grid.sync();
        }
        if (lane == 0 && worst > 0.0) atomic_max_nonneg(p.conv + sweep, worst);
        grid.sync();
        double w = *((volatile double*)(p.conv + sweep));
        if (w <= p.tol) { sweep++; break; }
    }
    if (gw == 0 && lane == 0) p.info[0] = sweep;
}

__global__ void jacobi_init_kernel(double* X, double* J, const double* R, int q, int n) {
    size_t total = (size_t)n * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(i / n), c = (int)(i - (size_t)r * n);
        X[i] = (r < q && c < q) ? R[(size_t)r * q + c] : 0.0;
        J[i] = (r == c) ? 1.0 : 0.0;
    }
}

// singular values = row norms; sort descending (rank by counting, stable); emit normalised rows.
__global__ void jacobi_finalize_kernel(const double* __restrict__ X, const double* __restrict__ J, int q, int n,
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
            for (int c = lane; c < n; c += 32) { double v = X[(size_t)r * n + c]; s += v * v; }
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
                Wt[(size_t)k * q + c] = X[(size_t)r * n + c] * inv;
                Jt[(size_t)k * q + c] = J[(size_t)r * n + c];
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

size_t jacobi_workspace_bytes(int q) {
    int n = q + (q & 1);
    return ws_round((size_t)n * n * sizeof(double)) * 2 + ws_round(64 * sizeof(double)) + ws_round((size_t)n * sizeof(double)) +
           ws_round((size_t)n * sizeof(int)) + ws_round(16 * sizeof(int)) + 1024;
}

int jacobi_svd_launch(const double* R, int q, double* S, double* Wt, double* Jt, int chi, double cutoff, int* count,
                      void* wsp, size_t ws_bytes, cudaStream_t s) {
    AB_REQUIRE(q >= 1 && q <= 4096, "jacobi_svd: q=%d out of range", q);
    const int n = q + (q & 1);
    const int max_sweeps = 60;
    Workspace ws(wsp, ws_bytes);
    double* X = ws.take<double>((size_t)n * n);
    double* J = ws.take<double>((size_t)n * n);
    double* conv = ws.take<double>(64);
    double* sig = ws.take<double>(n);
    int* rank = ws.take<int>(n);
    int* info = ws.take<int>(16);
    if (ws.overflow) { set_error("jacobi_svd: workspace too small"); return ERR_WORKSPACE; }
    AB_CHECK_CUDA(cudaMemsetAsync(conv, 0, 64 * sizeof(double), s));
    AB_CHECK_CUDA(cudaMemsetAsync(info, 0, 16 * sizeof(int), s));
    jacobi_init_kernel<<<148, 256, 0, s>>>(X, J, R, q, n);
    AB_LAUNCHED();
    if (n >= 2) {
        JacobiParams p;
        p.X = X; p.J = J; p.n = n; p.max_sweeps = max_sweeps; p.tol = 2.3e-16 * sqrt((double)(n > 64 ? n : 64)); p.conv = conv; p.info = info;
        int half = n / 2;
        int blocks = (half + JC_WARPS - 1) / JC_WARPS;
        int maxb = 0;
        AB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, jacobi_rows_kernel, JC_WARPS * 32, 0));
        int cap = maxb * device_sm_count();
        if (cap < 1) cap = 1;
        if (blocks > cap) blocks = cap;
        void* args[] = {&p};
        AB_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)jacobi_rows_kernel, dim3(blocks), dim3(JC_WARPS * 32), args, 0, s));
        note_launch(1);
    }
    for (int phase = 0; phase < 4; phase++) {
        jacobi_finalize_kernel<<<phase == 3 ? 1 : 64, 256, 0, s>>>(X, J, q, n, S, Wt, Jt, chi, cutoff, count, info, sig, rank, phase);
        AB_LAUNCHED();
    }
    return OK;
}

}  // namespace ab200
