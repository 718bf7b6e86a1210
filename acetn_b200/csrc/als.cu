// K6: persistent ALS inner solver of the full update (reference acetn/evolution/als_solver.py:55-82 = the C++ loop
// csrc/evolution/als_solve.cpp:55-105 with its 9 cuTENSOR plans + cuSOLVER potrf/potrs per iteration and one host
// sync per iteration).  Here the WHOLE loop -- normal-equation contractions, symmetrise + regularise, Cholesky,
// two triangular solves, cost, convergence test -- runs in ONE cooperative kernel; all operands (<= 1 MB) stay L2
// resident, the host is not involved until the loop has converged.
//
//   a1r[y,u,p], a2r[x,u,q] (nD,bD,pD);  n12[y,x,Y,X] (nD^4);  n12g[Y,X,p,q];  a12g[y,x,p,q]
//   S1[Y,U,p]   = sum n12g[Y,X,p,Q] a2r[X,U,Q]            R1[(Y,U),(y,u)] = sum n12[y,x,Y,X] a2r[x,u,q] a2r[X,U,q]
//   S2[X,V,q]   = sum n12g[Y,X,P,q] a1r[Y,V,P]            R2[(X,V),(x,v)] = sum n12[y,x,Y,X] a1r[y,v,p] a1r[Y,V,p]
//   method "cholesky": R <- (R+R^T)/2 + eps*max|R|*I ;  R a = S by Cholesky
//   method "pinv" (als_solver.py:226-228 = als_solve.cpp:47-50): a = pinv((R+R^T)/2, hermitian, rcond = eps) S -- symmetric
//     eigen-decomposition by a parallel two-sided Jacobi in the shared memory of CTA 0, eigenvalues below eps*max|w| dropped
//   cost = <a12n|N|a12n> - 2 <a12n|N|a12g>
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace cg = cooperative_groups;

namespace ab200 {

constexpr int ALS_THREADS = 512;

struct AlsParams {
    double* a1r; double* a2r;
    const double* n12g; const double* n12; const double* a12g;
    int nD, bD, pD, niter;
    double tol, epsilon;
    double* G;        // n x n
    double* R;        // n x n
    double* S;        // n x pD
    double* part;     // [grid][2] cost partials
    int* info;        // [0] iterations run, [1] cholesky failures
    int chol_in_smem;
    int method;       // 0 = cholesky, 1 = pinv
    double* V;        // n x n eigenvectors (pinv; global, L2 resident)
};

__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < ALS_THREADS / 32; i++) s += red[i];
    return s;
}
__device__ __forceinline__ double block_reduce_max(double v, double* red) {
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < ALS_THREADS / 32; i++) s = fmax(s, red[i]);
    return s;
}

// a = pinv(M, hermitian, rcond) S for the symmetric n x n matrix M (shared memory, pitch ldm), CTA-wide.  Parallel two-sided Jacobi:
// the n/2 disjoint pairs of a round-robin step get their rotation from (m_pp, m_qq, m_pq); row phase M <- J^T M, column phase
// M <- M J and V <- V J; a sweep that rotates nothing ends the iteration.  Absolute accuracy eps*||M|| like a LAPACK symmetric
// eigensolver.  S (n x pD, global) is overwritten with the solution.  rot: shared scratch [n] (cos, sin per pair).
__device__ void als_pinv_solve(double* M, int ldm, int n, double* V, double* S, int pD, double rcond, double* rot, double* red) {
    const int tid = threadIdx.x, T = blockDim.x;
    const int np = n + (n & 1);                      // an odd n gets a bye
    const int npairs = np / 2, nm1 = np - 1;
    for (int o = tid; o < n * n; o += T) V[o] = (o / n == o % n) ? 1.0 : 0.0;
    double fro = 0.0;
    for (int o = tid; o < n * n; o += T) { const double v = M[(o / n) * ldm + (o % n)]; fro += v * v; }
    fro = sqrt(block_reduce_sum(fro, red));
    const double small = 1e-17 * fro;                // couplings below eps*||M|| are zero to working accuracy
    __shared__ int s_rotated;
    for (int sweep = 0; sweep < 40; sweep++) {
        if (tid == 0) s_rotated = 0;
        __syncthreads();
        for (int t = 0; t < nm1; t++) {
            for (int k = tid; k < npairs; k += T) {
                int a, b;
                if (k == 0) { a = nm1; b = t; }
                else { a = (t + k) % nm1; b = (t - k + nm1) % nm1; }
                double c = 1.0, sn = 0.0;
                if (a < n && b < n) {
                    const double mpq = M[a * ldm + b];
                    if (fabs(mpq) > small) {
                        const double tau = (M[b * ldm + b] - M[a * ldm + a]) / (2.0 * mpq);
                        const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = rsqrt(1.0 + tt * tt);
                        sn = c * tt;
                        s_rotated = 1;
                    }
                }
                rot[2 * k] = c; rot[2 * k + 1] = sn;
            }
            __syncthreads();
            // rows:  M[a,:] <- c M[a,:] - s M[b,:] ;  M[b,:] <- s M[a,:] + c M[b,:]
            for (int o = tid; o < npairs * n; o += T) {
                const int k = o / n, col = o - k * n;
                int a, b;
                if (k == 0) { a = nm1; b = t; }
                else { a = (t + k) % nm1; b = (t - k + nm1) % nm1; }
                const double c = rot[2 * k], sn = rot[2 * k + 1];
                if (a < n && b < n && sn != 0.0) {
                    const double x = M[a * ldm + col], y = M[b * ldm + col];
                    M[a * ldm + col] = c * x - sn * y;
                    M[b * ldm + col] = sn * x + c * y;
                }
            }
            __syncthreads();
            // columns of M and of V
            for (int o = tid; o < npairs * n; o += T) {
                const int k = o / n, row = o - k * n;
                int a, b;
                if (k == 0) { a = nm1; b = t; }
                else { a = (t + k) % nm1; b = (t - k + nm1) % nm1; }
                const double c = rot[2 * k], sn = rot[2 * k + 1];
                if (a < n && b < n && sn != 0.0) {
                    const double x = M[row * ldm + a], y = M[row * ldm + b];
                    M[row * ldm + a] = c * x - sn * y;
                    M[row * ldm + b] = sn * x + c * y;
                    const double vx = V[row * n + a], vy = V[row * n + b];
                    V[row * n + a] = c * vx - sn * vy;
                    V[row * n + b] = sn * vx + c * vy;
                }
            }
            __syncthreads();
        }
        if (s_rotated == 0) break;
        __syncthreads();
    }
    // a = V diag(1/w, |w| > rcond max|w|) V^T S
    double wmax = 0.0;
    for (int i = tid; i < n; i += T) wmax = fmax(wmax, fabs(M[i * ldm + i]));
    wmax = block_reduce_max(wmax, red);
    // y = diag(1/w, kept) V^T S
    for (int o = tid; o < n * pD; o += T) {
        const int i = o / pD, ph = o - i * pD;
        const double w = M[i * ldm + i];
        double acc = 0.0;
        for (int r = 0; r < n; r++) acc += V[r * n + i] * S[r * pD + ph];
        // the diagonal entry is read by the other pD-1 threads of this row: the result is staged next to V, not in M
        V[n * n + o] = (fabs(w) > rcond * wmax) ? acc / w : 0.0;
    }
    __syncthreads();
    for (int o = tid; o < n * pD; o += T) {
        const int r = o / pD, ph = o - r * pD;
        double acc = 0.0;
        for (int i = 0; i < n; i++) acc += V[r * n + i] * V[n * n + i * pD + ph];
        S[o] = acc;
    }
    __syncthreads();
}

// which = 0: solve for a1r with a2r fixed ; which = 1: solve for a2r with a1r fixed
__device__ void als_half_step(const AlsParams& p, int which, cg::grid_group& grid, double* smem, double* red) {
    const int nD = p.nD, bD = p.bD, pD = p.pD, n = nD * bD;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const double* af = which == 0 ? p.a2r : p.a1r;      // the fixed tensor  [site, bond, phys]
    // ---- P1: Gram of the fixed tensor over its physical leg, and the right-hand side
    for (int o = gtid; o < n * n; o += gsz) {
        int r = o / n, c = o - r * n;
        double s = 0.0;
        for (int q = 0; q < pD; q++) s += af[r * pD + q] * af[c * pD + q];
        p.G[o] = s;
    }
    for (int o = gtid; o < n * pD; o += gsz) {
        int ph = o % pD, U = (o / pD) % bD, Z = o / (pD * bD);      // Z = Y (which 0) or X (which 1)
        double s = 0.0;
        for (int W = 0; W < nD; W++)
            for (int Q = 0; Q < pD; Q++) {
                // which 0: n12g[Y=Z, X=W, p=ph, Q] a2r[X=W,U,Q] ; which 1: n12g[Y=W, X=Z, P=Q, q=ph] a1r[Y=W,V=U,P=Q]
                double g = which == 0 ? p.n12g[((Z * nD + W) * pD + ph) * pD + Q] : p.n12g[((W * nD + Z) * pD + Q) * pD + ph];
                s += g * af[(W * bD + U) * pD + Q];
            }
        p.S[o] = s;
    }
    grid.sync();
    // ---- P2: R[(Z,U),(z,u)] = sum_{w,W} n12[...] G[(w,u),(W,U)]
    for (int o = gtid; o < n * n; o += gsz) {
        int row = o / n, col = o - row * n;
        int Z = row / bD, U = row - Z * bD, z = col / bD, u = col - z * bD;
        double s = 0.0;
        for (int w = 0; w < nD; w++) {
            const double* grow = p.G + (size_t)(w * bD + u) * n + U;
            // which 0: n12[y=z, x=w, Y=Z, X=W] ; which 1: n12[y=w, x=z, Y=W, X=Z]
            if (which == 0) {
                const double* nn = p.n12 + ((size_t)(z * nD + w) * nD + Z) * nD;
                for (int W = 0; W < nD; W++) s += nn[W] * grow[W * bD];
            } else {
                const double* nn = p.n12 + ((size_t)(w * nD + z) * nD) * nD + Z;
                for (int W = 0; W < nD; W++) s += nn[(size_t)W * nD] * grow[W * bD];
            }
        }
        p.R[o] = s;
    }
    grid.sync();
    // ---- P3 (CTA 0): symmetrise, regularise, Cholesky, solve
    if (blockIdx.x == 0) {
        const int tid = threadIdx.x, T = blockDim.x;
        const int ldm = n + 1;
        __shared__ double smem_rot[2 * 260];            // (cos, sin) per pair of a Jacobi step, n <= 512
        double* M = p.chol_in_smem ? smem : p.G;        // G is free now; (n x ldm needs n*(n+1) <= allocated (n+1)^2)
        double mx = 0.0;
        for (int o = tid; o < n * n; o += T) {
            int r = o / n, c = o - r * n;
            double v = 0.5 * (p.R[o] + p.R[c * n + r]);
            M[r * ldm + c] = v;
            mx = fmax(mx, fabs(v));
        }
        mx = block_reduce_max(mx, red);
        if (p.method == 1) {
            __syncthreads();
            als_pinv_solve(M, ldm, n, p.V, p.S, pD, p.epsilon, smem_rot, red);
            double* dstp = which == 0 ? p.a1r : p.a2r;
            for (int o = tid; o < n * pD; o += T) dstp[o] = p.S[o];
        } else {
        for (int i = tid; i < n; i += T) M[i * ldm + i] += p.epsilon * mx;
        __syncthreads();
        for (int j = 0; j < n; j++) {
            if (tid == 0) {
                double d = M[j * ldm + j];
                if (!(d > 0.0)) { atomicAdd(p.info + 1, 1); d = fabs(d) > 0.0 ? fabs(d) : 1.0; }
                M[j * ldm + j] = sqrt(d);
            }
            __syncthreads();
            const double inv = 1.0 / M[j * ldm + j];
            for (int i = j + 1 + tid; i < n; i += T) M[i * ldm + j] *= inv;
            __syncthreads();
            const int rem = n - j - 1;
            for (int o = tid; o < rem * rem; o += T) {
                int a = o / rem, b = o - a * rem;
                if (b <= a) { int i = j + 1 + a, k = j + 1 + b; M[i * ldm + k] -= M[i * ldm + j] * M[k * ldm + j]; }
            }
            __syncthreads();
        }
        // forward: L y = S ; backward: L^T a = y   (pD right-hand sides; y overwrites p.S)
        for (int j = 0; j < n; j++) {
            if (tid < pD) p.S[j * pD + tid] /= M[j * ldm + j];
            __syncthreads();
            for (int o = tid; o < (n - j - 1) * pD; o += T) {
                int i = j + 1 + o / pD, ph = o % pD;
                p.S[i * pD + ph] -= M[i * ldm + j] * p.S[j * pD + ph];
            }
            __syncthreads();
        }
        for (int j = n - 1; j >= 0; j--) {
            if (tid < pD) p.S[j * pD + tid] /= M[j * ldm + j];
            __syncthreads();
            for (int o = tid; o < j * pD; o += T) {
                int i = o / pD, ph = o % pD;
                p.S[i * pD + ph] -= M[j * ldm + i] * p.S[j * pD + ph];
            }
            __syncthreads();
        }
        double* dst = which == 0 ? p.a1r : p.a2r;
        for (int o = tid; o < n * pD; o += T) dst[o] = p.S[o];
        }
    }
    grid.sync();
}

// cost = <a12n|N|a12n> - 2 <a12n|N|a12g>   (als_solver.py:246-257); every thread returns the same value
__device__ double als_cost(const AlsParams& p, cg::grid_group& grid, double* smem, double* red) {
    const int nD = p.nD, bD = p.bD, pD = p.pD, pp = pD * pD, n2 = nD * nD;
    double* a12n = smem;                                 // [nD*nD][pD*pD]
    for (int o = threadIdx.x; o < n2 * pp; o += blockDim.x) {
        int q = o % pD, ph = (o / pD) % pD, x = (o / pp) % nD, y = o / (pp * nD);
        double s = 0.0;
        for (int u = 0; u < bD; u++) s += p.a1r[(y * bD + u) * pD + ph] * p.a2r[(x * bD + u) * pD + q];
        a12n[o] = s;
    }
    __syncthreads();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    double d2 = 0.0, d3 = 0.0;
    for (int o = gtid; o < n2 * n2; o += gsz) {
        int yx = o / n2, YX = o - yx * n2;
        double nv = p.n12[o];
        double s2 = 0.0, s3 = 0.0;
        for (int t = 0; t < pp; t++) {
            double b = a12n[YX * pp + t];
            s2 += a12n[yx * pp + t] * b;
            s3 += p.a12g[yx * pp + t] * b;
        }
        d2 += nv * s2;
        d3 += nv * s3;
    }
    d2 = block_reduce_sum(d2, red);
    d3 = block_reduce_sum(d3, red);
    if (threadIdx.x == 0) { p.part[2 * blockIdx.x] = d2; p.part[2 * blockIdx.x + 1] = d3; }
    grid.sync();
    double t2 = 0.0, t3 = 0.0;
    for (unsigned i = 0; i < gridDim.x; i++) { t2 += __ldcg(p.part + 2 * i); t3 += __ldcg(p.part + 2 * i + 1); }
    grid.sync();                                         // partials may be overwritten by the next call
    return t2 - 2.0 * t3;
}

__global__ void __launch_bounds__(ALS_THREADS, 1) als_kernel(AlsParams p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double smem[];
    __shared__ double red[ALS_THREADS / 32];
    double d1 = fabs(als_cost(p, grid, smem, red));
    int it = 0;
    for (int i = 0; i < p.niter; i++) {
        it = i + 1;
        als_half_step(p, 0, grid, smem, red);
        als_half_step(p, 1, grid, smem, red);
        double d2 = als_cost(p, grid, smem, red);
        double err = fabs(d2 - d1) / fabs(d1);
        if (err < p.tol && i > 1) break;
        d1 = d2;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) p.info[0] = it;
}

size_t als_workspace_bytes(int nD, int bD, int pD) {
    size_t n = (size_t)nD * bD;
    return ws_round((n + 1) * (n + 1) * 8) + ws_round(n * n * 8) + ws_round(n * pD * 8) + ws_round(148 * 2 * 8) +
           ws_round((n * n + n * pD) * 8) + 1024;          // + eigenvectors and the scaled V^T S of the pinv method
}

int als_solve_launch(double* a1r, double* a2r, const double* n12g, const double* n12, const double* a12g, int nD, int bD, int pD,
                     int niter, double tol, double epsilon, int method, int* info, void* wsp, size_t ws_bytes, cudaStream_t s) {
    AB_REQUIRE(nD >= 1 && bD >= 1 && pD >= 1 && pD <= 16, "als_solve: bad dims nD=%d bD=%d pD=%d", nD, bD, pD);
    AB_REQUIRE(method == 0 || method == 1, "als_solve: method must be 0 (cholesky) or 1 (pinv), got %d", method);
    AB_REQUIRE(method == 0 || nD * bD <= 512, "als_solve: pinv method supports nD*bD <= 512 (got %d)", nD * bD);
    const size_t n = (size_t)nD * bD;
    Workspace ws(wsp, ws_bytes);
    AlsParams p;
    p.a1r = a1r; p.a2r = a2r; p.n12g = n12g; p.n12 = n12; p.a12g = a12g; p.nD = nD; p.bD = bD; p.pD = pD; p.niter = niter;
    p.tol = tol; p.epsilon = epsilon; p.info = info;
    p.G = ws.take<double>((n + 1) * (n + 1));
    p.R = ws.take<double>(n * n);
    p.S = ws.take<double>(n * pD);
    p.part = ws.take<double>(148 * 2);
    p.V = ws.take<double>(n * n + n * pD);
    p.method = method;
    if (ws.overflow) { set_error("als_solve: workspace too small"); return ERR_WORKSPACE; }
    size_t chol_bytes = n * (n + 1) * 8, cost_bytes = (size_t)nD * nD * pD * pD * 8;
    p.chol_in_smem = chol_bytes <= 200 * 1024 ? 1 : 0;
    AB_REQUIRE(method == 0 || p.chol_in_smem, "als_solve: pinv method needs the normal matrix in shared memory (nD*bD <= 159)");
    size_t smem = p.chol_in_smem ? (chol_bytes > cost_bytes ? chol_bytes : cost_bytes) : cost_bytes;
    AB_REQUIRE(smem <= 220 * 1024, "als_solve: nD^2 pD^2 too large for shared memory");
    AB_ENSURE_SMEM(als_kernel, smem);
    AB_CHECK_CUDA(cudaMemsetAsync(info, 0, 2 * sizeof(int), s));
    int grid = 32;
    int cap = device_sm_count();
    if (grid > cap) grid = cap;
    void* args[] = {&p};
    AB_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)als_kernel, dim3(grid), dim3(ALS_THREADS), args, smem, s));
    note_launch(1);
    return OK;
}

}  // namespace ab200
