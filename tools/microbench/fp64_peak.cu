// FP64 issue-rate microbenchmark for B200 (sm_100a): DMMA.8x8x4 vs DFMA.
// Bench-only evidence for the roofline denominator (SURVEY.md §8d asks for a measured FP64 roof).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int NACC>
__global__ void dmma_loop(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_loop(double* out, int iters, double seed) {
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-12;
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warps with (warp % 2 == 0) issue DMMA, the others DFMA: if the two are separate pipes the chip total exceeds either alone
template <int NACC>
__global__ void mixed_loop(double* out, int iters, double seed, int dfma_per_dmma) {
    const int warp = threadIdx.x >> 5;
    double a = seed + threadIdx.x * 1e-9, b = 1.0 - 1e-12;
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
    if ((warp & 1) == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < NACC; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    } else {
        for (int it = 0; it < iters * dfma_per_dmma; it++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) { c[i][0] = fma(c[i][0], b, a); c[i][1] = fma(c[i][1], b, a); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock_khz %d\n", p.name, sms, p.clockRate);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    const int iters = 20000;
    if (getenv("FP64_MIX_ONLY")) {
        // 16 warps per CTA, 1 CTA per SM: 8 DMMA warps (enough to saturate the DMMA sub-pipe alone) + 8 DFMA warps
        for (int ratio = 1; ratio <= 8; ratio *= 2) {
            dim3 grid(sms), block(512);
            float ms = time_ms([&] { mixed_loop<8><<<grid, block>>>(out, iters, 1.0, ratio); }, 3);
            double fl_dmma = 2.0 * 256 * 8 * (double)iters * 8 * sms;
            double fl_dfma = 2.0 * 32 * 16 * (double)iters * ratio * 8 * sms;
            printf("MIX 8 DMMA warps + 8 DFMA warps per SM, %d x16 DFMA per 8 DMMA: %.3f ms  DMMA part %.2f TF + DFMA part %.2f TF = %.2f TF if fully overlapped\n",
                   ratio, ms, fl_dmma / ms * 1e-9, fl_dfma / ms * 1e-9, (fl_dmma + fl_dfma) / ms * 1e-9);
        }
        float md = time_ms([&] { dmma_loop<8><<<dim3(sms), dim3(256)>>>(out, iters, 1.0); }, 3);
        float mf = time_ms([&] { dfma_loop<16><<<dim3(sms), dim3(256)>>>(out, iters * 8, 1.0); }, 3);
        printf("alone: 8 DMMA warps %.3f ms (%.2f TF), 8 DFMA warps x8 iters %.3f ms (%.2f TF)\n", md,
               2.0 * 256 * 8 * (double)iters * 8 * sms / md * 1e-9, mf, 2.0 * 32 * 16 * (double)iters * 8 * 8 * sms / mf * 1e-9);
        return 0;
    }
    int warps_list[] = {1, 2, 4, 8, 16, 32};
    for (int wi = 0; wi < 6; wi++) {
        int warps = warps_list[wi];
        for (int ctas = 1; ctas <= 2; ctas++) {
            if (warps * ctas > 32) continue;
            dim3 grid(sms * ctas), block(warps * 32);
            float ms1 = time_ms([&] { dmma_loop<1><<<grid, block>>>(out, iters, 1.0); }, 3);
            float ms4 = time_ms([&] { dmma_loop<4><<<grid, block>>>(out, iters, 1.0); }, 3);
            float ms8 = time_ms([&] { dmma_loop<8><<<grid, block>>>(out, iters, 1.0); }, 3);
            double fl = 2.0 * 8 * 8 * 4 * (double)iters * warps * ctas * sms;
            printf("DMMA warps/cta %2d ctas/sm %d : acc1 %7.2f TF  acc4 %7.2f TF  acc8 %7.2f TF | cyc/dmma/warp(acc1)@1.965GHz %.1f\n",
                   warps, ctas, fl * 1 / ms1 * 1e-9, fl * 4 / ms4 * 1e-9, fl * 8 / ms8 * 1e-9,
                   ms1 * 1e-3 * 1.965e9 / iters);
            float f1 = time_ms([&] { dfma_loop<1><<<grid, block>>>(out, iters, 1.0); }, 3);
            float f8 = time_ms([&] { dfma_loop<8><<<grid, block>>>(out, iters, 1.0); }, 3);
            float f16 = time_ms([&] { dfma_loop<16><<<grid, block>>>(out, iters, 1.0); }, 3);
            double ff = 2.0 * 32 * (double)iters * warps * ctas * sms;
            printf("DFMA warps/cta %2d ctas/sm %d : acc1 %7.2f TF  acc8 %7.2f TF  acc16 %7.2f TF\n",
                   warps, ctas, ff * 1 / f1 * 1e-9, ff * 8 / f8 * 1e-9, ff * 16 / f16 * 1e-9);
        }
    }
    // sustained DMMA for ~3 s to see clocks under FP64 load
    {
        dim3 grid(sms * 2), block(256);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int n = 0;
        for (; n < 60; n++) dmma_loop<8><<<grid, block>>>(out, 200000, 1.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 256 * 8 * 200000.0 * 8 * 2 * sms * n;
        printf("DMMA sustained %.1f ms : %.2f TF\n", ms, fl / ms * 1e-9);
    }
    return 0;
}
