// acetn_b200 -- common device/host helpers for the sm_100a CTMRG kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <stdio.h>
#include <string.h>

namespace ab200 {

// ---- status / error reporting (C-ABI never throws; see include/acetn_b200.h) --------------------
enum Status : int { OK = 0, ERR_INVALID = 1, ERR_WORKSPACE = 2, ERR_CUDA = 3, ERR_UNSUPPORTED = 4 };

void set_error(const char* fmt, ...);
const char* get_error();

#define AB_CHECK_CUDA(expr)                                                                     \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ab200::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,              \
                             cudaGetErrorString(_e));                                           \
            return ab200::ERR_CUDA;                                                             \
        }                                                                                       \
    } while (0)

#define AB_REQUIRE(cond, ...)                                                                   \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            ab200::set_error(__VA_ARGS__);                                                      \
            return ab200::ERR_INVALID;                                                          \
        }                                                                                       \
    } while (0)

#define AB_TRY(expr)                                                                            \
    do {                                                                                        \
        int _s = (expr);                                                                        \
        if (_s != 0) return _s;                                                                 \
    } while (0)

// ---- two-level index: off(i) = (i / div) * s_hi + (i % div) * s_lo  (div == 0: i * s_lo) ---------
struct Idx2 {
    int64_t s_hi;
    int64_t s_lo;
    uint32_t div;
    uint32_t pad_;
    __host__ __device__ __forceinline__ int64_t off(uint32_t i) const {
        if (div == 0) return (int64_t)i * s_lo;
        uint32_t hi = i / div;
        uint32_t lo = i - hi * div;
        return (int64_t)hi * s_hi + (int64_t)lo * s_lo;
    }
};
static inline Idx2 idx1(int64_t stride) { Idx2 r; r.s_hi = 0; r.s_lo = stride; r.div = 0; r.pad_ = 0; return r; }
static inline Idx2 idx2(int64_t div, int64_t s_hi, int64_t s_lo) {
    Idx2 r; r.s_hi = s_hi; r.s_lo = s_lo; r.div = (uint32_t)div; r.pad_ = 0; return r;
}

// ---- workspace bump allocator over a caller-provided device buffer --------------------------------
struct Workspace {
    char* base;
    size_t bytes;
    size_t used;
    bool overflow;
    Workspace(void* p, size_t n) : base((char*)p), bytes(n), used(0), overflow(false) {}
    template <typename T>
    T* take(size_t count) {
        size_t need = (count * sizeof(T) + 255) & ~(size_t)255;
        if (base == nullptr || used + need > bytes) { overflow = true; used += need; return nullptr; }
        T* r = (T*)(base + used);
        used += need;
        return r;
    }
};
static inline size_t ws_round(size_t count_bytes) { return (count_bytes + 255) & ~(size_t)255; }

// ---- device helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16-byte async copy, src_bytes in {0,8,16}: the remainder is zero-filled
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// order-preserving atomic max for non-negative doubles
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}

int device_sm_count();
void note_launch(int n);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: the "already configured" state of a kernel is kept per
// device (the host shim initialises the library for every device index it meets) and is safe to race on from several host threads
// (setting the attribute twice is harmless).  One SmemConfig per kernel instantiation, as a function-local static.
constexpr int AB_MAX_DEVICES = 64;
struct SmemConfig {
    std::atomic<size_t> bytes[AB_MAX_DEVICES];
    SmemConfig() { for (auto& b : bytes) b.store(0, std::memory_order_relaxed); }
};
int ensure_dynamic_smem(const void* kernel, SmemConfig& cfg, size_t bytes);
#define AB_ENSURE_SMEM(kern, bytes)                                                              \
    do {                                                                                         \
        static ab200::SmemConfig _cfg;                                                           \
        AB_TRY(ab200::ensure_dynamic_smem((const void*)(kern), _cfg, (size_t)(bytes)));          \
    } while (0)
#define AB_LAUNCHED()                                   \
    do {                                               \
        ab200::note_launch(1);                         \
        AB_CHECK_CUDA(cudaGetLastError());             \
    } while (0)

}  // namespace ab200
