// TEST INFRASTRUCTURE -- host stand-ins for the kernel launchers that acetn_b200/csrc/environment.cu composes (K1 GEMM, gather),
// so that the HOST-SIDE index logic of the C-ABI environment contractions (acetn_b200_site_rdm / _bond_rdm / _norm_tensor: chi-leg
// validation, workspace carving, two-level index descriptors, leg gathers) can be executed in a container without a GPU: the
// "device" pointers are host pointers and every descriptor is evaluated literally by loops.  Linked ONLY into
// tests/mock_backend/_build/libenvmock.so by tests/test_environment_host_cpu.py; the product library never contains this file.
#include <stdarg.h>
#include <stdio.h>
#include <vector>

#include "../../acetn_b200/csrc/gemm.cuh"
#include "../../acetn_b200/csrc/kernels.cuh"

namespace ab200 {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

static long long g_launches = 0;
static size_t g_fake_splitk_bytes = 0;

// a non-zero scratch requirement for every GEMM exercises the workspace accounting of the composition
size_t gemm_workspace_bytes(const GemmDesc& d) {
    if (d.M <= 0 || d.N <= 0 || d.batch <= 0) return 0;
    return g_fake_splitk_bytes;
}

int gemm_launch(const GemmDesc& d, void* ws, size_t ws_bytes, cudaStream_t) {
    if (d.M <= 0 || d.N <= 0 || d.batch <= 0) return OK;
    if (g_fake_splitk_bytes > 0) {
        if (ws == nullptr || ws_bytes < g_fake_splitk_bytes) { set_error("mock gemm: scratch too small (%zu < %zu)", ws_bytes, g_fake_splitk_bytes); return ERR_WORKSPACE; }
        memset(ws, 0xff, g_fake_splitk_bytes);          // scribble: a scratch area overlapping a live buffer shows up as NaNs
    }
    g_launches++;
    for (int b = 0; b < d.batch; b++) {
        const int64_t ab = d.A.batch.off((uint32_t)b), bb = d.B.batch.off((uint32_t)b), cb = d.cb.off((uint32_t)b);
        for (int m = 0; m < d.M; m++) {
            const int64_t am = d.A.row.off((uint32_t)m), cm = d.cm.off((uint32_t)m);
            for (int n = 0; n < d.N; n++) {
                const int64_t bn = d.B.col.off((uint32_t)n);
                double acc = 0.0;
                for (int k = 0; k < d.K; k++) acc += d.A.ptr[ab + am + d.A.col.off((uint32_t)k)] * d.B.ptr[bb + d.B.row.off((uint32_t)k) + bn];
                double* c = d.C + cb + cm + d.cn.off((uint32_t)n);
                *c = d.alpha * acc + (d.beta != 0.0 ? d.beta * *c : 0.0);
            }
        }
    }
    return OK;
}

int gather_nd_launch(double* dst, const double* src, int nd, const int64_t* dims, const int64_t* strides, cudaStream_t) {
    if (nd < 1 || nd > 8) { set_error("gather_nd: 1..8 dims supported (got %d)", nd); return ERR_INVALID; }
    g_launches++;
    size_t total = 1;
    for (int i = 0; i < nd; i++) total *= (size_t)dims[i];
    std::vector<int64_t> idx(nd, 0);
    for (size_t t = 0; t < total; t++) {
        int64_t off = 0;
        for (int i = 0; i < nd; i++) off += idx[i] * strides[i];
        dst[t] = src[off];
        for (int i = nd - 1; i >= 0; i--) { if (++idx[i] < dims[i]) break; idx[i] = 0; }
    }
    return OK;
}

}  // namespace ab200

extern "C" {
const char* envmock_last_error(void) { return ab200::get_error(); }
long long envmock_launches(void) { return ab200::g_launches; }
void envmock_set_fake_splitk_bytes(size_t n) { ab200::g_fake_splitk_bytes = n; }
}
