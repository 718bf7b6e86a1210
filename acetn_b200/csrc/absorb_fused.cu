// K2 (fused form): double-layer site absorption with the intermediate kept on chip.  (specialisations are added per (D,d))
#include "kernels.cuh"

namespace ab200 {

int double_layer_fused_supported(int64_t D, int64_t d) { (void)D; (void)d; return 0; }

int double_layer_fused_launch(const double* X, int64_t n0, int64_t n1, int64_t in_s0, int64_t in_s1, const int64_t* in_es,
                              int order, const double* A, const int64_t* a_strides, int64_t D, int64_t d, double* Y,
                              int64_t out_s0, int64_t out_s1, const int64_t* out_es, double* absmax, cudaStream_t s) {
    set_error("double_layer_fused: no specialisation for D=%lld d=%lld", (long long)D, (long long)d);
    return ERR_UNSUPPORTED;
}

}  // namespace ab200
