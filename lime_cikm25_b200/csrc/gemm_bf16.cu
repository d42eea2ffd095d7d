// placeholder, replaced by the tcgen05 implementation
#include "common.cuh"
extern "C" int lime_linear_bf16(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias,
                                const float *residual, int64_t ldr, float *C, int64_t ldc, int64_t m,
                                int n, int k, int act, void *stream) {
    lime::set_error("lime_linear_bf16: not implemented yet");
    return 9;
}
