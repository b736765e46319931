// conv.cu — Conv2D entry points (layers/conv.py). Placeholder until the implicit-GEMM kernels land.
#include "common.cuh"
using namespace npm;
extern "C" {
size_t npm_conv2d_workspace(int64_t, int64_t, int64_t, int64_t, int64_t, int) { return 0; }
int npm_conv2d_fwd(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, int64_t, int,
                   int, void*, npm_stream_t) {
    set_error("conv2d_fwd: not built yet");
    return NPM_ERR_UNSUPPORTED;
}
int npm_conv2d_bwd_dx(const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, int64_t, int, void*,
                      npm_stream_t) {
    set_error("conv2d_bwd_dx: not built yet");
    return NPM_ERR_UNSUPPORTED;
}
int npm_conv2d_bwd_dw_db(const float*, const float*, float*, float*, int64_t, int64_t, int64_t, int64_t, int64_t, int,
                         void*, npm_stream_t) {
    set_error("conv2d_bwd_dw_db: not built yet");
    return NPM_ERR_UNSUPPORTED;
}
}
