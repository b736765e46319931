// elementwise.cu — HBM-bound elementwise kernels: ReLU fwd/bwd, DropOut fwd/bwd (Philox),
// residual add, 3-way add, scale, fill.
//
// All are grid-stride loops over 128-bit vectors (scalar head-free tail only), grids sized to
// a whole number of waves of the 148 SMs, streaming loads/stores that bypass L1.
// Algorithmic bytes per element (fp32): relu fwd 8, relu bwd 12, dropout fwd/bwd 8 (mask is
// recomputed from the Philox counter, never stored), add_inplace 12, add3 12/16.
#include "common.cuh"

namespace npm {
namespace {

constexpr int kThreads = 256;

template <class F>
__global__ void __launch_bounds__(kThreads) ew1_kernel(const float* x, float* y,
                                                       int64_t n, F f) {
    const int64_t nv = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        float4 v = ld_stream(reinterpret_cast<const float4*>(x) + i);
        v.x = f(v.x); v.y = f(v.y); v.z = f(v.z); v.w = f(v.w);
        st_stream(reinterpret_cast<float4*>(y) + i, v);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (nv << 2) + threadIdx.x;
        y[i] = f(x[i]);
    }
}

template <class F>
__global__ void __launch_bounds__(kThreads) ew2_kernel(const float* a, const float* b, float* y, int64_t n, F f) {
    const int64_t nv = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 u = ld_stream(reinterpret_cast<const float4*>(a) + i);
        const float4 w = ld_stream(reinterpret_cast<const float4*>(b) + i);
        float4 v;
        v.x = f(u.x, w.x); v.y = f(u.y, w.y); v.z = f(u.z, w.z); v.w = f(u.w, w.w);
        st_stream(reinterpret_cast<float4*>(y) + i, v);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (nv << 2) + threadIdx.x;
        y[i] = f(a[i], b[i]);
    }
}

__global__ void __launch_bounds__(kThreads) add3_kernel(const float* a, const float* b, const float* c, float* y, int64_t n) {
    const int64_t nv = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 u = ld_stream(reinterpret_cast<const float4*>(a) + i);
        const float4 w = ld_stream(reinterpret_cast<const float4*>(b) + i);
        const float4 q = ld_stream(reinterpret_cast<const float4*>(c) + i);
        float4 v;
        v.x = (u.x + w.x) + q.x; v.y = (u.y + w.y) + q.y; v.z = (u.z + w.z) + q.z; v.w = (u.w + w.w) + q.w;
        st_stream(reinterpret_cast<float4*>(y) + i, v);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (nv << 2) + threadIdx.x;
        y[i] = (a[i] + b[i]) + c[i];
    }
}

// Scalar fallbacks for pointers that are not 16-byte aligned (views at odd offsets).
template <class F>
__global__ void ew1_scalar_kernel(const float* x, float* y, int64_t n, F f) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = f(x[i]);
}
template <class F>
__global__ void ew2_scalar_kernel(const float* a, const float* b, float* y, int64_t n, F f) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = f(a[i], b[i]);
}

template <class F>
int launch_ew1(const char* name, const float* x, float* y, int64_t n, F f, cudaStream_t s) {
    if (n <= 0) return NPM_OK;
    if (aligned16(x) && aligned16(y)) {
        ew1_kernel<<<bw_grid((n + 3) / 4, kThreads), kThreads, 0, s>>>(x, y, n, f);
    } else {
        ew1_scalar_kernel<<<bw_grid(n, kThreads), kThreads, 0, s>>>(x, y, n, f);
    }
    count_launch();
    return check_launch(name);
}
template <class F>
int launch_ew2(const char* name, const float* a, const float* b, float* y, int64_t n, F f, cudaStream_t s) {
    if (n <= 0) return NPM_OK;
    if (aligned16(a) && aligned16(b) && aligned16(y)) {
        ew2_kernel<<<bw_grid((n + 3) / 4, kThreads), kThreads, 0, s>>>(a, b, y, n, f);
    } else {
        ew2_scalar_kernel<<<bw_grid(n, kThreads), kThreads, 0, s>>>(a, b, y, n, f);
    }
    count_launch();
    return check_launch(name);
}

struct ReluF { __device__ float operator()(float x) const { return x < 0.0f ? 0.0f : x; } };
struct ReluB { __device__ float operator()(float x, float dy) const { return x >= 0.0f ? dy : 0.0f; } };
// ReLU.backward from the fused Dense output: the GEMM epilogue stored -0.0 for x < 0 and +0.0 for x == 0
struct ReluBY { __device__ float operator()(float y, float dy) const { return (__float_as_uint(y) >> 31) ? 0.0f : dy; } };
struct AddF  { __device__ float operator()(float a, float b) const { return a + b; } };
struct ScaleF { float s; __device__ float operator()(float x) const { return x * s; } };

// fp32 view of a tensor that exists as split-bf16 planes: out = hi + mid (exact in fp32: 16 significant bits; signed zeros kept)
__global__ void __launch_bounds__(kThreads) planes_join_kernel(const uint2* __restrict__ planes, int64_t plane4, float4* __restrict__ out,
                                                               int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 h = planes[i], m = planes[plane4 + i];
        // a zero keeps the sign of its hi part: the fused Dense epilogue stores -0.0 for a negative pre-activation
        // (activations.py:19 gate), and (-0.0) + (+0.0) would be +0.0
        auto join = [](uint32_t hw, uint32_t mw) {
            const float s = __uint_as_float(hw) + __uint_as_float(mw);
            return s == 0.0f ? __uint_as_float(hw & 0x80000000u) : s;
        };
        out[i] = make_float4(join(h.x << 16, m.x << 16), join(h.x & 0xffff0000u, m.x & 0xffff0000u),
                             join(h.y << 16, m.y << 16), join(h.y & 0xffff0000u, m.y & 0xffff0000u));
    }
}
__global__ void __launch_bounds__(kThreads) fill_kernel(float* __restrict__ x, float v, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = v;
}

// ---------------------------------------------------------------- dropout
// keep element g  <=>  philox(g >> 2)[g & 3] < thr, thr = floor(keep_prob * 2^32).
__device__ __forceinline__ uint32_t pick(const Philox4& p, int lane) {
    return lane == 0 ? p.x : lane == 1 ? p.y : lane == 2 ? p.z : p.w;
}

template <bool kExtMask, bool kWriteMask>
__global__ void __launch_bounds__(kThreads) dropout_kernel(const float* x, float* y,
                                                           uint8_t* __restrict__ mask_out, int64_t n,
                                                           float keep_prob, uint64_t thr, uint64_t seed,
                                                           uint64_t offset, const uint8_t* __restrict__ ext) {
    const int64_t ngroups = (n + 3) >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec_ok = aligned16(x) && aligned16(y);
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += stride) {
        const int64_t i0 = gi << 2;
        bool keep[4];
        if (kExtMask) {
#pragma unroll
            for (int e = 0; e < 4; ++e) keep[e] = (i0 + e < n) ? (ext[i0 + e] != 0) : false;
        } else {
            const uint64_t g0 = offset + (uint64_t)i0;
            const Philox4 p0 = philox4x32_10(g0 >> 2, seed);
            if ((g0 & 3) == 0) {
                keep[0] = p0.x < thr; keep[1] = p0.y < thr; keep[2] = p0.z < thr; keep[3] = p0.w < thr;
            } else {
                const Philox4 p1 = philox4x32_10((g0 >> 2) + 1, seed);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint64_t g = g0 + e;
                    const uint32_t r = ((g >> 2) == (g0 >> 2)) ? pick(p0, (int)(g & 3)) : pick(p1, (int)(g & 3));
                    keep[e] = r < thr;
                }
            }
        }
        if (kWriteMask) {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i0 + e < n) mask_out[i0 + e] = keep[e] ? 1 : 0;
        } else if (vec_ok && i0 + 3 < n) {
            float4 v = ld_stream(reinterpret_cast<const float4*>(x + i0));
            v.x = keep[0] ? __fdiv_rn(v.x, keep_prob) : 0.0f;
            v.y = keep[1] ? __fdiv_rn(v.y, keep_prob) : 0.0f;
            v.z = keep[2] ? __fdiv_rn(v.z, keep_prob) : 0.0f;
            v.w = keep[3] ? __fdiv_rn(v.w, keep_prob) : 0.0f;
            st_stream(reinterpret_cast<float4*>(y + i0), v);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i0 + e < n) y[i0 + e] = keep[e] ? __fdiv_rn(x[i0 + e], keep_prob) : 0.0f;
        }
    }
}

int dropout_apply(const float* x, float* y, int64_t n, float keep_prob, uint64_t seed, uint64_t offset,
                  const uint8_t* ext, cudaStream_t s) {
    if (n <= 0) return NPM_OK;
    NPM_REQUIRE(keep_prob > 0.0f && keep_prob <= 1.0f, "dropout: keep_prob %g out of (0,1]", keep_prob);
    const uint64_t thr = (uint64_t)((double)keep_prob * 4294967296.0);
    const int grid = bw_grid((n + 3) / 4, kThreads);
    if (ext)
        dropout_kernel<true, false><<<grid, kThreads, 0, s>>>(x, y, nullptr, n, keep_prob, thr, seed, offset, ext);
    else
        dropout_kernel<false, false><<<grid, kThreads, 0, s>>>(x, y, nullptr, n, keep_prob, thr, seed, offset, nullptr);
    count_launch();
    return check_launch("dropout_kernel");
}

}  // namespace
}  // namespace npm

using namespace npm;


namespace npm {
namespace {
// Token embedding (beyond the reference: the reference has no embedding layer).  One warp per token
// row: out[i, :] = table[ids[i], :] (+ pos[i % S, :]); the backward scatters dy rows into dtable with red.global.add
// (rows of the same token id collide, so the sum order — and the last bits — depend on scheduling).
__global__ void __launch_bounds__(256) embedding_fwd_kernel(const float* __restrict__ table, const int32_t* __restrict__ ids,
                                                            float* __restrict__ out, int64_t n, int d, int vocab) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        int id = ids[i];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        const float* src = table + (int64_t)id * d;
        float* dst = out + i * d;
        for (int c = lane; c < d; c += 32) dst[c] = __ldg(src + c);
    }
}
__global__ void __launch_bounds__(256) embedding_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ ids,
                                                            float* __restrict__ dtable, int64_t n, int d, int vocab) {
    const int lane = threadIdx.x & 31;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        int id = ids[i];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        float* dst = dtable + (int64_t)id * d;
        const float* src = dy + i * d;
        for (int c = lane; c < d; c += 32) atomicAdd(dst + c, src[c]);
    }
}
}  // namespace
}  // namespace npm

extern "C" {

int npm_relu_fwd(const float* x, float* y, int64_t n, npm_stream_t stream) {
    return launch_ew1("relu_fwd", x, y, n, ReluF{}, (cudaStream_t)stream);
}
int npm_relu_bwd_y(const float* y, const float* dy, float* dx, int64_t n, npm_stream_t stream) {
    return launch_ew2("relu_bwd_y", y, dy, dx, n, ReluBY{}, (cudaStream_t)stream);
}
int npm_relu_bwd(const float* x, const float* dy, float* dx, int64_t n, npm_stream_t stream) {
    return launch_ew2("relu_bwd", x, dy, dx, n, ReluB{}, (cudaStream_t)stream);
}
int npm_add_inplace(float* y, const float* x, int64_t n, npm_stream_t stream) {
    return launch_ew2("add_inplace", y, x, y, n, AddF{}, (cudaStream_t)stream);
}
int npm_add3(const float* a, const float* b, const float* c, float* out, int64_t n, npm_stream_t stream) {
    if (c == nullptr) return launch_ew2("add2", a, b, out, n, AddF{}, (cudaStream_t)stream);
    if (n <= 0) return NPM_OK;
    if (aligned16(a) && aligned16(b) && aligned16(c) && aligned16(out)) {
        add3_kernel<<<bw_grid((n + 3) / 4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(a, b, c, out, n);
        count_launch();
        return check_launch("add3_kernel");
    }
    int rc = launch_ew2("add3a", a, b, out, n, AddF{}, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_ew2("add3b", out, c, out, n, AddF{}, (cudaStream_t)stream);
}
int npm_scale(float* x, float s, int64_t n, npm_stream_t stream) {
    return launch_ew1("scale", x, x, n, ScaleF{s}, (cudaStream_t)stream);
}
int npm_embedding_fwd(const float* table, const int32_t* ids, float* out, int64_t n, int64_t d, int64_t vocab, npm_stream_t stream) {
    NPM_REQUIRE(table && ids && out && d > 0 && vocab > 0 && d < (1ll << 30) && vocab < (1ll << 31), "embedding_fwd: bad arguments");
    if (n <= 0) return NPM_OK;
    npm::embedding_fwd_kernel<<<bw_grid((size_t)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(table, ids, out, n, (int)d, (int)vocab);
    count_launch();
    return check_launch("embedding_fwd_kernel");
}
int npm_embedding_bwd(const float* dy, const int32_t* ids, float* dtable, int64_t n, int64_t d, int64_t vocab, npm_stream_t stream) {
    NPM_REQUIRE(dy && ids && dtable && d > 0 && vocab > 0 && d < (1ll << 30) && vocab < (1ll << 31), "embedding_bwd: bad arguments");
    cudaError_t e = cudaMemsetAsync(dtable, 0, sizeof(float) * (size_t)vocab * d, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("embedding_bwd memset: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
    if (n <= 0) return NPM_OK;
    npm::embedding_bwd_kernel<<<bw_grid((size_t)n * 32, 256), 256, 0, (cudaStream_t)stream>>>(dy, ids, dtable, n, (int)d, (int)vocab);
    count_launch();
    return check_launch("embedding_bwd_kernel");
}
int npm_fill(float* x, float v, int64_t n, npm_stream_t stream) {
    if (n <= 0) return NPM_OK;
    fill_kernel<<<bw_grid(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(x, v, n);
    count_launch();
    return check_launch("fill_kernel");
}

int npm_planes_join(const void* planes, int64_t plane, float* out, int64_t n, npm_stream_t stream) {
    if (n <= 0) return NPM_OK;
    NPM_REQUIRE(planes && out && (n & 3) == 0 && (plane & 3) == 0 && (reinterpret_cast<uintptr_t>(planes) & 7u) == 0 &&
                (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "planes_join: needs n %% 4 == 0, plane %% 4 == 0 and aligned pointers");
    planes_join_kernel<<<bw_grid(n / 4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint2*>(planes), plane / 4,
                                                                                       reinterpret_cast<float4*>(out), n / 4);
    count_launch();
    return check_launch("planes_join_kernel");
}

int npm_dropout_fwd(const float* x, float* y, int64_t n, float keep_prob, uint64_t seed, uint64_t offset,
                    const uint8_t* ext_mask, npm_stream_t stream) {
    return dropout_apply(x, y, n, keep_prob, seed, offset, ext_mask, (cudaStream_t)stream);
}
int npm_dropout_bwd(const float* dy, float* dx, int64_t n, float keep_prob, uint64_t seed, uint64_t offset,
                    const uint8_t* ext_mask, npm_stream_t stream) {
    // where(mask, dy / keep, 0) — the same map as forward (normalizations.py:25-30)
    return dropout_apply(dy, dx, n, keep_prob, seed, offset, ext_mask, (cudaStream_t)stream);
}
int npm_dropout_mask(uint8_t* mask, int64_t n, float keep_prob, uint64_t seed, uint64_t offset,
                     npm_stream_t stream) {
    if (n <= 0) return NPM_OK;
    NPM_REQUIRE(keep_prob > 0.0f && keep_prob <= 1.0f, "dropout: keep_prob %g out of (0,1]", keep_prob);
    const uint64_t thr = (uint64_t)((double)keep_prob * 4294967296.0);
    dropout_kernel<false, true><<<bw_grid((n + 3) / 4, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        nullptr, nullptr, mask, n, keep_prob, thr, seed, offset, nullptr);
    count_launch();
    return check_launch("dropout_mask_kernel");
}

}  // extern "C"
