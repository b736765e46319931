// loss_optim.cu — MSE / cross-entropy value + gradient (loss.py:20-39) and the fused
// multi-tensor SGD / Adam update (optimizer.py:30-69).
//
// The optimizer is ONE launch over a device-resident table of tensors: CTAs walk fixed-size
// chunks, find their tensor by binary search over the per-entry chunk prefix, and stream
// param/grad/m/v once.  Algorithmic bytes: SGD 12 B/param, Adam 28 B/param.
#include "common.cuh"

namespace npm {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double block_sum_d(double v) {
    __shared__ double sm[kThreads / 32];
    v = warp_sum_d(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < kThreads / 32) t = sm[threadIdx.x];
    if (threadIdx.x < 32) t = warp_sum_d(t);
    return t;   // valid in thread 0
}

// loss_out must be zero on entry (the wrappers memset it on the stream)
__global__ void __launch_bounds__(kThreads) mse_fwd_kernel(const float* __restrict__ y, const float* __restrict__ t,
                                                           float* loss_out, int64_t n, double inv_n) {
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float d = y[i] - t[i];
        acc += (double)(d * d);
    }
    const double tot = block_sum_d(acc);
    if (threadIdx.x == 0) atomicAdd(loss_out, (float)(tot * inv_n));
}
__global__ void __launch_bounds__(kThreads) ce_fwd_kernel(const float* __restrict__ y, const float* __restrict__ t,
                                                          float* loss_out, int64_t n) {
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc -= (double)(t[i] * logf(y[i]));
    const double tot = block_sum_d(acc);
    if (threadIdx.x == 0) atomicAdd(loss_out, (float)tot);
}
__global__ void __launch_bounds__(kThreads) mse_bwd_kernel(const float* __restrict__ y, const float* __restrict__ t,
                                                           float* __restrict__ dy, int64_t n, float nf) {
    pdl_trigger();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dy[i] = __fdiv_rn(2.0f * (y[i] - t[i]), nf);
}
__global__ void __launch_bounds__(kThreads) ce_bwd_kernel(const float* __restrict__ y, const float* __restrict__ t,
                                                          float* __restrict__ dy, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dy[i] = __fdiv_rn(-t[i], y[i]);
}

// ------------------------------------------------------------- optimizers
constexpr int64_t kChunk = NPM_OPT_CHUNK;

__device__ __forceinline__ int find_tensor(const npm_tensor_entry* tab, int n, int64_t chunk) {
    int lo = 0, hi = n - 1;   // last entry with chunk_begin <= chunk
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].chunk_begin <= chunk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ void split2(float p0, float p1, uint32_t& hi, uint32_t& mid) {     // bf16 hi / mid pairs of two values
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(p1), "f"(p0));
    const float r0 = p0 - __uint_as_float(hi << 16), r1 = p1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(r1), "f"(r0));
}

template <bool kAdam>
__global__ void __launch_bounds__(kThreads) opt_multi_kernel(const npm_tensor_entry* __restrict__ tab, int n_tensors,
                                                             int64_t n_chunks, float lr, float b1, float b2, float eps,
                                                             float bc1, float bc2, float gscale) {
    pdl_trigger();
    pdl_wait();
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int ti = find_tensor(tab, n_tensors, chunk);
        const npm_tensor_entry e = tab[ti];
        const int64_t begin = (chunk - e.chunk_begin) * kChunk;
        const int64_t end = begin + kChunk < e.numel ? begin + kChunk : e.numel;
        float* p = e.param;
        const float* g = e.grad;
        const bool vec = aligned16(p) && aligned16(g) && (!kAdam || (aligned16(e.m) && aligned16(e.v)));
        // planes ride on the vector path only (the host attaches them to 16-byte aligned tensors of 4k elements)
        uint2* planes = vec ? reinterpret_cast<uint2*>(e.planes) : nullptr;
        if (vec) {
            const int64_t nv = (end - begin) >> 2;
            for (int64_t i = threadIdx.x; i < nv; i += kThreads) {
                const int64_t o = begin + (i << 2);
                float4 pv = *reinterpret_cast<float4*>(p + o);
                float4 gv = ld_stream(reinterpret_cast<const float4*>(g + o));
                gv.x *= gscale; gv.y *= gscale; gv.z *= gscale; gv.w *= gscale;
                if (kAdam) {
                    float4 mv = *reinterpret_cast<float4*>(e.m + o);
                    float4 vv = *reinterpret_cast<float4*>(e.v + o);
#define NPM_ADAM1(c)                                                             \
    mv.c = b1 * mv.c + (1.0f - b1) * gv.c;                                       \
    vv.c = b2 * vv.c + (1.0f - b2) * (gv.c * gv.c);                              \
    pv.c -= lr * ((mv.c / bc1) / sqrtf(vv.c / bc2 + eps));
                    NPM_ADAM1(x) NPM_ADAM1(y) NPM_ADAM1(z) NPM_ADAM1(w)
                    *reinterpret_cast<float4*>(e.m + o) = mv;
                    *reinterpret_cast<float4*>(e.v + o) = vv;
                } else {
                    pv.x -= lr * gv.x; pv.y -= lr * gv.y; pv.z -= lr * gv.z; pv.w -= lr * gv.w;
                }
                *reinterpret_cast<float4*>(p + o) = pv;
                if (planes != nullptr) {        // the updated weight's split-bf16 image (npm_weight_split layout)
                    uint2 hi, mid;
                    split2(pv.x, pv.y, hi.x, mid.x);
                    split2(pv.z, pv.w, hi.y, mid.y);
                    planes[o >> 2] = hi;
                    planes[(e.plane_stride + o) >> 2] = mid;
                }
            }
        }
        const int64_t tail0 = vec ? begin + (((end - begin) >> 2) << 2) : begin;
        for (int64_t o = tail0 + threadIdx.x; o < end; o += kThreads) {
            const float gg = g[o] * gscale;
            if (kAdam) {
                const float m = b1 * e.m[o] + (1.0f - b1) * gg;
                const float v = b2 * e.v[o] + (1.0f - b2) * (gg * gg);
                e.m[o] = m; e.v[o] = v;
                p[o] -= lr * ((m / bc1) / sqrtf(v / bc2 + eps));
            } else {
                p[o] -= lr * gg;
            }
        }
    }
}

int loss_grid(int64_t n) { return bw_grid(n, kThreads, 4); }

}  // namespace
}  // namespace npm

using namespace npm;

extern "C" {

int npm_mse_fwd(const float* y, const float* t, float* loss_out, int64_t n, npm_stream_t stream) {
    NPM_REQUIRE(n > 0, "mse_fwd: empty input");
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(loss_out, 0, sizeof(float), s) != cudaSuccess) return check_launch("mse_fwd memset");
    mse_fwd_kernel<<<loss_grid(n), kThreads, 0, s>>>(y, t, loss_out, n, 1.0 / (double)n);
    count_launch();
    return check_launch("mse_fwd_kernel");
}
int npm_ce_fwd(const float* y, const float* t, float* loss_out, int64_t n, npm_stream_t stream) {
    NPM_REQUIRE(n > 0, "ce_fwd: empty input");
    cudaStream_t s = (cudaStream_t)stream;
    if (cudaMemsetAsync(loss_out, 0, sizeof(float), s) != cudaSuccess) return check_launch("ce_fwd memset");
    ce_fwd_kernel<<<loss_grid(n), kThreads, 0, s>>>(y, t, loss_out, n);
    count_launch();
    return check_launch("ce_fwd_kernel");
}
int npm_mse_bwd(const float* y, const float* t, float* dy, int64_t n, npm_stream_t stream) {
    NPM_REQUIRE(n > 0, "mse_bwd: empty input");
    launch_pdl(mse_bwd_kernel, dim3(bw_grid(n, kThreads)), dim3(kThreads), 0, (cudaStream_t)stream, 1, y, t, dy, n, (float)n);
    count_launch();
    return check_launch("mse_bwd_kernel");
}
int npm_ce_bwd(const float* y, const float* t, float* dy, int64_t n, npm_stream_t stream) {
    NPM_REQUIRE(n > 0, "ce_bwd: empty input");
    ce_bwd_kernel<<<bw_grid(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(y, t, dy, n);
    count_launch();
    return check_launch("ce_bwd_kernel");
}

int npm_sgd_multi(const npm_tensor_entry* table_dev, int32_t n_tensors, int64_t n_chunks, float lr, float grad_scale,
                  npm_stream_t stream) {
    if (n_tensors <= 0 || n_chunks <= 0) return NPM_OK;
    const int64_t cap = (int64_t)num_sms() * 8;
    const int grid = (int)(n_chunks < cap ? n_chunks : cap);
    opt_multi_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(table_dev, n_tensors, n_chunks, lr, 0.f, 0.f,
                                                                         0.f, 1.f, 1.f, grad_scale);
    count_launch();
    return check_launch("sgd_multi_kernel");
}

int npm_adam_multi(const npm_tensor_entry* table_dev, int32_t n_tensors, int64_t n_chunks, float lr, float beta1,
                   float beta2, float epsilon, int32_t t, float grad_scale, npm_stream_t stream) {
    if (n_tensors <= 0 || n_chunks <= 0) return NPM_OK;
    NPM_REQUIRE(t >= 1, "adam: step t must start at 1 (optimizer.py:53), got %d", t);
    // bias corrections in double on the host, as the reference's float64 arithmetic does
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)t));
    const float bc2 = (float)(1.0 - pow((double)beta2, (double)t));
    const int64_t cap = (int64_t)num_sms() * 8;
    const int grid = (int)(n_chunks < cap ? n_chunks : cap);
    launch_pdl(opt_multi_kernel<true>, dim3(grid), dim3(kThreads), 0, (cudaStream_t)stream, 1, table_dev, n_tensors, n_chunks, lr,
               beta1, beta2, epsilon, bc1, bc2, grad_scale);
    count_launch();
    return check_launch("adam_multi_kernel");
}

}  // extern "C"
