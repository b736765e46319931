// conv.cu — Conv2D forward / input-gradient / filter-gradient (layers/conv.py:44-194) as
// implicit GEMMs: NHWC activations, HWIO filters, SAME padding, stride 1, odd kernel size.
//
//   fprop : y[p, o]      = sum_{tap,c} xpad[p + tap, c] * f[tap, c, o]      M = N*H*W, N = Cout, K = k*k*Cin
//   dgrad : dx[p, c]     = sum_{tap,o} dypad[p + tap, o] * f[flip(tap), c, o]   (conv.py:110-130)
//   wgrad : dw[tap,c,o]  = sum_p      xpad[p + tap, c] * dy[p, o]          M = k*k*Cin, N = Cout, K = N*H*W
//
// The im2col matrix is never materialised: the loaders compute shifted addresses on the fly and
// return 0 in the halo.  This file holds the CUDA-core fp32 implementation (64x64 tile, 16-deep
// K slices, 4x4 register micro-tile) that serves every shape, including channel counts TMA cannot
// describe (Cin = 3: 12-byte pixel stride); wgrad is split over pixel slabs and combined with
// fp32 atomics into a zero-initialised dw.
#include <stdlib.h>

#include "common.cuh"

namespace npm {
int current_precision();
// conv_tc.cu: tcgen05 implicit GEMM (TF32 mode)
bool conv_tc_supported(const void* p0, const void* p1, const void* p2, int64_t N, int64_t H, int64_t W, int64_t Cin,
                       int64_t Cout, int ks);
int conv_tc_fprop_dgrad(bool dgrad, const float* act, const float* f, const float* bias, float* out, int64_t N, int64_t H,
                        int64_t W, int64_t Cin, int64_t Cout, int ks, int relu, bool bx_mode, cudaStream_t stream);
int conv_tc_wgrad(const float* x, const float* dy, float* dw, int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                  int ks, bool bx_mode, cudaStream_t stream);
static bool bx_mode() { return current_precision() == NPM_PREC_BF16X3; }
static bool use_tc(const void* p0, const void* p1, const void* p2, int64_t N, int64_t H, int64_t W, int64_t Cin,
                   int64_t Cout, int ks) {
    // tensor cores in the single-pass TF32 mode and in the split-bf16 mode (bf16 mode: the TF32 kernels)
    const int prec = current_precision();
    static const bool simt_only = getenv("NPM_CONV_SIMT") != nullptr;      // A/B switch for tools/
    return (prec == NPM_PREC_TF32 || prec == NPM_PREC_BF16X3 || prec == NPM_PREC_BF16) && !simt_only &&
           conv_tc_supported(p0, p1, p2, N, H, W, Cin, Cout, ks);
}
size_t colsum_workspace_bytes(int64_t rows, int64_t cols);
int colsum_launch(const float* x, float* out, int64_t rows, int64_t cols, void* workspace, cudaStream_t s);

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct ConvArgs {
    const float* act;    // x (fprop, wgrad) or dy (dgrad)            [N,H,W,Ca]
    const float* other;  // filters (fprop, dgrad) or dy (wgrad)
    float* out;
    const float* bias;
    int N, H, W, Ca;     // Ca = channels of `act`
    int Co;              // channels of the output of this GEMM's N dimension
    int ks, pad;
    int relu;
    int64_t pixels;
    int64_t k_per_slab;  // wgrad: pixels per split-K slab
};

enum Mode { FPROP = 0, DGRAD = 1, WGRAD = 2 };

// shifted activation fetch: pixel p (linear n,h,w), tap t, channel c
__device__ __forceinline__ float act_at(const ConvArgs& a, int64_t p, int tap, int c) {
    const int w = (int)(p % a.W);
    const int64_t r = p / a.W;
    const int h = (int)(r % a.H);
    const int64_t n = r / a.H;
    const int hh = h + tap / a.ks - a.pad, ww = w + tap % a.ks - a.pad;
    if (hh < 0 || hh >= a.H || ww < 0 || ww >= a.W) return 0.0f;
    return __ldg(a.act + ((n * a.H + hh) * a.W + ww) * a.Ca + c);
}

template <int MODE>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvArgs a) {
    __shared__ float sA[TK][TM + 4];
    __shared__ float sB[TK][TN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int taps = a.ks * a.ks;

    // GEMM extents
    const int64_t M = MODE == WGRAD ? (int64_t)taps * a.Ca : a.pixels;
    const int64_t Nn = a.Co;
    const int64_t Kfull = MODE == WGRAD ? a.pixels : (int64_t)taps * a.Ca;
    const int64_t m0 = (int64_t)blockIdx.x * TM, n0 = (int64_t)blockIdx.y * TN;
    int64_t kbeg = 0, kend = Kfull;
    if (MODE == WGRAD) {
        kbeg = (int64_t)blockIdx.z * a.k_per_slab;
        kend = kbeg + a.k_per_slab < Kfull ? kbeg + a.k_per_slab : Kfull;
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
        // ---- A tile [TM x TK]
#pragma unroll
        for (int i = 0; i < (TM * TK) / 256; ++i) {
            const int e = tid + i * 256;
            int mm, kk;
            if (MODE == WGRAD) { mm = e % TM; kk = e / TM; }   // m = (tap, c): c contiguous in memory
            else               { kk = e % TK; mm = e / TK; }   // k = (tap, c): c contiguous in memory
            const int64_t gm = m0 + mm, gk = k0 + kk;
            float v = 0.0f;
            if (gm < M && gk < kend) {
                if (MODE == WGRAD) v = act_at(a, gk, (int)(gm / a.Ca), (int)(gm % a.Ca));
                else               v = act_at(a, gm, (int)(gk / a.Ca), (int)(gk % a.Ca));
            }
            sA[kk][mm] = v;
        }
        // ---- B tile [TK x TN]
#pragma unroll
        for (int i = 0; i < (TN * TK) / 256; ++i) {
            const int e = tid + i * 256;
            int nn, kk;
            if (MODE == DGRAD) { kk = e % TK; nn = e / TK; }   // k = (tap, o): o contiguous
            else               { nn = e % TN; kk = e / TN; }
            const int64_t gn = n0 + nn, gk = k0 + kk;
            float v = 0.0f;
            if (gn < Nn && gk < kend) {
                if (MODE == FPROP) {
                    v = __ldg(a.other + gk * a.Co + gn);                        // f[tap, c, o]
                } else if (MODE == DGRAD) {
                    const int tap = (int)(gk / a.Ca), o = (int)(gk % a.Ca);     // Ca = Cout here
                    v = __ldg(a.other + ((int64_t)(taps - 1 - tap) * a.Co + gn) * a.Ca + o);  // f[flip, c=gn, o]
                } else {
                    v = __ldg(a.other + gk * a.Co + gn);                        // dy[p, o]
                }
            }
            sB[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gn = n0 + tx * 4 + j;
            if (gn >= Nn) continue;
            float v = acc[i][j];
            float* dst = a.out + gm * a.Co + gn;
            if (MODE == WGRAD) {
                atomicAdd(dst, v);
            } else {
                if (a.bias) v += __ldg(a.bias + gn);
                if (a.relu) v = fmaxf(v, 0.0f);
                *dst = v;
            }
        }
    }
}

int check_conv(int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int ks) {
    NPM_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv2d: empty tensor");
    NPM_REQUIRE(ks > 0 && (ks & 1), "conv2d: kernel size must be odd (conv.py:94), got %d", ks);
    NPM_REQUIRE(N * H * W < (1ll << 40) && Cin < (1 << 20) && Cout < (1 << 20), "conv2d: tensor too large");
    return NPM_OK;
}

}  // namespace

// dgrad for a layer with very few INPUT channels (the image layer: Cin = 3): the GEMM's N dimension is Cin, so the
// 64-wide tile above wastes 95 % of its work.  Here one warp owns an output pixel: each lane takes two of every 64 dy
// channels, multiplies them with the (flipped) filters held in shared memory for all CIN outputs and the warp reduces
// the CIN sums with shuffles; taps in the halo are skipped warp-uniformly.            conv.py:110-153
template <int CIN>
__global__ void __launch_bounds__(256) conv_dgrad_smallc_kernel(const float* __restrict__ dy, const float* __restrict__ f,
                                                               float* __restrict__ dx, int N, int H, int W, int Cout,
                                                               int ks, int64_t pixels) {
    extern __shared__ float sf[];                 // [taps][CIN][Cout], tap already flipped
    const int taps = ks * ks, pad = ks / 2;
    for (int e = threadIdx.x; e < taps * CIN * Cout; e += blockDim.x) {
        const int o = e % Cout, c = (e / Cout) % CIN, tap = e / (Cout * CIN);
        sf[e] = __ldg(f + ((int64_t)(taps - 1 - tap) * CIN + c) * Cout + o);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npix = (int)pixels;                 // the launcher guarantees pixels < 2^31: 32-bit index arithmetic
    for (int p = blockIdx.x * 8 + warp; p < npix; p += gridDim.x * 8) {
        const int w = p % W;
        const int r = p / W;
        const int h = r % H;
        const int n = r / H;
        float acc[CIN];
#pragma unroll
        for (int c = 0; c < CIN; ++c) acc[c] = 0.0f;
        if (Cout == 64 && ks == 3) {
            // the image layer of cfg2: all nine taps' loads are issued before the FMAs (halo taps read pixel p, weight 0)
            float2 v[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
                const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
                const float2 t = __ldg(reinterpret_cast<const float2*>(dy + ((size_t)(n * H + (ok ? hh : h)) * W + (ok ? ww : w)) * 64 + lane * 2));
                v[tap] = ok ? t : make_float2(0.0f, 0.0f);
            }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                for (int c = 0; c < CIN; ++c) {
                    const float2 g = *reinterpret_cast<const float2*>(sf + (tap * CIN + c) * 64 + lane * 2);
                    acc[c] = fmaf(v[tap].x, g.x, fmaf(v[tap].y, g.y, acc[c]));
                }
        } else {
            for (int tap = 0; tap < taps; ++tap) {
                const int hh = h + tap / ks - pad, ww = w + tap % ks - pad;
                if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
                const float* row = dy + ((size_t)(n * H + hh) * W + ww) * Cout;
                const float* ft = sf + tap * CIN * Cout;
                for (int o = lane * 2; o < Cout; o += 64) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(row + o));
#pragma unroll
                    for (int c = 0; c < CIN; ++c) {
                        const float2 g = *reinterpret_cast<const float2*>(ft + c * Cout + o);
                        acc[c] = fmaf(v.x, g.x, fmaf(v.y, g.y, acc[c]));
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CIN; ++c)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < CIN; ++c) dx[(size_t)p * CIN + c] = acc[c];
        }
    }
}

template <int CIN>
int launch_dgrad_smallc(const float* dy, const float* f, float* dx, int64_t N, int64_t H, int64_t W, int64_t Cout, int ks,
                        cudaStream_t s) {
    const size_t smem = (size_t)ks * ks * CIN * Cout * sizeof(float);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_dgrad_smallc_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv_dgrad_smallc smem attribute: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
        configured = smem;
    }
    const int64_t pixels = N * H * W;
    int64_t grid = (pixels + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    conv_dgrad_smallc_kernel<CIN><<<(unsigned)grid, 256, smem, s>>>(dy, f, dx, (int)N, (int)H, (int)W, (int)Cout, ks, pixels);
    count_launch();
    return check_launch("conv_dgrad_smallc_kernel");
}
}  // namespace npm

using namespace npm;

extern "C" {

size_t npm_conv2d_workspace(int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int ks) {
    (void)Cin; (void)ks;
    if (N <= 0 || H <= 0 || W <= 0 || Cout <= 0) return 0;
    return colsum_workspace_bytes(N * H * W, Cout);   // bias-gradient reduction
}

int npm_conv2d_fwd(const float* x, const float* f, const float* b, float* y, int64_t N, int64_t H, int64_t W,
                   int64_t Cin, int64_t Cout, int ksize, int relu, void* workspace, npm_stream_t stream) {
    (void)workspace;
    int rc = check_conv(N, H, W, Cin, Cout, ksize);
    if (rc) return rc;
    NPM_REQUIRE(x && f && y, "conv2d_fwd: NULL pointer");
    if (use_tc(x, f, y, N, H, W, Cin, Cout, ksize))
        return conv_tc_fprop_dgrad(false, x, f, b, y, N, H, W, Cin, Cout, ksize, relu, bx_mode(), (cudaStream_t)stream);
    ConvArgs a{};
    a.act = x; a.other = f; a.out = y; a.bias = b;
    a.N = (int)N; a.H = (int)H; a.W = (int)W; a.Ca = (int)Cin; a.Co = (int)Cout;
    a.ks = ksize; a.pad = ksize / 2; a.relu = relu; a.pixels = N * H * W;
    dim3 grid((unsigned)((a.pixels + TM - 1) / TM), (unsigned)((Cout + TN - 1) / TN), 1);
    conv_simt_kernel<FPROP><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_launch("conv_fprop_kernel");
}

int npm_conv2d_bwd_dx(const float* dy, const float* f, float* dx, int64_t N, int64_t H, int64_t W, int64_t Cin,
                      int64_t Cout, int ksize, void* workspace, npm_stream_t stream) {
    (void)workspace;
    int rc = check_conv(N, H, W, Cin, Cout, ksize);
    if (rc) return rc;
    NPM_REQUIRE(dy && f && dx, "conv2d_bwd_dx: NULL pointer");
    if (use_tc(dy, f, dx, N, H, W, Cin, Cout, ksize))
        return conv_tc_fprop_dgrad(true, dy, f, nullptr, dx, N, H, W, Cin, Cout, ksize, 0, bx_mode(), (cudaStream_t)stream);
    if (Cin <= 4 && Cout % 2 == 0 && (reinterpret_cast<uintptr_t>(dy) & 7u) == 0 && N * H * W < (1ll << 31) &&
        (size_t)ksize * ksize * Cin * Cout * sizeof(float) <= 160 * 1024) {
        cudaStream_t s = (cudaStream_t)stream;
        switch (Cin) {
            case 1: return launch_dgrad_smallc<1>(dy, f, dx, N, H, W, Cout, ksize, s);
            case 2: return launch_dgrad_smallc<2>(dy, f, dx, N, H, W, Cout, ksize, s);
            case 3: return launch_dgrad_smallc<3>(dy, f, dx, N, H, W, Cout, ksize, s);
            default: return launch_dgrad_smallc<4>(dy, f, dx, N, H, W, Cout, ksize, s);
        }
    }
    ConvArgs a{};
    a.act = dy; a.other = f; a.out = dx; a.bias = nullptr;
    a.N = (int)N; a.H = (int)H; a.W = (int)W; a.Ca = (int)Cout; a.Co = (int)Cin;
    a.ks = ksize; a.pad = ksize / 2; a.relu = 0; a.pixels = N * H * W;
    dim3 grid((unsigned)((a.pixels + TM - 1) / TM), (unsigned)((Cin + TN - 1) / TN), 1);
    conv_simt_kernel<DGRAD><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    count_launch();
    return check_launch("conv_dgrad_kernel");
}

int npm_conv2d_bwd_dw_db(const float* x, const float* dy, float* dw, float* db, int64_t N, int64_t H, int64_t W,
                         int64_t Cin, int64_t Cout, int ksize, void* workspace, npm_stream_t stream) {
    int rc = check_conv(N, H, W, Cin, Cout, ksize);
    if (rc) return rc;
    NPM_REQUIRE(x && dy && dw, "conv2d_bwd_dw_db: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t pixels = N * H * W;
    const int64_t Mw = (int64_t)ksize * ksize * Cin;
    if ((rc = npm_fill(dw, 0.0f, Mw * Cout, stream))) return rc;
    if (use_tc(x, dy, dw, N, H, W, Cin, Cout, ksize)) {
        if ((rc = conv_tc_wgrad(x, dy, dw, N, H, W, Cin, Cout, ksize, bx_mode(), s))) return rc;
        if (db == nullptr) return NPM_OK;
        NPM_REQUIRE(workspace != nullptr, "conv2d_bwd_dw_db: workspace is NULL");
        return colsum_launch(dy, db, pixels, Cout, workspace, s);
    }
    ConvArgs a{};
    a.act = x; a.other = dy; a.out = dw; a.bias = nullptr;
    a.N = (int)N; a.H = (int)H; a.W = (int)W; a.Ca = (int)Cin; a.Co = (int)Cout;
    a.ks = ksize; a.pad = ksize / 2; a.relu = 0; a.pixels = pixels;
    const int64_t tiles = ((Mw + TM - 1) / TM) * ((Cout + TN - 1) / TN);
    int64_t slabs = ((int64_t)num_sms() * 8 + tiles - 1) / tiles;
    const int64_t max_slabs = (pixels + 255) / 256;
    if (slabs > max_slabs) slabs = max_slabs;
    if (slabs > 65535) slabs = 65535;
    if (slabs < 1) slabs = 1;
    a.k_per_slab = ((pixels + slabs - 1) / slabs + TK - 1) / TK * TK;
    slabs = (pixels + a.k_per_slab - 1) / a.k_per_slab;
    dim3 grid((unsigned)((Mw + TM - 1) / TM), (unsigned)((Cout + TN - 1) / TN), (unsigned)slabs);
    conv_simt_kernel<WGRAD><<<grid, 256, 0, s>>>(a);
    count_launch();
    if ((rc = check_launch("conv_wgrad_kernel"))) return rc;
    if (db == nullptr) return NPM_OK;
    NPM_REQUIRE(workspace != nullptr, "conv2d_bwd_dw_db: workspace is NULL");
    return colsum_launch(dy, db, pixels, Cout, workspace, s);
}

}  // extern "C"
