// gemm_simt.cu — CUDA-core fp32 GEMM with arbitrary strides.
//
// Serves the shapes TMA cannot describe (a leading dimension that is not a multiple of
// 16 bytes, e.g. the reference's Dense(10) head, layers/mlp.py:18 with N=10) and the
// NPM_PREC_FP32 mode.  Same contract as gemm_tc.cu; 64x64 tile, 16-deep K slices,
// 256 threads, 4x4 register micro-tile, fp32 FMA accumulation in ascending-k order.
#include "common.cuh"

namespace npm {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
    const float* a; const float* b; float* c; const float* bias;
    int64_t m, n, k;
    int64_t a_rs, a_cs, b_rs, b_cs, ldc;
    int nb1;
    int64_t a_bs1, a_bs2, b_bs1, b_bs2, c_bs1, c_bs2;
    float alpha;
    int relu, accum;
    const float* residual; int64_t ldr;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtArgs p) {
    __shared__ float sA[TK][TM + 4];
    __shared__ float sB[TK][TN + 4];

    const int z = blockIdx.z, z1 = z % p.nb1, z2 = z / p.nb1;
    const float* A = p.a + z1 * p.a_bs1 + z2 * p.a_bs2;
    const float* B = p.b + z1 * p.b_bs1 + z2 * p.b_bs2;
    float* C = p.c + z1 * p.c_bs1 + z2 * p.c_bs2;

    const int64_t m0 = (int64_t)blockIdx.x * TM, n0 = (int64_t)blockIdx.y * TN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, each 4x4 outputs

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    // Loader index maps: make the contiguous global index the fastest thread index.
    const bool a_k_contig = (p.a_cs == 1);
    const bool b_n_contig = (p.b_cs == 1);

    for (int64_t k0 = 0; k0 < p.k; k0 += TK) {
#pragma unroll
        for (int i = 0; i < (TM * TK) / 256; ++i) {
            const int e = tid + i * 256;
            int mm, kk;
            if (a_k_contig) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
            const int64_t gm = m0 + mm, gk = k0 + kk;
            sA[kk][mm] = (gm < p.m && gk < p.k) ? __ldg(A + gm * p.a_rs + gk * p.a_cs) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < (TN * TK) / 256; ++i) {
            const int e = tid + i * 256;
            int nn, kk;
            if (b_n_contig) { nn = e % TN; kk = e / TN; } else { kk = e % TK; nn = e / TK; }
            const int64_t gn = n0 + nn, gk = k0 + kk;
            sB[kk][nn] = (gn < p.n && gk < p.k) ? __ldg(B + gk * p.b_rs + gn * p.b_cs) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + ty * 4 + i;
        if (gm >= p.m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gn = n0 + tx * 4 + j;
            if (gn >= p.n) continue;
            float v = acc[i][j] * p.alpha;
            if (p.bias) v += __ldg(p.bias + gn);
            if (p.residual) v += __ldg(p.residual + gm * p.ldr + gn);
            if (p.relu) v = v < 0.0f ? -0.0f : fabsf(v);     // sign bit set <=> pre-activation < 0 (a genuine -0.0 is >= 0)
            float* dst = C + gm * p.ldc + gn;
            *dst = p.accum ? (*dst + v) : v;
        }
    }
}

}  // namespace

int gemm_simt_launch(const npm_gemm_desc& d, cudaStream_t stream) {
    SimtArgs p;
    p.a = d.a; p.b = d.b; p.c = d.c; p.bias = d.bias;
    p.m = d.m; p.n = d.n; p.k = d.k;
    p.a_rs = d.a_rs; p.a_cs = d.a_cs; p.b_rs = d.b_rs; p.b_cs = d.b_cs; p.ldc = d.ldc;
    p.nb1 = d.nb1 > 0 ? d.nb1 : 1;
    const int nb2 = d.nb2 > 0 ? d.nb2 : 1;
    p.a_bs1 = d.a_bs1; p.a_bs2 = d.a_bs2; p.b_bs1 = d.b_bs1; p.b_bs2 = d.b_bs2;
    p.c_bs1 = d.c_bs1; p.c_bs2 = d.c_bs2;
    p.alpha = d.alpha;
    p.relu = (d.flags & NPM_GEMM_RELU) ? 1 : 0;
    p.accum = (d.flags & NPM_GEMM_ACCUM) ? 1 : 0;
    p.residual = d.residual; p.ldr = d.ldr;
    const int64_t gx = (d.n + TN - 1) / TN, gy = (d.m + TM - 1) / TM, gz = (int64_t)p.nb1 * nb2;
    if (gx > 65535 || gz > 65535) {
        set_error("gemm_simt: grid too large (n tiles %lld, batches %lld)", (long long)gx, (long long)gz);
        return NPM_ERR_UNSUPPORTED;
    }
    gemm_simt_kernel<<<dim3((unsigned)gy, (unsigned)gx, (unsigned)gz), 256, 0, stream>>>(p);
    count_launch();
    return check_launch("gemm_simt_kernel");
}

}  // namespace npm
