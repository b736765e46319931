// tools/micro/mma_rate.cu — cycles per tcgen05.mma kind::tf32 for the operand sources / shapes the attention
// kernels use (garbage data; only the timing matters).  nvcc -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../np-modeling_b200/csrc/ptx.cuh"
using namespace npm;

template <int MODE>   // 0: SS N=128 K-major both; 1: TS N=64 B MN-major; 2: SS N=256; 3: SS N=64; 4: TS N=128 (B MN-major 128)
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ uint64_t barmem;
    const uint32_t bar = ptx::smem_u32(&barmem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { if (lane == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); } __syncwarp(); ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint64_t dk = ptx::umma_desc_base(2, 16, 1024), dmn = ptx::umma_desc_base(1, 16384, 512);
        constexpr int N = (MODE == 0 || MODE == 4) ? 128 : (MODE == 2 ? 256 : 64);
        const uint32_t idesc = ptx::umma_idesc_tf32(128, N, false, MODE == 1 || MODE == 4);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                if (MODE == 1 || MODE == 4)
                    ptx::umma_tf32_ts(tm + 256, tm + (kk & 15) * 8, ptx::umma_desc(dmn, base + 65536 + kk * 1024), idesc, 1u);
                else
                    ptx::umma_tf32(tm + 256, ptx::umma_desc(dk, base + (kk & 3) * 32 + (kk >> 2) * 16384),
                                   ptx::umma_desc(dk, base + 65536 + (kk & 3) * 32 + (kk >> 2) * 16384), idesc, 1u);
            }
        }
        ptx::umma_commit(bar);
        ptx::mbar_wait(bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int MODE> void run(const char* name, int grid) {
    long long* d; cudaMalloc(&d, sizeof(long long) * grid);
    const int smem = 200 * 1024, iters = 200;
    cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<MODE><<<grid, 128, smem>>>(d, iters);
    cudaDeviceSynchronize();
    rate_kernel<MODE><<<grid, 128, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(long long) * (grid < 148 ? grid : 148), cudaMemcpyDeviceToHost);
    printf("%-44s grid %3d: %7.1f cycles / MMA   (%s)\n", name, grid, (double)h[0] / (iters * 16), cudaGetErrorString(e));
    cudaFree(d);
}
// cta_group::2 rates: MODE 0: SS M256 N128 K8 (B = 64 rows per CTA); 1: TS M256 N64 K8 (A from TMEM, B = one 32-wide MN slab per CTA)
template <int MODE>
__global__ void __launch_bounds__(128, 1) rate2_kernel(long long* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint32_t slot;
    __shared__ uint64_t barmem;
    const uint32_t bar = ptx::smem_u32(&barmem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    if (warp == 0) { if (lane == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); } __syncwarp(); ptx::tmem_alloc_2sm(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish_2sm(); }
    ptx::tc_fence_before(); ptx::cluster_sync(); ptx::tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 1 && rank == 0 && ptx::elect_one()) {
        const uint64_t dk = ptx::umma_desc_base(2, 16, 1024), dmn = ptx::umma_desc_base(1, 16384, 512);
        constexpr int N = MODE == 0 ? 128 : 64;
        const uint32_t idesc = ptx::umma_idesc_tf32(256, N, false, MODE == 1);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                if (MODE == 1)
                    ptx::umma_tf32_ts_2sm(tm + 256, tm + (kk & 15) * 8, ptx::umma_desc(dmn, base + 65536 + kk * 1024), idesc, 1u);
                else
                    ptx::umma_tf32_2sm(tm + 256, ptx::umma_desc(dk, base + (kk & 3) * 32 + (kk >> 2) * 16384),
                                       ptx::umma_desc(dk, base + 65536 + (kk & 3) * 32 + (kk >> 2) * 8192), idesc, 1u);
            }
        }
        ptx::umma_commit_2sm(bar, 1);
        ptx::mbar_wait(bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    ptx::tc_fence_before(); ptx::cluster_sync();
    if (warp == 0) ptx::tmem_dealloc_2sm(tm, 512);
}
template <int MODE> void run2(const char* name, int grid) {
    long long* d; cudaMalloc(&d, sizeof(long long) * grid);
    const int smem = 200 * 1024, iters = 200;
    auto kern = rate2_kernel<MODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 2; ++rep) { cudaLaunchKernelEx(&cfg, kern, d, iters); e = cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, d, sizeof(long long) * (grid < 148 ? grid : 148), cudaMemcpyDeviceToHost);
    printf("%-44s grid %3d: %7.1f cycles / MMA   (%s)\n", name, grid, (double)h[0] / (iters * 16), cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    for (int grid : {2, 148}) {
        run2<0>("2SM SS M256 N128 K8", grid);
        run2<1>("2SM TS M256 N64  K8 (A from TMEM)", grid);
    }
    for (int grid : {1, 148}) {
        run<0>("SS  M128 N128 K8 (A,B K-major from smem)", grid);
        run<2>("SS  M128 N256 K8", grid);
        run<3>("SS  M128 N64  K8", grid);
        run<1>("TS  M128 N64  K8 (A from TMEM, B MN-major)", grid);
        run<4>("TS  M128 N128 K8 (A from TMEM, B MN-major)", grid);
    }
    return 0;
}
