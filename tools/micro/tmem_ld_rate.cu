// tools/micro/tmem_ld_rate.cu — cycles per tcgen05.ld / tcgen05.st (32x32b.x32 = 4 KB per warp instruction) with 1, 4, 8
// warps issuing: is TMEM read bandwidth per SM or per SM sub-partition?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../np-modeling_b200/csrc/ptx.cuh"
using namespace npm;

template <int MODE>   // 0 = ld with a wait after every load, 1 = 4 loads per wait, 2 = st (4 per wait)
__global__ void __launch_bounds__(256, 1) k(long long* out, int iters, int active_warps) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = slot + (uint32_t((warp & 3) * 32) << 16);
    uint32_t v[32], acc = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i;
    __syncthreads();
    long long t0 = clock64();
    if (warp < active_warps) {
        for (int it = 0; it < iters; ++it) {
            if (MODE == 0) {
                ptx::tmem_ld_32x32(tm + (it & 7) * 32, v); ptx::tmem_ld_wait(); acc += v[0];
            } else if (MODE == 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c) { ptx::tmem_ld_32x32(tm + ((it * 4 + c) & 15) * 32, v); }
                ptx::tmem_ld_wait(); acc += v[0];
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) ptx::tmem_st_32x32(tm + ((it * 4 + c) & 15) * 32, v);
                ptx::tmem_st_wait();
            }
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[warp] = (t1 - t0) + (acc == 12345 ? 1 : 0);
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(slot, 512);
}
template <int MODE> void run(const char* name, int warps) {
    long long* d; cudaMalloc(&d, 64);
    const int iters = 2000;
    k<MODE><<<1, 256>>>(d, iters, warps); cudaDeviceSynchronize();
    k<MODE><<<1, 256>>>(d, iters, warps); cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    const int per = MODE == 0 ? 1 : 4;
    printf("%-34s %d warps: %7.1f cycles per 4 KB warp-instruction  -> %6.1f B/clk/SM   (%s)\n", name, warps,
           (double)h[0] / (iters * per), 4096.0 * per * iters * warps / (double)h[0], cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    for (int w : {1, 4, 8}) run<0>("tcgen05.ld x32, wait each", w);
    for (int w : {1, 4, 8}) run<1>("tcgen05.ld x32, 4 per wait", w);
    for (int w : {1, 4, 8}) run<2>("tcgen05.st x32, 4 per wait", w);
    return 0;
}
