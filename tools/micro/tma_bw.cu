// tools/micro/tma_bw.cu — L2 -> shared-memory delivery rate of TMA tile loads, unicast vs cluster multicast.
// Question it answers: the fp32 GEMM / attention kernels sit on the L2->SM ceiling; does .multicast::cluster
// (each CTA of a cluster loads 1/CS of a tile and broadcasts it) raise the bytes a CTA RECEIVES per clock?
// Every CTA streams 16 KB boxes {32 fp32, 128 rows} of an L2-resident [16384, 1024] matrix through a smem ring.
//   CS = cluster size (1 = unicast), G = CTAs that read the same box (sharing group, as GEMM CTAs sharing A).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../np-modeling_b200/csrc/ptx.cuh"
using namespace npm;

constexpr int kBoxBytes = 16384;
constexpr int kRowsM = 16384, kColsK = 1024;

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

template <int CS, int STAGES, int NBOX, int PMODE>
__global__ void __launch_bounds__(256, 1) bw_kernel(const __grid_constant__ CUtensorMap tm, int iters, int share, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + STAGES * kBoxBytes;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
    const uint32_t rank = CS > 1 ? ptx::cluster_ctarank() : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), CS); }
        ptx::fence_mbar_init();
    }
    if (CS > 1) ptx::cluster_sync(); else __syncthreads();
    const int group = (int)blockIdx.x / share;             // CTAs of one sharing group stream the same boxes
    const int mblocks = kRowsM / 128, ksteps = kColsK / 32;
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    // issuing thread j handles boxes j, j+NISS, ...
    constexpr int NISS = PMODE == 0 ? 1 : NBOX;
    const int my_j = PMODE == 1 ? (ln == 0 && w >= 2 && w < 2 + NBOX ? w - 2 : -1) : (PMODE == 2 ? (w == 2 && ln < NBOX ? ln : -1) : (threadIdx.x == 64 ? 0 : -1));
    if (my_j >= 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            ptx::mbar_wait(empty(s), ph ^ 1);
            if (my_j == 0) ptx::mbar_arrive_expect_tx(full(s), kBoxBytes);
            const int mb = (group * 7 + it / ksteps) % mblocks, ks = it % ksteps;
            if (CS == 1) {
#pragma unroll
                for (int bx = my_j; bx < NBOX; bx += NISS)
                    tma_load_2d(base + s * kBoxBytes + bx * (kBoxBytes / NBOX), &tm, full(s), ks * 32, mb * 128 + bx * (128 / NBOX));
            } else {
                constexpr int rows = 128 / CS;
                tma_load_2d_mc(base + s * kBoxBytes + rank * (kBoxBytes / CS), &tm, full(s), ks * 32, mb * 128 + rank * rows,
                               (uint16_t)((1u << CS) - 1));
            }
        }
        if (my_j == 0) out[blockIdx.x] = clock64() - t0;
    } else if (threadIdx.x == 32) {
        for (int it = 0; it < iters; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            ptx::mbar_wait(full(s), ph);
            if (CS == 1) ptx::mbar_arrive(empty(s));
            else
                for (int r = 0; r < CS; ++r) ptx::mbar_arrive_cluster(ptx::mapa(empty(s), r));
        }
    }
    if (CS > 1) ptx::cluster_sync(); else __syncthreads();
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CS, int STAGES, int NBOX = 1, int PMODE = 0>
void run(EncodeFn enc, float* d, int share, int grid, int iters, bool tf32 = false, bool atom32 = false) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {kColsK, kRowsM};
    cuuint64_t strides[1] = {kColsK * 4};
    cuuint32_t box[2] = {32, 128 / CS / NBOX};
    cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(&tm, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return; }
    long long* out; cudaMalloc(&out, sizeof(long long) * grid);
    const int smem = STAGES * kBoxBytes + 1024 + 256;
    auto kern = bw_kernel<CS, STAGES, NBOX, PMODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (CS > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, kern, tm, iters, share, out);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CS=%d launch failed: %s\n", CS, cudaGetErrorString(e)); return; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long h[160]; cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double bytes = (double)iters * kBoxBytes;
    printf("%s%s pmode %d boxes/stage %d  cluster %2d  share %2d  stages %d  grid %3d: %6.1f B/clk/SM received (producer clocks)  %6.2f TB/s aggregate received, %6.2f TB/s issued from L2  (%.3f ms)\n",
           tf32 ? "TFLOAT32" : "FLOAT32 ", atom32 ? " atom32" : "       ", PMODE, NBOX, CS, share, STAGES, grid, bytes / avg, bytes * grid / (ms * 1e-3) / 1e12, bytes * grid / CS / (ms * 1e-3) / 1e12, ms);
    cudaFree(out);
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fn;
    float* d; cudaMalloc(&d, (size_t)kRowsM * kColsK * 4);
    cudaMemset(d, 0, (size_t)kRowsM * kColsK * 4);
    const int iters = 4000;
    run<1, 8, 1, 0>(enc, d, 1, 148, iters);
    run<1, 8, 4, 0>(enc, d, 1, 148, iters);
    run<1, 8, 4, 1>(enc, d, 1, 148, iters);
    run<1, 8, 4, 2>(enc, d, 1, 148, iters);
    run<1, 8, 2, 0>(enc, d, 1, 148, iters);
    run<1, 8, 2, 1>(enc, d, 1, 148, iters);
    run<1, 8, 2, 2>(enc, d, 1, 148, iters);
    run<1, 8, 8, 0>(enc, d, 1, 148, iters);
    run<1, 8, 8, 1>(enc, d, 1, 148, iters);
    run<1, 8, 8, 2>(enc, d, 1, 148, iters);
    return 0;
}
