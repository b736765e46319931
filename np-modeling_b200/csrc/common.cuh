// common.cuh — shared device/host helpers for libnpm_b200 (sm_100a only).
//
// Everything in this library is launched asynchronously on a caller-supplied
// stream, never allocates, and reports failure through an integer return code
// plus npm_last_error(). See include/npm_b200.h for the C-ABI contract.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

#include "../../include/npm_b200.h"

namespace npm {

// ---------------------------------------------------------------------------
// Error plumbing (host)
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what);       // cudaGetLastError → rc
int  num_sms();                            // cached SM count of current device
void count_launch(int n = 1);              // npm_launch_count() bookkeeping

#define NPM_REQUIRE(cond, ...)                                              \
    do {                                                                    \
        if (!(cond)) {                                                      \
            ::npm::set_error(__VA_ARGS__);                                  \
            return NPM_ERR_INVALID;                                         \
        }                                                                   \
    } while (0)

// ---------------------------------------------------------------------------
// Programmatic dependent launch.  ~1000 dependent kernels per training step leave the GPU idle ~2 us at every boundary
// (launch latency + the next kernel's prologue).  A kernel launched through launch_pdl() may start while its
// predecessor in the stream is still running: every such kernel calls pdl_trigger() first (its own successor may be
// scheduled) and pdl_wait() before it touches ANY global memory — only barrier / TMEM / tensor-map set-up may precede
// the wait.  pdl_wait() returns once all prerequisite grids have completed and their writes are visible, so stream
// order semantics are unchanged.  Without the launch attribute both instructions are no-ops.  NPM_NO_PDL=1 disables.
// ---------------------------------------------------------------------------
bool pdl_enabled();
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... Params, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = (unsigned)n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------------------
// Small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Streaming 128-bit accesses: data touched once should not pollute L1.  Deliberately NOT
// the .nc path: several callers run in place (y aliases x), which .nc forbids.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// Grid sizing for grid-stride bandwidth kernels: a whole number of waves.
inline int bw_grid(size_t work_items, int threads, int ctas_per_sm = 8) {
    size_t need = (work_items + threads - 1) / threads;
    size_t cap = (size_t)num_sms() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------------------
// Philox-4x32-10 (Salmon et al. 2011). Counter = (ctr_lo, ctr_hi, 0, 0),
// key = (seed_lo, seed_hi). oracle/philox.py restates this bit for bit.
// ---------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b,
                                                        uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t counter, uint64_t seed) {
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = 0u, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

}  // namespace npm
