// bx_convert.cuh — the in-shared-memory fp32 -> split-bf16 conversion shared by the "bf16x3" kernels
// (gemm_bx.cu, conv_tc.cu): 256 converter threads rewrite a TMA-landed fp32 operand stage in place as the bf16 hi / mid
// images the tensor core reads (see gemm_bx.cu for the layouts and the measurements behind this structure).
#pragma once

#include "common.cuh"
#include "ptx.cuh"

namespace npm {
namespace bx {

constexpr int kConvWarps = 8;

// {lo16 = bf16_rn(lo), hi16 = bf16_rn(hi)}
__device__ __forceinline__ uint32_t bf16x2_rn(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bar_sync_conv() { asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvWarps) : "memory"); }
// plain (release.cta) arrive on a barrier of another CTA of the cluster — the form CUTLASS's ClusterBarrier uses.  The
// release.cluster form lowers to MEMBAR.ALL.GPU (~2000 clk measured); nothing the leader reads depends on it: the data
// this signals sits in THIS CTA's shared memory, already proxy-fenced, and is read by THIS SM's tensor core.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}

// One operand of one CTA: R rows (A: 128 rows of M, B: BLOCK_N/2 rows of N) x 32 k per stage, 16 KB (R = 128) as
// fp32 and as bf16 hi + mid.  Each of the 256 converter threads owns NLD 16-byte fp32 chunks of the stage.
template <int R, bool MN, int NTERMS>
struct OperandConverter {
    static constexpr int NLD = R / 32;
    // K-major : TMA box {32 k, R rows}: row r = 128 B, 16-byte chunk c at position c ^ (r & 7).  A warp instruction
    //           covers 4 rows (8 lanes per row); the rows of one half-warp are {r, r+4} so that the two 64-byte hi
    //           halves it writes never share a bank group.
    // MN-major: TMA boxes {32 mn, 32 k}: 4 KB per 32-mn chunk, k-row kk = 128 B, chunk j at j ^ (kk & 7).  A warp
    //           instruction covers 4 k-rows of one chunk, again {kk, kk+4} per half-warp.
    uint32_t src[NLD];       // byte offset of the fp32 chunk in the staged image
    uint32_t dst[NLD];       // byte offset of the 8-byte hi store in the bf16 image

    __device__ __forceinline__ void init(int pw, int lane) {
        const int c = lane & 7, q = lane >> 3;
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            if (!MN) {
                const int g = i >> 1, j = i & 1;
                const int r = pw * (R / 8) + g * 8 + (q & 1) * 4 + (q >> 1) + 2 * j;
                src[i] = uint32_t(r) * 128u + (uint32_t(c ^ (r & 7)) << 4);
                dst[i] = uint32_t(r) * 128u + (uint32_t((c >> 1) ^ (r & 7)) << 4) + uint32_t(c & 1) * 8u;
            } else {
                const int t = pw * NLD + i;                       // warp instruction index: chunk t / 8, row group t % 8
                const int ch = t >> 3, g8 = t & 7;
                const int kk = (g8 >> 1) * 8 + (q & 1) * 4 + (q >> 1) + 2 * (g8 & 1);
                const int mn = ch * 32 + 4 * c;
                src[i] = uint32_t(ch) * 4096u + uint32_t(kk) * 128u + (uint32_t(c ^ (kk & 7)) << 4);
                dst[i] = uint32_t(mn >> 6) * 4096u + uint32_t(kk) * 128u + (uint32_t(((mn & 63) >> 3) ^ (kk & 7)) << 4) +
                         uint32_t((mn >> 2) & 1) * 8u;
            }
        }
    }
    __device__ __forceinline__ void load(float4 (&v)[NLD], uint32_t image) const {
#pragma unroll
        for (int i = 0; i < NLD; ++i) v[i] = lds_f4(image + src[i]);
    }
    __device__ __forceinline__ void store(const float4 (&v)[NLD], uint32_t image) const {
        constexpr uint32_t mid_delta = uint32_t(R / 64) * 4096u;   // MN-major: the mid image follows the hi image
#pragma unroll
        for (int i = 0; i < NLD; ++i) {
            const uint32_t h01 = bf16x2_rn(v[i].x, v[i].y), h23 = bf16x2_rn(v[i].z, v[i].w);
            const uint32_t a = image + dst[i];
            sts_v2(a, h01, h23);
            if (NTERMS == 3) {
                const float rx = v[i].x - __uint_as_float(h01 << 16), ry = v[i].y - __uint_as_float(h01 & 0xffff0000u);
                const float rz = v[i].z - __uint_as_float(h23 << 16), rw = v[i].w - __uint_as_float(h23 & 0xffff0000u);
                sts_v2(MN ? a + mid_delta : (a ^ 64u), bf16x2_rn(rx, ry), bf16x2_rn(rz, rw));
            }
        }
    }
};

// UMMA descriptors of the converted images and the byte offset of K16 slice s of the hi (t = 0) / mid (t = 1) image of
// an operand with R rows
constexpr uint64_t kDescK  = ptx::umma_desc_base(2 /*SWIZZLE_128B*/, 16, 1024);
constexpr uint64_t kDescMN = ptx::umma_desc_base(2 /*SWIZZLE_128B*/, 4096, 1024);   // LBO: next 64-mn slab, SBO: next 8 k-rows
template <int R, bool MN>
__device__ __forceinline__ constexpr uint32_t slice_off(int t, int s) {
    return MN ? uint32_t(t) * (R / 64) * 4096u + uint32_t(s) * 2048u : uint32_t(t) * 64u + uint32_t(s) * 32u;
}

}  // namespace bx
}  // namespace npm
