// attn_bwd.cu — fused attention core backward on tcgen05 / TMEM / TMA (TF32, head dim 64).
//
// Replaces the reference's  dV = P^T dO, dP = dO V^T, Softmax.backward Jacobian einsum, dQ = dS K,
// dK = dS^T Q  (layers/attentions.py:146-162, layers/activations.py:33-45) without ever
// materialising P, dP or dS in HBM.  P is recomputed from the saved log-sum-exp:
//     P = exp2(c * Q K^T - L),   dS = P o (dP - D) / sqrt(dk),   D = rowsum(dO o O)
// Two persistent kernels, each shaped like the forward one: TMA producers / one MMA issuer / four
// "exp" warps (P from S, in place in TMEM) / four "dS" warps (dS from P and dP, in place over dP),
// every thread of the last two groups owning one TMEM lane:
//   attn_bwd_dkdv : work item = (b, h, 128 kv rows), loops over 128-row q blocks, works on the
//                   TRANSPOSED score tile S^T = K Q^T (TMEM lane = kv row) so that P^T and dS^T are
//                   directly the TMEM A operands of  dV += P^T dO  and  dK += dS^T Q.
//   attn_bwd_dq   : work item = (b, h, 128 q rows), loops over kv blocks, S = Q K^T (lane = q row),
//                   dS is the TMEM A operand of  dQ += dS K.
// A [rows, 64] fp32 tile that is contracted over its 64 columns is landed with the plain 128-byte
// swizzle ("R" image, K-major operand); contracted over its rows it is landed with 32-byte swizzle
// atoms ("T" image, MN-major operand) — tf32 operands cannot be transposed at 16-byte granularity.
//
// BX = true: the split-bf16 ("bf16x3") variant of both kernels, as in attn_fwd.cu: q / k / v / dO arrive as bf16 hi + mid
// planes (attn_split_kernel below), every product runs mid*hi + hi*mid + hi*hi with kind::f16, P / dS go back to TMEM as
// packed bf16 pairs (per 64-column half: 32 columns hi, 32 columns mid).  A bf16 [128, 64] tile with the 128-byte
// swizzle serves as the "R" and as the "T" image at once; the T loads below then land the same bytes again.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace npm {

int gcd_int(int a, int b);
int make_tensor_map_4d(CUtensorMap* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                       uint64_t s1, uint64_t s2, uint64_t s3, uint32_t b0, uint32_t b1, bool round_tf32,
                       bool atom32b);
int make_tensor_map_bf16_planes(CUtensorMap* tm, const void* base, uint64_t S, uint64_t H, uint64_t B, uint64_t ld,
                                uint64_t plane_elems, uint32_t box_rows);

namespace {

constexpr int kBlk = 128;         // rows of every tile (q rows and kv rows)
constexpr int kD   = 64;          // head dim
constexpr int kTileBytes  = kBlk * kD * 4;   // 32 KB
constexpr int kChunkBytes = 16384;           // one {32 d, 128 rows} TMA box
constexpr int kThreads = 384;       // 4 control warps + 4 exp warps + 4 dS warps
constexpr int kThreadsBx = 640;     // split-bf16 variant: 8 exp warps + 8 dS warps (two warps per TMEM lane quarter, each half the columns)

struct BwdArgs {
    int B, H, Sq, Skv;
    int n_q, n_kv, total_items;
    float c;                  // log2(e) / sqrt(dk)
    float scale;              // 1 / sqrt(dk)
    const float* lse;         // [B, H, Sq]  (log2 domain)
    const float* dsum;        // [B, H, Sq]  D = rowsum(dO o O)
    float* dq;                // [B, Sq, H, 64]   token stride lddq
    float* dk;                // [B, Skv, H, 64]  token stride lddk
    float* dv;                // [B, Skv, H, 64]  token stride lddv
    int64_t lddq, lddk, lddv;
    int t0;                   // split-bf16 variant: first term to run (0 = bf16x3: mid*hi, hi*mid, hi*hi; 2 = plain bf16: hi*hi only)
    int causal;               // kv position t > q position s is masked (Sq == Skv): blocks above the diagonal are skipped
    long long* dbg;           // tools only: per-CTA cycle counters of the dK/dV MMA issuer's waits (NPM_ATTN_DEBUG_TIMES)
    int debug_skip;           // tools only: 1 = exp warps do no work, 2 = dS warps do no work, 3 = both (timing experiments)
    int halves;               // split-bf16 kernels: dP / dP^T as two N = 64 products interleaved with the dQ / dK halves (1) or one N = 128 product (0)
};

__device__ __forceinline__ uint32_t cvt_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// fp32 → tf32 round-to-nearest on the integer ALU (kind::tf32 ignores the low 13 bits); see attn_fwd.cu
__device__ __forceinline__ uint32_t rna_tf32(float x) { return __float_as_uint(x) + 0x1000u; }
template <int N>
__device__ __forceinline__ void bar_sync_n(int id) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory");
}
// both 16 KB boxes of a [128 rows, 64] tile: fp32 = d 0..31 and d 32..63; split bf16 = hi plane and mid plane
template <bool BX>
__device__ __forceinline__ void load_tile(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int row0, int h, int b) {
    if (BX) {
        ptx::tma_load_5d(dst, tm, bar, 0, row0, h, b, 0);
        ptx::tma_load_5d(dst + kChunkBytes, tm, bar, 0, row0, h, b, 1);
    } else {
        ptx::tma_load_4d(dst, tm, bar, 0, row0, h, b);
        ptx::tma_load_4d(dst + kChunkBytes, tm, bar, 32, row0, h, b);
    }
}
// D[tmem] (=|+=) A[smem, K-major R image] * B[smem, K-major R image]^T over the 64-wide head dim
template <bool BX>
__device__ __forceinline__ void mma_rr(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, int t0 = 0) {
    const uint64_t desc_k = ptx::umma_desc_base(2 /*SWIZZLE_128B*/, 16, 1024);
    if (BX) {
        const uint64_t da0 = ptx::umma_desc(desc_k, a_addr), db0 = ptx::umma_desc(desc_k, b_addr);
#pragma unroll
        for (int t = 0; t < 3; ++t) {          // mid*hi, hi*mid, hi*hi (the mid image follows the hi image); t0 = 2: hi*hi only
            if (t < t0) continue;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                ptx::umma_f16(d_tmem, ptx::umma_desc_off(da0, (t == 0 ? kChunkBytes : 0) + kk * 32),
                              ptx::umma_desc_off(db0, (t == 1 ? kChunkBytes : 0) + kk * 32), idesc, (t > t0 || kk != 0) ? 1u : 0u);
        }
        return;
    }
#pragma unroll
    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            ptx::umma_tf32(d_tmem, ptx::umma_desc(desc_k, a_addr + kb * kChunkBytes + kk * 32),
                           ptx::umma_desc(desc_k, b_addr + kb * kChunkBytes + kk * 32), idesc, (kb | kk) != 0 ? 1u : 0u);
}
// D[tmem, 128 x 64] (=|+=) A[tmem, 128 lanes x 128 cols] * B[smem T image: 128 rows (k) x 64 (n)]
template <bool BX>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, uint32_t idesc, bool accumulate, int t0 = 0) {
    if (BX) {
        // A: per 32-k quarter [16 columns hi | 16 columns mid], two bf16 per column; a K16 step = 8 columns of A and 16
        // rows (2048 B) of the B image (plain 128-byte swizzle, N = 64 = one 128-byte chunk, 8-row groups 1024 B apart)
        const uint64_t db0 = ptx::umma_desc(ptx::umma_desc_base(2 /*SWIZZLE_128B*/, kChunkBytes, 1024), b_addr);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (t < t0) continue;
#pragma unroll
            for (int kk = 0; kk < kBlk / 16; ++kk)
                ptx::umma_f16_ts(d_tmem, a_tmem + (kk >> 1) * 32 + (t == 0 ? 16 : 0) + (kk & 1) * 8,
                                 ptx::umma_desc_off(db0, (t == 1 ? kChunkBytes : 0) + kk * 2048), idesc,
                                 (accumulate || t > t0 || kk != 0) ? 1u : 0u);
        }
        return;
    }
    const uint64_t desc_mn = ptx::umma_desc_base(1 /*SWIZZLE_128B_BASE32B*/, kChunkBytes, 512);
#pragma unroll
    for (int kk = 0; kk < kBlk / 8; ++kk)
        ptx::umma_tf32_ts(d_tmem, a_tmem + kk * 8, ptx::umma_desc(desc_mn, b_addr + kk * 1024), idesc,
                          (accumulate || kk != 0) ? 1u : 0u);
}
// p0, p1 (consecutive k) -> packed bf16 hi pair and mid pair; and back
__device__ __forceinline__ void split_pack(float p0, float p1, uint32_t& hi, uint32_t& mid) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(p1), "f"(p0));
    const float r0 = p0 - __uint_as_float(hi << 16), r1 = p1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ float unpack_lo(uint32_t hi, uint32_t mid) { return __uint_as_float(hi << 16) + __uint_as_float(mid << 16); }
__device__ __forceinline__ float unpack_hi(uint32_t hi, uint32_t mid) {
    return __uint_as_float(hi & 0xffff0000u) + __uint_as_float(mid & 0xffff0000u);
}
// 32 columns of one accumulator row → 128 contiguous bytes of global memory
__device__ __forceinline__ void store_acc_half(uint32_t taddr, float* dst, bool live) {
    uint32_t o[32];
    ptx::tmem_ld_32x32(taddr, o);
    ptx::tmem_ld_wait();
    if (live) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            d4[k] = make_float4(__uint_as_float(o[4 * k]), __uint_as_float(o[4 * k + 1]), __uint_as_float(o[4 * k + 2]),
                                __uint_as_float(o[4 * k + 3]));
    }
}
// one accumulator row (64 fp32 in TMEM) → 256 contiguous bytes of global memory
__device__ __forceinline__ void store_acc_row(uint32_t taddr, float* dst, bool live) {
    uint32_t o[kD];
    ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
    ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
    ptx::tmem_ld_wait();
    if (live) {
        float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int k = 0; k < kD / 4; ++k)
            d4[k] = make_float4(__uint_as_float(o[4 * k]), __uint_as_float(o[4 * k + 1]), __uint_as_float(o[4 * k + 2]),
                                __uint_as_float(o[4 * k + 3]));
    }
}

// =============================================================================== dK, dV
// Measured design notes (NPM_ATTN_DEBUG_SKIP / NPM_ATTN_DEBUG_TIMES): with the MMAs and the elementwise work both
// skipped this kernel still takes ~3700 clk per 128-row q block — it is bound by the 128 KB of tile loads per block
// (Q and dO, each as an R and a T image) at ~35 B/clk/SM with 128 KB in flight.  A variant with 64-row q blocks and
// every stream double buffered (same bytes in flight) measured 204 us against 152 us for this one at B8 H16 S1024, and a
// CTA-pair variant (cta_group::2, M = 256: each CTA loads only half of every streamed tile and the TS products run at
// 32 instead of 45 clk per K8 step, tools/micro/mma_rate.cu) 168 us: what bounds a q block is the dependency chain
// dP^T MMA -> dS warps -> dK MMA -> next dP^T MMA through the single dP^T buffer (TMEM is full: 2 x S^T, dP^T, dV, dK =
// 512 columns), not the loads, and the pair's cross-CTA mbarrier hops lengthen exactly that chain.
// smem: K_R, V_R (resident per item) | Q_R, dO_R, Q_T, dO_T (one q block each) | L/D staging | barriers
constexpr int kKvSmem = 6 * kTileBytes + 2 * 2 * kBlk * 4 + 1024 + 256;

template <bool CAUSAL, bool BX>
__global__ void __launch_bounds__(BX ? kThreadsBx : kThreads, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQr, const __grid_constant__ CUtensorMap tmQt,
                     const __grid_constant__ CUtensorMap tmKr, const __grid_constant__ CUtensorMap tmVr,
                     const __grid_constant__ CUtensorMap tmDOr, const __grid_constant__ CUtensorMap tmDOt,
                     const BwdArgs args) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr  = ptx::smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr        = smem_raw + (base_addr - raw_addr);

    const uint32_t kr_addr = base_addr, vr_addr = kr_addr + kTileBytes;
    const uint32_t qr_addr = vr_addr + kTileBytes, dor_addr = qr_addr + kTileBytes;
    const uint32_t qt_addr = dor_addr + kTileBytes, dot_addr = qt_addr + kTileBytes;
    float* lsm = reinterpret_cast<float*>(base_ptr + 6 * kTileBytes);   // [2][128] L, then [2][128] D
    float* dsm = lsm + 2 * kBlk;
    const uint32_t bar_addr = base_addr + 6 * kTileBytes + 2 * 2 * kBlk * 4;
    enum { K_FULL = 0, K_EMPTY, V_FULL, V_EMPTY, QR_FULL, QR_EMPTY, DOR_FULL, DOR_EMPTY, QT_FULL, QT_EMPTY, DOT_FULL,
           DOT_EMPTY, ST_FULL0, ST_FULL1, P_READY0, P_READY1, DPT_FULL, DS_READY, DV_DONE, DK_DONE, NBAR };
    auto bar = [&](int i) { return bar_addr + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 6 * kTileBytes + 2 * 2 * kBlk * 4 + 8 * NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQr); ptx::prefetch_tensormap(&tmQt); ptx::prefetch_tensormap(&tmKr);
        ptx::prefetch_tensormap(&tmVr); ptx::prefetch_tensormap(&tmDOr); ptx::prefetch_tensormap(&tmDOt);
    }
    if (warp == 3) {
        if (lane == 0) {
            for (int i = 0; i < NBAR; ++i)
                ptx::mbar_init(bar(i), (i == P_READY0 || i == P_READY1 || i == DS_READY) ? (BX ? 256 : 128) : 1);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();               // only shared memory / TMEM set-up above
    const uint32_t tm_dpt = tmem_base + 256, tm_dv = tmem_base + 384, tm_dk = tmem_base + 448;

    const int n_q = args.n_q;
    // first q block of an item (b, h, kv tile nt): 0, or nt under the causal mask (q blocks above the diagonal see none
    // of this kv tile)
    auto first_q = [&](int item) -> int { return CAUSAL ? item % args.n_kv : 0; };
    uint32_t G = 0;                                            // q blocks this CTA walks, over all its items
    for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) G += (uint32_t)(n_q - first_q(item));

    if (warp == 0) {
        // ============ producer: K_R, V_R per item; Q_R, dO_R per q block ============
        // K_R is released by the LAST S^T of an item (issued two blocks early) and V_R by its last
        // dP^T, so the next item's K / V land while the current item is still finishing; V is queued
        // behind Q_R(0) because S^T(0) of the next item is issued before V_R is free.
        if (ptx::elect_one()) {
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
                const int nt = item % args.n_kv;
                const int bh = item / args.n_kv;
                const int h = bh % args.H, b = bh / args.H;
                ptx::mbar_wait(bar(K_EMPTY), (it & 1) ^ 1u);
                ptx::mbar_arrive_expect_tx(bar(K_FULL), kTileBytes);
                load_tile<BX>(kr_addr, &tmKr, bar(K_FULL), nt * kBlk, h, b);
                const int i0 = first_q(item);
                for (int i = i0; i < n_q; ++i, ++g) {
                    ptx::mbar_wait(bar(QR_EMPTY), (g & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(QR_FULL), kTileBytes);
                    load_tile<BX>(qr_addr, &tmQr, bar(QR_FULL), i * kBlk, h, b);
                    if (i == i0) {
                        ptx::mbar_wait(bar(V_EMPTY), (it & 1) ^ 1u);
                        ptx::mbar_arrive_expect_tx(bar(V_FULL), kTileBytes);
                        load_tile<BX>(vr_addr, &tmVr, bar(V_FULL), nt * kBlk, h, b);
                    }
                    ptx::mbar_wait(bar(DOR_EMPTY), (g & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(DOR_FULL), kTileBytes);
                    load_tile<BX>(dor_addr, &tmDOr, bar(DOR_FULL), i * kBlk, h, b);
                }
            }
        }
    } else if (warp == 1) {
        // ============ producer: dO_T, Q_T per q block ============
        if (ptx::elect_one()) {
            uint32_t g = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) {
                const int bh = item / args.n_kv;
                const int h = bh % args.H, b = bh / args.H;
                for (int i = first_q(item); i < n_q; ++i, ++g) {
                    ptx::mbar_wait(bar(DOT_EMPTY), (g & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(DOT_FULL), kTileBytes);
                    load_tile<BX>(dot_addr, &tmDOt, bar(DOT_FULL), i * kBlk, h, b);
                    ptx::mbar_wait(bar(QT_EMPTY), (g & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(QT_FULL), kTileBytes);
                    load_tile<BX>(qt_addr, &tmQt, bar(QT_FULL), i * kBlk, h, b);
                }
            }
        }
    } else if (warp == 2) {
        // ============ MMA issuer ============
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_s  = BX ? ptx::umma_idesc_bf16(kBlk, kBlk, false, false) : ptx::umma_idesc_tf32(kBlk, kBlk, false, false);
            constexpr uint32_t idesc_ts = BX ? ptx::umma_idesc_bf16(kBlk, kD, false, true) : ptx::umma_idesc_tf32(kBlk, kD, false, true);
            long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            long long* const dbg = args.dbg;
            auto twait = [&](int slot, int which, uint32_t parity) {
                if (dbg) {
                    const long long t0 = clock64();
                    ptx::mbar_wait(bar(which), parity);
                    acc[slot] += clock64() - t0;
                } else {
                    ptx::mbar_wait(bar(which), parity);
                }
            };
            // (item, q block) sequence of this CTA; S^T runs two blocks and dP^T one block ahead of dV / dK: three cursors
            struct Cur { int item, it, i, i0; };
            auto cur_init = [&](Cur& c) {
                c.item = blockIdx.x; c.it = 0;
                c.i0 = c.item < args.total_items ? first_q(c.item) : 0;
                c.i = c.i0;
            };
            auto cur_next = [&](Cur& c) {
                if (++c.i == n_q) {
                    c.item += gridDim.x; ++c.it;
                    c.i0 = c.item < args.total_items ? first_q(c.item) : 0;
                    c.i = c.i0;
                }
            };
            Cur c_st, c_dpt, c_main;
            cur_init(c_st); cur_init(c_dpt); cur_init(c_main);
            auto issue_st = [&](uint32_t g) {        // S^T(g) = K Q^T
                const int it = c_st.it, i = c_st.i;
                const bool first = i == c_st.i0;
                cur_next(c_st);
                if (first) twait(0, K_FULL, it & 1);
                twait(1, QR_FULL, g & 1);
                ptx::tc_fence_after();
                mma_rr<BX>(tmem_base + (g & 1u) * kBlk, kr_addr, qr_addr, idesc_s, args.t0);
                ptx::umma_commit(bar(QR_EMPTY));
                ptx::umma_commit(bar(ST_FULL0 + (g & 1u)));
                if (i == n_q - 1) ptx::umma_commit(bar(K_EMPTY));
            };
            auto issue_dpt = [&](uint32_t g) {       // dP^T(g) = V dO^T
                const int it = c_dpt.it, i = c_dpt.i;
                const bool first = i == c_dpt.i0;
                cur_next(c_dpt);
                if (first) twait(2, V_FULL, it & 1);
                twait(3, DOR_FULL, g & 1);
                ptx::tc_fence_after();
                mma_rr<BX>(tm_dpt, vr_addr, dor_addr, idesc_s, args.t0);
                ptx::umma_commit(bar(DOR_EMPTY));
                ptx::umma_commit(bar(DPT_FULL));
                if (i == n_q - 1) ptx::umma_commit(bar(V_EMPTY));
            };
            const long long t_begin = dbg ? clock64() : 0;
            if (G > 0) { issue_st(0); issue_dpt(0); }
            if (G > 1) issue_st(1);
            for (uint32_t g = 0; g < G; ++g) {
                const int i = c_main.i;
                const bool first = i == c_main.i0;
                cur_next(c_main);
                twait(4, P_READY0 + (g & 1u), (g >> 1) & 1);
                twait(5, DOT_FULL, g & 1);
                ptx::tc_fence_after();
                mma_ts<BX>(tm_dv, tmem_base + (g & 1u) * kBlk, dot_addr, idesc_ts, !first, args.t0);    // dV += P^T dO
                ptx::umma_commit(bar(DOT_EMPTY));
                if (i == n_q - 1) ptx::umma_commit(bar(DV_DONE));
                twait(6, DS_READY, g & 1);
                twait(7, QT_FULL, g & 1);
                ptx::tc_fence_after();
                mma_ts<BX>(tm_dk, tm_dpt, qt_addr, idesc_ts, !first, args.t0);                          // dK += dS^T Q
                ptx::umma_commit(bar(QT_EMPTY));
                if (i == n_q - 1) ptx::umma_commit(bar(DK_DONE));
                if (g + 1 < G) issue_dpt(g + 1);
                if (g + 2 < G) issue_st(g + 2);
            }
            if (dbg) {
                acc[8] = clock64() - t_begin;
                acc[9] = G;
                for (int k = 0; k < 12; ++k) dbg[blockIdx.x * 24 + k] = acc[k];
            }
        }
    } else if (warp >= 4) {
        // ============ warps 4-7: P^T = exp2(c S^T - L[q]) in place.  warps 8-11: dS^T = P^T o (dP^T - D[q]) / sqrt(dk)
        // in place over dP^T.  Thread = kv row = TMEM lane; the two groups run one block apart, so the
        // exponentials of block g+1 overlap the dS / dK / dP^T chain of block g. ============
        // BX: 8 warps per group — warps w and w + 4 of a group share a TMEM lane quarter and take half of the 128
        // columns each (quarters q0..q1 of 32 columns), which halves the serial elementwise time on the
        // S^T -> P^T -> dV and dP^T -> dS^T -> dK chains (the split to bf16 hi / mid doubles the work per element)
        constexpr int NSET = BX ? 2 : 1;
        const bool exp_group = warp < 4 + 4 * NSET;
        const int wq = warp & 3;
        const int hsel = BX ? ((warp - 4) >> 2) & 1 : 0;
        const int q0 = BX ? 2 * hsel : 0, q1 = BX ? 2 * hsel + 2 : 4;
        const int tid = wq * 32 + lane;             // kv row within the tile (also: which q column's L / D this thread stages)
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const float c = args.c, scale = args.scale;
        const float* vec = exp_group ? args.lse : args.dsum;      // per-q-row vector this group consumes
        float* stage = exp_group ? lsm : dsm;
        const float pad = exp_group ? INFINITY : 0.0f;            // exp2(-inf) = 0 for padded q columns
        const int bar_id = exp_group ? 1 : 2;
        auto fetch = [&](int item, int i) -> float {
            const int bh = item / args.n_kv;
            const int qi = i * kBlk + tid;
            return qi < args.Sq ? __ldg(vec + (size_t)bh * args.Sq + qi) : pad;
        };
        uint32_t g = 0;
        int it = 0;
        float next = (int)blockIdx.x < args.total_items ? fetch(blockIdx.x, first_q(blockIdx.x)) : 0.0f;
        for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
            const int nt = item % args.n_kv;
            const int bh = item / args.n_kv;
            const int h = bh % args.H, b = bh / args.H;
            for (int i = first_q(item); i < n_q; ++i, ++g) {
                const uint32_t buf = g & 1u;
                // stage this block's L (or D) — fetched one block ago — and prefetch the next block's
                const bool rec = args.dbg != nullptr && lane == 0 && (warp == 4 || warp == 4 + 4 * NSET);
                long long* const dbg = args.dbg + blockIdx.x * 24 + (exp_group ? 12 : 15);
                long long t0 = rec ? clock64() : 0, t1;
#define BWD_STAMP(i) { if (rec) { t1 = clock64(); dbg[i] += t1 - t0; t0 = t1; } __syncwarp(); }
                if (hsel == 0) stage[buf * kBlk + tid] = next;
                {
                    int ni = i + 1, nitem = item;
                    if (ni == n_q) { nitem = item + gridDim.x; ni = nitem < args.total_items ? first_q(nitem) : 0; }
                    if (nitem < args.total_items) next = fetch(nitem, ni);
                }
                bar_sync_n<128 * NSET>(bar_id);
                BWD_STAMP(0)
                const float4* V4 = reinterpret_cast<const float4*>(stage + buf * kBlk);
                if (exp_group) {
                    ptx::mbar_wait(bar(ST_FULL0 + buf), (g >> 1) & 1);
                    BWD_STAMP(1)
                    ptx::tc_fence_after();
                    const uint32_t s_tmem = tmem_base + lane_off + buf * kBlk;
                    if (BX) {
#pragma unroll
                        for (int qt = q0; qt < q1; ++qt) {          // 32 columns: S^T in, [16 columns hi | 16 columns mid] of P^T out
                            float p[32];
                            ptx::tmem_ld_32x32(s_tmem + qt * 32, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int k = 0; k < 32; k += 4) {
                                const float4 l4 = V4[(qt * 32 + k) >> 2];
                                p[k]     = ptx::ex2(fmaf(p[k], c, -l4.x));
                                p[k + 1] = ptx::ex2(fmaf(p[k + 1], c, -l4.y));
                                p[k + 2] = ptx::ex2(fmaf(p[k + 2], c, -l4.z));
                                p[k + 3] = ptx::ex2(fmaf(p[k + 3], c, -l4.w));
                            }
                            if (CAUSAL && i == nt) {
#pragma unroll
                                for (int k = 0; k < 32; ++k)
                                    if (qt * 32 + k < tid) p[k] = 0.0f;
                            }
                            uint32_t hm[32];
#pragma unroll
                            for (int k = 0; k < 16; ++k) split_pack(p[2 * k], p[2 * k + 1], hm[k], hm[16 + k]);
                            ptx::tmem_st_32x32(s_tmem + qt * 32, hm);
                        }
                    } else
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (args.debug_skip & 1) break;
                        float p[64];
                        ptx::tmem_ld_32x32(s_tmem + hf * 64, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                        ptx::tmem_ld_32x32(s_tmem + hf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&p[32]));
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 64; k += 4) {
                            const float4 l4 = V4[(hf * 64 + k) >> 2];
                            p[k]     = ptx::ex2(fmaf(p[k], c, -l4.x));
                            p[k + 1] = ptx::ex2(fmaf(p[k + 1], c, -l4.y));
                            p[k + 2] = ptx::ex2(fmaf(p[k + 2], c, -l4.z));
                            p[k + 3] = ptx::ex2(fmaf(p[k + 3], c, -l4.w));
                            if (!BX) {
                                p[k] = __uint_as_float(rna_tf32(p[k])); p[k + 1] = __uint_as_float(rna_tf32(p[k + 1]));
                                p[k + 2] = __uint_as_float(rna_tf32(p[k + 2])); p[k + 3] = __uint_as_float(rna_tf32(p[k + 3]));
                            }
                        }
                        if (CAUSAL && i == nt) {            // diagonal block: q position (column) before kv position (lane)
#pragma unroll
                            for (int k = 0; k < 64; ++k)
                                if (hf * 64 + k < tid) p[k] = 0.0f;
                        }
                        ptx::tmem_st_32x32(s_tmem + hf * 64, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                        ptx::tmem_st_32x32(s_tmem + hf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&p[32]));
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(bar(P_READY0 + buf));
                    BWD_STAMP(2)
                } else {
                    ptx::mbar_wait(bar(P_READY0 + buf), (g >> 1) & 1);
                    BWD_STAMP(1)
                    ptx::mbar_wait(bar(DPT_FULL), g & 1);
                    BWD_STAMP(2)
                    ptx::tc_fence_after();
                    const uint32_t p_tmem = tmem_base + lane_off + buf * kBlk;
                    if (BX) {
#pragma unroll
                        for (int qt = q0; qt < q1; ++qt) {
                            uint32_t pm[32], dp[32];               // pm: [16 hi | 16 mid] of P^T, then of dS^T
                            ptx::tmem_ld_32x32(p_tmem + qt * 32, pm);
                            ptx::tmem_ld_32x32(tm_dpt + lane_off + qt * 32, dp);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int k = 0; k < 16; k += 2) {      // packed columns k, k+1 = elements 2k .. 2k+3 = one float4 of D
                                const float4 d4 = V4[(qt * 32 + 2 * k) >> 2];
                                const float s0 = unpack_lo(pm[k], pm[16 + k]) * ((__uint_as_float(dp[2 * k]) - d4.x) * scale);
                                const float s1 = unpack_hi(pm[k], pm[16 + k]) * ((__uint_as_float(dp[2 * k + 1]) - d4.y) * scale);
                                const float s2 = unpack_lo(pm[k + 1], pm[17 + k]) * ((__uint_as_float(dp[2 * k + 2]) - d4.z) * scale);
                                const float s3 = unpack_hi(pm[k + 1], pm[17 + k]) * ((__uint_as_float(dp[2 * k + 3]) - d4.w) * scale);
                                split_pack(s0, s1, pm[k], pm[16 + k]);
                                split_pack(s2, s3, pm[k + 1], pm[17 + k]);
                            }
                            ptx::tmem_st_32x32(tm_dpt + lane_off + qt * 32, pm);
                        }
                    } else
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        if (args.debug_skip & 2) break;
                        uint32_t pp[32], dp[32];
                        ptx::tmem_ld_32x32(p_tmem + ch * 32, pp);
                        ptx::tmem_ld_32x32(tm_dpt + lane_off + ch * 32, dp);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 32; k += 4) {
                            const float4 d4 = V4[(ch * 32 + k) >> 2];
                            dp[k]     = rna_tf32(__uint_as_float(pp[k])     * ((__uint_as_float(dp[k])     - d4.x) * scale));
                            dp[k + 1] = rna_tf32(__uint_as_float(pp[k + 1]) * ((__uint_as_float(dp[k + 1]) - d4.y) * scale));
                            dp[k + 2] = rna_tf32(__uint_as_float(pp[k + 2]) * ((__uint_as_float(dp[k + 2]) - d4.z) * scale));
                            dp[k + 3] = rna_tf32(__uint_as_float(pp[k + 3]) * ((__uint_as_float(dp[k + 3]) - d4.w) * scale));
                        }
                        ptx::tmem_st_32x32(tm_dpt + lane_off + ch * 32, dp);
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(bar(DS_READY));
                    BWD_STAMP(3)
                }
#undef BWD_STAMP
            }
            // ---- epilogue: the exp warps store this kv tile's dV rows, the dS warps its dK rows ----
            const bool rec_e = args.dbg != nullptr && lane == 0 && (warp == 4 || warp == 4 + 4 * NSET);
            const long long te = rec_e ? clock64() : 0;
            ptx::mbar_wait(bar(exp_group ? DV_DONE : DK_DONE), it & 1);
            ptx::tc_fence_after();
            const int t = nt * kBlk + tid;
            const bool live = t < args.Skv;
            const size_t tok = (size_t)b * args.Skv + (live ? t : 0);
            if (BX) {       // each of the two warp sets stores 32 of the 64 columns
                if (exp_group) store_acc_half(tm_dv + lane_off + 32 * hsel, args.dv + tok * args.lddv + (size_t)h * kD + 32 * hsel, live);
                else           store_acc_half(tm_dk + lane_off + 32 * hsel, args.dk + tok * args.lddk + (size_t)h * kD + 32 * hsel, live);
            } else {
                if (exp_group) store_acc_row(tm_dv + lane_off, args.dv + tok * args.lddv + (size_t)h * kD, live);
                else           store_acc_row(tm_dk + lane_off, args.dk + tok * args.lddk + (size_t)h * kD, live);
            }
            ptx::tc_fence_before();
            if (rec_e) args.dbg[blockIdx.x * 24 + (exp_group ? 19 : 20)] += clock64() - te;
            __syncwarp();
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc(tmem_base, 512);
}

// =============================================================================== dK, dV (split-bf16)
// The split-bf16 variant is its own kernel.  Measured on the shared-template version (NPM_ATTN_DEBUG_TIMES, B8 H16
// S1024, cycles per q block of 5700-6400): the dS warps work 2500 clk on a block and then wait 2100 clk for
// dK(g) (1080 clk of MMAs) -> dP^T(g+1) (768 clk) to pass through the single dP^T buffer — the period of a block is that
// serial chain, the tensor pipe is busy 56 % of it.  Three changes:
//   * HALF-TILE CHAIN.  dP^T is produced and consumed as two 64-column halves (q rows 0..63 / 64..127 of the block), each
//     with its own full / ready barrier and its own set of dS warps (the two warp sets of the BX layout already own one
//     half each).  dK(g) runs its four K16 steps of half A, then dP^T_A(g+1) overwrites half A while the steps of half B
//     still read theirs: per set the chain is dS -> 540 clk -> 384 clk -> dS instead of dS -> 1080 -> 768 -> dS, and
//     the two sets run staggered, so the other 3700 clk of MMAs per block (S^T, dV, the other half) fill the pipe
//     while a set computes.
//   * ONE IMAGE PER TILE.  A bf16 [128, 64] tile with the 128-byte swizzle is the K-major ("R") and the MN-major ("T")
//     operand at once, so Q and dO are loaded once per block instead of twice (64 KB instead of 128 KB per block; the
//     shared-template version measured ~3700 clk per block for its loads alone).  Q lives in a ring of three (S^T runs
//     two blocks ahead of dK), dO in a ring of two (dP^T runs one block ahead of dV).
//   * The L / D vectors are staged per warp set (128-thread named barriers), so the sets never meet.
// smem: K, V (resident per item) | Q ring x3 | dO ring x2 | L / D staging | barriers = all 227 KB of the SM.
constexpr int kKvBxStage = 2 * 2 * kBlk * 4;
constexpr int kKvBxUsed  = 7 * kTileBytes + kKvBxStage + 256;
constexpr int kKvBxSmem  = 232448;

// D[tmem, 128 x 64] (=|+=) A[tmem, K16 steps kk0..kk1 of a packed 128-column tile] * B[smem, the same rows of the bf16 image]
__device__ __forceinline__ void mma_ts_bx_range(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_addr, uint32_t idesc, bool accumulate,
                                                int t0, int kk0, int kk1) {
    const uint64_t db0 = ptx::umma_desc(ptx::umma_desc_base(2 /*SWIZZLE_128B*/, kChunkBytes, 1024), b_addr);
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        if (t < t0) continue;
#pragma unroll
        for (int kk = kk0; kk < kk1; ++kk)
            ptx::umma_f16_ts(d_tmem, a_tmem + (kk >> 1) * 32 + (t == 0 ? 16 : 0) + (kk & 1) * 8,
                             ptx::umma_desc_off(db0, (t == 1 ? kChunkBytes : 0) + kk * 2048), idesc,
                             (accumulate || t > t0 || kk != kk0) ? 1u : 0u);
    }
}

template <bool CAUSAL>
__global__ void __launch_bounds__(kThreadsBx, 1)
attn_bwd_dkdv_bx_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const BwdArgs args) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr  = ptx::smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr        = smem_raw + (base_addr - raw_addr);
    {
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if (base_addr - raw_addr + kKvBxUsed > dyn) __trap();      // the launch gives the kernel the whole SM; fail loudly if the base moved
    }
    const uint32_t kr_addr = base_addr, vr_addr = kr_addr + kTileBytes;
    const uint32_t q_addr = vr_addr + kTileBytes;                    // ring of 3
    const uint32_t do_addr = q_addr + 3 * kTileBytes;                // ring of 2
    float* lsm = reinterpret_cast<float*>(base_ptr + 7 * kTileBytes);   // [2][128] L, then [2][128] D
    float* dsm = lsm + 2 * kBlk;
    const uint32_t bar_addr = base_addr + 7 * kTileBytes + kKvBxStage;
    enum { K_FULL = 0, K_EMPTY, V_FULL, V_EMPTY, Q_FULL0, Q_FULL1, Q_FULL2, Q_EMPTY0, Q_EMPTY1, Q_EMPTY2, DO_FULL0, DO_FULL1,
           DO_EMPTY0, DO_EMPTY1, ST_FULL0, ST_FULL1, P_READY0, P_READY1, DPT_FULL_A, DPT_FULL_B, DS_READY_A, DS_READY_B,
           DV_DONE, DK_DONE, DK_FREE, NBAR };
    static_assert(8 * NBAR + 4 <= 256, "barrier block");
    auto bar = [&](int i) { return bar_addr + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 7 * kTileBytes + kKvBxStage + 8 * NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmK); ptx::prefetch_tensormap(&tmV); ptx::prefetch_tensormap(&tmDO);
    }
    if (warp == 3) {
        if (lane == 0) {
            for (int i = 0; i < NBAR; ++i)
                ptx::mbar_init(bar(i), (i == P_READY0 || i == P_READY1 || i == DK_FREE) ? 256 : (i == DS_READY_A || i == DS_READY_B) ? 128 : 1);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();               // only shared memory / TMEM set-up above
    const uint32_t tm_dpt = tmem_base + 256, tm_dv = tmem_base + 384, tm_dk = tmem_base + 448;

    const int n_q = args.n_q;
    auto first_q = [&](int item) -> int { return CAUSAL ? item % args.n_kv : 0; };
    uint32_t G = 0;                                            // q blocks this CTA walks, over all its items
    for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) G += (uint32_t)(n_q - first_q(item));

    if (warp == 0) {
        // ============ producer: K, V per item; Q per q block (ring of 3) ============
        if (ptx::elect_one()) {
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
                const int nt = item % args.n_kv;
                const int bh = item / args.n_kv;
                const int h = bh % args.H, b = bh / args.H;
                ptx::mbar_wait(bar(K_EMPTY), (it & 1) ^ 1u);
                ptx::mbar_arrive_expect_tx(bar(K_FULL), kTileBytes);
                load_tile<true>(kr_addr, &tmK, bar(K_FULL), nt * kBlk, h, b);
                const int i0 = first_q(item);
                for (int i = i0; i < n_q; ++i, ++g) {
                    const uint32_t slot = g % 3u;
                    ptx::mbar_wait(bar(Q_EMPTY0 + slot), ((g / 3u) & 1u) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(Q_FULL0 + slot), kTileBytes);
                    load_tile<true>(q_addr + slot * kTileBytes, &tmQ, bar(Q_FULL0 + slot), i * kBlk, h, b);
                    if (i == i0) {
                        ptx::mbar_wait(bar(V_EMPTY), (it & 1) ^ 1u);
                        ptx::mbar_arrive_expect_tx(bar(V_FULL), kTileBytes);
                        load_tile<true>(vr_addr, &tmV, bar(V_FULL), nt * kBlk, h, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============ producer: dO per q block (ring of 2) ============
        if (ptx::elect_one()) {
            uint32_t g = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) {
                const int bh = item / args.n_kv;
                const int h = bh % args.H, b = bh / args.H;
                for (int i = first_q(item); i < n_q; ++i, ++g) {
                    const uint32_t slot = g & 1u;
                    ptx::mbar_wait(bar(DO_EMPTY0 + slot), ((g >> 1) & 1u) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(DO_FULL0 + slot), kTileBytes);
                    load_tile<true>(do_addr + slot * kTileBytes, &tmDO, bar(DO_FULL0 + slot), i * kBlk, h, b);
                }
            }
        }
    } else if (warp == 2) {
        // ============ MMA issuer ============
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_s   = ptx::umma_idesc_bf16(kBlk, kBlk, false, false);
            constexpr uint32_t idesc_s64 = ptx::umma_idesc_bf16(kBlk, kBlk / 2, false, false);
            constexpr uint32_t idesc_ts  = ptx::umma_idesc_bf16(kBlk, kD, false, true);
            long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            long long* const dbg = args.dbg;
            auto twait = [&](int slot, int which, uint32_t parity) {
                if (dbg) {
                    const long long t0 = clock64();
                    ptx::mbar_wait(bar(which), parity);
                    acc[slot] += clock64() - t0;
                } else {
                    ptx::mbar_wait(bar(which), parity);
                }
            };
            struct Cur { int item, it, i, i0; };
            auto cur_init = [&](Cur& c) {
                c.item = blockIdx.x; c.it = 0;
                c.i0 = c.item < args.total_items ? first_q(c.item) : 0;
                c.i = c.i0;
            };
            auto cur_next = [&](Cur& c) {
                if (++c.i == n_q) {
                    c.item += gridDim.x; ++c.it;
                    c.i0 = c.item < args.total_items ? first_q(c.item) : 0;
                    c.i = c.i0;
                }
            };
            Cur c_st, c_dpt, c_main;
            cur_init(c_st); cur_init(c_dpt); cur_init(c_main);
            auto issue_st = [&](uint32_t g) {        // S^T(g) = K Q^T
                const int it = c_st.it, i = c_st.i;
                const bool first = i == c_st.i0;
                cur_next(c_st);
                if (first) twait(0, K_FULL, it & 1);
                twait(1, Q_FULL0 + g % 3u, (g / 3u) & 1u);
                ptx::tc_fence_after();
                mma_rr<true>(tmem_base + (g & 1u) * kBlk, kr_addr, q_addr + (g % 3u) * kTileBytes, idesc_s, args.t0);
                ptx::umma_commit(bar(ST_FULL0 + (g & 1u)));
                if (i == n_q - 1) ptx::umma_commit(bar(K_EMPTY));
            };
            const bool halves = args.halves != 0;
            auto issue_dpt = [&](uint32_t g, int half) {       // dP^T(g)[:, half] = V dO[64 q rows of the half]^T
                const int it = c_dpt.it, i = c_dpt.i;
                const bool first = i == c_dpt.i0;
                if (!halves) {             // one N = 128 product when called for half 1, nothing for half 0
                    if (half == 0) return;
                    cur_next(c_dpt);
                    if (first) twait(2, V_FULL, it & 1);
                    twait(3, DO_FULL0 + (g & 1u), (g >> 1) & 1u);
                    ptx::tc_fence_after();
                    mma_rr<true>(tm_dpt, vr_addr, do_addr + (g & 1u) * kTileBytes, idesc_s, args.t0);
                    ptx::umma_commit(bar(DPT_FULL_A));
                    ptx::umma_commit(bar(DPT_FULL_B));
                    if (i == n_q - 1) ptx::umma_commit(bar(V_EMPTY));
                    return;
                }
                if (half == 0) {
                    if (first) twait(2, V_FULL, it & 1);
                    twait(3, DO_FULL0 + (g & 1u), (g >> 1) & 1u);
                    ptx::tc_fence_after();
                } else {
                    cur_next(c_dpt);
                }
                mma_rr<true>(tm_dpt + half * 64, vr_addr, do_addr + (g & 1u) * kTileBytes + half * 8192, idesc_s64, args.t0);
                ptx::umma_commit(bar(DPT_FULL_A + half));
                if (half == 1 && i == n_q - 1) ptx::umma_commit(bar(V_EMPTY));
            };
            const long long t_begin = dbg ? clock64() : 0;
            if (G > 0) { issue_st(0); issue_dpt(0, 0); issue_dpt(0, 1); }
            if (G > 1) issue_st(1);
            for (uint32_t g = 0; g < G; ++g) {
                const int i = c_main.i, it_main = c_main.it;
                const bool first = i == c_main.i0, last = i == n_q - 1;
                cur_next(c_main);
                const uint32_t qa = q_addr + (g % 3u) * kTileBytes, da = do_addr + (g & 1u) * kTileBytes;
                twait(4, P_READY0 + (g & 1u), (g >> 1) & 1);
                ptx::tc_fence_after();
                mma_ts<true>(tm_dv, tmem_base + (g & 1u) * kBlk, da, idesc_ts, !first, args.t0);          // dV += P^T dO
                ptx::umma_commit(bar(DO_EMPTY0 + (g & 1u)));
                if (last) ptx::umma_commit(bar(DV_DONE));
                twait(5, DS_READY_A, g & 1);
                // the first dK product of an item overwrites the accumulator: BOTH warp sets must have stored the previous
                // item's dK rows (set B arrives on DS_READY_B later than this point)
                if (first && it_main > 0) twait(7, DK_FREE, (it_main - 1) & 1);
                ptx::tc_fence_after();
                mma_ts_bx_range(tm_dk, tm_dpt, qa, idesc_ts, !first, args.t0, 0, 4);                      // dK += dS^T Q, q rows 0..63
                if (g + 1 < G) issue_dpt(g + 1, 0);
                twait(6, DS_READY_B, g & 1);
                ptx::tc_fence_after();
                mma_ts_bx_range(tm_dk, tm_dpt, qa, idesc_ts, true, args.t0, 4, 8);                        // q rows 64..127
                ptx::umma_commit(bar(Q_EMPTY0 + g % 3u));
                if (last) ptx::umma_commit(bar(DK_DONE));
                if (g + 1 < G) issue_dpt(g + 1, 1);
                if (g + 2 < G) issue_st(g + 2);
            }
            if (dbg) {
                acc[8] = clock64() - t_begin;
                acc[9] = G;
                for (int k = 0; k < 12; ++k) dbg[blockIdx.x * 24 + k] = acc[k];
            }
        }
    } else if (warp >= 4) {
        // ============ warps 4-11: P^T = exp2(c S^T - L[q]) in place, packed.  warps 12-19: dS^T = P^T o (dP^T - D[q]) /
        // sqrt(dk) in place over dP^T.  Thread = kv row = TMEM lane; warps w and w + 4 of a group share a TMEM lane quarter,
        // set A (hsel = 0) owns q columns 0..63 of every block, set B columns 64..127. ============
        const bool exp_group = warp < 12;
        const int wq = warp & 3;
        const int hsel = ((warp - 4) >> 2) & 1;
        const int q0 = 2 * hsel, q1 = 2 * hsel + 2;
        const int tid = wq * 32 + lane;             // kv row within the tile
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const float c = args.c, scale = args.scale;
        const float* vec = exp_group ? args.lse : args.dsum;      // per-q-row vector this group consumes
        float* stage = (exp_group ? lsm : dsm) + hsel * 64;       // this set's 64 columns of either buffer
        const float pad = exp_group ? INFINITY : 0.0f;            // exp2(-inf) = 0 for padded q columns
        const int bar_id = (exp_group ? 1 : 3) + hsel;
        const bool stager = tid < 64;                             // these threads fetch one L / D value per block
        auto fetch = [&](int item, int i) -> float {
            const int bh = item / args.n_kv;
            const int qi = i * kBlk + hsel * 64 + tid;
            return qi < args.Sq ? __ldg(vec + (size_t)bh * args.Sq + qi) : pad;
        };
        uint32_t g = 0;
        int it = 0;
        float next = (stager && (int)blockIdx.x < args.total_items) ? fetch(blockIdx.x, first_q(blockIdx.x)) : 0.0f;
        const bool rec = args.dbg != nullptr && lane == 0 && (warp == 4 || warp == 12);
        long long racc[5] = {0, 0, 0, 0, 0};          // tools only: kept in registers, written once at the end
        // Epilogue of a finished item: the exp warps store its dV rows, the dS warps its dK rows (32 columns per warp set).
        // It is DEFERRED into the first block of the next item, between that block's TMEM stores and its barrier arrival:
        // the accumulator is only overwritten by the MMA that this arrival releases, and the elementwise work of the
        // next item's first blocks no longer waits for the last MMA of this one (measured: the exp warps sat 7300 clk
        // per item in the epilogue wait and the issuer 750 clk per block on P_READY behind it).
        int prev_item = -1;
        auto epilogue = [&](int item, uint32_t parity) {
            const long long te = rec ? clock64() : 0;
            const int nt = item % args.n_kv;
            const int bh = item / args.n_kv;
            const int h = bh % args.H, b = bh / args.H;
            ptx::mbar_wait(bar(exp_group ? DV_DONE : DK_DONE), parity);
            ptx::tc_fence_after();
            const int t = nt * kBlk + tid;
            const bool live = t < args.Skv;
            const size_t tok = (size_t)b * args.Skv + (live ? t : 0);
            if (exp_group) store_acc_half(tm_dv + lane_off + 32 * hsel, args.dv + tok * args.lddv + (size_t)h * kD + 32 * hsel, live);
            else           store_acc_half(tm_dk + lane_off + 32 * hsel, args.dk + tok * args.lddk + (size_t)h * kD + 32 * hsel, live);
            ptx::tc_fence_before();
            if (!exp_group) ptx::mbar_arrive(bar(DK_FREE));
            if (rec) racc[4] += clock64() - te;
        };
        for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
            const int nt = item % args.n_kv;
            for (int i = first_q(item); i < n_q; ++i, ++g) {
                const uint32_t buf = g & 1u;
                long long t0 = rec ? clock64() : 0, t1;
#define BWD_STAMP(i) { if (rec) { t1 = clock64(); racc[i] += t1 - t0; t0 = t1; } }
                // stage this block's L (or D) — fetched one block ago — and prefetch the next block's
                if (stager) {
                    stage[buf * kBlk + tid] = next;
                    int ni = i + 1, nitem = item;
                    if (ni == n_q) { nitem = item + gridDim.x; ni = nitem < args.total_items ? first_q(nitem) : 0; }
                    if (nitem < args.total_items) next = fetch(nitem, ni);
                }
                bar_sync_n<128>(bar_id);
                BWD_STAMP(0)
                const float4* V4 = reinterpret_cast<const float4*>(stage + buf * kBlk) - hsel * 16;   // indexed by block column / 4
                if (exp_group) {
                    ptx::mbar_wait(bar(ST_FULL0 + buf), (g >> 1) & 1);
                    BWD_STAMP(1)
                    ptx::tc_fence_after();
                    const uint32_t s_tmem = tmem_base + lane_off + buf * kBlk;
#pragma unroll
                    for (int qt = q0; qt < q1; ++qt) {          // 32 columns: S^T in, [16 columns hi | 16 columns mid] of P^T out
                        float p[32];
                        ptx::tmem_ld_32x32(s_tmem + qt * 32, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 32; k += 4) {
                            const float4 l4 = V4[(qt * 32 + k) >> 2];
                            p[k]     = ptx::ex2(fmaf(p[k], c, -l4.x));
                            p[k + 1] = ptx::ex2(fmaf(p[k + 1], c, -l4.y));
                            p[k + 2] = ptx::ex2(fmaf(p[k + 2], c, -l4.z));
                            p[k + 3] = ptx::ex2(fmaf(p[k + 3], c, -l4.w));
                        }
                        if (CAUSAL && i == nt) {            // diagonal block: q position (column) before kv position (lane)
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (qt * 32 + k < tid) p[k] = 0.0f;
                        }
                        uint32_t hm[32];
#pragma unroll
                        for (int k = 0; k < 16; ++k) split_pack(p[2 * k], p[2 * k + 1], hm[k], hm[16 + k]);
                        ptx::tmem_st_32x32(s_tmem + qt * 32, hm);
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    if (prev_item >= 0) { epilogue(prev_item, (it - 1) & 1); prev_item = -1; }
                    ptx::mbar_arrive(bar(P_READY0 + buf));
                    BWD_STAMP(2)
                } else {
                    ptx::mbar_wait(bar(P_READY0 + buf), (g >> 1) & 1);
                    BWD_STAMP(1)
                    ptx::mbar_wait(bar(DPT_FULL_A + hsel), g & 1);
                    BWD_STAMP(2)
                    ptx::tc_fence_after();
                    const uint32_t p_tmem = tmem_base + lane_off + buf * kBlk;
#pragma unroll
                    for (int qt = q0; qt < q1; ++qt) {
                        uint32_t pm[32], dp[32];               // pm: [16 hi | 16 mid] of P^T, then of dS^T
                        ptx::tmem_ld_32x32(p_tmem + qt * 32, pm);
                        ptx::tmem_ld_32x32(tm_dpt + lane_off + qt * 32, dp);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 16; k += 2) {      // packed columns k, k+1 = elements 2k .. 2k+3 = one float4 of D
                            const float4 d4 = V4[(qt * 32 + 2 * k) >> 2];
                            const float s0 = unpack_lo(pm[k], pm[16 + k]) * ((__uint_as_float(dp[2 * k]) - d4.x) * scale);
                            const float s1 = unpack_hi(pm[k], pm[16 + k]) * ((__uint_as_float(dp[2 * k + 1]) - d4.y) * scale);
                            const float s2 = unpack_lo(pm[k + 1], pm[17 + k]) * ((__uint_as_float(dp[2 * k + 2]) - d4.z) * scale);
                            const float s3 = unpack_hi(pm[k + 1], pm[17 + k]) * ((__uint_as_float(dp[2 * k + 3]) - d4.w) * scale);
                            split_pack(s0, s1, pm[k], pm[16 + k]);
                            split_pack(s2, s3, pm[k + 1], pm[17 + k]);
                        }
                        ptx::tmem_st_32x32(tm_dpt + lane_off + qt * 32, pm);
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    if (prev_item >= 0) { epilogue(prev_item, (it - 1) & 1); prev_item = -1; }
                    ptx::mbar_arrive(bar(DS_READY_A + hsel));
                    BWD_STAMP(3)
                }
#undef BWD_STAMP
            }
            prev_item = item;
        }
        if (prev_item >= 0) epilogue(prev_item, (it - 1) & 1);
        if (rec) {
            long long* const dbg = args.dbg + blockIdx.x * 24;
            for (int k = 0; k < (exp_group ? 3 : 4); ++k) dbg[(exp_group ? 12 : 15) + k] = racc[k];
            dbg[exp_group ? 19 : 20] = racc[4];
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc(tmem_base, 512);
}

// =============================================================================== dQ
// smem: Q_R, dO_R (double buffered across items) | K_R, V_R, K_T (one kv block each) | barriers
constexpr int kDqSmem = 7 * kTileBytes + 1024 + 256;

template <bool CAUSAL, bool BX>
__global__ void __launch_bounds__(BX ? kThreadsBx : kThreads, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQr, const __grid_constant__ CUtensorMap tmDOr,
                   const __grid_constant__ CUtensorMap tmKr, const __grid_constant__ CUtensorMap tmKt,
                   const __grid_constant__ CUtensorMap tmVr, const BwdArgs args) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr  = ptx::smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr        = smem_raw + (base_addr - raw_addr);

    const uint32_t qr_addr = base_addr;                       // 2 x 32 KB
    const uint32_t dor_addr = qr_addr + 2 * kTileBytes;       // 2 x 32 KB
    const uint32_t kr_addr = dor_addr + 2 * kTileBytes, vr_addr = kr_addr + kTileBytes, kt_addr = vr_addr + kTileBytes;
    const uint32_t bar_addr = base_addr + 7 * kTileBytes;
    enum { QDO_FULL0 = 0, QDO_FULL1, QDO_EMPTY0, QDO_EMPTY1, KR_FULL, KR_EMPTY, VR_FULL, VR_EMPTY, KT_FULL, KT_EMPTY,
           S_FULL0, S_FULL1, P_READY0, P_READY1, DP_FULL, DS_READY, ACC_DONE, NBAR };
    auto bar = [&](int i) { return bar_addr + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 7 * kTileBytes + 8 * NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQr); ptx::prefetch_tensormap(&tmDOr); ptx::prefetch_tensormap(&tmKr);
        ptx::prefetch_tensormap(&tmKt); ptx::prefetch_tensormap(&tmVr);
    }
    if (warp == 3) {
        if (lane == 0) {
            for (int i = 0; i < NBAR; ++i)
                ptx::mbar_init(bar(i), (i == P_READY0 || i == P_READY1 || i == DS_READY) ? (BX ? 256 : 128) : 1);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();               // only shared memory / TMEM set-up above
    const uint32_t tm_dp = tmem_base + 256, tm_dq = tmem_base + 384;

    const int n_kv = args.n_kv;
    // kv blocks of an item (b, h, q tile mt): all, or 0..mt under the causal mask
    auto blocks_of = [&](int item) -> int { return CAUSAL ? (item % args.n_q) + 1 : n_kv; };
    uint32_t G = 0;
    for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) G += (uint32_t)blocks_of(item);

    if (warp == 0) {
        // ============ producer: Q_R + dO_R per item; K_R, V_R per kv block ============
        if (ptx::elect_one()) {
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
                const int mt = item % args.n_q;
                const int bh = item / args.n_q;
                const int h = bh % args.H, b = bh / args.H;
                const int qb = it & 1;
                ptx::mbar_wait(bar(QDO_EMPTY0 + qb), ((it >> 1) & 1) ^ 1u);
                ptx::mbar_arrive_expect_tx(bar(QDO_FULL0 + qb), 2 * kTileBytes);
                load_tile<BX>(qr_addr + qb * kTileBytes, &tmQr, bar(QDO_FULL0 + qb), mt * kBlk, h, b);
                load_tile<BX>(dor_addr + qb * kTileBytes, &tmDOr, bar(QDO_FULL0 + qb), mt * kBlk, h, b);
                const int nb = blocks_of(item);
                for (int j = 0; j < nb; ++j, ++g) {
                    ptx::mbar_wait(bar(KR_EMPTY), (g & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(KR_FULL), kTileBytes);
                    load_tile<BX>(kr_addr, &tmKr, bar(KR_FULL), j * kBlk, h, b);
                    ptx::mbar_wait(bar(VR_EMPTY), (g & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(VR_FULL), kTileBytes);
                    load_tile<BX>(vr_addr, &tmVr, bar(VR_FULL), j * kBlk, h, b);
                }
            }
        }
    } else if (warp == 1) {
        // ============ producer: K_T per kv block ============
        if (ptx::elect_one()) {
            uint32_t g = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) {
                const int bh = item / args.n_q;
                const int h = bh % args.H, b = bh / args.H;
                const int nb = blocks_of(item);
                for (int j = 0; j < nb; ++j, ++g) {
                    ptx::mbar_wait(bar(KT_EMPTY), (g & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(KT_FULL), kTileBytes);
                    load_tile<BX>(kt_addr, &tmKt, bar(KT_FULL), j * kBlk, h, b);
                }
            }
        }
    } else if (warp == 2) {
        // ============ MMA issuer ============
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_s  = BX ? ptx::umma_idesc_bf16(kBlk, kBlk, false, false) : ptx::umma_idesc_tf32(kBlk, kBlk, false, false);
            constexpr uint32_t idesc_ts = BX ? ptx::umma_idesc_bf16(kBlk, kD, false, true) : ptx::umma_idesc_tf32(kBlk, kD, false, true);
            struct Cur { int item, it, j, cnt; };
            auto cur_init = [&](Cur& c) {
                c.item = blockIdx.x; c.it = 0; c.j = 0;
                c.cnt = c.item < args.total_items ? blocks_of(c.item) : 0;
            };
            auto cur_next = [&](Cur& c) {
                if (++c.j == c.cnt) {
                    c.j = 0; c.item += gridDim.x; ++c.it;
                    c.cnt = c.item < args.total_items ? blocks_of(c.item) : 0;
                }
            };
            Cur c_s, c_dp, c_main;
            cur_init(c_s); cur_init(c_dp); cur_init(c_main);
            auto issue_s = [&](uint32_t g) {         // S(g) = Q K^T
                const int it = c_s.it, j = c_s.j;
                cur_next(c_s);
                if (j == 0) ptx::mbar_wait(bar(QDO_FULL0 + (it & 1)), (it >> 1) & 1);
                ptx::mbar_wait(bar(KR_FULL), g & 1);
                ptx::tc_fence_after();
                mma_rr<BX>(tmem_base + (g & 1u) * kBlk, qr_addr + (it & 1) * kTileBytes, kr_addr, idesc_s, args.t0);
                ptx::umma_commit(bar(KR_EMPTY));
                ptx::umma_commit(bar(S_FULL0 + (g & 1u)));
            };
            auto issue_dp = [&](uint32_t g) {        // dP(g) = dO V^T
                const int it = c_dp.it, j = c_dp.j;
                const bool last = j == c_dp.cnt - 1;
                cur_next(c_dp);
                ptx::mbar_wait(bar(VR_FULL), g & 1);
                ptx::tc_fence_after();
                mma_rr<BX>(tm_dp, dor_addr + (it & 1) * kTileBytes, vr_addr, idesc_s, args.t0);
                ptx::umma_commit(bar(VR_EMPTY));
                ptx::umma_commit(bar(DP_FULL));
                if (last) ptx::umma_commit(bar(QDO_EMPTY0 + (it & 1)));
            };
            if (G > 0) { issue_s(0); issue_dp(0); }
            if (G > 1) issue_s(1);
            for (uint32_t g = 0; g < G; ++g) {
                const int j = c_main.j;
                const bool last = j == c_main.cnt - 1;
                cur_next(c_main);
                ptx::mbar_wait(bar(DS_READY), g & 1);
                ptx::mbar_wait(bar(KT_FULL), g & 1);
                ptx::tc_fence_after();
                mma_ts<BX>(tm_dq, tm_dp, kt_addr, idesc_ts, j != 0, args.t0);                           // dQ += dS K
                ptx::umma_commit(bar(KT_EMPTY));
                if (last) ptx::umma_commit(bar(ACC_DONE));
                if (g + 1 < G) issue_dp(g + 1);
                if (g + 2 < G) issue_s(g + 2);
            }
        }
    } else if (warp >= 4) {
        // ============ warps 4-7: P = exp2(c S - L) in place.  warps 8-11: dS = P o (dP - D) / sqrt(dk) in place over
        // dP, then the dQ epilogue.  Thread = q row = TMEM lane; the groups run one block apart. ============
        constexpr int NSET = BX ? 2 : 1;               // BX: two warps per TMEM lane quarter, half of the columns each (see dkdv)
        const bool exp_group = warp < 4 + 4 * NSET;
        const int wq = warp & 3;
        const int hsel = BX ? ((warp - 4) >> 2) & 1 : 0;
        const int q0 = BX ? 2 * hsel : 0, q1 = BX ? 2 * hsel + 2 : 4;
        const int tid = wq * 32 + lane;
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const float c = args.c, scale = args.scale;
        const float* vec = exp_group ? args.lse : args.dsum;
        const float pad = exp_group ? INFINITY : 0.0f;
        auto fetch = [&](int item) -> float {
            const int sq = (item % args.n_q) * kBlk + tid;
            return sq < args.Sq ? __ldg(vec + (size_t)(item / args.n_q) * args.Sq + sq) : pad;
        };
        uint32_t g = 0;
        int it = 0;
        float next = (int)blockIdx.x < args.total_items ? fetch(blockIdx.x) : 0.0f;
        // dQ rows of a finished q tile (dS warps).  Deferred into the first block of the next item, between that block's
        // TMEM stores and its DS_READY arrival (which releases the MMA that overwrites the accumulator): the next item's
        // first dS no longer waits for the last dQ product of this one.
        int prev_item = -1;
        auto epilogue = [&](int item, uint32_t parity) {
            const int mt = item % args.n_q;
            const int bh = item / args.n_q;
            const int h = bh % args.H, b = bh / args.H;
            ptx::mbar_wait(bar(ACC_DONE), parity);
            ptx::tc_fence_after();
            const int sq = mt * kBlk + tid;
            const bool live = sq < args.Sq;
            float* drow = args.dq + ((size_t)b * args.Sq + (live ? sq : 0)) * args.lddq + (size_t)h * kD;
            if (BX) store_acc_half(tm_dq + lane_off + 32 * hsel, drow + 32 * hsel, live);      // two warp sets, 32 columns each
            else    store_acc_row(tm_dq + lane_off, drow, live);
            ptx::tc_fence_before();
        };
        for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
            const int mt = item % args.n_q;
            const float mine = next;                              // L (exp warps) or D (dS warps) of this thread's q row
            if (item + (int)gridDim.x < args.total_items) next = fetch(item + gridDim.x);
            const int nb = blocks_of(item);
            for (int j = 0; j < nb; ++j, ++g) {
                const uint32_t buf = g & 1u;
                const uint32_t s_tmem = tmem_base + lane_off + buf * kBlk;
                if (exp_group) {
                    ptx::mbar_wait(bar(S_FULL0 + buf), (g >> 1) & 1);
                    ptx::tc_fence_after();
                    const int kv_left = args.Skv - j * kBlk;
                    if (BX) {
#pragma unroll
                        for (int qt = q0; qt < q1; ++qt) {
                            float p[32];
                            ptx::tmem_ld_32x32(s_tmem + qt * 32, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                const float e = ptx::ex2(fmaf(p[k], c, -mine));
                                const bool masked = CAUSAL && j == mt && qt * 32 + k > tid;          // kv position after q position
                                p[k] = (qt * 32 + k < kv_left && !masked) ? e : 0.0f;   // zero-filled K rows past Skv
                            }
                            // P is no MMA operand in this kernel (only dS is): it goes back as plain fp32, no split
                            ptx::tmem_st_32x32(s_tmem + qt * 32, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                        }
                    } else
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        float p[64];
                        ptx::tmem_ld_32x32(s_tmem + hf * 64, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                        ptx::tmem_ld_32x32(s_tmem + hf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&p[32]));
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 64; ++k) {
                            const float e = __uint_as_float(rna_tf32(ptx::ex2(fmaf(p[k], c, -mine))));
                            const bool masked = CAUSAL && j == mt && hf * 64 + k > tid;          // kv position after q position
                            p[k] = (hf * 64 + k < kv_left && !masked) ? e : 0.0f;   // zero-filled K rows past Skv
                        }
                        ptx::tmem_st_32x32(s_tmem + hf * 64, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                        ptx::tmem_st_32x32(s_tmem + hf * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&p[32]));
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(bar(P_READY0 + buf));
                } else {
                    ptx::mbar_wait(bar(P_READY0 + buf), (g >> 1) & 1);
                    ptx::mbar_wait(bar(DP_FULL), g & 1);
                    ptx::tc_fence_after();
                    const float dscale = mine * scale;
                    if (BX) {
#pragma unroll
                        for (int qt = q0; qt < q1; ++qt) {
                            uint32_t pm[32], dp[32];               // pm: P (fp32) in, [16 hi | 16 mid] of dS out
                            ptx::tmem_ld_32x32(s_tmem + qt * 32, pm);
                            ptx::tmem_ld_32x32(tm_dp + lane_off + qt * 32, dp);
                            ptx::tmem_ld_wait();
                            uint32_t ds[32];
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
                                const float s0 = __uint_as_float(pm[2 * k]) * fmaf(__uint_as_float(dp[2 * k]), scale, -dscale);
                                const float s1 = __uint_as_float(pm[2 * k + 1]) * fmaf(__uint_as_float(dp[2 * k + 1]), scale, -dscale);
                                split_pack(s0, s1, ds[k], ds[16 + k]);
                            }
                            ptx::tmem_st_32x32(tm_dp + lane_off + qt * 32, ds);
                        }
                    } else
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t pp[32], dp[32];
                        ptx::tmem_ld_32x32(s_tmem + ch * 32, pp);
                        ptx::tmem_ld_32x32(tm_dp + lane_off + ch * 32, dp);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            dp[k] = rna_tf32(__uint_as_float(pp[k]) * fmaf(__uint_as_float(dp[k]), scale, -dscale));
                        ptx::tmem_st_32x32(tm_dp + lane_off + ch * 32, dp);
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    if (prev_item >= 0) { epilogue(prev_item, (it - 1) & 1); prev_item = -1; }
                    ptx::mbar_arrive(bar(DS_READY));
                }
            }
            if (!exp_group) prev_item = item;
        }
        if (prev_item >= 0) epilogue(prev_item, (it - 1) & 1);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc(tmem_base, 512);
}

// =============================================================================== dQ (split-bf16)
// The split-bf16 dQ kernel, restructured like attn_bwd_dkdv_bx_kernel: dP is produced and consumed as two 64-column halves
// (kv rows 0..63 / 64..127 of the block; per warp set the chain is dS -> 4 K16 steps of dQ -> dP half -> dS), K is loaded
// once per block into a ring of three (S runs two blocks ahead of dQ, and the bf16 image is the K-major operand of S and
// the MN-major operand of dQ at once), the dQ epilogue is deferred into the next item's first block.
// smem: Q x2 (per item) | dO | K ring x3 | V | barriers.
constexpr int kDqBxUsed = 7 * kTileBytes + 256;
constexpr int kDqBxSmem = 7 * kTileBytes + 1024 + 256;

template <bool CAUSAL>
__global__ void __launch_bounds__(kThreadsBx, 1)
attn_bwd_dq_bx_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const BwdArgs args) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr  = ptx::smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr        = smem_raw + (base_addr - raw_addr);

    const uint32_t q_addr = base_addr;                        // 2 x 32 KB (items alternate)
    const uint32_t do_addr = q_addr + 2 * kTileBytes;
    const uint32_t k_addr = do_addr + kTileBytes;             // ring of 3
    const uint32_t v_addr = k_addr + 3 * kTileBytes;
    const uint32_t bar_addr = base_addr + 7 * kTileBytes;
    enum { Q_FULL0 = 0, Q_FULL1, Q_EMPTY0, Q_EMPTY1, DO_FULL, DO_EMPTY, K_FULL0, K_FULL1, K_FULL2, K_EMPTY0, K_EMPTY1, K_EMPTY2,
           V_FULL, V_EMPTY, S_FULL0, S_FULL1, P_READY0, P_READY1, DP_FULL_A, DP_FULL_B, DS_READY_A, DS_READY_B, ACC_DONE,
           DQ_FREE, NBAR };
    static_assert(8 * NBAR + 4 <= 256, "barrier block");
    auto bar = [&](int i) { return bar_addr + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + 7 * kTileBytes + 8 * NBAR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQ); ptx::prefetch_tensormap(&tmDO); ptx::prefetch_tensormap(&tmK); ptx::prefetch_tensormap(&tmV);
    }
    if (warp == 3) {
        if (lane == 0) {
            for (int i = 0; i < NBAR; ++i)
                ptx::mbar_init(bar(i), (i == P_READY0 || i == P_READY1 || i == DQ_FREE) ? 256 : (i == DS_READY_A || i == DS_READY_B) ? 128 : 1);
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();               // only shared memory / TMEM set-up above
    const uint32_t tm_dp = tmem_base + 256, tm_dq = tmem_base + 384;

    const int n_kv = args.n_kv;
    auto blocks_of = [&](int item) -> int { return CAUSAL ? (item % args.n_q) + 1 : n_kv; };
    uint32_t G = 0;
    for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) G += (uint32_t)blocks_of(item);

    if (warp == 0) {
        // ============ producer of the streams that run ahead: Q per item, K per kv block (ring of 3) ============
        if (ptx::elect_one()) {
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
                const int mt = item % args.n_q;
                const int bh = item / args.n_q;
                const int h = bh % args.H, b = bh / args.H;
                const int qb = it & 1;
                ptx::mbar_wait(bar(Q_EMPTY0 + qb), ((it >> 1) & 1) ^ 1u);
                ptx::mbar_arrive_expect_tx(bar(Q_FULL0 + qb), kTileBytes);
                load_tile<true>(q_addr + qb * kTileBytes, &tmQ, bar(Q_FULL0 + qb), mt * kBlk, h, b);
                const int nb = blocks_of(item);
                for (int j = 0; j < nb; ++j, ++g) {
                    const uint32_t slot = g % 3u;
                    ptx::mbar_wait(bar(K_EMPTY0 + slot), ((g / 3u) & 1u) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(K_FULL0 + slot), kTileBytes);
                    load_tile<true>(k_addr + slot * kTileBytes, &tmK, bar(K_FULL0 + slot), j * kBlk, h, b);
                }
            }
        }
    } else if (warp == 1) {
        // ============ producer of the streams consumed one block ahead: dO per item, V per kv block ============
        if (ptx::elect_one()) {
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
                const int mt = item % args.n_q;
                const int bh = item / args.n_q;
                const int h = bh % args.H, b = bh / args.H;
                ptx::mbar_wait(bar(DO_EMPTY), (it & 1) ^ 1u);
                ptx::mbar_arrive_expect_tx(bar(DO_FULL), kTileBytes);
                load_tile<true>(do_addr, &tmDO, bar(DO_FULL), mt * kBlk, h, b);
                const int nb = blocks_of(item);
                for (int j = 0; j < nb; ++j, ++g) {
                    ptx::mbar_wait(bar(V_EMPTY), (g & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(bar(V_FULL), kTileBytes);
                    load_tile<true>(v_addr, &tmV, bar(V_FULL), j * kBlk, h, b);
                }
            }
        }
    } else if (warp == 2) {
        // ============ MMA issuer ============
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_s   = ptx::umma_idesc_bf16(kBlk, kBlk, false, false);
            constexpr uint32_t idesc_s64 = ptx::umma_idesc_bf16(kBlk, kBlk / 2, false, false);
            constexpr uint32_t idesc_ts  = ptx::umma_idesc_bf16(kBlk, kD, false, true);
            struct Cur { int item, it, j, cnt; };
            auto cur_init = [&](Cur& c) {
                c.item = blockIdx.x; c.it = 0; c.j = 0;
                c.cnt = c.item < args.total_items ? blocks_of(c.item) : 0;
            };
            auto cur_next = [&](Cur& c) {
                if (++c.j == c.cnt) {
                    c.j = 0; c.item += gridDim.x; ++c.it;
                    c.cnt = c.item < args.total_items ? blocks_of(c.item) : 0;
                }
            };
            Cur c_s, c_dp, c_main;
            cur_init(c_s); cur_init(c_dp); cur_init(c_main);
            long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            long long* const dbg = args.dbg;
            auto twait = [&](int slot, int which, uint32_t parity) {
                if (dbg) {
                    const long long t0 = clock64();
                    ptx::mbar_wait(bar(which), parity);
                    acc[slot] += clock64() - t0;
                } else {
                    ptx::mbar_wait(bar(which), parity);
                }
            };
            const long long t_begin = dbg ? clock64() : 0;
            auto issue_s = [&](uint32_t g) {         // S(g) = Q K^T
                const int it = c_s.it, j = c_s.j;
                const bool last = j == c_s.cnt - 1;
                cur_next(c_s);
                if (j == 0) twait(0, Q_FULL0 + (it & 1), (it >> 1) & 1);
                twait(1, K_FULL0 + g % 3u, (g / 3u) & 1u);
                ptx::tc_fence_after();
                mma_rr<true>(tmem_base + (g & 1u) * kBlk, q_addr + (it & 1) * kTileBytes, k_addr + (g % 3u) * kTileBytes, idesc_s, args.t0);
                ptx::umma_commit(bar(S_FULL0 + (g & 1u)));
                if (last) ptx::umma_commit(bar(Q_EMPTY0 + (it & 1)));
            };
            const bool halves = args.halves != 0;
            auto issue_dp = [&](uint32_t g, int half) {        // dP(g)[:, half] = dO V[64 kv rows of the half]^T
                const int it = c_dp.it, j = c_dp.j;
                const bool last = j == c_dp.cnt - 1;
                if (!halves) {             // one N = 128 product when called for half 1, nothing for half 0
                    if (half == 0) return;
                    cur_next(c_dp);
                    if (j == 0) twait(2, DO_FULL, it & 1);
                    twait(3, V_FULL, g & 1);
                    ptx::tc_fence_after();
                    mma_rr<true>(tm_dp, do_addr, v_addr, idesc_s, args.t0);
                    ptx::umma_commit(bar(DP_FULL_A));
                    ptx::umma_commit(bar(DP_FULL_B));
                    ptx::umma_commit(bar(V_EMPTY));
                    if (last) ptx::umma_commit(bar(DO_EMPTY));
                    return;
                }
                if (half == 0) {
                    if (j == 0) twait(2, DO_FULL, it & 1);
                    twait(3, V_FULL, g & 1);
                    ptx::tc_fence_after();
                } else {
                    cur_next(c_dp);
                }
                mma_rr<true>(tm_dp + half * 64, do_addr, v_addr + half * 8192, idesc_s64, args.t0);
                ptx::umma_commit(bar(DP_FULL_A + half));
                if (half == 1) {
                    ptx::umma_commit(bar(V_EMPTY));
                    if (last) ptx::umma_commit(bar(DO_EMPTY));
                }
            };
            if (G > 0) { issue_s(0); issue_dp(0, 0); issue_dp(0, 1); }
            if (G > 1) issue_s(1);
            for (uint32_t g = 0; g < G; ++g) {
                const int j = c_main.j, it_main = c_main.it;
                const bool last = j == c_main.cnt - 1;
                cur_next(c_main);
                const uint32_t ka = k_addr + (g % 3u) * kTileBytes;
                twait(4, DS_READY_A, g & 1);
                // the first dQ product of an item overwrites the accumulator: both warp sets must have stored the previous item's rows
                if (j == 0 && it_main > 0) twait(6, DQ_FREE, (it_main - 1) & 1);
                ptx::tc_fence_after();
                mma_ts_bx_range(tm_dq, tm_dp, ka, idesc_ts, j != 0, args.t0, 0, 4);                       // dQ += dS K, kv rows 0..63
                if (g + 1 < G) issue_dp(g + 1, 0);
                twait(5, DS_READY_B, g & 1);
                ptx::tc_fence_after();
                mma_ts_bx_range(tm_dq, tm_dp, ka, idesc_ts, true, args.t0, 4, 8);                         // kv rows 64..127
                ptx::umma_commit(bar(K_EMPTY0 + g % 3u));
                if (last) ptx::umma_commit(bar(ACC_DONE));
                if (g + 1 < G) issue_dp(g + 1, 1);
                if (g + 2 < G) issue_s(g + 2);
            }
            if (dbg) {
                acc[8] = clock64() - t_begin;
                acc[9] = G;
                for (int k = 0; k < 12; ++k) dbg[blockIdx.x * 24 + k] = acc[k];
            }
        }
    } else if (warp >= 4) {
        // ============ warps 4-11: P = exp2(c S - L) in place (fp32: P is no MMA operand here).  warps 12-19: dS = P o (dP - D) /
        // sqrt(dk), packed, in place over dP, then the dQ epilogue.  Thread = q row = TMEM lane; set A (hsel = 0) owns kv
        // columns 0..63 of every block, set B columns 64..127. ============
        const bool exp_group = warp < 12;
        const int wq = warp & 3;
        const int hsel = ((warp - 4) >> 2) & 1;
        const int q0 = 2 * hsel, q1 = 2 * hsel + 2;
        const int tid = wq * 32 + lane;
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const float c = args.c, scale = args.scale;
        const float* vec = exp_group ? args.lse : args.dsum;
        const float pad = exp_group ? INFINITY : 0.0f;
        auto fetch = [&](int item) -> float {
            const int sq = (item % args.n_q) * kBlk + tid;
            return sq < args.Sq ? __ldg(vec + (size_t)(item / args.n_q) * args.Sq + sq) : pad;
        };
        uint32_t g = 0;
        int it = 0;
        float next = (int)blockIdx.x < args.total_items ? fetch(blockIdx.x) : 0.0f;
        const bool rec = args.dbg != nullptr && lane == 0 && (warp == 4 || warp == 12);
        long long racc[5] = {0, 0, 0, 0, 0};
        int prev_item = -1;
        auto epilogue = [&](int item, uint32_t parity) {      // dQ rows of a finished q tile (dS warps), 32 columns per warp set
            const int mt = item % args.n_q;
            const int bh = item / args.n_q;
            const int h = bh % args.H, b = bh / args.H;
            ptx::mbar_wait(bar(ACC_DONE), parity);
            ptx::tc_fence_after();
            const int sq = mt * kBlk + tid;
            const bool live = sq < args.Sq;
            float* drow = args.dq + ((size_t)b * args.Sq + (live ? sq : 0)) * args.lddq + (size_t)h * kD;
            store_acc_half(tm_dq + lane_off + 32 * hsel, drow + 32 * hsel, live);
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar(DQ_FREE));
        };
        for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
            const int mt = item % args.n_q;
            const float mine = next;                              // L (exp warps) or D (dS warps) of this thread's q row
            if (item + (int)gridDim.x < args.total_items) next = fetch(item + gridDim.x);
            const int nb = blocks_of(item);
            for (int j = 0; j < nb; ++j, ++g) {
                const uint32_t buf = g & 1u;
                const uint32_t s_tmem = tmem_base + lane_off + buf * kBlk;
                long long t0 = rec ? clock64() : 0, t1;
#define BWD_STAMP(i) { if (rec) { t1 = clock64(); racc[i] += t1 - t0; t0 = t1; } }
                if (exp_group) {
                    ptx::mbar_wait(bar(S_FULL0 + buf), (g >> 1) & 1);
                    BWD_STAMP(0)
                    ptx::tc_fence_after();
                    const int kv_left = args.Skv - j * kBlk;
#pragma unroll
                    for (int qt = q0; qt < q1; ++qt) {
                        float p[32];
                        ptx::tmem_ld_32x32(s_tmem + qt * 32, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const float e = ptx::ex2(fmaf(p[k], c, -mine));
                            const bool masked = CAUSAL && j == mt && qt * 32 + k > tid;          // kv position after q position
                            p[k] = (qt * 32 + k < kv_left && !masked) ? e : 0.0f;   // zero-filled K rows past Skv
                        }
                        ptx::tmem_st_32x32(s_tmem + qt * 32, *reinterpret_cast<uint32_t(*)[32]>(&p[0]));
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(bar(P_READY0 + buf));
                    BWD_STAMP(1)
                } else {
                    ptx::mbar_wait(bar(P_READY0 + buf), (g >> 1) & 1);
                    BWD_STAMP(0)
                    ptx::mbar_wait(bar(DP_FULL_A + hsel), g & 1);
                    BWD_STAMP(1)
                    ptx::tc_fence_after();
                    const float dscale = mine * scale;
#pragma unroll
                    for (int qt = q0; qt < q1; ++qt) {
                        uint32_t pm[32], dp[32];               // pm: P (fp32) in, [16 hi | 16 mid] of dS out
                        ptx::tmem_ld_32x32(s_tmem + qt * 32, pm);
                        ptx::tmem_ld_32x32(tm_dp + lane_off + qt * 32, dp);
                        ptx::tmem_ld_wait();
                        uint32_t ds[32];
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            const float s0 = __uint_as_float(pm[2 * k]) * fmaf(__uint_as_float(dp[2 * k]), scale, -dscale);
                            const float s1 = __uint_as_float(pm[2 * k + 1]) * fmaf(__uint_as_float(dp[2 * k + 1]), scale, -dscale);
                            split_pack(s0, s1, ds[k], ds[16 + k]);
                        }
                        ptx::tmem_st_32x32(tm_dp + lane_off + qt * 32, ds);
                    }
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before();
                    if (prev_item >= 0) { epilogue(prev_item, (it - 1) & 1); prev_item = -1; }
                    ptx::mbar_arrive(bar(DS_READY_A + hsel));
                    BWD_STAMP(2)
                }
#undef BWD_STAMP
            }
            if (!exp_group) prev_item = item;
        }
        if (prev_item >= 0) epilogue(prev_item, (it - 1) & 1);
        if (rec) {
            long long* const dbg = args.dbg + blockIdx.x * 24;
            for (int k = 0; k < (exp_group ? 2 : 3); ++k) dbg[(exp_group ? 12 : 15) + k] = racc[k];
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc(tmem_base, 512);
}

// D[b,h,s] = sum_d dO[b,s,h,d] * O[b,s,h,d]: 16 lanes x float4 per (b,s,h) row.  A separate 18 us launch: computing D
// inside the dQ kernel (by its dS warps at item start, or by its exp warps one item ahead) was measured and costs the
// dQ kernel as much as or more than this launch (0.267 / 0.279 ms vs 0.265 ms for the whole backward at B8 H16 S1024).
// `planes` != nullptr (split-bf16 mode): dO is also written as bf16 hi / mid planes ([2][B,Sq,H,64]) for the backward
// kernels, and D uses the exact fp32 dO (hi + mid carries 16 bits, there is no common-mode rounding to cancel).
__global__ void __launch_bounds__(256) attn_dsum_kernel(const float* __restrict__ d_o, const float* __restrict__ o,
                                                        float* __restrict__ dsum, uint2* __restrict__ planes, int B, int H, int Sq) {
    pdl_trigger();
    pdl_wait();
    const int64_t rows = (int64_t)B * Sq * H;
    const int sub = threadIdx.x & 15;
    const int64_t step = ((int64_t)gridDim.x * blockDim.x) >> 4;
    // r0 is warp-uniform up to the +1 of the upper half-warp, so every lane runs the same trip count
    for (int64_t r0 = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) >> 4; r0 < rows; r0 += step) {
        const int64_t r = r0 + ((threadIdx.x >> 4) & 1);
        const bool ok = r < rows;
        float v = 0.0f;
        if (ok) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(d_o + r * kD) + sub);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(o + r * kD) + sub);
            if (planes != nullptr) {
                uint2 hi, mid;
                split_pack(a.x, a.y, hi.x, mid.x);
                split_pack(a.z, a.w, hi.y, mid.y);
                planes[r * (kD / 4) + sub] = hi;
                planes[(rows + r) * (kD / 4) + sub] = mid;
                v = a.x * bb.x + a.y * bb.y + a.z * bb.z + a.w * bb.w;
            } else {
                // dO rounded exactly as the tensor core will see it: D then equals sum_t P dP for the dP the
                // MMA computes, so the common-mode part of its TF32 error cancels in dS = P o (dP - D)
                v = __uint_as_float(cvt_tf32(a.x)) * bb.x + __uint_as_float(cvt_tf32(a.y)) * bb.y +
                    __uint_as_float(cvt_tf32(a.z)) * bb.z + __uint_as_float(cvt_tf32(a.w)) * bb.w;
            }
        }
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (ok && sub == 0) {
            const int h = (int)(r % H);
            const int64_t bs = r / H;
            const int s = (int)(bs % Sq), b = (int)(bs / Sq);
            dsum[((size_t)b * H + h) * Sq + s] = v;
        }
    }
}

// x [rows, HD] fp32 with row stride ld (a column block of a packed projection output) -> bf16 hi plane [rows, HD]
// followed by the mid plane [rows, HD]: hi = bf16_rn(x), mid = bf16_rn(x - hi).  8 B/elem, HBM bound.
__global__ void __launch_bounds__(256) attn_split_kernel(const float* __restrict__ x, int64_t ld, uint2* __restrict__ planes,
                                                         int64_t rows, int hd4) {
    pdl_trigger();
    pdl_wait();
    const int64_t n = rows * hd4;                        // float4 chunks
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / hd4;
        const int c = (int)(i - r * hd4);
        const float4 a = ld_stream(reinterpret_cast<const float4*>(x + r * ld) + c);
        uint2 hi, mid;
        split_pack(a.x, a.y, hi.x, mid.x);
        split_pack(a.z, a.w, hi.y, mid.y);
        planes[i] = hi;
        planes[n + i] = mid;
    }
}

// p[bh, s, t] = exp2(log2(e) * x[bh, s, t] - L[bh, s]) in place (x = scaled scores): debug view of P
__global__ void __launch_bounds__(256) attn_scores_from_lse_kernel(float* __restrict__ p, const float* __restrict__ lse,
                                                                   int64_t rows, int64_t cols) {
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = exp2f(fmaf(p[i], 1.4426950408889634f, -lse[i / cols]));
}

}  // namespace

// D [B,H,Sq]; split-bf16 mode: + the dO planes [2][B,Sq,H,64] bf16
size_t attn_bwd_scratch_bytes(int64_t B, int64_t H, int64_t Sq, bool bx) {
    const size_t d = ((size_t)B * H * Sq * sizeof(float) + 255) & ~(size_t)255;
    return bx ? d + (size_t)B * Sq * H * kD * 4 : d;
}

// split-bf16 planes of a [rows, HD] fp32 matrix (row stride ld): attn_split_kernel
int attn_split_launch(const float* x, int64_t ld, void* planes, int64_t rows, int64_t HD, cudaStream_t stream) {
    NPM_REQUIRE(aligned16(x) && aligned16(planes) && ld % 4 == 0 && HD % 4 == 0, "attention split: operands must be 16-byte aligned");
    const int64_t n = rows * (HD / 4);
    launch_pdl(attn_split_kernel, dim3(bw_grid((size_t)n, 256)), dim3(256), 0, stream, 1, x, ld, reinterpret_cast<uint2*>(planes), rows, (int)(HD / 4));
    count_launch();
    return check_launch("attn_split_kernel");
}

int attn_scores_from_lse_launch(float* p, const float* lse, int64_t rows, int64_t cols, cudaStream_t stream) {
    attn_scores_from_lse_kernel<<<bw_grid((size_t)(rows * cols), 256), 256, 0, stream>>>(p, lse, rows, cols);
    count_launch();
    return check_launch("attn_scores_from_lse_kernel");
}

// bx = false: q / k / v fp32 with token strides.  bx = true: q / k / v are the split-bf16 planes the forward pass
// saved ([2][B,S,H,64]); dO is split into `scratch` behind D by the dsum kernel.
int attn_bwd_launch(const void* q, const void* k, const void* v, const float* o, const float* d_o, const float* lse,
                    float* dq, float* dk, float* dv, float* dsum, int64_t B, int64_t H, int64_t Sq, int64_t Skv,
                    int64_t ldq, int64_t ldk, int64_t ldv, int64_t plq, int64_t plk, int64_t plv, int64_t lddq, int64_t lddk,
                    int64_t lddv, int causal, bool bx, int nterms, bool do_ready, cudaStream_t stream) {
    NPM_REQUIRE(!causal || Sq == Skv, "mha_core_bwd: the causal mask needs Sq == Skv");
    NPM_REQUIRE(o != nullptr || do_ready, "mha_core_bwd: the fused path needs the forward output o");
    NPM_REQUIRE(!do_ready || bx, "mha_core_bwd: do_ready is a split-bf16 feature");
    NPM_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o) && aligned16(d_o) && aligned16(dq) &&
                aligned16(dk) && aligned16(dv), "mha_core_bwd: pointers must be 16-byte aligned");
    const uint64_t HD = (uint64_t)H * kD;
    CUtensorMap tQr, tQt, tKr, tKt, tVr, tDOr, tDOt;
    int rc;
    for (int64_t ld : {lddq, lddk, lddv})
        NPM_REQUIRE(ld >= (int64_t)HD && ld % 4 == 0, "mha_core_bwd: token strides must be >= H*d and multiples of 4 floats");
    uint2* do_planes = nullptr;
    if (bx) {
        do_planes = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(dsum) + (((size_t)B * H * Sq * sizeof(float) + 255) & ~(size_t)255));
        NPM_REQUIRE(ldq >= (int64_t)HD && ldk >= (int64_t)HD && ldv >= (int64_t)HD && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 &&
                    plq % 8 == 0 && plk % 8 == 0 && plv % 8 == 0, "mha_core_bwd: plane strides must be multiples of 8 bf16 elements");
        if ((rc = make_tensor_map_bf16_planes(&tQr, q, Sq, H, B, ldq, plq, kBlk))) return rc;
        if ((rc = make_tensor_map_bf16_planes(&tKr, k, Skv, H, B, ldk, plk, kBlk))) return rc;
        if ((rc = make_tensor_map_bf16_planes(&tVr, v, Skv, H, B, ldv, plv, kBlk))) return rc;
        if ((rc = make_tensor_map_bf16_planes(&tDOr, do_planes, Sq, H, B, HD, (uint64_t)B * Sq * HD, kBlk))) return rc;
        tQt = tQr; tKt = tKr; tDOt = tDOr;          // one bf16 image serves as the R and the T operand
    } else {
    for (int64_t ld : {ldq, ldk, ldv})
        NPM_REQUIRE(ld >= (int64_t)HD && ld % 4 == 0, "mha_core_bwd: token strides must be >= H*d and multiples of 4 floats");
    const float* qf = reinterpret_cast<const float*>(q);
    const float* kf = reinterpret_cast<const float*>(k);
    const float* vf = reinterpret_cast<const float*>(v);
    if ((rc = make_tensor_map_4d(&tQr, qf, kD, Sq, H, B, ldq, kD, Sq * ldq, 32, kBlk, true, false))) return rc;
    if ((rc = make_tensor_map_4d(&tQt, qf, kD, Sq, H, B, ldq, kD, Sq * ldq, 32, kBlk, true, true))) return rc;
    if ((rc = make_tensor_map_4d(&tKr, kf, kD, Skv, H, B, ldk, kD, Skv * ldk, 32, kBlk, true, false))) return rc;
    if ((rc = make_tensor_map_4d(&tKt, kf, kD, Skv, H, B, ldk, kD, Skv * ldk, 32, kBlk, true, true))) return rc;
    if ((rc = make_tensor_map_4d(&tVr, vf, kD, Skv, H, B, ldv, kD, Skv * ldv, 32, kBlk, true, false))) return rc;
    if ((rc = make_tensor_map_4d(&tDOr, d_o, kD, Sq, H, B, HD, kD, Sq * HD, 32, kBlk, true, false))) return rc;
    if ((rc = make_tensor_map_4d(&tDOt, d_o, kD, Sq, H, B, HD, kD, Sq * HD, 32, kBlk, true, true))) return rc;
    }

    BwdArgs a;
    a.B = (int)B; a.H = (int)H; a.Sq = (int)Sq; a.Skv = (int)Skv;
    a.n_q = (int)((Sq + kBlk - 1) / kBlk);
    a.n_kv = (int)((Skv + kBlk - 1) / kBlk);
    a.c = (float)(1.4426950408889634 / sqrt((double)kD));
    a.scale = (float)(1.0 / sqrt((double)kD));
    a.lse = lse; a.dsum = dsum; a.dq = dq; a.dk = dk; a.dv = dv;
    a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
    a.causal = causal ? 1 : 0;
    a.t0 = nterms == 1 ? 2 : 0;
    static const int debug_skip_env = getenv("NPM_ATTN_DEBUG_SKIP") ? atoi(getenv("NPM_ATTN_DEBUG_SKIP")) : 0;
    a.debug_skip = debug_skip_env;
    static const int halves_env = getenv("NPM_ATTN_HALVES") ? atoi(getenv("NPM_ATTN_HALVES")) : 3;   // bit 0: dK/dV kernel, bit 1: dQ kernel
    const int halves_mask = halves_env;
    a.dbg = nullptr;
    static const bool dbg_times = getenv("NPM_ATTN_DEBUG_TIMES") != nullptr;       // tools only: synchronises and prints
    if (dbg_times) {
        cudaMalloc(&a.dbg, sizeof(long long) * 24 * num_sms());
        cudaMemset(a.dbg, 0, sizeof(long long) * 24 * num_sms());
    }

    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaSuccess;
        auto set = [&](auto kern, int bytes) { if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); };
        set(attn_bwd_dkdv_kernel<false, false>, kKvSmem); set(attn_bwd_dkdv_kernel<true, false>, kKvSmem);
        set(attn_bwd_dkdv_kernel<false, true>, kKvSmem);  set(attn_bwd_dkdv_kernel<true, true>, kKvSmem);
        set(attn_bwd_dkdv_bx_kernel<false>, kKvBxSmem);   set(attn_bwd_dkdv_bx_kernel<true>, kKvBxSmem);
        set(attn_bwd_dq_bx_kernel<false>, kDqBxSmem);     set(attn_bwd_dq_bx_kernel<true>, kDqBxSmem);
        set(attn_bwd_dq_kernel<false, false>, kDqSmem);   set(attn_bwd_dq_kernel<true, false>, kDqSmem);
        set(attn_bwd_dq_kernel<false, true>, kDqSmem);    set(attn_bwd_dq_kernel<true, true>, kDqSmem);
        if (e != cudaSuccess) { set_error("attn_bwd smem attribute: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
        configured = true;
    }
    if (!do_ready) {     // else: the output projection's dX GEMM already wrote D and the dO planes (npm_gemm_desc.rowdot_*)
        const int64_t rows = B * Sq * H;
        int grid = (int)((rows * 16 + 255) / 256);
        const int cap = num_sms() * 8;
        if (grid > cap) grid = cap;
        launch_pdl(attn_dsum_kernel, dim3(grid), dim3(256), 0, stream, 1, d_o, o, dsum, do_planes, (int)B, (int)H, (int)Sq);
        count_launch();
        if ((rc = check_launch("attn_dsum_kernel"))) return rc;
    }
    {
        const int64_t items = B * H * a.n_kv;
        NPM_REQUIRE(items < (1ll << 30), "mha_core_bwd: too many tiles");
        a.total_items = (int)items;
        int grid = (int)(items < num_sms() ? items : num_sms());
        if (causal) while (grid > 1 && gcd_int(grid, a.n_kv) != 1) --grid;     // see attn_fwd_launch
        static const bool old_bx = getenv("NPM_ATTN_OLD_DKDV") != nullptr;      // A/B: the shared-template split-bf16 kernel
        a.halves = halves_mask & 1;
        if (bx && !old_bx) {
            auto kern = causal ? attn_bwd_dkdv_bx_kernel<true> : attn_bwd_dkdv_bx_kernel<false>;
            launch_pdl(kern, dim3(grid), dim3(kThreadsBx), kKvBxSmem, stream, 1, tQr, tKr, tVr, tDOr, a);
        } else {
            auto kern = causal ? (bx ? attn_bwd_dkdv_kernel<true, true> : attn_bwd_dkdv_kernel<true, false>)
                               : (bx ? attn_bwd_dkdv_kernel<false, true> : attn_bwd_dkdv_kernel<false, false>);
            launch_pdl(kern, dim3(grid), dim3(bx ? kThreadsBx : kThreads), kKvSmem, stream, 1, tQr, tQt, tKr, tVr, tDOr, tDOt, a);
        }
        count_launch();
        if ((rc = check_launch("attn_bwd_dkdv_kernel"))) return rc;
        if (dbg_times) {
            cudaStreamSynchronize(stream);
            std::vector<long long> h(24 * (size_t)grid);
            cudaMemcpy(h.data(), a.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
            static const char* names_old[10] = {"K_FULL", "QR_FULL", "V_FULL", "DOR_FULL", "P_READY", "DOT_FULL", "DS_READY",
                                                "QT_FULL", "total", "blocks"};
            static const char* names_bx[10] = {"K_FULL", "Q_FULL", "V_FULL", "DO_FULL", "P_READY", "DS_READY_A", "DS_READY_B",
                                               "DK_FREE", "total", "blocks"};
            const char** names = (bx && !old_bx) ? names_bx : names_old;
            double sum[24] = {0};
            for (int c = 0; c < grid; ++c) for (int k = 0; k < 24; ++k) sum[k] += (double)h[24 * c + k];
            const double blocks = sum[9] > 0 ? sum[9] : 1;
            fprintf(stderr, "[attn_bwd_dkdv issuer, cycles per q block, mean over %d CTAs]", grid);
            for (int k = 0; k < 9; ++k) fprintf(stderr, " %s=%.0f", names[k], sum[k] / blocks);
            fprintf(stderr, "\n  elementwise warps, cycles per q block: exp(w4) stage+bar %.0f, wait S^T %.0f, work %.0f | dS(first warp) stage+bar %.0f, "
                    "wait P %.0f, wait dP^T %.0f, work %.0f | epilogues per item: dV %.0f dK %.0f\n",
                    sum[12] / blocks, sum[13] / blocks, sum[14] / blocks, sum[15] / blocks, sum[16] / blocks, sum[17] / blocks, sum[18] / blocks,
                    sum[19] * 8 / blocks, sum[20] * 8 / blocks);
            cudaFree(a.dbg);
            a.dbg = nullptr;
        }
    }
    {
        const int64_t items = B * H * a.n_q;
        NPM_REQUIRE(items < (1ll << 30), "mha_core_bwd: too many tiles");
        a.total_items = (int)items;
        int grid = (int)(items < num_sms() ? items : num_sms());
        if (causal) while (grid > 1 && gcd_int(grid, a.n_q) != 1) --grid;
        static const bool old_dq = getenv("NPM_ATTN_OLD_DQ") != nullptr;      // A/B: the shared-template split-bf16 kernel
        if (dbg_times && bx && !old_dq) {
            cudaMalloc(&a.dbg, sizeof(long long) * 24 * num_sms());
            cudaMemset(a.dbg, 0, sizeof(long long) * 24 * num_sms());
        }
        a.halves = (halves_mask >> 1) & 1;
        if (bx && !old_dq) {
            auto kern = causal ? attn_bwd_dq_bx_kernel<true> : attn_bwd_dq_bx_kernel<false>;
            launch_pdl(kern, dim3(grid), dim3(kThreadsBx), kDqBxSmem, stream, 1, tQr, tDOr, tKr, tVr, a);
        } else {
            auto kern = causal ? (bx ? attn_bwd_dq_kernel<true, true> : attn_bwd_dq_kernel<true, false>)
                               : (bx ? attn_bwd_dq_kernel<false, true> : attn_bwd_dq_kernel<false, false>);
            launch_pdl(kern, dim3(grid), dim3(bx ? kThreadsBx : kThreads), kDqSmem, stream, 1, tQr, tDOr, tKr, tKt, tVr, a);
        }
        count_launch();
        if ((rc = check_launch("attn_bwd_dq_kernel"))) return rc;
        if (dbg_times && a.dbg != nullptr) {
            cudaStreamSynchronize(stream);
            std::vector<long long> h(24 * (size_t)grid);
            cudaMemcpy(h.data(), a.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
            double sum[24] = {0};
            for (int c = 0; c < grid; ++c) for (int k = 0; k < 24; ++k) sum[k] += (double)h[24 * c + k];
            const double blocks = sum[9] > 0 ? sum[9] : 1;
            fprintf(stderr, "[attn_bwd_dq issuer, cycles per kv block] Q_FULL=%.0f K_FULL=%.0f DO_FULL=%.0f V_FULL=%.0f DS_READY_A=%.0f DS_READY_B=%.0f "
                    "DQ_FREE=%.0f total=%.0f\n  exp(w4): wait S %.0f, work %.0f | dS(first warp): wait P %.0f, wait dP %.0f, work %.0f\n",
                    sum[0] / blocks, sum[1] / blocks, sum[2] / blocks, sum[3] / blocks, sum[4] / blocks, sum[5] / blocks, sum[6] / blocks,
                    sum[8] / blocks, sum[12] / blocks, sum[13] / blocks, sum[15] / blocks, sum[16] / blocks, sum[17] / blocks);
            cudaFree(a.dbg);
            a.dbg = nullptr;
        }
    }
    return NPM_OK;
}

}  // namespace npm
