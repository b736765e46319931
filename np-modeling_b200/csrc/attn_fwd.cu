// attn_fwd.cu — fused attention core forward on tcgen05 / TMEM / TMA (TF32, head dim 64).
//
//   o[b,s,h,:] = softmax_t( q[b,s,h,:] . k[b,t,h,:] / sqrt(dk) ) v[b,t,h,:]
//
// replaces the reference's materialised  einsum → Softmax → einsum  chain
// (layers/attentions.py:103-112): the [B,H,Sq,Skv] score tensor never exists in HBM; only the
// per-row log-sum-exp is saved for the backward pass.
//
// One persistent CTA per SM; a work item is (batch, head, 128 query rows).  Per 128-row KV block:
//   warp 0   TMA producer: Q tile (double buffered across items) and the K ring
//   warp 1   TMA producer: the V ring (landed with 32-byte swizzle atoms = an MN-major B operand)
//   warp 2   MMA issuer  : S = Q K^T          (tcgen05.mma kind::tf32, A/B from smem, D in TMEM)
//                          O += P V           (A = P read straight from TMEM, B = V from smem)
//   warps 4-7 softmax    : thread t owns query row t = TMEM lane t: tcgen05.ld the S row, online
//                          max / exp2 / sum, write P back over S with tcgen05.st.  The running
//                          maximum is only raised when it grew by more than 2^8 (the O accumulator
//                          is then rescaled in TMEM), so the rescale is off the critical path.
// S/P is double buffered in TMEM so that S(j+1) is computed while the softmax of block j runs.
//
// BX = true is the split-bf16 ("bf16x3") variant of the same kernel: q / k / v arrive pre-split as bf16 hi + mid planes
// ([2][B,S,H,64], written by attn_split_kernel in attn_bwd.cu), a tile is its hi image (16 KB) followed by its mid image,
// every product runs as mid*hi + hi*mid + hi*hi with kind::f16 (K = 16), and P goes back to TMEM as packed bf16
// pairs, per 64-column half [32 columns hi | 32 columns mid].  In bf16 a [rows, 64] tile with the 128-byte swizzle is
// at once a K-major operand over its 64 columns and an MN-major operand over its rows, so V needs no second layout.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace npm {

int make_tensor_map_4d(CUtensorMap* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                       uint64_t s1, uint64_t s2, uint64_t s3, uint32_t b0, uint32_t b1, bool round_tf32,
                       bool atom32b);
int make_tensor_map_bf16_planes(CUtensorMap* tm, const void* base, uint64_t S, uint64_t H, uint64_t B, uint64_t ld,
                                uint64_t plane_elems, uint32_t box_rows);

namespace {

constexpr int kBM = 128;          // query rows per work item
constexpr int kBN = 128;          // kv rows per block
constexpr int kD  = 64;           // head dim (dk = dv)
constexpr int kTileBytes = kBM * kD * 4;     // 32 KB: two {32 d, 128 rows} boxes of 16 KB
constexpr int kChunkBytes = 16384;
constexpr int kKS = 3, kVS = 2;              // K / V ring depth (K is consumed one block ahead of V)
constexpr int kSmemBytes = (2 + kKS + kVS) * kTileBytes + 1024 /*align*/ + 512 /*barriers*/;
constexpr int kThreads = 256;
constexpr int kThreadsBx = 384;               // split-bf16 variant: 8 softmax warps (two per TMEM lane quarter, 64 columns each)
constexpr float kRescaleThreshold = 8.0f;    // log2 units

struct FwdArgs {
    int B, H, Sq, Skv;
    int q_tiles, n_kv, total_items;
    int t0;                  // split-bf16 variant: first of the three terms (mid*hi, hi*mid, hi*hi) to run: 0 = bf16x3, 2 = plain bf16 (hi*hi only)
    int causal;              // scores of kv position t > query position s are masked (Sq == Skv): kv blocks past the
                             // diagonal are never visited, the diagonal block is masked element-wise
    float c;                 // log2(e) / sqrt(dk)
    float* o;                // [B, Sq, H, 64]
    float* lse;              // [B, H, Sq]  log2-domain: m + log2(sum)
    long long* dbg;          // tools only (NPM_ATTN_DEBUG_TIMES): 16 cycle counters per CTA
};

// Round-to-nearest (ties away) fp32 → tf32 on the integer ALU: kind::tf32 reads only the upper 19 bits
// of each operand, so adding half a tf32 ulp to the bit pattern is the whole rounding.  (cvt.rna.tf32
// issues on the XU pipe, which the ex2 of the softmax already saturates.)
__device__ __forceinline__ uint32_t rna_tf32(float x) { return __float_as_uint(x) + 0x1000u; }

// p0, p1 (consecutive k) -> packed bf16 hi pair and mid pair
__device__ __forceinline__ void split_pack(float p0, float p1, uint32_t& hi, uint32_t& mid) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(p1), "f"(p0));
    const float r0 = p0 - __uint_as_float(hi << 16), r1 = p1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(r1), "f"(r0));
}

// one 32-bit TMEM cell per lane: how the two softmax warps of a lane quarter exchange row maxima / row sums
__device__ __forceinline__ void tmem_st_x1(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_x1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ void bar_sync_64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <bool CAUSAL, bool BX>
__global__ void __launch_bounds__(BX ? kThreadsBx : kThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const FwdArgs args) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr  = ptx::smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr        = smem_raw + (base_addr - raw_addr);

    const uint32_t q_addr = base_addr;                              // 2 x 32 KB
    const uint32_t k_addr = q_addr + 2 * kTileBytes;                // kKS x 32 KB
    const uint32_t v_addr = k_addr + kKS * kTileBytes;              // kVS x 32 KB
    const uint32_t bar_addr = v_addr + kVS * kTileBytes;
    auto q_full   = [&](int i) { return bar_addr + 8u * i; };
    auto q_empty  = [&](int i) { return bar_addr + 8u * (2 + i); };
    auto k_full   = [&](int i) { return bar_addr + 8u * (4 + i); };
    auto k_empty  = [&](int i) { return bar_addr + 8u * (4 + kKS + i); };
    auto v_full   = [&](int i) { return bar_addr + 8u * (4 + 2 * kKS + i); };
    auto v_empty  = [&](int i) { return bar_addr + 8u * (4 + 2 * kKS + kVS + i); };
    const uint32_t misc = bar_addr + 8u * (4 + 2 * kKS + 2 * kVS);
    auto s_full   = [&](int i) { return misc + 8u * i; };
    auto p_ready  = [&](int i) { return misc + 8u * (2 + i); };
    const uint32_t o_done = misc + 8u * 4;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(
        base_ptr + (2 + kKS + kVS) * kTileBytes + 8 * (4 + 2 * kKS + 2 * kVS + 5));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmQ);
        ptx::prefetch_tensormap(&tmK);
        ptx::prefetch_tensormap(&tmV);
    }
    if (warp == 3) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) { ptx::mbar_init(q_full(i), 1); ptx::mbar_init(q_empty(i), 1); }
            for (int i = 0; i < kKS; ++i) { ptx::mbar_init(k_full(i), 1); ptx::mbar_init(k_empty(i), 1); }
            for (int i = 0; i < kVS; ++i) { ptx::mbar_init(v_full(i), 1); ptx::mbar_init(v_empty(i), 1); }
            for (int i = 0; i < 2; ++i) { ptx::mbar_init(s_full(i), 1); ptx::mbar_init(p_ready(i), BX ? 256 : 128); }
            ptx::mbar_init(o_done, 1);
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
    const uint32_t tmem_o = tmem_base + 256;          // S/P buffers at columns [0,128) and [128,256)

    const int n_kv = args.n_kv;
    // kv blocks item (b, h, q tile mt) walks: all of them, or 0..mt under the causal mask (kBM == kBN, Sq == Skv)
    auto blocks_of = [&](int item) -> int { return CAUSAL ? (item % args.q_tiles) + 1 : n_kv; };

    if (warp == 0) {
        // ===================== Q + K producer =====================
        if (ptx::elect_one()) {
            uint32_t gk = 0;
            int it = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x, ++it) {
                const int mt = item % args.q_tiles;
                const int bh = item / args.q_tiles;
                const int h = bh % args.H, b = bh / args.H;
                const int qb = it & 1;
                ptx::mbar_wait(q_empty(qb), ((it >> 1) & 1) ^ 1u);
                ptx::mbar_arrive_expect_tx(q_full(qb), kTileBytes);
                if (BX) {
                    ptx::tma_load_5d(q_addr + qb * kTileBytes, &tmQ, q_full(qb), 0, mt * kBM, h, b, 0);
                    ptx::tma_load_5d(q_addr + qb * kTileBytes + kChunkBytes, &tmQ, q_full(qb), 0, mt * kBM, h, b, 1);
                } else {
                    ptx::tma_load_4d(q_addr + qb * kTileBytes, &tmQ, q_full(qb), 0, mt * kBM, h, b);
                    ptx::tma_load_4d(q_addr + qb * kTileBytes + kChunkBytes, &tmQ, q_full(qb), 32, mt * kBM, h, b);
                }
                const int nb = blocks_of(item);
                for (int j = 0; j < nb; ++j, ++gk) {
                    const int st = gk % kKS;
                    ptx::mbar_wait(k_empty(st), ((gk / kKS) & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(k_full(st), kTileBytes);
                    if (BX) {
                        ptx::tma_load_5d(k_addr + st * kTileBytes, &tmK, k_full(st), 0, j * kBN, h, b, 0);
                        ptx::tma_load_5d(k_addr + st * kTileBytes + kChunkBytes, &tmK, k_full(st), 0, j * kBN, h, b, 1);
                    } else {
                        ptx::tma_load_4d(k_addr + st * kTileBytes, &tmK, k_full(st), 0, j * kBN, h, b);
                        ptx::tma_load_4d(k_addr + st * kTileBytes + kChunkBytes, &tmK, k_full(st), 32, j * kBN, h, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= V producer =======================
        if (ptx::elect_one()) {
            uint32_t gv = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) {
                const int bh = item / args.q_tiles;
                const int h = bh % args.H, b = bh / args.H;
                const int nb = blocks_of(item);
                for (int j = 0; j < nb; ++j, ++gv) {
                    const int st = gv % kVS;
                    ptx::mbar_wait(v_empty(st), ((gv / kVS) & 1) ^ 1u);
                    ptx::mbar_arrive_expect_tx(v_full(st), kTileBytes);
                    if (BX) {
                        ptx::tma_load_5d(v_addr + st * kTileBytes, &tmV, v_full(st), 0, j * kBN, h, b, 0);
                        ptx::tma_load_5d(v_addr + st * kTileBytes + kChunkBytes, &tmV, v_full(st), 0, j * kBN, h, b, 1);
                    } else {
                        ptx::tma_load_4d(v_addr + st * kTileBytes, &tmV, v_full(st), 0, j * kBN, h, b);
                        ptx::tma_load_4d(v_addr + st * kTileBytes + kChunkBytes, &tmV, v_full(st), 32, j * kBN, h, b);
                    }
                }
            }
        }
    } else if (warp == 2) {
        // ======================= MMA issuer =======================
        if (ptx::elect_one()) {
            constexpr uint32_t idesc_s  = BX ? ptx::umma_idesc_bf16(kBM, kBN, false, false) : ptx::umma_idesc_tf32(kBM, kBN, false, false);
            constexpr uint32_t idesc_pv = BX ? ptx::umma_idesc_bf16(kBM, kD, false, true) : ptx::umma_idesc_tf32(kBM, kD, false, true);
            const uint64_t desc_k  = ptx::umma_desc_base(2 /*SWIZZLE_128B*/, 16, 1024);
            // MN-major B operand.  tf32: 32-byte swizzle atoms, 32-element chunks 16 KB apart, 4-row K groups.  bf16: the
            // plain 128-byte swizzle, N = 64 is one 128-byte chunk (LBO unused), 8-row K groups 1024 B apart.
            const uint64_t desc_mn = BX ? ptx::umma_desc_base(2 /*SWIZZLE_128B*/, kChunkBytes, 1024)
                                        : ptx::umma_desc_base(1 /*SWIZZLE_128B_BASE32B*/, kChunkBytes, 512);
            // (item, kv block) sequence of this CTA; the S products run one block ahead of the PV products, so two cursors
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
            uint32_t total_blocks = 0;
            for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) total_blocks += (uint32_t)blocks_of(item);

            // S(gs): the score block of sequence index gs = cursor cs
            auto issue_s = [&](uint32_t gs, const Cur& cs) {
                const int it = cs.it, j = cs.j;
                const int qb = it & 1;
                if (j == 0) ptx::mbar_wait(q_full(qb), (it >> 1) & 1);
                const int st = gs % kKS;
                const long long tk = args.dbg ? clock64() : 0;
                ptx::mbar_wait(k_full(st), (gs / kKS) & 1);
                if (args.dbg) args.dbg[blockIdx.x * 16 + 0] += clock64() - tk;
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (gs & 1u) * kBN;
                const uint32_t qa = q_addr + qb * kTileBytes, ka = k_addr + st * kTileBytes;
                if (BX) {
                    // (A image, B image): mid*hi, hi*mid, hi*hi; the mid image of a tile follows its hi image
                    const uint64_t da0 = ptx::umma_desc(desc_k, qa), db0 = ptx::umma_desc(desc_k, ka);
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        if (t < args.t0) continue;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t da = ptx::umma_desc_off(da0, (t == 0 ? kChunkBytes : 0) + kk * 32);
                            const uint64_t db = ptx::umma_desc_off(db0, (t == 1 ? kChunkBytes : 0) + kk * 32);
                            ptx::umma_f16(d_tmem, da, db, idesc_s, (t > args.t0 || kk != 0) ? 1u : 0u);
                        }
                    }
                } else {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t da = ptx::umma_desc(desc_k, qa + kb * kChunkBytes + kk * 32);
                            const uint64_t db = ptx::umma_desc(desc_k, ka + kb * kChunkBytes + kk * 32);
                            ptx::umma_tf32(d_tmem, da, db, idesc_s, (kb | kk) != 0 ? 1u : 0u);
                        }
                }
                ptx::umma_commit(k_empty(st));
                ptx::umma_commit(s_full(gs & 1u));
                if (j == cs.cnt - 1) ptx::umma_commit(q_empty(qb));
            };

            Cur cs, cm;
            cur_init(cs);
            cur_init(cm);
            if (total_blocks > 0) { issue_s(0, cs); cur_next(cs); }
            for (uint32_t g = 0; g < total_blocks; ++g) {
                if (g + 1 < total_blocks) { issue_s(g + 1, cs); cur_next(cs); }
                const int j = cm.j;
                cur_next(cm);
                const int st = g % kVS;
                const long long tp = args.dbg ? clock64() : 0;
                ptx::mbar_wait(p_ready(g & 1u), (g >> 1) & 1);
                const long long tv = args.dbg ? clock64() : 0;
                ptx::mbar_wait(v_full(st), (g / kVS) & 1);
                if (args.dbg) {
                    args.dbg[blockIdx.x * 16 + 1] += tv - tp;
                    args.dbg[blockIdx.x * 16 + 2] += clock64() - tv;
                    args.dbg[blockIdx.x * 16 + 3] += 1;
                }
                ptx::tc_fence_after();
                const uint32_t p_tmem = tmem_base + (g & 1u) * kBN;
                const uint32_t va = v_addr + st * kTileBytes;
                if (BX) {
                    // P in TMEM: per 64-k half [32 columns hi | 32 columns mid], two bf16 per column; a K16 step is 8 columns
                    // of P and 16 rows (2048 B) of the V image
                    const uint64_t db0 = ptx::umma_desc(desc_mn, va);
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        if (t < args.t0) continue;
#pragma unroll
                        for (int kk = 0; kk < kBN / 16; ++kk) {
                            const uint32_t pa = p_tmem + (kk >> 2) * 64 + (t == 0 ? 32 : 0) + (kk & 3) * 8;
                            const uint64_t db = ptx::umma_desc_off(db0, (t == 1 ? kChunkBytes : 0) + kk * 2048);
                            ptx::umma_f16_ts(tmem_o, pa, db, idesc_pv, (j != 0 || t > args.t0 || kk != 0) ? 1u : 0u);
                        }
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < kBN / 8; ++kk) {
                        const uint64_t db = ptx::umma_desc(desc_mn, va + kk * 1024);
                        ptx::umma_tf32_ts(tmem_o, p_tmem + kk * 8, db, idesc_pv, (j | kk) != 0 ? 1u : 0u);
                    }
                }
                ptx::umma_commit(v_empty(st));
                ptx::umma_commit(o_done);
            }
        }
    } else if (BX && warp >= 4) {
        // ===== softmax, split-bf16 variant: TWO warps per TMEM lane quarter (warps 4+wq and 8+wq), 64 of the 128 columns
        // each.  The split to bf16 hi / mid doubles the per-element work of these warps (cvt.rn.bf16x2 shares the XU pipe
        // with ex2) and with four warps they, not the tensor pipe, bounded the kernel.  The two threads of a row agree on
        // the running maximum through one TMEM cell each per block (columns 320.. of their own lane) and add their row
        // sums the same way in the epilogue. =====
        const int wq = warp & 3, hsel = ((warp - 4) >> 2) & 1;
        const int row = wq * 32 + lane;
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const uint32_t tmem_x = tmem_base + lane_off + 320;
        const int c0 = hsel * 64;
        const float c = args.c;
        uint32_t g = 0;
        // Epilogue of a finished item (add the two partial row sums, O / l -> global, log-sum-exp -> saved).  DEFERRED into
        // the first kv block of the next item, between that block's P stores and its p_ready arrival — the arrival is what
        // releases the P V product that overwrites O — so the softmax of the next item's first block no longer waits for the
        // last P V product of this one (measured: 4700 clk per item in this epilogue, most of it that wait).
        int prev_item = -1;
        float prev_l = 0.0f, prev_m = 0.0f;
        auto epilogue = [&](int item, float l, float m_ref, uint32_t parity) {
            const long long te = (args.dbg != nullptr && warp == 4 && lane == 0) ? clock64() : 0;
            const int mt = item % args.q_tiles;
            const int bh = item / args.q_tiles;
            const int h = bh % args.H, b = bh / args.H;
            tmem_st_x1(tmem_x + 4 + hsel, __float_as_uint(l));
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            bar_sync_64(1 + wq);
            ptx::tc_fence_after();
            l += __uint_as_float(tmem_ld_x1(tmem_x + 4 + (hsel ^ 1)));
            ptx::tmem_ld_wait();
            ptx::mbar_wait(o_done, parity);
            ptx::tc_fence_after();
            uint32_t o[32];
            ptx::tmem_ld_32x32(tmem_o + lane_off + hsel * 32, o);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            bar_sync_64(1 + wq);          // both partial sums have been read: the exchange cells may be rewritten
            const int sq = mt * kBM + row;
            if (sq < args.Sq) {
                const float inv = 1.0f / l;
                float4* dst = reinterpret_cast<float4*>(args.o + (((size_t)b * args.Sq + sq) * args.H + h) * kD + hsel * 32);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    dst[k] = make_float4(__uint_as_float(o[4 * k]) * inv, __uint_as_float(o[4 * k + 1]) * inv,
                                         __uint_as_float(o[4 * k + 2]) * inv, __uint_as_float(o[4 * k + 3]) * inv);
                if (hsel == 0) args.lse[((size_t)b * args.H + h) * args.Sq + sq] = m_ref + ptx::lg2(l);
            }
            if (args.dbg != nullptr && warp == 4 && lane == 0) args.dbg[blockIdx.x * 16 + 10] += clock64() - te;
            __syncwarp();
        };
        for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) {
            const int mt = item % args.q_tiles;
            float m_ref = -INFINITY, l = 0.0f;
            const int nb = blocks_of(item);
            for (int j = 0; j < nb; ++j, ++g) {
                const uint32_t buf = g & 1u;
                const bool rec = args.dbg != nullptr && warp == 4 && lane == 0;
                long long* dbg = args.dbg + blockIdx.x * 16;
                long long t0 = rec ? clock64() : 0, t1;
#define ATTN_STAMP(i) { if (rec) { t1 = clock64(); dbg[i] += t1 - t0; t0 = t1; } __syncwarp(); }
                ptx::mbar_wait(s_full(buf), (g >> 1) & 1);
                ATTN_STAMP(4)
                ptx::tc_fence_after();
                float s[64];
                const uint32_t s_tmem = tmem_base + lane_off + buf * kBN + c0;
                ptx::tmem_ld_32x32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
                ptx::tmem_ld_32x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
                ptx::tmem_ld_wait();
                ATTN_STAMP(5)
                const int kv_left = args.Skv - j * kBN;
                if (kv_left < kBN) {
#pragma unroll
                    for (int k = 0; k < 64; ++k)
                        if (c0 + k >= kv_left) s[k] = -INFINITY;
                }
                if (CAUSAL && j == mt) {
#pragma unroll
                    for (int k = 0; k < 64; ++k)
                        if (c0 + k > row) s[k] = -INFINITY;
                }
                float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
                for (int k = 4; k < 64; k += 4) {
                    mx0 = fmaxf(mx0, s[k]); mx1 = fmaxf(mx1, s[k + 1]);
                    mx2 = fmaxf(mx2, s[k + 2]); mx3 = fmaxf(mx3, s[k + 3]);
                }
                float pm = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                // exchange with the thread that holds the other 64 columns of this row
                tmem_st_x1(tmem_x + buf * 2 + hsel, __float_as_uint(pm));
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                bar_sync_64(1 + wq);
                ptx::tc_fence_after();
                pm = fmaxf(pm, __uint_as_float(tmem_ld_x1(tmem_x + buf * 2 + (hsel ^ 1))));
                ptx::tmem_ld_wait();
                ATTN_STAMP(6)
                const float mb = pm * c;
                const bool need = mb > m_ref + kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    float alpha = 1.0f;
                    if (need) {
                        alpha = ptx::ex2(m_ref - mb);
                        m_ref = mb;
                        l *= alpha;
                    }
                    if (j > 0) {       // see the four-warp variant below for why parity (g-1)&1 is unambiguous
                        ptx::mbar_wait(o_done, (g - 1) & 1);
                        ptx::tc_fence_after();
                        uint32_t o[32];                       // this thread rescales its 32 of the 64 O columns
                        ptx::tmem_ld_32x32(tmem_o + lane_off + hsel * 32, o);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
                        ptx::tmem_st_32x32(tmem_o + lane_off + hsel * 32, o);
                    }
                }
                ATTN_STAMP(7)
                // exponentials (XU pipe, 16 lanes / clk / SM: the 64 of a thread keep it busy ~1000 clk per block) interleaved
                // with the split of the PREVIOUS chunk of eight (ALU / FMA pipes), so the split rides in the XU shadow
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                uint32_t hi[32], mid[32];
#pragma unroll
                for (int ch = 0; ch <= 8; ++ch) {
                    if (ch < 8) {
#pragma unroll
                        for (int k = 8 * ch; k < 8 * ch + 8; ++k) s[k] = ptx::ex2(fmaf(s[k], c, -m_ref));
                    }
                    if (ch > 0) {
                        const int k0 = 8 * (ch - 1);
                        l0 += s[k0] + s[k0 + 4]; l1 += s[k0 + 1] + s[k0 + 5]; l2 += s[k0 + 2] + s[k0 + 6]; l3 += s[k0 + 3] + s[k0 + 7];
#pragma unroll
                        for (int i = k0 / 2; i < k0 / 2 + 4; ++i) split_pack(s[2 * i], s[2 * i + 1], hi[i], mid[i]);
                    }
                }
                l += (l0 + l1) + (l2 + l3);
                ATTN_STAMP(8)
                ptx::tmem_st_32x32(s_tmem, hi);                // this 64-k half: [32 columns hi | 32 columns mid]
                ptx::tmem_st_32x32(s_tmem + 32, mid);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                if (prev_item >= 0) { epilogue(prev_item, prev_l, prev_m, (g - 1) & 1); prev_item = -1; }
                ptx::mbar_arrive(p_ready(buf));
                ATTN_STAMP(9)
                if (rec) dbg[11] += 1;
                __syncwarp();
#undef ATTN_STAMP
            }
            prev_item = item; prev_l = l; prev_m = m_ref;
        }
        if (prev_item >= 0) epilogue(prev_item, prev_l, prev_m, (g - 1) & 1);
    } else if (warp >= 4) {
        // ========================= softmax =========================
        const int wq = warp & 3;                         // TMEM lane quarter this warp may access
        const int row = wq * 32 + lane;                  // query row within the tile
        const uint32_t lane_off = uint32_t(wq * 32) << 16;
        const float c = args.c;
        uint32_t g = 0;
        for (int item = blockIdx.x; item < args.total_items; item += gridDim.x) {
            const int mt = item % args.q_tiles;
            const int bh = item / args.q_tiles;
            const int h = bh % args.H, b = bh / args.H;
            float m_ref = -INFINITY, l = 0.0f;
            const int nb = blocks_of(item);
            for (int j = 0; j < nb; ++j, ++g) {
                const uint32_t buf = g & 1u;
                ptx::mbar_wait(s_full(buf), (g >> 1) & 1);
                ptx::tc_fence_after();
                float s[kBN];
                const uint32_t s_tmem = tmem_base + lane_off + buf * kBN;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    ptx::tmem_ld_32x32(s_tmem + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[ch * 32]));
                ptx::tmem_ld_wait();
                const int kv_left = args.Skv - j * kBN;
                if (kv_left < kBN) {
#pragma unroll
                    for (int k = 0; k < kBN; ++k)
                        if (k >= kv_left) s[k] = -INFINITY;
                }
                if (CAUSAL && j == mt) {                // the diagonal block: kv position k > query position row
#pragma unroll
                    for (int k = 0; k < kBN; ++k)
                        if (k > row) s[k] = -INFINITY;
                }
                float mx0 = s[0], mx1 = s[1], mx2 = s[2], mx3 = s[3];
#pragma unroll
                for (int k = 4; k < kBN; k += 4) {
                    mx0 = fmaxf(mx0, s[k]); mx1 = fmaxf(mx1, s[k + 1]);
                    mx2 = fmaxf(mx2, s[k + 2]); mx3 = fmaxf(mx3, s[k + 3]);
                }
                const float mb = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
                const bool need = mb > m_ref + kRescaleThreshold;
                if (__any_sync(0xffffffffu, need)) {
                    float alpha = 1.0f;
                    if (need) {
                        alpha = ptx::ex2(m_ref - mb);      // 0 on the first block (m_ref = -inf)
                        m_ref = mb;
                        l *= alpha;
                    }
                    if (j > 0) {
                        // PV(g-1) must have landed in O before O is rescaled.  o_done is only waited on
                        // here (rare) and in the epilogue, so PV(g-1) overlaps the exponentials of block
                        // g.  Skipping waits is phase-safe: s_full(g) (observed above) was committed
                        // after PV(g-2), and PV(g) cannot start before this thread's p_ready(g), so the
                        // barrier has completed either g-1 or g phases — parity (g-1)&1 is unambiguous.
                        ptx::mbar_wait(o_done, (g - 1) & 1);
                        ptx::tc_fence_after();
                        uint32_t o[kD];
                        ptx::tmem_ld_32x32(tmem_o + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
                        ptx::tmem_ld_32x32(tmem_o + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < kD; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
                        ptx::tmem_st_32x32(tmem_o + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
                        ptx::tmem_st_32x32(tmem_o + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
                    }
                }
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
                for (int k = 0; k < kBN; k += 4) {
                    const float p0 = ptx::ex2(fmaf(s[k], c, -m_ref));
                    const float p1 = ptx::ex2(fmaf(s[k + 1], c, -m_ref));
                    const float p2 = ptx::ex2(fmaf(s[k + 2], c, -m_ref));
                    const float p3 = ptx::ex2(fmaf(s[k + 3], c, -m_ref));
                    l0 += p0; l1 += p1; l2 += p2; l3 += p3;       // rounding to nearest is zero-mean: sum the exact p
                    if (BX) {
                        s[k] = p0; s[k + 1] = p1; s[k + 2] = p2; s[k + 3] = p3;
                    } else {
                        s[k] = __uint_as_float(rna_tf32(p0)); s[k + 1] = __uint_as_float(rna_tf32(p1));
                        s[k + 2] = __uint_as_float(rna_tf32(p2)); s[k + 3] = __uint_as_float(rna_tf32(p3));
                    }
                }
                l += (l0 + l1) + (l2 + l3);
                if (BX) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t hi[32], mid[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) split_pack(s[hf * 64 + 2 * i], s[hf * 64 + 2 * i + 1], hi[i], mid[i]);
                        ptx::tmem_st_32x32(s_tmem + hf * 64, hi);
                        ptx::tmem_st_32x32(s_tmem + hf * 64 + 32, mid);
                    }
                } else {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
                        ptx::tmem_st_32x32(s_tmem + ch * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[ch * 32]));
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(p_ready(buf));
            }
            // ---- epilogue: O / l → global, log-sum-exp → saved ----
            ptx::mbar_wait(o_done, (g - 1) & 1);
            ptx::tc_fence_after();
            uint32_t o[kD];
            ptx::tmem_ld_32x32(tmem_o + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
            ptx::tmem_ld_32x32(tmem_o + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            const int sq = mt * kBM + row;
            if (sq < args.Sq) {
                const float inv = 1.0f / l;
                float4* dst = reinterpret_cast<float4*>(args.o + (((size_t)b * args.Sq + sq) * args.H + h) * kD);
#pragma unroll
                for (int k = 0; k < kD / 4; ++k)
                    dst[k] = make_float4(__uint_as_float(o[4 * k]) * inv, __uint_as_float(o[4 * k + 1]) * inv,
                                         __uint_as_float(o[4 * k + 2]) * inv, __uint_as_float(o[4 * k + 3]) * inv);
                args.lse[((size_t)b * args.H + h) * args.Sq + sq] = m_ref + ptx::lg2(l);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace

int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }

bool attn_fused_supported(int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv) {
    return dk == kD && dv == kD && B > 0 && H > 0 && Sq > 0 && Skv > 0 && B < 65536 && H < 65536 &&
           Sq < (1ll << 30) && Skv < (1ll << 30);
}

// bx = false: q / k / v are fp32 (token strides ld*).  bx = true: q / k / v point at the hi plane of split-bf16 planes
// (token strides ld* in bf16 elements, the mid plane pl* elements after the hi plane).
int attn_fwd_launch(const void* q, const void* k, const void* v, float* o, float* lse, int64_t B, int64_t H,
                    int64_t Sq, int64_t Skv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t plq, int64_t plk, int64_t plv,
                    int causal, bool bx, int nterms, cudaStream_t stream) {
    NPM_REQUIRE(!causal || Sq == Skv, "mha_core_fwd: the causal mask needs Sq == Skv");
    NPM_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "mha_core_fwd: pointers must be 16-byte aligned");
    CUtensorMap tmQ, tmK, tmV;
    int rc;
    const uint64_t HD = (uint64_t)H * kD;
    if (bx) {
        NPM_REQUIRE(ldq >= (int64_t)HD && ldk >= (int64_t)HD && ldv >= (int64_t)HD && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 &&
                    plq % 8 == 0 && plk % 8 == 0 && plv % 8 == 0, "mha_core_fwd: plane strides must be multiples of 8 bf16 elements");
        if ((rc = make_tensor_map_bf16_planes(&tmQ, q, Sq, H, B, ldq, plq, kBM))) return rc;
        if ((rc = make_tensor_map_bf16_planes(&tmK, k, Skv, H, B, ldk, plk, kBN))) return rc;
        if ((rc = make_tensor_map_bf16_planes(&tmV, v, Skv, H, B, ldv, plv, kBN))) return rc;
    } else {
    NPM_REQUIRE(ldq >= (int64_t)HD && ldk >= (int64_t)HD && ldv >= (int64_t)HD && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0,
                "mha_core_fwd: token strides must be >= H*d and multiples of 4 floats");
    // token stride ld*: q / k / v may be column blocks of one packed [tokens, 3*H*d] projection output
    const float* qf = reinterpret_cast<const float*>(q);
    const float* kf = reinterpret_cast<const float*>(k);
    const float* vf = reinterpret_cast<const float*>(v);
    if ((rc = make_tensor_map_4d(&tmQ, qf, kD, Sq, H, B, ldq, kD, Sq * ldq, 32, kBM, true, false))) return rc;
    if ((rc = make_tensor_map_4d(&tmK, kf, kD, Skv, H, B, ldk, kD, Skv * ldk, 32, kBN, true, false))) return rc;
    if ((rc = make_tensor_map_4d(&tmV, vf, kD, Skv, H, B, ldv, kD, Skv * ldv, 32, kBN, true, true))) return rc;
    }
    FwdArgs a;
    a.B = (int)B; a.H = (int)H; a.Sq = (int)Sq; a.Skv = (int)Skv;
    a.q_tiles = (int)((Sq + kBM - 1) / kBM);
    a.n_kv = (int)((Skv + kBN - 1) / kBN);
    const int64_t items = B * H * a.q_tiles;
    NPM_REQUIRE(items < (1ll << 30), "mha_core_fwd: too many tiles");
    a.total_items = (int)items;
    a.causal = causal ? 1 : 0;
    a.t0 = nterms == 1 ? 2 : 0;
    a.c = (float)(1.4426950408889634 / sqrt((double)kD));
    a.o = o; a.lse = lse;
    a.dbg = nullptr;
    static const bool dbg_times = getenv("NPM_ATTN_DEBUG_TIMES") != nullptr;      // tools only: synchronises and prints
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attn_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attn_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) { set_error("attn_fwd smem attribute: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
        configured = true;
    }
    int grid = (int)(items < num_sms() ? items : num_sms());
    // causal: the work of an item grows with its tile index, and CTA c takes items c, c + grid, ...: a grid size
    // coprime with the tile count makes every CTA cycle through all tile indices (148 and 8 share the factor 4)
    if (causal) while (grid > 1 && gcd_int(grid, a.q_tiles) != 1) --grid;
    auto kern = causal ? (bx ? attn_fwd_kernel<true, true> : attn_fwd_kernel<true, false>)
                       : (bx ? attn_fwd_kernel<false, true> : attn_fwd_kernel<false, false>);
    if (dbg_times && bx) {
        cudaMalloc(&a.dbg, sizeof(long long) * 16 * grid);
        cudaMemset(a.dbg, 0, sizeof(long long) * 16 * grid);
    }
    cudaError_t le = launch_pdl(kern, dim3(grid), dim3(bx ? kThreadsBx : kThreads), kSmemBytes, stream, 1, tmQ, tmK, tmV, a);
    if (le != cudaSuccess) { set_error("attn_fwd_kernel launch: %s", cudaGetErrorString(le)); return NPM_ERR_CUDA; }
    count_launch();
    if (a.dbg != nullptr) {
        cudaStreamSynchronize(stream);
        std::vector<long long> h(16 * (size_t)grid);
        cudaMemcpy(h.data(), a.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        double t[16] = {0};
        for (int cta = 0; cta < grid; ++cta)
            for (int i = 0; i < 16; ++i) t[i] += (double)h[16 * cta + i];
        fprintf(stderr, "[attn_fwd bx] cycles per kv block: issuer waits k %.0f, p_ready %.0f, v %.0f | softmax(w4): s_full wait %.0f, "
                "tmem ld %.0f, max+exchange %.0f, rescale %.0f, exp %.0f, split+st+arrive %.0f; epilogue per item %.0f\n",
                t[0] / t[3], t[1] / t[3], t[2] / t[3], t[4] / t[11], t[5] / t[11], t[6] / t[11], t[7] / t[11], t[8] / t[11], t[9] / t[11],
                t[10] * 8 / t[11]);
        cudaFree(a.dbg);
    }
    return check_launch("attn_fwd_kernel");
}

}  // namespace npm
