// gemm_tc.cu — persistent, warp-specialised TF32 / 3xTF32 GEMM on tcgen05 + TMEM + TMA.
//
//   C[z][m,n] (=|+=) alpha * sum_k A[z][m,k] * B[z][k,n] (+ bias[n]) (ReLU)
//
// This single kernel family serves every contraction of the reference hot path:
// Linear fwd/dX/dW (layers/mlp.py:21-40), the MultiHeadAttention projections and
// the batched QK^T / PV / dP / dV / dQ / dK products (layers/attentions.py:88-184).
//
// Design (one CTA per SM, 128 x BLOCK_N output tile, K step of 32 fp32 = one
// 128-byte swizzle span):
//   warp 4      TMA producer   : cp.async.bulk.tensor → smem ring (mbarrier full/empty)
//   warp 5      MMA issuer     : one lane issues tcgen05.mma kind::tf32, accumulators in TMEM
//                                (two accumulator buffers so the epilogue of tile i overlaps
//                                the main loop of tile i+1)
//   warps 0-3   epilogue       : tcgen05.ld → alpha/bias/ReLU → swizzled smem → TMA store
//                                (or TMA reduce-add for C += ...)
//   warps 6-9   split (3xTF32) : rewrite each landed fp32 tile as hi = trunc_tf32(x) in place
//                                and lo = x - hi beside it; the issuer then runs
//                                lo*hi + hi*lo + hi*hi, recovering ~fp32 accuracy.
// Operands may be K-major or MN-major ("transposed"): the UMMA descriptors take the
// 128B-swizzled tiles exactly as TMA lands them, so no transposes are ever materialised.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "ptx.cuh"

namespace npm {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;   // fp32 elements per K step (128 bytes)
constexpr int kUmmaK  = 8;    // K of one tcgen05.mma kind::tf32

struct GemmTcArgs {
    int M, N, K;
    int tiles_m, tiles_n, nb1, total_tiles;
    int splits, kb_per_split, items_per_split;   // split-K: work item = (split, batch, n-tile, m-tile); partials are TMA-reduced into C
    float alpha;
    const float* bias;
    const float* residual;     // [M, N] with leading dimension ldr, or nullptr
    int64_t ldr;
    int relu, accum;
    uint64_t desc_a, desc_b;   // UMMA shared-memory descriptor templates (start address = 0)
    int band_m;                // pair kernel: tiles are walked in bands of band_m row tiles (n fastest inside a band) so that the
                               // tiles running concurrently share A AND B panels instead of all hammering one B panel; 0 = m fastest
    int dbg_mode;              // tools only: 1 = no TMA loads / no full-barrier waits (pure MMA issue rate)
    int a_chunked, b_chunked;  // pair kernel: MN-major operand described as a 5-D {32, K, MN/32, nb1, nb2} tensor, one box per stage
    long long* dbg;            // tools only (NPM_GEMM_DEBUG_TIMES): per-CTA wait-cycle counters of the pair kernel's roles
};

template <int BLOCK_N, int NPASS>
struct TcCfg {
    static constexpr int kABytes     = kBlockM * kBlockK * 4;           // 16 KB
    static constexpr int kBBytes     = BLOCK_N * kBlockK * 4;
    static constexpr int kRawBytes   = kABytes + kBBytes;
    static constexpr int kStageBytes = kRawBytes * (NPASS == 3 ? 2 : 1);  // + lo copies
    static constexpr int kEpiBytes   = 4 * 2 * 4096;                    // 4 warps x 2 buffers
    static constexpr int kBarBytes   = 512;
    static constexpr int kBudget     = 232448 - 1024;                   // 227 KB minus align slack
    static constexpr int kStagesRaw  = (kBudget - kEpiBytes - kBarBytes) / kStageBytes;
    static constexpr int kStages     = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kSmemBytes  = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;
    static constexpr int kTmemCols   = 2 * BLOCK_N;                     // 128 / 256 / 512
    static constexpr int kThreads    = NPASS == 3 ? 320 : 192;
    static_assert(kStages >= 2, "need at least a double-buffered ring");
    static_assert(kTmemCols >= 32 && kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "");
};

template <int BLOCK_N, bool A_MN, bool B_MN, int NPASS>
__global__ void __launch_bounds__(TcCfg<BLOCK_N, NPASS>::kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmTcArgs args) {
    using Cfg = TcCfg<BLOCK_N, NPASS>;
    constexpr int S = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr  = ptx::smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr        = smem_raw + (base_addr - raw_addr);

    const uint32_t stage_addr = base_addr;                         // S x kStageBytes
    const uint32_t epi_addr   = base_addr + S * Cfg::kStageBytes;  // 1024-aligned (stage sizes are)
    const uint32_t bar_addr   = epi_addr + Cfg::kEpiBytes;
    // barrier slots (8 bytes each)
    auto full_bar   = [&](int s) { return bar_addr + 8u * s; };
    auto xf_bar     = [&](int s) { return bar_addr + 8u * (S + s); };
    auto empty_bar  = [&](int s) { return bar_addr + 8u * (2 * S + s); };
    auto tfull_bar  = [&](int a) { return bar_addr + 8u * (3 * S + a); };
    auto tempty_bar = [&](int a) { return bar_addr + 8u * (3 * S + 2 + a); };
    volatile uint32_t* tmem_slot =
        reinterpret_cast<volatile uint32_t*>(base_ptr + S * Cfg::kStageBytes + Cfg::kEpiBytes +
                                             8 * (3 * S + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (args.K + kBlockK - 1) / kBlockK;

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        ptx::prefetch_tensormap(&tmC);
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) {
                ptx::mbar_init(full_bar(s), 1);
                ptx::mbar_init(xf_bar(s), 128);
                ptx::mbar_init(empty_bar(s), 1);
            }
            for (int a = 0; a < 2; ++a) {
                ptx::mbar_init(tfull_bar(a), 1);
                ptx::mbar_init(tempty_bar(a), 4);
            }
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_mn = args.tiles_m * args.tiles_n;

    if (warp == 4) {
        // ============================ TMA producer ============================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
                const int sp = tile / args.items_per_split, t2 = tile - sp * args.items_per_split;
                const int z  = t2 / tiles_mn;
                const int r  = t2 - z * tiles_mn;
                const int kb0 = sp * args.kb_per_split;
                const int kb1 = min(num_kb, kb0 + args.kb_per_split);
                const int m0 = (r % args.tiles_m) * kBlockM;
                const int n0 = (r / args.tiles_m) * BLOCK_N;
                const int z1 = z % args.nb1, z2 = z / args.nb1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                    const uint32_t sB = sA + Cfg::kABytes;
                    const uint32_t fb = full_bar(stage);
                    ptx::mbar_arrive_expect_tx(fb, Cfg::kRawBytes);
                    const int k0 = kb * kBlockK;
                    if (!A_MN) {
                        ptx::tma_load_4d(sA, &tmA, fb, k0, m0, z1, z2);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBlockM / 32; ++c)
                            ptx::tma_load_4d(sA + c * 4096, &tmA, fb, m0 + c * 32, k0, z1, z2);
                    }
                    if (!B_MN) {
                        ptx::tma_load_4d(sB, &tmB, fb, k0, n0, z1, z2);
                    } else {
#pragma unroll
                        for (int c = 0; c < BLOCK_N / 32; ++c)
                            ptx::tma_load_4d(sB + c * 4096, &tmB, fb, n0 + c * 32, k0, z1, z2);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 5) {
        // ============================= MMA issuer =============================
        constexpr uint32_t idesc = ptx::umma_idesc_tf32(kBlockM, BLOCK_N, A_MN, B_MN);
        const uint64_t descA = args.desc_a, descB = args.desc_b;   // see gemm_tc_launch()
        constexpr uint32_t a_kstep = A_MN ? 1024u : 32u;   // bytes per UMMA_K (8) step
        constexpr uint32_t b_kstep = B_MN ? 1024u : 32u;
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
            const int kb0 = (tile / args.items_per_split) * args.kb_per_split;
            const int kb1 = min(num_kb, kb0 + args.kb_per_split);
            ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(NPASS == 3 ? xf_bar(stage) : full_bar(stage), phase);
                ptx::tc_fence_after();
                __syncwarp();
                if (ptx::elect_one()) {
                    const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                    const uint32_t sB = sA + Cfg::kABytes;
#pragma unroll
                    for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
                        const uint64_t da = ptx::umma_desc(descA, sA + kk * a_kstep);
                        const uint64_t db = ptx::umma_desc(descB, sB + kk * b_kstep);
                        const uint32_t first = (kb != kb0 || kk != 0) ? 1u : 0u;
                        if (NPASS == 1) {
                            ptx::umma_tf32(d_tmem, da, db, idesc, first);
                        } else {
                            const uint64_t da_lo =
                                ptx::umma_desc(descA, sA + Cfg::kRawBytes + kk * a_kstep);
                            const uint64_t db_lo =
                                ptx::umma_desc(descB, sB + Cfg::kRawBytes + kk * b_kstep);
                            ptx::umma_tf32(d_tmem, da_lo, db, idesc, first);   // lo * hi
                            ptx::umma_tf32(d_tmem, da, db_lo, idesc, 1u);      // hi * lo
                            ptx::umma_tf32(d_tmem, da, db, idesc, 1u);         // hi * hi
                        }
                    }
                    ptx::umma_commit(empty_bar(stage));            // smem slot reusable
                    if (kb == kb1 - 1) ptx::umma_commit(tfull_bar(acc));  // accumulator ready
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
    } else if (warp < 4) {
        // ============================== epilogue ==============================
        uint8_t* stg_base = base_ptr + S * Cfg::kStageBytes + warp * 2 * 4096;
        const uint32_t stg_addr = epi_addr + warp * 2 * 4096;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t nstore = 0;
        for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
            const int sp = tile / args.items_per_split, t2 = tile - sp * args.items_per_split;
            const int z  = t2 / tiles_mn;
            const int r  = t2 - z * tiles_mn;
            const int m0 = (r % args.tiles_m) * kBlockM;
            const int n0 = (r / args.tiles_m) * BLOCK_N;
            const int z1 = z % args.nb1, z2 = z / args.nb1;
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const bool rows_live = (m0 + warp * 32) < args.M;
#pragma unroll 1
            for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
                const int nc = n0 + chunk * 32;
                if (nc >= args.N || !rows_live) break;
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + (uint32_t(warp * 32) << 16) + acc * BLOCK_N + chunk * 32, v);
                ptx::tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * args.alpha;
                if (args.bias != nullptr && sp == 0) {
                    // 8 x 128-bit broadcast loads: 32 scalar loads per chunk made this epilogue longer than the main
                    // loop of a K = 1024 tile (FFN up-projection 730 -> 590 TF with a bias)
                    const float* bp = args.bias + nc;
                    if (nc + 32 <= args.N && ((reinterpret_cast<uintptr_t>(bp) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
                            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nc + j < args.N) f[j] += __ldg(bp + j);
                    }
                }
                if (args.residual != nullptr && sp == 0) {
                    const int64_t row_g = (int64_t)m0 + warp * 32 + lane;
                    if (row_g < args.M) {
                        const float* rrow = args.residual + row_g * args.ldr + nc;
                        if (nc + 32 <= args.N && ((reinterpret_cast<uintptr_t>(rrow) & 15u) == 0)) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow) + j);
                                f[4 * j] += r4.x; f[4 * j + 1] += r4.y; f[4 * j + 2] += r4.z; f[4 * j + 3] += r4.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (nc + j < args.N) f[j] += __ldg(rrow + j);
                        }
                    }
                }
                if (args.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)      // sign bit set <=> pre-activation < 0 (a genuine -0.0 is >= 0, activations.py:19: store +0.0)
                        f[j] = f[j] < 0.0f ? -0.0f : __uint_as_float(__float_as_uint(f[j]) & 0x7fffffffu);
                }
                const uint32_t buf = nstore & 1u;
                // the store that last read this buffer (two stores ago) must have drained
                if (lane == 0) ptx::tma_wait_group_read<1>();
                __syncwarp();
                uint8_t* row = stg_base + buf * 4096 + lane * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    *reinterpret_cast<float4*>(row + ((j ^ (lane & 7)) << 4)) = o;
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (ptx::elect_one()) {
                    if (args.accum || args.splits > 1)
                        ptx::tma_reduce_add_4d(&tmC, stg_addr + buf * 4096, nc, m0 + warp * 32, z1, z2);
                    else
                        ptx::tma_store_4d(&tmC, stg_addr + buf * 4096, nc, m0 + warp * 32, z1, z2);
                    ptx::tma_commit_group();
                }
                ++nstore;
            }
            // every tcgen05.ld of this accumulator has completed (wait::ld above)
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (lane == 0) ptx::tma_wait_group<0>();
    } else if (NPASS == 3) {
        // ===================== 3xTF32 operand split (warps 6-9) ================
        const int t = threadIdx.x - 192;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
            const int kb0 = (tile / args.items_per_split) * args.kb_per_split;
            const int kb1 = min(num_kb, kb0 + args.kb_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                uint8_t* raw = base_ptr + stage * Cfg::kStageBytes;
#pragma unroll 4
                for (int i = t; i < Cfg::kRawBytes / 16; i += 128) {
                    float4 x = *reinterpret_cast<float4*>(raw + i * 16);
                    float4 hi, lo;
                    hi.x = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
                    hi.y = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
                    hi.z = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
                    hi.w = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
                    lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
                    *reinterpret_cast<float4*>(raw + i * 16) = hi;
                    *reinterpret_cast<float4*>(raw + Cfg::kRawBytes + i * 16) = lo;
                }
                ptx::fence_proxy_async_smem();
                ptx::mbar_arrive(xf_bar(stage));
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}


// ===========================================================================================
// CTA-pair variant (TF32 single pass): two CTAs of a cluster compute one 256 x BLOCK_N tile with
// tcgen05.mma.cta_group::2.  Each CTA stages only ITS 128 rows of A and ITS half of B's columns
// (32 KB per K step instead of 48 KB), the tensor cores of both SMs read both halves of B — shared
// memory bandwidth, the limiter of fp32-operand MMA, is spent on 1/3 fewer bytes per flop, and the
// ring is 6 stages deep.  The leader CTA (rank 0) issues every MMA; a multicast tcgen05.commit
// releases the ring slot / publishes the accumulator in both CTAs; both CTAs run their own
// producer and their own epilogue (TMEM lanes = their 128 rows).
// ===========================================================================================
// r-th tile of a z slice -> (row tile, column tile)
__device__ __forceinline__ void tile_coords(int r, int tiles_m, int tiles_n, int band, int& tm, int& tn) {
    if (band <= 1) { tm = r % tiles_m; tn = r / tiles_m; return; }
    const int per_band = band * tiles_n;
    const int b = r / per_band, idx = r - b * per_band;
    const int width = min(band, tiles_m - b * band);        // the last band may be narrower
    tm = b * band + idx % width;
    tn = idx / width;
}

template <int BLOCK_N, int KS>
struct Tc2Cfg {
    static constexpr int kHalfN      = BLOCK_N / 2;
    static constexpr int kABytes     = kBlockM * KS * 4;                // 16 KB per 32 k
    static constexpr int kBBytes     = kHalfN * KS * 4;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBytes   = 4 * 2 * 4096;
    static constexpr int kBarBytes   = 512;
    static constexpr int kBudget     = 232448 - 1024;
    static constexpr int kStagesRaw  = (kBudget - kEpiBytes - kBarBytes) / kStageBytes;
    static constexpr int kStages     = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kSmemBytes  = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;
    static constexpr int kTmemCols   = 2 * BLOCK_N;
    static constexpr int kThreads    = 192;
};

// KS = fp32 elements of K per ring stage: 32, or 64 with each operand landed by ONE 32 KB TMA box (K-major: {32 k, rows,
// 2 k-chunks} of a {32, rows, K/32} view; MN-major: {32 mn, 64 k, chunks}) — the TMA unit costs ~46 clk per box on top
// of bytes / 70 B/clk (tools/micro/tma_bw.cu), and with fp32 operands the ring is what limits the main loop.
template <int BLOCK_N, bool A_MN, bool B_MN, int KS>
__global__ void __launch_bounds__(192, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmTcArgs args) {
    using Cfg = Tc2Cfg<BLOCK_N, KS>;
    constexpr int S = Cfg::kStages;
    pdl_trigger();            // the next kernel in the stream may be scheduled; it waits for this grid before reading

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr  = ptx::smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr        = smem_raw + (base_addr - raw_addr);

    const uint32_t stage_addr = base_addr;
    const uint32_t epi_addr   = base_addr + S * Cfg::kStageBytes;
    const uint32_t bar_addr   = epi_addr + Cfg::kEpiBytes;
    auto full_bar   = [&](int s) { return bar_addr + 8u * s; };
    auto empty_bar  = [&](int s) { return bar_addr + 8u * (S + s); };
    auto tfull_bar  = [&](int a) { return bar_addr + 8u * (2 * S + a); };
    auto tempty_bar = [&](int a) { return bar_addr + 8u * (2 * S + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(
        base_ptr + S * Cfg::kStageBytes + Cfg::kEpiBytes + 8 * (2 * S + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();          // 0 = pair leader
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_kb = (args.K + KS - 1) / KS;

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        ptx::prefetch_tensormap(&tmC);
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) {
                ptx::mbar_init(full_bar(s), 1);       // the leader's arrive.expect_tx (bytes of both CTAs)
                ptx::mbar_init(empty_bar(s), 1);      // one multicast commit
            }
            for (int a = 0; a < 2; ++a) {
                ptx::mbar_init(tfull_bar(a), 1);
                ptx::mbar_init(tempty_bar(a), 8);     // 4 epilogue warps of each CTA (leader's copy is the one used)
            }
            ptx::fence_mbar_init();
        }
        __syncwarp();
        ptx::tmem_alloc_2sm(ptx::smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::kTmemCols);
        ptx::tmem_relinquish_2sm();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();               // everything above touched only this CTA's shared memory / TMEM: now wait for the predecessor

    const int tiles_mn = args.tiles_m * args.tiles_n;      // tiles_m counts 256-row tiles here

    if (warp == 4) {
        // ============================ TMA producer (both CTAs) ============================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < args.total_tiles; tile += num_clusters) {
                const int sp = tile / args.items_per_split, t2 = tile - sp * args.items_per_split;
                const int z  = t2 / tiles_mn;
                const int r  = t2 - z * tiles_mn;
                const int kb0 = sp * args.kb_per_split;
                const int kb1 = min(num_kb, kb0 + args.kb_per_split);
                int tm, tn;
                tile_coords(r, args.tiles_m, args.tiles_n, args.band_m, tm, tn);
                const int m0 = tm * (2 * kBlockM) + (int)rank * kBlockM;
                const int n0 = tn * BLOCK_N + (int)rank * Cfg::kHalfN;
                const int z1 = z % args.nb1, z2 = z / args.nb1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (args.dbg_mode == 1) break;
                    if (args.dbg) {
                        const long long t0 = clock64();
                        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                        args.dbg[blockIdx.x * 8 + 0] += clock64() - t0;
                    } else {
                        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    }
                    const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                    const uint32_t sB = sA + Cfg::kABytes;
                    const uint32_t fb = ptx::mapa(full_bar(stage), 0);       // the leader's barrier
                    if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
                    const int k0 = kb * KS;
                    if (!A_MN) {
                        if (KS == 32) ptx::tma_load_4d_2sm(sA, &tmA, fb, k0, m0, z1, z2);
                        else          ptx::tma_load_5d_2sm(sA, &tmA, fb, 0, m0, k0 / 32, z1, z2);
                    } else if (args.a_chunked) {
                        // one 16 KB box {32 m, 32 k, 4 chunks}: the TMA unit's cost is ~46 clk per box + bytes / 70 B/clk
                        // (tools/micro/tma_bw.cu), so four 4 KB boxes cost 1.5x the time of one 16 KB box
                        ptx::tma_load_5d_2sm(sA, &tmA, fb, 0, k0, m0 / 32, z1, z2);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBlockM / 32; ++c)
                            ptx::tma_load_4d_2sm(sA + c * 4096, &tmA, fb, m0 + c * 32, k0, z1, z2);
                    }
                    if (!B_MN) {
                        if (KS == 32) ptx::tma_load_4d_2sm(sB, &tmB, fb, k0, n0, z1, z2);
                        else          ptx::tma_load_5d_2sm(sB, &tmB, fb, 0, n0, k0 / 32, z1, z2);
                    } else if (args.b_chunked) {
                        ptx::tma_load_5d_2sm(sB, &tmB, fb, 0, k0, n0 / 32, z1, z2);
                    } else {
#pragma unroll
                        for (int c = 0; c < Cfg::kHalfN / 32; ++c)
                            ptx::tma_load_4d_2sm(sB + c * 4096, &tmB, fb, n0 + c * 32, k0, z1, z2);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 5) {
        // ============================= MMA issuer (leader CTA) =============================
        if (rank == 0 && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::umma_idesc_tf32(2 * kBlockM, BLOCK_N, A_MN, B_MN);
            const uint64_t descA = args.desc_a, descB = args.desc_b;
            // K-major: 32 B per K8 step inside a 128-byte row, the next 32-k slab after 4 steps; MN-major: 8 k-rows of 128 B
            auto a_off = [](int kk) -> uint32_t { return A_MN ? uint32_t(kk) * 1024u : uint32_t(kk >> 2) * (kBlockM * 128u) + uint32_t(kk & 3) * 32u; };
            auto b_off = [](int kk) -> uint32_t { return B_MN ? uint32_t(kk) * 1024u : uint32_t(kk >> 2) * (Cfg::kHalfN * 128u) + uint32_t(kk & 3) * 32u; };
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            long long w_full = 0, w_tempty = 0, n_ksteps = 0;
            const long long t_begin = args.dbg ? clock64() : 0;
            for (int tile = cluster_id; tile < args.total_tiles; tile += num_clusters) {
                const int kb0 = (tile / args.items_per_split) * args.kb_per_split;
                const int kb1 = min(num_kb, kb0 + args.kb_per_split);
                if (args.dbg) {
                    const long long t0 = clock64();
                    ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    w_tempty += clock64() - t0;
                } else {
                    ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                }
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (args.dbg_mode == 1) {
                        ++n_ksteps;
                    } else if (args.dbg) {
                        const long long t0 = clock64();
                        ptx::mbar_wait(full_bar(stage), phase);
                        w_full += clock64() - t0;
                        ++n_ksteps;
                    } else {
                        ptx::mbar_wait(full_bar(stage), phase);
                    }
                    ptx::tc_fence_after();
                    const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                    const uint32_t sB = sA + Cfg::kABytes;
#pragma unroll
                    for (int kk = 0; kk < KS / kUmmaK; ++kk) {
                        const uint64_t da = ptx::umma_desc(descA, sA + a_off(kk));
                        const uint64_t db = ptx::umma_desc(descB, sB + b_off(kk));
                        ptx::umma_tf32_2sm(d_tmem, da, db, idesc, (kb != kb0 || kk != 0) ? 1u : 0u);
                    }
                    ptx::umma_commit_2sm(empty_bar(stage), 3);                    // slot reusable in both CTAs
                    if (kb == kb1 - 1) ptx::umma_commit_2sm(tfull_bar(acc), 3);   // accumulator ready in both CTAs
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
            if (args.dbg) {
                args.dbg[blockIdx.x * 8 + 1] = w_full;
                args.dbg[blockIdx.x * 8 + 2] = w_tempty;
                args.dbg[blockIdx.x * 8 + 3] = clock64() - t_begin;
                args.dbg[blockIdx.x * 8 + 4] = n_ksteps;
            }
        }
    } else if (warp < 4) {
        // ============================== epilogue (both CTAs) ==============================
        uint8_t* stg_base = base_ptr + S * Cfg::kStageBytes + warp * 2 * 4096;
        const uint32_t stg_addr = epi_addr + warp * 2 * 4096;
        int acc = 0;
        uint32_t acc_phase = 0;
        uint32_t nstore = 0;
        for (int tile = cluster_id; tile < args.total_tiles; tile += num_clusters) {
            const int sp = tile / args.items_per_split, t2 = tile - sp * args.items_per_split;
            const int z  = t2 / tiles_mn;
            const int r  = t2 - z * tiles_mn;
            int tm, tn;
            tile_coords(r, args.tiles_m, args.tiles_n, args.band_m, tm, tn);
            const int m0 = tm * (2 * kBlockM) + (int)rank * kBlockM;
            const int n0 = tn * BLOCK_N;
            const int z1 = z % args.nb1, z2 = z / args.nb1;
            if (args.dbg && warp == 0 && lane == 0) {
                const long long t0 = clock64();
                ptx::mbar_wait(tfull_bar(acc), acc_phase);
                args.dbg[blockIdx.x * 8 + 5] += clock64() - t0;
                args.dbg[blockIdx.x * 8 + 6] -= clock64();
            }
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const bool rows_live = (m0 + warp * 32) < args.M;
#pragma unroll 1
            for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
                const int nc = n0 + chunk * 32;
                if (nc >= args.N || !rows_live) break;
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + (uint32_t(warp * 32) << 16) + acc * BLOCK_N + chunk * 32, v);
                ptx::tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * args.alpha;
                if (args.bias != nullptr && sp == 0) {
                    // 8 x 128-bit broadcast loads: 32 scalar loads per chunk made this epilogue longer than the main
                    // loop of a K = 1024 tile (FFN up-projection 730 -> 590 TF with a bias)
                    const float* bp = args.bias + nc;
                    if (nc + 32 <= args.N && ((reinterpret_cast<uintptr_t>(bp) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
                            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nc + j < args.N) f[j] += __ldg(bp + j);
                    }
                }
                if (args.residual != nullptr && sp == 0) {
                    const int64_t row_g = (int64_t)m0 + warp * 32 + lane;
                    if (row_g < args.M) {
                        const float* rrow = args.residual + row_g * args.ldr + nc;
                        if (nc + 32 <= args.N && ((reinterpret_cast<uintptr_t>(rrow) & 15u) == 0)) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow) + j);
                                f[4 * j] += r4.x; f[4 * j + 1] += r4.y; f[4 * j + 2] += r4.z; f[4 * j + 3] += r4.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (nc + j < args.N) f[j] += __ldg(rrow + j);
                        }
                    }
                }
                if (args.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)      // sign bit set <=> pre-activation < 0 (a genuine -0.0 is >= 0, activations.py:19: store +0.0)
                        f[j] = f[j] < 0.0f ? -0.0f : __uint_as_float(__float_as_uint(f[j]) & 0x7fffffffu);
                }
                const uint32_t buf = nstore & 1u;
                if (lane == 0) ptx::tma_wait_group_read<1>();
                __syncwarp();
                uint8_t* row = stg_base + buf * 4096 + lane * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    *reinterpret_cast<float4*>(row + ((j ^ (lane & 7)) << 4)) = o;
                }
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (ptx::elect_one()) {
                    if (args.accum || args.splits > 1)
                        ptx::tma_reduce_add_4d(&tmC, stg_addr + buf * 4096, nc, m0 + warp * 32, z1, z2);
                    else
                        ptx::tma_store_4d(&tmC, stg_addr + buf * 4096, nc, m0 + warp * 32, z1, z2);
                    ptx::tma_commit_group();
                }
                ++nstore;
            }
            if (args.dbg && warp == 0 && lane == 0) args.dbg[blockIdx.x * 8 + 6] += clock64();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) ptx::mbar_arrive(tempty_bar(acc));
                else           ptx::mbar_arrive_cluster(ptx::mapa(tempty_bar(acc), 0));
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (lane == 0) ptx::tma_wait_group<0>();
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();       // the peer's barriers / TMEM must outlive every remote arrive and multicast commit
    if (warp == 5) ptx::tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn driver_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Per-thread tensor-map cache (the one piece of hidden state SURVEY.md §8b allows): a tensor map is a pure function of
// (base pointer, data type, dims, strides, box, swizzle), and a training step re-creates the same ~3000 of them every
// step (three per GEMM launch) — torch's caching allocator hands the same addresses back.  Direct mapped, 2048 entries,
// the full key is compared on a hit.  NPM_NO_TMAP_CACHE=1 disables it (A/B).
struct TmKey {
    void* addr;
    cuuint64_t dims[5];
    cuuint64_t strides[4];
    cuuint32_t box[5];
    uint32_t meta;       // data type | rank << 8 | swizzle << 16 | interleave << 24
};
struct TmEntry { TmKey key; CUtensorMap map; bool valid; };
constexpr int kTmCacheSize = 2048;

CUresult cached_encode(CUtensorMap* tm, CUtensorMapDataType dt, cuuint32_t rank, void* addr, const cuuint64_t* dims,
                       const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                       CUtensorMapSwizzle sw, CUtensorMapL2promotion l2, CUtensorMapFloatOOBfill oob) {
    static const bool off = getenv("NPM_NO_TMAP_CACHE") != nullptr;
    EncodeTiledFn fn = driver_encode_fn();
    bool plain = off || rank > 5;
    for (cuuint32_t i = 0; i < rank && !plain; ++i) plain = estr[i] != 1;
    if (plain || l2 != CU_TENSOR_MAP_L2_PROMOTION_L2_256B || oob != CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        return fn(tm, dt, rank, addr, dims, strides, box, estr, il, sw, l2, oob);
    static thread_local TmEntry* cache = nullptr;
    if (!cache) cache = new TmEntry[kTmCacheSize]();
    TmKey k;
    memset(&k, 0, sizeof(k));
    k.addr = addr;
    for (cuuint32_t i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
    for (cuuint32_t i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
    k.meta = (uint32_t)dt | (rank << 8) | ((uint32_t)sw << 16) | ((uint32_t)il << 24);
    uint64_t h = 1469598103934665603ull;
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    for (size_t i = 0; i < sizeof(k) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    TmEntry& e = cache[(h >> 20) & (kTmCacheSize - 1)];
    if (e.valid && memcmp(&e.key, &k, sizeof(k)) == 0) { *tm = e.map; return CUDA_SUCCESS; }
    const CUresult rc = fn(tm, dt, rank, addr, dims, strides, box, estr, il, sw, l2, oob);
    if (rc == CUDA_SUCCESS) { e.key = k; e.map = *tm; e.valid = true; }
    return rc;
}

EncodeTiledFn encode_fn() { return driver_encode_fn() ? &cached_encode : nullptr; }

}  // namespace

// 4-D fp32 tensor map: dims (d0 contiguous, d1, d2, d3), strides in ELEMENTS for d1..d3,
// box (b0, b1, 1, 1), 128-byte swizzle, zero fill out of bounds.
int make_tensor_map_4d(CUtensorMap* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t d3, uint64_t s1, uint64_t s2, uint64_t s3, uint32_t b0, uint32_t b1,
                       bool round_tf32, bool atom32b) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return NPM_ERR_CUDA;
    }
    cuuint64_t dims[4]    = {d0, d1, d2, d3};
    cuuint64_t strides[3] = {s1 * 4, s2 * 4, s3 * 4};
    cuuint32_t box[4]     = {b0, b1, 1, 1};
    cuuint32_t estr[4]    = {1, 1, 1, 1};
    CUresult rc = fn(tm, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                     4, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) "
                  "box=(%u,%u) base=%p",
                  (int)rc, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                  (unsigned long long)d3, (unsigned long long)s1, (unsigned long long)s2,
                  (unsigned long long)s3, b0, b1, (const void*)base);
        return NPM_ERR_CUDA;
    }
    return NPM_OK;
}

// General 4-D fp32 tensor map (dims[0] contiguous; strides of dims 1..3 in ELEMENTS; any box).
int make_tensor_map_4d_box(CUtensorMap* tm, const float* base, const uint64_t dims[4], const uint64_t strides[3],
                           const uint32_t box[4], bool round_tf32, bool atom32b) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return NPM_ERR_CUDA;
    }
    cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t st[3] = {strides[0] * 4, strides[1] * 4, strides[2] * 4};
    cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = fn(tm, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                     const_cast<float*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) box=(%u,%u,%u,%u) base=%p",
                  (int)rc, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                  (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
                  (unsigned long long)strides[2], box[0], box[1], box[2], box[3], (const void*)base);
        return NPM_ERR_CUDA;
    }
    return NPM_OK;
}

// General fp32 tensor map of rank 3..5 (dims[0] contiguous; strides of dims 1.. in ELEMENTS; any box).
int make_tensor_map_nd(CUtensorMap* tm, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                       const uint32_t* box, bool round_tf32, bool atom32b) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return NPM_ERR_CUDA;
    }
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides[i] * 4;
    CUresult rc = fn(tm, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                     const_cast<float*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (rank %d) failed (%d): dims0..2=(%llu,%llu,%llu) box0..2=(%u,%u,%u) base=%p", rank,
                  (int)rc, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0],
                  box[1], box[2], (const void*)base);
        return NPM_ERR_CUDA;
    }
    return NPM_OK;
}

// K-major operand [rows, K contiguous] (leading dimension ld) as a 5-D tensor {32, rows, K/32, nb1, nb2}: one box
// {32, box_rows, 2} lands two consecutive 32-k slabs (each the usual 128B-swizzled K-major image).  Needs K % 32 == 0.
int make_tensor_map_k_chunked(CUtensorMap* tm, const float* base, uint64_t K, uint64_t rows, uint64_t nb1, uint64_t nb2,
                              uint64_t ld, uint64_t s2, uint64_t s3, uint32_t box_rows, bool round_tf32) {
    const uint64_t dims[5] = {32, rows, K / 32, nb1, nb2};
    const uint64_t strides[4] = {ld, 32, s2, s3};
    const uint32_t box[5] = {32, box_rows, 2, 1, 1};
    return make_tensor_map_nd(tm, base, 5, dims, strides, box, round_tf32, false);
}

// MN-major operand [K rows, MN contiguous] (leading dimension ld) as a 5-D tensor {32, K, MN/32, nb1, nb2}: one TMA box
// {32, box_k, chunks} lands `chunks` consecutive 4 KB swizzle-atom slabs — the same shared-memory image as `chunks`
// separate {32, box_k} boxes, in one TMA instruction.  Needs MN % 32 == 0.
int make_tensor_map_mn_chunked(CUtensorMap* tm, const float* base, uint64_t MN, uint64_t K, uint64_t nb1, uint64_t nb2,
                               uint64_t ld, uint64_t s2, uint64_t s3, uint32_t box_k, uint32_t chunks, bool round_tf32) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return NPM_ERR_CUDA;
    }
    cuuint64_t dims[5]    = {32, K, MN / 32, nb1, nb2};
    cuuint64_t strides[4] = {ld * 4, 32 * 4, s2 * 4, s3 * 4};
    cuuint32_t box[5]     = {32, box_k, chunks, 1, 1};
    cuuint32_t estr[5]    = {1, 1, 1, 1, 1};
    CUresult rc = fn(tm, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5,
                     const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (chunked MN-major) failed (%d): MN=%llu K=%llu ld=%llu", (int)rc,
                  (unsigned long long)MN, (unsigned long long)K, (unsigned long long)ld);
        return NPM_ERR_CUDA;
    }
    return NPM_OK;
}

// Split-bf16 planes [2][B, S, H, 64] (hi plane, then mid plane; token stride ld elements) as a 5-D bf16 tensor
// {64 d, S, H, B, 2}: box {64, box_rows, 1, 1, 1} = box_rows rows of 128 bytes, 128-byte swizzle — one image that the
// tensor core can read K-major (contraction over d) or MN-major (contraction over the rows).
int make_tensor_map_bf16_planes(CUtensorMap* tm, const void* base, uint64_t S, uint64_t H, uint64_t B, uint64_t ld,
                                uint64_t plane_elems, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return NPM_ERR_CUDA;
    }
    cuuint64_t dims[5]    = {64, S, H, B, 2};
    cuuint64_t strides[4] = {ld * 2, 64 * 2, S * ld * 2, plane_elems * 2};
    cuuint32_t box[5]     = {64, box_rows, 1, 1, 1};
    cuuint32_t estr[5]    = {1, 1, 1, 1, 1};
    CUresult rc = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (bf16 planes) failed (%d): S=%llu H=%llu B=%llu ld=%llu", (int)rc,
                  (unsigned long long)S, (unsigned long long)H, (unsigned long long)B, (unsigned long long)ld);
        return NPM_ERR_CUDA;
    }
    return NPM_OK;
}

// General bf16 tensor map of rank 3..5 (dims[0] contiguous; strides of dims 1.. in ELEMENTS); swizzle 0 (none), 64 or 128 bytes.
int make_tensor_map_bf16_nd(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                            const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return NPM_ERR_CUDA;
    }
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides_elems[i] * 2;
    CUresult rc = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (bf16, rank %d) failed (%d): dims0..2=(%llu,%llu,%llu) box0..2=(%u,%u,%u) base=%p", rank, (int)rc,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1], box[2], base);
        return NPM_ERR_CUDA;
    }
    return NPM_OK;
}

namespace {

template <int BN, bool AMN, bool BMN, int NP>
int launch_one(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmTcArgs& args,
               int grid, cudaStream_t stream) {
    using Cfg = TcCfg<BN, NP>;
    auto kern = gemm_tc_kernel<BN, AMN, BMN, NP>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
            return NPM_ERR_CUDA;
        }
        configured = true;
    }
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(a, b, c, args);
    count_launch();
    return check_launch("gemm_tc_kernel");
}

template <int BN, int NP>
int launch_major(bool amn, bool bmn, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                 const GemmTcArgs& args, int grid, cudaStream_t s) {
    if (!amn && !bmn) return launch_one<BN, false, false, NP>(a, b, c, args, grid, s);
    if (!amn && bmn) return launch_one<BN, false, true, NP>(a, b, c, args, grid, s);
    if (amn && !bmn) return launch_one<BN, true, false, NP>(a, b, c, args, grid, s);
    return launch_one<BN, true, true, NP>(a, b, c, args, grid, s);
}


template <int BN, bool AMN, bool BMN, int KS>
int launch_one2(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmTcArgs& args,
                int grid, cudaStream_t stream) {
    using Cfg = Tc2Cfg<BN, KS>;
    auto kern = gemm_tc2_kernel<BN, AMN, BMN, KS>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
            return NPM_ERR_CUDA;
        }
        configured = true;
    }
    GemmTcArgs largs = args;
    static const bool dbg_times = getenv("NPM_GEMM_DEBUG_TIMES") != nullptr;      // tools only: synchronises and prints
    if (dbg_times) {
        cudaMalloc(&largs.dbg, sizeof(long long) * 8 * grid);
        cudaMemset(largs.dbg, 0, sizeof(long long) * 8 * grid);
    }
    cudaError_t e = launch_pdl(kern, dim3((unsigned)grid, 1, 1), dim3(Cfg::kThreads, 1, 1), Cfg::kSmemBytes, stream, 2, a, b, c, largs);
    count_launch();
    if (e != cudaSuccess) { set_error("gemm_tc2_kernel launch: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
    if (dbg_times) {
        cudaStreamSynchronize(stream);
        std::vector<long long> h(8 * (size_t)grid);
        cudaMemcpy(h.data(), largs.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        double prod_empty = 0, w_full = 0, w_tempty = 0, total = 0, ks = 0, e_wait = 0, e_busy = 0;
        int leaders = 0;
        for (int cta = 0; cta < grid; ++cta) {
            prod_empty += (double)h[8 * cta]; e_wait += (double)h[8 * cta + 5]; e_busy += (double)h[8 * cta + 6];
            if ((cta & 1) == 0) { w_full += (double)h[8 * cta + 1]; w_tempty += (double)h[8 * cta + 2]; total += (double)h[8 * cta + 3]; ks += (double)h[8 * cta + 4]; ++leaders; }
        }
        fprintf(stderr, "[gemm_tc2 M=%d N=%d K=%d] issuer: %.0f clk total, %.0f k-steps per leader -> %.0f clk/k-step; waits per k-step: "
                "full %.0f, tmem-empty %.0f | producer empty-wait %.0f clk/k-step | epilogue warp0: tfull-wait %.0f, busy %.0f clk per CTA\n",
                args.M, args.N, args.K, total / leaders, ks / leaders, total / ks, w_full / ks, w_tempty / ks, prod_empty / (2 * ks),
                e_wait / grid, e_busy / grid);
        cudaFree(largs.dbg);
    }
    return check_launch("gemm_tc2_kernel");
}

template <int BN, int KS>
int launch_major2(bool amn, bool bmn, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                  const GemmTcArgs& args, int grid, cudaStream_t s) {
    if (!amn && !bmn) return launch_one2<BN, false, false, KS>(a, b, c, args, grid, s);
    if (!amn && bmn) return launch_one2<BN, false, true, KS>(a, b, c, args, grid, s);
    if (amn && !bmn) return launch_one2<BN, true, false, KS>(a, b, c, args, grid, s);
    return launch_one2<BN, true, true, KS>(a, b, c, args, grid, s);
}

inline bool mult4(int64_t v) { return (v & 3) == 0; }

}  // namespace

// Can the tensor-core path take this problem?  TMA needs 16-byte aligned bases and
// 16-byte multiples for every non-contiguous stride.
bool gemm_tc_supported(const npm_gemm_desc& d) {
    if (d.m <= 0 || d.n <= 0 || d.k <= 0) return false;
    if (!aligned16(d.a) || !aligned16(d.b) || !aligned16(d.c)) return false;
    const int64_t a_ld = d.a_cs == 1 ? d.a_rs : d.a_cs;
    const int64_t b_ld = d.b_rs == 1 ? d.b_cs : d.b_rs;
    if (!(d.a_cs == 1 || d.a_rs == 1) || !(d.b_cs == 1 || d.b_rs == 1)) return false;
    if (!mult4(a_ld) || !mult4(b_ld) || !mult4(d.ldc)) return false;
    if (a_ld <= 0 || b_ld <= 0 || d.ldc < d.n) return false;
    if (d.nb1 > 1 && !(mult4(d.a_bs1) && mult4(d.b_bs1) && mult4(d.c_bs1))) return false;
    if (d.nb2 > 1 && !(mult4(d.a_bs2) && mult4(d.b_bs2) && mult4(d.c_bs2))) return false;
    if (d.m > (1ll << 31) - 256 || d.n > (1ll << 31) - 256 || d.k > (1ll << 31) - 256) return false;
    return true;
}

static bool env_flag(const char* name, bool dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return v[0] != '0';
}

int gemm_tc_launch(const npm_gemm_desc& d, int precision, cudaStream_t stream) {
    const int nb1 = d.nb1 > 0 ? d.nb1 : 1, nb2 = d.nb2 > 0 ? d.nb2 : 1;
    // A is "MN-major" when m is the contiguous index of A[m,k]; when both strides are 1 the
    // matrix is a vector and either reading works — prefer K-major.
    const bool a_mn = !(d.a_cs == 1);
    const bool b_mn = !(d.b_rs == 1);   // B[k,n]: K-major when k is contiguous
    const int npass = precision == NPM_PREC_3XTF32 ? 3 : 1;
    // In single-pass mode let TMA round fp32 → tf32 (nearest) instead of the MMA truncating.
    static const bool tma_round = env_flag("NPM_TF32_TMA_ROUND", true);
    const bool round_ab = (npass == 1) && tma_round;

    // ---- tile shape and split-K: minimise waves x (main-loop + per-tile overhead), in SM cycles ----
    // Per K-step (32 fp32) cost of one CTA: the MMA floor is 2*BLOCK_N cycles; narrow tiles re-read
    // more operand bytes per flop from L2 and measured slower than that floor (profiles/r01_gemm_bench.txt).
    const int sms = num_sms();
    // CTA-pair kernel (256-row tiles) for single-pass TF32 whenever there are at least two row tiles
    static const bool pair_off = env_flag("NPM_GEMM_1CTA", false);
    const bool pair = (npass == 1) && !pair_off && d.m > kBlockM;
    const int tile_m = pair ? 2 * kBlockM : kBlockM;
    const int units = pair ? sms / 2 : sms;        // concurrently running tiles
    const int64_t tiles_m = (d.m + tile_m - 1) / tile_m;
    const int num_kb_total = (int)((d.k + kBlockK - 1) / kBlockK);
    const bool c_dense = (d.ldc == d.n) && (nb1 == 1 || d.c_bs1 == d.m * d.n) && (nb2 == 1 || d.c_bs2 == d.m * d.n * nb1);
    // environment switches are tools-only and read once per process (this function runs ~1000 times per training step)
    static const bool no_splitk = getenv("NPM_GEMM_NO_SPLITK") != nullptr;
    static const int forced_bn = getenv("NPM_GEMM_BLOCK_N_DYN") ? atoi(getenv("NPM_GEMM_BLOCK_N_DYN")) : 0;
    static const int dbg_mode_env = getenv("NPM_GEMM_DEBUG_MODE") ? atoi(getenv("NPM_GEMM_DEBUG_MODE")) : 0;
    static const int mn_lbo_env = getenv("NPM_MN_LBO") ? atoi(getenv("NPM_MN_LBO")) : 0;
    const bool may_split = c_dense && !(d.flags & (NPM_GEMM_RELU | NPM_GEMM_ACCUM)) && !no_splitk;
    int best_bn = 64, best_splits = 1;
    {
        const int forced = forced_bn;           // tuning hook (tools/gemm_bench.py)
        double best_cost = 1e300;
        const int cands[3] = {256, 128, 64};
        const double kstep[3] = {512.0, 400.0, 300.0};
        for (int i = 0; i < (pair ? 2 : 3); ++i) {
            const int bn = cands[i];
            if (forced && bn != forced) continue;
            const int64_t tn = (d.n + bn - 1) / bn;
            const int64_t tiles = tiles_m * tn * nb1 * nb2;
            for (int sp = 1; sp <= (may_split ? 16 : 1); sp *= 2) {
                const int kbs = (num_kb_total + sp - 1) / sp;
                if (sp > 1 && kbs < 16) break;
                const int64_t waves = (tiles * sp + units - 1) / units;
                double cost = double(waves) * (kbs * kstep[i] * npass + 1500.0 + 8.0 * bn);
                if (sp > 1) cost = cost * 1.08 + 3000.0;   // zero-fill + reduce traffic: split only for a clear win
                if (cost < best_cost - 1e-9) { best_cost = cost; best_bn = bn; best_splits = sp; }
            }
        }
    }
    const int bn = best_bn;
    static const bool chunk_off = env_flag("NPM_GEMM_NO_CHUNKED_MN", false);
    const bool a_chunked = pair && a_mn && !chunk_off && (d.m % 32 == 0);
    const bool b_chunked = pair && b_mn && !chunk_off && (d.n % 32 == 0);
    // 64-wide K stages (one 32 KB box per operand) when every operand has a chunked view.  Opt-in (NPM_GEMM_KS64=1):
    // measured equal or slightly slower than six 32-wide stages (FFN shapes 671/616 vs 678/639 TF) — three coarse
    // stages pipeline worse, which costs what the saved per-box overhead gains.
    static const bool ks64_on = env_flag("NPM_GEMM_KS64", false);
    const bool ks64 = pair && ks64_on && (a_mn ? a_chunked : d.k % 32 == 0) && (b_mn ? b_chunked : d.k % 32 == 0) && d.k >= 128;
    const uint32_t ks = ks64 ? 64 : kBlockK;
    const int num_kb_stage = (int)((d.k + ks - 1) / ks);        // ring stages along K
    const int kb_per_split = (num_kb_stage + best_splits - 1) / best_splits;
    const int splits = (num_kb_stage + kb_per_split - 1) / kb_per_split;
    const int64_t tiles_n = (d.n + bn - 1) / bn;
    const int64_t items_per_split = tiles_m * tiles_n * nb1 * nb2;
    const int64_t total = items_per_split * splits;
    if (total > (1ll << 30)) { set_error("gemm: too many tiles"); return NPM_ERR_INVALID; }
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(d.c, 0, sizeof(float) * (size_t)d.m * d.n * nb1 * nb2, stream);
        if (e != cudaSuccess) { set_error("gemm split-K memset: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
    }

    // ---- tensor maps ----
    // Probe knob: land K-major operands with the 32-byte-atom swizzle too and read them through a
    // SWIZZLE_128B_BASE32B descriptor (the attention kernels use one smem image of a [rows, d] tile
    // both as a K-major and as an MN-major operand).
    static const bool k_atom32 = env_flag("NPM_KMAJOR_ATOM32", false);
    CUtensorMap tmA, tmB, tmC;
    int rc;
    const uint64_t M = d.m, N = d.n, K = d.k;
    auto bs = [](int nb, int64_t s, uint64_t natural) -> uint64_t { return nb > 1 ? (uint64_t)s : natural; };
    if (!a_mn) {
        const uint64_t ld = d.a_rs;
        const uint64_t s2 = bs(nb1, d.a_bs1, ld * M), s3 = bs(nb2, d.a_bs2, s2 * nb1);
        if (ks64) rc = make_tensor_map_k_chunked(&tmA, d.a, K, M, nb1, nb2, ld, s2, s3, kBlockM, round_ab);
        else      rc = make_tensor_map_4d(&tmA, d.a, K, M, nb1, nb2, ld, s2, s3, kBlockK, kBlockM, round_ab, k_atom32);
    } else {
        const uint64_t ld = d.a_cs;
        const uint64_t s2 = bs(nb1, d.a_bs1, ld * K), s3 = bs(nb2, d.a_bs2, s2 * nb1);
        if (a_chunked) rc = make_tensor_map_mn_chunked(&tmA, d.a, M, K, nb1, nb2, ld, s2, s3, ks, kBlockM / 32, round_ab);
        else           rc = make_tensor_map_4d(&tmA, d.a, M, K, nb1, nb2, ld, s2, s3, 32, kBlockK, round_ab, true);
    }
    if (rc) return rc;
    if (!b_mn) {
        const uint64_t ld = d.b_cs;
        const uint64_t s2 = bs(nb1, d.b_bs1, ld * N), s3 = bs(nb2, d.b_bs2, s2 * nb1);
        if (ks64) rc = make_tensor_map_k_chunked(&tmB, d.b, K, N, nb1, nb2, ld, s2, s3, bn / 2, round_ab);
        else      rc = make_tensor_map_4d(&tmB, d.b, K, N, nb1, nb2, ld, s2, s3, kBlockK, pair ? bn / 2 : bn, round_ab, k_atom32);
    } else {
        const uint64_t ld = d.b_rs;
        const uint64_t s2 = bs(nb1, d.b_bs1, ld * K), s3 = bs(nb2, d.b_bs2, s2 * nb1);
        if (b_chunked) rc = make_tensor_map_mn_chunked(&tmB, d.b, N, K, nb1, nb2, ld, s2, s3, ks, bn / 2 / 32, round_ab);
        else           rc = make_tensor_map_4d(&tmB, d.b, N, K, nb1, nb2, ld, s2, s3, 32, kBlockK, round_ab, true);
    }
    if (rc) return rc;
    {
        const uint64_t ld = d.ldc;
        const uint64_t s2 = bs(nb1, d.c_bs1, ld * M), s3 = bs(nb2, d.c_bs2, s2 * nb1);
        rc = make_tensor_map_4d(&tmC, d.c, N, M, nb1, nb2, ld, s2, s3, 32, 32, false, false);
    }
    if (rc) return rc;

    GemmTcArgs args;
    args.M = (int)d.m; args.N = (int)d.n; args.K = (int)d.k;
    args.tiles_m = (int)tiles_m; args.tiles_n = (int)tiles_n; args.nb1 = nb1;
    args.total_tiles = (int)total;
    args.splits = splits; args.kb_per_split = kb_per_split; args.items_per_split = (int)items_per_split;
    args.alpha = d.alpha;
    args.bias = d.bias;
    args.residual = d.residual;
    args.ldr = d.ldr;
    args.dbg = nullptr;
    {
        static const int band_env = getenv("NPM_GEMM_BAND") ? atoi(getenv("NPM_GEMM_BAND")) : 8;
        args.band_m = (pair && tiles_m > band_env && tiles_n > 1) ? band_env : 0;
    }
    args.dbg_mode = dbg_mode_env;
    args.a_chunked = a_chunked ? 1 : 0;
    args.b_chunked = b_chunked ? 1 : 0;
    args.relu = (d.flags & NPM_GEMM_RELU) ? 1 : 0;
    args.accum = (d.flags & NPM_GEMM_ACCUM) ? 1 : 0;
    // Shared-memory descriptors.
    //  K-major operand : TMA box {32 k, rows} with the 128B swizzle (16 B atoms) lands rows of 128 B;
    //                    UMMA layout SWIZZLE_128B, 8-row groups 1024 B apart (SBO), LBO unused (1).
    //  MN-major operand: fp32/tf32 MN-major only exists with 32-byte swizzle atoms (a 4-byte element
    //                    cannot be transposed at 16 B granularity): TMA SWIZZLE_128B_ATOM_32B lands
    //                    each {32 mn, 32 k} box as 32 K-rows of 128 B; UMMA layout
    //                    SWIZZLE_128B_BASE32B, 32-element MN chunks 4096 B apart (LBO), 4-row K groups
    //                    512 B apart (SBO).
    const uint32_t mn_lbo = mn_lbo_env ? (uint32_t)mn_lbo_env : ks * 128u;   // one {32 mn, ks k} chunk
    static const uint32_t mn_sbo = getenv("NPM_MN_SBO") ? (uint32_t)atoi(getenv("NPM_MN_SBO")) : 512u;
    const uint64_t desc_k  = k_atom32 ? ptx::umma_desc_base(1, 16, 1024) : ptx::umma_desc_base(2 /*SWIZZLE_128B*/, 16, 1024);
    const uint64_t desc_mn = ptx::umma_desc_base(1 /*SWIZZLE_128B_BASE32B*/, mn_lbo, mn_sbo);
    args.desc_a = a_mn ? desc_mn : desc_k;
    args.desc_b = b_mn ? desc_mn : desc_k;
    if (pair) {
        const int grid2 = 2 * (int)(total < units ? total : units);
        if (ks64) {
            if (bn == 256) return launch_major2<256, 64>(a_mn, b_mn, tmA, tmB, tmC, args, grid2, stream);
            return launch_major2<128, 64>(a_mn, b_mn, tmA, tmB, tmC, args, grid2, stream);
        }
        if (bn == 256) return launch_major2<256, 32>(a_mn, b_mn, tmA, tmB, tmC, args, grid2, stream);
        return launch_major2<128, 32>(a_mn, b_mn, tmA, tmB, tmC, args, grid2, stream);
    }
    const int grid = (int)(total < sms ? total : sms);

    if (npass == 1) {
        if (bn == 256) return launch_major<256, 1>(a_mn, b_mn, tmA, tmB, tmC, args, grid, stream);
        if (bn == 128) return launch_major<128, 1>(a_mn, b_mn, tmA, tmB, tmC, args, grid, stream);
        return launch_major<64, 1>(a_mn, b_mn, tmA, tmB, tmC, args, grid, stream);
    } else {
        if (bn == 256) return launch_major<256, 3>(a_mn, b_mn, tmA, tmB, tmC, args, grid, stream);
        if (bn == 128) return launch_major<128, 3>(a_mn, b_mn, tmA, tmB, tmC, args, grid, stream);
        return launch_major<64, 3>(a_mn, b_mn, tmA, tmB, tmC, args, grid, stream);
    }
}

}  // namespace npm
