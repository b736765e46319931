// gemm_bx.cu — split-bf16 GEMM on tcgen05 (kind::f16) for fp32 operands: the "bf16x3" contraction mode.
//
//   C[z][m,n] (=|+=) alpha * sum_k A[z][m,k] * B[z][k,n] (+ bias[n]) (+ residual) (ReLU)       (fp32 in HBM, fp32 out)
//
// Every fp32 operand element x is split inside shared memory into two bf16 numbers
//   hi = bf16_rn(x),  mid = bf16_rn(x - hi)          (x - hi - mid is below 2^-17 |x|)
// and the tensor cores run the three products  mid*hi + hi*mid + hi*hi  with fp32 accumulation in TMEM.  The dropped
// terms (mid*mid and the two residuals) are <= 3 * 2^-18 of |a||b| per product, i.e. ~50x below a single TF32 pass
// (2^-11), which is what north_star's rtol 1e-3 / atol 1e-4 needs — at 1.5x the tensor time of one TF32 pass instead
// of the 3x of 3xTF32 (kind::f16 runs K = 16 per instruction where kind::tf32 runs K = 8).  NTERMS = 1 keeps only
// hi*hi: plain bf16 compute on fp32 storage.
//
// The reference contractions this serves are the same as gemm_tc.cu: Linear fwd/dX/dW (layers/mlp.py:21-40), the
// MultiHeadAttention projections and their gradients (layers/attentions.py:88-100,116,129-135,169-184).
//
// Structure = the CTA-pair kernel of gemm_tc.cu (cta_group::2, 256 x BLOCK_N tile per pair, two TMEM accumulator
// buffers, four epilogue warps per CTA) with a conversion step between TMA and the tensor core:
//   warp 4      TMA producer : lands this CTA's fp32 operand tiles (128-byte swizzle) in a ring stage
//   warps 5-12  converters   : LDS the stage into registers (256 threads x 8 float4 = the whole stage), barrier, and
//                              store the bf16 hi / mid images IN PLACE in the canonical UMMA layouts:
//     K-major operand  [rows, 32 k]: one 128-byte row per operand row = [hi k0..31 | mid k0..31], 128B swizzle;
//                                    an MMA K16 slice is a 32-byte column of that row (hi: 0,32; mid: 64,96)
//     MN-major operand [32 k, rows]: hi image then mid image, each rows/64 slabs of 32 k-rows x 128 B (64 mn), 128B
//                                    swizzle; an MMA K16 slice is 16 k-rows = 2048 B of a slab
//                              then arrive on their CTA's barrier
//   warp 13     MMA issuer (leader CTA) / proxy-fence relay (peer CTA)
// Why not convert on the way in from global memory (the first version of this kernel: LDG.128 into registers, two
// stages in flight per thread, split, STS)?  Measured (NPM_GEMM_DEBUG_TIMES): the LSU path accepts ~18 B/clk/SM of
// loads under load — issuing the 8 LDG.128 of a stage stalled 1800 clk — against the 42 B/clk/SM this kernel needs
// and the 55-64 B/clk/SM TMA sustains; 320-350 TF.  (Before that: fence.proxy.async lowers to MEMBAR.ALL.CTA +
// FENCE.VIEW.ASYNC, and a MEMBAR in a thread with global loads in flight waits for them: 200 TF.)
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "bx_convert.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace npm {

int make_tensor_map_4d(CUtensorMap* tm, const float* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3,
                       uint64_t s1, uint64_t s2, uint64_t s3, uint32_t b0, uint32_t b1, bool round_tf32, bool atom32b);
int make_tensor_map_nd(CUtensorMap* tm, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                       const uint32_t* box, bool round_tf32, bool atom32b);

namespace {

constexpr int kBM = 128;        // rows of A per CTA (256 per pair)
constexpr int kKS = 32;         // fp32 elements of K per ring stage
using bx::kConvWarps;
using bx::OperandConverter;
using bx::bar_sync_conv;
using bx::mbar_arrive_remote;
constexpr int kTmaWarp = 4;
constexpr int kFirstConvWarp = 5;
constexpr int kMmaWarp = kFirstConvWarp + kConvWarps;                    // 13
constexpr int kThreadsBx = 32 * (kMmaWarp + 1);                          // 448: 4 epilogue, TMA, 8 converters, MMA issuer

struct GemmBxArgs {
    int a_chunked, b_chunked;                  // MN-major operand described as a 5-D {32, K, MN/32, nb1, nb2} tensor: one box per stage
    int M, N, K;
    int tiles_m, tiles_n, nb1, total_tiles;
    int splits, kb_per_split, items_per_split;
    int band_m;
    float alpha;
    const float* bias;
    const float* residual;
    int64_t ldr;
    int relu, accum;
    int c_planes;              // the result is written as bf16 hi / mid planes (tmC is a bf16 {N, M, 2 planes} map) instead of fp32
    float* a_colsum;           // optional [M] (A MN-major only): += sum over k of A[m, k] — the bias gradient of a dW GEMM
    const float* rowdot_x;     // optional (c_planes only): [M, N] fp32, leading dimension rowdot_ld; the epilogue writes
    int64_t rowdot_ld;         // rowdot_out[(row / rowdot_seq) * (N / 64) + col / 64][row % rowdot_seq] = sum over the 64 columns of
    float* rowdot_out;         // the group of C[row, .] * rowdot_x[row, .] — D = rowsum(dO o O) of the attention backward
    int rowdot_seq;
    long long* dbg;            // tools only (NPM_GEMM_DEBUG_TIMES): 16 wait-cycle counters per CTA
};

template <int BLOCK_N>
struct BxCfg {
    static constexpr int kHalfN      = BLOCK_N / 2;
    static constexpr int kABytes     = kBM * 128;                      // 16 KB: hi + mid of a 128 x 32 tile
    static constexpr int kBBytes     = kHalfN * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBytes   = 4 * 2 * 4096;
    static constexpr int kBarBytes   = 512;
    static constexpr int kBudget     = 232448 - 1024;
    static constexpr int kStagesRaw  = (kBudget - kEpiBytes - kBarBytes - 1024) / kStageBytes;
    static constexpr int kStages     = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kCsBytes    = 1024;                           // column-sum exchange between converter warp pairs
    static constexpr int kSmemBytes  = kStages * kStageBytes + kEpiBytes + kBarBytes + kCsBytes + 1024;
    static constexpr int kTmemCols   = 2 * BLOCK_N;
};

__device__ __forceinline__ void tile_coords_bx(int r, int tiles_m, int tiles_n, int band, int& tm, int& tn) {
    if (band <= 1) { tm = r % tiles_m; tn = r / tiles_m; return; }
    const int per_band = band * tiles_n;
    const int b = r / per_band, idx = r - b * per_band;
    const int width = min(band, tiles_m - b * band);
    tm = b * band + idx % width;
    tn = idx / width;
}

// B_PRE: the B operand (a weight matrix) arrives already split — bf16 hi / mid planes written once per step by
// split_planes_kernel — and is landed by TMA directly as the image the tensor core reads: no LDS / STS for it, which
// is what bounds this kernel (the shared-memory data pipe: 32 KB LDS + 32 KB STS + 48 KB of MMA operand reads per
// stage = 896 wavefronts + arbitration against the 768 clk of its six MMAs, profiles/r02_ncu_gemm_bx_ffn.txt).
//   K-major  B_PRE: box {32 k, rows, 2 planes}, 64-byte swizzle: hi image rows x 64 B, then the mid image
//   MN-major B_PRE: box {64 n, 32 k, rows/64 slabs, 2 planes}, 128-byte swizzle: the same image the converters write
// A_PRE: the same for the A operand — an activation whose producer wrote it as bf16 planes in the first place (the FFN
// hidden activation out of the first FFN GEMM's epilogue, its gradient out of the ReLU backward kernel): with both
// operands pre-split the converters only pass the TMA completion on and the main loop is TMA -> MMA.
template <int BLOCK_N, bool A_MN, bool B_MN, int NTERMS, bool B_PRE, bool A_PRE = false>
__global__ void __launch_bounds__(kThreadsBx, 1)
gemm_bx_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmBxArgs args) {
    using Cfg = BxCfg<BLOCK_N>;
    constexpr int S = Cfg::kStages;
    pdl_trigger();

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr  = ptx::smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr        = smem_raw + (base_addr - raw_addr);

    const uint32_t stage_addr = base_addr;
    const uint32_t epi_addr   = base_addr + S * Cfg::kStageBytes;
    const uint32_t bar_addr   = epi_addr + Cfg::kEpiBytes;
    auto tma_bar    = [&](int s) { return bar_addr + 8u * s; };                 // this CTA's fp32 tiles of stage s have landed
    auto empty_bar  = [&](int s) { return bar_addr + 8u * (S + s); };           // the MMAs that read stage s have completed
    auto conv_bar   = [&](int s) { return bar_addr + 8u * (2 * S + s); };       // this CTA's converters have stored stage s
    auto peer_bar   = [&](int s) { return bar_addr + 8u * (3 * S + s); };       // leader only: the peer's stage s is stored and fenced
    auto tfull_bar  = [&](int a) { return bar_addr + 8u * (4 * S + a); };
    auto tempty_bar = [&](int a) { return bar_addr + 8u * (4 * S + 2 + a); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(
        base_ptr + S * Cfg::kStageBytes + Cfg::kEpiBytes + 8 * (4 * S + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();          // 0 = pair leader
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_kb = (args.K + kKS - 1) / kKS;

    if (warp == kTmaWarp && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        ptx::prefetch_tensormap(&tmC);
    }
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) {
                ptx::mbar_init(tma_bar(s), 1);                     // this CTA's arrive.expect_tx
                ptx::mbar_init(empty_bar(s), 1);                   // one multicast commit
                ptx::mbar_init(conv_bar(s), kConvWarps);           // one arrive per converter warp of this CTA
                ptx::mbar_init(peer_bar(s), 1);                    // the peer relay's remote arrive
            }
            for (int a = 0; a < 2; ++a) {
                ptx::mbar_init(tfull_bar(a), 1);
                ptx::mbar_init(tempty_bar(a), 8);
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
    pdl_wait();

    const int tiles_mn = args.tiles_m * args.tiles_n;      // tiles_m counts 256-row tiles

    if (warp == kTmaWarp) {
        // ============================ TMA producer (both CTAs): fp32 tiles ============================
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
                tile_coords_bx(r, args.tiles_m, args.tiles_n, args.band_m, tm, tn);
                const int m0 = tm * (2 * kBM) + (int)rank * kBM;
                const int n0 = tn * BLOCK_N + (int)rank * Cfg::kHalfN;
                const int z1 = z % args.nb1, z2 = z / args.nb1;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                    const uint32_t sB = sA + Cfg::kABytes;
                    const uint32_t fb = tma_bar(stage);
                    ptx::mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
                    const int k0 = kb * kKS;
                    if (A_PRE) {
                        if (!A_MN) ptx::tma_load_4d(sA, &tmA, fb, k0, m0, 0, 0);           // {k, rows, plane, 1}
                        else       ptx::tma_load_4d(sA, &tmA, fb, 0, k0, m0 / 64, 0);      // {64 m, k, slab, plane}
                    } else if (!A_MN) {
                        ptx::tma_load_4d(sA, &tmA, fb, k0, m0, z1, z2);
                    } else if (args.a_chunked) {
                        ptx::tma_load_5d(sA, &tmA, fb, 0, k0, m0 / 32, z1, z2);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBM / 32; ++c) ptx::tma_load_4d(sA + c * 4096, &tmA, fb, m0 + c * 32, k0, z1, z2);
                    }
                    if (B_PRE) {
                        if (!B_MN) ptx::tma_load_4d(sB, &tmB, fb, k0, n0, 0, 0);           // {k, rows, plane, 1}
                        else       ptx::tma_load_4d(sB, &tmB, fb, 0, k0, n0 / 64, 0);      // {64 n, k, slab, plane}
                    } else if (!B_MN) {
                        ptx::tma_load_4d(sB, &tmB, fb, k0, n0, z1, z2);
                    } else if (args.b_chunked) {
                        ptx::tma_load_5d(sB, &tmB, fb, 0, k0, n0 / 32, z1, z2);
                    } else {
#pragma unroll
                        for (int c = 0; c < Cfg::kHalfN / 32; ++c) ptx::tma_load_4d(sB + c * 4096, &tmB, fb, n0 + c * 32, k0, z1, z2);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= kFirstConvWarp && warp < kMmaWarp) {
        // ===================== converters: fp32 stage -> registers -> bf16 hi / mid images in place =====================
        const int pw = warp - kFirstConvWarp;
        OperandConverter<kBM, A_MN, NTERMS> ca;
        OperandConverter<Cfg::kHalfN, B_MN, NTERMS> cb;
        ca.init(pw, lane);
        cb.init(pw, lane);
        // Bias gradient folded into the dW GEMM (layers/attentions.py:129-135, mlp.py:34: db = sum over tokens of dy): for
        // dW = dy^T x the A operand IS dy, MN-major, and every element of it passes through these registers as fp32 —
        // a thread always holds the same four columns (mn = 32 * (pw / 2) + 4 * (lane & 7) ..+3), so one float4
        // accumulates them; the tiles of the first column of N tiles add their sums to args.a_colsum (zeroed by the
        // launcher; red.global.add, the splits of a split-K launch each contribute their rows).
        const bool do_colsum = A_MN && !A_PRE && args.a_colsum != nullptr;
        float4* cs_smem = reinterpret_cast<float4*>(base_ptr + S * Cfg::kStageBytes + Cfg::kEpiBytes + Cfg::kBarBytes);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < args.total_tiles; tile += num_clusters) {
            const int sp = tile / args.items_per_split, t2i = tile - sp * args.items_per_split;
            const int kb0 = sp * args.kb_per_split;
            const int nkb = min(num_kb, kb0 + args.kb_per_split) - kb0;
            int tm = 0, tn = 1;
            if (do_colsum) tile_coords_bx(t2i % tiles_mn, args.tiles_m, args.tiles_n, args.band_m, tm, tn);
            const bool sum_tile = do_colsum && tn == 0;
            float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int it = 0; it < nkb; ++it) {
            const bool dbg = args.dbg != nullptr && pw == 0 && lane == 0;
            long long t0 = 0, t1 = 0, t2 = 0;
            if (dbg) t0 = clock64();
            ptx::mbar_wait(tma_bar(stage), phase);
            if (dbg) t1 = clock64();
            const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
            float4 va[OperandConverter<kBM, A_MN, NTERMS>::NLD];
            float4 vb[OperandConverter<Cfg::kHalfN, B_MN, NTERMS>::NLD];
            if (!A_PRE) ca.load(va, sA);
            if (!B_PRE) cb.load(vb, sA + Cfg::kABytes);
            if (!(A_PRE && B_PRE)) bar_sync_conv();          // every fp32 chunk of the stage is in registers: overwrite it
            if (dbg) t2 = clock64();
            if (!A_PRE) ca.store(va, sA);
            if (!B_PRE) cb.store(vb, sA + Cfg::kABytes);
            // no proxy fence here: fence.proxy.async lowers to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and the MEMBAR costs
            // ~36 clk per store in flight (700 clk per stage measured with 16 STS per thread); the consumer of this
            // barrier — a thread with nothing in flight — executes it (issuer / peer relay below)
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(conv_bar(stage));
            if (!A_PRE && sum_tile) {
#pragma unroll
                for (int i = 0; i < OperandConverter<kBM, A_MN, NTERMS>::NLD; ++i) {
                    csum.x += va[i].x; csum.y += va[i].y; csum.z += va[i].z; csum.w += va[i].w;
                }
            }
            if (dbg) {
                const long long t3 = clock64();
                args.dbg[blockIdx.x * 16 + 0] += t1 - t0;            // wait for the TMA tiles
                args.dbg[blockIdx.x * 16 + 1] += t2 - t1;            // LDS + barrier
                args.dbg[blockIdx.x * 16 + 2] += t3 - t2;            // split + STS + fence + arrive
                args.dbg[blockIdx.x * 16 + 3] += 1;
            }
            if (++stage == S) { stage = 0; phase ^= 1u; }
        }
            if (do_colsum) {       // uniform over the 256 converter threads of the CTA (tile coordinates are CTA-wide)
                if (sum_tile) {
                    // rows of one warp instruction sit in the lane groups q = lane >> 3: fold them, then the warp pair
                    // (pw, pw ^ 1) that shares the 32-column chunk through shared memory
#pragma unroll
                    for (int o = 8; o <= 16; o <<= 1) {
                        csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o); csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
                        csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o); csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
                    }
                    if (lane < 8) cs_smem[pw * 8 + lane] = csum;
                }
                bar_sync_conv();
                if (sum_tile && (pw & 1) == 0 && lane < 8) {
                    const float4 o = cs_smem[(pw + 1) * 8 + lane];
                    const int m = tm * (2 * kBM) + (int)rank * kBM + (pw >> 1) * 32 + 4 * lane;
                    float* dst = args.a_colsum + m;
                    if (m < args.M)     atomicAdd(dst, csum.x + o.x);
                    if (m + 1 < args.M) atomicAdd(dst + 1, csum.y + o.y);
                    if (m + 2 < args.M) atomicAdd(dst + 2, csum.z + o.z);
                    if (m + 3 < args.M) atomicAdd(dst + 3, csum.w + o.w);
                }
                bar_sync_conv();       // cs_smem is reused by the next tile
            }
        }
    } else if (warp == kMmaWarp && rank == 1) {
        // ============ peer CTA: proxy-fence relay — the fence sits on the causality path stores -> conv_bar -> fence ->
        // peer_bar -> MMA, which is what the PTX memory model asks of a proxy fence ============
        if (ptx::elect_one()) {
            int total_it = 0;
            for (int tile = cluster_id; tile < args.total_tiles; tile += num_clusters) {
                const int kb0 = (tile / args.items_per_split) * args.kb_per_split;
                total_it += min(num_kb, kb0 + args.kb_per_split) - kb0;
            }
            const uint32_t peer_leader = ptx::mapa(peer_bar(0), 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < total_it; ++it) {
                ptx::mbar_wait(conv_bar(stage), phase);
                ptx::fence_proxy_async_smem();
                mbar_arrive_remote(peer_leader + 8u * stage);
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == kMmaWarp) {
        // ============================= MMA issuer (leader CTA) =============================
        if (rank == 0 && ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * kBM, BLOCK_N, A_MN, B_MN);
            constexpr uint64_t desc_k  = ptx::umma_desc_base(2 /*SWIZZLE_128B*/, 16, 1024);
            constexpr uint64_t desc_mn = ptx::umma_desc_base(2 /*SWIZZLE_128B*/, 4096, 1024);   // LBO: next 64-mn slab, SBO: next 8 k-rows
            constexpr uint64_t desc_k64 = ptx::umma_desc_base(4 /*SWIZZLE_64B*/, 16, 512);      // pre-split K-major: 64-byte rows, 8-row groups 512 B apart
            constexpr uint64_t descA = A_MN ? desc_mn : (A_PRE ? desc_k64 : desc_k), descB = B_MN ? desc_mn : (B_PRE ? desc_k64 : desc_k);
            // byte offset of K16 slice s of the hi (t = 0) / mid (t = 1) image
            auto a_off = [](int t, int s) -> uint32_t {
                return A_MN ? uint32_t(t) * (kBM / 64) * 4096u + uint32_t(s) * 2048u
                            : (A_PRE ? uint32_t(t) * (kBM * 64u) + uint32_t(s) * 32u : uint32_t(t) * 64u + uint32_t(s) * 32u);
            };
            auto b_off = [](int t, int s) -> uint32_t {
                return B_MN ? uint32_t(t) * (Cfg::kHalfN / 64) * 4096u + uint32_t(s) * 2048u
                            : (B_PRE ? uint32_t(t) * (Cfg::kHalfN * 64u) + uint32_t(s) * 32u : uint32_t(t) * 64u + uint32_t(s) * 32u);
            };
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = cluster_id; tile < args.total_tiles; tile += num_clusters) {
                const int kb0 = (tile / args.items_per_split) * args.kb_per_split;
                const int kb1 = min(num_kb, kb0 + args.kb_per_split);
                {
                    const long long t0 = args.dbg ? clock64() : 0;
                    ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    if (args.dbg) args.dbg[blockIdx.x * 16 + 12] += clock64() - t0;   // issuer: wait for a free accumulator
                }
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const long long t0 = args.dbg ? clock64() : 0;
                    ptx::mbar_wait(conv_bar(stage), phase);
                    ptx::mbar_wait_cluster(peer_bar(stage), phase);
                    ptx::fence_proxy_async_smem();                      // this CTA's converted stage -> async proxy
                    ptx::tc_fence_after();
                    if (args.dbg) {
                        args.dbg[blockIdx.x * 16 + 8] += clock64() - t0;     // issuer: wait for a converted stage
                        args.dbg[blockIdx.x * 16 + 11] += 1;
                    }
                    const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                    const uint32_t sB = sA + Cfg::kABytes;
                    uint32_t accum = (kb != kb0) ? 1u : 0u;
                    if (NTERMS == 3) {
#pragma unroll
                        for (int s = 0; s < 2; ++s) {          // mid * hi
                            ptx::umma_f16_2sm(d_tmem, ptx::umma_desc(descA, sA + a_off(1, s)), ptx::umma_desc(descB, sB + b_off(0, s)), idesc, accum);
                            accum = 1u;
                        }
#pragma unroll
                        for (int s = 0; s < 2; ++s)            // hi * mid
                            ptx::umma_f16_2sm(d_tmem, ptx::umma_desc(descA, sA + a_off(0, s)), ptx::umma_desc(descB, sB + b_off(1, s)), idesc, 1u);
                    }
#pragma unroll
                    for (int s = 0; s < 2; ++s) {              // hi * hi
                        ptx::umma_f16_2sm(d_tmem, ptx::umma_desc(descA, sA + a_off(0, s)), ptx::umma_desc(descB, sB + b_off(0, s)), idesc, accum);
                        accum = 1u;
                    }
                    ptx::umma_commit_2sm(empty_bar(stage), 3);
                    if (kb == kb1 - 1) ptx::umma_commit_2sm(tfull_bar(acc), 3);
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
            }
        }
    } else if (warp < 4) {
        // ============================== epilogue (both CTAs; same as gemm_tc2_kernel) ==============================
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
            tile_coords_bx(r, args.tiles_m, args.tiles_n, args.band_m, tm, tn);
            const int m0 = tm * (2 * kBM) + (int)rank * kBM;
            const int n0 = tn * BLOCK_N;
            const int z1 = z % args.nb1, z2 = z / args.nb1;
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
            const bool rows_live = (m0 + warp * 32) < args.M;
            float rowdot_acc = 0.0f;            // the even chunk's half of a 64-column row dot (rowdot_x)
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
                if (args.c_planes && args.rowdot_x != nullptr) {
                    // this thread holds 32 consecutive fp32 results of row row_g: their dot with the same 32 elements of
                    // rowdot_x; two consecutive chunks are one 64-column group (N % 64 == 0 is checked by the launcher)
                    const int64_t row_g = (int64_t)m0 + warp * 32 + lane;
                    if (row_g < args.M) {
                        const float4* xr = reinterpret_cast<const float4*>(args.rowdot_x + row_g * args.rowdot_ld + nc);
                        float p0 = 0.0f, p1 = 0.0f;
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {
                            const float4 a = __ldg(xr + j), b = __ldg(xr + j + 1);
                            p0 += (f[4 * j] * a.x + f[4 * j + 1] * a.y) + (f[4 * j + 2] * a.z + f[4 * j + 3] * a.w);
                            p1 += (f[4 * j + 4] * b.x + f[4 * j + 5] * b.y) + (f[4 * j + 6] * b.z + f[4 * j + 7] * b.w);
                        }
                        if ((chunk & 1) == 0) {
                            rowdot_acc = p0 + p1;
                        } else {
                            const int64_t bi = row_g / args.rowdot_seq, si = row_g - bi * args.rowdot_seq;
                            args.rowdot_out[(bi * (args.N >> 6) + (nc >> 6)) * args.rowdot_seq + si] = rowdot_acc + (p0 + p1);
                        }
                    }
                }
                if (args.c_planes) {
                    // split-bf16 output (the q | k | v projection feeding the fused attention): hi rows of 64 B in the first
                    // half of the staging buffer, mid rows in the second, two TMA stores into the planes
                    uint8_t* hrow = stg_base + buf * 4096 + lane * 64;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t h[4], m[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float a = f[8 * j + 2 * q], b = f[8 * j + 2 * q + 1];
                            h[q] = bx::bf16x2_rn(a, b);
                            m[q] = bx::bf16x2_rn(a - __uint_as_float(h[q] << 16), b - __uint_as_float(h[q] & 0xffff0000u));
                        }
                        *reinterpret_cast<uint4*>(hrow + j * 16) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(hrow + 2048 + j * 16) = make_uint4(m[0], m[1], m[2], m[3]);
                    }
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    if (ptx::elect_one()) {
                        ptx::tma_store_4d(&tmC, stg_addr + buf * 4096, nc, m0 + warp * 32, 0, 0);
                        ptx::tma_store_4d(&tmC, stg_addr + buf * 4096 + 2048, nc, m0 + warp * 32, 1, 0);
                        ptx::tma_commit_group();
                    }
                    ++nstore;
                    continue;
                }
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
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) ptx::mbar_arrive(tempty_bar(acc));
                else           mbar_arrive_remote(ptx::mapa(tempty_bar(acc), 0));
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (lane == 0) ptx::tma_wait_group<0>();
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == kMmaWarp) ptx::tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

template <int BN, bool AMN, bool BMN, int NT, bool BPRE, bool APRE = false>
int launch_bx(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmBxArgs& args, int grid, cudaStream_t stream) {
    using Cfg = BxCfg<BN>;
    auto kern = gemm_bx_kernel<BN, AMN, BMN, NT, BPRE, APRE>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gemm_bx, smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
            return NPM_ERR_CUDA;
        }
        configured = true;
    }
    GemmBxArgs largs = args;
    static const bool dbg_times = getenv("NPM_GEMM_DEBUG_TIMES") != nullptr;      // tools only: synchronises and prints
    if (dbg_times) {
        cudaMalloc(&largs.dbg, sizeof(long long) * 16 * grid);
        cudaMemset(largs.dbg, 0, sizeof(long long) * 16 * grid);
    }
    cudaError_t e = launch_pdl(kern, dim3((unsigned)grid, 1, 1), dim3(kThreadsBx, 1, 1), Cfg::kSmemBytes, stream, 2, a, b, c, largs);
    count_launch();
    if (e != cudaSuccess) { set_error("gemm_bx_kernel launch: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
    if (dbg_times) {
        cudaStreamSynchronize(stream);
        std::vector<long long> h(16 * (size_t)grid);
        cudaMemcpy(h.data(), largs.dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        double s[16] = {0};
        for (int cta = 0; cta < grid; ++cta)
            for (int i = 0; i < 16; ++i) s[i] += (double)h[16 * cta + i];
        fprintf(stderr, "[gemm_bx M=%d N=%d K=%d nt=%d] cycles per stage: converter(w0) tma-wait %.0f, LDS+barrier %.0f, split+STS+fence+arrive %.0f | "
                "issuer stage-wait %.0f, acc-wait %.0f\n",
                args.M, args.N, args.K, NT, s[0] / s[3], s[1] / s[3], s[2] / s[3], s[8] / s[11], s[12] / s[11]);
        cudaFree(largs.dbg);
    }
    return check_launch("gemm_bx_kernel");
}

template <int BN, bool AMN, bool BMN>
int launch_bx_pre(bool apre, bool bpre, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const GemmBxArgs& args, int grid,
                  cudaStream_t s) {
    if (apre && bpre) return launch_bx<BN, AMN, BMN, 3, true, true>(a, b, c, args, grid, s);
    if (apre) return launch_bx<BN, AMN, BMN, 3, false, true>(a, b, c, args, grid, s);
    if (bpre) return launch_bx<BN, AMN, BMN, 3, true, false>(a, b, c, args, grid, s);
    return launch_bx<BN, AMN, BMN, 3, false, false>(a, b, c, args, grid, s);
}
template <int BN, int NT>
int launch_bx_major(bool amn, bool bmn, bool apre, bool bpre, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                    const GemmBxArgs& args, int grid, cudaStream_t s) {
    if (NT == 3) {
        if (!amn && !bmn) return launch_bx_pre<BN, false, false>(apre, bpre, a, b, c, args, grid, s);
        if (!amn && bmn) return launch_bx_pre<BN, false, true>(apre, bpre, a, b, c, args, grid, s);
        if (amn && !bmn) return launch_bx_pre<BN, true, false>(apre, bpre, a, b, c, args, grid, s);
        return launch_bx_pre<BN, true, true>(apre, bpre, a, b, c, args, grid, s);
    }
    if (!amn && !bmn) return launch_bx<BN, false, false, 1, false>(a, b, c, args, grid, s);
    if (!amn && bmn) return launch_bx<BN, false, true, 1, false>(a, b, c, args, grid, s);
    if (amn && !bmn) return launch_bx<BN, true, false, 1, false>(a, b, c, args, grid, s);
    return launch_bx<BN, true, true, 1, false>(a, b, c, args, grid, s);
}

}  // namespace

// The split-bf16 kernel takes every problem the TMA GEMM takes that has at least two row tiles; everything else in
// these modes runs the TF32 kernels of the same or better accuracy class (3xTF32 for bf16x3).
bool gemm_tc_supported(const npm_gemm_desc& d);
bool gemm_bx_supported(const npm_gemm_desc& d) { return d.m > kBM && gemm_tc_supported(d); }

int make_tensor_map_bf16_nd(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                            const uint32_t* box, int swizzle_bytes);

int gemm_bx_launch(const npm_gemm_desc& d, int nterms, cudaStream_t stream) {
    const int nb1 = d.nb1 > 0 ? d.nb1 : 1, nb2 = d.nb2 > 0 ? d.nb2 : 1;
    const bool a_mn = !(d.a_cs == 1), b_mn = !(d.b_rs == 1);
    const int units = num_sms() / 2;
    const int64_t tiles_m = (d.m + 2 * kBM - 1) / (2 * kBM);
    const int num_kb_total = (int)((d.k + kKS - 1) / kKS);
    const bool c_dense = (d.ldc == d.n) && (nb1 == 1 || d.c_bs1 == d.m * d.n) && (nb2 == 1 || d.c_bs2 == d.m * d.n * nb1);
    static const bool no_splitk = getenv("NPM_GEMM_NO_SPLITK") != nullptr;
    static const int forced_bn = getenv("NPM_GEMM_BLOCK_N_DYN") ? atoi(getenv("NPM_GEMM_BLOCK_N_DYN")) : 0;
    const bool c_planes = d.c_split != nullptr;
    if (c_planes)
        NPM_REQUIRE(nb1 == 1 && nb2 == 1 && !(d.flags & NPM_GEMM_ACCUM) && d.residual == nullptr && aligned16(d.c_split) && d.ldc % 8 == 0 &&
                    d.c_split_plane % 8 == 0, "gemm: c_split needs an unbatched, non-accumulating problem and 16-byte aligned bf16 rows");
    const bool may_split = c_dense && !(d.flags & (NPM_GEMM_RELU | NPM_GEMM_ACCUM)) && !no_splitk && !c_planes;
    int best_bn = 128, best_splits = 1;
    {
        // cycles per K stage of one pair: 2 * nterms MMAs of N x K16 at N/2 clk each (tensor floor)
        double best_cost = 1e300;
        const int cands[2] = {256, 128};
        for (int i = 0; i < 2; ++i) {
            const int bn = cands[i];
            if (forced_bn && bn != forced_bn) continue;
            const double kstep = (nterms == 3 ? 3.0 : 1.0) * bn + (bn == 128 ? 120.0 : 0.0);
            const int64_t tn = (d.n + bn - 1) / bn;
            const int64_t tiles = tiles_m * tn * nb1 * nb2;
            for (int sp = 1; sp <= (may_split ? 16 : 1); sp *= 2) {
                const int kbs = (num_kb_total + sp - 1) / sp;
                if (sp > 1 && kbs < 16) break;
                const int64_t waves = (tiles * sp + units - 1) / units;
                double cost = double(waves) * (kbs * kstep + 1500.0 + 8.0 * bn);
                if (sp > 1) cost = cost * 1.08 + 3000.0;
                if (cost < best_cost - 1e-9) { best_cost = cost; best_bn = bn; best_splits = sp; }
            }
        }
    }
    const int bn = best_bn;
    const int kb_per_split = (num_kb_total + best_splits - 1) / best_splits;
    const int splits = (num_kb_total + kb_per_split - 1) / kb_per_split;
    const int64_t tiles_n = (d.n + bn - 1) / bn;
    const int64_t items_per_split = tiles_m * tiles_n * nb1 * nb2;
    const int64_t total = items_per_split * splits;
    if (total > (1ll << 30)) { set_error("gemm: too many tiles"); return NPM_ERR_INVALID; }
    if (splits > 1) {
        cudaError_t e = cudaMemsetAsync(d.c, 0, sizeof(float) * (size_t)d.m * d.n * nb1 * nb2, stream);
        if (e != cudaSuccess) { set_error("gemm split-K memset: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
    }
    CUtensorMap tmA, tmB, tmC;
    auto bs = [](int nb, int64_t s, uint64_t natural) -> uint64_t { return nb > 1 ? (uint64_t)s : natural; };
    const uint64_t M = d.m, N = d.n, K = d.k;
    const bool a_chunked = a_mn && (d.m % 32 == 0), b_chunked = b_mn && (d.n % 32 == 0);
    int rc;
    // fp32 staging images: K-major = one box {32 k, rows}; MN-major = {32 mn, 32 k} boxes of 4 KB, all of a stage in one
    // TMA instruction when the operand has the 5-D {32, K, MN/32, nb1, nb2} view (a box costs ~46 clk + bytes / 70 B/clk)
    // pre-split A (an activation its producer wrote as bf16 planes): unbatched, three terms; MN-major needs whole 64-row slabs
    const bool a_pre = d.a_split != nullptr && nterms == 3 && nb1 == 1 && nb2 == 1 && (!a_mn || d.m % 64 == 0) && aligned16(d.a_split) &&
                       d.a_split_plane % 8 == 0 && (a_mn ? d.a_cs : d.a_rs) % 8 == 0 && d.a_colsum == nullptr;
    if (d.a == nullptr && !a_pre) {
        set_error("gemm: a_split cannot be used for this problem (alignment / shape / mode) and no fp32 A was given");
        return NPM_ERR_UNSUPPORTED;
    }
    if (a_pre && !a_mn) {
        const uint64_t dims[4] = {K, M, 2, 1}, st[3] = {(uint64_t)d.a_rs, (uint64_t)d.a_split_plane, (uint64_t)d.a_split_plane * 2};
        const uint32_t box[4] = {kKS, kBM, 2, 1};
        rc = make_tensor_map_bf16_nd(&tmA, d.a_split, 4, dims, st, box, 64);
    } else if (a_pre) {
        const uint64_t dims[4] = {64, K, M / 64, 2}, st[3] = {(uint64_t)d.a_cs, 64, (uint64_t)d.a_split_plane};
        const uint32_t box[4] = {64, kKS, kBM / 64, 2};
        rc = make_tensor_map_bf16_nd(&tmA, d.a_split, 4, dims, st, box, 128);
    } else if (!a_mn) {
        const uint64_t ld = d.a_rs, s2 = bs(nb1, d.a_bs1, ld * M), s3 = bs(nb2, d.a_bs2, s2 * nb1);
        rc = make_tensor_map_4d(&tmA, d.a, K, M, nb1, nb2, ld, s2, s3, kKS, kBM, false, false);
    } else {
        const uint64_t ld = d.a_cs, s2 = bs(nb1, d.a_bs1, ld * K), s3 = bs(nb2, d.a_bs2, s2 * nb1);
        if (a_chunked) {
            const uint64_t dims[5] = {32, K, M / 32, (uint64_t)nb1, (uint64_t)nb2}, st[4] = {ld, 32, s2, s3};
            const uint32_t box[5] = {32, kKS, kBM / 32, 1, 1};
            rc = make_tensor_map_nd(&tmA, d.a, 5, dims, st, box, false, false);
        } else {
            rc = make_tensor_map_4d(&tmA, d.a, M, K, nb1, nb2, ld, s2, s3, 32, kKS, false, false);
        }
    }
    if (rc) return rc;
    // pre-split B (weights): unbatched, A K-major, three terms; MN-major needs whole 64-column slabs
    const bool b_pre = d.b_split != nullptr && nterms == 3 && nb1 == 1 && nb2 == 1 && (!b_mn || d.n % 64 == 0) &&
                       aligned16(d.b_split) && d.b_split_plane % 8 == 0 && (b_mn ? d.b_rs : d.b_cs) % 8 == 0;   // TMA: 16-byte strides in bf16
    if (d.b == nullptr && !b_pre) {
        set_error("gemm: b_split cannot be used for this problem (alignment / shape / mode) and no fp32 B was given");
        return NPM_ERR_UNSUPPORTED;
    }
    if (b_pre && !b_mn) {
        const uint64_t dims[4] = {K, N, 2, 1}, st[3] = {(uint64_t)d.b_cs, (uint64_t)d.b_split_plane, (uint64_t)d.b_split_plane * 2};
        const uint32_t box[4] = {kKS, (uint32_t)bn / 2, 2, 1};
        rc = make_tensor_map_bf16_nd(&tmB, d.b_split, 4, dims, st, box, 64);
    } else if (b_pre) {
        const uint64_t dims[4] = {64, K, N / 64, 2}, st[3] = {(uint64_t)d.b_rs, 64, (uint64_t)d.b_split_plane};
        const uint32_t box[4] = {64, kKS, (uint32_t)bn / 2 / 64, 2};
        rc = make_tensor_map_bf16_nd(&tmB, d.b_split, 4, dims, st, box, 128);
    } else if (!b_mn) {
        const uint64_t ld = d.b_cs, s2 = bs(nb1, d.b_bs1, ld * N), s3 = bs(nb2, d.b_bs2, s2 * nb1);
        rc = make_tensor_map_4d(&tmB, d.b, K, N, nb1, nb2, ld, s2, s3, kKS, bn / 2, false, false);
    } else {
        const uint64_t ld = d.b_rs, s2 = bs(nb1, d.b_bs1, ld * K), s3 = bs(nb2, d.b_bs2, s2 * nb1);
        if (b_chunked) {
            const uint64_t dims[5] = {32, K, N / 32, (uint64_t)nb1, (uint64_t)nb2}, st[4] = {ld, 32, s2, s3};
            const uint32_t box[5] = {32, kKS, (uint32_t)bn / 2 / 32, 1, 1};
            rc = make_tensor_map_nd(&tmB, d.b, 5, dims, st, box, false, false);
        } else {
            rc = make_tensor_map_4d(&tmB, d.b, N, K, nb1, nb2, ld, s2, s3, 32, kKS, false, false);
        }
    }
    if (rc) return rc;
    {
        const uint64_t ld = d.ldc;
        const uint64_t s2 = bs(nb1, d.c_bs1, ld * d.m), s3 = bs(nb2, d.c_bs2, s2 * nb1);
        if (c_planes) {
            const uint64_t dims[4] = {N, M, 2, 1}, st[3] = {ld, (uint64_t)d.c_split_plane, (uint64_t)d.c_split_plane * 2};
            const uint32_t box[4] = {32, 32, 1, 1};
            rc = make_tensor_map_bf16_nd(&tmC, d.c_split, 4, dims, st, box, 0);
        } else {
            rc = make_tensor_map_4d(&tmC, d.c, d.n, d.m, nb1, nb2, ld, s2, s3, 32, 32, false, false);
        }
        if (rc) return rc;
    }
    GemmBxArgs args;
    args.a_chunked = a_chunked ? 1 : 0;
    args.b_chunked = b_chunked ? 1 : 0;
    args.M = (int)d.m; args.N = (int)d.n; args.K = (int)d.k;
    args.tiles_m = (int)tiles_m; args.tiles_n = (int)tiles_n; args.nb1 = nb1;
    args.total_tiles = (int)total;
    args.splits = splits; args.kb_per_split = kb_per_split; args.items_per_split = (int)items_per_split;
    static const int band_env = getenv("NPM_GEMM_BAND") ? atoi(getenv("NPM_GEMM_BAND")) : 8;
    args.band_m = (tiles_m > band_env && tiles_n > 1) ? band_env : 0;
    args.alpha = d.alpha;
    args.bias = d.bias;
    args.residual = d.residual;
    args.ldr = d.ldr;
    args.relu = (d.flags & NPM_GEMM_RELU) ? 1 : 0;
    args.accum = (d.flags & NPM_GEMM_ACCUM) ? 1 : 0;
    args.c_planes = c_planes ? 1 : 0;
    args.rowdot_x = nullptr; args.rowdot_ld = 0; args.rowdot_out = nullptr; args.rowdot_seq = 1;
    if (d.rowdot_x != nullptr) {
        NPM_REQUIRE(c_planes && d.rowdot_out != nullptr && d.n % 64 == 0 && d.rowdot_ld % 4 == 0 && d.rowdot_ld >= d.n && aligned16(d.rowdot_x) &&
                    d.rowdot_seq > 0 && d.rowdot_seq < (1ll << 31) && d.m % d.rowdot_seq == 0,
                    "gemm: rowdot needs the c_split output, n %% 64 == 0, a 16-byte aligned rowdot_x and m a multiple of rowdot_seq");
        args.rowdot_x = d.rowdot_x; args.rowdot_ld = d.rowdot_ld; args.rowdot_out = d.rowdot_out; args.rowdot_seq = (int)d.rowdot_seq;
    }
    args.a_colsum = nullptr;
    if (d.a_colsum != nullptr) {
        NPM_REQUIRE(a_mn && nb1 == 1 && nb2 == 1, "gemm: a_colsum needs an unbatched problem with an MN-major A");
        cudaError_t e = cudaMemsetAsync(d.a_colsum, 0, sizeof(float) * (size_t)d.m, stream);
        if (e != cudaSuccess) { set_error("gemm a_colsum memset: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
        args.a_colsum = d.a_colsum;
    }
    args.dbg = nullptr;
    const int grid = 2 * (int)(total < units ? total : units);
    if (nterms == 3) {
        if (bn == 256) return launch_bx_major<256, 3>(a_mn, b_mn, a_pre, b_pre, tmA, tmB, tmC, args, grid, stream);
        return launch_bx_major<128, 3>(a_mn, b_mn, a_pre, b_pre, tmA, tmB, tmC, args, grid, stream);
    }
    if (bn == 256) return launch_bx_major<256, 1>(a_mn, b_mn, a_pre, b_pre, tmA, tmB, tmC, args, grid, stream);
    return launch_bx_major<128, 1>(a_mn, b_mn, a_pre, b_pre, tmA, tmB, tmC, args, grid, stream);
}

}  // namespace npm
