// conv_tc.cu — Conv2D on the tensor cores: TMA-staged implicit GEMM, tcgen05 kind::tf32, accumulators in TMEM.
//
// layers/conv.py computes SAME / stride-1 convolution as k*k shifted-window GEMMs over zero-padded copies
// (:97-107), the input gradient as the same convolution with flipped, IO-transposed filters (:110-153) and
// the filter gradient as k*k GEMMs  xpad[:, i:, j:, :]^T @ dy  (:185-194).  Here no padded copy and no im2col
// matrix exist: the activation tensor is described to TMA as a 4-D tensor (C, W, H, N) and every K step
// loads the box of the current filter tap at SHIFTED coordinates — the out-of-bounds part of a box is
// zero-filled by TMA, which is exactly the SAME padding.
//
//   fprop : y [pix, o] = sum_{tap, c} x [pix + tap, c] * f[tap, c, o]        A = x boxes  (K-major)
//                                                                            B = f[tap]   (MN-major, o contiguous)
//   dgrad : dx[pix, c] = sum_{tap, o} dy[pix + tap, o] * f[flip(tap), c, o]  A = dy boxes (K-major)
//                                                                            B = f[flip]  (K-major, o contiguous)
//   wgrad : dw[tap][c, o] = sum_pix x[pix + tap, c] * dy[pix, o]             A = x boxes  (MN-major), B = dy boxes
//                                                                            (MN-major); split over pixel ranges,
//                                                                            TMA reduce-add into a zeroed dw
// A 128-pixel M tile is a TW x TH rectangle of one image (TW = 32, TH = 4 for the 32 x 32 images of
// BASELINE cfg2), so its box lands in shared memory as 128 rows of 32 channels = the K-major operand
// layout of csrc/gemm_tc.cu; warp roles, ring, TMEM double buffering and epilogue are that kernel's.
// Served in the TF32 and split-bf16 modes for Cin % 4 == 0, Cout % 4 == 0, W >= 8 (TMA needs 16-byte pixel strides —
// the 3-channel input layer of cfg2 and the 3xTF32 / fp32 modes run the exact fp32 kernel in conv.cu).
//
// BX = true ("bf16x3" mode, rtol 1e-3 / atol 1e-4): the same boxes land as plain fp32 (128-byte swizzle), eight
// converter warps rewrite every stage in place as bf16 hi / mid images (bx_convert.cuh, as in gemm_bx.cu) and the
// issuer runs mid*hi + hi*mid + hi*hi with kind::f16.
#include "bx_convert.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace npm {

int make_tensor_map_4d_box(CUtensorMap* tm, const float* base, const uint64_t dims[4], const uint64_t strides[3],
                           const uint32_t box[4], bool round_tf32, bool atom32b);

namespace {

constexpr int kBlockM = 128, kBlockK = 32, kUmmaK = 8;
enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };

struct ConvTcArgs {
    int H, W, NB;              // image height / width / batch
    int ks, pad, taps;
    int kc_blocks;             // 32-channel blocks of the contraction (fprop / dgrad)
    int Nn;                    // GEMM N: Cout (fprop, wgrad) or Cin (dgrad)
    int Mw;                    // wgrad: GEMM M = Cin
    int TW, TH;                // pixel rectangle of an M tile (TW * TH = 128)
    int PW, PH;                // 32-pixel rectangle: epilogue store box (fprop / dgrad), K box (wgrad)
    int tiles_w, tiles_h, tiles_n, tiles_m;
    int rect_w, rect_h;        // wgrad: 32-pixel rectangles per image row / column
    int kb_total, kb_per_split, splits;   // wgrad K blocks (= rectangles over the whole batch)
    int tap_pack;              // wgrad with Cin <= 64: 128 / Cin filter taps share one M tile (rows = [tap][channel]), else 1
    int total_tiles;
    const float* bias;
    int relu;
    uint64_t desc_a, desc_b;
};

template <int BLOCK_N>
struct CCfg {
    static constexpr int kABytes = kBlockM * kBlockK * 4;
    static constexpr int kBBytes = BLOCK_N * kBlockK * 4;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBytes = 4 * 2 * 4096;
    static constexpr int kBarBytes = 512;
    static constexpr int kBudget = 232448 - 1024;
    static constexpr int kStagesRaw = (kBudget - kEpiBytes - kBarBytes) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;
    static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
};

struct Tile {      // decoded work item
    int n_img, h0, w0, n0;     // fprop / dgrad
    int tap, m0, kb0, kb1;     // wgrad
};

template <int MODE>
__device__ __forceinline__ Tile decode(const ConvTcArgs& a, int tile) {
    Tile t{};
    if (MODE != WGRAD) {
        int r = tile;
        const int tw = r % a.tiles_w; r /= a.tiles_w;
        const int th = r % a.tiles_h; r /= a.tiles_h;
        const int nt = r % a.tiles_n; r /= a.tiles_n;
        t.n_img = r; t.h0 = th * a.TH; t.w0 = tw * a.TW; t.n0 = nt;
    } else {
        int r = tile;
        const int sp = r % a.splits; r /= a.splits;
        const int nt = r % a.tiles_n; r /= a.tiles_n;
        const int mt = r % a.tiles_m; r /= a.tiles_m;
        t.tap = r; t.m0 = mt * kBlockM; t.n0 = nt;
        t.kb0 = sp * a.kb_per_split;
        t.kb1 = min(a.kb_total, t.kb0 + a.kb_per_split);
    }
    return t;
}

constexpr int kFirstConvWarp = 6;           // BX: warps 6-13 convert

template <int MODE, int BLOCK_N, bool BX>
__global__ void __launch_bounds__(BX ? 32 * (kFirstConvWarp + bx::kConvWarps) : 192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const ConvTcArgs args) {
    using Cfg = CCfg<BLOCK_N>;
    constexpr int S = Cfg::kStages;
    constexpr bool A_MN = MODE == WGRAD;
    constexpr bool B_MN = MODE != DGRAD;

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
    auto conv_bar   = [&](int s) { return bar_addr + 8u * (2 * S + 4 + s); };   // BX: the converters have rewritten stage s
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(
        base_ptr + S * Cfg::kStageBytes + Cfg::kEpiBytes + 8 * (3 * S + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 4 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        ptx::prefetch_tensormap(&tmC);
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < S; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); ptx::mbar_init(conv_bar(s), bx::kConvWarps); }
            for (int a = 0; a < 2; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), 4); }
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

    const int num_kb_fd = args.taps * args.kc_blocks;      // K blocks of one fprop / dgrad tile

    if (warp == 4) {
        // ============================ TMA producer ============================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
                const Tile t = decode<MODE>(args, tile);
                const int n0 = t.n0 * BLOCK_N;
                const int kb0 = MODE == WGRAD ? t.kb0 : 0;
                const int kb1 = MODE == WGRAD ? t.kb1 : num_kb_fd;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                    const uint32_t sB = sA + Cfg::kABytes;
                    const uint32_t fb = full_bar(stage);
                    ptx::mbar_arrive_expect_tx(fb, Cfg::kStageBytes);
                    if (MODE != WGRAD) {
                        const int tap = kb / args.kc_blocks, c0 = (kb - tap * args.kc_blocks) * kBlockK;
                        const int ti = tap / args.ks, tj = tap - ti * args.ks;
                        // the tap's window of the (virtually zero-padded) activation: a TW x TH box at shifted coordinates
                        ptx::tma_load_4d(sA, &tmA, fb, c0, t.w0 + tj - args.pad, t.h0 + ti - args.pad, t.n_img);
                        if (MODE == FPROP) {
#pragma unroll
                            for (int c = 0; c < BLOCK_N / 32; ++c)      // f[ti, tj, c0.., n0..]: o contiguous → MN-major boxes
                                ptx::tma_load_4d(sB + c * 4096, &tmB, fb, n0 + c * 32, c0, tj, ti);
                        } else {                                        // f[ks-1-ti, ks-1-tj, n0.., c0..]: o (= k) contiguous
                            ptx::tma_load_4d(sB, &tmB, fb, c0, n0, args.ks - 1 - tj, args.ks - 1 - ti);
                        }
                    } else {
                        // K block = one PW x PH rectangle of 32 pixels of image n
                        int r = kb;
                        const int rw = r % args.rect_w; r /= args.rect_w;
                        const int rh = r % args.rect_h; r /= args.rect_h;
                        const int w = rw * args.PW, h = rh * args.PH;
#pragma unroll
                        for (int c = 0; c < kBlockM / 32; ++c) {
                            // 32 rows of the M tile: channels c*32.. of the tile's tap, or (packed) of tap group*pack + (c*32)/Cin —
                            // a tap past the last one loads a valid window whose rows the epilogue drops
                            int tap = t.tap, ch = t.m0 + c * 32;
                            if (args.tap_pack > 1) {
                                tap = min(t.tap * args.tap_pack + (c * 32) / args.Mw, args.taps - 1);
                                ch = (c * 32) % args.Mw;
                            }
                            const int ti = tap / args.ks, tj = tap - ti * args.ks;
                            ptx::tma_load_4d(sA + c * 4096, &tmA, fb, ch, w + tj - args.pad, h + ti - args.pad, r);
                        }
#pragma unroll
                        for (int c = 0; c < BLOCK_N / 32; ++c)
                            ptx::tma_load_4d(sB + c * 4096, &tmB, fb, n0 + c * 32, w, h, r);
                    }
                    if (++stage == S) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 5) {
        // ============================= MMA issuer =============================
        constexpr uint32_t idesc = BX ? ptx::umma_idesc_bf16(kBlockM, BLOCK_N, A_MN, B_MN) : ptx::umma_idesc_tf32(kBlockM, BLOCK_N, A_MN, B_MN);
        constexpr uint32_t a_kstep = A_MN ? 1024u : 32u;
        constexpr uint32_t b_kstep = B_MN ? 1024u : 32u;
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
            int nkb = num_kb_fd;
            if (MODE == WGRAD) { const Tile t = decode<MODE>(args, tile); nkb = t.kb1 - t.kb0; }
            ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < nkb; ++kb) {
                ptx::mbar_wait(BX ? conv_bar(stage) : full_bar(stage), phase);
                if (BX) ptx::fence_proxy_async_smem();       // the converters' generic-proxy stores -> async proxy (see gemm_bx.cu)
                ptx::tc_fence_after();
                __syncwarp();
                if (ptx::elect_one()) {
                    const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                    const uint32_t sB = sA + Cfg::kABytes;
                    if (BX) {
                        constexpr uint64_t dA = A_MN ? bx::kDescMN : bx::kDescK, dB = B_MN ? bx::kDescMN : bx::kDescK;
#pragma unroll
                        for (int t = 0; t < 3; ++t)            // mid*hi, hi*mid, hi*hi
#pragma unroll
                            for (int sl = 0; sl < 2; ++sl)
                                ptx::umma_f16(d_tmem, ptx::umma_desc(dA, sA + bx::slice_off<kBlockM, A_MN>(t == 0 ? 1 : 0, sl)),
                                              ptx::umma_desc(dB, sB + bx::slice_off<BLOCK_N, B_MN>(t == 1 ? 1 : 0, sl)), idesc,
                                              (kb | t | sl) != 0 ? 1u : 0u);
                    } else
#pragma unroll
                    for (int kk = 0; kk < kBlockK / kUmmaK; ++kk)
                        ptx::umma_tf32(d_tmem, ptx::umma_desc(args.desc_a, sA + kk * a_kstep),
                                       ptx::umma_desc(args.desc_b, sB + kk * b_kstep), idesc, (kb | kk) != 0 ? 1u : 0u);
                    ptx::umma_commit(empty_bar(stage));
                    if (kb == nkb - 1) ptx::umma_commit(tfull_bar(acc));
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
            const Tile t = decode<MODE>(args, tile);
            const int n0 = t.n0 * BLOCK_N;
            // where this warp's 32 accumulator rows go
            int c1, c2, c3;
            bool rows_live;
            if (MODE != WGRAD) {
                const int p0 = warp * 32;                 // first pixel of the warp within the TW x TH tile
                c1 = t.w0 + p0 % args.TW; c2 = t.h0 + p0 / args.TW; c3 = t.n_img;
                rows_live = c1 < args.W && c2 < args.H;
            } else {
                int tap = t.tap;
                c1 = t.m0 + warp * 32;
                if (args.tap_pack > 1) {
                    tap = t.tap * args.tap_pack + (warp * 32) / args.Mw;
                    c1 = (warp * 32) % args.Mw;
                }
                const int ti = tap / args.ks, tj = tap - ti * args.ks;
                c2 = tj; c3 = ti;
                rows_live = c1 < args.Mw && tap < args.taps;
            }
            ptx::mbar_wait(tfull_bar(acc), acc_phase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
                const int nc = n0 + chunk * 32;
                if (nc >= args.Nn || !rows_live) break;
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + (uint32_t(warp * 32) << 16) + acc * BLOCK_N + chunk * 32, v);
                ptx::tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (MODE == FPROP && args.bias != nullptr) {
                    const float* bp = args.bias + nc;       // 128-bit broadcast loads (see gemm_tc.cu)
                    if (nc + 32 <= args.Nn && ((reinterpret_cast<uintptr_t>(bp) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp) + j);
                            f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nc + j < args.Nn) f[j] += __ldg(bp + j);
                    }
                }
                if (MODE == FPROP && args.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = f[j] < 0.0f ? -0.0f : __uint_as_float(__float_as_uint(f[j]) & 0x7fffffffu);
                }
                const uint32_t buf = nstore & 1u;
                if (lane == 0) ptx::tma_wait_group_read<1>();
                __syncwarp();
                uint8_t* row = stg_base + buf * 4096 + lane * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(row + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (ptx::elect_one()) {
                    if (MODE == WGRAD) ptx::tma_reduce_add_4d(&tmC, stg_addr + buf * 4096, nc, c1, c2, c3);
                    else               ptx::tma_store_4d(&tmC, stg_addr + buf * 4096, nc, c1, c2, c3);
                    ptx::tma_commit_group();
                }
                ++nstore;
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
        }
        if (lane == 0) ptx::tma_wait_group<0>();
    } else if (BX && warp >= kFirstConvWarp) {
        // ===================== converters: fp32 stage -> bf16 hi / mid images in place =====================
        const int pw = warp - kFirstConvWarp;
        bx::OperandConverter<kBlockM, A_MN, 3> ca;
        bx::OperandConverter<BLOCK_N, B_MN, 3> cb;
        ca.init(pw, lane);
        cb.init(pw, lane);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < args.total_tiles; tile += gridDim.x) {
            int nkb = num_kb_fd;
            if (MODE == WGRAD) { const Tile t = decode<MODE>(args, tile); nkb = t.kb1 - t.kb0; }
            for (int kb = 0; kb < nkb; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                const uint32_t sA = stage_addr + stage * Cfg::kStageBytes;
                float4 va[bx::OperandConverter<kBlockM, A_MN, 3>::NLD];
                float4 vb[bx::OperandConverter<BLOCK_N, B_MN, 3>::NLD];
                ca.load(va, sA);
                cb.load(vb, sA + Cfg::kABytes);
                bx::bar_sync_conv();
                ca.store(va, sA);
                cb.store(vb, sA + Cfg::kABytes);
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(conv_bar(stage));
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

int pow2_floor(int64_t v, int cap) {
    int p = 1;
    while (p * 2 <= v && p * 2 <= cap) p *= 2;
    return p;
}

template <int MODE, int BN, bool BX>
int launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const ConvTcArgs& args, cudaStream_t s) {
    using Cfg = CCfg<BN>;
    auto kern = conv_tc_kernel<MODE, BN, BX>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) { set_error("conv_tc smem attribute: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
        configured = true;
    }
    const int grid = args.total_tiles < num_sms() ? args.total_tiles : num_sms();
    kern<<<grid, BX ? 32 * (kFirstConvWarp + bx::kConvWarps) : 192, Cfg::kSmemBytes, s>>>(a, b, c, args);
    count_launch();
    return check_launch("conv_tc_kernel");
}
template <int MODE>
int launch_bn(int bn, bool bx_mode, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const ConvTcArgs& args, cudaStream_t s) {
    if (bx_mode) {
        if (bn == 256) return launch<MODE, 256, true>(a, b, c, args, s);
        if (bn == 128) return launch<MODE, 128, true>(a, b, c, args, s);
        return launch<MODE, 64, true>(a, b, c, args, s);
    }
    if (bn == 256) return launch<MODE, 256, false>(a, b, c, args, s);
    if (bn == 128) return launch<MODE, 128, false>(a, b, c, args, s);
    return launch<MODE, 64, false>(a, b, c, args, s);
}
int pick_bn(int64_t n) { return n > 128 ? 256 : (n > 64 ? 128 : 64); }

}  // namespace

bool conv_tc_supported(const void* p0, const void* p1, const void* p2, int64_t N, int64_t H, int64_t W, int64_t Cin,
                       int64_t Cout, int ks) {
    return aligned16(p0) && aligned16(p1) && aligned16(p2) && (Cin & 3) == 0 && (Cout & 3) == 0 && W >= 8 && H >= 1 &&
           N >= 1 && N < (1 << 20) && H < (1 << 15) && W < (1 << 15) && ks >= 1 && ks <= 15 &&
           N * H * W < (1ll << 31);
}

// fprop (act = x, Ca = Cin, Cn = Cout) and dgrad (act = dy, Ca = Cout, Cn = Cin)
int conv_tc_fprop_dgrad(bool dgrad, const float* act, const float* f, const float* bias, float* out, int64_t N, int64_t H,
                        int64_t W, int64_t Cin, int64_t Cout, int ks, int relu, bool bx_mode, cudaStream_t stream) {
    const bool rnd = !bx_mode;          // TF32 mode: TMA rounds to tf32; split-bf16 mode: plain fp32, 16-byte swizzle atoms
    const int64_t Ca = dgrad ? Cout : Cin, Cn = dgrad ? Cin : Cout;
    ConvTcArgs a{};
    a.H = (int)H; a.W = (int)W; a.NB = (int)N; a.ks = ks; a.pad = ks / 2; a.taps = ks * ks;
    a.kc_blocks = (int)((Ca + kBlockK - 1) / kBlockK);
    a.Nn = (int)Cn;
    a.TW = pow2_floor(W, 128); a.TH = kBlockM / a.TW;
    a.PW = a.TW >= 32 ? 32 : a.TW; a.PH = 32 / a.PW;
    a.tiles_w = (int)((W + a.TW - 1) / a.TW); a.tiles_h = (int)((H + a.TH - 1) / a.TH);
    const int bn = pick_bn(Cn);
    a.tiles_n = (int)((Cn + bn - 1) / bn);
    const int64_t total = (int64_t)N * a.tiles_h * a.tiles_w * a.tiles_n;
    NPM_REQUIRE(total < (1ll << 30), "conv2d: too many tiles");
    a.total_tiles = (int)total;
    a.bias = bias; a.relu = relu;
    a.desc_a = ptx::umma_desc_base(2, 16, 1024);
    a.desc_b = dgrad ? ptx::umma_desc_base(2, 16, 1024) : ptx::umma_desc_base(1, 4096, 512);

    CUtensorMap tA, tB, tC;
    int rc;
    {
        const uint64_t dims[4] = {(uint64_t)Ca, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        const uint64_t str[3] = {(uint64_t)Ca, (uint64_t)Ca * W, (uint64_t)Ca * W * H};
        const uint32_t box[4] = {32, (uint32_t)a.TW, (uint32_t)a.TH, 1};
        if ((rc = make_tensor_map_4d_box(&tA, act, dims, str, box, rnd, false))) return rc;
    }
    {   // filters f[kh, kw, Cin, Cout]
        const uint64_t str[3] = {(uint64_t)Cout, (uint64_t)Cout * Cin, (uint64_t)Cout * Cin * ks};
        if (!dgrad) {
            const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Cin, (uint64_t)ks, (uint64_t)ks};
            const uint32_t box[4] = {32, 32, 1, 1};
            if ((rc = make_tensor_map_4d_box(&tB, f, dims, str, box, rnd, !bx_mode))) return rc;
        } else {
            const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Cin, (uint64_t)ks, (uint64_t)ks};
            const uint32_t box[4] = {32, (uint32_t)bn, 1, 1};
            if ((rc = make_tensor_map_4d_box(&tB, f, dims, str, box, rnd, false))) return rc;
        }
    }
    {
        const uint64_t dims[4] = {(uint64_t)Cn, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        const uint64_t str[3] = {(uint64_t)Cn, (uint64_t)Cn * W, (uint64_t)Cn * W * H};
        const uint32_t box[4] = {32, (uint32_t)a.PW, (uint32_t)a.PH, 1};
        if ((rc = make_tensor_map_4d_box(&tC, out, dims, str, box, false, false))) return rc;
    }
    return dgrad ? launch_bn<DGRAD>(bn, bx_mode, tA, tB, tC, a, stream) : launch_bn<FPROP>(bn, bx_mode, tA, tB, tC, a, stream);
}

// dw[kh, kw, Cin, Cout] (must be zero on entry: partial sums are TMA-reduced into it)
int conv_tc_wgrad(const float* x, const float* dy, float* dw, int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                  int ks, bool bx_mode, cudaStream_t stream) {
    const bool rnd = !bx_mode;
    ConvTcArgs a{};
    a.H = (int)H; a.W = (int)W; a.NB = (int)N; a.ks = ks; a.pad = ks / 2; a.taps = ks * ks;
    a.Nn = (int)Cout; a.Mw = (int)Cin;
    a.PW = pow2_floor(W, 32); a.PH = 32 / a.PW;
    a.rect_w = (int)((W + a.PW - 1) / a.PW); a.rect_h = (int)((H + a.PH - 1) / a.PH);
    const int64_t kb_total = (int64_t)N * a.rect_w * a.rect_h;
    NPM_REQUIRE(kb_total < (1ll << 30), "conv2d wgrad: too many K blocks");
    a.kb_total = (int)kb_total;
    const int bn = pick_bn(Cout);
    a.tiles_n = (int)((Cout + bn - 1) / bn);
    a.tiles_m = (int)((Cin + kBlockM - 1) / kBlockM);
    // Cin = 32 / 64: the 128-row M tile would be 1/4 / 1/2 empty — 4 / 2 taps share it (their A boxes are the same
    // channels at differently shifted pixel coordinates), 9 taps in 3 / 5 tiles instead of 9
    static const bool no_pack = getenv("NPM_CONV_NO_TAP_PACK") != nullptr;
    a.tap_pack = (!no_pack && (Cin == 32 || Cin == 64)) ? (int)(kBlockM / Cin) : 1;
    const int tap_groups = (a.taps + a.tap_pack - 1) / a.tap_pack;
    const int64_t base_items = (int64_t)tap_groups * a.tiles_m * a.tiles_n;
    static const int waves_env = getenv("NPM_CONV_WGRAD_WAVES") ? atoi(getenv("NPM_CONV_WGRAD_WAVES")) : 2;
    int64_t splits = ((int64_t)num_sms() * waves_env) / base_items;              // at most `waves` whole waves of work items
    const int64_t max_splits = (kb_total + 15) / 16;                            // >= 16 K blocks per item
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.kb_per_split = (int)((kb_total + splits - 1) / splits);
    a.splits = (int)((kb_total + a.kb_per_split - 1) / a.kb_per_split);
    a.total_tiles = (int)(base_items * a.splits);
    a.desc_a = ptx::umma_desc_base(1, 4096, 512);
    a.desc_b = ptx::umma_desc_base(1, 4096, 512);

    CUtensorMap tA, tB, tC;
    int rc;
    {
        const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        const uint64_t str[3] = {(uint64_t)Cin, (uint64_t)Cin * W, (uint64_t)Cin * W * H};
        const uint32_t box[4] = {32, (uint32_t)a.PW, (uint32_t)a.PH, 1};
        if ((rc = make_tensor_map_4d_box(&tA, x, dims, str, box, rnd, !bx_mode))) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        const uint64_t str[3] = {(uint64_t)Cout, (uint64_t)Cout * W, (uint64_t)Cout * W * H};
        const uint32_t box[4] = {32, (uint32_t)a.PW, (uint32_t)a.PH, 1};
        if ((rc = make_tensor_map_4d_box(&tB, dy, dims, str, box, rnd, !bx_mode))) return rc;
    }
    {
        const uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)Cin, (uint64_t)ks, (uint64_t)ks};
        const uint64_t str[3] = {(uint64_t)Cout, (uint64_t)Cout * Cin, (uint64_t)Cout * Cin * ks};
        const uint32_t box[4] = {32, 32, 1, 1};
        if ((rc = make_tensor_map_4d_box(&tC, dw, dims, str, box, false, false))) return rc;
    }
    return launch_bn<WGRAD>(bn, bx_mode, tA, tB, tC, a, stream);
}

}  // namespace npm
