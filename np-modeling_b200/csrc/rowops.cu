// rowops.cu — HBM-bound row kernels: Softmax fwd/bwd, LayerNormalization fwd/bwd, column sums.
//
// One warp owns one row; a row of up to 4096 fp32 lives entirely in registers (128-bit loads,
// lane-interleaved so every warp request is one fully coalesced 512-byte line group), reductions
// are warp shuffles.  Rows wider than that (or with a width that is not a multiple of 4) take a
// multi-pass variant of the same kernels.  Algorithmic bytes per element: softmax fwd 8, bwd 12;
// layernorm fwd 8 (+8 B/row statistics), bwd 12 (+8 B/row, + 2*C*4 B of parameter gradients).
#include <atomic>

#include "common.cuh"

namespace npm {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kRowThreads = kWarpsPerCta * 32;

// ---- a row held in registers: NV float4 per lane, element index (v*32 + lane)*4 + {0..3}
template <int NV>
struct RowRegs {
    float4 v[NV];
    __device__ __forceinline__ void load(const float* row, int cols, int lane, float pad) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            v[i] = (c < cols) ? ld_stream(reinterpret_cast<const float4*>(row + c)) : make_float4(pad, pad, pad, pad);
        }
    }
    __device__ __forceinline__ void store(float* row, int cols, int lane) const {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < cols) st_stream(reinterpret_cast<float4*>(row + c), v[i]);
        }
    }
};

// =============================================================== softmax
template <int NV>
__global__ void __launch_bounds__(kRowThreads) softmax_fwd_kernel(const float* x, float* y, int64_t rows, int cols) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    RowRegs<NV> r;
    r.load(x + row * cols, cols, lane, -INFINITY);
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NV; ++i) m = fmaxf(m, fmaxf(fmaxf(r.v[i].x, r.v[i].y), fmaxf(r.v[i].z, r.v[i].w)));
    m = warp_max(m);
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        r.v[i].x = expf(r.v[i].x - m); r.v[i].y = expf(r.v[i].y - m);
        r.v[i].z = expf(r.v[i].z - m); r.v[i].w = expf(r.v[i].w - m);
        s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
    }
    s = warp_sum(s);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        r.v[i].x = __fdiv_rn(r.v[i].x, s); r.v[i].y = __fdiv_rn(r.v[i].y, s);
        r.v[i].z = __fdiv_rn(r.v[i].z, s); r.v[i].w = __fdiv_rn(r.v[i].w, s);
    }
    r.store(y + row * cols, cols, lane);
}

template <int NV>
__global__ void __launch_bounds__(kRowThreads) softmax_bwd_kernel(const float* y, const float* dy, float* dx,
                                                                  int64_t rows, int cols, float scale) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    RowRegs<NV> ry, rd;
    ry.load(y + row * cols, cols, lane, 0.0f);
    rd.load(dy + row * cols, cols, lane, 0.0f);
    float dot = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        dot += (ry.v[i].x * rd.v[i].x + ry.v[i].y * rd.v[i].y) + (ry.v[i].z * rd.v[i].z + ry.v[i].w * rd.v[i].w);
    dot = warp_sum(dot);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        rd.v[i].x = scale * ry.v[i].x * (rd.v[i].x - dot); rd.v[i].y = scale * ry.v[i].y * (rd.v[i].y - dot);
        rd.v[i].z = scale * ry.v[i].z * (rd.v[i].z - dot); rd.v[i].w = scale * ry.v[i].w * (rd.v[i].w - dot);
    }
    rd.store(dx + row * cols, cols, lane);
}

// generic (any width / alignment): multi-pass, lanes stride the row
__global__ void __launch_bounds__(kRowThreads) softmax_fwd_generic(const float* x, float* y, int64_t rows, int64_t cols) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * cols;
    float* yr = y + row * cols;
    float m = -INFINITY;
    for (int64_t c = lane; c < cols; c += 32) m = fmaxf(m, xr[c]);
    m = warp_max(m);
    float s = 0.0f;
    for (int64_t c = lane; c < cols; c += 32) s += expf(xr[c] - m);
    s = warp_sum(s);
    for (int64_t c = lane; c < cols; c += 32) yr[c] = __fdiv_rn(expf(xr[c] - m), s);
}
__global__ void __launch_bounds__(kRowThreads) softmax_bwd_generic(const float* y, const float* dy, float* dx,
                                                                   int64_t rows, int64_t cols, float scale) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* yr = y + row * cols;
    const float* dr = dy + row * cols;
    float* xr = dx + row * cols;
    float dot = 0.0f;
    for (int64_t c = lane; c < cols; c += 32) dot += yr[c] * dr[c];
    dot = warp_sum(dot);
    for (int64_t c = lane; c < cols; c += 32) xr[c] = scale * yr[c] * (dr[c] - dot);
}

inline int nv_for(int64_t cols) {   // float4 per lane needed to hold a row, rounded to a template size
    const int64_t need = (cols + 127) / 128;
    if (need <= 1) return 1;
    if (need <= 2) return 2;
    if (need <= 4) return 4;
    if (need <= 8) return 8;
    if (need <= 16) return 16;
    if (need <= 32) return 32;
    return 0;
}

// ============================================================= layernorm
// Optional DropOut fused in front of the normalisation (the pre-norm blocks of layers/transformer.py run
// DropOut -> LayerNormalization back to back, :35-37): the dropped tensor is never written, forward and
// backward recompute the Philox mask of elementwise.cu's dropout_kernel from (seed, offset + element index).
struct DropArgs {
    float keep_prob;      // 0 = no dropout
    float inv_keep;       // 1 / keep_prob rounded to fp32: the fused kernels scale by it (<= 1 ulp from x / keep_prob;
                          // the stand-alone DropOut kernel keeps the exact division of normalizations.py:22)
    uint64_t thr, seed, offset;   // offset is a multiple of 4: one Philox call per float4
};
// applies the mask to one float4 whose first element has global index e; returns the 4 keep bits
__device__ __forceinline__ uint32_t drop4(float4& v, const DropArgs& d, uint64_t e) {
    const Philox4 p = philox4x32_10((d.offset + e) >> 2, d.seed);
    const uint32_t k = (p.x < d.thr ? 1u : 0u) | (p.y < d.thr ? 2u : 0u) | (p.z < d.thr ? 4u : 0u) | (p.w < d.thr ? 8u : 0u);
    v.x = (k & 1u) ? v.x * d.inv_keep : 0.0f; v.y = (k & 2u) ? v.y * d.inv_keep : 0.0f;
    v.z = (k & 4u) ? v.z * d.inv_keep : 0.0f; v.w = (k & 8u) ? v.w * d.inv_keep : 0.0f;
    return k;
}
// PLANES: the result goes out ONLY as split-bf16 planes (`out` is the bf16 hi plane, the mid plane plane4 uint2 units
// further) — the A operand image of the GEMM that follows (the first FFN layer in a pre-norm block), same bytes as fp32.
__device__ __forceinline__ void ln_split2(float p0, float p1, uint32_t& hi, uint32_t& mid) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(p1), "f"(p0));
    const float r0 = p0 - __uint_as_float(hi << 16), r1 = p1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(r1), "f"(r0));
}
template <int NV, bool DROP = false, bool PLANES = false>
__global__ void __launch_bounds__(kRowThreads) layernorm_fwd_kernel(const float* __restrict__ x,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, float* __restrict__ out,
                                                                    float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                    int64_t rows, int cols, float eps, DropArgs drop = DropArgs{},
                                                                    uint32_t* __restrict__ maskbits = nullptr, int64_t plane4 = 0) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    RowRegs<NV> r;
    r.load(x + row * cols, cols, lane, 0.0f);
    if (DROP) {
        uint32_t keep_bits = 0;      // 4 bits per float4 of this lane; saved so that backward need not redo Philox
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < cols) keep_bits |= drop4(r.v[i], drop, (uint64_t)row * cols + c) << (4 * i);
        }
        maskbits[row * 32 + lane] = keep_bits;
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
    const float mean = warp_sum(s) / (float)cols;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < cols) {
            const float a = r.v[i].x - mean, b = r.v[i].y - mean, cc = r.v[i].z - mean, d = r.v[i].w - mean;
            q += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    const float var = warp_sum(q) / (float)cols;
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < cols) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
            r.v[i].x = g.x * ((r.v[i].x - mean) * rstd) + b.x;
            r.v[i].y = g.y * ((r.v[i].y - mean) * rstd) + b.y;
            r.v[i].z = g.z * ((r.v[i].z - mean) * rstd) + b.z;
            r.v[i].w = g.w * ((r.v[i].w - mean) * rstd) + b.w;
        }
    }
    if (PLANES) {
        uint2* pl = reinterpret_cast<uint2*>(out);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < cols) {
                uint2 h, m;
                ln_split2(r.v[i].x, r.v[i].y, h.x, m.x);
                ln_split2(r.v[i].z, r.v[i].w, h.y, m.y);
                const int64_t q = (row * cols + c) >> 2;
                pl[q] = h;
                pl[plane4 + q] = m;
            }
        }
    } else {
        r.store(out + row * cols, cols, lane);
    }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

__global__ void __launch_bounds__(kRowThreads) layernorm_fwd_generic(const float* x, const float* gamma, const float* beta,
                                                                     float* out, float* mean_out, float* rstd_out,
                                                                     int64_t rows, int64_t cols, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * cols;
    float s = 0.0f;
    for (int64_t c = lane; c < cols; c += 32) s += xr[c];
    const float mean = warp_sum(s) / (float)cols;
    float q = 0.0f;
    for (int64_t c = lane; c < cols; c += 32) { const float d = xr[c] - mean; q += d * d; }
    const float var = warp_sum(q) / (float)cols;
    const float rstd = 1.0f / sqrtf(var + eps);
    for (int64_t c = lane; c < cols; c += 32) out[row * cols + c] = gamma[c] * ((xr[c] - mean) * rstd) + beta[c];
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// Measured and rejected (round 2): a persistent forward (one wave of CTAs, the next row's loads issued before the
// current row is reduced) and a backward with one CTA per SM holding the next row's x / dz / statistics in flight.  Under
// ncu (one launch, cold L2) the fused dropout+LayerNorm backward went 39.1 -> 32.9 us and the plain backward 33.9 -> 31.3 us,
// the forward was unchanged and the Philox variant slower (21.6 -> 28.3 us: issue-bound, wants every warp the SM holds);
// inside the cfg5 step — inputs warm in L2, neighbours overlapped by PDL — the same library was 0.4 ms per step SLOWER
// (three alternating same-box runs: 87.23 vs 86.83 ms).  The occupancy-heavy forms below stay.
// Backward.  Each warp walks rows (grid-stride), writes dx, and accumulates its columns of
// dgamma/dbeta in registers; warps of a CTA are combined through shared memory and each CTA
// writes one partial row pair to the workspace [grid][2][cols]; colsum-style second stage
// finishes the reduction (no atomics → deterministic).
// XSUM: also accumulate the column sums of the dx this kernel writes (third partial row): dx is the gradient of the
// residual stream, whose column sum is the bias gradient of the projection / FFN layer that produced that stream
// (attentions.py:129, mlp.py:34) — summed here from registers instead of re-reading dx in a colsum launch.
template <int NV, bool DROP = false, bool XSUM = false>
__global__ void __launch_bounds__(kRowThreads, 2) layernorm_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ x,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    float* __restrict__ dx, float* __restrict__ partial,
                                                                    int64_t rows, int cols, DropArgs drop = DropArgs{},
                                                                    const float* __restrict__ dskip = nullptr,
                                                                    const uint32_t* __restrict__ maskbits = nullptr) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sred[];   // [kWarpsPerCta][NP][cols_padded], NP = 2 (+1 with XSUM)
    constexpr int NP = XSUM ? 3 : 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // dgamma / dbeta of this warp's rows accumulate in the warp's own shared-memory slice and gamma is
    // re-read through L1 for every row: with only the two row images in registers two CTAs fit an SM,
    // and 16 warps of loads in flight are what a latency-bound row kernel needs to approach HBM speed.
    constexpr int cpad = NV * 128;
    float* my = sred + (size_t)warp * NP * cpad;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        *reinterpret_cast<float4*>(my + c) = make_float4(0, 0, 0, 0);
        *reinterpret_cast<float4*>(my + cpad + c) = make_float4(0, 0, 0, 0);
        if (XSUM) *reinterpret_cast<float4*>(my + 2 * cpad + c) = make_float4(0, 0, 0, 0);
    }
    const float inv_c = 1.0f / (float)cols;
    for (int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + warp; row < rows; row += (int64_t)gridDim.x * kWarpsPerCta) {
        RowRegs<NV> rx, rz;
        rx.load(x + row * cols, cols, lane, 0.0f);
        rz.load(dz + row * cols, cols, lane, 0.0f);
        uint32_t keep_bits = 0;          // 4 bits per float4 of this lane (NV <= 8), written by the fused forward
        if (DROP) {
            keep_bits = __ldg(maskbits + row * 32 + lane);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const uint32_t k = keep_bits >> (4 * i);
                rx.v[i].x = (k & 1u) ? rx.v[i].x * drop.inv_keep : 0.0f;
                rx.v[i].y = (k & 2u) ? rx.v[i].y * drop.inv_keep : 0.0f;
                rx.v[i].z = (k & 4u) ? rx.v[i].z * drop.inv_keep : 0.0f;
                rx.v[i].w = (k & 8u) ? rx.v[i].w * drop.inv_keep : 0.0f;
            }
        }
        const float mu = mean[row], rs = rstd[row];
        float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < cols) {
                // rx ← y_hat ; rz stays dz
                rx.v[i].x = (rx.v[i].x - mu) * rs; rx.v[i].y = (rx.v[i].y - mu) * rs;
                rx.v[i].z = (rx.v[i].z - mu) * rs; rx.v[i].w = (rx.v[i].w - mu) * rs;
                float4 dg = *reinterpret_cast<float4*>(my + c), db = *reinterpret_cast<float4*>(my + cpad + c);
                dg.x += rz.v[i].x * rx.v[i].x; dg.y += rz.v[i].y * rx.v[i].y;
                dg.z += rz.v[i].z * rx.v[i].z; dg.w += rz.v[i].w * rx.v[i].w;
                db.x += rz.v[i].x; db.y += rz.v[i].y; db.z += rz.v[i].z; db.w += rz.v[i].w;
                *reinterpret_cast<float4*>(my + c) = dg;
                *reinterpret_cast<float4*>(my + cpad + c) = db;
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
                const float gx = rz.v[i].x * g.x, gy = rz.v[i].y * g.y, gz = rz.v[i].z * g.z, gw = rz.v[i].w * g.w;
                s1 += (gx + gy) + (gz + gw);
                s2 += (gx * rx.v[i].x + gy * rx.v[i].y) + (gz * rx.v[i].z + gw * rx.v[i].w);
                rz.v[i] = make_float4(gx, gy, gz, gw);   // rz ← g = dz * gamma
            }
        }
        s1 = warp_sum(s1) * inv_c;
        s2 = warp_sum(s2) * inv_c;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            rz.v[i].x = rs * (rz.v[i].x - s1 - rx.v[i].x * s2); rz.v[i].y = rs * (rz.v[i].y - s1 - rx.v[i].y * s2);
            rz.v[i].z = rs * (rz.v[i].z - s1 - rx.v[i].z * s2); rz.v[i].w = rs * (rz.v[i].w - s1 - rx.v[i].w * s2);
        }
        if (DROP) {       // DropOut.backward (normalizations.py:25-30) on the way out, then the residual branch
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const uint32_t k = keep_bits >> (4 * i);
                rz.v[i].x = (k & 1u) ? rz.v[i].x * drop.inv_keep : 0.0f;
                rz.v[i].y = (k & 2u) ? rz.v[i].y * drop.inv_keep : 0.0f;
                rz.v[i].z = (k & 4u) ? rz.v[i].z * drop.inv_keep : 0.0f;
                rz.v[i].w = (k & 8u) ? rz.v[i].w * drop.inv_keep : 0.0f;
            }
        }
        if (dskip != nullptr) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = (i * 32 + lane) * 4;
                if (c < cols) {
                    const float4 a = ld_stream(reinterpret_cast<const float4*>(dskip + row * cols + c));
                    rz.v[i].x += a.x; rz.v[i].y += a.y; rz.v[i].z += a.z; rz.v[i].w += a.w;
                }
            }
        }
        if (XSUM) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = (i * 32 + lane) * 4;
                if (c < cols) {
                    float4 a = *reinterpret_cast<float4*>(my + 2 * cpad + c);
                    a.x += rz.v[i].x; a.y += rz.v[i].y; a.z += rz.v[i].z; a.w += rz.v[i].w;
                    *reinterpret_cast<float4*>(my + 2 * cpad + c) = a;
                }
            }
        }
        rz.store(dx + row * cols, cols, lane);
    }
    // CTA reduce of the parameter-gradient partials
    __syncthreads();
    for (int c = threadIdx.x; c < NP * cpad; c += kRowThreads) {
        float acc = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) acc += sred[(size_t)w * NP * cpad + c];
        const int which = c / cpad, col = c - which * cpad;
        if (col < cols) partial[((size_t)blockIdx.x * NP + which) * cols + col] = acc;
    }
}

// generic backward: dx by multi-pass warps; parameter grads by a column-parallel second kernel
__global__ void __launch_bounds__(kRowThreads) layernorm_bwd_dx_generic(const float* dz, const float* x, const float* gamma,
                                                                        const float* mean, const float* rstd, float* dx,
                                                                        int64_t rows, int64_t cols) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + row * cols;
    const float* zr = dz + row * cols;
    float s1 = 0.0f, s2 = 0.0f;
    for (int64_t c = lane; c < cols; c += 32) {
        const float g = zr[c] * gamma[c];
        s1 += g;
        s2 += g * ((xr[c] - mu) * rs);
    }
    s1 = warp_sum(s1) / (float)cols;
    s2 = warp_sum(s2) / (float)cols;
    for (int64_t c = lane; c < cols; c += 32) {
        const float g = zr[c] * gamma[c];
        dx[row * cols + c] = rs * (g - s1 - ((xr[c] - mu) * rs) * s2);
    }
}
// partial[slab][0/1][col] = sum over the slab's rows of dz*y_hat / dz
__global__ void __launch_bounds__(256) layernorm_bwd_param_generic(const float* dz, const float* x, const float* mean,
                                                                   const float* rstd, float* partial, int64_t rows,
                                                                   int64_t cols, int64_t rows_per_slab) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
    const int64_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
    float a = 0.0f, b = 0.0f;
    for (int64_t r = r0; r < r1; ++r) {
        const float z = dz[r * cols + c];
        a += z * ((x[r * cols + c] - mean[r]) * rstd[r]);
        b += z;
    }
    partial[((size_t)blockIdx.y * 2 + 0) * cols + c] = a;
    partial[((size_t)blockIdx.y * 2 + 1) * cols + c] = b;
}

// out[c] = sum_s partial[s*slab_stride + c]  — second stage of every column reduction here.
// 32 columns x 8 slab lanes per CTA: coalesced 128-byte rows, the slab chain is 8x shorter than a
// thread-per-column loop, combined through shared memory.
// blockIdx.y = 1 (LayerNorm: dbeta) reads the partials `cols` further on and writes out2.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                              int64_t cols, int nslabs, int64_t slab_stride,
                                                              float* __restrict__ out2 = nullptr,
                                                              float* __restrict__ out3 = nullptr) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * 32 + tx;
    if (blockIdx.y == 1) { partial += cols; out = out2; }
    if (blockIdx.y == 2) { partial += 2 * cols; out = out3; }
    float acc = 0.0f;
    if (c < cols)
        for (int s = ty; s < nslabs; s += 8) acc += partial[(size_t)s * slab_stride + c];
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < cols) {
#pragma unroll
        for (int w = 1; w < 8; ++w) acc += sm[w][tx];
        out[c] = acc;
    }
}

// ================================================================ colsum
// One launch.  CTA = 8 warps over a slab of rows; a warp row covers 128 columns (float4 per lane).  Every
// CTA writes its partial row to the workspace; the LAST CTA of a column group to arrive (ticket counter)
// adds the partial rows in slab order — a fixed order, so the result is deterministic — and resets the
// counter.  `relu_y` != nullptr fuses ReLU.backward (activations.py:17-19): the value summed and written
// to `masked` is x where the forward output's sign bit is clear (the fused Dense epilogue stores -0.0 for
// negative pre-activations, +0.0 for zero ones, so "x >= 0" survives in the sign of y).
constexpr int kTicketSlots = 64, kTicketCols = 128;
__device__ unsigned int g_colsum_tickets[kTicketSlots * kTicketCols];

__global__ void __launch_bounds__(kRowThreads) colsum_kernel(const float* __restrict__ x, float* __restrict__ partial,
                                                            float* __restrict__ out, int64_t rows, int64_t cols,
                                                            int64_t rows_per_slab, unsigned ticket_base,
                                                            const float* __restrict__ relu_y, float* __restrict__ masked) {
    pdl_trigger();
    pdl_wait();
    __shared__ float4 sm[kWarpsPerCta][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = ((int64_t)blockIdx.x * 32 + lane) * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
    const int64_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
    float4 acc = make_float4(0, 0, 0, 0);
    if (c < cols) {
        constexpr int U = 4;       // rows in flight per thread: the loads of a batch are all issued before the adds
        int64_t r = r0 + warp;
        for (; r + (U - 1) * kWarpsPerCta < r1; r += U * kWarpsPerCta) {
            float4 v[U], y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_stream(reinterpret_cast<const float4*>(x + (r + u * kWarpsPerCta) * cols + c));
            if (relu_y != nullptr) {
#pragma unroll
                for (int u = 0; u < U; ++u) y[u] = ld_stream(reinterpret_cast<const float4*>(relu_y + (r + u * kWarpsPerCta) * cols + c));
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    v[u].x = (__float_as_uint(y[u].x) >> 31) ? 0.0f : v[u].x; v[u].y = (__float_as_uint(y[u].y) >> 31) ? 0.0f : v[u].y;
                    v[u].z = (__float_as_uint(y[u].z) >> 31) ? 0.0f : v[u].z; v[u].w = (__float_as_uint(y[u].w) >> 31) ? 0.0f : v[u].w;
                    st_stream(reinterpret_cast<float4*>(masked + (r + u * kWarpsPerCta) * cols + c), v[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; r < r1; r += kWarpsPerCta) {
            float4 v = ld_stream(reinterpret_cast<const float4*>(x + r * cols + c));
            if (relu_y != nullptr) {
                const float4 y = ld_stream(reinterpret_cast<const float4*>(relu_y + r * cols + c));
                v.x = (__float_as_uint(y.x) >> 31) ? 0.0f : v.x; v.y = (__float_as_uint(y.y) >> 31) ? 0.0f : v.y;
                v.z = (__float_as_uint(y.z) >> 31) ? 0.0f : v.z; v.w = (__float_as_uint(y.w) >> 31) ? 0.0f : v.w;
                st_stream(reinterpret_cast<float4*>(masked + r * cols + c), v);
            }
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    sm[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && c < cols) {
#pragma unroll
        for (int w = 1; w < kWarpsPerCta; ++w) {
            const float4 v = sm[w][lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * cols + c) = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&g_colsum_tickets[ticket_base + blockIdx.x], 1u);
        is_last = (t == gridDim.y - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    acc = make_float4(0, 0, 0, 0);
    if (c < cols)
        for (int sl = warp; sl < (int)gridDim.y; sl += kWarpsPerCta) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + (size_t)sl * cols + c));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    sm[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && c < cols) {
#pragma unroll
        for (int w = 1; w < kWarpsPerCta; ++w) {
            const float4 v = sm[w][lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(out + c) = acc;
    }
    if (threadIdx.x == 0) g_colsum_tickets[ticket_base + blockIdx.x] = 0;     // ready for the next launch on this slot
}
// ReLU.backward + bias gradient for an activation that exists only as split-bf16 planes (the FFN hidden activation out of
// the first FFN GEMM's epilogue, npm_gemm_desc.c_split): the gate is the sign bit of y's bf16 hi plane (bf16_rn keeps the
// sign, -0.0 included), the masked gradient goes out as bf16 hi / mid planes too — the A operand of the dX GEMM and the
// B operand of the dW GEMM that follow, which then land it by TMA without converting.  Same slab / ticket scheme as
// colsum_kernel.  Bytes per element: 4 (dy) + 2 (y hi) + 4 (dz planes) = 10 against 12 of the fp32 form.
__device__ __forceinline__ void split_pack2(float p0, float p1, uint32_t& hi, uint32_t& mid) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(p1), "f"(p0));
    const float r0 = p0 - __uint_as_float(hi << 16), r1 = p1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(r1), "f"(r0));
}
__global__ void __launch_bounds__(kRowThreads) relu_bwd_colsum_planes_kernel(const float* __restrict__ dy, const uint2* __restrict__ y_hi,
                                                                            uint2* __restrict__ dz, int64_t plane4,
                                                                            float* __restrict__ partial, float* __restrict__ out,
                                                                            int64_t rows, int64_t cols, int64_t rows_per_slab,
                                                                            unsigned ticket_base) {
    pdl_trigger();
    pdl_wait();
    __shared__ float4 sm[kWarpsPerCta][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = ((int64_t)blockIdx.x * 32 + lane) * 4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
    const int64_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
    float4 acc = make_float4(0, 0, 0, 0);
    if (c < cols) {
        constexpr int U = 4;       // rows in flight per thread (measured at [8192, 4096]: 2 -> 80.7 us, 4 -> 78.9 us, 8 -> 86.9 us)
        auto one = [&](float4 v, uint2 y, int64_t r) {
            v.x = (y.x & 0x8000u) ? 0.0f : v.x; v.y = (y.x & 0x80000000u) ? 0.0f : v.y;
            v.z = (y.y & 0x8000u) ? 0.0f : v.z; v.w = (y.y & 0x80000000u) ? 0.0f : v.w;
            uint2 h, m;
            split_pack2(v.x, v.y, h.x, m.x);
            split_pack2(v.z, v.w, h.y, m.y);
            const int64_t i = (r * cols + c) >> 2;
            dz[i] = h;
            dz[plane4 + i] = m;
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        };
        int64_t r = r0 + warp;
        for (; r + (U - 1) * kWarpsPerCta < r1; r += U * kWarpsPerCta) {
            float4 v[U];
            uint2 y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_stream(reinterpret_cast<const float4*>(dy + (r + u * kWarpsPerCta) * cols + c));
#pragma unroll
            for (int u = 0; u < U; ++u) y[u] = __ldg(y_hi + (((r + u * kWarpsPerCta) * cols + c) >> 2));
#pragma unroll
            for (int u = 0; u < U; ++u) one(v[u], y[u], r + u * kWarpsPerCta);
        }
        for (; r < r1; r += kWarpsPerCta)
            one(ld_stream(reinterpret_cast<const float4*>(dy + r * cols + c)), __ldg(y_hi + ((r * cols + c) >> 2)), r);
    }
    sm[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && c < cols) {
#pragma unroll
        for (int w = 1; w < kWarpsPerCta; ++w) {
            const float4 v = sm[w][lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(partial + (size_t)blockIdx.y * cols + c) = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&g_colsum_tickets[ticket_base + blockIdx.x], 1u);
        is_last = (t == gridDim.y - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    acc = make_float4(0, 0, 0, 0);
    if (c < cols)
        for (int sl = warp; sl < (int)gridDim.y; sl += kWarpsPerCta) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + (size_t)sl * cols + c));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    sm[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && c < cols) {
#pragma unroll
        for (int w = 1; w < kWarpsPerCta; ++w) {
            const float4 v = sm[w][lane];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(out + c) = acc;
    }
    if (threadIdx.x == 0) g_colsum_tickets[ticket_base + blockIdx.x] = 0;
}
__global__ void __launch_bounds__(256) colsum_generic(const float* x, float* partial, int64_t rows, int64_t cols,
                                                      int64_t rows_per_slab) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
    const int64_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
    float a = 0.0f;
    for (int64_t r = r0; r < r1; ++r) a += x[r * cols + c];
    partial[(size_t)blockIdx.y * cols + c] = a;
}

inline int colsum_slabs(int64_t rows, int64_t cols) {
    const int64_t col_ctas = (cols + 127) / 128;
    int64_t want = ((int64_t)num_sms() * 4 + col_ctas - 1) / col_ctas;   // ~4 CTAs per SM in total
    const int64_t max_by_rows = (rows + 63) / 64;                        // ≥ 64 rows per slab
    if (want > max_by_rows) want = max_by_rows;
    if (want < 1) want = 1;
    if (want > 1024) want = 1024;
    return (int)want;
}

}  // namespace

// internal C++ entry used by linear/conv/mha wrappers
size_t colsum_workspace_bytes(int64_t rows, int64_t cols) {
    return (size_t)colsum_slabs(rows, cols) * (size_t)cols * sizeof(float);
}
static unsigned next_ticket_base(int64_t col_ctas) {
    static std::atomic<unsigned> n{0};
    (void)col_ctas;
    return (n.fetch_add(1u) % kTicketSlots) * kTicketCols;
}

// out[c] = sum_r x[r,c]; with relu_y/masked: out[c] = sum_r m[r,c], m = x where sign(relu_y) clear else 0, m → masked
int colsum_relu_launch(const float* x, float* out, int64_t rows, int64_t cols, void* workspace, const float* relu_y,
                       float* masked, cudaStream_t s) {
    NPM_REQUIRE(rows > 0 && cols > 0, "colsum: empty input");
    NPM_REQUIRE(workspace != nullptr, "colsum: workspace is NULL");
    const int slabs = colsum_slabs(rows, cols);
    const int64_t rps = (rows + slabs - 1) / slabs;
    float* partial = reinterpret_cast<float*>(workspace);
    const int64_t col_ctas = (cols + 127) / 128;
    const bool vec = (cols & 3) == 0 && aligned16(x) && aligned16(partial) && aligned16(out) &&
                     (relu_y == nullptr || (aligned16(relu_y) && aligned16(masked)));
    if (vec && col_ctas <= kTicketCols) {
        launch_pdl(colsum_kernel, dim3((unsigned)col_ctas, slabs), dim3(kRowThreads), 0, s, 1, x, partial, out, rows, cols, rps,
                   next_ticket_base(col_ctas), relu_y, masked);
        count_launch();
        return check_launch("colsum_kernel");
    }
    NPM_REQUIRE(relu_y == nullptr, "relu_bwd_colsum: needs 16-byte aligned pointers and cols %% 4 == 0");
    colsum_generic<<<dim3((unsigned)((cols + 255) / 256), slabs), 256, 0, s>>>(x, partial, rows, cols, rps);
    count_launch();
    int rc = check_launch("colsum_generic");
    if (rc) return rc;
    reduce_partials_kernel<<<(unsigned)((cols + 31) / 32), 256, 0, s>>>(partial, out, cols, slabs, cols);
    count_launch();
    return check_launch("reduce_partials_kernel");
}
int relu_bwd_colsum_planes_launch(const void* y_hi, const float* dy, void* dz_planes, int64_t plane, float* db, int64_t rows,
                                  int64_t cols, void* workspace, cudaStream_t s) {
    NPM_REQUIRE(rows > 0 && cols > 0 && workspace != nullptr, "relu_bwd_colsum_planes: empty input or NULL workspace");
    const int64_t col_ctas = (cols + 127) / 128;
    if (!((cols & 3) == 0 && (plane & 3) == 0 && col_ctas <= kTicketCols && aligned16(dy) && aligned16(db) && aligned16(workspace) &&
          (reinterpret_cast<uintptr_t>(y_hi) & 7u) == 0 && (reinterpret_cast<uintptr_t>(dz_planes) & 7u) == 0)) {
        set_error("relu_bwd_colsum_planes: needs cols %% 4 == 0, plane %% 4 == 0, cols <= %d and aligned pointers", kTicketCols * 128);
        return NPM_ERR_UNSUPPORTED;
    }
    const int slabs = colsum_slabs(rows, cols);
    const int64_t rps = (rows + slabs - 1) / slabs;
    launch_pdl(relu_bwd_colsum_planes_kernel, dim3((unsigned)col_ctas, slabs), dim3(kRowThreads), 0, s, 1, dy,
               reinterpret_cast<const uint2*>(y_hi), reinterpret_cast<uint2*>(dz_planes), plane / 4, reinterpret_cast<float*>(workspace), db,
               rows, cols, rps, next_ticket_base(col_ctas));
    count_launch();
    return check_launch("relu_bwd_colsum_planes_kernel");
}
int colsum_launch(const float* x, float* out, int64_t rows, int64_t cols, void* workspace, cudaStream_t s) {
    return colsum_relu_launch(x, out, rows, cols, workspace, nullptr, nullptr, s);
}

}  // namespace npm

using namespace npm;

#define DISPATCH_NV(nv, CALL)                       \
    switch (nv) {                                   \
        case 1:  { constexpr int NV = 1;  CALL; } break;  \
        case 2:  { constexpr int NV = 2;  CALL; } break;  \
        case 4:  { constexpr int NV = 4;  CALL; } break;  \
        case 8:  { constexpr int NV = 8;  CALL; } break;  \
        case 16: { constexpr int NV = 16; CALL; } break;  \
        default: { constexpr int NV = 32; CALL; } break;  \
    }

extern "C" {

size_t npm_colsum_workspace(int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return 0;
    return colsum_workspace_bytes(rows, cols);
}
int npm_colsum(const float* x, float* out, int64_t rows, int64_t cols, void* workspace, npm_stream_t stream) {
    return colsum_launch(x, out, rows, cols, workspace, (cudaStream_t)stream);
}

int npm_relu_bwd_colsum(const float* y, const float* dy, float* dx, float* db, int64_t rows, int64_t cols,
                        void* workspace, npm_stream_t stream) {
    NPM_REQUIRE(y && dy && dx && db, "relu_bwd_colsum: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = (cols & 3) == 0 && cols <= (int64_t)kTicketCols * 128 && aligned16(y) && aligned16(dy) &&
                     aligned16(dx) && aligned16(db) && aligned16(workspace);
    if (vec) return colsum_relu_launch(dy, db, rows, cols, workspace, y, dx, s);
    int rc = npm_relu_bwd_y(y, dy, dx, rows * cols, stream);
    if (rc) return rc;
    return colsum_launch(dx, db, rows, cols, workspace, s);
}

int npm_relu_bwd_colsum_planes(const void* y_hi, const float* dy, void* dz_planes, int64_t plane, float* db, int64_t rows,
                               int64_t cols, void* workspace, npm_stream_t stream) {
    NPM_REQUIRE(y_hi && dy && dz_planes && db, "relu_bwd_colsum_planes: NULL pointer");
    return relu_bwd_colsum_planes_launch(y_hi, dy, dz_planes, plane, db, rows, cols, workspace, (cudaStream_t)stream);
}

int npm_softmax_fwd(const float* x, float* y, int64_t rows, int64_t cols, npm_stream_t stream) {
    if (rows <= 0 || cols <= 0) return NPM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
    const int nv = nv_for(cols);
    if (nv && (cols & 3) == 0 && aligned16(x) && aligned16(y)) {
        DISPATCH_NV(nv, (softmax_fwd_kernel<NV><<<grid, kRowThreads, 0, s>>>(x, y, rows, (int)cols)));
    } else {
        softmax_fwd_generic<<<grid, kRowThreads, 0, s>>>(x, y, rows, cols);
    }
    count_launch();
    return check_launch("softmax_fwd");
}

int npm_softmax_bwd(const float* y, const float* dy, float* dx, int64_t rows, int64_t cols, float scale,
                    npm_stream_t stream) {
    if (rows <= 0 || cols <= 0) return NPM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
    int nv = nv_for(cols);
    if (nv > 16) nv = 0;   // two rows in registers: cap at 2048 columns
    if (nv && (cols & 3) == 0 && aligned16(y) && aligned16(dy) && aligned16(dx)) {
        switch (nv) {
            case 1: softmax_bwd_kernel<1><<<grid, kRowThreads, 0, s>>>(y, dy, dx, rows, (int)cols, scale); break;
            case 2: softmax_bwd_kernel<2><<<grid, kRowThreads, 0, s>>>(y, dy, dx, rows, (int)cols, scale); break;
            case 4: softmax_bwd_kernel<4><<<grid, kRowThreads, 0, s>>>(y, dy, dx, rows, (int)cols, scale); break;
            case 8: softmax_bwd_kernel<8><<<grid, kRowThreads, 0, s>>>(y, dy, dx, rows, (int)cols, scale); break;
            default: softmax_bwd_kernel<16><<<grid, kRowThreads, 0, s>>>(y, dy, dx, rows, (int)cols, scale); break;
        }
    } else {
        softmax_bwd_generic<<<grid, kRowThreads, 0, s>>>(y, dy, dx, rows, cols, scale);
    }
    count_launch();
    return check_launch("softmax_bwd");
}

int npm_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* out, float* mean, float* rstd,
                      int64_t rows, int64_t cols, float epsilon, npm_stream_t stream) {
    if (rows <= 0 || cols <= 0) return NPM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
    int nv = nv_for(cols);
    if (nv > 16) nv = 0;
    if (nv && (cols & 3) == 0 && aligned16(x) && aligned16(out) && aligned16(gamma) && aligned16(beta)) {
        switch (nv) {
            case 1: layernorm_fwd_kernel<1><<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon); break;
            case 2: layernorm_fwd_kernel<2><<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon); break;
            case 4: layernorm_fwd_kernel<4><<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon); break;
            case 8: layernorm_fwd_kernel<8><<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon); break;
            default: layernorm_fwd_kernel<16><<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon); break;
        }
    } else {
        layernorm_fwd_generic<<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, cols, epsilon);
    }
    count_launch();
    return check_launch("layernorm_fwd");
}

static int ln_bwd_grid(int64_t rows) {
    int64_t g = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t cap = (int64_t)num_sms() * 2;      // two resident CTAs per SM, one wave
    return (int)(g < cap ? g : cap);
}
static int ln_generic_slabs(int64_t rows) {
    int64_t s = (rows + 255) / 256;
    if (s > 512) s = 512;
    if (s < 1) s = 1;
    return (int)s;
}
static bool ln_fast(int64_t cols) { return (cols & 3) == 0 && cols <= 1024; }

size_t npm_layernorm_bwd_workspace(int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t slabs = ln_fast(cols) ? ln_bwd_grid(rows) : ln_generic_slabs(rows);
    return (size_t)slabs * 3 * (size_t)cols * sizeof(float);    // 3 partial rows per slab: dgamma, dbeta, column sums of dx
}

int npm_layernorm_bwd(const float* dz, const float* x, const float* gamma, const float* mean, const float* rstd,
                      float* dx, float* dgamma, float* dbeta, int64_t rows, int64_t cols, void* workspace,
                      npm_stream_t stream) {
    if (rows <= 0 || cols <= 0) return NPM_OK;
    NPM_REQUIRE(workspace != nullptr, "layernorm_bwd: workspace is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(workspace);
    int slabs;
    const bool fast = ln_fast(cols) && aligned16(dz) && aligned16(x) && aligned16(gamma) && aligned16(dx);
    if (fast) {
        slabs = ln_bwd_grid(rows);
        const int nv = nv_for(cols);   // ≤ 8 for cols ≤ 1024
        const size_t smem = (size_t)kWarpsPerCta * 2 * nv * 128 * sizeof(float);
        switch (nv) {
            case 1: layernorm_bwd_kernel<1><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols); break;
            case 2: layernorm_bwd_kernel<2><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols); break;
            case 4: layernorm_bwd_kernel<4><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols); break;
            default: {
                static bool configured = false;
                if (!configured) {
                    cudaFuncSetAttribute(layernorm_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    configured = true;
                }
                layernorm_bwd_kernel<8><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols);
            } break;
        }
        count_launch();
        int rc = check_launch("layernorm_bwd_kernel");
        if (rc) return rc;
    } else {
        slabs = ln_generic_slabs(rows);
        // the workspace was sized by npm_layernorm_bwd_workspace(); if the fast path was refused only
        // because of alignment, the generic slab count may exceed it — clamp to what was promised.
        if (ln_fast(cols)) { const int promised = ln_bwd_grid(rows); if (slabs > promised) slabs = promised; }
        const int64_t rps = (rows + slabs - 1) / slabs;
        const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
        layernorm_bwd_dx_generic<<<grid, kRowThreads, 0, s>>>(dz, x, gamma, mean, rstd, dx, rows, cols);
        count_launch();
        int rc = check_launch("layernorm_bwd_dx_generic");
        if (rc) return rc;
        slabs = (int)((rows + rps - 1) / rps);
        layernorm_bwd_param_generic<<<dim3((unsigned)((cols + 255) / 256), slabs), 256, 0, s>>>(dz, x, mean, rstd, partial,
                                                                                             rows, cols, rps);
        count_launch();
        rc = check_launch("layernorm_bwd_param_generic");
        if (rc) return rc;
    }
    const unsigned g2 = (unsigned)((cols + 31) / 32);
    reduce_partials_kernel<<<dim3(g2, 2), 256, 0, s>>>(partial, dgamma, cols, slabs, 2 * cols, dbeta);
    count_launch();
    return check_launch("layernorm_bwd_reduce");
}

// ---- DropOut -> LayerNormalization fused (pre-norm transformer blocks) ----
int npm_dropout_layernorm_fused(int64_t rows, int64_t cols) {
    (void)rows;
    return ln_fast(cols) && nv_for(cols) > 0 ? 1 : 0;
}

static DropArgs make_drop(float keep_prob, uint64_t seed, uint64_t offset) {
    DropArgs d;
    d.keep_prob = keep_prob;
    d.inv_keep = 1.0f / keep_prob;
    d.thr = (uint64_t)((double)keep_prob * 4294967296.0);
    d.seed = seed;
    d.offset = offset;
    return d;
}

size_t npm_dropout_layernorm_mask_bytes(int64_t rows, int64_t cols) {
    (void)cols;
    return rows > 0 ? (size_t)rows * 32 * sizeof(uint32_t) : 0;     // 4 keep bits per float4, one word per lane and row
}

// LayerNormalization forward (and the fused DropOut -> LayerNormalization) whose result exists only as split-bf16 planes.
// keep_prob >= 1 (or maskbits == NULL) = no dropout.  NPM_ERR_UNSUPPORTED unless cols % 4 == 0, cols <= 1024, plane % 4 == 0.
int npm_layernorm_fwd_planes(const float* x, const float* gamma, const float* beta, void* out_planes, int64_t plane, float* mean,
                             float* rstd, uint32_t* maskbits, int64_t rows, int64_t cols, float epsilon, float keep_prob,
                             uint64_t seed, uint64_t offset, npm_stream_t stream) {
    if (rows <= 0 || cols <= 0) return NPM_OK;
    NPM_REQUIRE(x && gamma && beta && out_planes && mean && rstd, "layernorm_fwd_planes: NULL pointer");
    const bool drop = maskbits != nullptr && keep_prob < 1.0f;
    NPM_REQUIRE(!drop || keep_prob > 0.0f, "layernorm_fwd_planes: keep_prob %g out of (0,1]", keep_prob);
    const int nv = nv_for(cols);
    if (!(nv && nv <= 8 && (cols & 3) == 0 && (plane & 3) == 0 && (rows * cols) % 4 == 0 && aligned16(x) && aligned16(gamma) &&
          aligned16(beta) && (reinterpret_cast<uintptr_t>(out_planes) & 7u) == 0 && (!drop || (offset & 3) == 0))) {
        set_error("layernorm_fwd_planes: needs cols %% 4 == 0, cols <= 1024, plane %% 4 == 0, offset %% 4 == 0 and aligned pointers");
        return NPM_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
    float* out = reinterpret_cast<float*>(out_planes);
    const int64_t plane4 = plane / 4;
    if (drop) {
        const DropArgs d = make_drop(keep_prob, seed, offset);
        switch (nv) {
            case 1: launch_pdl(layernorm_fwd_kernel<1, true, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, maskbits, plane4); break;
            case 2: launch_pdl(layernorm_fwd_kernel<2, true, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, maskbits, plane4); break;
            case 4: launch_pdl(layernorm_fwd_kernel<4, true, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, maskbits, plane4); break;
            default: launch_pdl(layernorm_fwd_kernel<8, true, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, maskbits, plane4); break;
        }
    } else {
        const DropArgs d = DropArgs{};
        uint32_t* nomask = nullptr;
        switch (nv) {
            case 1: launch_pdl(layernorm_fwd_kernel<1, false, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, nomask, plane4); break;
            case 2: launch_pdl(layernorm_fwd_kernel<2, false, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, nomask, plane4); break;
            case 4: launch_pdl(layernorm_fwd_kernel<4, false, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, nomask, plane4); break;
            default: launch_pdl(layernorm_fwd_kernel<8, false, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, nomask, plane4); break;
        }
    }
    count_launch();
    return check_launch("layernorm_fwd_planes");
}

int npm_dropout_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* out, float* mean, float* rstd,
                              uint32_t* maskbits, int64_t rows, int64_t cols, float epsilon, float keep_prob, uint64_t seed,
                              uint64_t offset, npm_stream_t stream) {
    NPM_REQUIRE(maskbits != nullptr, "dropout_layernorm_fwd: maskbits is NULL");
    if (rows <= 0 || cols <= 0) return NPM_OK;
    NPM_REQUIRE(keep_prob > 0.0f && keep_prob <= 1.0f, "dropout_layernorm: keep_prob %g out of (0,1]", keep_prob);
    if (!(npm_dropout_layernorm_fused(rows, cols) && (offset & 3) == 0 && aligned16(x) && aligned16(out) &&
          aligned16(gamma) && aligned16(beta))) {
        set_error("dropout_layernorm_fwd: needs cols %% 4 == 0, cols <= 1024, offset %% 4 == 0 and 16-byte aligned pointers");
        return NPM_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
    const DropArgs d = make_drop(keep_prob, seed, offset);
    switch (nv_for(cols)) {
        case 1: layernorm_fwd_kernel<1, true><<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, maskbits); break;
        case 2: layernorm_fwd_kernel<2, true><<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, maskbits); break;
        case 4: layernorm_fwd_kernel<4, true><<<grid, kRowThreads, 0, s>>>(x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, maskbits); break;
        default: launch_pdl(layernorm_fwd_kernel<8, true>, dim3(grid), dim3(kRowThreads), 0, s, 1, x, gamma, beta, out, mean, rstd, rows, (int)cols, epsilon, d, maskbits, (int64_t)0); break;
    }
    count_launch();
    return check_launch("dropout_layernorm_fwd");
}

int npm_dropout_layernorm_bwd(const float* dz, const float* x, const float* gamma, const float* mean, const float* rstd,
                              const uint32_t* maskbits, const float* dskip, float* dx, float* dgamma, float* dbeta,
                              int64_t rows, int64_t cols, float keep_prob, void* workspace, npm_stream_t stream) {
    return npm_dropout_layernorm_bwd_colsum(dz, x, gamma, mean, rstd, maskbits, dskip, dx, dgamma, dbeta, nullptr, rows, cols,
                                            keep_prob, workspace, stream);
}

int npm_dropout_layernorm_bwd_colsum(const float* dz, const float* x, const float* gamma, const float* mean,
                                     const float* rstd, const uint32_t* maskbits, const float* dskip, float* dx,
                                     float* dgamma, float* dbeta, float* dx_colsum, int64_t rows, int64_t cols,
                                     float keep_prob, void* workspace, npm_stream_t stream) {
    NPM_REQUIRE(maskbits != nullptr, "dropout_layernorm_bwd: maskbits is NULL");
    const uint64_t seed = 0, offset = 0;      // the mask comes from the forward pass
    if (rows <= 0 || cols <= 0) return NPM_OK;
    NPM_REQUIRE(workspace != nullptr, "dropout_layernorm_bwd: workspace is NULL");
    NPM_REQUIRE(keep_prob > 0.0f && keep_prob <= 1.0f, "dropout_layernorm: keep_prob %g out of (0,1]", keep_prob);
    if (!(npm_dropout_layernorm_fused(rows, cols) && (offset & 3) == 0 && aligned16(dz) && aligned16(x) &&
          aligned16(gamma) && aligned16(dx) && (dskip == nullptr || aligned16(dskip)))) {
        set_error("dropout_layernorm_bwd: needs cols %% 4 == 0, cols <= 1024, offset %% 4 == 0 and 16-byte aligned pointers");
        return NPM_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(workspace);
    const int slabs = ln_bwd_grid(rows);
    const int nv = nv_for(cols);
    const bool xsum = dx_colsum != nullptr;
    const int np = xsum ? 3 : 2;
    const size_t smem = (size_t)kWarpsPerCta * np * nv * 128 * sizeof(float);
    const DropArgs d = make_drop(keep_prob, seed, offset);
    if (xsum && nv == 8) {
        static bool configured3 = false;
        if (!configured3) {
            cudaFuncSetAttribute(layernorm_bwd_kernel<8, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            configured3 = true;
        }
        launch_pdl(layernorm_bwd_kernel<8, true, true>, dim3(slabs), dim3(kRowThreads), smem, s, 1, dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols, d, dskip, maskbits);
    } else if (xsum && nv == 4) {
        layernorm_bwd_kernel<4, true, true><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols, d, dskip, maskbits);
    } else if (xsum && nv == 2) {
        layernorm_bwd_kernel<2, true, true><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols, d, dskip, maskbits);
    } else if (xsum) {
        layernorm_bwd_kernel<1, true, true><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols, d, dskip, maskbits);
    } else
    switch (nv) {
        case 1: layernorm_bwd_kernel<1, true><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols, d, dskip, maskbits); break;
        case 2: layernorm_bwd_kernel<2, true><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols, d, dskip, maskbits); break;
        case 4: layernorm_bwd_kernel<4, true><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols, d, dskip, maskbits); break;
        default: {
            static bool configured = false;
            if (!configured) {
                cudaFuncSetAttribute(layernorm_bwd_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                configured = true;
            }
            layernorm_bwd_kernel<8, true><<<slabs, kRowThreads, smem, s>>>(dz, x, gamma, mean, rstd, dx, partial, rows, (int)cols, d, dskip, maskbits);
        } break;
    }
    count_launch();
    int rc = check_launch("dropout_layernorm_bwd");
    if (rc) return rc;
    const unsigned g2 = (unsigned)((cols + 31) / 32);
    launch_pdl(reduce_partials_kernel, dim3(g2, np), dim3(256), 0, s, 1, (const float*)partial, dgamma, cols, slabs, (int64_t)np * cols, dbeta, dx_colsum);
    count_launch();
    return check_launch("dropout_layernorm_bwd_reduce");
}

}  // extern "C"
