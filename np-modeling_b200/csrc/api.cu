// api.cu — library state, error plumbing, GEMM dispatch and the Linear / attention-core
// entry points of the C-ABI (include/npm_b200.h).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace npm {

// ------------------------------------------------------------------- state
static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_precision{NPM_PREC_BF16X3};     // the default mode meets rtol 1e-3 / atol 1e-4 (as does 3xTF32)

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return NPM_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NPM_ERR_CUDA;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int current_precision() { return g_precision.load(); }

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
            sms = v;
        else
            return 148;   // B200; do not cache a failed query
    }
    return sms;
}

// implemented in gemm_tc.cu / gemm_simt.cu / rowops.cu
bool gemm_tc_supported(const npm_gemm_desc& d);
int gemm_tc_launch(const npm_gemm_desc& d, int precision, cudaStream_t stream);
int gemm_simt_launch(const npm_gemm_desc& d, cudaStream_t stream);
bool gemm_bx_supported(const npm_gemm_desc& d);
int gemm_bx_launch(const npm_gemm_desc& d, int nterms, cudaStream_t stream);
size_t colsum_workspace_bytes(int64_t rows, int64_t cols);
// attn_fwd.cu / attn_bwd.cu
bool attn_fused_supported(int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv);
int attn_fwd_launch(const void* q, const void* k, const void* v, float* o, float* lse, int64_t B, int64_t H,
                    int64_t Sq, int64_t Skv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t plq, int64_t plk, int64_t plv,
                    int causal, bool bx, int nterms, cudaStream_t stream);
size_t attn_bwd_scratch_bytes(int64_t B, int64_t H, int64_t Sq, bool bx);
int attn_split_launch(const float* x, int64_t ld, void* planes, int64_t rows, int64_t HD, cudaStream_t stream);
int attn_bwd_launch(const void* q, const void* k, const void* v, const float* o, const float* d_o, const float* lse,
                    float* dq, float* dk, float* dv, float* dsum, int64_t B, int64_t H, int64_t Sq, int64_t Skv,
                    int64_t ldq, int64_t ldk, int64_t ldv, int64_t plq, int64_t plk, int64_t plv, int64_t lddq, int64_t lddk,
                    int64_t lddv, int causal, bool bx, int nterms, bool do_ready, cudaStream_t stream);
int attn_scores_from_lse_launch(float* p, const float* lse, int64_t rows, int64_t cols, cudaStream_t stream);
int colsum_launch(const float* x, float* out, int64_t rows, int64_t cols, void* workspace, cudaStream_t s);

// scores[bh, s, t] = -inf for t > s (before the row softmax): the materialised-scores path of the causal mask
__global__ void __launch_bounds__(256) causal_mask_kernel(float* __restrict__ p, int64_t rows, int64_t Sq, int64_t Skv) {
    const int64_t n = rows * Skv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i % Skv, s = (i / Skv) % Sq;
        if (t > s) p[i] = -INFINITY;
    }
}

bool pdl_enabled() {
    static const bool on = !(getenv("NPM_NO_PDL") && getenv("NPM_NO_PDL")[0] != '0');
    return on;
}

static int require_sm100() {
    static int ok = -1;
    if (ok < 0) {
        int dev = 0, major = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
            set_error("no usable CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
            return NPM_ERR_CUDA;
        }
        ok = (major == 10) ? 1 : 0;
    }
    if (!ok) {
        set_error("libnpm_b200 is built for sm_100a (B200) only; current device is not compute capability 10.x");
        return NPM_ERR_UNSUPPORTED;
    }
    return NPM_OK;
}

int gemm_dispatch(const npm_gemm_desc& d, cudaStream_t stream) {
    NPM_REQUIRE((d.a || d.a_split) && (d.b || d.b_split) && (d.c || d.c_split), "gemm: NULL operand");
    NPM_REQUIRE(d.m > 0 && d.n > 0 && d.k > 0, "gemm: empty problem m=%lld n=%lld k=%lld", (long long)d.m,
                (long long)d.n, (long long)d.k);
    NPM_REQUIRE((d.a_rs == 1 || d.a_cs == 1) && (d.b_rs == 1 || d.b_cs == 1),
                "gemm: A and B need one unit stride each");
    NPM_REQUIRE(d.residual == nullptr || (d.nb1 <= 1 && d.nb2 <= 1 && d.ldr >= d.n && !(d.flags & NPM_GEMM_RELU)),
                "gemm: residual needs an unbatched problem, ldr >= n and no ReLU");
    int rc = require_sm100();
    if (rc) return rc;
    int prec = d.precision >= 0 ? d.precision : g_precision.load();
    const bool bx_mode = prec == NPM_PREC_BF16X3 || prec == NPM_PREC_BF16;
    static const bool bx_off = getenv("NPM_GEMM_NO_BX") != nullptr;      // A/B switch for tools/
    NPM_REQUIRE(d.a_colsum == nullptr || (bx_mode && !bx_off && gemm_bx_supported(d) && d.a_rs == 1 && d.nb1 <= 1 && d.nb2 <= 1),
                "gemm: a_colsum is served by the split-bf16 kernel only (precision bf16x3 / bf16, m-contiguous A, unbatched, m > 128)");
    if ((d.c_split != nullptr || d.a == nullptr || d.b == nullptr || d.rowdot_x != nullptr) && !(bx_mode && !bx_off && gemm_bx_supported(d))) {
        set_error("gemm: split-bf16 input / output planes are served by the split-bf16 kernel only (precision bf16x3 / bf16, m > 128)");
        return NPM_ERR_UNSUPPORTED;
    }
    if (bx_mode) {
        // split-bf16 CTA-pair kernel; problems it does not take (a single row tile, ragged 16-byte chunks) run the
        // TF32 kernels at the same or better accuracy class (3xTF32 for bf16x3, one TF32 pass for bf16)
        if (!bx_off && gemm_bx_supported(d)) return gemm_bx_launch(d, prec == NPM_PREC_BF16X3 ? 3 : 1, stream);
        prec = prec == NPM_PREC_BF16X3 ? NPM_PREC_3XTF32 : NPM_PREC_TF32;
    }
    if (prec != NPM_PREC_FP32 && gemm_tc_supported(d)) return gemm_tc_launch(d, prec, stream);
    return gemm_simt_launch(d, stream);
}

// Which implementation serves an attention-core call (the value npm_mha_core_path() reports and npm_mha_strides.path
// pins): head dim 64 runs the fused online-softmax kernels — in TF32 mode with tf32 operands (1), in the split-bf16
// modes with pre-split bf16 operands (2); the 3xTF32 / fp32 modes and other head dims run the batched GEMM + softmax
// chain with the scores materialised (0).
enum { ATTN_MATERIALISED = 0, ATTN_FUSED_TF32 = 1, ATTN_FUSED_BX = 2 };
static int attn_path_now(int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv) {
    static const bool off = getenv("NPM_ATTN_UNFUSED") != nullptr;     // A/B switch for tools/ and tests
    if (off || !attn_fused_supported(B, H, Sq, Skv, dk, dv)) return ATTN_MATERIALISED;
    const int prec = g_precision.load();
    if (prec == NPM_PREC_TF32) return ATTN_FUSED_TF32;
    if (prec == NPM_PREC_BF16X3 || prec == NPM_PREC_BF16) return ATTN_FUSED_BX;      // bf16: the same kernels, hi*hi term only
    return ATTN_MATERIALISED;
}
// the path a call runs: the one pinned in the strides struct (1 + path), else the current mode's
static int attn_path_of(const npm_mha_strides* ld, int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv) {
    if (ld && ld->path > 0) return (int)ld->path - 1;
    return attn_path_now(B, H, Sq, Skv, dk, dv);
}
static int attn_terms() { return g_precision.load() == NPM_PREC_BF16 ? 1 : 3; }
static size_t pad256(size_t n) { return (n + 255) & ~(size_t)255; }
// split-bf16 path: `saved` = log-sum-exp | q planes | k planes | v planes (each [2][B,S,H,64] bf16)
struct BxSaved { float* lse; uint8_t* q; uint8_t* k; uint8_t* v; };
static BxSaved bx_saved(void* saved, int64_t B, int64_t H, int64_t Sq, int64_t Skv) {
    BxSaved r;
    uint8_t* p = reinterpret_cast<uint8_t*>(saved);
    r.lse = reinterpret_cast<float*>(p);
    r.q = p + pad256((size_t)B * H * Sq * sizeof(float));
    r.k = r.q + (size_t)B * Sq * H * 64 * 4;
    r.v = r.k + (size_t)B * Skv * H * 64 * 4;
    return r;
}

static npm_gemm_desc blank_desc() {
    npm_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.nb1 = d.nb2 = 1;
    d.alpha = 1.0f;
    d.precision = -1;
    return d;
}

}  // namespace npm

using namespace npm;

extern "C" {

const char* npm_last_error(void) { return g_err; }
int npm_version(void) { return 100; }
uint64_t npm_launch_count(void) { return g_launches.load(); }
void npm_reset_launch_count(void) { g_launches.store(0); }
int npm_set_precision(int precision) {
    if (precision < NPM_PREC_TF32 || precision > NPM_PREC_BF16) return g_precision.load();
    return g_precision.exchange(precision);
}
int npm_get_precision(void) { return g_precision.load(); }
int npm_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0, a = 0, b = 0, c = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
        set_error("device query failed: %s", cudaGetErrorString(cudaGetLastError()));
        return NPM_ERR_CUDA;
    }
    if (sm_count) *sm_count = a;
    if (cc_major) *cc_major = b;
    if (cc_minor) *cc_minor = c;
    return NPM_OK;
}

int npm_gemm(const npm_gemm_desc* d, npm_stream_t stream) {
    NPM_REQUIRE(d != nullptr, "gemm: NULL descriptor");
    return gemm_dispatch(*d, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ Linear
int npm_linear_fwd(const float* x, const float* w, const float* b, float* y, int64_t m, int64_t k, int64_t n,
                   int w_out_major, int relu, npm_stream_t stream) {
    npm_gemm_desc d = blank_desc();
    d.a = x; d.b = w; d.c = y; d.bias = b;
    d.m = m; d.n = n; d.k = k;
    d.a_rs = k; d.a_cs = 1;
    if (w_out_major) { d.b_rs = 1; d.b_cs = k; }   // W[n,k]: B(k,n) at n*k + k
    else             { d.b_rs = n; d.b_cs = 1; }   // W[k,n]
    d.ldc = n;
    d.flags = relu ? NPM_GEMM_RELU : 0;
    return gemm_dispatch(d, (cudaStream_t)stream);
}

int npm_linear_fwd_residual(const float* x, const float* w, const float* b, const float* residual, float* y, int64_t m,
                            int64_t k, int64_t n, int w_out_major, npm_stream_t stream) {
    npm_gemm_desc d = blank_desc();
    d.a = x; d.b = w; d.c = y; d.bias = b;
    d.m = m; d.n = n; d.k = k;
    d.a_rs = k; d.a_cs = 1;
    if (w_out_major) { d.b_rs = 1; d.b_cs = k; } else { d.b_rs = n; d.b_cs = 1; }
    d.ldc = n;
    d.residual = residual; d.ldr = n;
    return gemm_dispatch(d, (cudaStream_t)stream);
}

size_t npm_weight_split_bytes(int64_t rows, int64_t cols) { return (size_t)rows * cols * 4; }
int npm_weight_split(const float* w, void* planes, int64_t rows, int64_t cols, npm_stream_t stream) {
    NPM_REQUIRE(w && planes && rows > 0 && cols > 0 && (rows * cols) % 4 == 0 && rows * cols < (1ll << 32),
                "weight_split: needs a multiple of 4 (and fewer than 2^32) elements");
    return attn_split_launch(w, rows * cols, planes, 1, rows * cols, (cudaStream_t)stream);     // one "row" of rows * cols elements
}
int npm_linear_fwd_presplit(const float* x, const float* w, const void* w_planes, int64_t plane, const float* b,
                            const float* residual, float* y, int64_t m, int64_t k, int64_t n, int w_out_major, int relu,
                            npm_stream_t stream) {
    npm_gemm_desc d = blank_desc();
    d.a = x; d.b = w; d.c = y; d.bias = b;
    d.m = m; d.n = n; d.k = k;
    d.a_rs = k; d.a_cs = 1;
    if (w_out_major) { d.b_rs = 1; d.b_cs = k; } else { d.b_rs = n; d.b_cs = 1; }
    d.ldc = n;
    d.flags = relu ? NPM_GEMM_RELU : 0;
    d.residual = residual; d.ldr = n;
    d.b_split = w_planes; d.b_split_plane = plane;
    return gemm_dispatch(d, (cudaStream_t)stream);
}
int npm_linear_fwd_planes(const float* x, const float* w, const void* w_planes, int64_t plane, const float* b, void* y_planes,
                          int64_t y_plane, int64_t m, int64_t k, int64_t n, int w_out_major, npm_stream_t stream) {
    NPM_REQUIRE(y_planes != nullptr, "linear_fwd_planes: NULL output");
    if (n % 8 != 0 || y_plane % 8 != 0) { set_error("linear_fwd_planes: n and the plane stride must be multiples of 8"); return NPM_ERR_UNSUPPORTED; }
    npm_gemm_desc d = blank_desc();
    d.a = x; d.b = w; d.c = nullptr; d.bias = b;
    d.m = m; d.n = n; d.k = k;
    d.a_rs = k; d.a_cs = 1;
    if (w_out_major) { d.b_rs = 1; d.b_cs = k; } else { d.b_rs = n; d.b_cs = 1; }
    d.ldc = n;
    d.b_split = w_planes; d.b_split_plane = plane;
    d.c_split = y_planes; d.c_split_plane = y_plane;
    return gemm_dispatch(d, (cudaStream_t)stream);
}
int npm_linear_bwd_dx_presplit(const float* dy, const float* w, const void* w_planes, int64_t plane, float* dx, int64_t m,
                               int64_t k, int64_t n, int w_out_major, npm_stream_t stream) {
    npm_gemm_desc d = blank_desc();
    d.a = dy; d.b = w; d.c = dx;
    d.m = m; d.n = k; d.k = n;
    d.a_rs = n; d.a_cs = 1;
    if (w_out_major) { d.b_rs = k; d.b_cs = 1; } else { d.b_rs = 1; d.b_cs = n; }
    d.ldc = k;
    d.b_split = w_planes; d.b_split_plane = plane;
    return gemm_dispatch(d, (cudaStream_t)stream);
}

int npm_linear_bwd_dx_planes_rowdot(const float* dy, const float* w, const void* w_planes, int64_t plane, void* dx_planes,
                                    int64_t dx_plane, int64_t m, int64_t k, int64_t n, int w_out_major, const float* o, int64_t ldo,
                                    float* rowdot_out, int64_t seq, npm_stream_t stream) {
    NPM_REQUIRE(dy && w && dx_planes && o && rowdot_out, "linear_bwd_dx_planes_rowdot: NULL pointer");
    if (k % 64 != 0 || dx_plane % 8 != 0 || seq <= 0 || m % seq != 0) {
        set_error("linear_bwd_dx_planes_rowdot: k must be a multiple of 64, m of seq, the plane stride of 8");
        return NPM_ERR_UNSUPPORTED;
    }
    npm_gemm_desc d = blank_desc();
    d.a = dy; d.b = w; d.c = nullptr;
    d.m = m; d.n = k; d.k = n;
    d.a_rs = n; d.a_cs = 1;
    if (w_out_major) { d.b_rs = k; d.b_cs = 1; } else { d.b_rs = 1; d.b_cs = n; }
    d.ldc = k;
    d.b_split = w_planes; d.b_split_plane = plane;
    d.c_split = dx_planes; d.c_split_plane = dx_plane;
    d.rowdot_x = o; d.rowdot_ld = ldo; d.rowdot_out = rowdot_out; d.rowdot_seq = seq;
    return gemm_dispatch(d, (cudaStream_t)stream);
}

int npm_linear_bwd_dx(const float* dy, const float* w, float* dx, int64_t m, int64_t k, int64_t n, int w_out_major,
                      npm_stream_t stream) {
    // dx[m,k] = sum_n dy[m,n] * W(k,n): contraction over n
    npm_gemm_desc d = blank_desc();
    d.a = dy; d.b = w; d.c = dx;
    d.m = m; d.n = k; d.k = n;
    d.a_rs = n; d.a_cs = 1;
    if (w_out_major) { d.b_rs = k; d.b_cs = 1; }   // B(kk=n, nn=k) = W[n,k] at n*k + k
    else             { d.b_rs = 1; d.b_cs = n; }   // B(kk=n, nn=k) = W[k,n] at k*n + n
    d.ldc = k;
    return gemm_dispatch(d, (cudaStream_t)stream);
}

int npm_linear_bwd_dw_db(const float* x, const float* dy, float* dw, float* db, int64_t m, int64_t k, int64_t n,
                         int w_out_major, void* workspace, npm_stream_t stream) {
    npm_gemm_desc d = blank_desc();
    if (!w_out_major) {
        // dw[k,n] = sum_m x[m,k] * dy[m,n]
        d.a = x; d.b = dy; d.c = dw;
        d.m = k; d.n = n; d.k = m;
        d.a_rs = 1; d.a_cs = k;    // A(kk, mm) = x[mm, kk]
        d.b_rs = n; d.b_cs = 1;
        d.ldc = n;
    } else {
        // dw[n,k] = sum_m dy[m,n] * x[m,k]
        d.a = dy; d.b = x; d.c = dw;
        d.m = n; d.n = k; d.k = m;
        d.a_rs = 1; d.a_cs = n;
        d.b_rs = k; d.b_cs = 1;
        d.ldc = k;
    }
    // split-bf16 mode, output-major weights: dy is the MN-major A operand of this GEMM, and its column sums come out
    // of the operand tiles the kernel converts anyway (gemm_bx.cu) — no separate pass over dy
    static const bool fold_off = getenv("NPM_NO_COLSUM_FOLD") != nullptr || getenv("NPM_GEMM_NO_BX") != nullptr;      // A/B switch
    const int prec = g_precision.load();
    const bool fold = db != nullptr && w_out_major && !fold_off && (prec == NPM_PREC_BF16X3 || prec == NPM_PREC_BF16) && gemm_bx_supported(d);
    if (fold) d.a_colsum = db;
    int rc = gemm_dispatch(d, (cudaStream_t)stream);
    if (rc || db == nullptr || fold) return rc;
    return colsum_launch(dy, db, m, n, workspace, (cudaStream_t)stream);
}

// ---------------------------------------------------------- attention core
// Two implementations behind one ABI.  TF32 mode with dk = dv = 64: the fused online-softmax kernels of attn_fwd.cu /
// attn_bwd.cu (`saved` = one log-sum-exp per row).  Otherwise (3xTF32 / fp32 modes, other head dims): the batched
// products run on the GEMM with the scores materialised ([B,H,Sq,Skv], as the reference does at attentions.py:103-111)
// and `saved` holds the probabilities P.
int npm_mha_core_path(int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv) {
    return attn_path_now(B, H, Sq, Skv, dk, dv);
}
size_t npm_mha_core_saved_bytes_for(int path, int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv) {
    if (path == ATTN_FUSED_TF32) return (size_t)B * H * Sq * sizeof(float);             // log-sum-exp per row
    if (path == ATTN_FUSED_BX) return pad256((size_t)B * H * Sq * sizeof(float)) + (size_t)B * (Sq + 2 * Skv) * H * 64 * 4;
    return (size_t)B * H * Sq * Skv * sizeof(float);                                      // P
}
size_t npm_mha_core_bwd_scratch_bytes_for(int path, int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv) {
    if (path != ATTN_MATERIALISED) return attn_bwd_scratch_bytes(B, H, Sq, path == ATTN_FUSED_BX);   // D (+ dO planes)
    return (size_t)B * H * Sq * Skv * sizeof(float);   // dP / dS
}
size_t npm_mha_core_saved_bytes(int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv) {
    return npm_mha_core_saved_bytes_for(attn_path_now(B, H, Sq, Skv, dk, dv), B, H, Sq, Skv, dk, dv);
}
size_t npm_mha_core_bwd_scratch_bytes(int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv) {
    return npm_mha_core_bwd_scratch_bytes_for(attn_path_now(B, H, Sq, Skv, dk, dv), B, H, Sq, Skv, dk, dv);
}

int npm_mha_core_fwd(const float* q, const float* k, const float* v, float* o, void* saved, int64_t B, int64_t H,
                     int64_t Sq, int64_t Skv, int64_t dk, int64_t dv, npm_stream_t stream) {
    return npm_mha_core_fwd_strided(q, k, v, o, saved, B, H, Sq, Skv, dk, dv, nullptr, stream);
}

int npm_mha_core_fwd_strided(const float* q, const float* k, const float* v, float* o, void* saved, int64_t B,
                             int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv, const npm_mha_strides* ld,
                             npm_stream_t stream) {
    NPM_REQUIRE(q && k && v && o && saved, "mha_core_fwd: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    float* P = reinterpret_cast<float*>(saved);
    const int64_t ldq = ld && ld->q ? ld->q : H * dk, ldk = ld && ld->k ? ld->k : H * dk,
                  ldv = ld && ld->v ? ld->v : H * dv;
    NPM_REQUIRE(ldq >= H * dk && ldk >= H * dk && ldv >= H * dv, "mha_core_fwd: token strides must be >= H*d");
    const int causal = ld && ld->causal ? 1 : 0;
    NPM_REQUIRE(!causal || Sq == Skv, "mha_core_fwd: the causal mask needs Sq == Skv");
    const int path = attn_path_of(ld, B, H, Sq, Skv, dk, dv);
    NPM_REQUIRE(path == ATTN_MATERIALISED || attn_fused_supported(B, H, Sq, Skv, dk, dv), "mha_core_fwd: the pinned path does not serve this shape");
    NPM_REQUIRE(!(ld && ld->planes) || path == ATTN_FUSED_BX, "mha_core_fwd: pre-split q / k / v need the split-bf16 path (strides.path = 3)");
    if (path == ATTN_FUSED_TF32) return attn_fwd_launch(q, k, v, o, P, B, H, Sq, Skv, ldq, ldk, ldv, 0, 0, 0, causal, false, 3, s);
    if (path == ATTN_FUSED_BX && ld && ld->planes)      // q / k / v already are bf16 planes (npm_linear_fwd_planes); saved = the log-sum-exp
        return attn_fwd_launch(q, k, v, o, P, B, H, Sq, Skv, ldq, ldk, ldv, ld->q_plane, ld->k_plane, ld->v_plane, causal, true, attn_terms(), s);
    if (path == ATTN_FUSED_BX) {
        const BxSaved sv = bx_saved(saved, B, H, Sq, Skv);
        int rc;
        if ((rc = attn_split_launch(q, ldq, sv.q, B * Sq, H * dk, s))) return rc;
        if ((rc = attn_split_launch(k, ldk, sv.k, B * Skv, H * dk, s))) return rc;
        if ((rc = attn_split_launch(v, ldv, sv.v, B * Skv, H * dv, s))) return rc;
        return attn_fwd_launch(sv.q, sv.k, sv.v, o, sv.lse, B, H, Sq, Skv, H * dk, H * dk, H * dv, B * Sq * H * dk, B * Skv * H * dk,
                               B * Skv * H * dv, causal, true, attn_terms(), s);
    }
    // S[b,h] = (1/sqrt(dk)) q[b,:,h,:] k[b,:,h,:]^T
    npm_gemm_desc d = blank_desc();
    d.a = q; d.b = k; d.c = P;
    d.m = Sq; d.n = Skv; d.k = dk;
    d.a_rs = ldq; d.a_cs = 1;
    d.b_rs = 1; d.b_cs = ldk;             // B(c, t) = k[t, h, c]
    d.ldc = Skv;
    d.nb1 = (int)H; d.nb2 = (int)B;
    d.a_bs1 = dk; d.a_bs2 = Sq * ldq;
    d.b_bs1 = dk; d.b_bs2 = Skv * ldk;
    d.c_bs1 = Sq * Skv; d.c_bs2 = H * Sq * Skv;
    d.alpha = (float)(1.0 / sqrt((double)dk));
    int rc = gemm_dispatch(d, s);
    if (rc) return rc;
    if (causal) {
        causal_mask_kernel<<<bw_grid((size_t)(B * H * Sq * Skv), 256), 256, 0, s>>>(P, B * H * Sq, Sq, Skv);
        count_launch();
        if ((rc = check_launch("causal_mask_kernel"))) return rc;
    }
    rc = npm_softmax_fwd(P, P, B * H * Sq, Skv, stream);
    if (rc) return rc;
    // o[b,:,h,:] = P[b,h] v[b,:,h,:]
    npm_gemm_desc e = blank_desc();
    e.a = P; e.b = v; e.c = o;
    e.m = Sq; e.n = dv; e.k = Skv;
    e.a_rs = Skv; e.a_cs = 1;
    e.b_rs = ldv; e.b_cs = 1;             // B(t, c) = v[t, h, c]
    e.ldc = H * dv;
    e.nb1 = (int)H; e.nb2 = (int)B;
    e.a_bs1 = Sq * Skv; e.a_bs2 = H * Sq * Skv;
    e.b_bs1 = dv; e.b_bs2 = Skv * ldv;
    e.c_bs1 = dv; e.c_bs2 = Sq * H * dv;
    return gemm_dispatch(e, s);
}

int npm_mha_core_bwd(const float* q, const float* k, const float* v, const float* o, const float* d_o,
                     const void* saved, float* dq, float* dk_out, float* dv_out, void* scratch, int64_t B, int64_t H,
                     int64_t Sq, int64_t Skv, int64_t dk, int64_t dv, npm_stream_t stream) {
    return npm_mha_core_bwd_strided(q, k, v, o, d_o, saved, dq, dk_out, dv_out, scratch, B, H, Sq, Skv, dk, dv, nullptr,
                                    stream);
}

int npm_mha_core_bwd_strided(const float* q, const float* k, const float* v, const float* o, const float* d_o,
                             const void* saved, float* dq, float* dk_out, float* dv_out, void* scratch, int64_t B,
                             int64_t H, int64_t Sq, int64_t Skv, int64_t dk, int64_t dv, const npm_mha_strides* ld,
                             npm_stream_t stream) {
    const bool do_ready = ld && ld->do_ready;
    NPM_REQUIRE(q && k && v && (d_o || do_ready) && saved && dq && dk_out && dv_out && scratch, "mha_core_bwd: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ldq = ld && ld->q ? ld->q : H * dk, ldk = ld && ld->k ? ld->k : H * dk,
                  ldv = ld && ld->v ? ld->v : H * dv, lddq = ld && ld->dq ? ld->dq : H * dk,
                  lddk = ld && ld->dk ? ld->dk : H * dk, lddv = ld && ld->dv ? ld->dv : H * dv;
    NPM_REQUIRE(ldq >= H * dk && ldk >= H * dk && ldv >= H * dv && lddq >= H * dk && lddk >= H * dk && lddv >= H * dv,
                "mha_core_bwd: token strides must be >= H*d");
    const int path = attn_path_of(ld, B, H, Sq, Skv, dk, dv);
    NPM_REQUIRE(path == ATTN_MATERIALISED || attn_fused_supported(B, H, Sq, Skv, dk, dv), "mha_core_bwd: the pinned path does not serve this shape");
    NPM_REQUIRE(!(ld && ld->planes) || path == ATTN_FUSED_BX, "mha_core_bwd: pre-split q / k / v need the split-bf16 path (strides.path = 3)");
    NPM_REQUIRE(!do_ready || path == ATTN_FUSED_BX, "mha_core_bwd: do_ready needs the split-bf16 path (strides.path = 3)");
    if (path == ATTN_FUSED_TF32)
        return attn_bwd_launch(q, k, v, o, d_o, reinterpret_cast<const float*>(saved), dq, dk_out, dv_out,
                               reinterpret_cast<float*>(scratch), B, H, Sq, Skv, ldq, ldk, ldv, 0, 0, 0, lddq, lddk, lddv,
                               ld && ld->causal ? 1 : 0, false, 3, false, s);
    if (path == ATTN_FUSED_BX && ld && ld->planes)
        return attn_bwd_launch(q, k, v, o, d_o, reinterpret_cast<const float*>(saved), dq, dk_out, dv_out,
                               reinterpret_cast<float*>(scratch), B, H, Sq, Skv, ldq, ldk, ldv, ld->q_plane, ld->k_plane, ld->v_plane,
                               lddq, lddk, lddv, ld->causal ? 1 : 0, true, attn_terms(), do_ready, s);
    if (path == ATTN_FUSED_BX) {
        const BxSaved sv = bx_saved(const_cast<void*>(saved), B, H, Sq, Skv);
        return attn_bwd_launch(sv.q, sv.k, sv.v, o, d_o, sv.lse, dq, dk_out, dv_out, reinterpret_cast<float*>(scratch), B, H,
                               Sq, Skv, H * dk, H * dk, H * dv, B * Sq * H * dk, B * Skv * H * dk, B * Skv * H * dv, lddq, lddk, lddv,
                               ld && ld->causal ? 1 : 0, true, attn_terms(), do_ready, s);
    }
    const float* P = reinterpret_cast<const float*>(saved);
    float* dP = reinterpret_cast<float*>(scratch);
    int rc;
    {   // dV[b,:,h,:] = P[b,h]^T dO[b,:,h,:]                      attentions.py:147-148
        npm_gemm_desc d = blank_desc();
        d.a = P; d.b = d_o; d.c = dv_out;
        d.m = Skv; d.n = dv; d.k = Sq;
        d.a_rs = 1; d.a_cs = Skv;            // A(t, s) = P[s, t]
        d.b_rs = H * dv; d.b_cs = 1;         // B(s, c) = dO[s, h, c]
        d.ldc = lddv;
        d.nb1 = (int)H; d.nb2 = (int)B;
        d.a_bs1 = Sq * Skv; d.a_bs2 = H * Sq * Skv;
        d.b_bs1 = dv; d.b_bs2 = Sq * H * dv;
        d.c_bs1 = dv; d.c_bs2 = Skv * lddv;
        if ((rc = gemm_dispatch(d, s))) return rc;
    }
    {   // dP[b,h] = dO[b,:,h,:] v[b,:,h,:]^T                      attentions.py:146
        npm_gemm_desc d = blank_desc();
        d.a = d_o; d.b = v; d.c = dP;
        d.m = Sq; d.n = Skv; d.k = dv;
        d.a_rs = H * dv; d.a_cs = 1;
        d.b_rs = 1; d.b_cs = ldv;            // B(c, t) = v[t, h, c]
        d.ldc = Skv;
        d.nb1 = (int)H; d.nb2 = (int)B;
        d.a_bs1 = dv; d.a_bs2 = Sq * H * dv;
        d.b_bs1 = dv; d.b_bs2 = Skv * ldv;
        d.c_bs1 = Sq * Skv; d.c_bs2 = H * Sq * Skv;
        if ((rc = gemm_dispatch(d, s))) return rc;
    }
    // dS = P * (dP - rowsum(dP * P)) / sqrt(dk)                    attentions.py:150-155
    if ((rc = npm_softmax_bwd(P, dP, dP, B * H * Sq, Skv, (float)(1.0 / sqrt((double)dk)), stream))) return rc;
    {   // dQ[b,:,h,:] = dS[b,h] k[b,:,h,:]                         attentions.py:161
        npm_gemm_desc d = blank_desc();
        d.a = dP; d.b = k; d.c = dq;
        d.m = Sq; d.n = dk; d.k = Skv;
        d.a_rs = Skv; d.a_cs = 1;
        d.b_rs = ldk; d.b_cs = 1;            // B(t, c) = k[t, h, c]
        d.ldc = lddq;
        d.nb1 = (int)H; d.nb2 = (int)B;
        d.a_bs1 = Sq * Skv; d.a_bs2 = H * Sq * Skv;
        d.b_bs1 = dk; d.b_bs2 = Skv * ldk;
        d.c_bs1 = dk; d.c_bs2 = Sq * lddq;
        if ((rc = gemm_dispatch(d, s))) return rc;
    }
    {   // dK[b,:,h,:] = dS[b,h]^T q[b,:,h,:]                       attentions.py:162
        npm_gemm_desc d = blank_desc();
        d.a = dP; d.b = q; d.c = dk_out;
        d.m = Skv; d.n = dk; d.k = Sq;
        d.a_rs = 1; d.a_cs = Skv;            // A(t, s) = dS[s, t]
        d.b_rs = ldq; d.b_cs = 1;            // B(s, c) = q[s, h, c]
        d.ldc = lddk;
        d.nb1 = (int)H; d.nb2 = (int)B;
        d.a_bs1 = Sq * Skv; d.a_bs2 = H * Sq * Skv;
        d.b_bs1 = dk; d.b_bs2 = Sq * ldq;
        d.c_bs1 = dk; d.c_bs2 = Skv * lddk;
        if ((rc = gemm_dispatch(d, s))) return rc;
    }
    return NPM_OK;
}

int npm_mha_core_scores(const float* q, const float* k, const void* saved, float* p_out, int64_t B, int64_t H,
                        int64_t Sq, int64_t Skv, int64_t dk, int64_t dv, npm_stream_t stream) {
    NPM_REQUIRE(saved && p_out, "mha_core_scores: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (attn_path_now(B, H, Sq, Skv, dk, dv) != ATTN_MATERIALISED) {
        // recompute P = exp(q k^T / sqrt(dk) - lse) from the saved log-sum-exp (first thing in `saved` on both fused paths)
        NPM_REQUIRE(q && k, "mha_core_scores: q and k are needed to recompute the probabilities");
        npm_gemm_desc d = blank_desc();
        d.a = q; d.b = k; d.c = p_out;
        d.m = Sq; d.n = Skv; d.k = dk;
        d.a_rs = H * dk; d.a_cs = 1;
        d.b_rs = 1; d.b_cs = H * dk;
        d.ldc = Skv;
        d.nb1 = (int)H; d.nb2 = (int)B;
        d.a_bs1 = dk; d.a_bs2 = Sq * H * dk;
        d.b_bs1 = dk; d.b_bs2 = Skv * H * dk;
        d.c_bs1 = Sq * Skv; d.c_bs2 = H * Sq * Skv;
        d.alpha = (float)(1.0 / sqrt((double)dk));
        int rc = gemm_dispatch(d, s);
        if (rc) return rc;
        return attn_scores_from_lse_launch(p_out, reinterpret_cast<const float*>(saved), B * H * Sq, Skv, s);
    }
    cudaError_t e = cudaMemcpyAsync(p_out, saved, (size_t)B * H * Sq * Skv * sizeof(float), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) { set_error("mha_core_scores: %s", cudaGetErrorString(e)); return NPM_ERR_CUDA; }
    return NPM_OK;
}

}  // extern "C"
