/* npm_b200.h — C-ABI of libnpm_b200.so: the B200 (sm_100a) replacement for the
 * NumPy calls on np-modeling's layer forward/backward hot path.
 *
 * The reference (levendlee/np-modeling) has no FFI layer: its "operator API" is
 * the Python Layer protocol (layers/layer.py:27-45) whose primitives call NumPy
 * directly.  Each entry point below replaces the NumPy expression(s) at the
 * cited reference file:line; the Python mirror in np-modeling_b200/layers/ package
 * keeps the reference's class / attribute names and calls these through ctypes
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Contract (all functions):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer to fp32
 *     (row-major, densely packed unless a stride is passed) unless the name
 *     says `host`;
 *   - asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     allocates, never synchronises; workspaces are caller-provided, sized by
 *     the matching *_workspace() query;
 *   - returns NPM_OK (0) or a negative code; npm_last_error() returns a
 *     thread-local human-readable message for the last failure;
 *   - no CPU fallback: without a usable sm_100 device every compute call fails.
 */
#ifndef NPM_B200_H_
#define NPM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPM_OK               0
#define NPM_ERR_INVALID     (-1)  /* bad argument / shape / alignment            */
#define NPM_ERR_CUDA        (-2)  /* CUDA runtime or driver error                */
#define NPM_ERR_UNSUPPORTED (-3)  /* valid request this build cannot serve       */

/* Contraction precision of the tensor-core paths (GEMM / conv / attention). */
#define NPM_PREC_TF32   0   /* one tcgen05 kind::tf32 pass                      */
#define NPM_PREC_3XTF32 1   /* hi/lo split, three passes, ~fp32 accuracy        */
#define NPM_PREC_FP32   2   /* CUDA-core fp32 FMA (exact-order SIMT kernel)     */
#define NPM_PREC_BF16X3 3   /* fp32 operands split into bf16 hi + mid on the way into shared memory,
                             * three tcgen05 kind::f16 products (mid*hi + hi*mid + hi*hi), fp32
                             * accumulate: ~2^-17 relative per product, 1.5x the time of one TF32 pass.
                             * Meets rtol 1e-3 / atol 1e-4; the mode bench.py reports.            */
#define NPM_PREC_BF16   4   /* bf16 hi*hi only (bf16 compute on fp32 storage, fp32 accumulate):
                             * SURVEY.md §8 f3; stated tolerance 2e-2 of each tensor's magnitude  */

typedef void* npm_stream_t; /* cudaStream_t */

/* ---- library state -------------------------------------------------------- */
const char* npm_last_error(void);
int         npm_version(void);
/* Number of kernels this library launched since the last reset (bench.py's
 * gpu_launches claim is read from here). */
uint64_t    npm_launch_count(void);
void        npm_reset_launch_count(void);
/* Default contraction precision used when a call passes precision < 0.
 * Returns the previous value. */
int         npm_set_precision(int precision);
int         npm_get_precision(void);
/* Device properties of the current device as this library sees them. */
int         npm_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- general strided-batched GEMM ---------------------------------------- */
/* C[z][m,n] (=|+=) alpha * sum_k A[z][m,k] * B[z][k,n] (+ bias[n]) (then ReLU)
 * A element (m,k) lives at a + z1*a_bs1 + z2*a_bs2 + m*a_rs + k*a_cs, with
 * exactly one of a_rs/a_cs equal to 1; same for B element (k,n) with b_rs/b_cs.
 * C is row-major with leading dimension ldc.  z = z2*nb1 + z1.
 * Replaces np.matmul / np.einsum throughout the reference hot path. */
#define NPM_GEMM_RELU   1   /* apply max(.,0) after bias                        */
#define NPM_GEMM_ACCUM  2   /* C += result instead of C = result                */
typedef struct npm_gemm_desc {
    const float* a;
    const float* b;
    float*       c;
    const float* bias;        /* length n, or NULL                              */
    int64_t m, n, k;
    int64_t a_rs, a_cs;
    int64_t b_rs, b_cs;
    int64_t ldc;
    int32_t nb1, nb2;
    int64_t a_bs1, a_bs2, b_bs1, b_bs2, c_bs1, c_bs2;
    float   alpha;
    int32_t flags;
    int32_t precision;        /* NPM_PREC_* or <0 for the library default       */
    const float* residual;    /* NULL, or [m, n] (leading dimension ldr) added after bias: the
                               * `out += skip` of layers/transformer.py:39,53 in the epilogue;
                               * unbatched problems only                            */
    int64_t ldr;
    float*  a_colsum;         /* NULL, or [m]: receives sum_k A[m,k] (overwritten).  For the dW GEMM of a
                               * projection (dw = dy^T x, A = dy^T) this is the bias gradient
                               * np.sum(dy, axis=0) of attentions.py:129-135 / mlp.py:34, taken from the
                               * operand tiles the GEMM stages anyway.  Served by the split-bf16 kernel
                               * (precision NPM_PREC_BF16X3 / BF16, A m-contiguous, unbatched, m > 128);
                               * otherwise NPM_ERR_UNSUPPORTED — npm_linear_bwd_dw_db chooses.      */
    const void* b_split;      /* NULL, or the bf16 hi / mid planes of B written by npm_weight_split
                               * (same element (k,n) addressing as `b`, in bf16; the mid plane starts
                               * b_split_plane elements after the hi plane).  A hint: the split-bf16
                               * kernel then lands B without converting it; every other path ignores it
                               * and reads `b`.                                                       */
    int64_t b_split_plane;
    const void* a_split;      /* the same hint for A: an activation that exists as bf16 planes because its
                               * producer wrote it that way (c_split of the GEMM before it, the split
                               * output of npm_relu_bwd_colsum_planes); `a` may then be NULL.            */
    int64_t a_split_plane;
    void*   c_split;          /* NULL, or bf16 planes [2][m, ldc] (mid plane c_split_plane elements after
                               * the hi plane): the result is written ONLY in this split form (c may be
                               * NULL) — what the fused split-bf16 attention reads, so a q | k | v
                               * projection feeds it without an fp32 round trip.  Split-bf16 kernel only
                               * (else NPM_ERR_UNSUPPORTED); unbatched, no ACCUM / residual.          */
    int64_t c_split_plane;
    const float* rowdot_x;    /* NULL, or [m, n] fp32 (leading dimension rowdot_ld): together with c_split, the epilogue
                               * also writes rowdot_out[(row / rowdot_seq) * (n / 64) + j][row % rowdot_seq] =
                               * sum over the 64 columns of column group j of C[row, .] * rowdot_x[row, .] (the fp32
                               * result, before the split).  For the dX GEMM of the attention output projection
                               * (C = dO, rowdot_x = O, rowdot_seq = Sq) that is D = rowsum(dO o O) per head, the row
                               * term of Softmax.backward (activations.py:42-45 in closed form) in the [B, H, Sq]
                               * layout npm_mha_core_bwd reads — no separate pass over dO and O.  n % 64 == 0.        */
    int64_t rowdot_ld;
    float*  rowdot_out;
    int64_t rowdot_seq;
} npm_gemm_desc;
int npm_gemm(const npm_gemm_desc* d, npm_stream_t stream);

/* ---- Linear / Dense (layers/mlp.py) --------------------------------------- */
/* y[m,n] = x[m,k] @ W + b.  w_out_major = 0: W is [k,n] (Linear._w, mlp.py:18);
 * w_out_major = 1: W is [n,k] (the MultiHeadAttention projection weights viewed
 * as [H*dk, D], attentions.py:46-57).  relu != 0 fuses Dense's ReLU
 * (mlp.py:70-72) — the pre-activation is then NOT written.   mlp.py:21-25 */
int npm_linear_fwd(const float* x, const float* w, const float* b, float* y,
                   int64_t m, int64_t k, int64_t n, int w_out_major, int relu,
                   npm_stream_t stream);
/* y = x @ W + b + residual[m,n]: the residual connection that follows the
 * output projection / second FFN layer in layers/transformer.py:39,53,129,143,
 * 155, fused into the GEMM epilogue (residual may alias nothing it writes). */
int npm_linear_fwd_residual(const float* x, const float* w, const float* b,
                            const float* residual, float* y, int64_t m,
                            int64_t k, int64_t n, int w_out_major,
                            npm_stream_t stream);
/* Split-bf16 planes of a weight matrix for the NPM_PREC_BF16X3 mode: hi = bf16_rn(w),
 * mid = bf16_rn(w - hi), as two [rows, cols] bf16 matrices (mid plane right after the hi plane:
 * npm_weight_split_bytes = 4 * rows * cols).  A weight is the B operand of its forward GEMM and
 * of its dX GEMM; split once per step it is landed by TMA in tensor-core layout instead of being
 * converted in shared memory by both.  The *_presplit entry points take the planes as a hint
 * (`w_planes` may be NULL; `plane` = elements between the hi and the mid plane, so that a row
 * block of a packed weight can be addressed); they compute exactly what npm_linear_fwd /
 * npm_linear_fwd_residual / npm_linear_bwd_dx compute. */
size_t npm_weight_split_bytes(int64_t rows, int64_t cols);
int npm_weight_split(const float* w, void* planes, int64_t rows, int64_t cols,
                     npm_stream_t stream);
int npm_linear_fwd_presplit(const float* x, const float* w, const void* w_planes,
                            int64_t plane, const float* b, const float* residual,
                            float* y, int64_t m, int64_t k, int64_t n,
                            int w_out_major, int relu, npm_stream_t stream);
int npm_linear_bwd_dx_presplit(const float* dy, const float* w, const void* w_planes,
                               int64_t plane, float* dx, int64_t m, int64_t k,
                               int64_t n, int w_out_major, npm_stream_t stream);
/* y = x @ W + b written ONLY as bf16 hi / mid planes (y_planes: [2][m, n], mid plane y_plane
 * elements after the hi plane): the projections in front of the fused split-bf16 attention
 * (attentions.py:90-100).  NPM_ERR_UNSUPPORTED when the split-bf16 GEMM does not take the
 * problem (m <= 128, n % 8 != 0, other precision modes): the caller then projects to fp32. */
int npm_linear_fwd_planes(const float* x, const float* w, const void* w_planes,
                          int64_t plane, const float* b, void* y_planes, int64_t y_plane,
                          int64_t m, int64_t k, int64_t n, int w_out_major,
                          npm_stream_t stream);
/* dx[m,k] = dy[m,n] @ W^T.                                         mlp.py:36 */
int npm_linear_bwd_dx(const float* dy, const float* w, float* dx,
                      int64_t m, int64_t k, int64_t n, int w_out_major,
                      npm_stream_t stream);
/* dw = x^T @ dy (layout per w_out_major), db[n] = sum_m dy (db may be NULL).
 * workspace: npm_colsum_workspace(m, n) bytes (may be NULL when db is NULL).
 *                                                               mlp.py:34-35 */
/* dx = dy @ W^T written ONLY as bf16 hi / mid planes [2][m, k] (mid plane dx_plane elements after the hi plane), plus the
 * per-64-column row dots with `o` (npm_gemm_desc.rowdot_*): the backward of MultiHeadAttention's output projection
 * (attentions.py:129-136) feeding the fused split-bf16 attention backward, whose `scratch` (D, then the planes) it fills.
 * NPM_ERR_UNSUPPORTED when the split-bf16 GEMM does not take the problem (callers then use npm_linear_bwd_dx). */
int npm_linear_bwd_dx_planes_rowdot(const float* dy, const float* w, const void* w_planes, int64_t plane,
                                    void* dx_planes, int64_t dx_plane, int64_t m, int64_t k, int64_t n,
                                    int w_out_major, const float* o, int64_t ldo, float* rowdot_out,
                                    int64_t seq, npm_stream_t stream);
/* An activation that exists ONLY as split-bf16 planes (NPM_PREC_BF16X3): the FFN hidden activation written by the
 * first FFN GEMM's epilogue (npm_gemm_desc.c_split, ReLU applied: mlp.py:70-72) is the A operand of the second FFN
 * GEMM and of its dW GEMM as is.
 * npm_relu_bwd_colsum_planes: ReLU.backward (activations.py:17-19) + the bias gradient (mlp.py:34) for such a y — the
 *   gate is the sign bit of y's bf16 hi plane `y_hi` ([rows, cols] bf16; bf16 rounding keeps the sign, -0.0 included);
 *   dz = where(gate, dy, 0) is written as bf16 hi / mid planes (mid `plane` elements after hi), db[cols] = column sums
 *   of dz.  workspace: npm_colsum_workspace(rows, cols).  NPM_ERR_UNSUPPORTED unless cols %% 4 == 0 and cols <= 16384.
 * npm_planes_join: out[n] fp32 = hi + mid (exact), for consumers outside the split-bf16 kernels. */
/* LayerNormalization.forward (normalizations.py:43-58), optionally with DropOut.forward (:14-23) in front of it (maskbits
 * != NULL and keep_prob < 1: the same Philox mask and keep bits as npm_dropout_layernorm_fwd), whose result exists only
 * as split-bf16 planes: the normalised activation of a pre-norm transformer block feeds nothing but the first FFN GEMM
 * and its dW GEMM, which take the planes as their operand image.  mean / rstd as npm_layernorm_fwd. */
int npm_layernorm_fwd_planes(const float* x, const float* gamma, const float* beta, void* out_planes, int64_t plane,
                             float* mean, float* rstd, uint32_t* maskbits, int64_t rows, int64_t cols, float epsilon,
                             float keep_prob, uint64_t seed, uint64_t offset, npm_stream_t stream);
int npm_relu_bwd_colsum_planes(const void* y_hi, const float* dy, void* dz_planes, int64_t plane, float* db,
                               int64_t rows, int64_t cols, void* workspace, npm_stream_t stream);
int npm_planes_join(const void* planes, int64_t plane, float* out, int64_t n, npm_stream_t stream);
int npm_linear_bwd_dw_db(const float* x, const float* dy, float* dw, float* db,
                         int64_t m, int64_t k, int64_t n, int w_out_major,
                         void* workspace, npm_stream_t stream);

/* ---- column sum: out[c] = sum_r x[r,c]   (bias / beta gradients) ---------- */
size_t npm_colsum_workspace(int64_t rows, int64_t cols);
int npm_colsum(const float* x, float* out, int64_t rows, int64_t cols,
               void* workspace, npm_stream_t stream);

/* ---- activations (layers/activations.py) ---------------------------------- */
int npm_relu_fwd(const float* x, float* y, int64_t n, npm_stream_t stream);      /* :13-15 */
/* dx = dy where x >= 0 else 0 (note >=, activations.py:19) */
int npm_relu_bwd(const float* x, const float* dy, float* dx, int64_t n,
                 npm_stream_t stream);
/* ReLU.backward from the OUTPUT of a fused Dense forward (npm_linear_fwd with
 * relu != 0): dx = dy where the sign bit of y is clear, else 0. */
int npm_relu_bwd_y(const float* y, const float* dy, float* dx, int64_t n,
                   npm_stream_t stream);
/* The same on a [rows, cols] matrix, fused with the bias gradient that follows
 * it in Dense.backward (mlp.py:74-77 → :34): db[c] = sum_r dx[r,c].
 * workspace: npm_colsum_workspace(rows, cols) bytes. */
int npm_relu_bwd_colsum(const float* y, const float* dy, float* dx, float* db,
                        int64_t rows, int64_t cols, void* workspace,
                        npm_stream_t stream);
/* row softmax over the last axis, max-shifted (activations.py:23-31). In place ok. */
int npm_softmax_fwd(const float* x, float* y, int64_t rows, int64_t cols,
                    npm_stream_t stream);
/* dx = scale * y * (dy - sum(dy*y)) — closed form of the Jacobian einsum at
 * activations.py:33-45. In place on dy ok. */
int npm_softmax_bwd(const float* y, const float* dy, float* dx, int64_t rows,
                    int64_t cols, float scale, npm_stream_t stream);

/* ---- LayerNormalization (layers/normalizations.py:33-75) ------------------ */
/* out = gamma*(x-mean)*rstd + beta; mean/rstd [rows] are saved for backward. */
int npm_layernorm_fwd(const float* x, const float* gamma, const float* beta,
                      float* out, float* mean, float* rstd, int64_t rows,
                      int64_t cols, float epsilon, npm_stream_t stream);
size_t npm_layernorm_bwd_workspace(int64_t rows, int64_t cols);
/* dx (closed form of the [..,C,C] Jacobian, :60-71), dgamma, dbeta (:55-56). */
int npm_layernorm_bwd(const float* dz, const float* x, const float* gamma,
                      const float* mean, const float* rstd, float* dx,
                      float* dgamma, float* dbeta, int64_t rows, int64_t cols,
                      void* workspace, npm_stream_t stream);

/* DropOut -> LayerNormalization fused (the pre-norm blocks of
 * layers/transformer.py:35-37,48-50,125-127,136-138,149-151 call them back to
 * back): out = LN(dropout(x)) without writing the dropped tensor; the mask is
 * the Philox mask of npm_dropout_fwd for (keep_prob, seed, offset).  Backward:
 * dx = dropout_bwd(LN_bwd(dz)) + dskip (dskip may be NULL; it is the residual
 * branch the reference adds right after, transformer.py:176,190,201).
 * Served when npm_dropout_layernorm_fused(rows, cols) != 0 (cols % 4 == 0,
 * cols <= 1024), offset % 4 == 0 and pointers are 16-byte aligned; otherwise
 * NPM_ERR_UNSUPPORTED and the caller runs the two layers separately.
 * workspace: npm_layernorm_bwd_workspace(rows, cols) bytes. */
int npm_dropout_layernorm_fused(int64_t rows, int64_t cols);
/* `maskbits` (npm_dropout_layernorm_mask_bytes): the keep bits, bit-packed by
 * the forward kernel and consumed by the backward one (private layout). */
size_t npm_dropout_layernorm_mask_bytes(int64_t rows, int64_t cols);
int npm_dropout_layernorm_fwd(const float* x, const float* gamma,
                              const float* beta, float* out, float* mean,
                              float* rstd, uint32_t* maskbits, int64_t rows,
                              int64_t cols, float epsilon, float keep_prob,
                              uint64_t seed, uint64_t offset,
                              npm_stream_t stream);
int npm_dropout_layernorm_bwd(const float* dz, const float* x,
                              const float* gamma, const float* mean,
                              const float* rstd, const uint32_t* maskbits,
                              const float* dskip, float* dx, float* dgamma,
                              float* dbeta, int64_t rows, int64_t cols,
                              float keep_prob, void* workspace,
                              npm_stream_t stream);

/* Same, and dx_colsum[cols] = column sums of the dx written (NULL: not computed).
 * dx is the gradient of the residual stream, so its column sum is the bias
 * gradient of the projection / second FFN layer that wrote that stream
 * (attentions.py:129 `dbo`, mlp.py:34 `db`): taken from registers here instead of
 * re-reading dx in a npm_colsum launch. */
int npm_dropout_layernorm_bwd_colsum(const float* dz, const float* x,
                                     const float* gamma, const float* mean,
                                     const float* rstd, const uint32_t* maskbits,
                                     const float* dskip, float* dx, float* dgamma,
                                     float* dbeta, float* dx_colsum, int64_t rows,
                                     int64_t cols, float keep_prob, void* workspace,
                                     npm_stream_t stream);

/* ---- DropOut (layers/normalizations.py:9-30) ------------------------------ */
/* Element i is kept iff philox4x32_10(counter=(offset+i)/4, key=seed)[(offset+i)%4]
 * < floor(keep_prob * 2^32); kept values are scaled by 1/keep_prob.  If
 * ext_mask != NULL it is used instead (1 byte per element, non-zero = keep):
 * the reference's own test injects the layer's mask into its oracle the same
 * way (normalizations_test.py:28). */
int npm_dropout_fwd(const float* x, float* y, int64_t n, float keep_prob,
                    uint64_t seed, uint64_t offset, const uint8_t* ext_mask,
                    npm_stream_t stream);
int npm_dropout_bwd(const float* dy, float* dx, int64_t n, float keep_prob,
                    uint64_t seed, uint64_t offset, const uint8_t* ext_mask,
                    npm_stream_t stream);
int npm_dropout_mask(uint8_t* mask, int64_t n, float keep_prob, uint64_t seed,
                     uint64_t offset, npm_stream_t stream);

/* ---- residual / elementwise glue (layers/transformer.py `out += skip`) ---- */
int npm_add_inplace(float* y, const float* x, int64_t n, npm_stream_t stream);
/* out = a + b + c (c may be NULL): sums MHA's (dquery,dkey,dvalue), :85,:196 */
int npm_add3(const float* a, const float* b, const float* c, float* out,
             int64_t n, npm_stream_t stream);
int npm_scale(float* x, float s, int64_t n, npm_stream_t stream);
int npm_fill(float* x, float v, int64_t n, npm_stream_t stream);

/* ---- token embedding (the reference has none) ------------- */
/* out[i,:] = table[ids[i],:] for n int32 token ids (clamped to [0, vocab)). */
int npm_embedding_fwd(const float* table, const int32_t* ids, float* out,
                      int64_t n, int64_t d, int64_t vocab, npm_stream_t stream);
/* dtable = 0; dtable[ids[i],:] += dy[i,:] (red.global.add: the order in which
 * rows of the same id are summed is not fixed). */
int npm_embedding_bwd(const float* dy, const int32_t* ids, float* dtable,
                      int64_t n, int64_t d, int64_t vocab, npm_stream_t stream);

/* ---- attention core (layers/attentions.py:103-112, :146-162) -------------- */
/* q [B,Sq,H,dk], k [B,Skv,H,dk], v [B,Skv,H,dv]  → o [B,Sq,H,dv]
 * o = softmax(q k^T / sqrt(dk)) v per (b,h); unmasked (the reference's mask
 * path raises, attentions.py:84).  `saved` (npm_mha_core_saved_bytes) must be
 * kept by the caller from fwd to bwd; its content is private to the library
 * (TF32 mode with dk = dv = 64: the fused tcgen05 kernels, `saved` = one
 * log-sum-exp per row; otherwise the probabilities P).  The size queries, fwd
 * and bwd of one attention call must run under the same precision mode. */
/* Implementation the CURRENT precision mode selects for this shape: 0 = batched
 * GEMMs + row softmax with the scores materialised (`saved` = P), 1 = fused TF32
 * kernels (`saved` = log-sum-exp), 2 = fused split-bf16 kernels (`saved` =
 * log-sum-exp + the bf16 hi/mid planes of q, k, v).  Callers that keep state
 * from forward to backward record it and pass 1 + path in npm_mha_strides.path;
 * the *_for size queries take it explicitly. */
int    npm_mha_core_path(int64_t B, int64_t H, int64_t Sq, int64_t Skv,
                         int64_t dk, int64_t dv);
size_t npm_mha_core_saved_bytes_for(int path, int64_t B, int64_t H, int64_t Sq,
                                    int64_t Skv, int64_t dk, int64_t dv);
size_t npm_mha_core_bwd_scratch_bytes_for(int path, int64_t B, int64_t H,
                                          int64_t Sq, int64_t Skv, int64_t dk,
                                          int64_t dv);
size_t npm_mha_core_saved_bytes(int64_t B, int64_t H, int64_t Sq, int64_t Skv,
                                int64_t dk, int64_t dv);
size_t npm_mha_core_bwd_scratch_bytes(int64_t B, int64_t H, int64_t Sq,
                                      int64_t Skv, int64_t dk, int64_t dv);
int npm_mha_core_fwd(const float* q, const float* k, const float* v, float* o,
                     void* saved, int64_t B, int64_t H, int64_t Sq, int64_t Skv,
                     int64_t dk, int64_t dv, npm_stream_t stream);
int npm_mha_core_bwd(const float* q, const float* k, const float* v,
                     const float* o, const float* d_o, const void* saved,
                     float* dq, float* dk_out, float* dv_out, void* scratch,
                     int64_t B, int64_t H, int64_t Sq, int64_t Skv, int64_t dk,
                     int64_t dv, npm_stream_t stream);
/* Strided variants: q / k / v (and dq / dk / dv) may be column blocks of a wider
 * row-major buffer, e.g. one [tokens, 3*H*dk] array written by a single packed
 * q|k|v projection GEMM (the three einsums of attentions.py:90-100 share their
 * input in self-attention).  Each field is the number of floats between
 * consecutive tokens (>= H*d, multiple of 4); 0 = dense (H*d); ld == NULL = all
 * dense.  o and d_o are always dense [B,Sq,H,dv]. */
typedef struct npm_mha_strides {
    int64_t q, k, v;          /* inputs                                         */
    int64_t dq, dk, dv;       /* gradients written by the backward              */
    int64_t causal;           /* != 0: key position t > query position s is masked (needs Sq == Skv).
                               * BEYOND the reference, whose mask argument is unusable (`if mask:` on an
                               * ndarray raises, attentions.py:84; backward NotImplementedError :152-153);
                               * SURVEY.md §8 f1.  The fused kernels skip the blocks above the diagonal. */
    int64_t path;             /* 0: the implementation the current precision mode selects; otherwise
                               * 1 + the value npm_mha_core_path() returned when the forward ran — pins
                               * forward, backward and the layout of `saved` to one implementation even if
                               * the precision mode changes in between.                              */
    int64_t planes;           /* != 0 (path 2 only): q / k / v are NOT fp32 but the hi planes of bf16 hi / mid
                               * planes (npm_linear_fwd_planes), token strides q / k / v in bf16 elements, the
                               * mid plane `planes` elements (per tensor: q_plane, k_plane, v_plane below)
                               * after the hi plane; the core then skips its own operand split.       */
    int64_t q_plane, k_plane, v_plane;
    int64_t do_ready;         /* != 0 (path 2 backward only): `scratch` already holds D [B,H,Sq] and, 256-byte aligned
                               * behind it, dO as bf16 hi / mid planes [2][B*Sq, H*64] — written by
                               * npm_linear_bwd_dx_planes_rowdot, the output projection's dX GEMM; d_o may be NULL
                               * and the core skips its own D / split pass.                                        */
} npm_mha_strides;
int npm_mha_core_fwd_strided(const float* q, const float* k, const float* v,
                             float* o, void* saved, int64_t B, int64_t H,
                             int64_t Sq, int64_t Skv, int64_t dk, int64_t dv,
                             const npm_mha_strides* ld, npm_stream_t stream);
int npm_mha_core_bwd_strided(const float* q, const float* k, const float* v,
                             const float* o, const float* d_o, const void* saved,
                             float* dq, float* dk_out, float* dv_out,
                             void* scratch, int64_t B, int64_t H, int64_t Sq,
                             int64_t Skv, int64_t dk, int64_t dv,
                             const npm_mha_strides* ld, npm_stream_t stream);
/* Writes the attention probabilities [B,H,Sq,Skv] to p_out (debug / parity with
 * MultiHeadAttention._attention_scores, attentions.py:108-111): copied out of
 * `saved` when it holds them, recomputed from q, k and the saved log-sum-exp
 * when the fused kernels ran. */
int npm_mha_core_scores(const float* q, const float* k, const void* saved,
                        float* p_out, int64_t B, int64_t H, int64_t Sq,
                        int64_t Skv, int64_t dk, int64_t dv, npm_stream_t stream);

/* ---- Conv2D (layers/conv.py) ---------------------------------------------- */
/* NHWC activations, HWIO filters, SAME padding, stride 1, odd ksize.
 * y = conv(x, f) + b (then ReLU if relu != 0).                 conv.py:44-48 */
size_t npm_conv2d_workspace(int64_t N, int64_t H, int64_t W, int64_t Cin,
                            int64_t Cout, int ksize);
int npm_conv2d_fwd(const float* x, const float* f, const float* b, float* y,
                   int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                   int ksize, int relu, void* workspace, npm_stream_t stream);
/* dx = conv(dy, flipHW(f)^T_IO).                             conv.py:110-153 */
int npm_conv2d_bwd_dx(const float* dy, const float* f, float* dx, int64_t N,
                      int64_t H, int64_t W, int64_t Cin, int64_t Cout,
                      int ksize, void* workspace, npm_stream_t stream);
/* dw[i,j] = xpad[:,i:,j:,:]^T @ dy ; db = sum_{n,h,w} dy.  conv.py:55,156-194 */
int npm_conv2d_bwd_dw_db(const float* x, const float* dy, float* dw, float* db,
                         int64_t N, int64_t H, int64_t W, int64_t Cin,
                         int64_t Cout, int ksize, void* workspace,
                         npm_stream_t stream);

/* ---- losses (loss.py) ------------------------------------------------------ */
/* loss_out: one device float. MSE: sum((y-t)^2)/n, grad 2(y-t)/n  (:20-29).
 * CE (on probabilities): -sum(t*log y), grad -t/y                  (:32-39). */
int npm_mse_fwd(const float* y, const float* t, float* loss_out, int64_t n,
                npm_stream_t stream);
int npm_mse_bwd(const float* y, const float* t, float* dy, int64_t n,
                npm_stream_t stream);
int npm_ce_fwd(const float* y, const float* t, float* loss_out, int64_t n,
               npm_stream_t stream);
int npm_ce_bwd(const float* y, const float* t, float* dy, int64_t n,
               npm_stream_t stream);

/* ---- optimizers (optimizer.py) — one multi-tensor launch ------------------- */
/* A device-resident table of tensors; entry i updates param[i][0..numel[i]).
 * Work is cut into chunks of NPM_OPT_CHUNK elements; chunk_begin[i] is the number
 * of chunks of entries 0..i-1 (ascending), n_chunks their total. */
#define NPM_OPT_CHUNK 8192
typedef struct npm_tensor_entry {
    float*       param;
    const float* grad;
    float*       m;           /* Adam first moment  (NULL for SGD)              */
    float*       v;           /* Adam second moment (NULL for SGD)              */
    int64_t      numel;
    int64_t      chunk_begin;
    void*        planes;      /* NULL, or the bf16 hi plane of this parameter's split-bf16 image (npm_weight_split layout,
                               * the mid plane `plane_stride` elements further): the update rewrites it from the new value,
                               * so the next forward pass needs no split pass over the weights (NPM_PREC_BF16X3)      */
    int64_t      plane_stride;
} npm_tensor_entry;
/* param -= lr * grad_scale * grad                              optimizer.py:30-33 */
int npm_sgd_multi(const npm_tensor_entry* table_dev, int32_t n_tensors,
                  int64_t n_chunks, float lr, float grad_scale,
                  npm_stream_t stream);
/* g = grad_scale*grad; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 * param -= lr * (m/(1-b1^t)) / sqrt(v/(1-b2^t) + eps)   (eps INSIDE the sqrt)
 *                                                           optimizer.py:50-69 */
int npm_adam_multi(const npm_tensor_entry* table_dev, int32_t n_tensors,
                   int64_t n_chunks, float lr, float beta1, float beta2,
                   float epsilon, int32_t t, float grad_scale,
                   npm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NPM_B200_H_ */
