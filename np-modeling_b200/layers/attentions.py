"""MultiHeadAttention — drop-in for layers/attentions.py (:11-199).

Parameter attributes and layouts are the reference's: `_wq,_wk [H,dk,D]`, `_wv [H,dv,D]`,
`_wo [D,H,dv]`, `_bq,_bk [H,dk]`, `_bv [H,dv]`, `_bo [D]`; `backward` returns the 3-tuple
`(dquery, dkey, dvalue)`.  The projection weights viewed as `[H*dk, D]` / `[D, H*dv]` are already
"output-major" GEMM operands (contraction index contiguous), so every projection and its two
gradients are single tcgen05 GEMMs on the arrays as stored; the attention core runs behind
`npm_mha_core_fwd/bwd`.
"""
import ctypes
import os

import optimizer
from layers import activations, layer
from npm_b200 import device
from npm_b200 import _lib
from npm_b200._lib import C, GemmDesc, MhaStrides


# Opt-in (NPM_DO_ROWDOT=1; tests flip the attribute): the output projection's dX GEMM writes dO as bf16 planes and
# D = rowsum(dO o O) from its epilogue (npm_linear_bwd_dx_planes_rowdot) instead of fp32 dO + the D / split kernel.
# Measured neutral-to-slower inside the cfg5 step (87.4-87.9 vs 87.1-87.7 ms, same box): the K = 1024 GEMM's epilogue
# becomes its bottleneck with the extra 32 x 128-byte row loads of O per chunk.
_NO_DO_ROWDOT = not os.environ.get('NPM_DO_ROWDOT')
_NO_QKV_PLANES = bool(os.environ.get('NPM_NO_QKV_PLANES'))      # A/B switch for tools


def _add3(parts):
    out = device.empty(parts[0].shape)
    C.npm_add3(parts[0].ptr, parts[1].ptr, parts[2].ptr, out.ptr, out.size, device.stream())
    return out


class KVCache:
    """Decode-time cache of one MultiHeadAttention (the reference's `# TODO: support cache`,
    transformer.py:120).  Self-attention: the projected keys / values of every token decoded so far, TIME-major
    `[max_len, B, H*d]` — with the batch folded into the head axis this is the `[S, B*H, d]` layout the attention core
    already takes (token stride B*H*d), so a step attends over the first `length` rows without any copy.
    Cross-attention: the projected memory, computed once."""

    def __init__(self, max_len: int, batch: int, hk: int, hv: int):
        self.k = device.empty((max_len, batch, hk))
        self.v = device.empty((max_len, batch, hv))
        self.max_len = max_len
        self.length = 0
        self.memory_kv = None      # cross-attention: [B*Skv, 2*H*d] projected memory (k | v)


class MultiHeadAttention(layer.StatefulLayer):
    def __init__(self, num_heads: int, *args, causal: bool = False, **kwargs):
        """`causal=True` (keyword-only, BEYOND the reference — SURVEY.md §8 f1): key position t > query position s is
        masked, forward and backward; needs seq_len_q == seq_len_kv.  The reference's own `mask=` argument cannot be
        used (attentions.py:84 raises for arrays, :152-153 has no backward) and keeps raising here."""
        super().__init__(*args, **kwargs)
        self._num_heads = num_heads
        self._causal = bool(causal)
        self._softmax = activations.Softmax()

    def initialize(self, query, key=None, value=None, *args, **kwargs) -> None:
        # query: [batch, seq_len_q,  num_heads * key_dim]
        # key:   [batch, seq_len_kv, num_heads * key_dim]
        # value: [batch, seq_len_kv, num_heads * value_dim]
        if key is None:
            key = query
        if value is None:
            value = key

        assert query.shape[0] == key.shape[0]
        assert query.shape[2] == key.shape[2]
        assert query.shape[0] == value.shape[0]
        assert key.shape[1] == value.shape[1]

        self._seq_len_q = query.shape[1]
        self._seq_len_kv = key.shape[1]

        assert key.shape[2] % self._num_heads == 0
        self._key_dim = key.shape[2] // self._num_heads
        assert value.shape[2] % self._num_heads == 0
        self._value_dim = value.shape[2] // self._num_heads

        h, dk, dv = self._num_heads, self._key_dim, self._value_dim
        self._wq = self._initializer([h, dk, h * dk])
        self._wk = self._initializer([h, dk, h * dk])
        self._wv = self._initializer([h, dv, h * dv])
        self._wo = self._initializer([h * dk, h, dv])
        self._bq = self._initializer([h, dk])
        self._bk = self._initializer([h, dk])
        self._bv = self._initializer([h, dv])
        self._bo = self._initializer([h * dk])

    # ---- packed projection parameters --------------------------------------------------------
    # `_wq,_wk,_wv` ([H,dk,D] each) and `_bq,_bk,_bv` are kept as the three leading-axis slices of one
    # [3,H,dk,D] / [3,H,dk] block, so that the projections that share an input run as ONE GEMM:
    # self-attention q|k|v = x @ [Wq;Wk;Wv]^T (N = 3*H*dk), cross-attention k|v (N = 2*H*dk), and in
    # backward dW / db of the block in one GEMM + one column sum and dx = [dq|dk|dv] @ [Wq;Wk;Wv]
    # (K = 3*H*dk), which is already the sum the transformer blocks take (transformer.py:85,184,196).
    # The attribute names, shapes and values are the reference's; only their placement in HBM is chosen.
    def _packed_params(self):
        """(w_pack [3,H,dk,D], b_pack [3,H,dk]) with `_wq/_wk/_wv` and `_bq/_bk/_bv` as their slices, or
        None when dk != dv.  Re-packs when a parameter was rebound (tests assign NumPy arrays by name,
        layers/utils.py:41-101) or the layer was deep-copied."""
        if self._key_dim != self._value_dim:
            return None
        ws = [self._p(n) for n in ('_wq', '_wk', '_wv')]
        bs = [self._p(n) for n in ('_bq', '_bk', '_bv')]
        packs = getattr(self, '_packs', None)
        if packs is not None:
            wp, bp = packs
            wsz, bsz = ws[0].size * 4, bs[0].size * 4
            if all(w.ptr == wp.ptr + i * wsz and w.size * 4 == wsz for i, w in enumerate(ws)) and \
               all(b.ptr == bp.ptr + i * bsz and b.size * 4 == bsz for i, b in enumerate(bs)):
                return packs
        if len({w.shape for w in ws}) != 1 or len({b.shape for b in bs}) != 1:
            return None
        wp = device.empty((3,) + ws[0].shape)
        bp = device.empty((3,) + bs[0].shape)
        for i, (wn, bn) in enumerate((('_wq', '_bq'), ('_wk', '_bk'), ('_wv', '_bv'))):
            setattr(self, wn, wp[i].copy_from(ws[i]))
            setattr(self, bn, bp[i].copy_from(bs[i]))
        self._packs = (wp, bp)
        return self._packs

    def _project(self, x2d, w, b, n, residual=None, planes=None):
        """y = x @ W^T + b (+ residual) for an output-major W [n, k]; `planes` = (pointer, plane stride) of W's bf16
        hi / mid split (bf16x3 mode) or None."""
        m, k = x2d.shape
        y = device.empty((m, n))
        assert residual is None or residual.size == m * n, 'residual must match the attention output'
        pp, ps = planes if planes is not None else (None, 0)
        C.npm_linear_fwd_presplit(x2d.ptr, w.ptr, pp, ps, b.ptr, residual.ptr if residual is not None else None, y.ptr, m, k, n,
                                  1, 0, device.stream())
        return y

    def _split_params(self, packs, wo, tokens):
        """bf16x3 mode: split the packed projection block and `_wo` once per forward; returns a function
        (weight DeviceArray) -> (pointer, plane stride) that also resolves row blocks of the packed block."""
        self._planes = None
        if tokens <= 128:
            return lambda w: None
        blocks = []
        for w in ([packs[0]] if packs is not None else [self._p('_wq'), self._p('_wk'), self._p('_wv')]) + [wo]:
            t = device.split_weight(w)
            if t is not None:
                blocks.append((w.ptr, w.size, t))
        self._planes = blocks

        def find(w):
            for base, size, t in self._planes or ():
                if base <= w.ptr and w.ptr + 4 * w.size <= base + 4 * size:
                    return (t.data_ptr() + (w.ptr - base) // 2, size)       # bf16: half the byte offset of the fp32 block
            return None
        return find

    def forward(self, query, key=None, value=None, mask=None, _residual=None):
        # _residual (B200 extension): the `out += skip` of the transformer blocks, run in the epilogue of the
        # output projection
        if mask is not None:
            # the reference's `if mask:` raises for arrays (attentions.py:84) and its backward is
            # NotImplementedError (:152-153): attention is unmasked.
            raise ValueError('MultiHeadAttention: mask is not supported (reference attentions.py:84,152)')
        query_in, key_in = query, key
        query = device.asdevice(query)
        key = query if (key is None or key is query_in) else device.asdevice(key)
        value = key if (value is None or value is key_in) else (query if value is query_in else device.asdevice(value))

        self._query, self._key, self._value, self._mask = query, key, value, mask
        batch, sq, dmodel = query.shape
        skv = key.shape[1]
        h, dk, dv = self._num_heads, self._key_dim, self._value_dim
        assert sq == self._seq_len_q and skv == self._seq_len_kv, 'sequence lengths are fixed at init (:142-145)'

        packs = self._packed_params()
        wq, wk, wv, wo = self._p('_wq'), self._p('_wk'), self._p('_wv'), self._p('_wo')
        bq, bk, bv, bo = self._p('_bq'), self._p('_bk'), self._p('_bv'), self._p('_bo')
        hd = h * dk
        query2 = query.reshape(batch * sq, dmodel)
        key2 = key.reshape(batch * skv, key.shape[2])
        pl = self._find_planes = self._split_params(packs, wo, min(batch * sq, batch * skv))

        # the implementation (and with it the layout of `saved`) is chosen HERE, under the precision mode of the forward
        # call, and pinned for the backward call: a set_precision() in between cannot make backward misread `saved`
        self._path = int(C.npm_mha_core_path(batch, h, sq, skv, dk, dv))
        # split-bf16 fused attention (path 2): the projections write q / k / v directly as the bf16 hi / mid planes the
        # attention kernels read (npm_linear_fwd_planes) — no fp32 q / k / v, no separate split pass
        self._planes_mode = (self._path == 2 and dk == dv and hd % 8 == 0 and min(batch * sq, batch * skv) > 128
                             and not _NO_QKV_PLANES)

        def project(x2d, w, b, n):
            """-> (buffer, pointer of q-like column 0, elements per token, elements between hi and mid plane)"""
            if self._planes_mode:
                m, k = x2d.shape
                buf = device.workspace(4 * m * n)                         # [2][m, n] bf16
                pp, ps = pl(w) or (None, 0)
                rc = _lib.load().npm_linear_fwd_planes(x2d.ptr, w.ptr, pp, ps, b.ptr, buf.data_ptr(), m * n, m, k, n, 1,
                                                       device.stream())
                if rc == 0:
                    return buf, buf.data_ptr(), 2
                if rc != -3:                                               # NPM_ERR_UNSUPPORTED: project to fp32 instead
                    raise _lib.NpmError(f'npm_linear_fwd_planes failed (rc={rc}): {_lib.last_error()}')
                self._planes_mode = False
            y = self._project(x2d, w, b, n, planes=pl(w))
            return y, y.ptr, 4

        # input projections (attentions.py:88-100); q/k/v are [B,S,H,dk] column blocks of the outputs, token stride ld
        if packs is not None and key is query and value is query:
            self._mode = 'qkv'
            qkv, p0, es = project(query2, packs[0], packs[1], 3 * hd)                      # [B*S, 3*H*dk]
            self._proj = (qkv,)
            self._qkv_ptrs = (p0, p0 + es * hd, p0 + 2 * es * hd)
            self._qkv_ld = (3 * hd, 3 * hd, 3 * hd)
            self._qkv_plane = (batch * sq * 3 * hd,) * 3
        elif packs is not None and value is key:
            self._mode = 'kv'
            q2, pq, es = project(query2, wq, bq, hd)
            kv_w = packs[0][1:3]
            if es == 2:
                kv2, pk, es2 = project(key2, kv_w, packs[1][1:3], 2 * hd)                  # [B*Skv, 2*H*dk]
                assert es2 == 2, 'q projected to planes but k | v could not be'
            else:
                kv2 = self._project(key2, kv_w, packs[1][1:3], 2 * hd, planes=pl(kv_w))
                pk = kv2.ptr
            self._proj = (q2, kv2)
            self._qkv_ptrs = (pq, pk, pk + es * hd)
            self._qkv_ld = (hd, 2 * hd, 2 * hd)
            self._qkv_plane = (batch * sq * hd, batch * skv * 2 * hd, batch * skv * 2 * hd)
        else:
            self._mode = 'separate'
            self._planes_mode = False
            q2 = self._project(query2, wq, bq, hd, planes=pl(wq))
            k2 = self._project(key2, wk, bk, hd, planes=pl(wk))
            v2 = self._project(value.reshape(batch * skv, value.shape[2]), wv, bv, h * dv, planes=pl(wv))
            self._proj = (q2, k2, v2)
            self._qkv_ptrs = (q2.ptr, k2.ptr, v2.ptr)
            self._qkv_ld = (hd, hd, h * dv)
            self._qkv_plane = (0, 0, 0)

        # planes mode: `saved` is only the log-sum-exp (the size the TF32 fused path reports)
        self._saved = device.workspace(C.npm_mha_core_saved_bytes_for(1 if self._planes_mode else self._path, batch, h, sq, skv, dk, dv))
        values = device.empty((batch, sq, h, dv))          # [B, Sq, H, dv] (reference keeps [B,H,Sq,dv])
        assert not self._causal or sq == skv, 'causal attention needs seq_len_q == seq_len_kv'
        ld = self._strides()
        qp, kp, vp = self._qkv_ptrs
        C.npm_mha_core_fwd_strided(qp, kp, vp, values.ptr, self._saved.data_ptr(), batch, h, sq, skv, dk, dv,
                                   ctypes.byref(ld), device.stream())
        self._values = values

        o = self._project(values.reshape(batch * sq, h * dv), wo, bo, wo.shape[0], _residual, planes=pl(wo))
        return o.reshape(batch, sq, wo.shape[0])

    def _strides(self, **grads):
        pm = self._planes_mode
        return MhaStrides(q=self._qkv_ld[0], k=self._qkv_ld[1], v=self._qkv_ld[2], causal=int(self._causal), path=1 + self._path,
                          planes=int(pm), q_plane=self._qkv_plane[0] if pm else 0, k_plane=self._qkv_plane[1] if pm else 0,
                          v_plane=self._qkv_plane[2] if pm else 0, **grads)

    # ---- decode-time step with a key/value cache (inference) ---------------------------------------
    def decode_step(self, x_t, cache: KVCache, memory=None, _residual=None):
        """One new token per sequence: `x_t` [B, 1, D].  memory is None: causal self-attention over the tokens decoded so
        far (the new token's k / v are appended to `cache`); else cross-attention over `memory` [B, Skv, D], whose k / v
        projections are computed on the first call and kept in `cache`.  Equals row t of `forward` on the full
        sequence with `causal=True` (tests/test_parity_gpu.py::test_kv_cache_decode_equals_full_forward)."""
        x_t = device.asdevice(x_t)
        batch, one, dmodel = x_t.shape
        assert one == 1, 'decode_step takes one token per sequence'
        h, dk, dv = self._num_heads, self._key_dim, self._value_dim
        assert dk == dv, 'the cached path keeps k | v packed'
        hd = h * dk
        packs = self._packed_params()
        assert packs is not None
        wq, wo, bq, bo = self._p('_wq'), self._p('_wo'), self._p('_bq'), self._p('_bo')
        x2 = x_t.reshape(batch, dmodel)
        q2 = self._project(x2, wq, bq, hd)                                            # [B, H*dk]
        values = device.empty((batch, 1, h, dv))
        s = device.stream()
        if memory is None:
            assert cache.length < cache.max_len, 'KVCache is full'
            kv2 = self._project(x2, packs[0][1:3], packs[1][1:3], 2 * hd)             # [B, 2*H*dk]: k_t | v_t
            t = cache.length
            cache.k.t[t].copy_(kv2.t[:, :hd])
            cache.v.t[t].copy_(kv2.t[:, hd:])
            cache.length = t + 1
            # batch folded into heads: q [1, 1, B*H, d], k / v [length, B*H, d] with token stride B*H*d
            fb, fh, skv = 1, batch * h, cache.length
            ld = MhaStrides(q=batch * hd, k=batch * hd, v=batch * hd, causal=0)       # the last position sees every cached token
            kp, vp = cache.k.ptr, cache.v.ptr
        else:
            memory = device.asdevice(memory)
            skv = memory.shape[1]
            if cache.memory_kv is None:
                cache.memory_kv = self._project(memory.reshape(batch * skv, memory.shape[2]), packs[0][1:3], packs[1][1:3], 2 * hd)
            fb, fh = batch, h
            ld = MhaStrides(q=hd, k=2 * hd, v=2 * hd, causal=0)
            kp, vp = cache.memory_kv.ptr, cache.memory_kv.ptr + 4 * hd
        path = int(C.npm_mha_core_path(fb, fh, 1, skv, dk, dv))
        ld.path = 1 + path
        saved = device.workspace(C.npm_mha_core_saved_bytes_for(path, fb, fh, 1, skv, dk, dv))
        C.npm_mha_core_fwd_strided(q2.ptr, kp, vp, values.ptr, saved.data_ptr(), fb, fh, 1, skv, dk, dv, ctypes.byref(ld), s)
        o = self._project(values.reshape(batch, h * dv), wo, bo, wo.shape[0], _residual)
        return o.reshape(batch, 1, wo.shape[0])

    def backward(self, dy, optimizer_: optimizer.Optimizer, _sum_inputs: bool = False):
        """Returns `(dquery, dkey, dvalue)` (attentions.py:199).  `_sum_inputs=True` (B200 extension used by
        the transformer blocks, which only ever sum these — transformer.py:85,184-185,196) returns
        `(dquery + dkey + dvalue, None)` for self-attention and `(dquery, dkey + dvalue)` when key is value,
        computed by one GEMM over the concatenated contraction instead of separate GEMMs and adds."""
        dy = device.asdevice(dy)
        batch, sq, dmodel = dy.shape
        h, dk, dv = self._num_heads, self._key_dim, self._value_dim
        skv = self._seq_len_kv
        assert sq == self._seq_len_q
        s = device.stream()
        hd = h * dk
        mode = self._mode
        packs = self._packed_params() if mode != 'separate' else None
        if mode != 'separate' and packs is None:
            raise RuntimeError('MultiHeadAttention: projection parameters changed shape between forward and backward')
        wq, wk, wv, wo = self._p('_wq'), self._p('_wk'), self._p('_wv'), self._p('_wo')
        for name in ('_bq', '_bk', '_bv', '_bo'):
            self._p(name)

        def grads_into(x2d, dy2d, dw, db):
            """dw [n, k] (output-major) = dy^T x, db [n] = column sums of dy, for y = x @ W^T + b."""
            m, k = x2d.shape
            n = dy2d.shape[1]
            if dy2d.colsum is not None:
                # the kernel that produced dy (fused LayerNorm backward) already summed its columns (attentions.py:129)
                db.copy_from(dy2d.colsum)
                C.npm_linear_bwd_dw_db(x2d.ptr, dy2d.ptr, dw.ptr, None, m, k, n, 1, None, s)
                return
            ws = device.workspace(C.npm_colsum_workspace(m, n))
            C.npm_linear_bwd_dw_db(x2d.ptr, dy2d.ptr, dw.ptr, db.ptr, m, k, n, 1, ws.data_ptr(), s)

        def grads(x2d, dy2d, w_attr, b_attr):
            dw = optimizer_.grad_buffer(self, w_attr, getattr(self, w_attr).shape)
            db = optimizer_.grad_buffer(self, b_attr, getattr(self, b_attr).shape)
            grads_into(x2d, dy2d, dw, db)
            return dw, db

        def dinput(dy2d, w, k, n=None, ld=None, ptr=None):
            """dx [m, k] = dy [m, n] @ W for an output-major W [n, k]; dy may be a column block (ptr, ld)."""
            m = dy2d.shape[0]
            n = dy2d.shape[1] if n is None else n
            dx = device.empty((m, k))
            find = getattr(self, '_find_planes', None)
            pp, ps = (find(w) if find is not None else None) or (None, 0)
            if ld is None:
                C.npm_linear_bwd_dx_presplit(dy2d.ptr, w.ptr, pp, ps, dx.ptr, m, k, n, 1, s)
            else:
                d = GemmDesc(a=ptr, b=w.ptr, c=dx.ptr, bias=None, m=m, n=k, k=n, a_rs=ld, a_cs=1, b_rs=k, b_cs=1,
                             ldc=k, nb1=1, nb2=1, alpha=1.0, flags=0, precision=-1, residual=None, ldr=0,
                             b_split=pp, b_split_plane=ps)
                C.npm_gemm(ctypes.byref(d), s)
            return dx

        # output projection (attentions.py:129-136)
        dy2 = dy.reshape(batch * sq, dmodel)
        values2 = self._values.reshape(batch * sq, h * dv)
        dwo, dbo = grads(values2, dy2, '_wo', '_bo')
        scratch = device.workspace(C.npm_mha_core_bwd_scratch_bytes_for(self._path, batch, h, sq, skv, dk, dv))
        # split-bf16 fused path: the dX GEMM writes dO straight as the bf16 planes the backward kernels read, and
        # D = rowsum(dO o O) per head from its epilogue registers, both into `scratch` — no fp32 dO, no D / split pass
        do_ready = 0
        if self._path == 2 and (h * dv) % 64 == 0 and batch * sq > 128 and not _NO_DO_ROWDOT:
            find = getattr(self, '_find_planes', None)
            pp, ps = (find(wo) if find is not None else None) or (None, 0)
            d_bytes = (batch * h * sq * 4 + 255) & ~255
            rc = _lib.load().npm_linear_bwd_dx_planes_rowdot(dy2.ptr, wo.ptr, pp, ps, scratch.data_ptr() + d_bytes,
                                                             batch * sq * h * dv, batch * sq, h * dv, dmodel, 1,
                                                             values2.ptr, h * dv, scratch.data_ptr(), sq, s)
            if rc == 0:
                do_ready = 1
            elif rc != -3:                                     # NPM_ERR_UNSUPPORTED: the fp32 route below
                raise _lib.NpmError(f'npm_linear_bwd_dx_planes_rowdot failed (rc={rc}): {_lib.last_error()}')
        dvalues = None if do_ready else dinput(dy2, wo, h * dv)                     # [B*Sq, H*dv]

        # attention core (attentions.py:146-162): dq / dk / dv are written as row blocks mirroring forward's layout
        if mode == 'qkv':
            dqkv = device.empty((batch * sq, 3 * hd))
            dptrs, dld, dbufs = (dqkv.ptr, dqkv.ptr + 4 * hd, dqkv.ptr + 8 * hd), (3 * hd,) * 3, (dqkv,)
        elif mode == 'kv':
            dq2 = device.empty((batch * sq, hd))
            dkv2 = device.empty((batch * skv, 2 * hd))
            dptrs, dld, dbufs = (dq2.ptr, dkv2.ptr, dkv2.ptr + 4 * hd), (hd, 2 * hd, 2 * hd), (dq2, dkv2)
        else:
            dq2 = device.empty((batch * sq, hd))
            dk2 = device.empty((batch * skv, hd))
            dv2 = device.empty((batch * skv, h * dv))
            dptrs, dld, dbufs = (dq2.ptr, dk2.ptr, dv2.ptr), (hd, hd, h * dv), (dq2, dk2, dv2)
        ld = self._strides(dq=dld[0], dk=dld[1], dv=dld[2], do_ready=do_ready)
        qp, kp, vp = self._qkv_ptrs
        C.npm_mha_core_bwd_strided(qp, kp, vp, self._values.ptr, dvalues.ptr if dvalues is not None else None,
                                   self._saved.data_ptr(), dptrs[0],
                                   dptrs[1], dptrs[2], scratch.data_ptr(), batch, h, sq, skv, dk, dv,
                                   ctypes.byref(ld), s)

        # input projections (attentions.py:169-188)
        query2 = self._query.reshape(batch * sq, self._query.shape[2])
        key2 = self._key.reshape(batch * skv, self._key.shape[2])
        value2 = self._value.reshape(batch * skv, self._value.shape[2])
        din = query2.shape[1]
        if mode == 'separate':
            dwq, dbq = grads(query2, dq2, '_wq', '_bq')
            dwk, dbk = grads(key2, dk2, '_wk', '_bk')
            dwv, dbv = grads(value2, dv2, '_wv', '_bv')
            dquery = dinput(dq2, wq, din).reshape(self._query.shape)
            dkey = dinput(dk2, wk, key2.shape[1]).reshape(self._key.shape)
            dvalue = dinput(dv2, wv, value2.shape[1]).reshape(self._value.shape)
            result = (dquery, dkey, dvalue)
        else:
            dwp = optimizer_.grad_buffer_pack(self, ('_wq', '_wk', '_wv'), wq.shape)      # [3,H,dk,D]
            dbp = optimizer_.grad_buffer_pack(self, ('_bq', '_bk', '_bv'), (h, dk))        # [3,H,dk]
            (dwq, dwk, dwv), (dbq, dbk, dbv) = (dwp[0], dwp[1], dwp[2]), (dbp[0], dbp[1], dbp[2])
            if mode == 'qkv':
                grads_into(query2, dqkv, dwp, dbp)
                if _sum_inputs:
                    result = (dinput(dqkv, packs[0], din).reshape(self._query.shape), None)
                else:
                    result = tuple(dinput(dqkv, w, din, n=hd, ld=3 * hd, ptr=dptrs[i]).reshape(self._query.shape)
                                   for i, w in enumerate((wq, wk, wv)))
            else:
                grads_into(query2, dq2, dwq, dbq)
                grads_into(key2, dkv2, dwp[1:3], dbp[1:3])
                dquery = dinput(dq2, wq, din).reshape(self._query.shape)
                if _sum_inputs:
                    result = (dquery, dinput(dkv2, packs[0][1:3], key2.shape[1]).reshape(self._key.shape))
                else:
                    result = (dquery,) + tuple(
                        dinput(dkv2, w, key2.shape[1], n=hd, ld=2 * hd, ptr=dptrs[1 + i]).reshape(self._key.shape)
                        for i, w in enumerate((wk, wv)))
        if _sum_inputs and mode == 'separate':
            assert self._value is self._key, '_sum_inputs is for self-attention and key-is-value cross-attention'
            result = (result[0], result[1] + result[2]) if self._key is not self._query else (_add3(result), None)

        optimizer_.update(self, '_wq', dwq)
        optimizer_.update(self, '_wk', dwk)
        optimizer_.update(self, '_wv', dwv)
        optimizer_.update(self, '_wo', dwo)
        optimizer_.update(self, '_bq', dbq)
        optimizer_.update(self, '_bk', dbk)
        optimizer_.update(self, '_bv', dbv)
        optimizer_.update(self, '_bo', dbo)

        return result

    # ---- views matching the reference's cached intermediates (debug / parity only) ----------
    @property
    def _attention_scores(self):
        """softmax probabilities [B, H, Sq, Skv] (attentions.py:108-111)."""
        q, k = self._q, self._k
        b, sq, h, _ = q.shape
        out = device.empty((b, h, sq, self._seq_len_kv))
        C.npm_mha_core_scores(q.ptr, k.ptr, self._saved.data_ptr(), out.ptr, b, h, sq, self._seq_len_kv,
                              self._key_dim, self._value_dim, device.stream())
        if getattr(self, '_causal', False):     # debug view only: the recomputed probabilities ignore the mask
            import torch
            out.t.masked_fill_(torch.triu(torch.ones(sq, sq, dtype=torch.bool, device=out.t.device), 1), 0.0)
        return out

    def _block(self, i, seq):
        """Dense copy of projected q (0) / k (1) / v (2) as [B, S, H, d] (reference `_q,_k,_v`, attentions.py:92-100)."""
        h, d = self._num_heads, (self._key_dim if i < 2 else self._value_dim)
        batch = self._query.shape[0]
        if getattr(self, '_planes_mode', False):
            # the projections exist only as bf16 hi / mid planes: hi + mid is the value the attention kernels used
            import torch
            j, width = (0, 1) if (self._mode == 'kv' and i == 0) else ((1, 2) if self._mode == 'kv' else (0, 3))
            pl2 = self._proj[j].view(torch.bfloat16).view(2, batch * seq, width * h * d).float().sum(0)
            col = i if self._mode == 'qkv' else (0 if i == 0 else i - 1)
            return device.DeviceArray(pl2.view(batch, seq, width, h, d)[:, :, col].contiguous())
        if self._mode == 'qkv':
            t = self._proj[0].t.view(batch, seq, 3, h, d)[:, :, i]
        elif self._mode == 'kv' and i > 0:
            t = self._proj[1].t.view(batch, seq, 2, h, d)[:, :, i - 1]
        else:
            t = self._proj[i if self._mode == 'separate' else 0].t.view(batch, seq, h, d)
        return device.DeviceArray(t.contiguous())

    @property
    def _q(self):
        return self._block(0, self._seq_len_q)

    @property
    def _k(self):
        return self._block(1, self._seq_len_kv)

    @property
    def _v(self):
        return self._block(2, self._seq_len_kv)

    @property
    def _attention_values(self):
        """[B, H, Sq, dv] as the reference stores them (attentions.py:112-114), on the host."""
        return self._values.numpy().transpose(0, 2, 1, 3)
