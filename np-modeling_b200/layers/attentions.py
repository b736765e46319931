"""MultiHeadAttention — drop-in for layers/attentions.py (:11-199).

Parameter attributes and layouts are the reference's: `_wq,_wk [H,dk,D]`, `_wv [H,dv,D]`,
`_wo [D,H,dv]`, `_bq,_bk [H,dk]`, `_bv [H,dv]`, `_bo [D]`; `backward` returns the 3-tuple
`(dquery, dkey, dvalue)`.  The projection weights viewed as `[H*dk, D]` / `[D, H*dv]` are already
"output-major" GEMM operands (contraction index contiguous), so every projection and its two
gradients are single tcgen05 GEMMs on the arrays as stored; the attention core runs behind
`npm_mha_core_fwd/bwd`.
"""
from typing import Optional

import optimizer
from layers import activations, layer
from npm_b200 import device
from npm_b200._lib import C


class MultiHeadAttention(layer.StatefulLayer):
    def __init__(self, num_heads: int, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._num_heads = num_heads
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

    def _project(self, x2d, w, b, n):
        m, k = x2d.shape
        y = device.empty((m, n))
        C.npm_linear_fwd(x2d.ptr, w.ptr, b.ptr, y.ptr, m, k, n, 1, 0, device.stream())
        return y

    def forward(self, query, key=None, value=None, mask=None):
        if mask is not None:
            # the reference's `if mask:` raises for arrays (attentions.py:84) and its backward is
            # NotImplementedError (:152-153): attention is unmasked.
            raise ValueError('MultiHeadAttention: mask is not supported (reference attentions.py:84,152)')
        query = device.asdevice(query)
        key = query if key is None else device.asdevice(key)
        value = key if value is None else device.asdevice(value)

        self._query, self._key, self._value, self._mask = query, key, value, mask
        batch, sq, dmodel = query.shape
        skv = key.shape[1]
        h, dk, dv = self._num_heads, self._key_dim, self._value_dim
        assert sq == self._seq_len_q and skv == self._seq_len_kv, 'sequence lengths are fixed at init (:142-145)'

        wq, wk, wv, wo = self._p('_wq'), self._p('_wk'), self._p('_wv'), self._p('_wo')
        bq, bk, bv, bo = self._p('_bq'), self._p('_bk'), self._p('_bv'), self._p('_bo')

        q2 = self._project(query.reshape(batch * sq, dmodel), wq, bq, h * dk)
        k2 = self._project(key.reshape(batch * skv, key.shape[2]), wk, bk, h * dk)
        v2 = self._project(value.reshape(batch * skv, value.shape[2]), wv, bv, h * dv)
        self._q = q2.reshape(batch, sq, h, dk)
        self._k = k2.reshape(batch, skv, h, dk)
        self._v = v2.reshape(batch, skv, h, dv)

        self._saved = device.workspace(C.npm_mha_core_saved_bytes(batch, h, sq, skv, dk, dv))
        values = device.empty((batch, sq, h, dv))          # [B, Sq, H, dv] (reference keeps [B,H,Sq,dv])
        C.npm_mha_core_fwd(q2.ptr, k2.ptr, v2.ptr, values.ptr, self._saved.data_ptr(), batch, h, sq, skv, dk, dv,
                           device.stream())
        self._values = values

        o = self._project(values.reshape(batch * sq, h * dv), wo, bo, wo.shape[0])
        return o.reshape(batch, sq, wo.shape[0])

    def backward(self, dy, optimizer_: optimizer.Optimizer):
        dy = device.asdevice(dy)
        batch, sq, dmodel = dy.shape
        h, dk, dv = self._num_heads, self._key_dim, self._value_dim
        skv = self._seq_len_kv
        assert sq == self._seq_len_q
        s = device.stream()
        wq, wk, wv, wo = self._p('_wq'), self._p('_wk'), self._p('_wv'), self._p('_wo')
        for name in ('_bq', '_bk', '_bv', '_bo'):
            self._p(name)

        def grads(x2d, dy2d, w_attr, b_attr):
            """(dw, db) of y = x @ W^T + b for an output-major W [n, k]."""
            m, k = x2d.shape
            n = dy2d.shape[1]
            dw = optimizer_.grad_buffer(self, w_attr, getattr(self, w_attr).shape)
            db = optimizer_.grad_buffer(self, b_attr, getattr(self, b_attr).shape)
            ws = device.workspace(C.npm_colsum_workspace(m, n))
            C.npm_linear_bwd_dw_db(x2d.ptr, dy2d.ptr, dw.ptr, db.ptr, m, k, n, 1, ws.data_ptr(), s)
            return dw, db

        def dinput(dy2d, w, k):
            m, n = dy2d.shape
            dx = device.empty((m, k))
            C.npm_linear_bwd_dx(dy2d.ptr, w.ptr, dx.ptr, m, k, n, 1, s)
            return dx

        # output projection (attentions.py:129-136)
        dy2 = dy.reshape(batch * sq, dmodel)
        values2 = self._values.reshape(batch * sq, h * dv)
        dwo, dbo = grads(values2, dy2, '_wo', '_bo')
        dvalues = dinput(dy2, wo, h * dv)                     # [B*Sq, H*dv]

        # attention core (attentions.py:146-162)
        dq = device.empty((batch, sq, h, dk))
        dk_ = device.empty((batch, skv, h, dk))
        dv_ = device.empty((batch, skv, h, dv))
        scratch = device.workspace(C.npm_mha_core_bwd_scratch_bytes(batch, h, sq, skv, dk, dv))
        C.npm_mha_core_bwd(self._q.ptr, self._k.ptr, self._v.ptr, self._values.ptr, dvalues.ptr,
                           self._saved.data_ptr(), dq.ptr, dk_.ptr, dv_.ptr, scratch.data_ptr(), batch, h, sq, skv,
                           dk, dv, s)

        # input projections (attentions.py:169-188)
        dq2, dk2, dv2 = dq.reshape(batch * sq, h * dk), dk_.reshape(batch * skv, h * dk), dv_.reshape(batch * skv, h * dv)
        query2 = self._query.reshape(batch * sq, self._query.shape[2])
        key2 = self._key.reshape(batch * skv, self._key.shape[2])
        value2 = self._value.reshape(batch * skv, self._value.shape[2])
        dwq, dbq = grads(query2, dq2, '_wq', '_bq')
        dquery = dinput(dq2, wq, query2.shape[1]).reshape(self._query.shape)
        dwk, dbk = grads(key2, dk2, '_wk', '_bk')
        dkey = dinput(dk2, wk, key2.shape[1]).reshape(self._key.shape)
        dwv, dbv = grads(value2, dv2, '_wv', '_bv')
        dvalue = dinput(dv2, wv, value2.shape[1]).reshape(self._value.shape)

        optimizer_.update(self, '_wq', dwq)
        optimizer_.update(self, '_wk', dwk)
        optimizer_.update(self, '_wv', dwv)
        optimizer_.update(self, '_wo', dwo)
        optimizer_.update(self, '_bq', dbq)
        optimizer_.update(self, '_bk', dbk)
        optimizer_.update(self, '_bv', dbv)
        optimizer_.update(self, '_bo', dbo)

        return dquery, dkey, dvalue

    # ---- views matching the reference's cached intermediates (debug / parity only) ----------
    @property
    def _attention_scores(self):
        """softmax probabilities [B, H, Sq, Skv] (attentions.py:108-111)."""
        b, sq, h, _ = self._q.shape
        out = device.empty((b, h, sq, self._seq_len_kv))
        C.npm_mha_core_scores(self._q.ptr, self._k.ptr, self._saved.data_ptr(), out.ptr, b, h, sq, self._seq_len_kv,
                              self._key_dim, self._value_dim, device.stream())
        return out

    @property
    def _attention_values(self):
        """[B, H, Sq, dv] as the reference stores them (attentions.py:112-114), on the host."""
        return self._values.numpy().transpose(0, 2, 1, 3)
