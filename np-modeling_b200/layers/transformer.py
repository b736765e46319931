"""TransformerEncoder / TransformerDecoder — drop-in for layers/transformer.py (:8-203).

Glue only, exactly the reference's dataflow (pre-norm = DropOut → LayerNorm → sublayer → +skip;
post-norm = sublayer → +skip → DropOut → LayerNorm; no causal mask), over device buffers: every
`+=`, reshape and gradient sum below is a device operation.
"""
from layers import attentions, layer, mlp, normalizations
from npm_b200 import device
from npm_b200._lib import C


def _sum3(t):
    """np.sum(dy, axis=0) over MultiHeadAttention.backward's 3-tuple (transformer.py:85,196)."""
    a, b, c = t
    out = device.empty(a.shape)
    C.npm_add3(a.ptr, b.ptr, c.ptr, out.ptr, a.size, device.stream())
    return out


class TransformerEncoder(layer.Layer):
    def __init__(self,
                 num_heads: int,
                 hidden_units: int,
                 norm_first: bool,
                 drop_rate: float = 0.0,
                 *args,
                 causal: bool = False,
                 **kwargs):
        # causal (keyword-only, beyond the reference — SURVEY.md §8 f1): causal mask in the self-attention
        super().__init__(*args, **kwargs)
        self._self_attention = attentions.MultiHeadAttention(num_heads, causal=causal)
        self._dense1 = mlp.Dense(units=hidden_units)
        self._norm1 = normalizations.LayerNormalization()
        self._norm2 = normalizations.LayerNormalization()
        self._norm_first = norm_first
        self._dropout1 = normalizations.DropOut(drop_rate)
        self._dropout2 = normalizations.DropOut(drop_rate)

    def initialize(self, qkv):
        features = qkv.shape[-1]
        self._dense2 = mlp.Linear(units=features)  # No activation

    def forward(self, qkv):
        qkv = device.asdevice(qkv)
        batch, seq_len_q, features = qkv.shape

        skip = qkv
        if self._norm_first:
            qkv = normalizations.dropout_layernorm_forward(self._dropout1, self._norm1, qkv)
        out = self._self_attention(qkv, _residual=skip)      # `out += skip` (transformer.py:39) in the GEMM epilogue
        if not self._norm_first:
            out = self._dropout1(out)
            out = self._norm1(out)

        # Linear takes 2-D inputs only (mlp.py:33)
        out = out.reshape(-1, features)
        skip = out

        if self._norm_first:
            out = normalizations.dropout_layernorm_forward(self._dropout2, self._norm2, out, _planes_out=not mlp._NO_FFN_PLANES)
        out = self._dense1(out, _alias_ok=True, _planes_ok=True)
        out = self._dense2(out, _residual=skip)               # `out += skip` (transformer.py:53,155)
        if not self._norm_first:
            out = self._dropout2(out)
            out = self._norm2(out)

        return out.reshape(batch, seq_len_q, features)

    def backward(self, dy, optimizer_):
        dy = device.asdevice(dy)
        batch, seq_len_q, features = dy.shape

        dy = dy.reshape(-1, features)
        if not self._norm_first:
            dy = self._norm2.backward(dy, optimizer_)
            dy = self._dropout2.backward(dy)
        dskip = dy
        dy = self._dense2.backward(dy, optimizer_)
        dy = self._dense1.backward(dy, optimizer_)
        if self._norm_first:
            dy = normalizations.dropout_layernorm_backward(self._dropout2, self._norm2, dy, dskip, optimizer_)
        else:
            dy += dskip
        dy = dy.reshape(batch, seq_len_q, features)

        if not self._norm_first:
            dy = self._norm1.backward(dy, optimizer_)
            dy = self._dropout1.backward(dy)
        dskip = dy
        dy, _ = self._self_attention.backward(dy, optimizer_, _sum_inputs=True)   # np.sum(dy, axis=0) (:85,196)
        if self._norm_first:
            dy = normalizations.dropout_layernorm_backward(self._dropout1, self._norm1, dy, dskip, optimizer_)
        else:
            dy += dskip

        return dy


class TransformerDecoder(layer.Layer):
    def __init__(self,
                 num_heads: int,
                 hidden_units: int,
                 norm_first: bool,
                 drop_rate: float = 0.0,
                 *args,
                 causal: bool = False,
                 **kwargs):
        # causal (keyword-only, beyond the reference — SURVEY.md §8 f1): causal mask in the SELF-attention (the
        # GPT-style decoder of north_star); the cross-attention over `kv` stays unmasked
        super().__init__(*args, **kwargs)
        self._self_attention = attentions.MultiHeadAttention(num_heads, causal=causal)
        self._cross_attention = attentions.MultiHeadAttention(num_heads)
        self._dense1 = mlp.Dense(units=hidden_units)
        self._norm1 = normalizations.LayerNormalization()
        self._norm2 = normalizations.LayerNormalization()
        self._norm3 = normalizations.LayerNormalization()
        self._norm_first = norm_first
        self._dropout1 = normalizations.DropOut(drop_rate)
        self._dropout2 = normalizations.DropOut(drop_rate)
        self._dropout3 = normalizations.DropOut(drop_rate)

    def initialize(self, q, kv):
        features = q.shape[-1]
        self._dense2 = mlp.Linear(units=features)  # No activation

    def forward(self, q, kv):
        q = device.asdevice(q)
        kv = device.asdevice(kv)
        batch, seq_len_q, features = q.shape

        skip = q
        if self._norm_first:
            q = normalizations.dropout_layernorm_forward(self._dropout1, self._norm1, q)
        out = self._self_attention(q, _residual=skip)        # `out += skip` (transformer.py:129) in the GEMM epilogue
        if not self._norm_first:
            out = self._dropout1(out)
            out = self._norm1(out)

        skip = out

        if self._norm_first:
            out = normalizations.dropout_layernorm_forward(self._dropout2, self._norm2, out)
        out = self._cross_attention(out, kv, _residual=skip)  # `out += skip` (transformer.py:143)
        if not self._norm_first:
            out = self._dropout2(out)
            out = self._norm2(out)

        out = out.reshape(-1, features)
        skip = out

        if self._norm_first:
            out = normalizations.dropout_layernorm_forward(self._dropout3, self._norm3, out, _planes_out=not mlp._NO_FFN_PLANES)
        out = self._dense1(out, _alias_ok=True, _planes_ok=True)
        out = self._dense2(out, _residual=skip)               # `out += skip` (transformer.py:53,155)
        if not self._norm_first:
            out = self._dropout3(out)
            out = self._norm3(out)

        return out.reshape(batch, seq_len_q, features)

    # ---- decode-time step with key/value caches (inference; SURVEY.md §8 f4, transformer.py:120 "TODO: support cache")
    def new_cache(self, batch: int, max_len: int):
        h, dk = self._self_attention._num_heads, self._self_attention._key_dim
        return (attentions.KVCache(max_len, batch, h * dk, h * dk), attentions.KVCache(0, batch, h * dk, h * dk))

    def decode_step(self, x_t, kv, cache):
        """`forward` for ONE new token per sequence (`x_t` [B, 1, D]) given the tokens already decoded into `cache`
        (from `new_cache`); dropout is the identity (inference).  The layer must have been initialised by a forward
        call (or a loaded checkpoint) with `causal=True` semantics in mind: position t sees positions <= t."""
        x_t = device.asdevice(x_t)
        batch, one, features = x_t.shape
        self_cache, cross_cache = cache
        skip = x_t
        out = self._norm1(x_t) if self._norm_first else x_t
        out = self._self_attention.decode_step(out, self_cache, _residual=skip)
        if not self._norm_first:
            out = self._norm1(out)
        skip = out
        h = self._norm2(out) if self._norm_first else out
        out = self._cross_attention.decode_step(h, cross_cache, memory=kv, _residual=skip)
        if not self._norm_first:
            out = self._norm2(out)
        out = out.reshape(-1, features)
        skip = out
        h = self._norm3(out) if self._norm_first else out
        h = self._dense1(h, _alias_ok=True)
        out = self._dense2(h, _residual=skip)
        if not self._norm_first:
            out = self._norm3(out)
        return out.reshape(batch, 1, features)

    def backward(self, dy, optimizer_):
        dy = device.asdevice(dy)
        batch, seq_len_q, features = dy.shape

        dy = dy.reshape(-1, features)
        if not self._norm_first:
            dy = self._norm3.backward(dy, optimizer_)
            dy = self._dropout3.backward(dy)
        dskip = dy
        dy = self._dense2.backward(dy, optimizer_)
        dy = self._dense1.backward(dy, optimizer_)
        if self._norm_first:
            dy = normalizations.dropout_layernorm_backward(self._dropout3, self._norm3, dy, dskip, optimizer_)
        else:
            dy += dskip
        dy = dy.reshape(batch, seq_len_q, features)

        if not self._norm_first:
            dy = self._norm2.backward(dy, optimizer_)
            dy = self._dropout2.backward(dy)
        dskip = dy
        dy, dkv = self._cross_attention.backward(dy, optimizer_, _sum_inputs=True)   # dkv = dkey + dvalue (:184)
        if self._norm_first:
            dy = normalizations.dropout_layernorm_backward(self._dropout2, self._norm2, dy, dskip, optimizer_)
        else:
            dy += dskip
        if not self._norm_first:
            dy = self._norm1.backward(dy, optimizer_)
            dy = self._dropout1.backward(dy)
        dskip = dy
        dy, _ = self._self_attention.backward(dy, optimizer_, _sum_inputs=True)   # np.sum(dy, axis=0) (:85,196)
        if self._norm_first:
            dy = normalizations.dropout_layernorm_backward(self._dropout1, self._norm1, dy, dskip, optimizer_)
        else:
            dy += dskip

        return dy, dkv
