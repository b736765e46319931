"""Adapters that are NOT in the reference (clearly labelled extensions).

The reference's Trainer chains single-input layers (train.py:28-38), so a TransformerDecoder
(two inputs, tuple gradient — transformer.py:119,203) cannot be trained by it.  `DecoderStack`
is the thinnest container that makes the BASELINE cfg5 workload (24 decoder layers sharing one
memory `kv`) expressible: it follows the decoder's own protocol — `forward(q, kv)`,
`backward(dy, optimizer_) -> (dq, dkv)` — and adds no arithmetic of its own beyond summing the
per-layer `dkv`.
"""
import numpy as np
import torch

import optimizer
from layers import activations, layer, mlp, transformer
from npm_b200 import device
from npm_b200._lib import C


class DecoderStack(layer.Layer):
    def __init__(self, num_layers: int, num_heads: int, hidden_units: int, norm_first: bool,
                 drop_rate: float = 0.0, *args, causal: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self._layers = [transformer.TransformerDecoder(num_heads, hidden_units, norm_first, drop_rate, causal=causal)
                        for _ in range(num_layers)]

    def forward(self, q, kv):
        q = device.asdevice(q)
        kv = device.asdevice(kv)
        for dec in self._layers:
            q = dec(q, kv)            # Layer.__call__: lazy initialisation on the first pass
        return q

    def backward(self, dy, optimizer_):
        dkv = None
        for dec in reversed(self._layers):
            dy, d = dec.backward(dy, optimizer_)
            if dkv is None:
                dkv = d
            else:
                dkv += d
        return dy, dkv


    # ---- decode-time -------------------------------------------------------------------------
    def new_cache(self, batch: int, max_len: int):
        return [dec.new_cache(batch, max_len) for dec in self._layers]

    def decode_step(self, x_t, kv, cache):
        for dec, c in zip(self._layers, cache):
            x_t = dec.decode_step(x_t, kv, c)
        return x_t


class Embedding(layer.StatefulLayer):
    """Token embedding (BEYOND the reference, which has none — SURVEY.md §8 f2): int token ids `[B, S]` ->
    `[B, S, features]` rows of a learned table `_w [vocab, features]`, plus a learned positional table
    `_pos [max_len, features]` when `max_len` is given.  Backward scatters the gradient rows into the tables."""
    PARAMETERS = ('_w', '_pos')

    def __init__(self, vocab: int, features: int, max_len: int = 0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._vocab, self._features, self._max_len = int(vocab), int(features), int(max_len)

    def initialize(self, ids, **kwargs) -> None:
        self._w = self._initializer([self._vocab, self._features])
        if self._max_len:
            self._pos = self._initializer([self._max_len, self._features])

    def _ids(self, ids):
        t = ids if isinstance(ids, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(ids)).astype(np.int32))
        return t.to(device=device._device(), dtype=torch.int32).contiguous()

    def forward(self, ids, offset: int = 0):
        self._tok = self._ids(ids)
        b, s = self._tok.shape
        w = self._p('_w')
        out = device.empty((b, s, self._features))
        C.npm_embedding_fwd(w.ptr, self._tok.data_ptr(), out.ptr, b * s, self._features, self._vocab, device.stream())
        if self._max_len:
            assert offset + s <= self._max_len, 'sequence longer than the positional table'
            self._posids = (torch.arange(s, device=self._tok.device, dtype=torch.int32) + offset).repeat(b).contiguous()
            pos = device.empty((b, s, self._features))
            C.npm_embedding_fwd(self._p('_pos').ptr, self._posids.data_ptr(), pos.ptr, b * s, self._features, self._max_len,
                                device.stream())
            out += pos
        return out

    def backward(self, dy, optimizer_: optimizer.Optimizer):
        dy = device.asdevice(dy)
        b, s = self._tok.shape
        dw = optimizer_.grad_buffer(self, '_w', (self._vocab, self._features))
        C.npm_embedding_bwd(dy.ptr, self._tok.data_ptr(), dw.ptr, b * s, self._features, self._vocab, device.stream())
        optimizer_.update(self, '_w', dw)
        if self._max_len:
            dp = optimizer_.grad_buffer(self, '_pos', (self._max_len, self._features))
            C.npm_embedding_bwd(dy.ptr, self._posids.data_ptr(), dp.ptr, b * s, self._features, self._max_len, device.stream())
            optimizer_.update(self, '_pos', dp)
        return None      # token ids carry no gradient


class GPTStack(layer.Layer):
    """A GPT-shaped model out of the reference's own blocks: Embedding -> `num_layers` x
    TransformerDecoder (causal self-attention; cross-attention over the fixed memory `kv`, which the reference's
    decoder always has, transformer.py:136-143) -> LM head `Dense(vocab, activation=Softmax())`, i.e. probabilities
    for the reference's CrossEntropyLoss (loss.py:32-39).  `forward(ids, kv)`, `backward(dprobs, optimizer_)`."""

    def __init__(self, vocab: int, features: int, num_layers: int, num_heads: int, hidden_units: int, norm_first: bool = True,
                 drop_rate: float = 0.0, max_len: int = 0, *args, causal: bool = True, **kwargs):
        super().__init__(*args, **kwargs)
        self._embed = Embedding(vocab, features, max_len)
        self._stack = DecoderStack(num_layers, num_heads, hidden_units, norm_first, drop_rate, causal=causal)
        self._head = mlp.Dense(vocab, activation=activations.Softmax())
        self._vocab = vocab

    def forward(self, ids, kv):
        x = self._embed(ids)
        b, s, d = x.shape
        x = self._stack(x, kv)
        probs = self._head(x.reshape(b * s, d))
        return probs.reshape(b, s, self._vocab)

    def backward(self, dprobs, optimizer_):
        dprobs = device.asdevice(dprobs)
        b, s, v = dprobs.shape
        dx = self._head.backward(dprobs.reshape(b * s, v), optimizer_)
        dx, dkv = self._stack.backward(dx.reshape(b, s, -1), optimizer_)
        self._embed.backward(dx, optimizer_)
        return None, dkv

    # ---- greedy decoding with key/value caches ---------------------------------------------------
    def new_cache(self, batch: int, max_len: int):
        return dict(layers=self._stack.new_cache(batch, max_len), pos=0)

    def decode_step(self, ids_t, kv, cache):
        """ids_t [B, 1] -> next-token probabilities [B, vocab] given the tokens already fed through `cache`."""
        x = self._embed.forward(ids_t, offset=cache['pos'])
        cache['pos'] += 1
        x = self._stack.decode_step(x, kv, cache['layers'])
        b = x.shape[0]
        return self._head(x.reshape(b, -1))


class EncoderStack(layer.Layer):
    """N TransformerEncoder layers in sequence (a plain list in a Trainer does the same)."""

    def __init__(self, num_layers: int, num_heads: int, hidden_units: int, norm_first: bool,
                 drop_rate: float = 0.0, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._layers = [transformer.TransformerEncoder(num_heads, hidden_units, norm_first, drop_rate)
                        for _ in range(num_layers)]

    def forward(self, x):
        x = device.asdevice(x)
        for enc in self._layers:
            x = enc(x)
        return x

    def backward(self, dy, optimizer_):
        for enc in reversed(self._layers):
            dy = enc.backward(dy, optimizer_)
        return dy
