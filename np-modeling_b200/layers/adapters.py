"""Adapters that are NOT in the reference (clearly labelled extensions).

The reference's Trainer chains single-input layers (train.py:28-38), so a TransformerDecoder
(two inputs, tuple gradient — transformer.py:119,203) cannot be trained by it.  `DecoderStack`
is the thinnest container that makes the BASELINE cfg5 workload (24 decoder layers sharing one
memory `kv`) expressible: it follows the decoder's own protocol — `forward(q, kv)`,
`backward(dy, optimizer_) -> (dq, dkv)` — and adds no arithmetic of its own beyond summing the
per-layer `dkv`.
"""
from layers import layer, transformer
from npm_b200 import device


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
