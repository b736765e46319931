"""DropOut / LayerNormalization — drop-in for layers/normalizations.py."""
import os

import numpy as np
import torch

import optimizer
from layers import layer
from npm_b200 import _lib, device
from npm_b200 import dist as npm_dist
from npm_b200._lib import C

# Process-wide Philox state: every DropOut.forward consumes a fresh counter range of the stream
# keyed by `seed`; `set_dropout_seed` restarts it (the analogue of np.random.seed for the
# reference's legacy-MT19937 binomial stream, normalizations.py:20).
_philox = {'seed': 0x5EED5EED, 'offset': 0}


def set_dropout_seed(seed: int, offset: int = 0) -> None:
    _philox['seed'] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _philox['offset'] = int(offset)


class DropOut(layer.Layer):
    def __init__(self, drop_prob: float, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._drop_prob = drop_prob
        self._ext_mask = None    # injected mask (uint8 device tensor) — parity path
        self._rng = None         # (seed, offset, n) of the last Philox forward
        self._shape = None

    def forward(self, x, training: bool = True):
        if training and self._drop_prob != 0.0:
            x = device.asdevice(x)
            keep_prob = np.float32(1 - self._drop_prob)
            y = device.empty(x.shape)
            self._shape = x.shape
            if self._ext_mask is not None:
                assert self._ext_mask.numel() == x.size, 'injected mask has the wrong size'
                C.npm_dropout_fwd(x.ptr, y.ptr, x.size, keep_prob, 0, 0, self._ext_mask.data_ptr(), device.stream())
            else:
                seed = _philox['seed']
                # this rank's slice of the global tensor's counter range (world-size independent)
                offset, _philox['offset'] = npm_dist.dropout_range(x.size, _philox['offset'], *npm_dist.world())
                self._rng = (seed, offset)
                C.npm_dropout_fwd(x.ptr, y.ptr, x.size, keep_prob, seed, offset, None, device.stream())
            return y
        return x

    def backward(self, dl_dy, *args, **kwargs):
        if self._drop_prob != 0.0:
            # Backward pass only run in training.
            dl_dy = device.asdevice(dl_dy)
            keep_prob = np.float32(1 - self._drop_prob)
            dx = device.empty(dl_dy.shape)
            if self._ext_mask is not None:
                C.npm_dropout_bwd(dl_dy.ptr, dx.ptr, dl_dy.size, keep_prob, 0, 0, self._ext_mask.data_ptr(),
                                  device.stream())
            else:
                seed, offset = self._rng
                C.npm_dropout_bwd(dl_dy.ptr, dx.ptr, dl_dy.size, keep_prob, seed, offset, None, device.stream())
            return dx
        return dl_dy

    # The reference stores an int64 mask array (normalizations.py:20); here the mask is a pure
    # function of (seed, offset) and is only materialised when somebody looks at it.
    @property
    def _mask(self):
        if self._ext_mask is not None:
            return self._ext_mask.cpu().numpy().astype(np.int64).reshape(self._shape or -1)
        if self._rng is None:
            raise AttributeError('_mask')
        n = int(np.prod(self._shape))
        m = torch.empty(n, dtype=torch.uint8, device=device._device())
        C.npm_dropout_mask(m.data_ptr(), n, np.float32(1 - self._drop_prob), self._rng[0], self._rng[1],
                           device.stream())
        return m.cpu().numpy().astype(np.int64).reshape(self._shape)

    @_mask.setter
    def _mask(self, value):
        """Inject a mask (what the reference's own test does, normalizations_test.py:28)."""
        if value is None:
            self._ext_mask = None
            return
        host = np.ascontiguousarray(np.asarray(value) != 0).astype(np.uint8)
        self._ext_mask = torch.from_numpy(host.reshape(-1)).to(device._device())
        self._shape = tuple(np.asarray(value).shape)


class LayerNormalization(layer.StatefulLayer):
    def __init__(self, epsilon: float = 1e-3, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._epsilon = epsilon

    def initialize(self, x):
        self._col = x.shape[-1]
        self._gamma = self._initializer([self._col])
        self._beta = self._initializer([self._col])

    def forward(self, x):
        x = device.asdevice(x)
        self._x = x
        self._fused = None      # set by dropout_layernorm_forward when DropOut ran inside the kernel
        cols = x.shape[-1]
        rows = x.size // cols
        gamma, beta = self._p('_gamma'), self._p('_beta')
        out = device.empty(x.shape)
        self._mean = device.empty((rows,))
        self._rstd = device.empty((rows,))
        C.npm_layernorm_fwd(x.ptr, gamma.ptr, beta.ptr, out.ptr, self._mean.ptr, self._rstd.ptr, rows, cols,
                            float(self._epsilon), device.stream())
        return out

    def backward(self, dl_dz, optimizer_: optimizer.Optimizer):
        dl_dz = device.asdevice(dl_dz)
        x = self._x
        assert dl_dz.shape == x.shape, f'{dl_dz.shape} vs {x.shape}'
        cols = x.shape[-1]
        rows = x.size // cols
        gamma = self._p('_gamma')
        self._p('_beta')
        dx = device.empty(x.shape)
        dgamma = optimizer_.grad_buffer(self, '_gamma', (cols,))
        dbeta = optimizer_.grad_buffer(self, '_beta', (cols,))
        ws = device.workspace(C.npm_layernorm_bwd_workspace(rows, cols))
        # closed form of the [..., C, C] Jacobian einsum (normalizations.py:60-71)
        C.npm_layernorm_bwd(dl_dz.ptr, x.ptr, gamma.ptr, self._mean.ptr, self._rstd.ptr, dx.ptr, dgamma.ptr,
                            dbeta.ptr, rows, cols, ws.data_ptr(), device.stream())
        optimizer_.update(self, '_gamma', dgamma)
        optimizer_.update(self, '_beta', dbeta)
        return dx


# ---- DropOut -> LayerNormalization as ONE kernel (pre-norm transformer blocks) ----------------------
# layers/transformer.py calls `dropout(x)` then `norm(x)` back to back (:35-37, :125-127, ...) and, in
# backward, norm.backward -> dropout.backward -> `dy += dskip` (:90-92, :199-201).  These two helpers are
# result-identical to those call sequences (same Philox mask, same arithmetic) but never write the dropped
# tensor: 1 launch instead of 2 forward, 1 instead of 3 backward.  They fall back to the plain sequence
# whenever the fused kernel does not apply (no dropout, injected mask, wide or unaligned rows).
def dropout_layernorm_forward(drop: DropOut, norm: LayerNormalization, x, _planes_out: bool = False):
    """`_planes_out` (the transformer blocks, for the normalisation in front of the FFN): in split-bf16 mode the result
    may be a device.PlanesArray — it then exists only as the bf16 hi / mid planes the first FFN GEMM and its dW GEMM take
    as their operand image (npm_layernorm_fwd_planes)."""
    x = device.asdevice(x)
    if not norm._initialized:
        norm.initialize(x)
        norm._initialized = True
    cols = x.shape[-1]
    rows = x.size // cols
    if _planes_out and not _NO_LN_PLANES and drop._ext_mask is None and rows > 128 and cols % 8 == 0 and \
            _lib.load().npm_get_precision() == _lib.PREC_BF16X3 and C.npm_dropout_layernorm_fused(rows, cols):
        out = _layernorm_planes(drop, norm, x, rows, cols)
        if out is not None:
            return out
    if drop._drop_prob == 0.0 or drop._ext_mask is not None or not C.npm_dropout_layernorm_fused(rows, cols):
        return norm(drop(x))
    keep_prob = np.float32(1 - drop._drop_prob)
    seed = _philox['seed']
    offset, _philox['offset'] = npm_dist.dropout_range(x.size, _philox['offset'], *npm_dist.world())
    drop._rng, drop._shape = (seed, offset), x.shape
    gamma, beta = norm._p('_gamma'), norm._p('_beta')
    out = device.empty(x.shape)
    norm._x = x                               # the PRE-dropout input: backward re-applies the mask
    norm._mean, norm._rstd = device.empty((rows,)), device.empty((rows,))
    norm._maskbits = device.workspace(C.npm_dropout_layernorm_mask_bytes(rows, cols))   # bit-packed keep mask
    norm._fused = (keep_prob, seed, offset)
    C.npm_dropout_layernorm_fwd(x.ptr, gamma.ptr, beta.ptr, out.ptr, norm._mean.ptr, norm._rstd.ptr,
                                norm._maskbits.data_ptr(), rows, cols, float(norm._epsilon), keep_prob, seed, offset,
                                device.stream())
    return out


_NO_LN_PLANES = bool(os.environ.get('NPM_NO_LN_PLANES'))          # A/B switch for tools


def _layernorm_planes(drop: DropOut, norm: LayerNormalization, x, rows, cols):
    """(DropOut ->) LayerNormalization with the result written only as split-bf16 planes; None = not taken.  Backward is
    the fp32 one in both cases: it reads the layer's INPUT (`norm._x`), never its output."""
    with_drop = drop._drop_prob != 0.0
    gamma, beta = norm._p('_gamma'), norm._p('_beta')
    buf = device.workspace(4 * rows * cols)
    mean, rstd = device.empty((rows,)), device.empty((rows,))
    keep_prob, seed, offset, maskbits = np.float32(1.0), 0, 0, None
    if with_drop:
        keep_prob = np.float32(1 - drop._drop_prob)
        seed = _philox['seed']
        offset, new_offset = npm_dist.dropout_range(x.size, _philox['offset'], *npm_dist.world())
        maskbits = device.workspace(C.npm_dropout_layernorm_mask_bytes(rows, cols))
    rc = _lib.load().npm_layernorm_fwd_planes(x.ptr, gamma.ptr, beta.ptr, buf.data_ptr(), rows * cols, mean.ptr, rstd.ptr,
                                             maskbits.data_ptr() if maskbits is not None else None, rows, cols,
                                             float(norm._epsilon), keep_prob, seed, offset, device.stream())
    if rc == -3:                                   # NPM_ERR_UNSUPPORTED
        return None
    if rc != 0:
        raise _lib.NpmError(f'npm_layernorm_fwd_planes failed (rc={rc}): {_lib.last_error()}')
    norm._x, norm._mean, norm._rstd = x, mean, rstd
    if with_drop:
        _philox['offset'] = new_offset
        drop._rng, drop._shape = (seed, offset), x.shape
        norm._maskbits = maskbits
        norm._fused = (keep_prob, seed, offset)
    else:
        norm._fused = None
    return device.PlanesArray(buf, x.shape)


_NO_COLSUM_RIDE = bool(os.environ.get('NPM_NO_COLSUM_RIDE'))      # A/B switch for tools


def dropout_layernorm_backward(drop: DropOut, norm: LayerNormalization, dz, dskip, optimizer_):
    """dropout.backward(norm.backward(dz)) + dskip"""
    dz = device.asdevice(dz)
    if getattr(norm, '_fused', None) is None:
        dy = drop.backward(norm.backward(dz, optimizer_))
        dy += dskip
        return dy
    keep_prob, seed, offset = norm._fused
    x = norm._x
    assert dz.size == x.size and dskip.size == x.size, f'{dz.shape} / {dskip.shape} vs {x.shape}'
    cols = x.shape[-1]
    rows = x.size // cols
    gamma = norm._p('_gamma')
    dx = device.empty(dz.shape)
    dgamma = optimizer_.grad_buffer(norm, '_gamma', (cols,))
    dbeta = optimizer_.grad_buffer(norm, '_beta', (cols,))
    ws = device.workspace(C.npm_layernorm_bwd_workspace(rows, cols))
    # dx is the gradient of the residual stream: its column sums are the bias gradient of whichever layer wrote the
    # stream (the previous sublayer's output projection / second FFN layer), so they ride along with dx
    if not _NO_COLSUM_RIDE:
        dx.colsum = device.empty((cols,))
    C.npm_dropout_layernorm_bwd_colsum(dz.ptr, x.ptr, gamma.ptr, norm._mean.ptr, norm._rstd.ptr,
                                       norm._maskbits.data_ptr(), dskip.ptr, dx.ptr, dgamma.ptr, dbeta.ptr,
                                       dx.colsum.ptr if dx.colsum is not None else None,
                                       rows, cols, keep_prob, ws.data_ptr(), device.stream())
    optimizer_.update(norm, '_gamma', dgamma)
    optimizer_.update(norm, '_beta', dbeta)
    return dx
