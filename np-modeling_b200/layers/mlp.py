"""Linear / Dense — drop-in for layers/mlp.py; GEMMs run on tcgen05 (csrc/gemm_tc.cu)."""
from typing import Optional

import optimizer
from layers import activations, layer
from npm_b200 import device
from npm_b200._lib import C


class Linear(layer.StatefulLayer):
    def __init__(self, units: int, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._output_units = units

    def initialize(self, x, **kwargs) -> None:
        self._input_units = x.shape[-1]
        self._w = self._initializer([self._input_units, self._output_units])
        self._b = self._initializer([self._output_units])

    def forward(self, x, _relu: bool = False, _residual=None):
        """_relu / _residual are B200 extensions used by Dense and the transformer blocks: the activation and
        the `out += skip` that follow this layer in the reference run in the GEMM epilogue."""
        x = device.asdevice(x)
        assert x.ndim == 2, 'Linear takes [m, k] inputs (mlp.py:33)'
        self._x = x
        w, b = self._p('_w'), self._p('_b')
        m, k = x.shape
        n = w.shape[1]
        assert w.shape[0] == k, f'{w.shape} vs input features {k}'
        y = device.empty((m, n))
        assert _residual is None or (not _relu and _residual.size == m * n), 'residual must match the output'
        # bf16x3 mode: the weight is split once here and serves this GEMM and the dX GEMM of backward (the optimizer
        # applies updates after the whole backward pass, so the weight cannot change in between)
        self._w_planes = device.split_weight(w) if m > 128 else None
        C.npm_linear_fwd_presplit(x.ptr, w.ptr, self._w_planes.data_ptr() if self._w_planes is not None else None, w.size,
                                  b.ptr, _residual.ptr if _residual is not None else None, y.ptr, m, k, n, 0,
                                  1 if _relu else 0, device.stream())
        return y

    def backward(self, dy, optimizer_: optimizer.Optimizer, _db=None):
        # dy: [m, n]   b/db: [n]   w/dw: [k, n]   x/dx: [m, k]
        # _db: bias gradient already computed by the caller (Dense fuses it with ReLU.backward)
        dy = device.asdevice(dy)
        w = self._p('_w')
        self._p('_b')
        x = self._x
        assert dy.shape == (x.shape[0], w.shape[1])
        m, k = x.shape
        n = w.shape[1]
        s = device.stream()
        dw = optimizer_.grad_buffer(self, '_w', (k, n))
        if _db is None and dy.colsum is not None:
            # the kernel that produced dy (fused LayerNorm backward) already summed its columns: mlp.py:34 for free
            db = optimizer_.grad_buffer(self, '_b', (n,)).copy_from(dy.colsum)
            C.npm_linear_bwd_dw_db(x.ptr, dy.ptr, dw.ptr, None, m, k, n, 0, None, s)
        elif _db is None:
            db = optimizer_.grad_buffer(self, '_b', (n,))
            ws = device.workspace(C.npm_colsum_workspace(m, n))
            C.npm_linear_bwd_dw_db(x.ptr, dy.ptr, dw.ptr, db.ptr, m, k, n, 0, ws.data_ptr(), s)
        else:
            db = _db
            C.npm_linear_bwd_dw_db(x.ptr, dy.ptr, dw.ptr, None, m, k, n, 0, None, s)
        dx = device.empty((m, k))
        planes = getattr(self, '_w_planes', None)
        C.npm_linear_bwd_dx_presplit(dy.ptr, w.ptr, planes.data_ptr() if planes is not None else None, w.size, dx.ptr, m, k, n,
                                     0, s)
        assert dx.shape == x.shape
        optimizer_.update(self, '_w', dw)
        optimizer_.update(self, '_b', db)
        return dx

    @property
    def w(self):
        assert self._initialized
        return self._w

    @property
    def b(self):
        assert self._initialized
        return self._b


class Dense(layer.StatefulLayer):
    """Dense w/ ReLU activation."""
    def __init__(self,
                 units: int,
                 activation: Optional[activations.Activation] = None,
                 *args,
                 **kwargs):
        super().__init__(*args, **kwargs)
        self._linear = Linear(units=units)
        self._activation = activation or activations.ReLU()

    def initialize(self, x, **kwargs) -> None:
        self._linear.initialize(x)
        self._linear._initialized = True
        self._activation.initialize()
        self._activation._initialized = True

    def _fused_relu(self):
        # The default activation is fused into the GEMM epilogue (bias + ReLU, pre-activation never
        # written); anything else (e.g. Softmax, a user subclass) runs as its own layer.
        return type(self._activation) is activations.ReLU

    def forward(self, x, _alias_ok: bool = False):
        """`_alias_ok=True` (the transformer blocks, which never write into this output) returns the buffer the
        backward pass reads its ReLU mask from; any other caller gets its own copy, so that the reference's idiom
        `out = layer(x); out += skip` (in-place on a layer output) cannot destroy the mask."""
        if self._fused_relu():
            self._y = self._linear.forward(x, _relu=True)
            return self._y if _alias_ok else self._y.copy()
        y = self._linear.forward(x)
        return self._activation.forward(y)

    def backward(self, dy, optimizer_: optimizer.Optimizer):
        if self._fused_relu():
            # ReLU.backward (activations.py:17-19) and the bias gradient (mlp.py:34) in one pass over dy
            dy = device.asdevice(dy)
            y = self._y
            assert dy.shape == y.shape, f'{dy.shape} vs {y.shape}'
            m, n = y.shape
            db = optimizer_.grad_buffer(self._linear, '_b', (n,))
            dz = device.empty((m, n))
            ws = device.workspace(C.npm_colsum_workspace(m, n))
            C.npm_relu_bwd_colsum(y.ptr, dy.ptr, dz.ptr, db.ptr, m, n, ws.data_ptr(), device.stream())
            return self._linear.backward(dz, optimizer_, _db=db)
        dy = self._activation.backward(dy)
        return self._linear.backward(dy, optimizer_)

    @property
    def linear(self) -> Linear:
        assert self._initialized
        return self._linear
