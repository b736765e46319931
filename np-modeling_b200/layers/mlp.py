"""Linear / Dense — drop-in for layers/mlp.py; GEMMs run on tcgen05 (csrc/gemm_tc.cu, csrc/gemm_bx.cu)."""
import ctypes
import os
from typing import Optional

import optimizer
from layers import activations, layer
from npm_b200 import _lib, device
from npm_b200._lib import C, GemmDesc

# Split-bf16 ('bf16x3') mode, transformer FFN only (`_planes_ok`): the hidden activation is written by the first GEMM's
# epilogue ONLY as bf16 hi / mid planes (device.PlanesArray) and its gradient by the ReLU backward the same way, so the
# second FFN GEMM, the first layer's dX GEMM and both dW GEMMs land those operands by TMA without converting them in
# shared memory.  Measured one kernel at a time (tools/ffn_planes_probe.py): dW GEMMs 144 -> 133 us, first-layer dX
# 137 -> 133 us, ReLU backward 82 -> 78 us; cfg5 step (three alternating same-box pairs) 87.57 -> 86.68 ms.  Results equal
# the fp32 route's to fp32 round-off (the planes hold exactly the hi / mid pairs the GEMM's converters would make).
# NPM_NO_FFN_PLANES=1 keeps everything fp32 (A/B; tests flip the attribute).
_NO_FFN_PLANES = bool(os.environ.get('NPM_NO_FFN_PLANES'))
_UNSUPPORTED = -3


def _gemm(**kw):
    """npm_gemm with keyword fields -> return code (0, or NPM_ERR_UNSUPPORTED for the caller's fp32 fallback)."""
    d = GemmDesc(nb1=1, nb2=1, alpha=1.0, precision=-1, **kw)
    rc = _lib.load().npm_gemm(ctypes.byref(d), device.stream())
    if rc not in (0, _UNSUPPORTED):
        raise _lib.NpmError(f'npm_gemm failed (rc={rc}): {_lib.last_error()}')
    return rc


class Linear(layer.StatefulLayer):
    def __init__(self, units: int, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._output_units = units

    def initialize(self, x, **kwargs) -> None:
        self._input_units = x.shape[-1]
        self._w = self._initializer([self._input_units, self._output_units])
        self._b = self._initializer([self._output_units])

    def forward(self, x, _relu: bool = False, _residual=None, _planes_out: bool = False):
        """_relu / _residual are B200 extensions used by Dense and the transformer blocks: the activation and
        the `out += skip` that follow this layer in the reference run in the GEMM epilogue.  `_planes_out` asks for the
        result as a device.PlanesArray (split-bf16 mode only; fp32 otherwise); `x` may be one."""
        x = device.asdevice(x)
        assert x.ndim == 2, 'Linear takes [m, k] inputs (mlp.py:33)'
        w, b = self._p('_w'), self._p('_b')
        m, k = x.shape
        n = w.shape[1]
        assert w.shape[0] == k, f'{w.shape} vs input features {k}'
        assert _residual is None or (not _relu and _residual.size == m * n), 'residual must match the output'
        # bf16x3 mode: the weight is split once here and serves this GEMM and the dX GEMM of backward (the optimizer
        # applies updates after the whole backward pass, so the weight cannot change in between)
        self._w_planes = device.split_weight(w) if m > 128 else None
        wp = self._w_planes.data_ptr() if self._w_planes is not None else None
        x_planes = isinstance(x, device.PlanesArray)
        if x_planes or (_planes_out and wp is not None and n % 8 == 0 and _residual is None):
            # y[m,n] = x[m,k] @ W[k,n] + b through the general descriptor: A and / or C as split-bf16 planes
            common = dict(b=w.ptr, bias=b.ptr, m=m, n=n, k=k, a_rs=k, a_cs=1, b_rs=n, b_cs=1, ldc=n, flags=1 if _relu else 0,
                          b_split=wp, b_split_plane=w.size)
            a_args = dict(a=None, a_split=x.hi_ptr, a_split_plane=x.size) if x_planes else dict(a=x.ptr)
            if _planes_out and wp is not None and n % 8 == 0 and _residual is None:
                buf = device.workspace(4 * m * n)
                if _gemm(c=None, c_split=buf.data_ptr(), c_split_plane=m * n, **a_args, **common) == 0:
                    self._x = x
                    return device.PlanesArray(buf, (m, n))
            if x_planes:
                y = device.empty((m, n))
                if _gemm(c=y.ptr, residual=_residual.ptr if _residual is not None else None, ldr=n, **a_args, **common) == 0:
                    self._x = x
                    return y
                x = x.to_fp32()       # the split-bf16 kernel does not take this problem (mode changed, shape): fp32 route
        self._x = x
        y = device.empty((m, n))
        C.npm_linear_fwd_presplit(x.ptr, w.ptr, wp, w.size, b.ptr, _residual.ptr if _residual is not None else None, y.ptr,
                                  m, k, n, 0, 1 if _relu else 0, device.stream())
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
        planes = getattr(self, '_w_planes', None)
        if isinstance(x, device.PlanesArray) or isinstance(dy, device.PlanesArray):
            done = self._backward_planes(x, dy, w, dw, _db, planes, optimizer_)
            if done is not None:
                return done
            x, dy = device.asfp32(x), device.asfp32(dy)       # fp32 route (the split-bf16 kernel declined)
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
        C.npm_linear_bwd_dx_presplit(dy.ptr, w.ptr, planes.data_ptr() if planes is not None else None, w.size, dx.ptr, m, k, n,
                                     0, s)
        assert dx.shape == x.shape
        optimizer_.update(self, '_w', dw)
        optimizer_.update(self, '_b', db)
        return dx

    def _backward_planes(self, x, dy, w, dw, db, planes, optimizer_):
        """backward with x and / or dy as split-bf16 planes (mlp.py:34-38): dw[k,n] = x^T dy with the planes as the
        MN-major A / B operand images, dx[m,k] = dy @ W^T with dy as the K-major A image.  Returns dx, or None when the
        split-bf16 GEMM does not take one of the problems (the caller then joins the planes and runs the fp32 route)."""
        m, k = x.shape
        n = w.shape[1]
        is_planes = lambda t: isinstance(t, device.PlanesArray)
        if db is None:
            if is_planes(dy):
                return None                                 # a planes-only dy always comes with its bias gradient (Dense)
            db = optimizer_.grad_buffer(self, '_b', (n,))
            if dy.colsum is not None:
                db.copy_from(dy.colsum)                     # summed by the kernel that produced dy (fused LayerNorm backward)
            else:
                ws = device.workspace(C.npm_colsum_workspace(m, n))
                C.npm_colsum(dy.ptr, db.ptr, m, n, ws.data_ptr(), device.stream())
        xa = dict(a=None, a_split=x.hi_ptr, a_split_plane=x.size) if is_planes(x) else dict(a=x.ptr)
        yb = dict(b=None, b_split=dy.hi_ptr, b_split_plane=dy.size) if is_planes(dy) else dict(b=dy.ptr)
        if _gemm(c=dw.ptr, m=k, n=n, k=m, a_rs=1, a_cs=k, b_rs=n, b_cs=1, ldc=n, flags=0, **xa, **yb) != 0:
            return None
        dx = device.empty((m, k))
        wp = planes.data_ptr() if planes is not None else None
        if is_planes(dy):
            # dx[m,k] = dy[m,n] @ W^T: A = dy planes (K-major), B(kk=n, nn=k) = W[k,n] at k*n + n
            if _gemm(a=None, a_split=dy.hi_ptr, a_split_plane=dy.size, b=w.ptr, c=dx.ptr, m=m, n=k, k=n, a_rs=n, a_cs=1,
                     b_rs=1, b_cs=n, ldc=k, flags=0, b_split=wp, b_split_plane=w.size) != 0:
                return None
        else:
            C.npm_linear_bwd_dx_presplit(dy.ptr, w.ptr, wp, w.size, dx.ptr, m, k, n, 0, device.stream())
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

    def forward(self, x, _alias_ok: bool = False, _planes_ok: bool = False):
        """`_alias_ok=True` (the transformer blocks, which never write into this output) returns the buffer the
        backward pass reads its ReLU mask from; any other caller gets its own copy, so that the reference's idiom
        `out = layer(x); out += skip` (in-place on a layer output) cannot destroy the mask.  `_planes_ok=True` (the
        same callers, whose next layer is a Linear of this package): in split-bf16 mode the output may be a
        device.PlanesArray — the activation then exists only as bf16 hi / mid planes."""
        if self._fused_relu():
            self._y = self._linear.forward(x, _relu=True, _planes_out=_alias_ok and _planes_ok and not _NO_FFN_PLANES)
            if isinstance(self._y, device.PlanesArray):
                return self._y
            return self._y if _alias_ok else self._y.copy()
        y = self._linear.forward(x)
        return self._activation.forward(y)

    def backward(self, dy, optimizer_: optimizer.Optimizer):
        if self._fused_relu():
            # ReLU.backward (activations.py:17-19) and the bias gradient (mlp.py:34) in one pass over dy
            dy = device.asfp32(dy)
            y = self._y
            assert dy.shape == y.shape, f'{dy.shape} vs {y.shape}'
            m, n = y.shape
            db = optimizer_.grad_buffer(self._linear, '_b', (n,))
            ws = device.workspace(C.npm_colsum_workspace(m, n))
            if isinstance(y, device.PlanesArray):
                # the gate is the sign of y's bf16 hi plane; dz goes out as planes too (operand of the dX and dW GEMMs)
                buf = device.workspace(4 * m * n)
                rc = _lib.load().npm_relu_bwd_colsum_planes(y.hi_ptr, dy.ptr, buf.data_ptr(), m * n, db.ptr, m, n, ws.data_ptr(),
                                                            device.stream())
                if rc == 0:
                    return self._linear.backward(device.PlanesArray(buf, (m, n)), optimizer_, _db=db)
                if rc != _UNSUPPORTED:
                    raise _lib.NpmError(f'npm_relu_bwd_colsum_planes failed (rc={rc}): {_lib.last_error()}')
                y = y.to_fp32()
            dz = device.empty((m, n))
            C.npm_relu_bwd_colsum(y.ptr, dy.ptr, dz.ptr, db.ptr, m, n, ws.data_ptr(), device.stream())
            return self._linear.backward(dz, optimizer_, _db=db)
        dy = self._activation.backward(dy)
        return self._linear.backward(dy, optimizer_)

    @property
    def linear(self) -> Linear:
        assert self._initialized
        return self._linear
