"""Conv2D — drop-in for layers/conv.py (:11-194). NHWC activations, HWIO filters, SAME padding,
stride (1, 1), odd kernel size; kernels in csrc/conv.cu."""
from typing import Optional, Sequence

import optimizer
from layers import activations, layer
from npm_b200 import device
from npm_b200._lib import C


class Conv2D(layer.StatefulLayer):
    """Conv2D w/ ReLU activation.

    Assumes:
      - Padding as 'SAME'.
      - Strides as (1, 1).
    """
    def __init__(self,
                 channels: int,
                 kernel_size: int,
                 padding: str = 'SAME',
                 strides: Sequence[int] = (1, 1),
                 activation: Optional[activations.Activation] = None,
                 *args,
                 **kwargs):
        super().__init__(*args, **kwargs)
        assert padding == 'SAME'
        assert strides == (1, 1)
        self._output_channels = channels
        self._kernel_size = kernel_size
        self._activation = activation or activations.ReLU()

    def initialize(self, x) -> None:
        # x in NHWC format, filters in HWIO format.
        self._input_channels = x.shape[-1]
        self._w = self._initializer([
            self._kernel_size, self._kernel_size, self._input_channels,
            self._output_channels
        ])
        self._b = self._initializer([self._output_channels])
        self._activation.initialize()

    def _dims(self, x):
        n, h, w, c0 = x.shape
        return n, h, w, c0, self._output_channels, self._kernel_size

    def forward(self, x):
        x = device.asdevice(x)
        assert x.ndim == 4 and x.shape[-1] == self._input_channels
        assert self._kernel_size % 2
        self._x = x
        n, h, w, c0, c1, k = self._dims(x)
        f, b = self._p('_w'), self._p('_b')
        y = device.empty((n, h, w, c1))
        ws = device.workspace(C.npm_conv2d_workspace(n, h, w, c0, c1, k))
        C.npm_conv2d_fwd(x.ptr, f.ptr, b.ptr, y.ptr, n, h, w, c0, c1, k, 0, ws.data_ptr(), device.stream())
        return self._activation.forward(y)

    def backward(self, dy, optimizer_: optimizer.Optimizer):
        dy = device.asdevice(dy)
        x = self._x
        assert dy.shape[:3] == x.shape[:3]
        assert dy.shape[3] == self._output_channels
        n, h, w, c0, c1, k = self._dims(x)
        f = self._p('_w')
        self._p('_b')
        s = device.stream()
        dy = self._activation.backward(dy)
        ws = device.workspace(C.npm_conv2d_workspace(n, h, w, c0, c1, k))
        db = optimizer_.grad_buffer(self, '_b', (c1,))
        dw = optimizer_.grad_buffer(self, '_w', (k, k, c0, c1))
        C.npm_conv2d_bwd_dw_db(x.ptr, dy.ptr, dw.ptr, db.ptr, n, h, w, c0, c1, k, ws.data_ptr(), s)
        dx = device.empty(x.shape)
        C.npm_conv2d_bwd_dx(dy.ptr, f.ptr, dx.ptr, n, h, w, c0, c1, k, ws.data_ptr(), s)
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
