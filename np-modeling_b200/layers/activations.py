"""Activations — drop-in for layers/activations.py; math runs in rowops.cu / elementwise.cu."""
from layers import layer
from npm_b200 import device
from npm_b200._lib import C


class Activation(layer.Layer):
    pass


class ReLU(Activation):
    def forward(self, x):
        x = device.asdevice(x)
        self._x = x                      # pre-activation, as the reference caches (activations.py:14)
        y = device.empty(x.shape)
        C.npm_relu_fwd(x.ptr, y.ptr, x.size, device.stream())
        return y

    def backward(self, dy):
        dy = device.asdevice(dy)
        assert dy.shape == self._x.shape, f'{dy.shape} vs {self._x.shape}'
        dx = device.empty(dy.shape)
        # where(x >= 0, dy, 0): gradient passes at x == 0 (activations.py:19)
        C.npm_relu_bwd(self._x.ptr, dy.ptr, dx.ptr, dy.size, device.stream())
        return dx


class Softmax(Activation):
    def forward(self, x):
        x = device.asdevice(x)
        self._x = x
        y = device.empty(x.shape)
        cols = x.shape[-1]
        C.npm_softmax_fwd(x.ptr, y.ptr, x.size // cols, cols, device.stream())
        self._y = y
        return self._y

    def backward(self, dy, *args, **kwargs):
        dy = device.asdevice(dy)
        assert dy.shape == self._y.shape, f'{dy.shape} vs {self._y.shape}'
        cols = dy.shape[-1]
        dx = device.empty(dy.shape)
        # closed form of the [..., n, n] Jacobian einsum (activations.py:33-45)
        C.npm_softmax_bwd(self._y.ptr, dy.ptr, dx.ptr, dy.size // cols, cols, 1.0, device.stream())
        return dx
