"""Losses — drop-in for the reference's `loss` module (loss.py:10-39)."""
import abc

from layers import layer
from npm_b200 import device
from npm_b200._lib import C


class Loss(layer.Layer):
    # How a data-parallel run must scale the SUM of per-rank gradients so the update equals the
    # single-process one: 1/world for a mean over the local shard, 1 for a plain sum.
    dp_mean = True

    @abc.abstractmethod
    def forward(self, *args, **kwargs) -> float:
        pass

    @abc.abstractmethod
    def backward(self, *args, **kwargs):
        pass


class MSELoss(Loss):
    dp_mean = True   # loss.py:25,29 divide by the (local) y.size

    def forward(self, y, targets):
        y = device.asdevice(y)
        targets = device.asdevice(targets)
        assert y.shape == targets.shape, f'{y.shape} vs {targets.shape}'
        self._y = y
        self._targets = targets
        out = device.empty((1,))
        C.npm_mse_fwd(y.ptr, targets.ptr, out.ptr, y.size, device.stream())
        return device.DeviceScalar(out)

    def backward(self, *args, **kwargs):
        dy = device.empty(self._y.shape)
        C.npm_mse_bwd(self._y.ptr, self._targets.ptr, dy.ptr, dy.size, device.stream())
        return dy


class CrossEntropyLoss(Loss):
    dp_mean = False  # loss.py:36 is an un-normalised sum over the batch

    def forward(self, y, targets):
        y = device.asdevice(y)
        targets = device.asdevice(targets)
        assert y.shape == targets.shape, f'{y.shape} vs {targets.shape}'
        self._y = y
        self._targets = targets
        out = device.empty((1,))
        C.npm_ce_fwd(y.ptr, targets.ptr, out.ptr, y.size, device.stream())
        return device.DeviceScalar(out)

    def backward(self, *args, **kwargs):
        dy = device.empty(self._y.shape)
        C.npm_ce_bwd(self._y.ptr, self._targets.ptr, dy.ptr, dy.size, device.stream())
        return dy
