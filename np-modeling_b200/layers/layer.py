"""Layer protocol — drop-in for the reference's layers/layer.py (:11-69).

Identical surface and behaviour: lazy `initialize(*args)` on the first call, `backprop=True`
dispatches to `backward(*args, optimizer_, **kwargs)`, `learning_rate=` is sugar for
`SGDOptimizer(learning_rate)`, passing both raises ValueError.  The only addition: the outermost
backprop call brackets the optimizer so that all parameter updates issued inside it are applied
by one fused multi-tensor kernel when it returns (see optimizer.py).
"""
import abc
from typing import Optional, Sequence

import numpy as np

import optimizer
from npm_b200 import device


class Layer(metaclass=abc.ABCMeta):
    def __init__(self, name: str = '', *args, **kwargs):
        self._name = name
        self._initialized = False

    def initialize(self, *args, **kwargs) -> None:
        pass

    @abc.abstractmethod
    def forward(self, *args, **kwargs):
        pass

    @abc.abstractmethod
    def backward(self, *args, optimizer_, **kwargs):
        pass

    def __call__(self,
                 *args,
                 backprop: bool = False,
                 learning_rate: Optional[float] = None,
                 optimizer_: Optional[optimizer.Optimizer] = None,
                 **kwargs):
        if not self._initialized:
            self.initialize(*args, **kwargs)
            self._initialized = True

        if backprop:
            if learning_rate is not None and optimizer_ is not None:
                raise ValueError(
                    'Optimizer and learning rate cannot both be specified!')
            if learning_rate is not None:
                optimizer_ = optimizer.SGDOptimizer(learning_rate)
            bracket = isinstance(optimizer_, optimizer.Optimizer)
            if bracket:
                optimizer_._enter()
            try:
                return self.backward(*args, optimizer_, **kwargs)
            finally:
                if bracket:
                    optimizer_._exit()
        else:
            return self.forward(*args, **kwargs)

    @property
    def name(self):
        return self._name


class Initializer(metaclass=abc.ABCMeta):
    def __call__(self, shape: Sequence[int]):
        pass


class RandomInitializer(Initializer):
    """clip(N(0,1), -1, 1) float32 drawn from the global legacy np.random stream on the host
    (layer.py:57-60) and uploaded once."""

    def __call__(self, shape: Sequence[int]):
        data = np.random.normal(size=shape).astype(np.float32)
        return device.asdevice(np.minimum(np.maximum(data, -1.0), 1.0))


class StatefulLayer(Layer):
    def __init__(self,
                 initializer: Optional[Initializer] = None,
                 *args,
                 **kwargs):
        super().__init__(*args, **kwargs)
        self._initializer = initializer or RandomInitializer()

    def _p(self, attr: str):
        """Parameter `attr` as a DeviceArray (NumPy arrays assigned by the user are uploaded once)."""
        v = getattr(self, attr)
        if not isinstance(v, device.DeviceArray):
            v = device.asdevice(v)
            setattr(self, attr, v)
        return v
