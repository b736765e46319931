"""Shared helpers of the GPU parity tests (the CUDA path is always reached through the layer mirror →
ctypes → C-ABI; the oracle is only ever the checker)."""
import numpy as np

import optimizer

# north_star tolerances
TC = dict(rtol=1e-3, atol=1e-4)     # TF32-class contractions (run in 3xTF32 unless stated)
EW = dict(rtol=1e-5, atol=1e-5)     # elementwise / normalisation


def close(a, b, **tol):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), **(tol or TC))


class Recorder(optimizer.Optimizer):
    """Gradient-recording optimizer: leaves parameters untouched, keeps what `update` receives."""

    def __init__(self):
        self.grads = {}

    def update_variable(self, identifier, variable, gradient):
        self.grads[identifier] = np.array(np.asarray(gradient))
        return variable


def resolve(layer, path):
    parts = path.split('.')
    obj = layer
    for p in parts[:-1]:
        obj = getattr(obj, p)
    return obj, parts[-1]


def bind(layer, params):
    """Overwrite parameters by dotted attribute path with host arrays (as the reference's tests do)."""
    for path, value in params.items():
        obj, attr = resolve(layer, path)
        setattr(obj, attr, np.array(value, dtype=np.float32))


def grads_of(layer, rec, paths):
    out = {}
    for path in paths:
        obj, attr = resolve(layer, path)
        out[path] = rec.grads[f'{id(obj)}.{attr}']
    return out


def param_values(layer, paths):
    return {p: np.asarray(getattr(*resolve(layer, p))) for p in paths}


def use_device_relu_gates(block, cache, eps=1e-4):
    """A pre-activation within rounding of zero may fall on the other side of ReLU's `x >= 0` gate (activations.py:19)
    on the device than in the float64 oracle, which moves a whole gradient row by O(1).  The backward pass of a
    transformer block is therefore judged on the oracle evaluated with the gates the device actually used — after
    checking that they differ only where |z| is at rounding level (the conv tests do the same)."""
    x2d, z, hdn = cache['ffn']
    gate = ~np.signbit(np.asarray(block._dense1._y)).reshape(z.shape)
    assert np.abs(z[gate != (z >= 0)]).max(initial=0.0) < eps, 'a ReLU gate differs where the pre-activation is not ~0'
    cache['ffn'] = (x2d, np.where(gate, np.abs(z), -np.abs(z) - 1e-30), hdn)
    return cache
