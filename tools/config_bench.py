"""tools/config_bench.py — step time of BASELINE.json's configs 1-4 (SURVEY.md §8d shapes) on one B200.
cfg5 is bench.py's workload.  usage: python tools/config_bench.py [bf16x3|tf32|3xtf32]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
import npm_b200  # noqa: E402
import loss  # noqa: E402
import optimizer  # noqa: E402
from layers import Conv2D, Dense, LayerNormalization, MultiHeadAttention, Softmax  # noqa: E402
from layers.adapters import EncoderStack  # noqa: E402
from npm_b200 import device  # noqa: E402
from train import Trainer, iter_parameters  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16x3'
print(f'# contraction mode {prec}', flush=True)
npm_b200.set_precision(prec)
rng = np.random.default_rng(0)


def scale_weights(layers_):
    for owner, name in iter_parameters(layers_):
        v = np.asarray(getattr(owner, name))
        if name in ('_w', '_wq', '_wk', '_wv', '_wo'):
            fan = v.shape[-1] if name in ('_wq', '_wk', '_wv') else (v.shape[1] * v.shape[2] if name == '_wo' else int(np.prod(v.shape[:-1])))
            setattr(owner, name, (v / np.sqrt(fan)).astype(np.float32))
        elif name == '_gamma':
            setattr(owner, name, np.ones_like(v))
        elif name in ('_beta',):
            setattr(owner, name, np.zeros_like(v))


def timeit(name, layers_, loss_, opt, x, t, steps, flop=None, tokens=None, cuda_graph=False):
    tr = Trainer(layers_, loss_, verbose=False, cuda_graph=cuda_graph)
    xd = device.asdevice(x) if not isinstance(x, tuple) else tuple(device.asdevice(a) for a in x)
    td = device.asdevice(t)
    tr._forward(xd)                     # lazy init only (an update with the unscaled reference init would overflow)
    scale_weights(layers_)
    for _ in range(3):
        tr.train(xd, td, 1, opt)
    if cuda_graph:
        tr.train(xd, td, 6, opt)        # captures (and instantiates) the step once; the timed call below replays it
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.train(xd, td, steps, opt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    extra = ''
    if flop:
        extra += f'  {flop / ms / 1e9:7.1f} TFLOP/s model flops'
    if tokens:
        extra += f'  {tokens / ms * 1e3 / 1e3:9.1f} k tokens/s'
    print(f'{name:58s} {ms:9.3f} ms/step  loss {float(tr.last_loss):.4g}{extra}', flush=True)


# cfg1: MLP 784 -> 256 -> 10, batch 64, CE, SGD
x = rng.standard_normal((64, 784)).astype(np.float32)
t = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 64)]
timeit('cfg1 MLP 784-256-10 b64 CE SGD', [Dense(256), Dense(10, activation=Softmax())], loss.CrossEntropyLoss(),
       optimizer.SGDOptimizer(1e-4), x, t, 50, flop=78.1e6)
timeit('cfg1 (same, step replayed as a CUDA graph)', [Dense(256), Dense(10, activation=Softmax())], loss.CrossEntropyLoss(),
       optimizer.SGDOptimizer(1e-4), x, t, 50, flop=78.1e6, cuda_graph=True)
# cfg2: conv stack
x = rng.standard_normal((256, 32, 32, 3)).astype(np.float32)
t = rng.standard_normal((256, 32, 32, 128)).astype(np.float32)
timeit('cfg2 Conv2D 3->64->128 k3 b256 MSE Adam', [Conv2D(64, 3), Conv2D(128, 3)], loss.MSELoss(),
       optimizer.AdamOptimizer(1e-3), x, t, 10, flop=118.7e9)
# cfg3: MHA + LN
x = rng.standard_normal((32, 1024, 768)).astype(np.float32)
t = rng.standard_normal((32, 1024, 768)).astype(np.float32)
timeit('cfg3 MHA(12) + LayerNorm b32 s1024 d768 MSE Adam', [MultiHeadAttention(12), LayerNormalization()], loss.MSELoss(),
       optimizer.AdamOptimizer(1e-4), x, t, 10, flop=773.1e9, tokens=32 * 1024)
# cfg4: 12 encoder layers
x = rng.standard_normal((32, 512, 768)).astype(np.float32)
t = rng.standard_normal((32, 512, 768)).astype(np.float32)
timeit('cfg4 12 x TransformerEncoder(12, 3072, pre, 0.1) b32 s512', [EncoderStack(12, 12, 3072, True, 0.1)], loss.MSELoss(),
       optimizer.AdamOptimizer(1e-4), x, t, 5, flop=9.277e12, tokens=32 * 512)
