"""oracle/make_golden.py — generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (the reference is at /root/reference there; it does not exist on the
GPU box):  python oracle/make_golden.py
It imports the reference's own `layers`, `optimizer`, `loss`, `train` modules with /root/reference
first on sys.path (never the mirrors under np-modeling_b200/), drives each layer on seeded inputs,
records outputs, the gradients every `optimizer_.update` receives, and the parameters after real
SGD / Adam steps, and stores them as small .npz fixtures.  tests/test_oracle.py pins
oracle/np_oracle.py to these; the GPU parity tests compare the CUDA path against them too.
"""
import contextlib
import copy
import io
import os
import sys

REF = os.environ.get('NPM_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

import loss as ref_loss  # noqa: E402
import optimizer as ref_optimizer  # noqa: E402
import train as ref_train  # noqa: E402
from layers import (Conv2D, Dense, LayerNormalization, Linear, MultiHeadAttention, ReLU, Softmax,  # noqa: E402
                    TransformerDecoder, TransformerEncoder)
from layers.normalizations import DropOut  # noqa: E402

assert ref_optimizer.__file__.startswith(REF), ref_optimizer.__file__

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


class Recorder(ref_optimizer.Optimizer):
    """Leaves variables untouched and keeps every gradient, keyed by (object id, attribute)."""

    def __init__(self):
        self.grads = {}

    def update_variable(self, identifier, variable, gradient):
        self.grads[identifier] = np.array(gradient)
        return variable


def rand(*shape, scale=1.0):
    return (np.random.normal(size=shape) * scale).astype(np.float32)


def named_params(obj, prefix=''):
    """{dotted attribute path: (owner, attr)} for every ndarray parameter below a reference layer."""
    out = {}
    for name, value in vars(obj).items():
        if isinstance(value, np.ndarray) and name in ('_w', '_b', '_gamma', '_beta', '_wq', '_wk', '_wv', '_wo',
                                                        '_bq', '_bk', '_bv', '_bo'):
            out[prefix + name] = (obj, name)
        elif hasattr(value, 'forward') and hasattr(value, '_initialized'):
            out.update(named_params(value, prefix + name + '.'))
    return out


def grads_by_name(layer, rec):
    return {path: rec.grads[f'{id(o)}.{a}'] for path, (o, a) in named_params(layer).items()
            if f'{id(o)}.{a}' in rec.grads}


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **{k: np.asarray(v) for k, v in arrays.items()})
    print(f'{name}: {len(arrays)} arrays')


def with_prefix(prefix, d):
    return {prefix + k: v for k, v in d.items()}


def case_dense():
    np.random.seed(0)
    x, dy = rand(64, 32), rand(64, 16)
    layer = Dense(16)
    y = layer(x)
    params = {k: getattr(o, a).copy() for k, (o, a) in named_params(layer).items()}
    rec = Recorder()
    dx = layer(dy, backprop=True, optimizer_=rec)
    g = grads_by_name(layer, rec)
    sgd = copy.deepcopy(layer)
    sgd(x)
    sgd(dy, backprop=True, learning_rate=0.05)
    adam = copy.deepcopy(layer)
    opt = ref_optimizer.AdamOptimizer(learning_rate=0.01)
    for _ in range(3):
        adam(x)
        adam(dy, backprop=True, optimizer_=opt)
    save('dense', x=x, dy=dy, y=y, dx=dx, **with_prefix('p.', params), **with_prefix('g.', g),
         **with_prefix('sgd.', {k: getattr(o, a) for k, (o, a) in named_params(sgd).items()}),
         **with_prefix('adam3.', {k: getattr(o, a) for k, (o, a) in named_params(adam).items()}))


def case_activations():
    np.random.seed(1)
    x, dy = rand(16, 32), rand(16, 32)
    sm = Softmax()
    y = sm(x)
    dx = sm(dy, backprop=True)
    r = ReLU()
    xr = rand(8, 16)
    xr[0, :4] = 0.0   # the x == 0 branch of `x >= 0`
    dyr = rand(8, 16)
    yr = r(xr)
    dxr = r.backward(dyr)
    save('activations', sm_x=x, sm_dy=dy, sm_y=y, sm_dx=dx, relu_x=xr, relu_y=yr, relu_dy=dyr, relu_dx=dxr)


def case_layernorm():
    np.random.seed(2)
    x, dz = rand(32, 128), rand(32, 128)
    layer = LayerNormalization()
    z = layer(x)
    params = {k: getattr(o, a).copy() for k, (o, a) in named_params(layer).items()}
    rec = Recorder()
    dx = layer(dz, backprop=True, optimizer_=rec)
    x3, dz3 = rand(2, 5, 24), rand(2, 5, 24)
    layer3 = LayerNormalization(epsilon=1e-6)
    z3 = layer3(x3)
    p3 = {k: getattr(o, a).copy() for k, (o, a) in named_params(layer3).items()}
    rec3 = Recorder()
    dx3 = layer3(dz3, backprop=True, optimizer_=rec3)
    save('layernorm', x=x, dz=dz, z=z, dx=dx, **with_prefix('p.', params), **with_prefix('g.', grads_by_name(layer, rec)),
         x3=x3, dz3=dz3, z3=z3, dx3=dx3, **with_prefix('p3.', p3), **with_prefix('g3.', grads_by_name(layer3, rec3)))


def case_dropout():
    np.random.seed(3)
    x, dy = rand(128, 32), rand(128, 32)
    layer = DropOut(0.5)
    y = layer(x)
    dx = layer(dy, backprop=True)
    layer2 = DropOut(0.1)
    y2 = layer2(x)
    save('dropout', x=x, dy=dy, mask=layer._mask, y=y, dx=dx, mask2=layer2._mask, y2=y2)


def case_mha():
    for tag, skv in (('self', None), ('cross', 12)):
        np.random.seed(4)
        b, sq, h, d = 2, 8, 2, 16
        query = rand(b, sq, d)
        kv = None if skv is None else rand(b, skv, d)
        dy = rand(b, sq, d)
        layer = MultiHeadAttention(h)
        out = layer(query) if kv is None else layer(query, kv)
        # re-scale the unscaled +-1 init so softmax is not saturated, then recompute
        for o, a in named_params(layer).values():
            if a.startswith('_w'):
                setattr(o, a, (getattr(o, a) * 0.25).astype(np.float32))
        out = layer(query) if kv is None else layer(query, kv)
        params = {k: getattr(o, a).copy() for k, (o, a) in named_params(layer).items()}
        rec = Recorder()
        dq, dk, dv = layer(dy, backprop=True, optimizer_=rec)
        arrays = dict(query=query, dy=dy, out=out, dquery=dq, dkey=dk, dvalue=dv, scores=layer._attention_scores,
                      **with_prefix('p.', params), **with_prefix('g.', grads_by_name(layer, rec)))
        if kv is not None:
            arrays['kv'] = kv
        save('mha_' + tag, **arrays)


def _rescale(layer, w_scale=0.25):
    for path, (o, a) in named_params(layer).items():
        if a.startswith('_w'):
            setattr(o, a, (getattr(o, a) * w_scale).astype(np.float32))


def case_transformer():
    b, sq, skv, h, d, f = 2, 8, 12, 2, 16, 32
    for kind in ('encoder', 'decoder'):
        for norm_first in (True, False):
            for drop in (0.0, 0.25):
                np.random.seed(5)
                q, kv, dy = rand(b, sq, d), rand(b, skv, d), rand(b, sq, d)
                cls = TransformerEncoder if kind == 'encoder' else TransformerDecoder
                layer = cls(h, f, norm_first, drop)
                args = (q,) if kind == 'encoder' else (q, kv)
                layer(*args)
                _rescale(layer)
                np.random.seed(55)     # dropout masks of the recorded forward
                out = layer(*args)
                params = {k: getattr(o, a).copy() for k, (o, a) in named_params(layer).items()}
                masks = {}
                if drop:
                    for i in (1, 2, 3):
                        dl = getattr(layer, f'_dropout{i}', None)
                        if dl is not None:
                            masks[f'mask{i}'] = dl._mask
                rec = Recorder()
                res = layer(dy, backprop=True, optimizer_=rec)
                arrays = dict(q=q, dy=dy, out=out, **masks, **with_prefix('p.', params),
                              **with_prefix('g.', grads_by_name(layer, rec)))
                if kind == 'encoder':
                    arrays['dq'] = res
                else:
                    arrays.update(kv=kv, dq=res[0], dkv=res[1])
                save(f'{kind}_{"pre" if norm_first else "post"}_{"drop" if drop else "nodrop"}', **arrays)


def case_conv():
    for tag, (n, hh, ww, c0, c1, k) in dict(c3=(2, 6, 5, 3, 4, 3), c8=(2, 7, 6, 8, 8, 3), k5=(1, 6, 6, 4, 4, 5),
                                            k1=(2, 4, 4, 8, 12, 1)).items():
        np.random.seed(6)
        x, dy = rand(n, hh, ww, c0), rand(n, hh, ww, c1)
        layer = Conv2D(c1, k)
        y = layer(x)
        params = {kk: getattr(o, a).copy() for kk, (o, a) in named_params(layer).items()}
        rec = Recorder()
        dx = layer(dy, backprop=True, optimizer_=rec)
        save('conv_' + tag, x=x, dy=dy, y=y, dx=dx, **with_prefix('p.', params),
             **with_prefix('g.', grads_by_name(layer, rec)))


def case_loss():
    np.random.seed(2024)
    y, t = rand(16, 8), rand(16, 8)
    mse = ref_loss.MSELoss()
    lm = mse(y, t)
    dm = mse(backprop=True)
    prob = np.exp(y) / np.exp(y).sum(axis=-1, keepdims=True)
    onehot = np.eye(8, dtype=np.float32)[np.random.randint(0, 8, size=16)]
    ce = ref_loss.CrossEntropyLoss()
    lc = ce(prob, onehot)
    dc = ce(backprop=True)
    save('loss', y=y, t=t, mse=lm, mse_dy=dm, prob=prob, onehot=onehot, ce=lc, ce_dy=dc)


def case_optimizer():
    np.random.seed(7)
    w0 = rand(7, 5)
    gs = [rand(7, 5) for _ in range(4)]

    class Box:
        pass

    box = Box()
    box.w = w0.copy()
    sgd = ref_optimizer.SGDOptimizer(0.1)
    sgd.update(box, 'w', gs[0])
    w_sgd = box.w.copy()
    box.w = w0.copy()
    adam = ref_optimizer.AdamOptimizer(learning_rate=0.01)
    traj = []
    for g in gs:
        adam.update(box, 'w', g)
        traj.append(box.w.copy())
    save('optimizer', w0=w0, grads=np.stack(gs), sgd=w_sgd, adam=np.stack(traj))


def case_trainer():
    np.random.seed(8)
    x = rand(16, 12)
    t = np.eye(4, dtype=np.float32)[np.random.randint(0, 4, size=16)]
    layers = [Dense(8), Dense(4, activation=Softmax())]
    trainer = ref_train.Trainer(layers, ref_loss.CrossEntropyLoss())
    with contextlib.redirect_stdout(io.StringIO()):
        trainer.eval(x, t)
    for layer in layers:
        _rescale(layer, 0.3)
    p0 = {}
    for i, layer in enumerate(layers):
        p0.update({f'{i}.{k}': getattr(o, a).copy() for k, (o, a) in named_params(layer).items()})
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        trainer.train(x, t, 4, ref_optimizer.SGDOptimizer(1e-2))
    losses = [float(line.split()[-1]) for line in buf.getvalue().splitlines() if line.startswith('Loss')]
    p1 = {}
    for i, layer in enumerate(layers):
        p1.update({f'{i}.{k}': getattr(o, a).copy() for k, (o, a) in named_params(layer).items()})
    save('trainer_mlp', x=x, t=t, losses=np.array(losses), **with_prefix('p0.', p0), **with_prefix('p1.', p1))


def _train_case(name, layers, x, t, steps, lr, w_scale):
    """The reference's own Trainer + AdamOptimizer for a few steps: losses printed per step and every parameter
    before / after (end-to-end pin of forward, backward, in-layer updates and Adam state on composite layers)."""
    trainer = ref_train.Trainer(layers, ref_loss.MSELoss())
    with contextlib.redirect_stdout(io.StringIO()):
        trainer.eval(x, t)                      # lazy initialisation
    for layer in layers:
        _rescale(layer, w_scale)
    p0 = {}
    for i, layer in enumerate(layers):
        p0.update({f'{i}.{k}': getattr(o, a).copy() for k, (o, a) in named_params(layer).items()})
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        trainer.train(x, t, steps, ref_optimizer.AdamOptimizer(learning_rate=lr))
    losses = [float(line.split()[-1]) for line in buf.getvalue().splitlines() if line.startswith('Loss')]
    p1 = {}
    for i, layer in enumerate(layers):
        p1.update({f'{i}.{k}': getattr(o, a).copy() for k, (o, a) in named_params(layer).items()})
    save(name, x=x, t=t, losses=np.array(losses), **with_prefix('p0.', p0), **with_prefix('p1.', p1))


def case_trainer_conv():
    np.random.seed(9)
    x, t = rand(2, 6, 6, 3), rand(2, 6, 6, 6)
    _train_case('trainer_conv', [Conv2D(4, 3), Conv2D(6, 3)], x, t, 3, 1e-2, 0.3)


def case_trainer_encoder():
    np.random.seed(10)
    x, t = rand(2, 8, 16), rand(2, 8, 16)
    layers = [TransformerEncoder(2, 32, True, 0.0), TransformerEncoder(2, 32, False, 0.0)]
    _train_case('trainer_encoder', layers, x, t, 3, 1e-2, 0.25)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'trainers':      # only the end-to-end Trainer fixtures added later
        case_trainer_conv()
        case_trainer_encoder()
        sys.exit(0)
    for fn in (case_dense, case_activations, case_layernorm, case_dropout, case_mha, case_transformer, case_conv,
               case_loss, case_optimizer, case_trainer, case_trainer_conv, case_trainer_encoder):
        fn()
