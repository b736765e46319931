"""GPU parity: every layer of the hot path, driven through the reference's own Layer API, against the
golden vectors produced by the unmodified reference (tests/golden) and against the oracle on fresh
seeded inputs.  Tolerances are north_star's: rtol 1e-3 / atol 1e-4 for tensor-core contractions
(3xTF32), 1e-5 elementwise, masks bit-exact."""
import copy

import numpy as np
import pytest

from conftest import load_golden, sub
from helpers import EW, TC, Recorder, bind, close, grads_of, param_values, use_device_relu_gates

pytestmark = pytest.mark.gpu


# Every test of this file runs in BOTH contraction modes that claim north_star's tolerance: 'bf16x3' (the mode bench.py
# reports: split-bf16 tcgen05 GEMM, fused attention, tensor-core conv) and '3xtf32'.
@pytest.fixture(autouse=True, params=['bf16x3', '3xtf32'])
def _precision(request):
    import npm_b200
    npm_b200.set_precision(request.param)
    yield request.param
    npm_b200.set_precision('bf16x3')


def test_library_loaded_and_device():
    import ctypes
    from npm_b200 import _lib
    lib = _lib.load()
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.npm_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)) == 0
    assert major.value == 10, 'these kernels are sm_100a only'
    assert sm.value >= 100


# ------------------------------------------------------------------ Dense / Linear
def test_dense_golden():
    import optimizer
    from layers import Dense
    g = load_golden('dense')
    layer = Dense(16)
    layer(g['x'])
    bind(layer, sub(g, 'p.'))
    close(layer(g['x']), g['y'])
    rec = Recorder()
    dx = layer(g['dy'], backprop=True, optimizer_=rec)
    close(dx, g['dx'])
    got = grads_of(layer, rec, sub(g, 'g.').keys())
    for k, v in sub(g, 'g.').items():
        close(got[k], v)
    # parameters are untouched by the recorder
    close(layer.linear.w, g['p._linear._w'], rtol=0, atol=0)

    # real SGD step through `learning_rate=` sugar; an alias captured before sees the update
    sgd = Dense(16)
    sgd(g['x'])
    bind(sgd, sub(g, 'p.'))
    sgd(g['x'])
    w_alias = sgd.linear.w
    sgd(g['dy'], backprop=True, learning_rate=0.05)
    close(w_alias, g['sgd._linear._w'])
    close(sgd.linear.b, g['sgd._linear._b'])

    adam = Dense(16)
    adam(g['x'])
    bind(adam, sub(g, 'p.'))
    opt = optimizer.AdamOptimizer(learning_rate=0.01)
    for _ in range(3):
        adam(g['x'])
        adam(g['dy'], backprop=True, optimizer_=opt)
    close(adam.linear.w, g['adam3._linear._w'], rtol=1e-3, atol=2e-4)
    close(adam.linear.b, g['adam3._linear._b'], rtol=1e-3, atol=2e-4)


def test_layer_protocol_errors():
    import optimizer
    from layers import Dense
    layer = Dense(4)
    x = np.ones((3, 5), np.float32)
    layer(x)
    with pytest.raises(ValueError):
        layer(np.ones((3, 4), np.float32), backprop=True, learning_rate=0.1, optimizer_=optimizer.SGDOptimizer(0.1))
    with pytest.raises(AssertionError):
        layer(np.ones((2, 4), np.float32), backprop=True, learning_rate=0.1)   # wrong batch


@pytest.mark.parametrize('m,k,n', [(64, 784, 256), (64, 256, 10), (513, 100, 36), (1, 8, 4), (300, 1024, 520)])
def test_linear_vs_oracle_shapes(m, k, n):
    from layers import Linear
    from oracle import np_oracle as O
    rng = np.random.default_rng(m * 7 + n)
    x = rng.standard_normal((m, k), dtype=np.float32)
    dy = rng.standard_normal((m, n), dtype=np.float32)
    w = (rng.standard_normal((k, n)) / np.sqrt(k)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    layer = Linear(n)
    layer(x)
    bind(layer, {'_w': w, '_b': b})
    close(layer(x), O.linear_fwd(x, w, b))
    rec = Recorder()
    dx = layer(dy, backprop=True, optimizer_=rec)
    odx, odw, odb = O.linear_bwd(x, w, dy)
    close(dx, odx)
    g = grads_of(layer, rec, ['_w', '_b'])
    close(g['_w'], odw, rtol=1e-3, atol=1e-4 * np.sqrt(m))
    close(g['_b'], odb, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('prec', ['bf16x3', '3xtf32', 'tf32', 'fp32', 'bf16'])
@pytest.mark.parametrize('m,k,n', [(256, 128, 128), (513, 100, 36), (1024, 1024, 1024), (1, 8, 4), (300, 64, 520)])
def test_linear_residual_epilogue(prec, m, k, n):
    """`out = dense2(out); out += skip` (transformer.py:52-53) with the add run in the GEMM epilogue."""
    import npm_b200
    from layers import Linear
    from oracle import np_oracle as O
    npm_b200.set_precision(prec)
    rng = np.random.default_rng(m + 3 * n)
    x = rng.standard_normal((m, k), dtype=np.float32)
    skip = rng.standard_normal((m, n), dtype=np.float32)
    w = (rng.standard_normal((k, n)) / np.sqrt(k)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    layer = Linear(n)
    layer(x)
    bind(layer, {'_w': w, '_b': b})
    from npm_b200 import device
    skip_d = device.asdevice(skip)
    want = O.linear_fwd(x, w, b) + skip
    # one-pass TF32: 10-bit operand mantissas; one-pass bf16: 8-bit
    tol = dict(rtol=2e-3, atol=5e-3) if prec == 'tf32' else dict(rtol=2e-2, atol=5e-2) if prec == 'bf16' else TC
    close(layer(x, _residual=skip_d), want, **tol)
    close(skip_d, skip, rtol=0, atol=0)           # the skip branch is read, never written


# ------------------------------------------------------------------ activations / norm / dropout
def test_activations_golden():
    from layers import ReLU, Softmax
    g = load_golden('activations')
    sm = Softmax()
    close(sm(g['sm_x']), g['sm_y'], **EW)
    close(sm(g['sm_dy'], backprop=True), g['sm_dx'], **EW)
    r = ReLU()
    close(r(g['relu_x']), g['relu_y'], rtol=0, atol=0)
    close(r.backward(g['relu_dy']), g['relu_dx'], rtol=0, atol=0)     # includes the x == 0 branch


@pytest.mark.parametrize('rows,cols', [(64, 10), (33, 1000), (128, 1024), (7, 4100), (5, 3)])
def test_softmax_shapes(rows, cols):
    from layers import Softmax
    from oracle import np_oracle as O
    rng = np.random.default_rng(rows + cols)
    x = (rng.standard_normal((rows, cols)) * 3).astype(np.float32)
    dy = rng.standard_normal((rows, cols)).astype(np.float32)
    sm = Softmax()
    y = sm(x)
    close(y, O.softmax_fwd(x), **EW)
    close(np.asarray(y).sum(-1), np.ones(rows), rtol=1e-5, atol=1e-5)
    close(sm(dy, backprop=True), O.softmax_bwd(O.softmax_fwd(x), dy), **EW)


def test_layernorm_golden():
    from layers import LayerNormalization
    g = load_golden('layernorm')
    for sfx, eps in (('', 1e-3), ('3', 1e-6)):
        layer = LayerNormalization(epsilon=eps)
        layer(g['x' + sfx])
        bind(layer, sub(g, f'p{sfx}.'))
        close(layer(g['x' + sfx]), g['z' + sfx], **EW)
        rec = Recorder()
        close(layer(g['dz' + sfx], backprop=True, optimizer_=rec), g['dx' + sfx], rtol=1e-4, atol=1e-5)
        got = grads_of(layer, rec, ['_gamma', '_beta'])
        close(got['_gamma'], g[f'g{sfx}._gamma'], rtol=1e-4, atol=1e-4)
        close(got['_beta'], g[f'g{sfx}._beta'], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('rows,cols', [(4096, 1024), (1000, 768), (50, 36), (3, 2050), (17, 7)])
def test_layernorm_shapes(rows, cols):
    from layers import LayerNormalization
    from oracle import np_oracle as O
    rng = np.random.default_rng(rows * 3 + cols)
    x = (rng.standard_normal((rows, cols)) * 2 + 0.5).astype(np.float32)
    dz = rng.standard_normal((rows, cols)).astype(np.float32)
    gamma = rng.standard_normal(cols).astype(np.float32)
    beta = rng.standard_normal(cols).astype(np.float32)
    layer = LayerNormalization()
    layer(x)
    bind(layer, {'_gamma': gamma, '_beta': beta})
    close(layer(x), O.layernorm_fwd(x, gamma, beta)[0], rtol=1e-4, atol=1e-5)
    rec = Recorder()
    dx = layer(dz, backprop=True, optimizer_=rec)
    odx, odg, odb = O.layernorm_bwd(x, gamma, dz)
    close(dx, odx, rtol=1e-4, atol=1e-5)
    got = grads_of(layer, rec, ['_gamma', '_beta'])
    close(got['_gamma'], odg, rtol=1e-4, atol=1e-5 * np.sqrt(rows) * 4)
    close(got['_beta'], odb, rtol=1e-4, atol=1e-5 * np.sqrt(rows) * 4)


def test_dropout_mask_injection_golden():
    from layers.normalizations import DropOut
    g = load_golden('dropout')
    layer = DropOut(0.5)
    layer._mask = g['mask']
    close(layer(g['x']), g['y'], rtol=0, atol=0)
    close(layer(g['dy'], backprop=True), g['dx'], rtol=0, atol=0)
    assert np.array_equal(layer._mask, g['mask'])
    layer2 = DropOut(0.1)
    layer2._mask = g['mask2']
    close(layer2(g['x']), g['y2'], rtol=0, atol=0)   # x / float32(0.9), correctly rounded division


@pytest.mark.parametrize('n,offset', [(4096, 0), (1001, 0), (777, 5), (64, 2 ** 33 + 3)])
def test_dropout_philox_bit_exact(n, offset):
    """The GPU mask equals the CPU Philox restatement bit for bit for a given (seed, offset)."""
    import torch
    from npm_b200 import device
    from npm_b200._lib import C
    from oracle import philox
    seed = 0x1234ABCD5678EF01
    keep = np.float32(0.9)
    x = np.random.default_rng(0).standard_normal(n).astype(np.float32)
    xd = device.asdevice(x)
    yd = device.empty((n,))
    C.npm_dropout_fwd(xd.ptr, yd.ptr, n, keep, seed, offset, None, device.stream())
    mask = philox.dropout_mask(n, keep, seed, offset)
    want = np.where(mask != 0, x / keep, np.float32(0)).astype(np.float32)
    assert np.array_equal(np.asarray(yd), want)
    md = torch.empty(n, dtype=torch.uint8, device='cuda')
    C.npm_dropout_mask(md.data_ptr(), n, keep, seed, offset, device.stream())
    assert np.array_equal(md.cpu().numpy(), mask)
    # backward applies the same mask
    dd = device.empty((n,))
    C.npm_dropout_bwd(xd.ptr, dd.ptr, n, keep, seed, offset, None, device.stream())
    assert np.array_equal(np.asarray(dd), want)


def test_dropout_layer_semantics():
    from layers.normalizations import DropOut, set_dropout_seed
    x = np.random.default_rng(1).standard_normal((64, 48)).astype(np.float32)
    set_dropout_seed(42)
    a = DropOut(0.25)
    ya = np.asarray(a(x))
    set_dropout_seed(42)
    b = DropOut(0.25)
    yb = np.asarray(b(x))
    assert np.array_equal(ya, yb)                           # same seed → same mask
    yc = np.asarray(b(x))
    assert not np.array_equal(yb, yc)                       # the stream advances between calls
    m = a._mask
    assert m.dtype == np.int64 and m.shape == x.shape and 0.6 < m.mean() < 0.9
    assert np.array_equal(ya != 0, m != 0)
    dy = np.ones_like(x)
    assert np.array_equal(np.asarray(a(dy, backprop=True)) != 0, m != 0)
    ident = DropOut(0.0)
    assert ident(x) is x                                    # identity when drop_prob == 0 (:14,23)
    assert np.array_equal(np.asarray(DropOut(0.5).forward(x, training=False)), x)


# ------------------------------------------------------------------ attention / transformer
@pytest.mark.parametrize('tag', ['self', 'cross'])
def test_mha_golden(tag):
    from layers import MultiHeadAttention
    from oracle import np_oracle as O
    g = load_golden('mha_' + tag)
    layer = MultiHeadAttention(2)
    args = (g['query'],) if tag == 'self' else (g['query'], g['kv'])
    layer(*args)
    bind(layer, sub(g, 'p.'))
    close(layer(*args), g['out'])
    close(layer._attention_scores, g['scores'], rtol=1e-3, atol=1e-5)
    rec = Recorder()
    dq, dk, dv = layer(g['dy'], backprop=True, optimizer_=rec)
    close(dq, g['dquery']); close(dk, g['dkey']); close(dv, g['dvalue'])
    got = grads_of(layer, rec, O.MHA_PARAMS)
    for k in O.MHA_PARAMS:
        assert got[k].shape == g['g.' + k].shape
        close(got[k], g['g.' + k])
    assert copy.deepcopy(layer) is not layer                # attentions_test.py:72 deep-copies layers


@pytest.mark.parametrize('prec', ['bf16x3', '3xtf32', 'tf32'])
@pytest.mark.parametrize('mode', ['qkv', 'kv', 'separate'])
@pytest.mark.parametrize('dims', [(2, 40, 56, 4, 64), (1, 130, 130, 2, 64), (2, 9, 17, 3, 8)])
def test_mha_packed_projection_modes(prec, mode, dims):
    """Self-attention runs q|k|v as one packed projection GEMM, key-is-value cross-attention k|v; three distinct
    inputs take three GEMMs.  All must equal the reference algorithm (attentions.py:67-199), through the public
    3-tuple API and through the summed form the transformer blocks use (transformer.py:85,184-185,196)."""
    import npm_b200
    from layers import MultiHeadAttention
    from oracle import np_oracle as O
    npm_b200.set_precision(prec)
    b, sq, skv, h, d = dims
    if mode == 'qkv':
        skv = sq
    rng = np.random.default_rng(sq * 31 + d)
    dm = h * d
    query = rng.standard_normal((b, sq, dm), dtype=np.float32)
    key = rng.standard_normal((b, skv, dm), dtype=np.float32)
    value = rng.standard_normal((b, skv, dm), dtype=np.float32)
    args = {'qkv': (query,), 'kv': (query, key), 'separate': (query, key, value)}[mode]
    dy = rng.standard_normal((b, sq, dm), dtype=np.float32)
    layer = MultiHeadAttention(h)
    layer(*args)
    params = {k: (np.asarray(getattr(layer, k)) * (1.0 / np.sqrt(dm) if k.startswith('_w') else 0.1)).astype(np.float32)
              for k in O.MHA_PARAMS}
    bind(layer, params)
    tol = dict(rtol=2e-3, atol=2e-2) if prec == 'tf32' else TC
    want, cache = O.mha_fwd(params, *args)
    close(layer(*args), want, **tol)
    assert layer._mode == mode
    (dq_, dk_, dv_), grads = O.mha_bwd(params, cache, dy)
    gtol = dict(rtol=2e-3, atol=2e-2 * np.sqrt(b * sq)) if prec == 'tf32' else dict(rtol=1e-3, atol=1e-4 * np.sqrt(b * sq))
    for summed in (False, True):
        rec = Recorder()
        clone = copy.deepcopy(layer)                        # the copy re-packs its parameters on first use
        close(clone(*args), want, **tol)
        if summed:
            if mode == 'separate':
                continue
            got = clone.backward(dy, rec, _sum_inputs=True)
            if mode == 'qkv':
                assert got[1] is None
                close(got[0], dq_ + dk_ + dv_, **tol)
            else:
                close(got[0], dq_, **tol)
                close(got[1], dk_ + dv_, **tol)
        else:
            got = clone(dy, backprop=True, optimizer_=rec)
            close(got[0], dq_, **tol); close(got[1], dk_, **tol); close(got[2], dv_, **tol)
        gg = grads_of(clone, rec, O.MHA_PARAMS)
        for k in O.MHA_PARAMS:
            assert gg[k].shape == grads[k].shape
            close(gg[k], grads[k], **gtol)
    # a real optimizer step through the packed gradient arena updates the three slices in place
    import optimizer
    opt = optimizer.SGDOptimizer(0.1)
    alias = layer._wk
    layer(*args)
    layer(dy, backprop=True, optimizer_=opt)
    close(alias, params['_wk'] - 0.1 * grads['_wk'], **gtol)
    close(layer._bv, params['_bv'] - 0.1 * grads['_bv'], **gtol)


@pytest.mark.parametrize('mode', ['qkv', 'kv'])
@pytest.mark.parametrize('dims', [(2, 200, 136, 2, 64), (1, 384, 384, 3, 64)])
def test_mha_backward_with_dO_planes_and_row_dots_from_the_gemm(mode, dims, monkeypatch):
    """Opt-in route (layers/attentions.py `_NO_DO_ROWDOT`): the output projection's dX GEMM writes dO as bf16 planes and
    D = rowsum(dO o O) per head from its epilogue (npm_gemm_desc.rowdot_*, npm_mha_strides.do_ready) instead of fp32 dO
    plus the D / split kernel.  Same gradients as the reference algorithm (attentions.py:129-188)."""
    import npm_b200
    from layers import MultiHeadAttention, attentions
    from npm_b200._lib import C
    from oracle import np_oracle as O
    npm_b200.set_precision('bf16x3')
    monkeypatch.setattr(attentions, '_NO_DO_ROWDOT', False)
    b, sq, skv, h, d = dims
    if mode == 'qkv':
        skv = sq
    rng = np.random.default_rng(sq + 7 * d)
    dm = h * d
    query = rng.standard_normal((b, sq, dm), dtype=np.float32)
    key = rng.standard_normal((b, skv, dm), dtype=np.float32)
    args = {'qkv': (query,), 'kv': (query, key)}[mode]
    dy = rng.standard_normal((b, sq, dm), dtype=np.float32)
    layer = MultiHeadAttention(h)
    layer(*args)
    params = {k: (np.asarray(getattr(layer, k)) * (1.0 / np.sqrt(dm) if k.startswith('_w') else 0.1)).astype(np.float32)
              for k in O.MHA_PARAMS}
    bind(layer, params)
    want, cache = O.mha_fwd(params, *args)
    close(layer(*args), want, **TC)
    assert layer._path == 2, 'the fused split-bf16 path serves head dim 64'
    (dq_, dk_, dv_), grads = O.mha_bwd(params, cache, dy)
    rec = Recorder()
    launches = C.npm_launch_count()
    got = layer(dy, backprop=True, optimizer_=rec)
    used = C.npm_launch_count() - launches
    close(got[0], dq_, **TC); close(got[1], dk_, **TC); close(got[2], dv_, **TC)
    gg = grads_of(layer, rec, O.MHA_PARAMS)
    for k in O.MHA_PARAMS:
        close(gg[k], grads[k], rtol=1e-3, atol=1e-4 * np.sqrt(b * sq))
    # the same backward through the default route launches one kernel more (the D / dO-split pass)
    monkeypatch.setattr(attentions, '_NO_DO_ROWDOT', True)
    layer(*args)
    launches = C.npm_launch_count()
    layer(dy, backprop=True, optimizer_=Recorder())
    assert C.npm_launch_count() - launches == used + 1


@pytest.mark.parametrize('prec', ['bf16x3', '3xtf32', 'tf32'])
@pytest.mark.parametrize('dims', [(2, 40, 4, 64), (1, 130, 2, 64), (2, 384, 3, 64), (2, 9, 3, 8)])
def test_mha_causal_extension(prec, dims):
    """causal=True (beyond the reference): fused kernels (head dim 64, tf32) and the materialised path
    against the oracle, whose causal variant is pinned against torch in tests/test_oracle.py."""
    import npm_b200
    from layers import MultiHeadAttention
    from oracle import np_oracle as O
    npm_b200.set_precision(prec)
    b, sq, h, d = dims
    rng = np.random.default_rng(sq * 7 + d)
    dm = h * d
    x = rng.standard_normal((b, sq, dm), dtype=np.float32)
    dy = rng.standard_normal((b, sq, dm), dtype=np.float32)
    layer = MultiHeadAttention(h, causal=True)
    layer(x)
    params = {k: (np.asarray(getattr(layer, k)) * (1.0 / np.sqrt(dm) if k.startswith('_w') else 0.1)).astype(np.float32)
              for k in O.MHA_PARAMS}
    bind(layer, params)
    tol = dict(rtol=2e-3, atol=2e-2) if prec == 'tf32' else TC
    want, cache = O.mha_fwd(params, x, causal=True)
    close(layer(x), want, **tol)
    close(layer._attention_scores, cache['prob'], rtol=2e-3, atol=1e-3 if prec == 'tf32' else 1e-5)
    (dq_, dk_, dv_), grads = O.mha_bwd(params, cache, dy)
    rec = Recorder()
    got = layer(dy, backprop=True, optimizer_=rec)
    close(got[0], dq_, **tol); close(got[1], dk_, **tol); close(got[2], dv_, **tol)
    gtol = dict(rtol=2e-3, atol=2e-2 * np.sqrt(b * sq)) if prec == 'tf32' else dict(rtol=1e-3, atol=1e-4 * np.sqrt(b * sq))
    gg = grads_of(layer, rec, O.MHA_PARAMS)
    for k in O.MHA_PARAMS:
        close(gg[k], grads[k], **gtol)
    # unmasked and causal layers differ (the mask is really applied)
    plain = MultiHeadAttention(h)
    plain(x)
    bind(plain, params)
    assert np.abs(np.asarray(plain(x)) - want).max() > 1e-3


def test_decoder_causal_self_attention():
    import npm_b200
    from layers import TransformerDecoder
    from oracle import np_oracle as O
    from train import iter_parameters
    rng = np.random.default_rng(11)
    b, s_, d, h, f = 2, 48, 128, 2, 256
    q = rng.standard_normal((b, s_, d)).astype(np.float32)
    kv = rng.standard_normal((b, 40, d)).astype(np.float32)
    layer = TransformerDecoder(h, f, True, 0.0, causal=True)
    layer(q, kv)
    params = {}
    for owner, name in iter_parameters(layer):
        v = np.asarray(getattr(owner, name))
        if name.startswith('_w'):
            v = (v * 0.1).astype(np.float32)
            setattr(owner, name, v)
    def path_of(owner, name):
        for attr, val in vars(layer).items():
            if val is owner:
                return f'{attr}.{name}'
            for attr2, val2 in (vars(val).items() if hasattr(val, '__dict__') else []):
                if val2 is owner:
                    return f'{attr}.{attr2}.{name}'
        raise KeyError(name)
    for owner, name in iter_parameters(layer):
        params[path_of(owner, name)] = np.asarray(getattr(owner, name))
    want, _ = O.decoder_fwd(params, q, kv, True, causal=True)
    close(layer(q, kv), want)
    plain, _ = O.decoder_fwd(params, q, kv, True)
    assert np.abs(plain - want).max() > 1e-3


def test_mha_mask_raises():
    from layers import MultiHeadAttention
    layer = MultiHeadAttention(2)
    q = np.zeros((1, 4, 8), np.float32)
    with pytest.raises(ValueError):
        layer(q, mask=np.ones((1, 2, 4, 4)))


def _transformer_case(kind, norm, drop):
    from layers import TransformerDecoder, TransformerEncoder
    g = load_golden(f'{kind}_{norm}_{drop}')
    cls = TransformerEncoder if kind == 'encoder' else TransformerDecoder
    layer = cls(2, 32, norm == 'pre', 0.25 if drop == 'drop' else 0.0)
    args = (g['q'],) if kind == 'encoder' else (g['q'], g['kv'])
    layer(*args)
    bind(layer, sub(g, 'p.'))
    if drop == 'drop':
        for i in (1, 2, 3):
            if f'mask{i}' in g:
                getattr(layer, f'_dropout{i}')._mask = g[f'mask{i}']
    close(layer(*args), g['out'])
    rec = Recorder()
    res = layer(g['dy'], backprop=True, optimizer_=rec)
    if kind == 'encoder':
        close(res, g['dq'])
    else:
        assert isinstance(res, tuple) and len(res) == 2
        close(res[0], g['dq']); close(res[1], g['dkv'])
    gg = sub(g, 'g.')
    got = grads_of(layer, rec, gg.keys())
    for k, v in gg.items():
        close(got[k], v, rtol=1e-3, atol=2e-4)


@pytest.mark.parametrize('norm', ['pre', 'post'])
@pytest.mark.parametrize('drop', ['nodrop', 'drop'])
def test_encoder_golden(norm, drop):
    _transformer_case('encoder', norm, drop)


@pytest.mark.parametrize('norm', ['pre', 'post'])
@pytest.mark.parametrize('drop', ['nodrop', 'drop'])
def test_decoder_golden(norm, drop):
    _transformer_case('decoder', norm, drop)


def test_decoder_vs_oracle_medium():
    """A decoder layer at a size the oracle finishes in seconds (B2, Sq 64, Skv 96, D 256, H 4, F 512),
    with Philox dropout: the GPU's own masks are injected into the oracle (normalizations_test.py:28)."""
    import optimizer
    from layers import TransformerDecoder
    from layers.normalizations import set_dropout_seed
    from oracle import np_oracle as O
    rng = np.random.default_rng(11)
    b, sq, skv, d, h, f = 2, 64, 96, 256, 4, 512
    q = rng.standard_normal((b, sq, d)).astype(np.float32)
    kv = rng.standard_normal((b, skv, d)).astype(np.float32)
    dy = rng.standard_normal((b, sq, d)).astype(np.float32)
    layer = TransformerDecoder(h, f, True, 0.1)
    set_dropout_seed(5)
    np.random.seed(11)          # the lazy initialisation draws from the global legacy stream (layer.py:57-60)
    layer(q, kv)
    # fan-in scaled weights so deep paths stay O(1)
    from train import iter_parameters
    params = {}
    for owner, name in iter_parameters(layer):
        v = np.asarray(getattr(owner, name))
        if name.startswith('_w'):
            fan_in = v.shape[-1] if v.ndim == 3 and name != '_wo' else (v.shape[1] * v.shape[2] if name == '_wo' else v.shape[0])
            v = (v / np.sqrt(fan_in)).astype(np.float32)
            setattr(owner, name, v)
    set_dropout_seed(6)
    out = layer(q, kv)
    masks = tuple(getattr(layer, f'_dropout{i}')._mask for i in (1, 2, 3))
    p = {}
    for path in ['_self_attention.' + k for k in O.MHA_PARAMS] + ['_cross_attention.' + k for k in O.MHA_PARAMS] + \
            ['_dense1._linear._w', '_dense1._linear._b', '_dense2._w', '_dense2._b'] + \
            [f'_norm{i}.{n}' for i in (1, 2, 3) for n in ('_gamma', '_beta')]:
        p[path] = param_values(layer, [path])[path]
    keep = np.float32(0.9)
    oout, cache = O.decoder_fwd(p, q, kv, True, masks, keep)
    close(out, oout)
    use_device_relu_gates(layer, cache)
    rec = Recorder()
    dq, dkv = layer(dy, backprop=True, optimizer_=rec)
    (odq, odkv), ograds = O.decoder_bwd(p, cache, dy, True, masks, keep)
    close(dq, odq); close(dkv, odkv)
    got = grads_of(layer, rec, ograds.keys())
    for k, v in ograds.items():
        close(got[k], v, rtol=1e-3, atol=1e-4 * max(1.0, np.abs(v).max()))


def test_trainer_cuda_graph_replay_equals_the_eager_loop():
    """Trainer(cuda_graph=True): after two eager steps the whole step (forward chain, loss, reverse chain, fused SGD update)
    is captured once and replayed; parameters and losses must equal the eager loop's (to the last bits: atomic sums), with
    different data in every step (train.py:20-39).  Adam (host-side bias correction) and stochastic layers stay eager."""
    import loss
    import optimizer
    from layers import Dense, Softmax
    from npm_b200._lib import C
    from train import Trainer, iter_parameters
    rng = np.random.default_rng(21)
    xs = [rng.standard_normal((64, 96)).astype(np.float32) for _ in range(7)]
    ts = [np.eye(10, dtype=np.float32)[rng.integers(0, 10, 64)] for _ in range(7)]
    runs = []
    for graphed in (False, True):
        np.random.seed(21)
        layers_ = [Dense(48), Dense(10, activation=Softmax())]
        tr = Trainer(layers_, loss.CrossEntropyLoss(), verbose=False, cuda_graph=graphed)
        opt = optimizer.SGDOptimizer(1e-3)
        tr._forward(tr._to_device('inputs', xs[0]))
        for owner, name in iter_parameters(layers_):
            v = np.asarray(getattr(owner, name))
            setattr(owner, name, (v / np.sqrt(v.shape[0])).astype(np.float32) if name == '_w' else v)
        losses = []
        # one call with the same batch six times (graph from the third step on), then per-batch calls of four steps each
        tr.train(xs[0], ts[0], 6, opt)
        losses.append(float(tr.last_loss))
        for x, t in zip(xs[1:], ts[1:]):
            before = C.npm_launch_count()
            tr.train(x, t, 4, opt)
            launched = C.npm_launch_count() - before
            losses.append(float(tr.last_loss))
        runs.append((losses, [np.asarray(getattr(o, n)) for o, n in iter_parameters(layers_)], launched, len(tr._graphs)))
    (l0, p0, n0, g0), (l1, p1, n1, g1) = runs
    assert g0 == 0 and g1 == 1, 'one capture, reused by the later train() calls'
    assert n1 < n0, 'replayed steps launch nothing through the C-ABI'
    # same kernels on the same data; the loss sum and the split-K reductions add in an order that is not fixed run to run
    np.testing.assert_allclose(np.array(l1), np.array(l0), rtol=1e-6)
    for a, e in zip(p1, p0):
        np.testing.assert_allclose(a, e, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('opt_name', ['adam', 'sgd'])
def test_optimizer_maintains_the_weight_planes(opt_name, monkeypatch):
    """Split-bf16 mode: the fused optimizers rewrite each weight's bf16 hi / mid image inside their update kernel
    (npm_tensor_entry.planes), so the forward after an update launches no split pass over the weights.  Same training
    trajectory, bit for bit, as with a fresh npm_weight_split every forward (optimizer.py:30-33, 50-69)."""
    import npm_b200
    import optimizer
    from layers import TransformerEncoder
    from npm_b200 import device
    from npm_b200._lib import C
    from train import iter_parameters
    npm_b200.set_precision('bf16x3')
    rng = np.random.default_rng(9)
    b, s_, d, h, f = 2, 136, 128, 2, 256
    x = rng.standard_normal((b, s_, d)).astype(np.float32)
    dy = (rng.standard_normal((b, s_, d)) * 0.1).astype(np.float32)
    np.random.seed(9)
    layer = TransformerEncoder(h, f, True, 0.0)
    layer(x)
    for owner, name in iter_parameters(layer):
        v = np.asarray(getattr(owner, name))
        if name.startswith('_w'):
            setattr(owner, name, (v / np.sqrt(max(v.shape[-1], v.shape[0]))).astype(np.float32))
    runs = []
    for maintained in (False, True):
        monkeypatch.setattr(device, '_NO_OPT_PLANES', not maintained)
        clone = copy.deepcopy(layer)
        opt = optimizer.AdamOptimizer(learning_rate=1e-2) if opt_name == 'adam' else optimizer.SGDOptimizer(1e-2)
        counts = []
        for _ in range(3):
            before = C.npm_launch_count()
            clone(x)
            counts.append(C.npm_launch_count() - before)
            clone(dy, backprop=True, optimizer_=opt)
        # a raw write between an update and the next forward must be announced; then the forward splits again
        p = clone._dense2._p('_w')
        p.t.mul_(1.0)
        p.touched()
        before = C.npm_launch_count()
        out = clone(x)
        counts.append(C.npm_launch_count() - before)
        runs.append((counts, np.asarray(out), {f'{i}.{n}': np.asarray(getattr(o, n)) for i, (o, n) in enumerate(iter_parameters(clone))}))
    (c0, o0, p0), (c1, o1, p1) = runs
    n_weights = 4                                           # packed q|k|v, wo, dense1, dense2
    assert c0[0] == c0[1] == c0[2], 'without the feature every forward splits every weight'
    assert c1[0] == c0[0] and c1[1] == c1[2] == c0[0] - n_weights, (c0, c1)
    assert c1[3] == c0[0] - n_weights + 1, 'the touched weight, and only it, is split again'
    np.testing.assert_array_equal(o1, o0)
    for k in p0:
        np.testing.assert_array_equal(p1[k], p0[k])


@pytest.mark.parametrize('norm_first', [True, False])
def test_ffn_hidden_activation_as_split_bf16_planes(norm_first, monkeypatch):
    """The split-bf16 FFN route of layers/mlp.py (`_NO_FFN_PLANES` off): the FFN hidden activation and its gradient exist only as bf16
    hi / mid planes (device.PlanesArray) — written by the first GEMM's epilogue / the ReLU backward, landed by the
    consuming GEMMs without conversion.  The planes hold exactly the pairs the converters would make, so outputs,
    input gradients and every parameter gradient must equal the fp32 route's to fp32 round-off (mlp.py:21-38, 70-77)."""
    import npm_b200
    from layers import TransformerEncoder, mlp
    from npm_b200 import device
    from train import iter_parameters
    npm_b200.set_precision('bf16x3')
    rng = np.random.default_rng(3)
    b, s_, d, h, f = 2, 136, 128, 2, 512                   # 272 tokens: the split-bf16 GEMM takes every FFN problem
    x = rng.standard_normal((b, s_, d)).astype(np.float32)
    dy = rng.standard_normal((b, s_, d)).astype(np.float32)
    np.random.seed(3)
    layer = TransformerEncoder(h, f, norm_first, 0.0)
    layer(x)
    for owner, name in iter_parameters(layer):
        v = np.asarray(getattr(owner, name))
        if name.startswith('_w'):
            setattr(owner, name, (v / np.sqrt(max(v.shape[-1], v.shape[0]))).astype(np.float32))
    results = []
    for planes_on in (False, True):
        monkeypatch.setattr(mlp, '_NO_FFN_PLANES', not planes_on)
        clone = copy.deepcopy(layer)
        out = clone(x)
        assert isinstance(clone._dense1._y, device.PlanesArray) == planes_on
        rec = Recorder()
        dx = clone(dy, backprop=True, optimizer_=rec)
        names = [f'{i}.{name}' for i, (owner, name) in enumerate(iter_parameters(clone))]
        grads = [np.asarray(rec.grads[f'{id(owner)}.{name}']) for owner, name in iter_parameters(clone)]
        results.append((np.asarray(out), np.asarray(dx), dict(zip(names, grads))))
    (o0, d0, g0), (o1, d1, g1) = results
    # same operand values, but the GEMM may pick another tile shape / K split for pre-split operands: fp32 round-off
    close(o1, o0, rtol=1e-5, atol=2e-5)
    close(d1, d0, rtol=1e-5, atol=2e-5)
    for k in g0:
        close(g1[k], g0[k], rtol=1e-5, atol=2e-5 * max(1.0, np.abs(g0[k]).max()))
    # the planes join back to fp32 exactly (hi + mid of a value that was split from fp32 has 16 significant bits)
    hidden = clone._dense1._y
    assert hidden.shape == (b * s_, f) and np.asarray(hidden).shape == (b * s_, f)


def test_ffn_planes_survive_a_precision_change_between_forward_and_backward():
    """The hidden activation written as split-bf16 planes in bf16x3 mode must still back-propagate when the mode is
    changed before backward (nothing may misread what forward saved): the GEMMs of the other modes decline the
    planes (NPM_ERR_UNSUPPORTED), the layer joins them to fp32 and takes the plain route."""
    import npm_b200
    from layers import TransformerEncoder
    from npm_b200 import device
    from train import iter_parameters
    rng = np.random.default_rng(5)
    b, s_, d, h, f = 2, 136, 128, 2, 256
    x = rng.standard_normal((b, s_, d)).astype(np.float32)
    dy = rng.standard_normal((b, s_, d)).astype(np.float32)
    npm_b200.set_precision('bf16x3')
    np.random.seed(5)
    layer = TransformerEncoder(h, f, True, 0.0)
    layer(x)
    for owner, name in iter_parameters(layer):
        v = np.asarray(getattr(owner, name))
        if name.startswith('_w'):
            setattr(owner, name, (v / np.sqrt(max(v.shape[-1], v.shape[0]))).astype(np.float32))
    got = []
    for bwd_mode in ('bf16x3', '3xtf32'):
        npm_b200.set_precision('bf16x3')
        clone = copy.deepcopy(layer)
        clone(x)
        assert isinstance(clone._dense1._y, device.PlanesArray)
        npm_b200.set_precision(bwd_mode)
        rec = Recorder()
        dx = clone(dy, backprop=True, optimizer_=rec)
        got.append((np.asarray(dx), [np.asarray(rec.grads[f'{id(o)}.{n}']) for o, n in iter_parameters(clone)]))
    npm_b200.set_precision('bf16x3')
    (dx0, g0), (dx1, g1) = got
    close(dx1, dx0)
    for a, e in zip(g1, g0):
        close(a, e, rtol=1e-3, atol=1e-4 * np.sqrt(b * s_) * max(1.0, np.abs(e).max()))


# ------------------------------------------------------------------ conv
@pytest.mark.parametrize('tag', ['c3', 'c8', 'k5', 'k1'])
def test_conv_golden(tag):
    from layers import Conv2D
    g = load_golden('conv_' + tag)
    k, _, _, c1 = g['p._w'].shape
    layer = Conv2D(c1, k)
    layer(g['x'])
    bind(layer, sub(g, 'p.'))
    close(layer(g['x']), g['y'])
    rec = Recorder()
    close(layer(g['dy'], backprop=True, optimizer_=rec), g['dx'])
    got = grads_of(layer, rec, ['_w', '_b'])
    close(got['_w'], g['g._w']); close(got['_b'], g['g._b'])


@pytest.mark.parametrize('shape', [(4, 32, 32, 3, 64, 3), (2, 32, 32, 64, 128, 3), (3, 9, 7, 16, 20, 5), (2, 16, 16, 32, 32, 1),
                                   (2, 8, 8, 3, 20, 5), (2, 5, 6, 1, 128, 3), (2, 6, 5, 4, 64, 1), (1, 7, 9, 2, 192, 3)])
def test_conv_vs_oracle(shape):
    from layers import Conv2D
    from oracle import np_oracle as O
    n, hh, ww, c0, c1, k = shape
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal((n, hh, ww, c0)).astype(np.float32)
    dy = rng.standard_normal((n, hh, ww, c1)).astype(np.float32)
    f = (rng.standard_normal((k, k, c0, c1)) / np.sqrt(k * k * c0)).astype(np.float32)
    b = rng.standard_normal(c1).astype(np.float32)
    layer = Conv2D(c1, k)
    layer(x)
    bind(layer, {'_w': f, '_b': b})
    y, z = O.conv_layer_fwd(x, f, b)
    close(layer(x), y)
    rec = Recorder()
    dx = layer(dy, backprop=True, optimizer_=rec)
    odx, odw, odb = O.conv_layer_bwd(x, f, z, dy)
    close(dx, odx)
    got = grads_of(layer, rec, ['_w', '_b'])
    close(got['_w'], odw, rtol=1e-3, atol=1e-4 * np.sqrt(n * hh * ww))
    close(got['_b'], odb, rtol=1e-4, atol=1e-4 * np.sqrt(n * hh * ww))


# ------------------------------------------------------------------ losses / optimizers / trainer
def test_losses_golden():
    import loss
    g = load_golden('loss')
    mse = loss.MSELoss()
    close(float(mse(g['y'], g['t'])), g['mse'], rtol=1e-5, atol=1e-6)
    close(mse(backprop=True), g['mse_dy'], **EW)
    ce = loss.CrossEntropyLoss()
    close(float(ce(g['prob'], g['onehot'])), g['ce'], rtol=1e-5, atol=1e-5)
    close(ce(backprop=True), g['ce_dy'], **EW)


def test_optimizers_golden():
    import optimizer
    from npm_b200 import device
    g = load_golden('optimizer')

    class Box:
        pass

    box = Box()
    box.w = device.asdevice(g['w0'])
    optimizer.SGDOptimizer(0.1).update(box, 'w', g['grads'][0])
    close(box.w, g['sgd'], rtol=1e-6, atol=1e-6)
    box.w = device.asdevice(g['w0'])
    adam = optimizer.AdamOptimizer(learning_rate=0.01)
    for t, grad in enumerate(g['grads']):
        adam.update(box, 'w', grad)
        close(box.w, g['adam'][t], rtol=1e-5, atol=2e-6)


def test_adam_many_tensors_one_launch():
    """The fused update touches tensors of ragged sizes (incl. non-multiples of 4 and > one chunk)."""
    import npm_b200
    import optimizer
    from npm_b200 import device
    from oracle import np_oracle as O
    rng = np.random.default_rng(3)
    sizes = [1, 3, 7, 64, 1000, 8192, 8193, 50000, 12]

    class Box:
        pass

    box = Box()
    host = {}
    for i, n in enumerate(sizes):
        host[f'p{i}'] = rng.standard_normal(n).astype(np.float32)
        setattr(box, f'p{i}', device.asdevice(host[f'p{i}']))
    opt = optimizer.AdamOptimizer(learning_rate=0.05)
    state = {k: (np.zeros_like(v, np.float64), np.zeros_like(v, np.float64)) for k, v in host.items()}
    for t in (1, 2):
        grads = {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in host.items()}
        opt._enter()
        for k in host:
            opt.update(box, k, grads[k])
        before = npm_b200.launch_count()
        opt._exit()
        # ONE multi-tensor kernel (+ on the first step the library's own zero fills of the 2 x 9 moment buffers)
        assert npm_b200.launch_count() - before == (1 if t > 1 else 1 + 2 * len(sizes))
        for k in host:
            host[k], m, v = O.adam_step(host[k], grads[k], *state[k], t, 0.05)
            state[k] = (m, v)
            close(getattr(box, k), host[k], rtol=1e-5, atol=2e-6)


def test_trainer_mlp_golden(capsys):
    import loss
    import optimizer
    from layers import Dense, Softmax
    from train import Trainer
    g = load_golden('trainer_mlp')
    layers = [Dense(8), Dense(4, activation=Softmax())]
    trainer = Trainer(layers, loss.CrossEntropyLoss())
    trainer.eval(g['x'], g['t'])
    for i, layer in enumerate(layers):
        bind(layer, sub(g, f'p0.{i}.'))
    capsys.readouterr()
    trainer.train(g['x'], g['t'], 4, optimizer.SGDOptimizer(1e-2))
    out = capsys.readouterr().out.splitlines()
    assert [line.split()[0] for line in out] == ['Step:', 'Loss:'] * 4     # train.py:24,32
    losses = [float(line.split()[-1]) for line in out if line.startswith('Loss')]
    close(losses, g['losses'], rtol=1e-4, atol=1e-4)
    for i, layer in enumerate(layers):
        for k, v in sub(g, f'p1.{i}.').items():
            close(param_values(layer, [k])[k], v, rtol=1e-3, atol=1e-4)


def _trainer_golden(name, make_layers, steps=3, lr=1e-2, ptol=None, ltol=None):
    import loss
    import optimizer
    from train import Trainer
    g = load_golden(name)
    layers = make_layers()
    trainer = Trainer(layers, loss.MSELoss(), verbose=False)
    trainer.eval(g['x'], g['t'])                                   # lazy initialisation, as the generator did
    for i, layer in enumerate(layers):
        bind(layer, sub(g, f'p0.{i}.'))
    losses = []
    adam = optimizer.AdamOptimizer(learning_rate=lr)
    for _ in range(steps):
        trainer.train(g['x'], g['t'], 1, adam)
        losses.append(float(trainer.last_loss))
    close(losses, g['losses'], **(ltol or dict(rtol=1e-3, atol=1e-4)))
    for i, layer in enumerate(layers):
        for k, v in sub(g, f'p1.{i}.').items():
            close(param_values(layer, [k])[k], v, **(ptol or dict(rtol=1e-3, atol=2e-4)))


def test_trainer_conv_golden():
    """The reference's Trainer + AdamOptimizer on two Conv2D layers, 3 steps (tests/golden/trainer_conv.npz): losses and
    every parameter after training."""
    from layers import Conv2D
    _trainer_golden('trainer_conv', lambda: [Conv2D(4, 3), Conv2D(6, 3)])


def test_trainer_encoder_golden():
    """The reference's Trainer + AdamOptimizer on a pre-norm and a post-norm TransformerEncoder, 3 steps."""
    from layers import TransformerEncoder
    # Adam normalises the update to ~lr per element whatever the gradient's size, so a gradient element that is tiny
    # relative to the 3xTF32 error can move by a sizeable fraction of lr = 1e-2: compare parameters to 1e-3 absolute
    _trainer_golden('trainer_encoder', lambda: [TransformerEncoder(2, 32, True, 0.0), TransformerEncoder(2, 32, False, 0.0)],
                    ptol=dict(rtol=1e-3, atol=1e-3))


# ------------------------------------------------------------------ fused bandwidth kernels
def test_dense_fused_relu_keeps_reference_gradient_at_zero():
    """Dense's default ReLU runs in the GEMM epilogue (pre-activation never written).  The reference's
    backward passes the gradient where x >= 0 (activations.py:19) — including x == 0 exactly, which
    the epilogue keeps apart from x < 0 through the sign of the stored zero."""
    from layers import Dense
    from oracle import np_oracle as O
    rng = np.random.default_rng(5)
    m, k, n = 96, 64, 48
    x = rng.standard_normal((m, k)).astype(np.float32)
    w = rng.standard_normal((k, n)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    w[:, :8] = 0.0
    b[:8] = 0.0                                   # pre-activation exactly 0 in the first 8 columns
    dy = rng.standard_normal((m, n)).astype(np.float32)
    layer = Dense(n)
    layer(x)
    bind(layer, {'_linear._w': w, '_linear._b': b})
    y = np.asarray(layer(x))
    z = O.linear_fwd(x, w, b)
    close(y, np.maximum(z, 0.0))
    assert (y >= 0).all() and (y[:, :8] == 0).all()
    rec = Recorder()
    dx = layer(dy, backprop=True, optimizer_=rec)
    dz = np.where(z >= 0.0, dy.astype(np.float64), 0.0)          # gradient passes in the zero columns
    assert (dz[:, :8] == dy[:, :8]).all()
    odx, odw, odb = O.linear_bwd(x, w, dz)
    close(dx, odx)
    g = grads_of(layer, rec, ['_linear._w', '_linear._b'])
    close(g['_linear._w'], odw, rtol=1e-3, atol=1e-3)
    close(g['_linear._b'], odb, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('rows,cols', [(1, 4), (8192, 1024), (8192, 4096), (333, 260), (70, 130), (5000, 36)])
def test_colsum_single_launch_matches_numpy_and_is_deterministic(rows, cols):
    import torch
    from npm_b200 import device
    from npm_b200._lib import C
    rng = np.random.default_rng(rows + cols)
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    tx = torch.from_numpy(x).cuda()
    outs = []
    for _ in range(3):
        out = torch.full((cols,), float('nan'), device='cuda')
        ws = device.workspace(C.npm_colsum_workspace(rows, cols))
        C.npm_colsum(tx.data_ptr(), out.data_ptr(), rows, cols, ws.data_ptr(), device.stream())
        outs.append(out.cpu().numpy())
    close(outs[0], x.astype(np.float64).sum(0), rtol=1e-5, atol=1e-3)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])     # fixed summation order


@pytest.mark.parametrize('shape', [(4, 32, 128), (2, 100, 1024), (64, 260)])
def test_fused_dropout_layernorm_equals_the_two_layers(shape):
    """dropout_layernorm_forward/backward (one kernel each) == DropOut then LayerNormalization, backward then
    `+= dskip`: the same Philox mask bit for bit; values to within the 1 ulp that separates the fused
    kernels' multiply by fl(1 / keep) from DropOut's exact division (normalizations.py:22)."""
    from layers.normalizations import (DropOut, LayerNormalization, dropout_layernorm_backward,
                                       dropout_layernorm_forward, set_dropout_seed)
    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape).astype(np.float32)
    dz = rng.standard_normal(shape).astype(np.float32)
    dskip = rng.standard_normal(shape).astype(np.float32)
    gamma = rng.standard_normal(shape[-1]).astype(np.float32)
    beta = rng.standard_normal(shape[-1]).astype(np.float32)

    def run(fused):
        set_dropout_seed(77, 40)
        drop, norm = DropOut(0.25), LayerNormalization()
        norm(x)
        bind(norm, {'_gamma': gamma, '_beta': beta})
        set_dropout_seed(77, 40)
        rec = Recorder()
        if fused:
            y = dropout_layernorm_forward(drop, norm, x)
            assert norm._fused is not None
            dx = dropout_layernorm_backward(drop, norm, dz, dskip_d(), rec)
        else:
            y = norm(drop(x))
            dx = drop.backward(norm.backward(dz, rec))
            dx += dskip_d()
        g = grads_of(norm, rec, ['_gamma', '_beta'])
        return np.asarray(y), np.asarray(dx), g['_gamma'], g['_beta'], drop._mask

    from npm_b200 import device
    dskip_d = lambda: device.asdevice(dskip)
    a, b = run(True), run(False)
    assert np.array_equal(a[4], b[4])
    for u, v in zip(a[:4], b[:4]):
        np.testing.assert_allclose(u, v, rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(v).max())))
    assert abs(a[4].mean() - 0.75) < 0.05


def test_checkpoint_resume_is_bit_exact(tmp_path):
    """parameters + Adam state + dropout stream position round-trip through Trainer.save_checkpoint /
    load_checkpoint; training resumed from the file continues exactly like the uninterrupted run."""
    import npm_b200
    import loss
    import optimizer
    from layers.adapters import EncoderStack
    from layers.normalizations import set_dropout_seed
    from train import Trainer, iter_parameters
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 24, 64)).astype(np.float32)
    t = rng.standard_normal((2, 24, 64)).astype(np.float32)

    def fresh():
        np.random.seed(5)
        stack = EncoderStack(2, 2, 96, True, 0.1)
        tr = Trainer([stack], loss.MSELoss(), verbose=False)
        tr._forward(x)
        for owner, name in iter_parameters([stack]):
            if name.startswith('_w'):
                setattr(owner, name, (np.asarray(getattr(owner, name)) * 0.1).astype(np.float32))
        return stack, tr

    set_dropout_seed(99)
    stack_a, tr_a = fresh()
    adam_a = optimizer.AdamOptimizer(learning_rate=1e-2)
    tr_a.train(x, t, 2, adam_a)
    tr_a.save_checkpoint(str(tmp_path / 'ck'), adam_a)
    tr_a.train(x, t, 2, adam_a)
    want = [np.asarray(getattr(o, n)) for o, n in iter_parameters([stack_a])]

    set_dropout_seed(1)                                   # a different stream position: the checkpoint must restore it
    stack_b, tr_b = fresh()
    adam_b = optimizer.AdamOptimizer(learning_rate=1e-2)
    tr_b.load_checkpoint(str(tmp_path / 'ck'), adam_b)
    tr_b.train(x, t, 2, adam_b)
    got = [np.asarray(getattr(o, n)) for o, n in iter_parameters([stack_b])]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    # the loss itself is a sum of per-CTA partials combined with atomics: equal up to the summation order
    assert abs(float(tr_a.last_loss) - float(tr_b.last_loss)) <= 1e-6 * abs(float(tr_a.last_loss))


# ------------------------------------------------------------------ round-2 regressions
def test_dense_output_may_be_mutated_in_place():
    """The reference caches the pre-activation, so `out = layer(x); out += skip` (its own block idiom) is safe there.
    Dense's fused ReLU reads its mask from the forward output: a caller must get its own buffer."""
    from layers import Dense
    from oracle import np_oracle as O
    rng = np.random.default_rng(21)
    x = rng.standard_normal((300, 64)).astype(np.float32)
    dy = rng.standard_normal((300, 48)).astype(np.float32)
    w = (rng.standard_normal((64, 48)) / 8).astype(np.float32)
    b = rng.standard_normal(48).astype(np.float32)
    layer = Dense(48)
    layer(x)
    bind(layer, {'_linear._w': w, '_linear._b': b})
    out = layer(x)
    out += np.full((300, 48), -100.0, np.float32)          # destroys every sign bit of the returned buffer
    rec = Recorder()
    dx = layer(dy, backprop=True, optimizer_=rec)
    z = O.linear_fwd(x, w, b)
    dz = O.relu_bwd(z, dy)
    odx, odw, odb = O.linear_bwd(x, w, dz)
    close(dx, odx)
    g = grads_of(layer, rec, ['_linear._w', '_linear._b'])
    close(g['_linear._w'], odw, rtol=1e-3, atol=1e-3)
    close(g['_linear._b'], odb, rtol=1e-4, atol=1e-3)


def test_negative_zero_preactivation_passes_the_gradient():
    """activations.py:19 is `x >= 0`: a pre-activation of exactly -0.0 keeps its gradient."""
    from layers import Dense
    x = np.zeros((200, 8), np.float32)
    w = np.zeros((8, 4), np.float32)
    b = np.array([-0.0, 0.0, 1.0, -1.0], np.float32)
    layer = Dense(4)
    layer(x)
    bind(layer, {'_linear._w': w, '_linear._b': b})
    layer(x)
    rec = Recorder()
    layer(np.ones((200, 4), np.float32), backprop=True, optimizer_=rec)
    g = grads_of(layer, rec, ['_linear._b'])['_linear._b']
    np.testing.assert_array_equal(g, np.array([200.0, 200.0, 200.0, 0.0], np.float32))


def test_trainer_minibatch_loop_uploads_each_batch(capsys):
    """`for xb, yb in data: trainer.train(xb, yb, 1, opt)` with verbose=False never synchronises: the pinned staging
    buffer of a step must not be rewritten before its asynchronous H2D copy has run."""
    import loss
    import optimizer
    from layers import Dense
    from train import Trainer
    rng = np.random.default_rng(5)
    batches = [(rng.standard_normal((512, 256)).astype(np.float32), rng.standard_normal((512, 8)).astype(np.float32)) for _ in range(6)]
    w0 = (rng.standard_normal((256, 8)) / 16).astype(np.float32)

    def run(sync_every_step):
        import torch
        layer = Dense(8)
        layer(batches[0][0])
        bind(layer, {'_linear._w': w0, '_linear._b': np.zeros(8, np.float32)})
        tr = Trainer([layer], loss.MSELoss(), verbose=False)
        opt = optimizer.SGDOptimizer(0.05)
        for xb, yb in batches:
            tr.train(xb, yb, 1, opt)
            if sync_every_step:
                torch.cuda.synchronize()
        return np.asarray(layer.linear.w)

    np.testing.assert_array_equal(run(False), run(True))


def test_colsum_attachment_is_invalidated_through_aliases():
    from npm_b200 import device
    a = device.asdevice(np.ones((4, 6, 8), np.float32))
    a.colsum = device.asdevice(np.full(8, 24.0, np.float32))
    b = a.reshape(24, 8)
    assert b.colsum is not None and a.reshape(4, 48).colsum is None
    a += np.ones((4, 6, 8), np.float32)                    # a write through ONE alias ...
    assert a.colsum is None and b.colsum is None           # ... invalidates the attachment for all of them
    a.colsum = device.asdevice(np.full(8, 48.0, np.float32))
    v = a[0]
    v += np.ones((6, 8), np.float32)                       # and through a leading-axis view
    assert a.colsum is None


def test_same_layer_twice_in_one_backward_keeps_sequential_updates():
    """optimizer.py:13-18 applies an update when it is issued; a layer used twice inside one bracket must see its first
    update applied before the second gradient overwrites the persistent buffer."""
    import optimizer
    from layers import Linear
    rng = np.random.default_rng(9)
    x1, x2 = (rng.standard_normal((160, 32)).astype(np.float32) for _ in range(2))
    dy1, dy2 = (rng.standard_normal((160, 16)).astype(np.float32) for _ in range(2))
    w = rng.standard_normal((32, 16)).astype(np.float32)
    layer = Linear(16)
    layer(x1)
    bind(layer, {'_w': w, '_b': np.zeros(16, np.float32)})
    opt = optimizer.SGDOptimizer(0.1)
    opt._enter()
    layer(x1)
    layer.backward(dy1, opt)
    layer(x2)
    layer.backward(dy2, opt)
    opt._exit()
    want = w.astype(np.float64) - 0.1 * (x1.astype(np.float64).T @ dy1) - 0.1 * (x2.astype(np.float64).T @ dy2)
    close(layer.w, want, rtol=1e-3, atol=1e-3)


def test_user_optimizer_in_the_reference_style():
    """A user-defined Optimizer written like the reference's SGD (`variable -= lr * gradient`, optimizer.py:30-33)."""
    import optimizer
    from layers import Linear

    class PlainSGD(optimizer.Optimizer):
        def update_variable(self, identifier, variable, gradient):
            variable -= 0.05 * gradient
            return variable

    rng = np.random.default_rng(10)
    x = rng.standard_normal((64, 12)).astype(np.float32)
    dy = rng.standard_normal((64, 5)).astype(np.float32)
    w = rng.standard_normal((12, 5)).astype(np.float32)
    layer = Linear(5)
    layer(x)
    bind(layer, {'_w': w, '_b': np.zeros(5, np.float32)})
    layer(x)
    layer(dy, backprop=True, optimizer_=PlainSGD())
    close(layer.w, w - 0.05 * (x.T @ dy), rtol=1e-4, atol=1e-4)
    close(layer.b, -0.05 * dy.sum(0), rtol=1e-4, atol=1e-4)


def test_attention_path_is_pinned_from_forward_to_backward():
    """`saved` is a log-sum-exp per row on the fused paths and the full probability tensor on the materialised one:
    backward must run the implementation forward ran even if the precision mode changes in between."""
    import npm_b200
    from layers import MultiHeadAttention
    rng = np.random.default_rng(12)
    x = rng.standard_normal((2, 130, 128)).astype(np.float32)
    dy = rng.standard_normal((2, 130, 128)).astype(np.float32)

    def grads(fwd_mode, bwd_mode):
        np.random.seed(3)
        npm_b200.set_precision(fwd_mode)
        layer = MultiHeadAttention(2)
        layer(x)
        for name in ('_wq', '_wk', '_wv', '_wo'):       # the reference's unscaled init saturates the softmax at d = 128
            setattr(layer, name, (np.asarray(getattr(layer, name)) * 0.1).astype(np.float32))
        layer(x)
        npm_b200.set_precision(bwd_mode)
        rec = Recorder()
        dq, dk, dv = layer(dy, backprop=True, optimizer_=rec)
        return np.asarray(dq, dtype=np.float64) + np.asarray(dk) + np.asarray(dv)

    want = grads('3xtf32', '3xtf32')
    for fwd_mode, bwd_mode in (('tf32', '3xtf32'), ('3xtf32', 'tf32'), ('bf16x3', '3xtf32'), ('3xtf32', 'bf16x3'), ('tf32', 'bf16x3')):
        got = grads(fwd_mode, bwd_mode)
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert np.isfinite(got).all() and err < 5e-3, (fwd_mode, bwd_mode, err)


def test_trainer_prefetch_overlaps_and_matches_plain_training():
    """Trainer.prefetch uploads the next call's batch on a copy stream; training with it equals training without."""
    import loss
    import optimizer
    import torch
    from layers import Dense
    from train import Trainer
    rng = np.random.default_rng(6)
    batches = [(torch.from_numpy(rng.standard_normal((256, 64)).astype(np.float32)).pin_memory(),
                torch.from_numpy(rng.standard_normal((256, 4)).astype(np.float32)).pin_memory()) for _ in range(5)]
    w0 = (rng.standard_normal((64, 4)) / 8).astype(np.float32)

    def run(prefetch):
        layer = Dense(4)
        layer(batches[0][0].numpy())
        bind(layer, {'_linear._w': w0, '_linear._b': np.zeros(4, np.float32)})
        tr = Trainer([layer], loss.MSELoss(), verbose=False)
        opt = optimizer.SGDOptimizer(0.05)
        for i, (xb, yb) in enumerate(batches):
            tr.train(xb, yb, 1, opt)
            if prefetch and i + 1 < len(batches):
                tr.prefetch(*batches[i + 1])
        torch.cuda.synchronize()
        return np.asarray(layer.linear.w)

    np.testing.assert_array_equal(run(True), run(False))


@pytest.mark.parametrize('majors', ['kk', 'km', 'mk', 'mm'])
@pytest.mark.parametrize('pre', ['a', 'b', 'ab', 'c'])
def test_gemm_split_planes_as_operands_and_output(majors, pre):
    """npm_gemm_desc.a_split / b_split / c_split (split-bf16 kernel): operands handed over as bf16 hi / mid planes
    (npm_weight_split) and the result written as planes give what the fp32 entry computes — against float64."""
    import ctypes
    import torch
    from npm_b200 import device
    from npm_b200._lib import C, GemmDesc
    M, N, K = 384, 320, 264
    rng = np.random.default_rng(40)
    a = rng.standard_normal((M, K) if majors[0] == 'k' else (K, M)).astype(np.float32)
    b = rng.standard_normal((N, K) if majors[1] == 'k' else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ta, tb, tbias = (torch.from_numpy(x).cuda() for x in (a, b, bias))
    tc = torch.full((M, N), float('nan'), device='cuda')
    st = device.stream()

    def planes(t):
        p = torch.empty(t.numel() * 4, dtype=torch.uint8, device='cuda')
        C.npm_weight_split(t.data_ptr(), p.data_ptr(), t.shape[0], t.shape[1], st)
        return p

    d = GemmDesc(a=ta.data_ptr(), b=tb.data_ptr(), c=tc.data_ptr(), bias=tbias.data_ptr(), m=M, n=N, k=K,
                 a_rs=K if majors[0] == 'k' else 1, a_cs=1 if majors[0] == 'k' else M,
                 b_rs=1 if majors[1] == 'k' else N, b_cs=K if majors[1] == 'k' else 1, ldc=N, nb1=1, nb2=1, alpha=1.0, flags=0,
                 precision=3)
    keep = []
    if 'a' in pre:
        keep.append(planes(ta)); d.a_split, d.a_split_plane, d.a = keep[-1].data_ptr(), M * K, None
    if 'b' in pre:
        keep.append(planes(tb)); d.b_split, d.b_split_plane = keep[-1].data_ptr(), N * K
    if pre == 'c':
        out = torch.empty(2, M, N, dtype=torch.bfloat16, device='cuda')
        d.c_split, d.c_split_plane, d.c = out.data_ptr(), M * N, None
    C.npm_gemm(ctypes.byref(d), st)
    torch.cuda.synchronize()
    A = a.astype(np.float64) if majors[0] == 'k' else a.astype(np.float64).T
    B = b.astype(np.float64).T if majors[1] == 'k' else b.astype(np.float64)
    want = A @ B + bias
    got = out.float().sum(0).cpu().numpy() if pre == 'c' else tc.cpu().numpy()
    close(got, want, rtol=1e-3, atol=4e-4)
