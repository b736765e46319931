"""The oracle (oracle/np_oracle.py, oracle/philox.py) reproduces every golden vector produced by the
unmodified reference (oracle/make_golden.py) — this is what pins it.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, sub
from oracle import np_oracle as O
from oracle import philox

TOL = dict(rtol=2e-5, atol=2e-6)   # reference mixes float32/float64; oracle is float64 throughout


def close(a, b, **kw):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), **(kw or TOL))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), '6627e8d5 e169c58d bc57ac4c 9b00dbd8'),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), '408f276d 41c83b0e a20bc7c6 6d5451fd'),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            'd16cfe09 94fdcceb 5001e420 24126ea1')]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert ' '.join('%08x' % int(w[0]) for w in got) == want


def test_philox_stream_offsets_and_mask_rate():
    a = philox.random_u32(64, seed=99, offset=0)
    b = philox.random_u32(60, seed=99, offset=4)
    assert np.array_equal(a[4:], b)                     # the stream is a pure function of the index
    c = philox.random_u32(61, seed=99, offset=3)
    assert np.array_equal(a[3:], c)                     # also at unaligned offsets
    m = philox.dropout_mask(1 << 16, 0.9, seed=7)
    assert abs(m.mean() - 0.9) < 0.01
    assert philox.dropout_mask(100, 1.0, seed=7).all()


def test_dense():
    g = load_golden('dense')
    w, b = g['p._linear._w'], g['p._linear._b']
    y, z = O.dense_fwd(g['x'], w, b)
    close(y, g['y'])
    dx, dw, db = O.dense_bwd(g['x'], w, z, g['dy'])
    close(dx, g['dx']); close(dw, g['g._linear._w']); close(db, g['g._linear._b'])
    close(O.sgd_step(w, dw, 0.05), g['sgd._linear._w'])
    close(O.sgd_step(b, db, 0.05), g['sgd._linear._b'])
    # three Adam steps on the same batch
    m = v = np.zeros_like(w, dtype=np.float64)
    mb = vb = np.zeros_like(b, dtype=np.float64)
    wa, ba = w.astype(np.float64), b.astype(np.float64)
    for t in (1, 2, 3):
        y, z = O.dense_fwd(g['x'], wa, ba)
        _, dw, db = O.dense_bwd(g['x'], wa, z, g['dy'])
        wa, m, v = O.adam_step(wa, dw, m, v, t, 0.01)
        ba, mb, vb = O.adam_step(ba, db, mb, vb, t, 0.01)
    close(wa, g['adam3._linear._w'], rtol=1e-4, atol=1e-6)
    close(ba, g['adam3._linear._b'], rtol=1e-4, atol=1e-6)


def test_activations():
    g = load_golden('activations')
    y = O.softmax_fwd(g['sm_x'])
    close(y, g['sm_y'])
    close(O.softmax_bwd(y, g['sm_dy']), g['sm_dx'])     # closed form == Jacobian einsum
    close(O.relu_fwd(g['relu_x']), g['relu_y'])
    close(O.relu_bwd(g['relu_x'], g['relu_dy']), g['relu_dx'])
    assert (g['relu_dx'][0, :4] == g['relu_dy'][0, :4]).all()   # gradient passes at x == 0


def test_layernorm():
    g = load_golden('layernorm')
    for sfx, eps in (('', 1e-3), ('3', 1e-6)):
        gamma, beta = g[f'p{sfx}._gamma'], g[f'p{sfx}._beta']
        z, *_ = O.layernorm_fwd(g['x' + sfx], gamma, beta, eps)
        close(z, g['z' + sfx])
        dx, dgamma, dbeta = O.layernorm_bwd(g['x' + sfx], gamma, g['dz' + sfx], eps)
        close(dx, g['dx' + sfx], rtol=1e-4, atol=1e-5)  # closed form == Jacobian einsum
        close(dgamma, g[f'g{sfx}._gamma']); close(dbeta, g[f'g{sfx}._beta'])


def test_dropout_mask_injection():
    g = load_golden('dropout')
    close(O.dropout_apply(g['x'], g['mask'], 0.5), g['y'])
    close(O.dropout_apply(g['dy'], g['mask'], 0.5), g['dx'])
    close(O.dropout_apply(g['x'], g['mask2'], np.float32(0.9)), g['y2'])


@pytest.mark.parametrize('tag', ['self', 'cross'])
def test_mha(tag):
    g = load_golden('mha_' + tag)
    p = sub(g, 'p.')
    out, cache = O.mha_fwd(p, g['query'], g.get('kv'))
    close(out, g['out'])
    close(cache['prob'], g['scores'])
    (dq, dk, dv), grads = O.mha_bwd(p, cache, g['dy'])
    close(dq, g['dquery']); close(dk, g['dkey']); close(dv, g['dvalue'])
    for k in O.MHA_PARAMS:
        close(grads[k], g['g.' + k], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('norm', ['pre', 'post'])
@pytest.mark.parametrize('drop', ['nodrop', 'drop'])
def test_encoder(norm, drop):
    g = load_golden(f'encoder_{norm}_{drop}')
    p = sub(g, 'p.')
    masks = (g.get('mask1'), g.get('mask2'))
    keep = 0.75 if drop == 'drop' else 1.0
    out, c = O.encoder_fwd(p, g['q'], norm == 'pre', masks, keep)
    close(out, g['out'], rtol=1e-4, atol=1e-5)
    dx, grads = O.encoder_bwd(p, c, g['dy'], norm == 'pre', masks, keep)
    close(dx, g['dq'], rtol=1e-4, atol=1e-5)
    gg = sub(g, 'g.')
    assert set(gg) == set(grads)
    for k in gg:
        close(grads[k], gg[k], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('norm', ['pre', 'post'])
@pytest.mark.parametrize('drop', ['nodrop', 'drop'])
def test_decoder(norm, drop):
    g = load_golden(f'decoder_{norm}_{drop}')
    p = sub(g, 'p.')
    masks = (g.get('mask1'), g.get('mask2'), g.get('mask3'))
    keep = 0.75 if drop == 'drop' else 1.0
    out, c = O.decoder_fwd(p, g['q'], g['kv'], norm == 'pre', masks, keep)
    close(out, g['out'], rtol=1e-4, atol=1e-5)
    (dq, dkv), grads = O.decoder_bwd(p, c, g['dy'], norm == 'pre', masks, keep)
    close(dq, g['dq'], rtol=1e-4, atol=1e-5)
    close(dkv, g['dkv'], rtol=1e-4, atol=1e-5)
    gg = sub(g, 'g.')
    assert set(gg) == set(grads)
    for k in gg:
        close(grads[k], gg[k], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('tag', ['c3', 'c8', 'k5', 'k1'])
def test_conv(tag):
    g = load_golden('conv_' + tag)
    f, b = g['p._w'], g['p._b']
    y, z = O.conv_layer_fwd(g['x'], f, b)
    close(y, g['y'])
    dx, dw, db = O.conv_layer_bwd(g['x'], f, z, g['dy'])
    close(dx, g['dx']); close(dw, g['g._w']); close(db, g['g._b'])


def test_losses():
    g = load_golden('loss')
    close(O.mse_fwd(g['y'], g['t']), g['mse'])
    close(O.mse_bwd(g['y'], g['t']), g['mse_dy'])
    close(O.ce_fwd(g['prob'], g['onehot']), g['ce'])
    close(O.ce_bwd(g['prob'], g['onehot']), g['ce_dy'])


def test_optimizers():
    g = load_golden('optimizer')
    close(O.sgd_step(g['w0'], g['grads'][0], 0.1), g['sgd'])
    w = g['w0'].astype(np.float64)
    m = v = np.zeros_like(w)
    for t, grad in enumerate(g['grads'], start=1):
        w, m, v = O.adam_step(w, grad, m, v, t, 0.01)
        close(w, g['adam'][t - 1], rtol=1e-5, atol=1e-6)


def test_trainer_mlp_losses():
    g = load_golden('trainer_mlp')
    p = {k: v.astype(np.float64) for k, v in sub(g, 'p0.').items()}
    losses = []
    for _ in range(4):
        h, z0 = O.dense_fwd(g['x'], p['0._linear._w'], p['0._linear._b'])
        y, z1 = O.dense_fwd(h, p['1._linear._w'], p['1._linear._b'], activation='softmax')
        losses.append(O.ce_fwd(y, g['t']))
        dy = O.ce_bwd(y, g['t'])
        dh, dw1, db1 = O.dense_bwd(h, p['1._linear._w'], z1, dy, activation='softmax', y=y)
        _, dw0, db0 = O.dense_bwd(g['x'], p['0._linear._w'], z0, dh)
        for k, gr in (('1._linear._w', dw1), ('1._linear._b', db1), ('0._linear._w', dw0), ('0._linear._b', db0)):
            p[k] = O.sgd_step(p[k], gr, 1e-2)
    close(losses, g['losses'], rtol=1e-4, atol=1e-5)
    for k, v in sub(g, 'p1.').items():
        close(p[k], v, rtol=1e-4, atol=1e-5)


def _adam_all(p, grads, state, t, lr):
    for k in grads:
        p[k], m, v = O.adam_step(p[k], grads[k], *state.setdefault(k, (np.zeros_like(p[k]), np.zeros_like(p[k]))), t, lr)
        state[k] = (m, v)


def test_trainer_conv_adam_steps():
    """The reference's Trainer on [Conv2D(4,3), Conv2D(6,3)] + MSELoss + AdamOptimizer, 3 steps, end to end."""
    g = load_golden('trainer_conv')
    p = {k: v.astype(np.float64) for k, v in sub(g, 'p0.').items()}
    state, losses = {}, []
    for step in range(3):
        y0, z0 = O.conv_layer_fwd(g['x'], p['0._w'], p['0._b'])
        y1, z1 = O.conv_layer_fwd(y0, p['1._w'], p['1._b'])
        losses.append(O.mse_fwd(y1, g['t']))
        dy = O.mse_bwd(y1, g['t'])
        d0, dw1, db1 = O.conv_layer_bwd(y0, p['1._w'], z1, dy)
        _, dw0, db0 = O.conv_layer_bwd(g['x'], p['0._w'], z0, d0)
        _adam_all(p, {'1._w': dw1, '1._b': db1, '0._w': dw0, '0._b': db0}, state, step + 1, 1e-2)
    close(losses, g['losses'], rtol=1e-5, atol=1e-6)
    for k, v in sub(g, 'p1.').items():
        close(p[k], v, rtol=1e-4, atol=1e-5)


def test_trainer_encoder_adam_steps():
    """The reference's Trainer on [TransformerEncoder(pre-norm), TransformerEncoder(post-norm)] + MSELoss + Adam."""
    g = load_golden('trainer_encoder')
    p = [{k: v.astype(np.float64) for k, v in sub(g, f'p0.{i}.').items()} for i in range(2)]
    state, losses = [{}, {}], []
    for step in range(3):
        h0, c0 = O.encoder_fwd(p[0], g['x'], True)
        h1, c1 = O.encoder_fwd(p[1], h0, False)
        losses.append(O.mse_fwd(h1, g['t']))
        dy = O.mse_bwd(h1, g['t'])
        d1, g1 = O.encoder_bwd(p[1], c1, dy, False)
        _, g0 = O.encoder_bwd(p[0], c0, d1, True)
        _adam_all(p[1], g1, state[1], step + 1, 1e-2)
        _adam_all(p[0], g0, state[0], step + 1, 1e-2)
    close(losses, g['losses'], rtol=1e-5, atol=1e-6)
    for i in range(2):
        for k, v in sub(g, f'p1.{i}.').items():
            close(p[i][k], v, rtol=1e-4, atol=2e-5)


def test_causal_extension_matches_torch_sdpa():
    """SURVEY.md §8 f1 is beyond the reference (its mask argument is unusable), so the oracle's `causal=True` is pinned
    against an independent implementation: torch scaled_dot_product_attention(is_causal=True) and its autograd, fp64."""
    import torch
    from oracle import np_oracle as O
    rng = np.random.default_rng(5)
    b, s_, h, d = 2, 9, 3, 4
    dm = h * d
    p = {'_wq': rng.standard_normal((h, d, dm)) * 0.3, '_wk': rng.standard_normal((h, d, dm)) * 0.3,
         '_wv': rng.standard_normal((h, d, dm)) * 0.3, '_wo': rng.standard_normal((dm, h, d)) * 0.3,
         '_bq': rng.standard_normal((h, d)) * 0.1, '_bk': rng.standard_normal((h, d)) * 0.1,
         '_bv': rng.standard_normal((h, d)) * 0.1, '_bo': rng.standard_normal(dm) * 0.1}
    x = rng.standard_normal((b, s_, dm))
    dy = rng.standard_normal((b, s_, dm))
    out, cache = O.mha_fwd(p, x, causal=True)
    (dq, dk, dv), g = O.mha_bwd(p, cache, dy)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    q = torch.einsum('bsd,hkd->bhsk', tx, tp['_wq']) + tp['_bq'][None, :, None, :]
    k = torch.einsum('bsd,hkd->bhsk', tx, tp['_wk']) + tp['_bk'][None, :, None, :]
    v = torch.einsum('bsd,hkd->bhsk', tx, tp['_wv']) + tp['_bv'][None, :, None, :]
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
    tout = torch.einsum('bhsc,dhc->bsd', o, tp['_wo']) + tp['_bo']
    tout.backward(torch.tensor(dy))
    np.testing.assert_allclose(out, tout.detach().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dq + dk + dv, tx.grad.numpy(), rtol=1e-9, atol=1e-11)
    for name in O.MHA_PARAMS:
        np.testing.assert_allclose(g[name], tp[name].grad.numpy(), rtol=1e-9, atol=1e-11, err_msg=name)
    # the first query row only sees the first key: its output is v[0] projected
    assert np.allclose(cache['prob'][:, :, 0, 1:], 0.0) and np.allclose(cache['prob'][:, :, 0, 0], 1.0)
