"""SURVEY.md §8 f2 / f4 (beyond the reference, built from its own blocks): the GPT-shaped container (Embedding ->
N x TransformerDecoder -> LM head -> CrossEntropyLoss) against the oracle chained layer by layer, and decode-time
key/value caches against the full causal forward."""
import numpy as np
import pytest

from helpers import Recorder, close, grads_of, param_values, use_device_relu_gates

pytestmark = pytest.mark.gpu

DEC_PATHS = None


def _dec_paths(O):
    return ['_self_attention.' + k for k in O.MHA_PARAMS] + ['_cross_attention.' + k for k in O.MHA_PARAMS] + \
        ['_dense1._linear._w', '_dense1._linear._b', '_dense2._w', '_dense2._b'] + \
        [f'_norm{i}.{n}' for i in (1, 2, 3) for n in ('_gamma', '_beta')]


def _scale_weights(layer_obj):
    """Fan-in scaled weights so that a stack stays O(1)."""
    from train import iter_parameters
    for owner, name in iter_parameters(layer_obj):
        v = np.asarray(getattr(owner, name))
        if name.startswith('_w') and v.ndim >= 2:
            fan_in = v.shape[-1] if (v.ndim == 3 and name != '_wo') else (v.shape[1] * v.shape[2] if name == '_wo' else v.shape[0])
            setattr(owner, name, (v / np.sqrt(fan_in)).astype(np.float32))
        elif name in ('_gamma',):
            setattr(owner, name, np.ones_like(v))
        elif name.startswith('_b') or name == '_beta':
            setattr(owner, name, (0.1 * v).astype(np.float32))


@pytest.fixture(autouse=True, params=['bf16x3', '3xtf32'])
def _precision(request):
    import npm_b200
    npm_b200.set_precision(request.param)
    yield request.param
    npm_b200.set_precision('bf16x3')


@pytest.mark.parametrize('heads,features', [(2, 128), (2, 64)])      # head dim 64 (fused attention kernels) and 32 (GEMM chain)
def test_three_layer_gpt_stack_vs_chained_oracle(heads, features):
    import loss
    from layers.adapters import GPTStack
    from oracle import np_oracle as O
    rng = np.random.default_rng(31)
    vocab, L, hidden, b, s, skv = 40, 3, 2 * features, 2, 12, 9
    ids = rng.integers(0, vocab, size=(b, s))
    kv = rng.standard_normal((b, skv, features)).astype(np.float32)
    tgt = np.eye(vocab, dtype=np.float32)[rng.integers(0, vocab, size=(b, s))]
    np.random.seed(4)
    model = GPTStack(vocab, features, L, heads, hidden, True, 0.0, max_len=16, causal=True)
    model(ids, kv)
    _scale_weights(model)
    probs = model(ids, kv)

    # ---- oracle: the same computation chained out of np_oracle's layer functions
    table, pos = np.asarray(model._embed._w, dtype=np.float64), np.asarray(model._embed._pos, dtype=np.float64)
    x = table[ids] + pos[np.arange(s)][None]
    caches, ps = [], []
    for dec in model._stack._layers:
        p = param_values(dec, _dec_paths(O))
        ps.append(p)
        x, c = O.decoder_fwd(p, x, kv, True, causal=True)
        caches.append(c)
    hw, hb = np.asarray(model._head._linear._w), np.asarray(model._head._linear._b)
    x2 = x.reshape(b * s, features)
    oprobs, z = O.dense_fwd(x2, hw, hb, activation='softmax')
    close(probs, oprobs.reshape(b, s, vocab), rtol=1e-3, atol=1e-5)
    ce = loss.CrossEntropyLoss()
    close(float(ce(probs, tgt)), O.ce_fwd(oprobs.reshape(b, s, vocab), tgt), rtol=1e-4, atol=1e-4)

    rec = Recorder()
    dprobs = ce(backprop=True)
    _, dkv = model(dprobs, backprop=True, optimizer_=rec)
    odp = O.ce_bwd(oprobs.reshape(b, s, vocab), tgt).reshape(b * s, vocab)
    dx, odhw, odhb = O.dense_bwd(x2, hw, z, odp, activation='softmax', y=oprobs)
    dx = dx.reshape(b, s, features)
    odkv = 0.0
    ograds = [None] * L
    for i in reversed(range(L)):
        use_device_relu_gates(model._stack._layers[i], caches[i])
        (dx, d), ograds[i] = O.decoder_bwd(ps[i], caches[i], dx, True)
        odkv = odkv + d
    close(dkv, odkv)
    otable = np.zeros_like(table)
    np.add.at(otable, ids.reshape(-1), dx.reshape(b * s, features))
    opos = np.zeros_like(pos)
    np.add.at(opos, np.tile(np.arange(s), b), dx.reshape(b * s, features))
    g = grads_of(model._embed, rec, ['_w', '_pos'])
    close(g['_w'], otable); close(g['_pos'], opos)
    gh = grads_of(model._head, rec, ['_linear._w', '_linear._b'])
    close(gh['_linear._w'], odhw); close(gh['_linear._b'], odhb)
    for i, dec in enumerate(model._stack._layers):
        got = grads_of(dec, rec, ograds[i].keys())
        for k, v in ograds[i].items():
            close(got[k], v, rtol=1e-3, atol=1e-4 * max(1.0, np.abs(v).max()))


@pytest.mark.parametrize('norm_first', [True, False])
@pytest.mark.parametrize('heads,features', [(2, 128), (2, 64)])
def test_kv_cache_decode_equals_full_forward(norm_first, heads, features):
    """transformer.py:120 `# TODO: support cache`: decoding token by token with cached keys / values gives row t of the
    full causal forward pass, and the last row matches the oracle."""
    from layers.adapters import DecoderStack
    from oracle import np_oracle as O
    rng = np.random.default_rng(32)
    b, s, skv, L = 2, 11, 7, 2
    x = rng.standard_normal((b, s, features)).astype(np.float32)
    kv = rng.standard_normal((b, skv, features)).astype(np.float32)
    np.random.seed(8)
    stack = DecoderStack(L, heads, 2 * features, norm_first, 0.0, causal=True)
    stack(x, kv)
    _scale_weights(stack)
    full = np.asarray(stack(x, kv))
    cache = stack.new_cache(b, s)
    for t in range(s):
        step = np.asarray(stack.decode_step(x[:, t:t + 1], kv, cache))
        close(step[:, 0], full[:, t], rtol=1e-3, atol=2e-4)
    want = x.astype(np.float64)
    for dec in stack._layers:
        want, _ = O.decoder_fwd(param_values(dec, _dec_paths(O)), want, kv, norm_first, causal=True)
    close(full, want, rtol=1e-3, atol=2e-4)


def test_gpt_greedy_decode_matches_full_forward_probabilities():
    from layers.adapters import GPTStack
    rng = np.random.default_rng(33)
    vocab, features, b, s = 30, 128, 2, 8
    ids = rng.integers(0, vocab, size=(b, s))
    kv = rng.standard_normal((b, 5, features)).astype(np.float32)
    np.random.seed(9)
    model = GPTStack(vocab, features, 2, 2, 256, True, 0.0, max_len=16, causal=True)
    model(ids, kv)
    _scale_weights(model)
    probs = np.asarray(model(ids, kv))
    cache = model.new_cache(b, s)
    for t in range(s):
        p_t = np.asarray(model.decode_step(ids[:, t:t + 1], kv, cache))
        close(p_t, probs[:, t], rtol=1e-3, atol=1e-5)
        assert (p_t.argmax(-1) == probs[:, t].argmax(-1)).all()


def test_embedding_gather_is_exact_and_clamps():
    from layers.adapters import Embedding
    rng = np.random.default_rng(34)
    emb = Embedding(17, 20)
    ids = rng.integers(0, 17, size=(3, 9))
    emb(ids)
    table = np.asarray(emb._w)
    np.testing.assert_array_equal(np.asarray(emb(ids)), table[ids])        # integer indexing: bit-exact
    rec = Recorder()
    dy = rng.standard_normal((3, 9, 20)).astype(np.float32)
    emb(dy, backprop=True, optimizer_=rec)
    want = np.zeros_like(table, dtype=np.float64)
    np.add.at(want, ids.reshape(-1), dy.reshape(-1, 20).astype(np.float64))
    close(grads_of(emb, rec, ['_w'])['_w'], want, rtol=1e-5, atol=1e-5)
