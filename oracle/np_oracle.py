"""oracle/np_oracle.py — CPU restatement of np-modeling's layer forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under np-modeling_b200/ imports this; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the
checker or the CPU baseline — never as the product path.

What it is: the reference's algorithm (levendlee/np-modeling, pure NumPy) restated as pure
functions over float64 arrays and parameter dicts keyed by the reference's attribute names.  Each
function cites the reference file:line it follows.  Two deliberate differences from the reference
code, both result-equivalent:
  * Softmax.backward and LayerNormalization.backward use the closed forms instead of materialising
    the [..., n, n] Jacobians (activations.py:42-45, normalizations.py:60-71) — the Jacobian route
    needs 3 TiB at BASELINE cfg3.  tests/test_oracle.py checks the closed forms against golden
    vectors produced by the reference's Jacobian code.
  * Everything is float64 (the reference drifts between float32 and float64 under NumPy >= 2,
    SURVEY.md appendix A.13); comparisons cast to float32.

Pinning: tests/golden/*.npz hold inputs/outputs/gradients/updated parameters produced by the
UNMODIFIED reference (oracle/make_golden.py imports it from /root/reference); tests/test_oracle.py
requires this restatement to reproduce every one of them.
"""
import numpy as np

F64 = np.float64


def _f(x):
    return np.asarray(x, dtype=F64)


def _es(spec, *ops):
    """np.einsum as the reference calls it, routed through BLAS (optimize=True) so the CPU baseline
    is not handicapped by einsum's scalar loops."""
    return np.einsum(spec, *ops, optimize=True)


# ----------------------------------------------------------------------------- Linear / Dense
def linear_fwd(x, w, b):
    """y = x @ w + b                                                    layers/mlp.py:21-25"""
    return _f(x) @ _f(w) + _f(b)


def linear_bwd(x, w, dy):
    """(dx, dw, db) = (dy @ w^T, x^T @ dy, sum_m dy)                    layers/mlp.py:27-40"""
    x, w, dy = _f(x), _f(w), _f(dy)
    return dy @ w.T, x.T @ dy, dy.sum(axis=0)


def relu_fwd(x):
    """max(x, 0)                                                layers/activations.py:13-15"""
    return np.maximum(_f(x), 0.0)


def relu_bwd(x, dy):
    """where(x >= 0, dy, 0) — passes the gradient at x == 0      layers/activations.py:17-19"""
    return np.where(_f(x) >= 0.0, _f(dy), 0.0)


def softmax_fwd(x):
    """max-shifted exp / sum over the last axis                 layers/activations.py:23-31"""
    x = _f(x)
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def softmax_bwd(y, dy):
    """dx_b = sum_a dy_a y_a (delta_ab - y_b) = y * (dy - sum(dy*y)) — the einsum of the Jacobian
    built at layers/activations.py:33-45, in closed form."""
    y, dy = _f(y), _f(dy)
    return y * (dy - (dy * y).sum(axis=-1, keepdims=True))


def dense_fwd(x, w, b, activation='relu'):
    """Linear then activation (default ReLU)                           layers/mlp.py:70-72"""
    z = linear_fwd(x, w, b)
    return (relu_fwd(z) if activation == 'relu' else softmax_fwd(z)), z


def dense_bwd(x, w, z, dy, activation='relu', y=None):
    """activation.backward then Linear.backward                        layers/mlp.py:74-77"""
    dz = relu_bwd(z, dy) if activation == 'relu' else softmax_bwd(y, dy)
    return linear_bwd(x, w, dz)


# ----------------------------------------------------------------------------- normalisation
def dropout_apply(x, mask, keep_prob):
    """where(mask, x / keep, 0) — forward and backward are the same map
    layers/normalizations.py:14-30"""
    return np.where(np.asarray(mask) != 0, _f(x) / keep_prob, 0.0)


def layernorm_fwd(x, gamma, beta, eps=1e-3):
    """biased variance over the last axis; returns (out, y_hat, mean, var)
    layers/normalizations.py:43-48"""
    x = _f(x)
    mean = x.mean(axis=-1, keepdims=True)
    var = x.var(axis=-1, keepdims=True)
    yhat = (x - mean) / np.sqrt(var + eps)
    return _f(gamma) * yhat + _f(beta), yhat, mean, var


def layernorm_bwd(x, gamma, dz, eps=1e-3):
    """(dx, dgamma, dbeta).  dgamma/dbeta as layers/normalizations.py:55-56; dx is the closed form
    of dl_dy @ dy_dx with the Jacobian of :60-71:
        dx = (g - mean(g) - y_hat * mean(g * y_hat)) / sqrt(var + eps),  g = dz * gamma."""
    x, gamma, dz = _f(x), _f(gamma), _f(dz)
    _, yhat, _, var = layernorm_fwd(x, gamma, np.zeros_like(gamma), eps)
    batch_axes = tuple(range(x.ndim - 1))
    dbeta = dz.sum(axis=batch_axes)
    dgamma = (dz * yhat).sum(axis=batch_axes)
    g = dz * gamma
    dx = (g - g.mean(axis=-1, keepdims=True) - yhat * (g * yhat).mean(axis=-1, keepdims=True)) / np.sqrt(var + eps)
    return dx, dgamma, dbeta


# ----------------------------------------------------------------------------- attention
MHA_PARAMS = ('_wq', '_wk', '_wv', '_wo', '_bq', '_bk', '_bv', '_bo')


def mha_fwd(p, query, key=None, value=None, causal=False):
    """MultiHeadAttention.forward, unmasked (the reference's only usable mode).  `causal=True` is the extension of
    SURVEY.md §8 f1 — scores of key position t > query position s are set to -inf before the softmax, which is what the
    reference's `np.where(mask, scores, -np.inf)` (attentions.py:106-107) would do with a lower-triangular mask if its
    `if mask:` test accepted arrays; pinned against torch's scaled_dot_product_attention(is_causal=True) in tests.  p: dict with _wq,_wk [H,dk,D], _wv [H,dv,D],
    _wo [D,H,dv], _bq,_bk [H,dk], _bv [H,dv], _bo [D].  Returns (out, cache).
    layers/attentions.py:67-120"""
    query = _f(query)
    key = query if key is None else _f(key)
    value = key if value is None else _f(value)
    wq, wk, wv, wo = (_f(p[k]) for k in ('_wq', '_wk', '_wv', '_wo'))
    dk = wq.shape[1]
    q = _es('bsd,hkd->bshk', query, wq) + _f(p['_bq'])          # :90-92
    k = _es('btd,hkd->bthk', key, wk) + _f(p['_bk'])            # :94-96
    v = _es('btd,hcd->bthc', value, wv) + _f(p['_bv'])          # :98-100
    s = _es('bshk,bthk->bhst', q, k) / np.sqrt(dk)              # :103-104
    if causal:
        assert s.shape[-1] == s.shape[-2], 'causal mask needs Sq == Skv'
        s = np.where(np.tril(np.ones(s.shape[-2:], dtype=bool)), s, -np.inf)   # :106-107 with a lower-triangular mask
    prob = softmax_fwd(s)                                             # :108
    vals = _es('bhst,bthc->bhsc', prob, v)                      # :112
    out = _es('bhsc,dhc->bsd', vals, wo) + _f(p['_bo'])         # :116-117
    cache = dict(query=query, key=key, value=value, q=q, k=k, v=v, prob=prob, vals=vals)
    return out, cache


def mha_bwd(p, cache, dy):
    """MultiHeadAttention.backward → ((dquery, dkey, dvalue), grads dict).
    layers/attentions.py:122-199"""
    dy = _f(dy)
    wq, wk, wv, wo = (_f(p[k]) for k in ('_wq', '_wk', '_wv', '_wo'))
    dk_dim = wq.shape[1]
    c = cache
    g = {}
    g['_bo'] = dy.sum(axis=(0, 1))                                    # :129
    g['_wo'] = _es('bhsc,bsd->dhc', c['vals'], dy)              # :133-135 (batch-summed)
    dvals = _es('bsd,dhc->bhsc', dy, wo)                        # :136
    dprob = _es('bhsc,bthc->bhst', dvals, c['v'])               # :146
    dv = _es('bhst,bhsc->bthc', c['prob'], dvals)               # :147-148
    ds = softmax_bwd(c['prob'], dprob) / np.sqrt(dk_dim)              # :150-155
    dq = _es('bhst,bthk->bshk', ds, c['k'])                     # :161
    dk = _es('bshk,bhst->bthk', c['q'], ds)                     # :162
    g['_wq'] = _es('bsd,bshk->hkd', c['query'], dq)             # :169-171
    dquery = _es('bshk,hkd->bsd', dq, wq)                       # :172
    g['_wk'] = _es('btd,bthk->hkd', c['key'], dk)               # :173-175
    dkey = _es('bthk,hkd->btd', dk, wk)                         # :176
    g['_wv'] = _es('btd,bthc->hcd', c['value'], dv)             # :181-183
    dvalue = _es('bthc,hcd->btd', dv, wv)                       # :184
    g['_bq'] = dq.sum(axis=(0, 1))                                    # :186
    g['_bk'] = dk.sum(axis=(0, 1))                                    # :187
    g['_bv'] = dv.sum(axis=(0, 1))                                    # :188
    return (dquery, dkey, dvalue), g


# ----------------------------------------------------------------------------- transformer blocks
def _sub(p, prefix):
    return {k[len(prefix):]: v for k, v in p.items() if k.startswith(prefix)}


def _ffn_fwd(p, x2d):
    h, z = dense_fwd(x2d, p['_dense1._linear._w'], p['_dense1._linear._b'])
    return linear_fwd(h, p['_dense2._w'], p['_dense2._b']), (x2d, z, h)


def _ffn_bwd(p, cache, dy, g):
    x2d, z, h = cache
    dh, g['_dense2._w'], g['_dense2._b'] = linear_bwd(h, p['_dense2._w'], dy)
    dx, g['_dense1._linear._w'], g['_dense1._linear._b'] = dense_bwd(x2d, p['_dense1._linear._w'], z, dh)
    return dx


def _ln_fwd(p, name, x, eps):
    out, *_ = layernorm_fwd(x, p[name + '._gamma'], p[name + '._beta'], eps)
    return out


def _ln_bwd(p, name, x, dz, g, eps):
    dx, g[name + '._gamma'], g[name + '._beta'] = layernorm_bwd(x, p[name + '._gamma'], dz, eps)
    return dx


def _drop(x, mask, keep):
    return x if mask is None else dropout_apply(x, mask, keep)


def encoder_fwd(p, x, norm_first, masks=(None, None), keep_prob=1.0, eps=1e-3, causal=False):
    """TransformerEncoder.forward.  p keys: '_self_attention._wq', '_norm1._gamma', '_dense1._w',
    '_dense2._w', ... ; masks: dropout masks for (_dropout1, _dropout2) or None.
    layers/transformer.py:29-62"""
    x = _f(x)
    b, s, d = x.shape
    c = {}
    skip = x
    h = x
    if norm_first:
        c['ln1_in'] = _drop(h, masks[0], keep_prob)                    # :35-37 Dropout THEN LayerNorm
        h = _ln_fwd(p, '_norm1', c['ln1_in'], eps)
    out, c['att'] = mha_fwd(_sub(p, '_self_attention.'), h, causal=causal)           # :38
    out = out + skip                                                  # :39
    if not norm_first:
        c['ln1_in'] = _drop(out, masks[0], keep_prob)                  # :40-42
        out = _ln_fwd(p, '_norm1', c['ln1_in'], eps)
    out = out.reshape(-1, d)                                          # :45
    skip = out
    h = out
    if norm_first:
        c['ln2_in'] = _drop(h, masks[1] if masks[1] is None else np.reshape(masks[1], h.shape), keep_prob)
        h = _ln_fwd(p, '_norm2', c['ln2_in'], eps)                    # :48-50
    h, c['ffn'] = _ffn_fwd(p, h)                                      # :51-52
    out = h + skip                                                    # :53
    if not norm_first:
        c['ln2_in'] = _drop(out, masks[1] if masks[1] is None else np.reshape(masks[1], out.shape), keep_prob)
        out = _ln_fwd(p, '_norm2', c['ln2_in'], eps)                  # :54-56
    return out.reshape(b, s, d), c


def encoder_bwd(p, c, dy, norm_first, masks=(None, None), keep_prob=1.0, eps=1e-3):
    """TransformerEncoder.backward → (dx, grads).                layers/transformer.py:64-92"""
    dy = _f(dy)
    b, s, d = dy.shape
    g = {}
    dy = dy.reshape(-1, d)
    m2 = None if masks[1] is None else np.reshape(masks[1], dy.shape)
    if not norm_first:
        dy = _drop(_ln_bwd(p, '_norm2', c['ln2_in'], dy, g, eps), m2, keep_prob)     # :70-72
    dskip = dy
    dy = _ffn_bwd(p, c['ffn'], dy, g)                                                # :74-75
    if norm_first:
        dy = _drop(_ln_bwd(p, '_norm2', c['ln2_in'], dy, g, eps), m2, keep_prob)     # :76-78
    dy = (dy + dskip).reshape(b, s, d)                                               # :80-81
    if not norm_first:
        dy = _drop(_ln_bwd(p, '_norm1', c['ln1_in'], dy, g, eps), masks[0], keep_prob)
    dskip = dy
    (dq, dk, dv), ga = mha_bwd(_sub(p, '_self_attention.'), c['att'], dy)            # :87
    g.update({'_self_attention.' + k: v for k, v in ga.items()})
    dy = dq + dk + dv                                                                # :88
    if norm_first:
        dy = _drop(_ln_bwd(p, '_norm1', c['ln1_in'], dy, g, eps), masks[0], keep_prob)
    return dy + dskip, g                                                             # :92


def decoder_fwd(p, q, kv, norm_first, masks=(None, None, None), keep_prob=1.0, eps=1e-3, causal=False):
    """TransformerDecoder.forward (unmasked self-attention, cross-attention on kv).
    layers/transformer.py:119-160"""
    q, kv = _f(q), _f(kv)
    b, s, d = q.shape
    c = {'kv': kv}
    skip = q
    h = q
    if norm_first:
        c['ln1_in'] = _drop(h, masks[0], keep_prob)                    # :125-127
        h = _ln_fwd(p, '_norm1', c['ln1_in'], eps)
    out, c['self'] = mha_fwd(_sub(p, '_self_attention.'), h, causal=causal)          # :128 (causal: §8 f1 extension)
    out = out + skip
    if not norm_first:
        c['ln1_in'] = _drop(out, masks[0], keep_prob)                  # :130-132
        out = _ln_fwd(p, '_norm1', c['ln1_in'], eps)
    skip = out
    h = out
    if norm_first:
        c['ln2_in'] = _drop(h, masks[1], keep_prob)                    # :136-138
        h = _ln_fwd(p, '_norm2', c['ln2_in'], eps)
    out, c['cross'] = mha_fwd(_sub(p, '_cross_attention.'), h, kv)    # :139
    out = out + skip
    if not norm_first:
        c['ln2_in'] = _drop(out, masks[1], keep_prob)                  # :141-143
        out = _ln_fwd(p, '_norm2', c['ln2_in'], eps)
    out = out.reshape(-1, d)                                          # :146
    skip = out
    h = out
    m3 = None if masks[2] is None else np.reshape(masks[2], out.shape)
    if norm_first:
        c['ln3_in'] = _drop(h, m3, keep_prob)                          # :149-151
        h = _ln_fwd(p, '_norm3', c['ln3_in'], eps)
    h, c['ffn'] = _ffn_fwd(p, h)                                      # :152-153
    out = h + skip
    if not norm_first:
        c['ln3_in'] = _drop(out, m3, keep_prob)                        # :155-157
        out = _ln_fwd(p, '_norm3', c['ln3_in'], eps)
    return out.reshape(b, s, d), c


def decoder_bwd(p, c, dy, norm_first, masks=(None, None, None), keep_prob=1.0, eps=1e-3):
    """TransformerDecoder.backward → ((dq, dkv), grads).        layers/transformer.py:162-203"""
    dy = _f(dy)
    b, s, d = dy.shape
    g = {}
    dy = dy.reshape(-1, d)
    m3 = None if masks[2] is None else np.reshape(masks[2], dy.shape)
    if not norm_first:
        dy = _drop(_ln_bwd(p, '_norm3', c['ln3_in'], dy, g, eps), m3, keep_prob)     # :166-168
    dskip = dy
    dy = _ffn_bwd(p, c['ffn'], dy, g)                                                # :170-171
    if norm_first:
        dy = _drop(_ln_bwd(p, '_norm3', c['ln3_in'], dy, g, eps), m3, keep_prob)     # :172-174
    dy = (dy + dskip).reshape(b, s, d)                                               # :176-177
    if not norm_first:
        dy = _drop(_ln_bwd(p, '_norm2', c['ln2_in'], dy, g, eps), masks[1], keep_prob)
    dskip = dy
    (dq_, dk_, dv_), ga = mha_bwd(_sub(p, '_cross_attention.'), c['cross'], dy)      # :183
    g.update({'_cross_attention.' + k: v for k, v in ga.items()})
    dkv = dk_ + dv_                                                                  # :184
    dy = dq_                                                                         # :185
    if norm_first:
        dy = _drop(_ln_bwd(p, '_norm2', c['ln2_in'], dy, g, eps), masks[1], keep_prob)
    dy = dy + dskip                                                                  # :190
    if not norm_first:
        dy = _drop(_ln_bwd(p, '_norm1', c['ln1_in'], dy, g, eps), masks[0], keep_prob)
    dskip = dy
    (dq_, dk_, dv_), ga = mha_bwd(_sub(p, '_self_attention.'), c['self'], dy)        # :195
    g.update({'_self_attention.' + k: v for k, v in ga.items()})
    dy = dq_ + dk_ + dv_                                                             # :196
    if norm_first:
        dy = _drop(_ln_bwd(p, '_norm1', c['ln1_in'], dy, g, eps), masks[0], keep_prob)
    return (dy + dskip, dkv), g                                                      # :201-203


# ----------------------------------------------------------------------------- convolution
def conv2d_fwd(x, f):
    """SAME, stride 1, odd k; NHWC x HWIO → NHWC, as k*k shifted GEMMs.
    layers/conv.py:74-107"""
    x, f = _f(x), _f(f)
    n, h, w, c0 = x.shape
    k, _, _, c1 = f.shape
    assert k % 2 == 1 and f.shape[1] == k and f.shape[2] == c0
    pad = k // 2
    xp = np.zeros((n, h + 2 * pad, w + 2 * pad, c0))
    xp[:, pad:pad + h, pad:pad + w, :] = x
    y = np.zeros((n, h, w, c1))
    for i in range(k):
        for j in range(k):
            y += xp[:, i:i + h, j:j + w, :].reshape(-1, c0).dot(f[i, j]).reshape(n, h, w, c1)
    return y


def conv2d_bwd_x(dy, f):
    """dx = conv(dy, flipHW(f) with I/O swapped)                layers/conv.py:110-153"""
    return conv2d_fwd(dy, np.transpose(_f(f)[::-1, ::-1], (0, 1, 3, 2)))


def conv2d_bwd_w(dy, x, k):
    """dw[i,j] = xpad[:, i:i+H, j:j+W, :]^T @ dy                layers/conv.py:156-194"""
    dy, x = _f(dy), _f(x)
    n, h, w, c1 = dy.shape
    c0 = x.shape[3]
    pad = k // 2
    xp = np.zeros((n, h + 2 * pad, w + 2 * pad, c0))
    xp[:, pad:pad + h, pad:pad + w, :] = x
    dw = np.zeros((k, k, c0, c1))
    d2 = dy.reshape(-1, c1)
    for i in range(k):
        for j in range(k):
            dw[i, j] = xp[:, i:i + h, j:j + w, :].reshape(-1, c0).T.dot(d2)
    return dw


def conv_layer_fwd(x, f, b):
    """Conv2D.forward with the default ReLU → (y, z)              layers/conv.py:44-48"""
    z = conv2d_fwd(x, f) + _f(b)
    return relu_fwd(z), z


def conv_layer_bwd(x, f, z, dy):
    """Conv2D.backward → (dx, dw, db)                             layers/conv.py:50-61"""
    dz = relu_bwd(z, dy)
    return conv2d_bwd_x(dz, f), conv2d_bwd_w(dz, x, f.shape[0]), dz.sum(axis=(0, 1, 2))


# ----------------------------------------------------------------------------- losses
def mse_fwd(y, t):
    """sum((y-t)^2) / y.size                                              loss.py:21-25"""
    y, t = _f(y), _f(t)
    return float(((y - t) ** 2).sum() / y.size)


def mse_bwd(y, t):
    """2 (y-t) / y.size                                                   loss.py:27-29"""
    y, t = _f(y), _f(t)
    return 2.0 * (y - t) / y.size


def ce_fwd(y, t):
    """-sum(t log y) on probabilities, summed over the batch              loss.py:33-36"""
    return float(-(_f(t) * np.log(_f(y))).sum())


def ce_bwd(y, t):
    """-t / y                                                             loss.py:38-39"""
    return -_f(t) / _f(y)


# ----------------------------------------------------------------------------- optimizers
def sgd_step(var, grad, lr):
    """var - lr * grad                                              optimizer.py:30-33"""
    return _f(var) - lr * _f(grad)


def adam_step(var, grad, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """One Adam update at step t (t starts at 1), epsilon INSIDE the sqrt → (var, m, v).
    optimizer.py:50-69"""
    var, grad, m, v = _f(var), _f(grad), _f(m), _f(v)
    m = beta1 * m + (1 - beta1) * grad
    v = beta2 * v + (1 - beta2) * grad ** 2
    mh = m / (1 - beta1 ** t)
    vh = v / (1 - beta2 ** t)
    return var - lr * (mh / np.sqrt(vh + eps)), m, v
