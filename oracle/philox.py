"""oracle/philox.py — NumPy restatement of Philox-4x32-10 and of the DropOut mask rule.

TEST INFRASTRUCTURE ONLY (see np_oracle.py).  The reference's DropOut draws its mask from the legacy
global MT19937 stream (layers/normalizations.py:20), which a counter-based GPU generator cannot
reproduce; parity on masks is therefore (i) mask injection, as the reference's own test does
(normalizations_test.py:28), and (ii) THIS restatement of the generator the CUDA kernels use
(np-modeling_b200/csrc/common.cuh philox4x32_10), which must match bit for bit.

Pinned by the Random123 known-answer vectors (Salmon et al., SC'11, kat_vectors) in
tests/test_oracle.py.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over uint32 arrays c0..c3; scalar keys. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def random_u32(n, seed, offset=0):
    """u32 stream element g = offset+i is word (g & 3) of philox(counter = g >> 2, key = seed)."""
    g = np.arange(n, dtype=np.uint64) + np.uint64(offset)
    ctr = g >> np.uint64(2)
    words = philox4x32_10(ctr & _MASK, ctr >> np.uint64(32), 0 * ctr, 0 * ctr, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lane = (g & np.uint64(3)).astype(np.int64)
    return np.choose(lane, words)


def dropout_mask(n, keep_prob, seed, offset=0):
    """keep element i  <=>  u32_i < floor(float32(keep_prob) * 2^32)   (npm_dropout_fwd contract)."""
    thr = int(float(np.float32(keep_prob)) * 4294967296.0)
    return (random_u32(n, seed, offset).astype(np.uint64) < np.uint64(thr)).astype(np.uint8)
