"""Philox4x32-10 counter-based RNG, NumPy restatement (TEST INFRASTRUCTURE ONLY).

The reference draws corrupt entities with host-side Python ``random.choice`` over a
per-step subsample (holE.py:343-347) and ``tf.random_uniform`` (holE.py:108-110,
127-129, 137).  Neither stream is reproducible without TensorFlow 1.2, so the build
replaces them (BASELINE.json north_star) with a counter-based Philox draw whose
*marginal distribution* equals the reference's.  This file is the
host-side statement of that draw; the CUDA sampler must match it bit for bit.

Algorithm: Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11),
Philox-4x32 with 10 rounds.  Known-answer vectors from the Random123 distribution are
checked in tests/test_philox.py.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays.

    Returns four uint32 arrays (the four output words).
    """
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(
        *(np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3, k0, k1)))
    c0 = c0.copy(); c1 = c1.copy(); c2 = c2.copy(); c3 = c3.copy()
    k0 = k0.copy(); k1 = k1.copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> _S32).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> _S32).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = k0 + _W0
            k1 = k1 + _W1
    return c0, c1, c2, c3


def _split64(x):
    x = np.asarray(x, dtype=np.uint64)
    return (x & _MASK).astype(np.uint32), (x >> _S32).astype(np.uint32)


def mulhi64_u32(r_lo, r_hi, cnt):
    """floor(((r_hi<<32 | r_lo) * cnt) / 2**64) for cnt < 2**32, in uint64 arithmetic."""
    cnt = np.asarray(cnt, dtype=np.uint64)
    lo = r_lo.astype(np.uint64) * cnt
    hi = r_hi.astype(np.uint64) * cnt
    return (hi + (lo >> _S32)) >> _S32


#: stream tag placed in counter word 3's high bits to separate the draws
STREAM_SIDE = 0x5EED0001
STREAM_ENTITY = 0x5EED0002


def side_coin(seed, step):
    """One fair coin per step: 1 = corrupt heads, 0 = corrupt tails (holE.py:137-140).

    counter = (step_lo, step_hi, 0, STREAM_SIDE), key = (seed_lo, seed_hi); bit 0 of word 0.
    """
    s_lo, s_hi = _split64(step)
    k_lo, k_hi = _split64(seed)
    w0, _, _, _ = philox4x32_10(s_lo, s_hi, np.uint32(0), np.uint32(STREAM_SIDE), k_lo, k_hi)
    return int(w0) & 1


def entity_draw(seed, step, index, count):
    """Uniform draw in [0, count) for batch slot ``index`` of global step ``step``.

    counter = (index, step_lo, step_hi, STREAM_ENTITY), key = (seed_lo, seed_hi);
    r64 = word1<<32 | word0; result = mulhi64(r64, count).
    """
    s_lo, s_hi = _split64(step)
    k_lo, k_hi = _split64(seed)
    idx = np.asarray(index, dtype=np.uint32)
    w0, w1, _, _ = philox4x32_10(idx, s_lo, s_hi, np.uint32(STREAM_ENTITY), k_lo, k_hi)
    return mulhi64_u32(w0, w1, count).astype(np.int64)
