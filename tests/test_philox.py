"""Philox4x32-10 host restatement vs the Random123 known-answer vectors."""
import numpy as np

from oracle import philox


def _hex(words):
    return [int(w) for w in words]


def test_known_answer_vectors():
    # Random123 kat_vectors: philox4x32 10 rounds
    assert _hex(philox.philox4x32_10(0, 0, 0, 0, 0, 0)) == \
        [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert _hex(philox.philox4x32_10(f, f, f, f, f, f)) == \
        [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _hex(philox.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344,
                                     0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_vectorised_matches_scalar():
    idx = np.arange(100, dtype=np.uint32)
    w = philox.philox4x32_10(idx, 7, 9, 11, 13, 17)
    for i in (0, 1, 50, 99):
        s = philox.philox4x32_10(i, 7, 9, 11, 13, 17)
        assert [int(x[i]) for x in w] == [int(x) for x in s]


def test_mulhi_matches_python_ints():
    rng = np.random.default_rng(0)
    lo = rng.integers(0, 2**32, size=200, dtype=np.uint64).astype(np.uint32)
    hi = rng.integers(0, 2**32, size=200, dtype=np.uint64).astype(np.uint32)
    cnt = rng.integers(1, 2**32, size=200, dtype=np.uint64)
    got = philox.mulhi64_u32(lo, hi, cnt)
    for l, h, c, g in zip(lo, hi, cnt, got):
        assert ((int(h) << 32 | int(l)) * int(c)) >> 64 == int(g)


def test_entity_draw_range_and_uniformity():
    d = philox.entity_draw(123, 5, np.arange(200000), 10)
    assert d.min() == 0 and d.max() == 9
    counts = np.bincount(d, minlength=10)
    assert np.all(np.abs(counts - 20000) < 600)       # ~4 sigma
    # singleton type always returns the only member
    assert np.all(philox.entity_draw(1, 2, np.arange(50), 1) == 0)
    # step and seed change the stream
    assert not np.array_equal(d[:1000], philox.entity_draw(123, 6, np.arange(1000), 10))
    assert not np.array_equal(d[:1000], philox.entity_draw(124, 5, np.arange(1000), 10))


def test_side_coin_is_fair():
    s = [philox.side_coin(99, step) for step in range(2000)]
    assert 900 < sum(s) < 1100
