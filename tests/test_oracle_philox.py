"""Pins oracle/philox.py: Random123 known-answer vectors for philox4x32_10 (Random123 kat_vectors), plus the
mask / modality-mask spec's structural properties."""
import numpy as np

from oracle import philox as px

KAT = [  # (counter, key) -> output, Random123 examples/kat_vectors "philox4x32 10"
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_random123_known_answers():
    for ctr, key, want in KAT:
        got = px.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_vectorised_matches_scalar():
    c0 = np.arange(7, dtype=np.uint64)
    got = px.philox4x32_10(c0, 5, 9, 3, 42, 0)
    for i in range(7):
        one = px.philox4x32_10(i, 5, 9, 3, 42, 0)
        assert [int(g[i]) for g in got] == [int(o) for o in one]


def test_element_words_layout():
    w = px.element_words(10, np.array([3, 4], dtype=np.uint64), 2, px.SITE_GATE, 42)
    assert w.shape == (2, 10)
    k0, k1 = px.seed_key(42)
    blk1 = px.philox4x32_10(1, 4, 2, px.SITE_GATE, k0, k1)
    assert int(w[1, 6]) == int(blk1[2])          # element 6 = block 1, word 2


def test_dropout_keep_rate_and_determinism():
    ids = np.arange(64, dtype=np.uint64)
    k1 = px.dropout_keep(4096, ids, 3, px.SITE_FRAME_MLP, 0.1, 42)
    k2 = px.dropout_keep(4096, ids, 3, px.SITE_FRAME_MLP, 0.1, 42)
    assert (k1 == k2).all()
    assert abs(k1.mean() - 0.9) < 0.005
    # batch-composition independence: a clip's mask does not depend on which batch it is in
    k3 = px.dropout_keep(4096, ids[10:20], 3, px.SITE_FRAME_MLP, 0.1, 42)
    assert (k3 == k1[10:20]).all()
    assert (px.dropout_keep(4096, ids, 4, px.SITE_FRAME_MLP, 0.1, 42) != k1).any()
    assert px.drop_threshold(0.0) == 0 and px.drop_threshold(0.5) == 1 << 31
    assert float(px.keep_scale(0.3)) == float(np.float32(1.0) / np.float32(0.7))


def test_modality_mask_rules():
    ids = np.arange(200000, dtype=np.uint64)
    m = px.modality_mask(ids, 2, 0.05, 42)
    assert m.shape == (200000, 2) and set(np.unique(m)) <= {0.0, 1.0}
    assert (m.sum(axis=1) >= 1).all()                       # src/data.py:115-118: never all missing
    assert abs((m == 0).mean() - (0.05 - 0.05 * 0.05 / 2)) < 0.002
    assert (px.modality_mask(ids[:10], 2, 0.0, 42) == 1).all()
    heavy = px.modality_mask(ids[:20000], 2, 0.9, 42)       # rescue path exercised
    assert (heavy.sum(axis=1) >= 1).all() and (heavy.sum(axis=1) == 1).mean() > 0.8
