"""Philox-4x32-10 and the specs built on it.  TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference draws dropout masks from torch's global CPU mt19937 stream
(src/uncertainty.py:61-65, no generator argument; src/train.py:472 is the only
seeding), which cannot be reproduced on a GPU.  SURVEY.md F6: "masks bit-exact
for a given seed" therefore means one written counter-based spec implemented
twice - here (numpy) and in csrc/philox.cuh.  Algorithm: Philox-4x32-10 of
Salmon et al., "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11), the
Random123 library's philox4x32_R(10, ...).  Pinned by Random123's
known-answer vectors in tests/test_oracle_philox.py.

Spec (all little pieces are pure functions of their arguments, so results do
not depend on batch size, sharding or launch geometry):

    words(c0, c1, c2, c3; seed) = philox4x32_10(ctr=(c0,c1,c2,c3),
                                               key=(seed & 0xffffffff, seed >> 32))

    dropout keep bit for element e of clip g, MC pass t, site s, drop-prob p:
        keep = words(e >> 2, g, t, s)[e & 3] >= floor(p * 2**32)
        (dense sites, >= 1024 mask elements per clip and pass - the (F, H) frame_mlp mask: 16-bit draws,
         keep = half(e & 1) of words(e >> 3, g, t, s)[(e >> 1) & 3] >= floor(p * 2**32) >> 16; see dropout_keep)
        kept activations are multiplied by fp32(1) / fp32(1 - p)   (nn.Dropout)

Site ids (element index is the row-major index inside
the per-clip mask tensor):
    SITE_FRAME_MLP   0   encoders.video.frame_mlp.2   (F, H)      src/encoders.py:228
    SITE_FRAME_POOL  1   encoders.video._dropout      (H,)        src/encoders.py:230,296
    SITE_GATE        2   fusion_model.fusion_gate.2   (M, G)      src/fusion.py:448
    SITE_MODMASK     3   modality dropout             (M,)        src/data.py:111-118
    SITE_EARLY0      4   EarlyFusion mlp.3            (hidden,)   src/fusion.py:65
    SITE_EARLY1      5   EarlyFusion mlp.7            (hidden,)   src/fusion.py:69
    SITE_CNN0        6   SequenceEncoder(cnn).dropout_layer, 1st use  (H, T)   src/encoders.py:161
    SITE_CNN1        7   SequenceEncoder(cnn).dropout_layer, 2nd use  (H,)     src/encoders.py:167
    SITE_ATTN + i    16+i fusion_model.attn_layers[i].attn_dropout (heads, Tq, Tk)
                                                                  src/attention.py:51,130
    SITE_LATE + i    32+i LateFusion classifiers[i].2  (hidden,)   src/fusion.py:184
    SITE_UNC + 4i+j  48+  LateFusionWithUncertainty modality i: j = 0 classifiers[i][0] (dim,), 1 classifiers[i][3]
                          (hidden,), 2 uncert_heads[i][0] (dim,), 3 uncert_heads[i][3] (hidden,)   src/fusion.py:613-620
    SITE_TFM + 3l+j  64+  SequenceEncoder(transformer) layer l: j = 0 dropout1 (T, H), 1 dropout (T, 4H), 2 dropout2 (T, H)
                          torch/nn/modules/transformer.py (TransformerEncoderLayer), built at src/encoders.py:101-108
    SITE_PAIR_MSG + i 80+i PairwiseModalityAttention.msg_dropout for target modality i  (hidden,)  src/attention.py:468,536
Synthetic-data streams use c3 >= 0x100 (see oracle/synth.py).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

SITE_FRAME_MLP = 0
SITE_FRAME_POOL = 1
SITE_GATE = 2
SITE_MODMASK = 3
SITE_EARLY0 = 4
SITE_EARLY1 = 5
SITE_CNN0 = 6
SITE_CNN1 = 7
SITE_ATTN = 16
SITE_LATE = 32
SITE_UNC = 48
SITE_TFM = 64
SITE_PAIR_MSG = 80


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox-4x32-10.  Counter words broadcast against each other.

    Returns four uint32 arrays of the broadcast shape.
    """
    c0, c1, c2, c3 = np.broadcast_arrays(
        np.asarray(c0, dtype=np.uint64), np.asarray(c1, dtype=np.uint64),
        np.asarray(c2, dtype=np.uint64), np.asarray(c3, dtype=np.uint64))
    c0 = c0 & MASK32
    c1 = c1 & MASK32
    c2 = c2 & MASK32
    c3 = c3 & MASK32
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0            # 32x32 -> 64 bit, exact in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1,
                          hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def seed_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def element_words(n_elems: int, c1, c2, c3, seed: int) -> np.ndarray:
    """uint32 word for elements e = 0..n_elems-1: words(e>>2, c1, c2, c3)[e&3].

    c1 may be an array of shape (B,) (clip indices); result is (B, n_elems)
    in that case, else (n_elems,).
    """
    k0, k1 = seed_key(seed)
    nblk = (n_elems + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    if c1.ndim == 0:
        w = philox4x32_10(blk, c1, c2, c3, k0, k1)
        return np.stack(w, axis=-1).reshape(-1)[:n_elems]
    w = philox4x32_10(blk[None, :], c1[:, None], c2, c3, k0, k1)
    return np.stack(w, axis=-1).reshape(c1.shape[0], -1)[:, :n_elems]


def drop_threshold(p: float) -> int:
    """floor(p * 2**32) clamped to uint32; an element is KEPT iff word >= threshold."""
    return max(0, min(0xFFFFFFFF, int(float(p) * 4294967296.0)))


def keep_scale(p: float) -> np.float32:
    """nn.Dropout's multiplier for kept elements, in fp32: 1 / (1 - p)."""
    return np.float32(1.0) / np.float32(1.0 - float(p))


DENSE_SITE_ELEMS = 1024


def dropout_keep(n_elems: int, clip_ids, t: int, site: int, p: float, seed: int) -> np.ndarray:
    """Boolean keep mask, shape (B, n_elems), for MC pass t at a dropout site.

    Sites whose mask has fewer than DENSE_SITE_ELEMS elements per (clip, pass) draw one 32-bit word per element:
    word(e) = words(e >> 2)[e & 3], kept iff word >= floor(p * 2**32).  Dense sites (the (F, H) = 4096-element
    frame_mlp mask, 94 % of all draws) use 16 bits per element: half-word e & 1 (0 = low) of word (e >> 1) & 3 of
    block e >> 3, kept iff half >= floor(p * 2**32) >> 16 = floor(p * 2**16); p is then realised to within 2**-16."""
    thr = drop_threshold(p)
    if n_elems < DENSE_SITE_ELEMS:
        w = element_words(n_elems, np.asarray(clip_ids), t, site, seed)
        return w >= np.uint32(thr)
    w = element_words((n_elems + 1) // 2, np.asarray(clip_ids), t, site, seed)      # word of element pair e >> 1
    halves = np.stack([w & np.uint32(0xFFFF), w >> np.uint32(16)], axis=-1).reshape(*w.shape[:-1], -1)[..., :n_elems]
    return halves >= np.uint32(thr >> 16)


def modality_mask(clip_ids, num_modalities: int, p: float, seed: int) -> np.ndarray:
    """Modality-availability mask (B, M) float32, 1 = available.

    Follows src/data.py:108-118: mask = ones(M); mask *= (rand(M) > p); if every
    modality was dropped, one uniformly chosen modality is switched back on.
    rand(M) is word * 2**-32 of words(0, g, 0, SITE_MODMASK)[m] (so M <= 3 here;
    word 3 picks the rescued modality).  "u > p" is evaluated on integers as
    word > floor(p * 2**32), which equals the real-number comparison for every
    p that is not a multiple of 2**-32.
    """
    assert 1 <= num_modalities <= 3
    clip_ids = np.asarray(clip_ids, dtype=np.uint64)
    k0, k1 = seed_key(seed)
    w = philox4x32_10(np.uint64(0), clip_ids, 0, SITE_MODMASK, k0, k1)
    w = np.stack(w, axis=-1)                                   # (B, 4)
    if p <= 0:
        return np.ones((clip_ids.shape[0], num_modalities), dtype=np.float32)
    keep = w[:, :num_modalities] > np.uint32(drop_threshold(p))
    none = ~keep.any(axis=1)
    rescue = (w[:, 3] % np.uint32(num_modalities)).astype(np.int64)
    keep[none, rescue[none]] = True
    return keep.astype(np.float32)
