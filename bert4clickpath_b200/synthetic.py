"""Synthetic clickstreams with the shapes of the reference's Cloze input contract
(examples/BERT4Rec/source/input_pipeline.py:59-133, :198-220): per session, TRAIN drops the last
item, masks n = clip(int(len * p), 0, max_masked) unique sorted positions with [MASK] and emits
the masked items' label-vocabulary ids (id - 10) as float32 labels padded with -1; EVAL masks the
last item only.  Item popularity is Zipf-like, session lengths either dense (= max_len) or drawn
from the Amazon-Beauty histogram.  Host-side NumPy, not on the timed path.
"""
import numpy as np

from .constants import CLS, LABEL_PAD, MASK_ID, NUM_RESERVED_TOKENS, SEP

# session-length histogram of beauty.txt after the "first 50 per user" cut, lengths 5..50
BEAUTY_LEN_HIST = np.array(
    [12832, 7587, 4867, 3234, 2327, 1727, 1337, 1070, 789, 614, 489, 409, 365, 280, 257, 205, 184,
     154, 134, 126, 103, 91, 81, 60, 71, 53, 57, 55, 44, 32, 29, 29, 42, 31, 24, 23, 25, 27, 21, 17,
     18, 15, 11, 11, 16, 253], dtype=np.float64)


def n_masked_for(length, masked_percentage, max_masked):
    """clip(int(float32(len) * p), 0, max)  (input_pipeline.py:68-70)."""
    return int(min(max(int(np.float32(length) * np.float32(masked_percentage)), 0), max_masked))


def _rank_to_item(vocab):
    """Multiplier of the bijection rank -> (rank * A) % vocab (golden-ratio hashing, gcd(A, vocab)
    = 1) that scatters popularity ranks over the id space.  Real catalogues are not sorted by
    popularity; with ids in popularity order a model that has learned the popularity prior emits
    scores that DECREASE along the vocabulary, and every threshold-based top-k degenerates to its
    no-candidate fast path (the first k ids win), which flatters the inference benchmark."""
    from math import gcd
    a = max(1, int(0.6180339887498949 * vocab))
    while gcd(a, vocab) != 1:
        a += 1
    return a


def zipf_items(rng, size, vocab, s=0.8):
    """Item ids in [10, vocab+9] with P(popularity rank r) ~ r^-s; ranks are scattered over the id
    space by a fixed bijection (see _rank_to_item)."""
    w = 1.0 / np.power(np.arange(1, vocab + 1, dtype=np.float64), s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    rank = np.searchsorted(cdf, rng.random(size)).astype(np.int64)
    item = (rank * _rank_to_item(vocab)) % vocab
    return (item + NUM_RESERVED_TOKENS).astype(np.int32)


def make_cloze_batch(rng, batch, vocab, max_len=50, mode="train", masked_percentage=0.15,
                     max_masked=10, lengths="dense", zipf_s=0.8):
    """Returns dict(ids (B, S) int32 chained `[CLS][SEP] items.. [SEP]`, labels (B, Mmax) float32
    padded with -1, n_masked int, items (B, L) int32 un-chained)."""
    if lengths == "dense":
        lens = np.full(batch, max_len, dtype=np.int64)
    else:
        support = np.arange(5, 51)
        lens = rng.choice(support, size=batch, p=BEAUTY_LEN_HIST / BEAUTY_LEN_HIST.sum())
        lens = np.minimum(np.round(lens * (max_len / 50.0)).astype(np.int64), max_len)
    if mode == "train":
        lens = lens - 1  # the last item is held out for validation (input_pipeline.py:101-104)
    L = int(lens.max())
    items = np.zeros((batch, L), dtype=np.int32)
    raw = zipf_items(rng, (batch, L), vocab, zipf_s)
    valid = np.arange(L)[None, :] < lens[:, None]
    items[valid] = raw[valid]
    labels_list = []
    for b in range(batch):
        n = int(lens[b])
        if mode == "train":
            k = n_masked_for(n, masked_percentage, max_masked)
            pos = np.sort(rng.permutation(n)[:k])
        else:
            pos = np.array([n - 1])
        labels_list.append(items[b, pos] - NUM_RESERVED_TOKENS)
        items[b, pos] = MASK_ID
    mmax = max(1, max(len(l) for l in labels_list))
    labels = np.full((batch, mmax), LABEL_PAD, dtype=np.float32)
    for b, l in enumerate(labels_list):
        labels[b, :len(l)] = l
    ids = np.concatenate([np.full((batch, 1), CLS, np.int32), np.full((batch, 1), SEP, np.int32),
                          items, np.full((batch, 1), SEP, np.int32)], axis=1)
    return dict(ids=ids, labels=labels, n_masked=int(sum(len(l) for l in labels_list)), items=items)
