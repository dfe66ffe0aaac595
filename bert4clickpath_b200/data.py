"""Real-data input side of the Cloze path (host NumPy; the step immediately before the hot path,
SURVEY.md N2): the BERT4Rec text format reader and session prep of
examples/BERT4Rec/data_prep/main.py:45-76, and the Cloze example / batch builder of
examples/BERT4Rec/source/input_pipeline.py:21-32, :59-133, :198-220, producing the same dict as
`synthetic.make_cloze_batch` (already-chained int32 ids + the reference's padded float32 labels),
i.e. what `ClozeTrainStep.to_device` / `ClickstreamTransformer.cloze_forward_backward` consume.

Not on the timed path and not a tf.data replacement: no TFRecord I/O, no shuffling buffers.
"""
import numpy as np

from .constants import (CLS, INPUT_MASKING_TOKEN, INPUT_PADDING_TOKEN, LABEL_PAD, MASK_ID,
                        NUM_RESERVED_TOKENS, SEP)
from .synthetic import n_masked_for

MASKED_PERCENTAGE = 0.4      # cloze_constants.py:1
MAX_MASKED_ITEMS = 10        # cloze_constants.py:2
MAX_SEQ_LEN = 50             # data_prep/main.py:57


def read_bert4rec_text_data(path):
    """`user item` pairs, one per line, in interaction order (data_prep/main.py:45-49).
    Returns (users, items): two lists of strings of equal length."""
    users, items = [], []
    with open(path) as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 2:
                users.append(parts[0])
                items.append(parts[1])
    return users, items


def prepare_sessions(users, items, max_seq_len=MAX_SEQ_LEN):
    """data_prep/main.py:62-76: keep each user's first `max_seq_len` interactions (cumcount <
    max), item vocabulary = pd.unique over the kept rows (order of first appearance), one session
    per user in order of first appearance.  Returns (sessions: list of lists of item strings,
    item_vocab: list of strings, user_ids: list of strings)."""
    per_user, order = {}, []
    vocab, seen = [], set()
    for u, it in zip(users, items):
        s = per_user.get(u)
        if s is None:
            s = per_user[u] = []
            order.append(u)
        if len(s) >= max_seq_len:
            continue
        s.append(it)
        if it not in seen:
            seen.add(it)
            vocab.append(it)
    return [per_user[u] for u in order], vocab, order


def cloze_example(item_ids, mode, rng, masked_percentage=MASKED_PERCENTAGE, max_masked=MAX_MASKED_ITEMS):
    """One session of INPUT-vocabulary ids (>= 10) -> (masked ids, labels in the label vocabulary
    = id - 10, as the reference's separate target lookup table gives: input_pipeline.py:189-192).
    TRAIN: the last item is held out (:101-104), n = clip(int(len * p), 0, max) (:68-70) distinct
    positions are replaced by [MASK] and the labels follow in ascending position (:21-32, :77-90).
    EVAL: only the last item is masked (:118-121)."""
    ids = np.asarray(item_ids, dtype=np.int32)
    if mode == "train":
        ids = ids[:-1].copy()
        n = n_masked_for(len(ids), masked_percentage, max_masked)
        pos = np.sort(rng.permutation(len(ids))[:n])
    elif mode == "eval":
        ids = ids.copy()
        pos = np.array([len(ids) - 1])
    else:
        raise ValueError(f"Unrecognized mode: {mode}")
    labels = (ids[pos] - NUM_RESERVED_TOKENS).astype(np.float32)
    ids[pos] = MASK_ID
    return ids, labels


def cloze_batch(sessions_ids, mode, rng, masked_percentage=MASKED_PERCENTAGE, max_masked=MAX_MASKED_ITEMS):
    """Batch of sessions (lists of input-vocabulary ids) -> dict(ids (B, S) int32 chained
    `[CLS] [SEP] items... [PAD]... [SEP]` - every sequence is padded BEFORE chaining, so the trailing
    [SEP] sits after the pad run (input_pipeline.py:198-214, clickstream_transformer.py:54-61,
    SURVEY.md T7) -, labels (B, Mmax) float32 padded with -1, n_masked, items (B, L))."""
    ex = [cloze_example(s, mode, rng, masked_percentage, max_masked) for s in sessions_ids]
    B = len(ex)
    L = max(len(e[0]) for e in ex)
    items = np.zeros((B, L), dtype=np.int32)
    mmax = max(1, max(len(e[1]) for e in ex))
    labels = np.full((B, mmax), LABEL_PAD, dtype=np.float32)
    for b, (ids, lab) in enumerate(ex):
        items[b, :len(ids)] = ids
        labels[b, :len(lab)] = lab
    ids = np.concatenate([np.full((B, 1), CLS, np.int32), np.full((B, 1), SEP, np.int32), items,
                          np.full((B, 1), SEP, np.int32)], axis=1)
    return dict(ids=ids, labels=labels, n_masked=int(sum(len(e[1]) for e in ex)), items=items)


class ClozeDataset:
    """Sessions of the BERT4Rec text format as Cloze batches.  `vocab` defaults to the file's own
    item vocabulary (order of first appearance, as data_prep writes item_vocab.txt); input ids are
    index + 10 (clickstream_transformer.py:247-258), labels index (no offset)."""

    def __init__(self, path, max_seq_len=MAX_SEQ_LEN, vocab=None):
        users, items = read_bert4rec_text_data(path)
        sessions, file_vocab, users = prepare_sessions(users, items, max_seq_len)
        self._bind(sessions, users, vocab if vocab is not None else file_vocab)

    @classmethod
    def from_sessions(cls, sessions, users=None, vocab=None):
        """Already-prepared sessions (lists of item strings).  `vocab`: list of item strings or
        the path of an `item_vocab.txt` (one token per line, data_prep/main.py:77-80); default =
        order of first appearance, which is what data_prep writes."""
        self = cls.__new__(cls)
        if isinstance(vocab, str):
            with open(vocab) as f:
                vocab = [line.strip() for line in f if line.strip()]
        if vocab is None:
            vocab, seen = [], set()
            for s in sessions:
                for t in s:
                    if t not in seen:
                        seen.add(t)
                        vocab.append(t)
        self._bind([list(s) for s in sessions],
                   list(users) if users is not None else [str(i) for i in range(len(sessions))], vocab)
        return self

    @classmethod
    def from_tfrecord(cls, paths, vocab=None, verify_crc=True):
        """The TFRecord files the reference's data prep writes (one Example per user with
        `reviewerID` and the `asin` list; data_prep/main.py:86-98, input_pipeline.py:149-156),
        read without TensorFlow (tfrecord.py)."""
        from .tfrecord import read_sessions
        users, sessions = read_sessions(paths, verify_crc=verify_crc)
        return cls.from_sessions(sessions, users, vocab)

    def _bind(self, sessions, users, vocab):
        self.sessions, self.users = sessions, users
        self.vocab = list(vocab)
        index = {t: i for i, t in enumerate(self.vocab)}
        oov = len(self.vocab) + NUM_RESERVED_TOKENS   # StaticVocabularyTable's single OOV bucket
        self.session_ids = [np.array([index[t] + NUM_RESERVED_TOKENS if t in index else oov for t in s],
                                     dtype=np.int32) for s in self.sessions]

    def __len__(self):
        return len(self.session_ids)

    def device_builder(self):
        """The sessions uploaded once as a CSR for on-device batch building (DeviceClozeBuilder)."""
        return DeviceClozeBuilder(self.session_ids)

    def batches(self, batch_size, mode, rng, masked_percentage=MASKED_PERCENTAGE,
                max_masked=MAX_MASKED_ITEMS, shuffle=None, drop_remainder=False):
        """Yields batch dicts; TRAIN shuffles the session order by default."""
        order = np.arange(len(self))
        if shuffle if shuffle is not None else mode == "train":
            order = rng.permutation(len(self))
        for a in range(0, len(order), batch_size):
            idx = order[a:a + batch_size]
            if drop_remainder and len(idx) < batch_size:
                return
            yield cloze_batch([self.session_ids[i] for i in idx], mode, rng, masked_percentage, max_masked)


def _shuffled_epochs(load, rng, buffer_size):
    """dataset.shuffle(buffer_size, reshuffle_each_iteration=True).repeat(None)
    (input_pipeline.py:183-185): a sliding shuffle buffer per pass over the source, forever."""
    while True:
        buf, seen = [], False
        for ex in load():
            seen = True
            if len(buf) < buffer_size:
                buf.append(ex)
                continue
            j = int(rng.integers(len(buf)))
            out, buf[j] = buf[j], ex
            yield out
        if not seen:
            raise ValueError("create_cloze_dataset: the source yielded no examples")
        while buf:
            j = int(rng.integers(len(buf)))
            buf[j], buf[-1] = buf[-1], buf[j]
            yield buf.pop()


def create_cloze_dataset(source, mode, batch_size, target_vocab_file, rng=None,
                         masked_percentage=MASKED_PERCENTAGE, max_masked=MAX_MASKED_ITEMS,
                         shuffle_buffer=20000):
    """examples/BERT4Rec/source/input_pipeline.py:136-232 with the same arguments: an endless
    iterator of `(features, labels)` batches in the reference's own contract -
    features = {'reviewerID': (B,) str, 'asin': (B, L) str padded with '[PAD]', masked positions
    replaced by '[MASK]'}, labels = (B, M) float32 target-vocabulary indices (one OOV bucket =
    len(vocab), :189-192) padded with -1 - i.e. what `ClickstreamTransformer.fit / train_step`
    takes.  `source`: a glob pattern of TFRecord files (read without TensorFlow, tfrecord.py) or a
    callable returning an iterator of {'reviewerID': str, 'asin': list of str}.  Shuffle buffer,
    repeat, per-example masking (cloze_example's rules), batches padded to their longest row."""
    import glob
    if mode not in ("train", "eval"):
        raise ValueError(f"Unrecognized mode: {mode}")
    rng = rng if rng is not None else np.random.default_rng()
    if isinstance(source, str):
        files = sorted(glob.glob(source))
        if not files:
            raise FileNotFoundError(source)

        def load():
            from .tfrecord import decode_example, read_records
            for f in files:
                for rec in read_records(f):
                    ex = decode_example(rec)
                    yield {"reviewerID": ex["reviewerID"][0].decode("utf-8"),
                           "asin": [v.decode("utf-8") for v in ex.get("asin", [])]}
    elif callable(source):
        load = source
    else:
        raise TypeError("Source must be either str or callable.")
    with open(target_vocab_file) as f:
        label_vocab = [line.strip() for line in f if line.strip()]
    label_index = {t: i for i, t in enumerate(label_vocab)}
    oov = len(label_vocab)

    def examples():
        for ex in _shuffled_epochs(load, rng, shuffle_buffer):
            items = list(ex["asin"])
            if mode == "train":
                items = items[:-1]
                n = n_masked_for(len(items), masked_percentage, max_masked)
                pos = np.sort(rng.permutation(len(items))[:n])
            else:
                pos = np.array([len(items) - 1])
            labels = [float(label_index.get(items[p], oov)) for p in pos]
            for p in pos:
                items[p] = INPUT_MASKING_TOKEN
            yield ex["reviewerID"], items, labels

    it = examples()
    while True:
        rows = [next(it) for _ in range(batch_size)]
        L = max(len(r[1]) for r in rows)
        M = max(len(r[2]) for r in rows)
        asin = np.full((batch_size, L), INPUT_PADDING_TOKEN, dtype=object)
        labels = np.full((batch_size, M), LABEL_PAD, dtype=np.float32)
        for b, (_, items, lab) in enumerate(rows):
            asin[b, :len(items)] = items
            labels[b, :len(lab)] = lab
        yield {"reviewerID": np.array([r[0] for r in rows], dtype=object), "asin": asin}, labels


class DeviceClozeBuilder:
    """Sessions resident in HBM (one CSR of input-vocabulary ids) -> Cloze batches built by
    `b4cp_cloze_build` on the device: masking, label extraction, padding and chaining
    (input_pipeline.py:21-32, :59-133, :198-214; clickstream_transformer.py:38-63) without host
    work per step.  The host only chooses the session indices; because the number of masked items
    is a function of the session length alone, batch shapes and `n_masked` are known on the host
    without a device round trip.  Mask positions follow the keyed rule stated in include/b4cp.h
    (the reference's shuffle stream cannot be matched), so a batch is a pure function of
    (seed, session indices, mode)."""

    def __init__(self, session_ids):
        import torch
        self.lengths = np.array([len(s) for s in session_ids], dtype=np.int64)
        offsets = np.zeros(len(session_ids) + 1, dtype=np.int64)
        np.cumsum(self.lengths, out=offsets[1:])
        flat = (np.concatenate([np.asarray(s, dtype=np.int32) for s in session_ids])
                if len(session_ids) else np.zeros(0, np.int32))
        self.items = torch.from_numpy(flat).cuda()
        self.offsets = torch.from_numpy(offsets).cuda()
        self._status = torch.zeros(1, dtype=torch.int32, device="cuda")

    def __len__(self):
        return len(self.lengths)

    def shapes(self, session_idx, mode, masked_percentage=MASKED_PERCENTAGE,
               max_masked=MAX_MASKED_ITEMS):
        """(L, Mmax, n_masked) of the batch `cloze_batch` would build from these sessions."""
        lens = self.lengths[np.asarray(session_idx, dtype=np.int64)]
        if mode == "train":
            lens = lens - 1
            counts = [n_masked_for(int(n), masked_percentage, max_masked) for n in lens]
        elif mode == "eval":
            counts = [1] * len(lens)
        else:
            raise ValueError(f"Unrecognized mode: {mode}")
        if len(lens) and lens.min() < (0 if mode == "train" else 1):
            raise ValueError("empty session in the batch")
        return int(max(1, lens.max())), max(1, max(counts)), int(sum(counts))

    def build(self, session_idx, mode, seed, masked_percentage=MASKED_PERCENTAGE,
              max_masked=MAX_MASKED_ITEMS, L=None, Mmax=None, check=False):
        """Returns dict(ids int32 (B, L + 3) device, labels float32 (B, Mmax) device, n_masked
        int, B, S).  L / Mmax default to the batch's own maxima (what the host builder produces);
        pass fixed values for shape-stable CUDA-graph steps.  check=True reads the status word
        back (a device sync) and raises if a row did not fit."""
        import torch
        from . import ops
        idx = np.ascontiguousarray(session_idx, dtype=np.int32)
        B = len(idx)
        l_need, m_need, n_masked = self.shapes(idx, mode, masked_percentage, max_masked)
        L = l_need if L is None else int(L)
        Mmax = m_need if Mmax is None else int(Mmax)
        if L < l_need or Mmax < m_need:
            raise ValueError(f"batch needs L >= {l_need}, Mmax >= {m_need}; got {L}, {Mmax}")
        idx_dev = torch.from_numpy(idx).cuda()
        ids = torch.empty((B, L + 3), dtype=torch.int32, device="cuda")
        labels = torch.empty((B, Mmax), dtype=torch.float32, device="cuda")
        count = torch.zeros(1, dtype=torch.int32, device="cuda")
        ops.cloze_build(self.items, self.offsets, idx_dev, B, L, Mmax, mode == "train",
                        masked_percentage, max_masked, seed,
                        (CLS, SEP, MASK_ID, 0, NUM_RESERVED_TOKENS), LABEL_PAD, ids, labels, count,
                        self._status)
        if check:
            bad = int(self._status.item())
            if bad or int(count.item()) != n_masked:
                raise RuntimeError(f"cloze_build: row {bad - 1} did not fit / count mismatch")
        return dict(ids=ids, labels=labels, n_masked=n_masked, B=B, S=L + 3)
