"""Cloze training step driver: the call a user of the reference's `model.fit` loop makes per
batch (examples/BERT4Rec/source/main.py:159-165), as one fixed launch sequence on one stream.

`step_device` runs on batches already resident in HBM; `step_host` is the end-to-end call:
pinned host ids / labels -> H2D -> forward, backward, gradient all-reduce, Adam -> D2H of the
loss statistics.
"""
import numpy as np
import torch

from .clickstream_transformer import Adam
from .ops import F32, I32


class DeviceBatch:
    __slots__ = ("ids", "labels", "B", "S", "n_masked")

    def __init__(self, ids, labels, B, S, n_masked):
        self.ids, self.labels, self.B, self.S, self.n_masked = ids, labels, B, S, n_masked


class ClozeTrainStep:
    """use_graph=True captures the whole step (forward, backward, NCCL gradient all-reduce, Adam)
    into one CUDA graph per (B, S, masked-row CAPACITY) after two eager steps, and replays it: the
    ~140 kernel launches of a step cost one host call.  Dropout seeds are then read on the device
    from the Adam step counter (B4CP_SEED_FROM_DEVICE), so every replay draws new masks.

    Real Cloze batches have a different total mask count almost every step, so graphs are keyed on
    a row capacity - the count rounded up to ROW_BUCKET rows - not on the count itself: the
    selection / label-compaction kernels pad the rows past the true count with -1 and every later
    kernel treats those rows as dead.  At most MAX_GRAPHS captured graphs are kept (least recently
    used is dropped).  With a vocabulary-parallel head every rank must present the same capacity:
    pass `row_capacity` (fixed for the run) or let each step negotiate it (one small all-reduce
    and host read per step, outside the captured region)."""

    GRAPH_WARMUP_STEPS = 2
    ROW_BUCKET = 128
    MAX_GRAPHS = 8

    def __init__(self, model, optimizer=None, use_graph=False, row_capacity=None):
        self.model = model
        self.opt = optimizer or Adam()
        self.seed = 0
        self.iterations = 0
        self.use_graph = bool(use_graph)
        self.row_capacity = row_capacity
        self._graphs = {}
        self._dev_ids = self._dev_labels = None
        self._host_stats = torch.empty(2, dtype=F32).pin_memory() if torch.cuda.is_available() else None
        self._copy_stream = None
        self._slots = {}
        self._staging = {}

    def _lr(self):
        lr = self.opt.learning_rate
        if callable(lr):        # training_utils.CustomLRSchedule / CustomExponentialDecayLR
            lr = lr(self.iterations)
        self.iterations += 1
        return float(lr)

    def to_device(self, batch):
        """batch: dict from synthetic.make_cloze_batch (host NumPy)."""
        ids = [torch.from_numpy(np.ascontiguousarray(batch["ids"])).cuda().view(-1)]
        labels = torch.from_numpy(np.ascontiguousarray(batch["labels"])).cuda()
        B, S = batch["ids"].shape
        return DeviceBatch(ids, labels, B, S, batch["n_masked"])

    def capacity(self, n_masked):
        """Masked-row capacity the step runs at (>= n_masked)."""
        if self.row_capacity is not None:
            if n_masked > self.row_capacity:
                raise ValueError(f"batch has {n_masked} [MASK] rows > row_capacity {self.row_capacity}")
            return int(self.row_capacity)
        cap = max(self.ROW_BUCKET, -(-int(n_masked) // self.ROW_BUCKET) * self.ROW_BUCKET)
        vocab = getattr(self.model.head, "vocab", None)
        if hasattr(vocab, "common_rows"):
            cap = vocab.common_rows(cap)      # host read; never inside a capture
        return cap

    def _eager(self, db, seed, cap=None):
        cap = self.capacity(db.n_masked) if cap is None else cap
        stats = self.model.cloze_forward_backward(db.ids, db.labels, db.B, db.S,
                                                  n_masked=cap, training=True, seed=seed,
                                                  rows_are_common=True)
        self.model.store.adam(self._lr(), self.opt.beta_1, self.opt.beta_2,
                              self.opt.epsilon)
        return stats

    def _step_graph(self, db):
        from . import ops
        cap = self.capacity(db.n_masked)
        db = DeviceBatch(db.ids, db.labels, db.B, db.S, cap)
        key = (db.B, db.S, cap, len(db.ids))
        st = self._graphs.pop(key, None)
        if st is None:
            st = dict(calls=0, graph=None)
            while len(self._graphs) >= self.MAX_GRAPHS:   # LRU: dicts keep insertion order
                self._graphs.pop(next(iter(self._graphs)))
        self._graphs[key] = st                            # (re)insert as most recently used
        seed = ops.device_seed(self.model.store.step_dev)
        if callable(self.opt.learning_rate):
            raise TypeError("a learning-rate schedule changes Adam's scalar every step: "
                            "use use_graph=False")
        hp = (float(self.opt.learning_rate), self.opt.beta_1, self.opt.beta_2, self.opt.epsilon)
        if st["graph"] is not None and st["hp"] != hp:
            st["graph"] = None      # Adam's scalars are baked into the captured launches
                                    # (ReduceLROnPlateau changes lr between epochs): re-capture
        if st["graph"] is None:
            st["calls"] += 1
            if st["calls"] <= self.GRAPH_WARMUP_STEPS:   # real steps; they also size every buffer
                return self._eager(db, seed, cap)
            if "ids" not in st:     # kept across re-captures: step_host copies into them
                st["ids"] = [torch.empty_like(t) for t in db.ids]
                st["labels"] = torch.empty_like(db.labels)
            sdb = DeviceBatch(st["ids"], st["labels"], db.B, db.S, db.n_masked)
            timer_was, ops.TIMER.enabled = ops.TIMER.enabled, False
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st["stats"] = self._eager(sdb, seed, cap)
            ops.TIMER.enabled = timer_was
            st["graph"], st["hp"] = g, hp
        for dst, src in zip(st["ids"], db.ids):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if st["labels"].data_ptr() != db.labels.data_ptr():
            st["labels"].copy_(db.labels, non_blocking=True)
        st["graph"].replay()
        return st["stats"]

    def step_device(self, db):
        """forward + backward + all-reduce + Adam; returns the device loss statistics."""
        if self.use_graph:
            return self._step_graph(db)
        self.seed += 1
        stats = self.model.cloze_forward_backward(db.ids, db.labels, db.B, db.S,
                                                  n_masked=self.capacity(db.n_masked),
                                                  training=True, seed=self.seed,
                                                  rows_are_common=True)
        self.model.store.adam(self._lr(), self.opt.beta_1, self.opt.beta_2,
                              self.opt.epsilon)
        return stats

    def step_host(self, ids_pinned, labels_pinned, n_masked):
        """ids_pinned: int32 (B, S) pinned host tensor; labels_pinned: float32 (B, Mmax) pinned.
        Returns the global mean loss as a Python float (forces the D2H read)."""
        B, S = ids_pinned.shape
        if self._dev_ids is None or self._dev_ids.shape != ids_pinned.shape:
            self._dev_ids = torch.empty(ids_pinned.shape, dtype=I32, device="cuda")
        if self._dev_labels is None or self._dev_labels.shape != labels_pinned.shape:
            self._dev_labels = torch.empty(labels_pinned.shape, dtype=F32, device="cuda")
        dev_ids, dev_labels = self._dev_ids.view(-1), self._dev_labels
        if self.use_graph:  # copy straight into the captured graph's input buffers
            st = self._graphs.get((B, S, self.capacity(n_masked), 1))
            if st is not None and st.get("graph") is not None:
                dev_ids, dev_labels = st["ids"][0], st["labels"]
        dev_ids.view(B, S).copy_(ids_pinned, non_blocking=True)
        dev_labels.copy_(labels_pinned, non_blocking=True)
        db = DeviceBatch([dev_ids], dev_labels, B, S, n_masked)
        stats = self.step_device(db)
        self._host_stats.copy_(stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        s0, s1 = float(self._host_stats[0]), float(self._host_stats[1])
        return s0 / s1 if s1 > 0 else 0.0

    def encode_host(self, features, labels, parity=0):
        """A batch as the reference's dataset yields it - {raw feature: (B, L) string array},
        (B, max_n_masked) float labels padded with -1 (input_pipeline.py:198-214) - to what
        `step_host` / `run_host` take: (ids_pinned int32 (B, S), labels_pinned, n_masked).  The
        chaining and the vocabulary lookup (native table, host threads) write straight into pinned
        staging memory; two staging sets alternate by `parity`, which is what `run_host` needs
        (it copies batch k while the caller encodes batch k + 1)."""
        m = self.model
        name = list(m.sequential_input_config.keys())[0]
        assert len(m.sequential_input_config) == 1, "the Cloze step takes one sequential feature"
        first_raw = np.asarray(features[m.sequential_input_config[name][0]])
        B = first_raw.shape[0]
        S = 1 + sum(np.asarray(features[r]).shape[1] + 1 for r in m.sequential_input_config[name]) + 1
        labels = np.asarray(labels, dtype=np.float32)
        key = (parity & 1, B, S, labels.shape)
        st = self._staging.get(key)
        if st is None:
            st = (torch.empty((B, S), dtype=I32).pin_memory(), torch.empty(labels.shape, dtype=F32).pin_memory())
            self._staging[key] = st
        ids_pinned, labels_pinned = st
        m.prepare_host(features, out={name: ids_pinned.numpy()})
        labels_pinned.numpy()[...] = labels
        return ids_pinned, labels_pinned, int((labels != -1.0).sum())

    def run_host(self, batches):
        """The end-to-end loop a `model.fit(dataset.prefetch(...))` of the reference runs
        (examples/BERT4Rec/source/main.py:140-165): for every (ids_pinned, labels_pinned,
        n_masked) of `batches` the H2D copy of its inputs from pinned memory, the training step,
        and the D2H read of its loss statistics.  Yields the global mean loss of every step, ONE
        STEP BEHIND the launches: step k+1 is copied (on a copy stream, into the other of two
        landing buffers) and enqueued while step k runs, so neither the copy nor the host's launch
        work leaves the device idle.  `step_host` is the same thing strictly one step at a time."""
        cur = torch.cuda.current_stream()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        cs = self._copy_stream
        pending = None

        def read(slot):
            slot["done"].synchronize()
            s0, s1 = float(slot["stats"][0]), float(slot["stats"][1])
            return s0 / s1 if s1 > 0 else 0.0

        for k, (ids_pinned, labels_pinned, n_masked) in enumerate(batches):
            key = (k & 1, tuple(ids_pinned.shape), tuple(labels_pinned.shape))
            slot = self._slots.get(key)
            if slot is None:
                slot = dict(ids=torch.empty(ids_pinned.shape, dtype=I32, device="cuda"),
                            labels=torch.empty(labels_pinned.shape, dtype=F32, device="cuda"),
                            stats=torch.empty(2, dtype=F32).pin_memory(),
                            copied=torch.cuda.Event(), consumed=None, done=torch.cuda.Event())
                self._slots[key] = slot
            B, S = ids_pinned.shape
            with torch.cuda.stream(cs):
                if slot["consumed"] is not None:     # the step that last read this landing buffer
                    cs.wait_event(slot["consumed"])
                slot["ids"].copy_(ids_pinned, non_blocking=True)
                slot["labels"].copy_(labels_pinned, non_blocking=True)
                slot["copied"].record(cs)
            cur.wait_event(slot["copied"])
            stats = self.step_device(DeviceBatch([slot["ids"].view(-1)], slot["labels"], B, S, n_masked))
            if slot["consumed"] is None:
                slot["consumed"] = torch.cuda.Event()
            slot["consumed"].record(cur)
            slot["stats"].copy_(stats, non_blocking=True)
            slot["done"].record(cur)
            if pending is not None:
                yield read(pending)
            pending = slot
        if pending is not None:
            yield read(pending)

    @staticmethod
    def h2d_bytes(ids_pinned, labels_pinned):
        return ids_pinned.numel() * 4 + labels_pinned.numel() * 4
