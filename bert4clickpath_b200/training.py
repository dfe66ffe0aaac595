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
    def __init__(self, model, optimizer=None):
        self.model = model
        self.opt = optimizer or Adam()
        self.seed = 0
        self._dev_ids = self._dev_labels = None
        self._host_stats = torch.empty(2, dtype=F32).pin_memory() if torch.cuda.is_available() else None

    def to_device(self, batch):
        """batch: dict from synthetic.make_cloze_batch (host NumPy)."""
        ids = [torch.from_numpy(np.ascontiguousarray(batch["ids"])).cuda().view(-1)]
        labels = torch.from_numpy(np.ascontiguousarray(batch["labels"])).cuda()
        B, S = batch["ids"].shape
        return DeviceBatch(ids, labels, B, S, batch["n_masked"])

    def step_device(self, db):
        """forward + backward + all-reduce + Adam; returns the device loss statistics."""
        self.seed += 1
        stats = self.model.cloze_forward_backward(db.ids, db.labels, db.B, db.S,
                                                  n_masked=db.n_masked, training=True,
                                                  seed=self.seed)
        self.model.store.adam(self.opt.learning_rate, self.opt.beta_1, self.opt.beta_2,
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
        self._dev_ids.copy_(ids_pinned, non_blocking=True)
        self._dev_labels.copy_(labels_pinned, non_blocking=True)
        db = DeviceBatch([self._dev_ids.view(-1)], self._dev_labels, B, S, n_masked)
        stats = self.step_device(db)
        self._host_stats.copy_(stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        s0, s1 = float(self._host_stats[0]), float(self._host_stats[1])
        return s0 / s1 if s1 > 0 else 0.0

    @staticmethod
    def h2d_bytes(ids_pinned, labels_pinned):
        return ids_pinned.numel() * 4 + labels_pinned.numel() * 4
