"""Cloze adaptor, loss and ranking metrics with the reference's names
(examples/BERT4Rec/source/utils.py:56-259)."""
import numpy as np
import torch

from . import ops
from .constants import LABEL_PAD
from .head import ClozeOutput
from .losses import MaskedLoss, _dev_f32
from .ops import F32, I32


def cloze_output_adaptor(y_true, y_pred):
    """Flatten (B, Mmax[, V]) and drop rows whose label is LABEL_PAD (utils.py:56-113).
    Returns device tensors (n, 1) labels and (n, V) predictions."""
    yt = _dev_f32(y_true).view(-1)
    yp = _dev_f32(y_pred)
    V = yp.shape[-1]
    yp = yp.view(-1, V)
    M = yt.numel()
    # positions of the valid rows via the selection kernel (labels != pad), in row-major order
    keep, count = ops.compact_labels(torch.where(yt == LABEL_PAD, yt, torch.arange(
        M, dtype=F32, device="cuda")), M, LABEL_PAD)
    n = int(count.item())
    rows = keep[:n]
    out_p = torch.empty((n, V), dtype=F32, device="cuda")
    ops.gather_rows(yp, rows, out_p, None)
    out_t = torch.empty((n, 1), dtype=F32, device="cuda")
    ops.gather_rows(yt.view(-1, 1), rows, out_t, None)
    return out_t, out_p


def _compact(y_true, y_pred):
    """(labels int32 [M] with -1 pads, score rows fp32 [M, V], V) for either a lazy ClozeOutput
    or a materialised (B, Mmax, V) prediction tensor."""
    if isinstance(y_pred, ClozeOutput):
        yt = _dev_f32(y_true)
        labels, _ = ops.compact_labels(yt.view(-1), y_pred.M)
        return labels, None, y_pred.head.output_vocab_size
    yt, yp = cloze_output_adaptor(y_true, y_pred)
    return yt.view(-1).to(I32), yp, yp.shape[-1]


class ClozeMaskedLoss:
    """ClozeMaskedLoss(item_wise_loss_fn, label_pad=LABEL_PAD) — utils.py:116-134."""

    def __init__(self, item_wise_loss_fn, label_pad=LABEL_PAD):
        self.masked_loss = MaskedLoss(item_wise_loss_fn=item_wise_loss_fn, label_pad=label_pad)

    def call(self, y_true, y_pred):
        if isinstance(y_pred, ClozeOutput):
            labels, _, _ = _compact(y_true, y_pred)
            stats = torch.empty(2, dtype=F32, device="cuda")
            y_pred.head.vocab.loss_forward(y_pred.ab, y_pred.M, labels, stats, need_grad=False)
            s = stats.cpu().numpy()
            return float(s[0] / s[1]) if s[1] > 0 else 0.0
        yt, yp = cloze_output_adaptor(y_true, y_pred)
        return self.masked_loss(yt, yp)

    __call__ = call


class _ClozeRankMetric:
    MAX_K = 256   # b4cp_topk_rows / b4cp_rank_metrics limit (include/b4cp.h)

    def __init__(self, k, name):
        self.k, self.name = int(k), name
        if not 1 <= self.k <= self.MAX_K:
            raise ValueError(f"k={k}: ranking metrics support 1 <= k <= {self.MAX_K}")
        self.counters = None

    def _counters(self):
        if self.counters is None:
            self.counters = torch.zeros(3, dtype=F32, device="cuda")
        return self.counters

    def get_config(self):
        return {'k': self.k, 'name': self.name}

    def update_state(self, y_true, y_pred, sample_weight=None):
        labels, scores, V = _compact(y_true, y_pred)
        if isinstance(y_pred, ClozeOutput):
            ids = y_pred.head.vocab.topk(y_pred.ab, y_pred.M, self.k)
        else:
            if scores.shape[0] == 0:
                return
            ids, _ = ops.topk_rows(scores, V, self.k)
        ops.rank_metrics(ids, ids.shape[1], labels, self._counters())

    def reset_states(self):
        if self.counters is not None:
            self.counters.zero_()


class ClozeMaskedRecall(_ClozeRankMetric):
    """ClozeMaskedRecall(k, name=None): hits@k / n (utils.py:137-194)."""

    def __init__(self, k, name=None):
        super().__init__(k, name if name is not None else f'Recall_at_{k}')

    def result(self):
        c = self._counters().cpu().numpy()
        return c[0] / c[2]


class ClozeMaskedNDCG(_ClozeRankMetric):
    """ClozeMaskedNDCG(k, name=None): sum of 1/log2(rank+2) at the hit / n (utils.py:197-259)."""

    def __init__(self, k, name=None):
        super().__init__(k, name if name is not None else f'NDCG_at_{k}')

    def result(self):
        c = self._counters().cpu().numpy()
        return c[1] / c[2]
