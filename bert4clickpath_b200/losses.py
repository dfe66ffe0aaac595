"""MaskedLoss with the reference's constructor (clickstream_transformer/losses.py:5-98).

`item_wise_loss_fn` is one of the two un-reduced Keras backend losses the reference documents
(losses.py:12-15, examples/BERT4Rec/source/main.py:89); they are exported here as named
callables so that call sites read like the reference's.  The arithmetic runs in libb4cp kernels.
"""
import numpy as np
import torch

from . import ops
from .constants import LABEL_PAD
from .ops import F32, I32


class _ItemWiseLoss:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"<item-wise loss {self.name}>"


sparse_categorical_crossentropy = _ItemWiseLoss("sparse_categorical_crossentropy")
binary_crossentropy = _ItemWiseLoss("binary_crossentropy")


def _dev_f32(t):
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float32))
    return t.to(device="cuda", dtype=F32).contiguous()


class MaskedLoss:
    """MaskedLoss(item_wise_loss_fn, pos_weight=None, label_pad=LABEL_PAD).__call__(y_true, y_pred)
    -> masked mean of the item-wise loss over labels != label_pad; 0.0 for an empty tensor
    (losses.py:89-91); divided by (pos_weight + 1)/2 when pos_weight is given (losses.py:94-96)."""

    def __init__(self, item_wise_loss_fn, pos_weight=None, label_pad=LABEL_PAD):
        self.item_wise_loss_fn = item_wise_loss_fn
        assert label_pad < 0, "label_pad must be less than zero, to distinguish it from actual labels."
        self.label_pad = float(label_pad)
        if pos_weight is not None:
            print('*' * 80)
            print('WARNING: providing pos_weight to a masked loss only works as expected for binary labels.')
            print('*' * 80)
        self.pos_weight = float(pos_weight) if pos_weight is not None else None
        self._negative_weight = 1.0

    def call(self, y_true, y_pred):
        yt = _dev_f32(y_true)
        yp = _dev_f32(y_pred)
        if yt.numel() == 0:
            return 0.0
        if self.item_wise_loss_fn is binary_crossentropy:
            assert yt.numel() == yp.numel()
            s = ops.masked_bce(yt.view(-1), yp.view(-1), self.label_pad, self.pos_weight).cpu().numpy()
        elif self.item_wise_loss_fn is sparse_categorical_crossentropy:
            V = yp.shape[-1]
            probs = yp.view(-1, V)
            M = probs.shape[0]
            if self.pos_weight is not None:
                raise NotImplementedError("pos_weight is defined for binary labels only")
            # rows keep their position: build per-row int labels (-1 where padded)
            lab_rows = torch.where(yt.view(-1) == self.label_pad, torch.full_like(yt.view(-1), -1.0),
                                   yt.view(-1)).to(I32)
            z = ops.clip_log(probs)  # TF 2.3: clip, log, then softmax-CE on the log-probabilities
            lse = torch.empty(M, dtype=F32, device="cuda")
            tgt = torch.empty(M, dtype=F32, device="cuda")
            stats = torch.empty(2, dtype=F32, device="cuda")
            ops.ce_rows_stats(z, V, lab_rows, lse, tgt)
            ops.ce_loss_reduce(lse, tgt, lab_rows, stats)
            s = stats.cpu().numpy()
        else:
            raise TypeError("item_wise_loss_fn must be losses.sparse_categorical_crossentropy or "
                            "losses.binary_crossentropy")
        mean = float(s[0]) / float(s[1]) if s[1] > 0 else float('nan')  # 0/0 like the reference
        if self.pos_weight is not None:
            mean = mean / ((self.pos_weight + self._negative_weight) / 2)
        return mean

    __call__ = call
