"""Binary-task metrics with the reference's names and semantics
(clickstream_transformer/metrics.py:5-106): PositiveRate, PredictedPositives, F1Score and the
MaskedMetric wrapper.  Counters live on the device and are accumulated by one kernel
(`b4cp_binary_metric_counts`); `result()` reads them back."""
import numpy as np
import torch

from . import ops
from .constants import LABEL_PAD
from .ops import F32


def _dev(t):
    if not torch.is_tensor(t):
        t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float32))
    return t.to(device="cuda", dtype=F32).contiguous().view(-1)


class _CountMetric:
    def __init__(self, name):
        self.name = name
        self.counters = None

    def _c(self):
        if self.counters is None:
            self.counters = torch.zeros(6, dtype=F32, device="cuda")
        return self.counters

    def update_state(self, y_true, y_pred, sample_weight=None):
        yt, yp = _dev(y_true), _dev(y_pred)
        assert yt.numel() == yp.numel()
        ops.binary_metric_counts(yt, yp, LABEL_PAD, self._c())

    def reset_states(self):
        if self.counters is not None:
            self.counters.zero_()

    def _host(self):
        return self._c().cpu().numpy().astype(np.float64)


class PositiveRate(_CountMetric):
    """sum(y_true over unpadded items) / number of unpadded items (metrics.py:5-27)."""

    def __init__(self, name='positive_rate', **kwargs):
        super().__init__(name)

    def result(self):
        c = self._host()
        return np.float32(c[1]) / np.float32(c[0])


class PredictedPositives(_CountMetric):
    """sum(round(y_pred) over unpadded items) / number of unpadded items (metrics.py:30-53)."""

    def __init__(self, name='pred_positives', **kwargs):
        super().__init__(name)

    def result(self):
        c = self._host()
        return np.float32(c[2]) / np.float32(c[0])


class F1Score(_CountMetric):
    """2 tp / (condition_true + predicted_true) with the reference's unmasked int32 comparisons
    (metrics.py:56-86).  Like the reference, `sample_weight` is ignored, so MaskedMetric(F1Score())
    counts padded positions whose rounded prediction is 1 as predicted-true."""

    def __init__(self, name='F1Score', **kwargs):
        super().__init__(name)

    def result(self):
        c = self._host()
        return np.float32(2.0) * np.float32(c[3]) / (np.float32(c[4]) + np.float32(c[5]))


class MaskedMetric:
    """MaskedMetric(metric, name): passes mask = (y_true != LABEL_PAD) as sample_weight to the
    wrapped metric (metrics.py:89-106); sample_weight itself is rejected with ValueError."""

    def __init__(self, metric, name, **kwargs):
        self._metric, self.name = metric, name

    def update_state(self, y_true, y_pred, sample_weight=None):
        if sample_weight is not None:
            raise ValueError("Masked metrics do not support sample_weight.")
        yt = _dev(y_true)
        self._metric.update_state(yt, y_pred, sample_weight=(yt != LABEL_PAD))

    def result(self):
        return self._metric.result()

    def reset_states(self):
        self._metric.reset_states()
