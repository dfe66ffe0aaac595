"""Host-side training control around the step: learning-rate schedules and the epoch-end
callbacks the reference's training script installs.

Mirrors clickstream_transformer/training_utils.py (CustomLRSchedule :15-36,
CustomExponentialDecayLR :39-60, BestModelSaverCallback :63-75) and the two stock Keras
callbacks of examples/BERT4Rec/source/main.py:134 (`ReduceLROnPlateau(monitor='val_loss',
patience=10, factor=0.317)`) and :156 (`EarlyStopping(monitor='val_loss', patience=30)`), whose
behaviour is restated from the published tf.keras 2.3 semantics (TensorFlow is not installable
here).  Everything in this file is plain Python over floats: no device work.

A callback sees `model.optimizer.learning_rate` (a float, or a schedule called with the optimizer's
0-based iteration count, as Keras does) and may set `model.stop_training`; `run_fit` is the loop `ClickstreamTransformer.fit`
delegates to.
"""
import json
import math
import os

import numpy as np


# ----------------------------------------------------------------------------- LR schedules
class CustomLRSchedule:
    """rsqrt(d_model) * min(rsqrt(step), step * warmup^-1.5) * scale, and then `* scale` once
    more on return: the reference multiplies by `scale` twice (training_utils.py:31-36) and that
    is kept.  Arithmetic in float32 like the TF ops it stands for."""

    def __init__(self, d_model, warmup_steps=4000, scale=1):
        self.d_model = float(d_model)
        self.warmup_steps = warmup_steps
        self.scale = scale

    def get_config(self):
        return {'d_model': self.d_model, 'warmup_steps': self.warmup_steps, 'scale': self.scale}

    def __call__(self, step):
        step = np.asarray(step, dtype=np.float32)
        with np.errstate(divide='ignore'):
            arg1 = np.float32(1.0) / np.sqrt(step)
        arg2 = step * np.float32(self.warmup_steps ** -1.5)
        lr = (np.float32(1.0) / np.sqrt(np.float32(self.d_model))) * np.minimum(arg1, arg2) \
            * np.float32(self.scale)
        out = (lr * np.float32(self.scale)).astype(np.float32)
        return float(out) if out.ndim == 0 else out


class CustomExponentialDecayLR:
    """(init - limit) * decay_rate^(step / decay_steps) + limit (training_utils.py:39-60);
    get_config keeps the reference's key names."""

    def __init__(self, initial_learning_rate, limiting_learning_rate, decay_steps, decay_rate):
        self.initial_learning_rate = initial_learning_rate
        self.limiting_learning_rate = limiting_learning_rate
        self.decay_steps = decay_steps
        self.decay_rate = decay_rate

    def get_config(self):
        return {'init_lr': self.initial_learning_rate, 'limit_lr': self.limiting_learning_rate,
                'decay_steps': self.decay_steps, 'decay_rate': self.decay_rate}

    def __call__(self, step):
        step = np.asarray(step, dtype=np.float32)
        p = np.power(np.float32(self.decay_rate), step / np.float32(self.decay_steps))
        out = (np.float32(self.initial_learning_rate - self.limiting_learning_rate) * p
               + np.float32(self.limiting_learning_rate)).astype(np.float32)
        return float(out) if out.ndim == 0 else out


def current_learning_rate(optimizer, step):
    """The scalar the Adam kernel is launched with at optimizer iteration `step` (0-based)."""
    lr = optimizer.learning_rate
    return float(lr(step)) if callable(lr) else float(lr)


# ----------------------------------------------------------------------------- callbacks
class Callback:
    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass


def _direction(mode, monitor):
    if mode not in ('auto', 'min', 'max'):
        mode = 'auto'            # Keras warns and falls back
    if mode == 'auto':
        mode = 'max' if 'acc' in monitor else 'min'
    return mode


class ReduceLROnPlateau(Callback):
    """tf.keras.callbacks.ReduceLROnPlateau (2.3): when `monitor` has not improved by more than
    `min_delta` for `patience` epochs, lr <- max(lr * factor, min_lr), then `cooldown` epochs
    during which the wait counter stays at 0.  The reference uses factor=0.317, patience=10
    (main.py:134).  Needs a scalar learning rate (Keras raises on a schedule too)."""

    def __init__(self, monitor='val_loss', factor=0.1, patience=10, verbose=0, mode='auto',
                 min_delta=1e-4, cooldown=0, min_lr=0):
        if factor >= 1.0:
            raise ValueError('ReduceLROnPlateau does not support a factor >= 1.0.')
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.mode, self.min_delta, self.cooldown, self.min_lr = mode, min_delta, cooldown, min_lr
        self._reset()

    def _reset(self):
        if _direction(self.mode, self.monitor) == 'min':
            self.monitor_op = lambda a, b: a < b - self.min_delta
            self.best = math.inf
        else:
            self.monitor_op = lambda a, b: a > b + self.min_delta
            self.best = -math.inf
        self.cooldown_counter = 0
        self.wait = 0

    def in_cooldown(self):
        return self.cooldown_counter > 0

    def on_train_begin(self, logs=None):
        self._reset()

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        opt = self.model.optimizer
        if callable(opt.learning_rate):
            raise TypeError('ReduceLROnPlateau needs a scalar learning rate, not a schedule')
        logs['lr'] = float(opt.learning_rate)
        current = logs.get(self.monitor)
        if current is None:
            return
        if self.in_cooldown():
            self.cooldown_counter -= 1
            self.wait = 0
        if self.monitor_op(current, self.best):
            self.best = current
            self.wait = 0
        elif not self.in_cooldown():
            self.wait += 1
            if self.wait >= self.patience:
                old_lr = float(np.float32(opt.learning_rate))     # Keras keeps lr in float32
                if old_lr > float(np.float32(self.min_lr)):
                    new_lr = max(float(np.float32(old_lr * self.factor)), float(self.min_lr))
                    opt.learning_rate = new_lr
                    if self.verbose:
                        print(f'\nEpoch {epoch + 1:05d}: ReduceLROnPlateau reducing learning '
                              f'rate to {new_lr}.')
                    self.cooldown_counter = self.cooldown
                    self.wait = 0


class EarlyStopping(Callback):
    """tf.keras.callbacks.EarlyStopping (2.3): stop when `monitor` has not improved (by more than
    `min_delta`) for `patience` epochs; the reference uses patience=30 on val_loss (main.py:156)."""

    def __init__(self, monitor='val_loss', min_delta=0, patience=0, verbose=0, mode='auto',
                 baseline=None, restore_best_weights=False):
        self.monitor, self.patience, self.verbose, self.baseline = monitor, patience, verbose, baseline
        self.restore_best_weights = restore_best_weights
        self.mode = _direction(mode, monitor)
        self.min_delta = abs(min_delta) * (1 if self.mode == 'max' else -1)
        self.wait = self.stopped_epoch = 0
        self.best_weights = None
        self.best = None

    def _better(self, a, b):
        return a > b if self.mode == 'max' else a < b

    def on_train_begin(self, logs=None):
        self.wait = self.stopped_epoch = 0
        self.best_weights = None
        if self.baseline is not None:
            self.best = self.baseline
        else:
            self.best = -math.inf if self.mode == 'max' else math.inf

    def on_epoch_end(self, epoch, logs=None):
        current = (logs or {}).get(self.monitor)
        if current is None:
            return
        if self._better(current - self.min_delta, self.best):
            self.best = current
            self.wait = 0
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch = epoch
                self.model.stop_training = True
                if self.restore_best_weights and self.best_weights is not None:
                    self.model.set_weights(self.best_weights)

    def on_train_end(self, logs=None):
        if self.stopped_epoch > 0 and self.verbose:
            print(f'Epoch {self.stopped_epoch + 1:05d}: early stopping')


def _jsonable(obj):
    """Constructor config -> JSON: head units become {class, dense_layer_dims, output_vocab_size},
    in-memory vocabularies their size."""
    if isinstance(obj, dict):
        return {str(k): _jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        if len(obj) > 64:
            return {'len': len(obj)}
        return [_jsonable(v) for v in obj]
    if isinstance(obj, (str, int, float, bool)) or obj is None:
        return obj
    if isinstance(obj, (np.integer, np.floating)):
        return obj.item()
    out = {'class': type(obj).__name__}
    for attr in ('dense_layer_dims', 'output_vocab_size'):
        if hasattr(obj, attr):
            out[attr] = _jsonable(getattr(obj, attr))
    return out


class BestModelSaverCallback(Callback):
    """Writes the model whenever val_loss improves (training_utils.py:63-75).  The reference
    writes a TF SavedModel with its serving signature; here the artefact is `variables.npz` (the
    parameters under the reference's checkpoint keys, model.save_weights) plus `model.json`
    (epoch, val_loss, serving signature and constructor config)."""

    def __init__(self, savedmodel_path):
        self.savedmodel_path = savedmodel_path
        self.best_val_loss = math.inf

    def on_epoch_end(self, epoch, logs=None):
        if logs['val_loss'] < self.best_val_loss:       # KeyError without validation, as upstream
            os.makedirs(self.savedmodel_path, exist_ok=True)
            self.model.save_weights(os.path.join(self.savedmodel_path, 'variables.npz'))
            meta = {'epoch': epoch, 'val_loss': float(logs['val_loss'])}
            if hasattr(self.model, 'get_serving_signature'):
                meta['serving_signature'] = self.model.get_serving_signature()
            if hasattr(self.model, 'get_config'):
                meta['config'] = _jsonable(self.model.get_config())
            with open(os.path.join(self.savedmodel_path, 'model.json'), 'w') as f:
                json.dump(meta, f, indent=1)
            self.best_val_loss = logs['val_loss']


# ----------------------------------------------------------------------------- data parallel
def average_logs(logs, group=None):
    """Mean over the ranks of every scalar in the epoch logs (one small all-reduce; identity
    without an initialised process group).  Under MirroredStrategy Keras evaluates on the global
    batch (main.py:46-57), so every replica's callbacks see the same numbers; with one process per
    GPU each rank validates its own shard, and without this a plateau or early-stopping decision
    could differ between ranks and leave them in different collectives."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return logs
    keys = sorted(k for k, v in logs.items() if isinstance(v, (int, float)))
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(logs[k]) for k in keys], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = (t / dist.get_world_size(group)).cpu()
    out = dict(logs)
    for k, v in zip(keys, t.tolist()):
        out[k] = v
    return out


# ----------------------------------------------------------------------------- the loop
def run_fit(model, dataset, steps_per_epoch, epochs=1, verbose=0, validation_data=None,
            validation_steps=None, callbacks=()):
    """Keras `Model.fit` for an iterator of (inputs, labels) batches: per epoch
    `steps_per_epoch` train steps, then `validation_steps` test steps whose results enter the
    epoch logs as `val_*`, then the callbacks.  Like Keras, the training loss in the logs is the
    running mean of the per-step losses and a metric entry is the metric's own running result;
    metrics are reset at the start of each epoch and before validation.  In a data-parallel run
    the epoch logs are averaged over the ranks before the callbacks see them (average_logs).
    Returns the list of epoch logs (`History.history` transposed)."""
    callbacks = list(callbacks)
    model.stop_training = False
    for cb in callbacks:
        cb.set_model(model)
        cb.on_train_begin()
    history = []
    it = iter(dataset)
    for epoch in range(epochs):
        for m in getattr(model, 'metrics', ()):
            m.reset_states()
        loss_sum, logs = 0.0, {}
        for i in range(steps_per_epoch):
            logs = dict(model.train_step(next(it)))
            loss_sum += logs.get('loss', 0.0)
            logs['loss'] = loss_sum / (i + 1)
        if validation_data is not None:
            for m in getattr(model, 'metrics', ()):
                m.reset_states()
            vit = iter(validation_data)
            vsum, vlast, nv = 0.0, {}, 0
            while validation_steps is None or nv < validation_steps:
                try:
                    batch = next(vit)
                except StopIteration:
                    break
                vlast = model.test_step(batch)
                vsum += vlast.get('loss', 0.0)
                nv += 1
            if nv == 0:
                raise ValueError("validation_data yielded no batches (a generator is exhausted "
                                 "after one epoch: pass a list or another re-iterable)")
            for k, v in vlast.items():
                logs['val_' + k] = vsum / nv if k == 'loss' else v
        logs = average_logs(logs, getattr(model, 'process_group', None))
        for cb in callbacks:
            cb.on_epoch_end(epoch, logs)
        history.append(logs)
        if verbose:
            print(f'Epoch {epoch + 1}/{epochs}', logs)
        if model.stop_training:
            break
    for cb in callbacks:
        cb.on_train_end()
    return history
