"""Host-side training control (SURVEY §8f N1 callbacks, N4 weight names): pure-Python logic, no GPU.

The Keras callbacks are checked against hand-worked traces of the tf.keras 2.3 rules (patience
counting, min_delta, cooldown, float32 learning rate); the schedules against the closed forms of
clickstream_transformer/training_utils.py:31-36 and :58-60."""
import math

import numpy as np
import pytest

from bert4clickpath_b200 import training_utils as T
from bert4clickpath_b200 import weights as W


class _Opt:
    def __init__(self, lr):
        self.learning_rate = lr


class _FakeModel:
    """train_step / test_step replay scripted losses; everything else is what run_fit touches."""

    def __init__(self, train_losses, val_losses, lr=1e-3):
        self.optimizer = _Opt(lr)
        self.metrics = []
        self.stop_training = False
        self._t, self._v = iter(train_losses), iter(val_losses)
        self.lrs_seen = []
        self.weights = {"w": np.zeros(1)}
        self.saved = []

    def train_step(self, batch):
        self.lrs_seen.append(self.optimizer.learning_rate)
        self.weights = {"w": self.weights["w"] + 1}
        return {"loss": next(self._t)}

    def test_step(self, batch):
        return {"loss": next(self._v)}

    def get_weights(self):
        return {k: v.copy() for k, v in self.weights.items()}

    def set_weights(self, w):
        self.weights = {k: v.copy() for k, v in w.items()}

    def save_weights(self, path):
        self.saved.append((path, float(self.weights["w"][0])))

    def get_serving_signature(self):
        return {"asin": ([None, None], "string")}

    def get_config(self):
        class _Head:
            dense_layer_dims, output_vocab_size = [8, 4], 99
        return {"head_unit": _Head(), "feature_vocabs": {"items": [str(i) for i in range(100)]},
                "dropout_rate": np.float32(0.5)}


def _batches():
    while True:
        yield None


def test_reduce_lr_on_plateau_follows_the_keras_rules():
    cb = T.ReduceLROnPlateau(monitor="val_loss", patience=2, factor=0.317)
    m = _FakeModel([], [])
    cb.set_model(m)
    cb.on_train_begin()
    # improvement must exceed min_delta=1e-4: 0.99995 after 1.0 does not count
    trace = [1.0, 0.99995, 0.99994, 0.5, 0.5, 0.5, 0.5]
    lrs = []
    for e, v in enumerate(trace):
        logs = {"val_loss": v}
        cb.on_epoch_end(e, logs)
        lrs.append(m.optimizer.learning_rate)
        assert "lr" in logs
    f32 = lambda x: float(np.float32(x))
    l1 = f32(f32(1e-3) * 0.317)
    l2 = f32(l1 * 0.317)
    #       best   wait1   wait2->cut  best  wait1  wait2->cut  wait1
    assert lrs == [1e-3, 1e-3, l1, l1, l1, l2, l2]
    with pytest.raises(ValueError):
        T.ReduceLROnPlateau(factor=1.0)


def test_reduce_lr_cooldown_and_min_lr():
    cb = T.ReduceLROnPlateau(monitor="val_loss", patience=1, factor=0.5, cooldown=2, min_lr=4e-4)
    m = _FakeModel([], [])
    cb.set_model(m)
    cb.on_train_begin()
    lrs = []
    for e in range(8):
        cb.on_epoch_end(e, {"val_loss": 1.0})
        lrs.append(m.optimizer.learning_rate)
    # e0 best; e1 wait=1 -> 5e-4, cooldown 2; e2 (cd 2->1, still cooling), e3 (cd 1->0, wait 1 -> cut
    # to max(2.5e-4, 4e-4) = 4e-4); afterwards lr == min_lr: no further change
    assert lrs[0] == 1e-3 and lrs[1] == pytest.approx(5e-4) and lrs[2] == pytest.approx(5e-4)
    assert lrs[3] == pytest.approx(4e-4) and all(l == pytest.approx(4e-4) for l in lrs[3:])
    m.optimizer.learning_rate = T.CustomLRSchedule(64)
    with pytest.raises(TypeError):
        cb.on_epoch_end(9, {"val_loss": 1.0})


def test_early_stopping_patience_and_restore():
    val = [1.0, 0.9, 0.95, 0.91, 0.92, 0.1]
    m = _FakeModel([0.0] * 10, val)
    es = T.EarlyStopping(monitor="val_loss", patience=3, restore_best_weights=True)
    hist = T.run_fit(m, _batches(), steps_per_epoch=1, epochs=10, validation_data=[None],
                     validation_steps=1, callbacks=[es])
    # best at epoch 1 (0.9); epochs 2, 3, 4 do not improve -> stop after epoch index 4
    assert len(hist) == 5 and es.stopped_epoch == 4 and m.stop_training
    assert [h["val_loss"] for h in hist] == val[:5]
    assert m.weights["w"][0] == 2          # weights of the best epoch (two train steps done)
    # min_delta: an improvement smaller than it does not reset the wait
    es2 = T.EarlyStopping(monitor="val_loss", patience=1, min_delta=0.1)
    m2 = _FakeModel([], [])
    es2.set_model(m2)
    es2.on_train_begin()
    es2.on_epoch_end(0, {"val_loss": 1.0})
    es2.on_epoch_end(1, {"val_loss": 0.95})
    assert m2.stop_training and es2.stopped_epoch == 1
    # monitor missing: nothing happens
    es3 = T.EarlyStopping(patience=0)
    es3.set_model(_FakeModel([], []))
    es3.on_train_begin()
    es3.on_epoch_end(0, {"loss": 1.0})
    assert not es3.model.stop_training


def test_fit_loop_order_reference_callbacks(tmp_path):
    """The reference's callback set (main.py:134-157): val loss plateaus, lr is cut after 10 stale
    epochs, training stops after 30, the best model is written only on improvement."""
    n = 60
    val = [1.0 - 0.1 * e for e in range(5)] + [0.7] * (n - 5)
    m = _FakeModel([2.0, 1.0] * n, val)
    saver = T.BestModelSaverCallback(str(tmp_path / "savedmodel"))
    cbs = [T.EarlyStopping(monitor="val_loss", patience=30),
           T.ReduceLROnPlateau(monitor="val_loss", patience=10, factor=0.317), saver]
    hist = T.run_fit(m, _batches(), steps_per_epoch=2, epochs=n, validation_data=[None],
                     callbacks=cbs)
    assert all(h["loss"] == 1.5 for h in hist)                 # running mean of the step losses
    assert len(hist) == 5 + 30                                 # best at epoch 4, 30 stale epochs
    assert [s[1] for s in m.saved] == [2.0, 4.0, 6.0, 8.0, 10.0]
    import json
    meta = json.load(open(tmp_path / "savedmodel" / "model.json"))
    assert meta["epoch"] == 4 and meta["val_loss"] == pytest.approx(0.6)
    assert meta["serving_signature"] == {"asin": [[None, None], "string"]}
    assert meta["config"]["head_unit"] == {"class": "_Head", "dense_layer_dims": [8, 4], "output_vocab_size": 99}
    assert meta["config"]["feature_vocabs"] == {"items": {"len": 100}}
    f32 = lambda x: float(np.float32(x))
    l1 = f32(f32(1e-3) * 0.317); l2 = f32(l1 * 0.317); l3 = f32(l2 * 0.317)
    # cuts after epochs 14, 24, 34 (0-based); epoch e trains with the lr left by epoch e-1
    assert m.lrs_seen[2 * 14] == 1e-3 and m.lrs_seen[2 * 15] == l1 and m.lrs_seen[2 * 25] == l2
    assert m.optimizer.learning_rate == l3 and hist[-1]["lr"] == l2
    with pytest.raises(KeyError):                               # upstream: logs['val_loss'] unguarded
        T.run_fit(_FakeModel([1.0], []), _batches(), 1, callbacks=[T.BestModelSaverCallback("x")])


def test_fit_rejects_an_exhausted_validation_generator():
    m = _FakeModel([1.0] * 4, [1.0])
    gen = (b for b in [None])
    with pytest.raises(ValueError):
        T.run_fit(m, _batches(), steps_per_epoch=1, epochs=2, validation_data=gen)


def test_lr_schedules_closed_forms():
    s = T.CustomLRSchedule(d_model=64, warmup_steps=4000)
    steps = np.array([1, 100, 4000, 40000], dtype=np.float32)
    want = 64 ** -0.5 * np.minimum(steps ** -0.5, steps * 4000 ** -1.5)
    np.testing.assert_allclose(s(steps), want, rtol=1e-6)
    assert s(0) == 0.0                                         # rsqrt(0)=inf, min(inf, 0) = 0
    # `scale` enters twice (training_utils.py:33 and :36)
    assert T.CustomLRSchedule(64, scale=3)(100) == pytest.approx(9 * s(100), rel=1e-6)
    assert s.get_config() == {"d_model": 64.0, "warmup_steps": 4000, "scale": 1}
    d = T.CustomExponentialDecayLR(1e-3, 1e-5, decay_steps=1000, decay_rate=0.5)
    assert d(0) == pytest.approx(1e-3) and d(1000) == pytest.approx(5.05e-4, rel=1e-6)
    assert d(1e7) == pytest.approx(1e-5)
    assert set(d.get_config()) == {"init_lr", "limit_lr", "decay_steps", "decay_rate"}
    assert T.current_learning_rate(_Opt(d), 1000) == pytest.approx(5.05e-4, rel=1e-6)
    assert T.current_learning_rate(_Opt(0.01), 5) == 0.01


def test_reference_checkpoint_key_map_round_trips():
    rng = np.random.default_rng(0)
    d, dff = 8, 12
    ref = {"emb.items": rng.normal(size=(21, 6)), "emb.events": rng.normal(size=(13, 2))}
    for l in range(2):
        for n in "qkv":
            ref[f"enc.{l}.w{n}"] = rng.normal(size=(d, d)); ref[f"enc.{l}.b{n}"] = rng.normal(size=d)
        ref[f"enc.{l}.wo"] = rng.normal(size=(d, d)); ref[f"enc.{l}.bo"] = rng.normal(size=d)
        ref[f"enc.{l}.w1"] = rng.normal(size=(d, dff)); ref[f"enc.{l}.b1"] = rng.normal(size=dff)
        ref[f"enc.{l}.w2"] = rng.normal(size=(dff, d)); ref[f"enc.{l}.b2"] = rng.normal(size=d)
        for k in ("ln1", "ln2"):
            ref[f"enc.{l}.{k}_g"] = rng.normal(size=d); ref[f"enc.{l}.{k}_b"] = rng.normal(size=d)
    ref["head.0.w"] = rng.normal(size=(d, 5)); ref["head.0.b"] = rng.normal(size=5)
    ref["head.out.w"] = rng.normal(size=(5, 10)); ref["head.out.b"] = rng.normal(size=10)
    ref = {k: v.astype(np.float32) for k, v in ref.items()}
    store = W.to_store_layout(ref)
    assert store["enc.1.wqkv"].shape == (d, 3 * d) and "enc.1.wq" not in store
    tf_vars = W.export_reference_variables(store)
    keys = set(tf_vars)
    sfx = "/.ATTRIBUTES/VARIABLE_VALUE"
    # attribute-path keys of the reference's object graph (transformer.py:112-116, :181-184, :245,
    # :338, :346; head.py:10-11)
    for k in ["transformer/embedding_layers/items/embeddings",
              "transformer/encoder/enc_layers/1/mha/wq/kernel",
              "transformer/encoder/enc_layers/0/mha/dense/bias",
              "transformer/encoder/enc_layers/0/ffn/layer_with_weights-1/kernel",
              "transformer/encoder/enc_layers/1/layernorm2/gamma",
              "head/intermediate_layers/0/kernel", "head/output_layer/bias"]:
        assert k + sfx in keys
    assert len(keys) == len(ref)
    tf_vars["optimizer/iter" + sfx] = np.array(3)                    # ignored on import
    tf_vars["optimizer/transformer/x/.OPTIMIZER_SLOT/m" + sfx] = np.zeros(2)
    tf_vars["save_counter" + sfx] = np.array(1)
    back = W.import_reference_variables(tf_vars)
    assert set(back) == set(store)
    for k in store:
        np.testing.assert_array_equal(back[k], store[k])
    for k, v in W.to_reference_layout(back).items():
        np.testing.assert_array_equal(v, ref[k])
    with pytest.raises(KeyError):
        W.tf_checkpoint_key("decoder.0.w")
    # the store registers tables by POSITION (emb.0, emb.1); the reference's checkpoint keys them
    # by FEATURE NAME: with the model's feature list the round trip lands on the store's names
    feats = ["items", "events"]
    pos = {("emb.0" if k == "emb.items" else "emb.1" if k == "emb.events" else k): v
           for k, v in store.items()}
    tf2 = W.export_reference_variables(pos, feats)
    assert set(tf2) == keys                          # same feature-named keys as above
    back2 = W.import_reference_variables(tf_vars, feats)   # a REAL reference checkpoint's keys
    assert set(back2) == set(pos)
    for k in pos:
        np.testing.assert_array_equal(back2[k], pos[k])
    with pytest.raises(KeyError):
        W.import_reference_variables({"transformer/embedding_layers/basket/embeddings" + sfx: np.zeros((2, 2))}, feats)


def _fit_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0 sees an improving validation loss, rank 1 a flat one: alone, rank 1 would cut its
    # learning rate and stop early while rank 0 trains on
    val = [1.0, 0.8, 0.6, 0.4, 0.2, 0.1] if rank == 0 else [1.0] * 6
    m = _FakeModel([0.5 + rank] * 6, val)
    m.process_group = None
    es = T.EarlyStopping(monitor="val_loss", patience=2)
    rl = T.ReduceLROnPlateau(monitor="val_loss", patience=1, factor=0.5)
    hist = T.run_fit(m, _batches(), steps_per_epoch=1, epochs=6, validation_data=[None],
                     validation_steps=1, callbacks=[es, rl])
    q.put((rank, [h["val_loss"] for h in hist], [h["loss"] for h in hist],
           m.optimizer.learning_rate, m.stop_training))
    dist.destroy_process_group()


def test_two_rank_gloo_fit_keeps_callback_decisions_identical():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fit_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, v0, l0, lr0, stop0), (_, v1, l1, lr1, stop1) = res
    assert v0 == v1 == pytest.approx([1.0, 0.9, 0.8, 0.7, 0.6, 0.55])
    assert l0 == l1 == pytest.approx([1.0] * 6)
    assert lr0 == lr1 == 1e-3 and stop0 is stop1 is False
    assert T.average_logs({"a": 1.0, "tag": "x"}) == {"a": 1.0, "tag": "x"}   # no process group
