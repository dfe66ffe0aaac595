"""Keras-style fit loop on the device: validation logs, the reference's callbacks, weight files in
the reference's variable naming, and re-capture of the graph step when the learning rate changes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

V = 1237


def _model(seed=5, dropout=0.0):
    import bert4clickpath_b200 as bc
    head = bc.SoftMaxHead(dense_layer_dims=[64, 128], output_vocab_size=V)
    return bc.ClickstreamTransformer(
        sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": V},
        embedding_dims={"items": 64}, head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
        num_encoder_layers=1, num_attention_heads=2, dropout_rate=dropout, seed=seed)


def _stream(rng, mode, n=None):
    from bert4clickpath_b200.synthetic import make_cloze_batch
    i = 0
    while n is None or i < n:
        b = make_cloze_batch(rng, 32, V, max_len=30, mode=mode)
        yield {"asin": b["items"]}, b["labels"]
        i += 1


def test_fit_validation_callbacks_and_weight_files(cuda_lib, tmp_path):
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200 import training_utils as T
    model = _model()
    model.compile(optimizer=bc.Adam(3e-3), loss=bc.ClozeMaskedLoss(bc.sparse_categorical_crossentropy),
                  metrics=[bc.ClozeMaskedRecall(k=10)])
    val = list(_stream(np.random.default_rng(1), "eval", 2))
    saver = T.BestModelSaverCallback(str(tmp_path / "savedmodel"))
    # patience 0 on a tiny min_delta-free monitor: every non-improving epoch cuts the lr
    cbs = [T.EarlyStopping(monitor="val_loss", patience=30),
           T.ReduceLROnPlateau(monitor="val_loss", patience=1, factor=0.317, min_delta=10.0), saver]
    hist = model.fit(_stream(np.random.default_rng(0), "train"), steps_per_epoch=4, epochs=4,
                     validation_data=val, callbacks=cbs)
    assert len(hist) == 4 and not model.stop_training
    for h in hist:
        assert set(h) >= {"loss", "val_loss", "Recall_at_10", "val_Recall_at_10", "lr"}
        assert np.isfinite(h["loss"]) and np.isfinite(h["val_loss"])
    assert hist[-1]["loss"] < hist[0]["loss"]
    # min_delta=10 makes every epoch after the first "stale": lr is cut after epochs 1, 2, 3
    f32 = lambda x: float(np.float32(x))
    l1 = f32(f32(3e-3) * 0.317); l2 = f32(l1 * 0.317); l3 = f32(l2 * 0.317)
    assert [h["lr"] for h in hist] == [3e-3, 3e-3, l1, l2] and model.optimizer.learning_rate == l3
    # the saved file holds the reference's variable names and reloads bit-exactly into a new model
    path = tmp_path / "savedmodel" / "variables.npz"
    with np.load(path) as z:
        names = set(z.files)
    assert "transformer/encoder/enc_layers/0/mha/wq/kernel/.ATTRIBUTES/VARIABLE_VALUE" in names
    assert "head/output_layer/kernel/.ATTRIBUTES/VARIABLE_VALUE" in names
    model.save_weights(str(tmp_path / "now.npz"))
    other = _model(seed=99)
    other.load_weights(str(tmp_path / "now.npz"))
    a, b = model.store.get_weights(), other.store.get_weights()
    assert set(a) == set(b) and all(np.array_equal(a[k], b[k]) for k in a)
    other.compile(loss=bc.ClozeMaskedLoss(bc.sparse_categorical_crossentropy))
    model.metrics = []
    assert other.test_step(val[0])["loss"] == model.test_step(val[0])["loss"]
    w = model.get_weights()
    assert "enc.0.wq" in w and w["enc.0.wq"].shape == (64, 64)
    other.set_weights({k: v * 0 for k, v in w.items()})
    assert not other.store.get_weights()["enc.0.wqkv"].any()


def test_schedule_drives_the_update_size(cuda_lib):
    """With a CustomLRSchedule the first update uses lr(0) = 0 (weights unchanged), later ones
    grow with the step; ClozeTrainStep in graph mode refuses a schedule."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200 import training_utils as T
    from bert4clickpath_b200.training import ClozeTrainStep
    model = _model()
    model.compile(optimizer=bc.Adam(T.CustomLRSchedule(d_model=64, warmup_steps=10)))
    it = _stream(np.random.default_rng(0), "train")
    w0 = model.store.get_weights()["head.0.w"]
    model.train_step(next(it))
    w1 = model.store.get_weights()["head.0.w"]
    assert np.array_equal(w0, w1)
    model.train_step(next(it))
    w2 = model.store.get_weights()["head.0.w"]
    model.train_step(next(it))
    w3 = model.store.get_weights()["head.0.w"]
    s1, s2 = np.abs(w2 - w1).max(), np.abs(w3 - w2).max()
    lr1, lr2 = T.CustomLRSchedule(64, 10)(1), T.CustomLRSchedule(64, 10)(2)
    assert 0 < s1 <= lr1 * 1.01 + 1e-9 and s1 < s2 <= lr2 * 1.5
    step = ClozeTrainStep(model, model.optimizer, use_graph=True)
    from bert4clickpath_b200.synthetic import make_cloze_batch
    b = make_cloze_batch(np.random.default_rng(2), 32, V, max_len=30, mode="train")
    with pytest.raises(TypeError):
        step.step_device(step.to_device(b))


def test_graph_step_recaptures_when_the_learning_rate_changes(cuda_lib):
    """Adam's scalars are launch parameters baked into the captured graph: after a callback changes
    lr the replayed step must use the new value — bit-identical to an eager twin."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200 import ops
    from bert4clickpath_b200.synthetic import make_cloze_batch
    from bert4clickpath_b200.training import ClozeTrainStep
    rng = np.random.default_rng(0)
    batches = [make_cloze_batch(rng, 32, V, max_len=30, mode="train") for _ in range(2)]
    mg, me = _model(dropout=0.1), _model(dropout=0.1)
    og, oe = bc.Adam(1e-3), bc.Adam(1e-3)
    tg, te = ClozeTrainStep(mg, og, use_graph=True), ClozeTrainStep(me, oe)
    for i in range(8):
        if i == 5:
            og.learning_rate = oe.learning_rate = 1e-4
        b = batches[i % 2]
        sg = tg.step_device(tg.to_device(b)).clone()
        se = te._eager(te.to_device(b), ops.device_seed(me.store.step_dev)).clone()
        torch.cuda.synchronize()
        assert torch.equal(sg, se), i
    st = next(iter(tg._graphs.values()))
    assert st["graph"] is not None and st["hp"][0] == 1e-4
    wg, we = mg.store.get_weights(), me.store.get_weights()
    assert all(np.array_equal(wg[k], we[k]) for k in wg)


@pytest.mark.parametrize("use_graph", [False, True])
def test_pipelined_host_loop_equals_the_blocking_one(cuda_lib, use_graph):
    """run_host (copies on a second stream, losses read one step late) must train exactly like
    step_host called once per batch: same losses, same weights, with batches of different mask
    counts and a partial last pair of landing buffers."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.synthetic import make_cloze_batch
    from bert4clickpath_b200.training import ClozeTrainStep
    rng = np.random.default_rng(3)
    host = [make_cloze_batch(rng, 32, V, max_len=30, mode="train", lengths="beauty") for _ in range(3)]
    pinned = [(torch.from_numpy(b["ids"]).pin_memory(), torch.from_numpy(b["labels"]).pin_memory(),
               b["n_masked"]) for b in host]
    order = [0, 1, 2, 1, 0, 2, 2]
    ma, mb = _model(dropout=0.0), _model(dropout=0.0)
    ta = ClozeTrainStep(ma, bc.Adam(1e-3), use_graph=use_graph)
    tb = ClozeTrainStep(mb, bc.Adam(1e-3), use_graph=use_graph)
    blocking = [ta.step_host(*pinned[i]) for i in order]
    piped = list(tb.run_host(pinned[i] for i in order))
    assert piped == blocking
    wa, wb = ma.store.get_weights(), mb.store.get_weights()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)
    assert list(tb.run_host(iter(()))) == []


def test_weight_stream_schedule_is_bit_identical(cuda_lib):
    """Weight gradients on the second stream (engine.WeightStream) are the same kernels in the
    same order per tensor: every gradient and the loss must equal the single-stream schedule
    bit for bit, eagerly and through the captured graph, with dropout on."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200 import engine
    from bert4clickpath_b200.synthetic import make_cloze_batch
    from bert4clickpath_b200.training import ClozeTrainStep
    rng = np.random.default_rng(11)
    batches = [make_cloze_batch(rng, 32, V, max_len=30, mode="train") for _ in range(2)]
    results = []
    was = engine.WSTREAM.enabled
    try:
        for enabled in (True, False):
            engine.WSTREAM.enabled = enabled
            m = _model(dropout=0.1)
            t = ClozeTrainStep(m, bc.Adam(1e-3), use_graph=True)
            losses = []
            for i in range(6):     # 2 eager warm-up steps, capture, replays
                st = t.step_device(t.to_device(batches[i % 2])).clone()
                torch.cuda.synchronize()
                losses.append(st.cpu().numpy().copy())
            results.append((losses, m.store.get_grads(), m.store.get_weights()))
    finally:
        engine.WSTREAM.enabled = was
    (la, ga, wa), (lb, gb, wb) = results
    assert all(np.array_equal(x, y) for x, y in zip(la, lb))
    assert all(np.array_equal(ga[k], gb[k]) for k in ga)
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)


def test_string_batches_through_encode_host_train_like_the_id_batches(cuda_lib):
    """The reference-shaped input ({raw feature: (B, L) strings}, padded labels) through
    encode_host (chaining + native vocabulary lookup into pinned memory) + run_host must give the
    ids of the pre-chained batch and train identically: same losses, same weights - for '<U'
    arrays and for object arrays, including an out-of-vocabulary string."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.constants import RESERVED_TOKENS
    from bert4clickpath_b200.synthetic import make_cloze_batch
    from bert4clickpath_b200.training import ClozeTrainStep
    rng = np.random.default_rng(8)
    host = [make_cloze_batch(rng, 32, V, max_len=30, mode="train", lengths="beauty") for _ in range(3)]
    tok = np.asarray(list(RESERVED_TOKENS) + [f"item_{j}" for j in range(V)] + ["never-seen"], dtype=np.str_)
    col = int(np.nonzero(host[1]["ids"][0] >= 10)[0][0])     # an ordinary item (not [MASK] / [PAD])
    host[1]["ids"][0, col] = len(tok) - 1                    # -> the OOV bucket = 10 + V
    strings = [({"asin": tok[b["ids"][:, 2:-1]]}, b["labels"]) for b in host]
    strings[2] = ({"asin": strings[2][0]["asin"].astype(object)}, strings[2][1])
    pinned = [(torch.from_numpy(b["ids"]).pin_memory(), torch.from_numpy(b["labels"]).pin_memory(),
               b["n_masked"]) for b in host]
    order = [0, 1, 2, 1, 0]
    ma, mb = _model(dropout=0.0), _model(dropout=0.0)
    ta, tb = ClozeTrainStep(ma, bc.Adam(1e-3), use_graph=True), ClozeTrainStep(mb, bc.Adam(1e-3), use_graph=True)
    for i in range(3):
        ids_p, lab_p, n = tb.encode_host(*strings[i], parity=i)
        assert np.array_equal(ids_p.numpy(), host[i]["ids"]) and n == host[i]["n_masked"]
        assert np.array_equal(lab_p.numpy(), host[i]["labels"])
    want = list(ta.run_host(pinned[i] for i in order))
    got = list(tb.run_host(tb.encode_host(*strings[i], parity=k) for k, i in enumerate(order)))
    assert got == want
    wa, wb = ma.store.get_weights(), mb.store.get_weights()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)
