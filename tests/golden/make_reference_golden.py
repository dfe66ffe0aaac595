"""Golden vectors from the REFERENCE'S OWN SOURCE, executed in this container.

    python tests/golden/make_reference_golden.py          # writes tests/golden/reference_*.npz

TensorFlow 2.3.1 cannot be installed here, so the reference's modules are imported unmodified
from /root/reference (clickstream_transformer/{clickstream_transformer,transformer,head,losses,
metrics,training_utils}.py and examples/BERT4Rec/source/{utils,input_pipeline}.py) on top of
tests/golden/tf_shim/tensorflow - a small eager implementation of the TensorFlow calls they make
(torch CPU tensors, torch autograd for the gradients).  The reference decides WHAT is computed;
the shim only says what each TensorFlow op means.  Every case is produced twice: with
tf.float32 = float32 (what TensorFlow would hold, up to summation order) and, TFSHIM_FLOAT64=1,
with the same graph in float64 (what the oracle's float64 mode must reproduce to ~1e-12).

/root/reference does not exist on the GPU box and TensorFlow exists nowhere: the outputs are
committed.  tests/test_reference_golden.py pins the NumPy oracle to them on CPU;
tests/test_zz_reference_golden_gpu.py compares the CUDA path with them directly.

Cases
  cloze        create_cloze_dataset (generator source, train + eval modes) -> ClickstreamTransformer(
               value_to_head='[MASK]', SoftMaxHead) -> ClozeMaskedLoss / ClozeMaskedNDCG /
               ClozeMaskedRecall; forward, loss, every gradient; inference and training (dropout
               masks recorded).  `tf.random.shuffle` is given the keyed permutation the device
               batch builder uses (oracle.keyed_mask_positions), so the batches are the reference's
               masking code applied to known positions.
  segment      two chained features, segment_to_head -> BinaryClassificationHead ->
               MaskedLoss(K.binary_crossentropy, pos_weight) + PositiveRate / PredictedPositives /
               F1Score
  multilabel   segment 0 ([CLS]) -> MultiLabel_MultiClass_classification -> MaskedLoss(BCE)
  schedules    CustomLRSchedule / CustomExponentialDecayLR values (float32 only)
  masking      random_item_mask's count rule on a 90-item session at 0.7 (float32 product -> 63)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
SEED = 20260118


def _import_reference():
    sys.path[:0] = [os.path.join(HERE, "tf_shim"), REF, os.path.join(REF, "examples", "BERT4Rec"), ROOT]
    import tensorflow as tf
    assert tf.__version__.endswith("shim")
    return tf


def _write_vocab(tmp, name, tokens):
    path = os.path.join(tmp, name)
    with open(path, "w") as f:
        f.write("\n".join(tokens) + "\n")
    return path


def _snapshot(tf, model):
    """{checkpoint key: (value, gradient)} over the reference model's object graph."""
    out = {}
    for path, p in tf.tracked_variables(model).items():
        g = p.tensor.grad
        out[path + SUFFIX] = (p.tensor.detach().numpy().copy(),
                              None if g is None else g.detach().numpy().copy())
    return out


def _zero_grads(tf, model):
    for p in tf.tracked_variables(model).values():
        p.tensor.grad = None


def _perturb_1d(tf, model, rng):
    """Biases / LayerNorm parameters off their 0 / 1 initial values (float32 values)."""
    for p in tf.tracked_variables(model).values():
        if p.tensor.dim() == 1:
            p.assign(p.tensor.detach().numpy().astype(np.float32)
                     + rng.normal(scale=0.05, size=tuple(p.tensor.shape)).astype(np.float32))


def _pack(prefix, snap, out):
    """Parameters once (they do not change between the passes of a case), gradients per pass."""
    for k, (v, g) in snap.items():
        if f"param:{k}" in out:
            assert np.array_equal(out[f"param:{k}"], v)
        out[f"param:{k}"] = v
        if g is not None:
            out[f"{prefix}grad:{k}"] = g


# ----------------------------------------------------------------------------------- cloze
def case_cloze(tf, tmp):
    from clickstream_transformer.clickstream_transformer import ClickstreamTransformer
    from clickstream_transformer.constants import INPUT_MASKING_TOKEN, LABEL_PAD
    from clickstream_transformer.head import SoftMaxHead
    from source.cloze_constants import modes
    from source.input_pipeline import create_cloze_dataset
    from source.utils import ClozeMaskedLoss, ClozeMaskedNDCG, ClozeMaskedRecall
    from oracle import clickpath_oracle as O

    rng = np.random.default_rng(SEED)
    V = 37
    vocab = [f"B00{j:04d}" for j in range(V)]
    vocab_file = _write_vocab(tmp, "item_vocab.txt", vocab)
    lengths = [2, 3, 5, 9, 14, 7, 4, 26, 11, 6, 3, 8]
    sessions = [rng.integers(0, V, size=n).tolist() for n in lengths]
    sessions[7][3] = V + 5          # an item that is not in the vocabulary -> the OOV bucket
    names = vocab + [f"UNSEEN{j}" for j in range(10)]
    current = {"session": None}

    def source():
        for s, items in enumerate(sessions):
            current["session"] = s
            yield {"asin": [names[i] for i in items], "reviewerID": f"user{s}"}

    def keyed_permutation(n):
        # all n positions ordered by the builder's key: its first k entries are the k smallest keys
        return O.keyed_mask_positions_order(SEED, current["session"], n)

    tf.random.set_shuffle(keyed_permutation)
    B = 6
    out = {"vocab": np.array(vocab), "seed": np.int64(SEED), "batch_size": np.int64(B),
           "sessions_flat": np.concatenate(sessions).astype(np.int64),
           "session_lengths": np.array(lengths, dtype=np.int64)}
    batches = {}
    for mode in (modes.TRAIN, modes.EVAL):
        ds = create_cloze_dataset(source, mode, B, vocab_file)
        for i, (features, labels) in enumerate(ds.take(2)):
            batches[(mode, i)] = (features, labels)
            out[f"{mode}{i}:asin"] = features["asin"].astype(str)
            out[f"{mode}{i}:labels"] = labels.numpy()

    tf.set_initializer_seed(SEED)
    head = SoftMaxHead(dense_layer_dims=[24, 12], output_vocab_size=V)
    model = ClickstreamTransformer(
        sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": vocab_file},
        embedding_dims={"items": 16}, head_unit=head, value_to_head=INPUT_MASKING_TOKEN,
        num_encoder_layers=2, num_attention_heads=2, dropout_rate=0.1)
    out["config"] = np.array(json.dumps(dict(V=V, d=16, layers=2, heads=2, dff=100, head=[24, 12],
                                             dropout=0.1)))
    feats0, labels0 = batches[(modes.TRAIN, 0)]
    model({"asin": feats0["asin"]}, training=False)          # builds the lazily created weights
    _perturb_1d(tf, model, rng)
    loss_fn = ClozeMaskedLoss(tf.keras.backend.sparse_categorical_crossentropy, label_pad=LABEL_PAD)

    # what the lookup tables make of the chained strings (ids the encoder sees)
    raw, _, _ = model.transformer_input_prep(features={"asin": feats0["asin"]})
    out["train0:ids"] = model.vocab_lookup_tables["items"].lookup(raw["items"]).numpy()

    for key, training in (("infer", False), ("train", True)):
        for (mode, i), (features, labels) in batches.items():
            if training and (mode, i) != (modes.TRAIN, 0):
                continue
            tag = f"{key}:{mode}{i}:"
            _zero_grads(tf, model)
            del tf.DROPOUT_LOG[:]
            tf.random.set_seed(SEED + 1)
            y_pred = model({"asin": features["asin"]}, training=training)
            loss = loss_fn(labels, y_pred)
            loss.backward()
            out[tag + "probs"] = y_pred.numpy()
            out[tag + "loss"] = loss.numpy()
            _pack(tag, _snapshot(tf, model), out)
            if training:
                # call order: Encoder.dropout, then (dropout1, dropout2) of every layer
                keys = ["in"] + [f"{l}.{j}" for l in range(2) for j in (1, 2)]
                assert len(tf.DROPOUT_LOG) == len(keys)
                for k, (_, m) in zip(keys, tf.DROPOUT_LOG):
                    out[tag + "dropout:" + k] = m
            else:
                for k in (1, 5, 10):
                    nd, rc = ClozeMaskedNDCG(k=k), ClozeMaskedRecall(k=k)
                    nd.update_state(labels, y_pred)
                    rc.update_state(labels, y_pred)
                    out[tag + f"ndcg@{k}"] = nd.result().numpy()
                    out[tag + f"recall@{k}"] = rc.result().numpy()
                    out[tag + f"ndcg_sum@{k}"] = nd.ndcg.numpy()
                    out[tag + f"n_examples@{k}"] = nd.n_examples.numpy()
    tf.random.set_shuffle(None)
    return out


# --------------------------------------------------------------------------------- segment
def _two_feature_inputs(rng, B, L1, L2, n_items, n_events):
    item_names = [f"item{j}" for j in range(n_items)]
    event_names = [f"ev{j}" for j in range(n_events)]

    def seq(n_tok, names, L):
        a = np.empty((B, L), dtype=object)
        a[...] = "[PAD]"
        lens = rng.integers(1, L + 1, size=B)
        lens[0] = L
        for b in range(B):
            a[b, :lens[b]] = [names[i] for i in rng.integers(0, n_tok, size=lens[b])]
        return a, lens

    s_items, l1 = seq(n_items, item_names, L1)
    b_items, l2 = seq(n_items, item_names, L2)
    s_events = np.where(s_items == "[PAD]", "[PAD]", None)
    b_events = np.where(b_items == "[PAD]", "[PAD]", None)
    for arr in (s_events, b_events):
        idx = np.argwhere(arr == None)  # noqa: E711
        for r, c in idx:
            arr[r, c] = event_names[rng.integers(0, n_events)]
    return item_names, event_names, dict(s_items=s_items, b_items=b_items, s_events=s_events,
                                         b_events=b_events), l1, l2


def case_segment(tf, tmp):
    from clickstream_transformer.clickstream_transformer import ClickstreamTransformer
    from clickstream_transformer.head import BinaryClassificationHead
    from clickstream_transformer.losses import MaskedLoss
    from clickstream_transformer.metrics import F1Score, PositiveRate, PredictedPositives
    rng = np.random.default_rng(SEED + 2)
    B, L1, L2 = 5, 9, 6
    items, events, feats, l1, l2 = _two_feature_inputs(rng, B, L1, L2, 29, 5)
    item_file = _write_vocab(tmp, "items.txt", items)
    event_file = _write_vocab(tmp, "events.txt", events)
    tf.set_initializer_seed(SEED + 2)
    head = BinaryClassificationHead(dense_layer_dims=[8])
    model = ClickstreamTransformer(
        sequential_input_config={"items": ["s_items", "b_items"], "events": ["s_events", "b_events"]},
        feature_vocabs={"items": item_file, "events": event_file},
        embedding_dims={"items": 12, "events": 4}, head_unit=head, segment_to_head=2,
        num_encoder_layers=2, num_attention_heads=4, dropout_rate=0.0)
    model(feats, training=False)
    _perturb_1d(tf, model, rng)
    y = np.full((B, L2), -1.0, dtype=np.float32)
    for b in range(B):
        y[b, :l2[b]] = rng.integers(0, 2, size=l2[b])
    out = {"item_vocab": np.array(items), "event_vocab": np.array(events), "labels": y,
           "config": np.array(json.dumps(dict(d_items=12, d_events=4, layers=2, heads=4, dff=100,
                                              head=[8], segment=2, pos_weight=3.0)))}
    for k, v in feats.items():
        out["feature:" + k] = v.astype(str)
    raw, starts, ends = model.transformer_input_prep(features=feats)
    out["segment_starts"], out["segment_ends"] = starts.numpy(), ends.numpy()
    for f in ("items", "events"):
        out["ids:" + f] = model.vocab_lookup_tables[f].lookup(raw[f]).numpy()
    for tag, pw in (("pw3:", 3.0), ("pw_none:", None)):
        _zero_grads(tf, model)
        probs = model(feats, training=False)
        loss = MaskedLoss(tf.keras.backend.binary_crossentropy, pos_weight=pw)(y, probs)
        loss.backward()
        out[tag + "probs"], out[tag + "loss"] = probs.numpy(), loss.numpy()
        _pack(tag, _snapshot(tf, model), out)
    for name, metric in (("positive_rate", PositiveRate()), ("pred_positives", PredictedPositives()),
                         ("f1", F1Score())):
        metric.update_state(tf.convert_to_tensor(y), probs.detach())
        out["metric:" + name] = metric.result().numpy()
    return out


def case_multilabel(tf, tmp):
    from clickstream_transformer.clickstream_transformer import ClickstreamTransformer
    from clickstream_transformer.head import MultiLabel_MultiClass_classification
    from clickstream_transformer.losses import MaskedLoss
    rng = np.random.default_rng(SEED + 3)
    B, L1, L2, C = 4, 7, 3, 11
    items, events, feats, _, _ = _two_feature_inputs(rng, B, L1, L2, 19, 4)
    item_file = _write_vocab(tmp, "items.txt", items)
    event_file = _write_vocab(tmp, "events.txt", events)
    tf.set_initializer_seed(SEED + 3)
    head = MultiLabel_MultiClass_classification(dense_layer_dims=[10], output_vocab_size=C)
    model = ClickstreamTransformer(
        sequential_input_config={"items": ["s_items", "b_items"], "events": ["s_events", "b_events"]},
        feature_vocabs={"items": item_file, "events": event_file},
        embedding_dims={"items": 12, "events": 4}, head_unit=head, segment_to_head=0,
        num_encoder_layers=1, num_attention_heads=2, dropout_rate=0.0)
    model(feats, training=False)
    _perturb_1d(tf, model, rng)
    y = rng.integers(0, 2, size=(B, C)).astype(np.float32)
    y[1, 4:] = -1.0
    out = {"item_vocab": np.array(items), "event_vocab": np.array(events), "labels": y,
           "config": np.array(json.dumps(dict(d_items=12, d_events=4, layers=1, heads=2, dff=100,
                                              head=[10], classes=C, segment=0, pos_weight=2.0)))}
    for k, v in feats.items():
        out["feature:" + k] = v.astype(str)
    raw, _, _ = model.transformer_input_prep(features=feats)
    for f in ("items", "events"):
        out["ids:" + f] = model.vocab_lookup_tables[f].lookup(raw[f]).numpy()
    _zero_grads(tf, model)
    probs = model(feats, training=False)
    loss = MaskedLoss(tf.keras.backend.binary_crossentropy, pos_weight=2.0)(y, probs)
    loss.backward()
    out["probs"], out["loss"] = probs.numpy(), loss.numpy()
    _pack("", _snapshot(tf, model), out)
    return out


def case_schedules_and_masking(tf, tmp):
    from clickstream_transformer.training_utils import CustomExponentialDecayLR, CustomLRSchedule
    from source.input_pipeline import random_item_mask
    steps = np.array([1, 2, 10, 100, 3999, 4000, 4001, 10000, 123456], dtype=np.float32)
    out = {"steps": steps}
    out["custom_lr:d64_w4000_s1"] = CustomLRSchedule(d_model=64)(tf.convert_to_tensor(steps)).numpy()
    out["custom_lr:d20_w100_s2"] = CustomLRSchedule(d_model=20, warmup_steps=100, scale=2)(
        tf.convert_to_tensor(steps)).numpy()
    out["exp_decay:1e-3_1e-5_1000_0.9"] = CustomExponentialDecayLR(1e-3, 1e-5, 1000, 0.9)(
        tf.convert_to_tensor(steps)).numpy()
    counts = {}
    for n, p, cap in ((90, 0.7, 100), (90, 0.7, 10), (50, 0.4, 10), (3, 0.4, 10), (1, 0.4, 10), (10, 0.2, 10)):
        items = np.array([f"x{j}" for j in range(n)], dtype=object)
        masked, labels = random_item_mask(items, p, cap)
        assert (masked == "[MASK]").sum() == len(labels)
        counts[f"{n}|{p}|{cap}"] = int(len(labels))
    out["n_masked"] = np.array(json.dumps(counts))
    out.update(_clipped_losses(tf))
    return out


def _clipped_losses(tf):
    """ClozeMaskedLoss / MaskedLoss on PROBABILITIES where TF 2.3's clip to [1e-7, 1 - 1e-7] is
    active (peaked rows: the label's probability far below 1e-7, or above 1 - 1e-7) - the regime
    where K.sparse_categorical_crossentropy's clip -> log -> softmax-CE differs from -log p_t
     - and MaskedLoss(K.binary_crossentropy) at saturated sigmoids."""
    from clickstream_transformer.constants import LABEL_PAD
    from clickstream_transformer.losses import MaskedLoss
    from source.utils import ClozeMaskedLoss
    rng = np.random.default_rng(SEED + 9)
    B, M, V = 5, 4, 23
    z = rng.normal(size=(B, M, V)) * np.array([1.0, 8.0, 30.0, 60.0])[None, :, None]
    p = np.exp(z - z.max(-1, keepdims=True))
    p = (p / p.sum(-1, keepdims=True)).astype(np.float32)
    y = rng.integers(0, V, size=(B, M)).astype(np.float32)
    y[0, 2:] = LABEL_PAD
    y[3, :] = LABEL_PAD
    y[1, 3] = float(np.argmax(p[1, 3]))          # a label whose probability is clipped from above
    out = {"clip:probs": p, "clip:labels": y}
    out["clip:cloze_loss"] = ClozeMaskedLoss(tf.keras.backend.sparse_categorical_crossentropy,
                                             label_pad=LABEL_PAD)(y, p).numpy()
    flat_y, flat_p = y.reshape(-1), p.reshape(-1, V)
    out["clip:masked_scc_loss"] = MaskedLoss(tf.keras.backend.sparse_categorical_crossentropy)(
        flat_y, flat_p).numpy()
    q = (1.0 / (1.0 + np.exp(-rng.normal(size=(B, M)) * 25.0))).astype(np.float32)   # many exact 0 / 1
    t = rng.integers(0, 2, size=(B, M)).astype(np.float32)
    t[2, 1:] = LABEL_PAD
    out["clip:sigmoid_probs"], out["clip:binary_labels"] = q, t
    out["clip:masked_bce_loss"] = MaskedLoss(tf.keras.backend.binary_crossentropy)(t, q).numpy()
    out["clip:masked_bce_loss_pw"] = MaskedLoss(tf.keras.backend.binary_crossentropy, pos_weight=4.0)(t, q).numpy()
    # an empty batch: the reference returns 0.0 instead of 0 / 0 (losses.py:89-91)
    out["clip:empty_loss"] = np.float32(MaskedLoss(tf.keras.backend.binary_crossentropy)(
        np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)))
    return out


CASES = {"cloze": case_cloze, "segment": case_segment, "multilabel": case_multilabel,
         "misc": case_schedules_and_masking}


def child(case, path):
    tf = _import_reference()
    with tempfile.TemporaryDirectory() as tmp:
        out = CASES[case](tf, tmp)
    np.savez_compressed(path, **out)


def main():
    if len(sys.argv) == 4 and sys.argv[1] == "--child":
        return child(sys.argv[2], sys.argv[3])
    for case in CASES:
        for mode in ("f32", "f64"):
            path = os.path.join(HERE, f"reference_{case}_{mode}.npz")
            env = dict(os.environ, TFSHIM_FLOAT64="1" if mode == "f64" else "0")
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", case, path],
                           check=True, env=env)
            print("wrote", os.path.relpath(path, ROOT), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
