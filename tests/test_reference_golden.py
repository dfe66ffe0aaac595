"""The oracle and the host logic against the REFERENCE'S OWN SOURCE run in the build container.

tests/golden/reference_*.npz were written by tests/golden/make_reference_golden.py: the
reference's unmodified modules (ClickstreamTransformer, Transformer, the three heads, MaskedLoss,
ClozeMaskedLoss / NDCG / Recall, the metrics, create_cloze_dataset and its masking functions, the
learning-rate schedules) executed on tests/golden/tf_shim - float32 as TensorFlow would hold it
and the same graph in float64.  Here:

  * the NumPy oracle in float64 must reproduce the float64 run to rounding (RTOL64) - forward
    probabilities, loss, every gradient tensor, metrics - and in float32 the float32 run to
    summation-order error (RTOL32);
  * the reference's masking pipeline (keyed permutation in place of tf.random.shuffle) must give
    the ids / labels of oracle.keyed_cloze_batch, the device batch builder's specification;
  * the checkpoint-key map of bert4clickpath_b200/weights.py must list exactly the variables of
    the reference's object graph (row N4), and the host-side learning-rate schedules must return
    the reference's float32 values.
"""
import json
import os

import numpy as np
import pytest

from oracle import clickpath_oracle as O
from bert4clickpath_b200 import weights as W

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
RTOL64 = 1e-12      # float64 oracle vs the float64 run of the reference graph (measured 2.3e-15)
RTOL32 = 2e-5       # float32 oracle vs the float32 run: summation order only (measured 9.4e-7)


def load(case, mode):
    return np.load(os.path.join(G, f"reference_{case}_{mode}.npz"))


def oracle_params(d, features, dtype):
    """Golden {checkpoint key: array} -> the oracle's parameter names, through the product's own
    import map (so the map is exercised on keys that come from the reference's object graph)."""
    variables = {k[len("param:"):]: d[k] for k in d.files if k.startswith("param:")}
    P = W.to_reference_layout(W.import_reference_variables(variables, features=features))
    assert len(P) == len(variables), (sorted(P), sorted(variables))
    return {k: np.asarray(v, dtype=np.float32).astype(dtype) for k, v in P.items()}, variables


def golden_grads(d, tag, features):
    variables = {k[len(tag + "grad:"):]: d[k] for k in d.files if k.startswith(tag + "grad:")}
    return W.to_reference_layout(W.import_reference_variables(variables, features=features, dtype=None))


def rel(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-300))


def assert_grads(Gd, want, tol):
    assert set(Gd) == set(want)
    worst = max(((k, rel(Gd[k].reshape(want[k].shape), want[k])) for k in want
                 if not k.endswith(".bk")), key=lambda kv: kv[1])
    assert worst[1] < tol, worst
    for k in want:
        if k.endswith(".bk"):     # identically zero in exact arithmetic: measured against bq
            assert np.abs(Gd[k] - want[k]).max() < tol * np.abs(want[k[:-2] + "bq"]).max()


# ------------------------------------------------------------------------------------- cloze
@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_cloze_model_loss_and_gradients_match_the_reference_run(mode):
    d = load("cloze", mode)
    dt, tol = (np.float64, RTOL64) if mode == "f64" else (np.float32, RTOL32)
    cfg = json.loads(str(d["config"]))
    vocab = [str(v) for v in d["vocab"]]
    P, variables = oracle_params(d, ["items"], dt)
    pe = O.positional_encoding(10000, cfg["d"])
    for tag in ("infer:train0:", "infer:train1:", "infer:eval0:", "infer:eval1:", "train:train0:"):
        batch = tag.split(":")[1]
        tokens = d[batch + ":asin"]
        ids = O.chain_sequences([O.lookup_ids(tokens, vocab)])
        labels = d[batch + ":labels"]
        masks = None
        if tag.startswith("train:"):
            masks = {"in": d[tag + "dropout:in"].astype(dt)}
            for l in range(cfg["layers"]):
                for j in (1, 2):
                    masks[(l, j)] = d[tag + f"dropout:{l}.{j}"].astype(dt)
        loss, Gd, ex = O.cloze_train_step([ids], labels, P, cfg["layers"], cfg["heads"], pe, dt,
                                          masks=masks, embed_dtype=dt)
        assert rel(loss, d[tag + "loss"]) < tol, (tag, loss, d[tag + "loss"])
        # probabilities of every (row, slot) of the padded head input, pads included
        z = ex["logits"].astype(np.float64)
        p = np.exp(z - z.max(-1, keepdims=True))
        p /= p.sum(-1, keepdims=True)
        want_p = d[tag + "probs"]
        assert p.reshape(want_p.shape).shape == want_p.shape
        assert rel(p.reshape(want_p.shape), want_p) < tol
        assert_grads(Gd, golden_grads(d, tag, ["items"]), tol)


def test_reference_ids_are_the_lookup_of_the_chained_strings():
    d = load("cloze", "f32")
    vocab = [str(v) for v in d["vocab"]]
    ids = O.chain_sequences([O.lookup_ids(d["train0:asin"], vocab)])
    assert np.array_equal(ids, d["train0:ids"])
    assert (d["train1:asin"] == "UNSEEN5").sum() == 1          # the out-of-vocabulary item
    oov = O.lookup_ids(d["train1:asin"], vocab).max()
    assert oov == O.NUM_RESERVED_TOKENS + len(vocab)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_reference_masking_pipeline_equals_the_keyed_batch_builder_spec(mode):
    """create_cloze_dataset / cloze_data_prep / random_item_mask / mask_items of the reference
    (input_pipeline.py:21-133, :198-214), fed the keyed permutation, against
    oracle.keyed_cloze_batch - the specification the device builder is bit-exact to."""
    d = load("cloze", "f32")
    vocab = [str(v) for v in d["vocab"]]
    V, B, seed = len(vocab), int(d["batch_size"]), int(d["seed"])
    lens = d["session_lengths"]
    flat = d["sessions_flat"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    sessions = [np.minimum(flat[offs[i]:offs[i + 1]], V) + O.NUM_RESERVED_TOKENS for i in range(len(lens))]
    for i in range(2):
        idx = list(range(i * B, (i + 1) * B))
        ids, labels, n = O.keyed_cloze_batch(sessions, idx, mode, seed, 0.4, 10)
        want_ids = O.chain_sequences([O.lookup_ids(d[f"{mode}{i}:asin"], vocab)])
        want_labels = d[f"{mode}{i}:labels"]
        assert np.array_equal(ids, want_ids)
        assert n == (want_labels >= 0).sum()
        if want_labels.shape[1] == 0:      # a batch without a single mask: padded_batch gives width 0
            assert n == 0
        else:
            assert np.array_equal(labels, want_labels)


def test_mask_count_rule_is_the_float32_product():
    counts = json.loads(str(load("misc", "f32")["n_masked"]))
    assert counts["90|0.7|100"] == 63          # 62 in float64
    for key, want in counts.items():
        n, p, cap = key.split("|")
        assert O.n_masked_for(int(n), float(p), int(cap)) == want, key


@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_cloze_metrics_match_the_reference_run(mode):
    d = load("cloze", mode)
    for batch in ("train0", "train1", "eval0", "eval1"):
        tag = f"infer:{batch}:"
        labels, probs = d[batch + ":labels"], d[tag + "probs"].astype(np.float32)
        for k in (1, 5, 10):
            s, n = O.cloze_ndcg_update(labels, probs, k)
            assert n == d[tag + f"n_examples@{k}"]
            assert abs(s - d[tag + f"ndcg_sum@{k}"]) < 1e-5 * max(n, 1)
            assert abs(s / n - d[tag + f"ndcg@{k}"]) < 1e-5
            h, n2 = O.cloze_recall_update(labels, probs, k)
            assert n2 == n and abs(h / n - d[tag + f"recall@{k}"]) < 1e-6


# --------------------------------------------------------------------- segment / multilabel
@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_segment_binary_model_matches_the_reference_run(mode):
    d = load("segment", mode)
    dt, tol = (np.float64, RTOL64) if mode == "f64" else (np.float32, RTOL32)
    cfg = json.loads(str(d["config"]))
    feats = ["items", "events"]
    P, _ = oracle_params(d, feats, dt)
    ivocab, evocab = [str(v) for v in d["item_vocab"]], [str(v) for v in d["event_vocab"]]
    ids_items = O.chain_sequences([O.lookup_ids(d["feature:s_items"], ivocab), O.lookup_ids(d["feature:b_items"], ivocab)])
    ids_events = O.chain_sequences([O.lookup_ids(d["feature:s_events"], evocab), O.lookup_ids(d["feature:b_events"], evocab)])
    assert np.array_equal(ids_items, d["ids:items"]) and np.array_equal(ids_events, d["ids:events"])
    starts, ends = O.segment_bounds(ids_items[0])
    assert np.array_equal(starts, d["segment_starts"]) and np.array_equal(ends, d["segment_ends"])
    pe = O.positional_encoding(10000, cfg["d_items"] + cfg["d_events"])
    for tag, pw in (("pw3:", 3.0), ("pw_none:", None)):
        loss, Gd, ex = O.segment_binary_train_step(
            [ids_items, ids_events], d["labels"], P, cfg["layers"], cfg["heads"], pe, cfg["segment"],
            pos_weight=pw, dtype=dt, embed_dtype=dt)
        assert rel(loss, d[tag + "loss"]) < tol
        assert rel(ex["probs"], d[tag + "probs"]) < tol
        assert_grads(Gd, golden_grads(d, tag, feats), tol)
    # PositiveRate / PredictedPositives / F1Score (metrics.py) from the oracle's counters
    c = O.binary_metric_counts(d["labels"], d["pw3:probs"].astype(np.float32))
    assert abs(c[1] / c[0] - d["metric:positive_rate"]) < 1e-6
    assert abs(c[2] / c[0] - d["metric:pred_positives"]) < 1e-6
    assert abs(2 * c[3] / (c[4] + c[5]) - d["metric:f1"]) < 1e-6


@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_multilabel_model_matches_the_reference_run(mode):
    d = load("multilabel", mode)
    dt, tol = (np.float64, RTOL64) if mode == "f64" else (np.float32, RTOL32)
    cfg = json.loads(str(d["config"]))
    feats = ["items", "events"]
    P, _ = oracle_params(d, feats, dt)
    pe = O.positional_encoding(10000, cfg["d_items"] + cfg["d_events"])
    loss, Gd, ex = O.segment_binary_train_step(
        [d["ids:items"], d["ids:events"]], d["labels"], P, cfg["layers"], cfg["heads"], pe,
        cfg["segment"], pos_weight=cfg["pos_weight"], dtype=dt, embed_dtype=dt, head_kind="multilabel")
    assert rel(loss, d["loss"]) < tol
    assert rel(ex["probs"], d["probs"]) < tol
    assert_grads(Gd, golden_grads(d, "", feats), tol)


# ----------------------------------------------------------------------- N4 and host logic
@pytest.mark.parametrize("case,features", [("cloze", ["items"]), ("segment", ["items", "events"]),
                                           ("multilabel", ["items", "events"])])
def test_checkpoint_keys_are_the_reference_object_graph(case, features):
    """export_reference_variables must name every variable exactly as the walk over the
    reference's own model object does (attribute path + /.ATTRIBUTES/VARIABLE_VALUE), with the
    reference's shapes."""
    d = load(case, "f32")
    variables = {k[len("param:"):]: d[k] for k in d.files if k.startswith("param:")}
    store = W.import_reference_variables(variables, features=features)
    back = W.export_reference_variables(store, features=features)
    assert set(back) == set(variables)
    for k, v in variables.items():
        assert back[k].shape == v.shape and np.array_equal(back[k], v.astype(np.float32)), k


def test_learning_rate_schedules_return_the_reference_values():
    from bert4clickpath_b200.training_utils import CustomExponentialDecayLR, CustomLRSchedule
    d = load("misc", "f32")
    steps = d["steps"]
    for key, sched in (("custom_lr:d64_w4000_s1", CustomLRSchedule(d_model=64)),
                       ("custom_lr:d20_w100_s2", CustomLRSchedule(d_model=20, warmup_steps=100, scale=2)),
                       ("exp_decay:1e-3_1e-5_1000_0.9", CustomExponentialDecayLR(1e-3, 1e-5, 1000, 0.9))):
        got = np.array([np.float32(sched(float(s))) for s in steps], dtype=np.float32)
        assert np.allclose(got, d[key], rtol=3e-7, atol=0), (key, got, d[key])


def test_host_input_prep_and_lookup_give_the_reference_ids():
    """The drop-in's host side (TransformerInputPrep + StaticVocabularyTable) on the reference's
    string features -> the ids the reference's own prep + tf.lookup tables produced, and the same
    segment bounds."""
    from bert4clickpath_b200.clickstream_transformer import StaticVocabularyTable, TransformerInputPrep
    from bert4clickpath_b200.constants import RESERVED_TOKENS
    d = load("cloze", "f32")
    table = StaticVocabularyTable(list(RESERVED_TOKENS) + [str(v) for v in d["vocab"]])
    raw, starts, ends = TransformerInputPrep({"items": ["asin"]})(features={"asin": d["train0:asin"].astype(object)})
    assert np.array_equal(table.lookup(raw["items"]), d["train0:ids"])
    assert table.size() == O.NUM_RESERVED_TOKENS + len(d["vocab"]) + 1
    d = load("segment", "f32")
    feats = {k[len("feature:"):]: d[k].astype(object) for k in d.files if k.startswith("feature:")}
    prep = TransformerInputPrep({"items": ["s_items", "b_items"], "events": ["s_events", "b_events"]})
    raw, starts, ends = prep(features=feats)
    assert set(raw) == {"items", "events"}
    assert np.array_equal(starts, d["segment_starts"]) and np.array_equal(ends, d["segment_ends"])
    for f, vocab in (("items", d["item_vocab"]), ("events", d["event_vocab"])):
        table = StaticVocabularyTable(list(RESERVED_TOKENS) + [str(v) for v in vocab])
        assert np.array_equal(table.lookup(raw[f]), d["ids:" + f])


@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_losses_where_the_tf_clip_is_active_match_the_reference_run(mode):
    """K.sparse_categorical_crossentropy / K.binary_crossentropy on PROBABILITIES with the label's
    probability far outside [1e-7, 1 - 1e-7]: the oracle's 'exact_tf23' forms
    (clip -> log -> softmax-CE; clip -> log(p + eps)) against ClozeMaskedLoss / MaskedLoss of the
    reference, and the logits-mode value must DIFFER there (that is the regime the two part)."""
    d = load("misc", mode)
    dt = np.float64 if mode == "f64" else np.float32
    tol = 1e-12 if mode == "f64" else 2e-6
    p, y = d["clip:probs"].astype(dt), d["clip:labels"]
    assert (p[np.arange(5)[:, None], np.arange(4)[None, :], np.maximum(y, 0).astype(int)] < 1e-7).any()
    got = O.cloze_masked_loss(y, p, ce_mode="exact_tf23")
    assert rel(got, d["clip:cloze_loss"]) < tol
    assert rel(O.cloze_masked_loss(y, p, ce_mode="logits"), d["clip:cloze_loss"]) > 1e-2
    flat = O.masked_loss(y.reshape(-1).astype(dt), p.reshape(-1, p.shape[-1]),
                         lambda yy, pp: O.sparse_categorical_crossentropy_probs(yy, pp, "exact_tf23"))
    assert rel(flat, d["clip:masked_scc_loss"]) < tol
    q, t = d["clip:sigmoid_probs"].astype(dt), d["clip:binary_labels"].astype(dt)
    assert ((q == 0) | (q == 1)).any()
    assert rel(O.masked_loss(t, q, O.binary_crossentropy_probs), d["clip:masked_bce_loss"]) < tol
    assert rel(O.masked_loss(t, q, O.binary_crossentropy_probs, pos_weight=4.0), d["clip:masked_bce_loss_pw"]) < tol
    assert float(d["clip:empty_loss"]) == 0.0
    assert float(O.masked_loss(np.zeros((0, 3), dt), np.zeros((0, 3), dt), O.binary_crossentropy_probs)) == 0.0
