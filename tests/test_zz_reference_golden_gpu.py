"""The CUDA path against the REFERENCE'S OWN SOURCE - no oracle in between.

tests/golden/reference_*_f64.npz hold what MiladShahidi/BERT4ClickPath's unmodified modules
computed in this repository's build container (tests/golden/make_reference_golden.py: the
reference's ClickstreamTransformer / heads / losses / metrics / input pipeline executed on a small
eager implementation of the TensorFlow calls they make, in float64): string batches produced by
the reference's create_cloze_dataset, the weights of the reference's model object keyed by their
checkpoint paths, probabilities, loss, every gradient, NDCG / recall.

Here the drop-in model is built from the same constructor arguments, loads those weights through
`load_weights` (the checkpoint-key map, row N4), is fed the same STRING features and labels, and
must reproduce

  precision="fp32"  loss, probabilities and every gradient within 1e-3 (north_star's bar; both
                    norms, relative to the tensor's own scale)
  precision="bf16"  the stated fast-path tolerances (these toy models have 8- and 16-wide layers:
                    one ReLU unit flipped by bf16 operand rounding is a percent of a tensor)

and the reference's metric values.  Measured errors are written to gpurun_out/parity_reference_*.json.
"""
import json
import os

import numpy as np
import pytest
import torch

from bert4clickpath_b200 import weights as W

from .test_zz_parity_configs_gpu import FP32_TOL, report, tensor_errors

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BF16_LOSS_TOL = 1e-2       # measured worst 2.6e-3 (segment, pos_weight None)
BF16_FRO_TOL = 0.25        # measured worst 0.11 (cloze train1, head.0.w: 24 units, one gate)
BF16_PROB_TOL = 2e-2       # measured worst 3.9e-3


def load(case):
    return np.load(os.path.join(G, f"reference_{case}_f64.npz"))


def load_reference_weights(model, d, tmp_path):
    """The golden variables -> an .npz keyed like a reference checkpoint -> model.load_weights."""
    path = os.path.join(str(tmp_path), "reference_variables.npz")
    np.savez(path, **{k[len("param:"):]: d[k].astype(np.float32) for k in d.files if k.startswith("param:")})
    model.load_weights(path)


def reference_grads(d, tag, features):
    variables = {k[len(tag + "grad:"):]: d[k] for k in d.files if k.startswith(tag + "grad:")}
    return W.to_reference_layout(W.import_reference_variables(variables, features=features, dtype=None))


def rel_pair(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return (float(np.abs(got - want).max() / np.abs(want).max()),
            float(np.linalg.norm(got - want) / np.linalg.norm(want)))


def judge(name, precision, loss_err, errs, prob_err):
    loss_err = float(loss_err)
    worst_max = max(errs.items(), key=lambda kv: kv[1][0])
    worst_fro = max(errs.items(), key=lambda kv: kv[1][1])
    report(f"reference_{name}_{precision}", dict(
        loss_rel_err=loss_err, probs=dict(max=prob_err[0], fro=prob_err[1]), worst_max=worst_max,
        worst_fro=worst_fro, per_tensor={k: dict(max=v[0], fro=v[1]) for k, v in errs.items()}))
    if precision == "fp32":
        assert loss_err < FP32_TOL, loss_err
        assert max(prob_err) < FP32_TOL, prob_err
        assert worst_max[1][0] < FP32_TOL, worst_max
        assert worst_fro[1][1] < FP32_TOL, worst_fro
    else:
        assert loss_err < BF16_LOSS_TOL, loss_err
        assert prob_err[0] < BF16_PROB_TOL, prob_err
        assert worst_fro[1][1] < BF16_FRO_TOL, worst_fro


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cloze_step_on_reference_batches_matches_the_reference_run(cuda_lib, tmp_path, precision):
    import bert4clickpath_b200 as bc
    d = load("cloze")
    cfg = json.loads(str(d["config"]))
    vocab = [str(v) for v in d["vocab"]]
    head = bc.SoftMaxHead(dense_layer_dims=cfg["head"], output_vocab_size=cfg["V"])
    model = bc.ClickstreamTransformer(
        sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": vocab},
        embedding_dims={"items": cfg["d"]}, head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
        num_encoder_layers=cfg["layers"], num_attention_heads=cfg["heads"],
        dropout_rate=cfg["dropout"], precision=precision)
    load_reference_weights(model, d, tmp_path)
    for batch in ("train0", "train1", "eval0", "eval1"):
        tag = f"infer:{batch}:"
        feats = {"asin": d[batch + ":asin"].astype(object)}
        labels = d[batch + ":labels"].astype(np.float32)
        ids_list, B, S, _, _ = model.prepare_inputs(feats)
        if batch == "train0":
            assert np.array_equal(ids_list[0].view(B, S).cpu().numpy(), d["train0:ids"])
        n = int((labels >= 0).sum())
        stats = model.cloze_forward_backward(ids_list, torch.from_numpy(labels).cuda(), B, S,
                                             n_masked=n, training=False).cpu().numpy()
        assert stats[1] == n
        want_loss = float(d[tag + "loss"])
        got = W.to_reference_layout(model.store.get_grads())
        want = reference_grads(d, tag, ["items"])
        errs = tensor_errors(got, want, want)
        out = model.call(feats, training=False)
        probs = out.materialize().cpu().numpy()
        want_p = d[tag + "probs"]
        assert probs.shape == want_p.shape
        # every (row, slot) of the padded head input, including the slots past an example's own
        # masks, where the reference feeds zero vectors through the head (to_tensor's padding)
        judge(f"cloze_{batch}", precision, abs(stats[0] / stats[1] - want_loss) / want_loss, errs,
              rel_pair(probs, want_p))
        for k in (1, 5, 10):
            nd, rc = bc.ClozeMaskedNDCG(k=k), bc.ClozeMaskedRecall(k=k)
            nd.update_state(labels, out)
            rc.update_state(labels, out)
            if precision == "fp32":     # ranks are discrete: a near-tie may swap under bf16
                assert abs(float(nd.result()) - float(d[tag + f"ndcg@{k}"])) < 1e-5
                assert abs(float(rc.result()) - float(d[tag + f"recall@{k}"])) < 1e-6


def _two_feature_model(bc, d, cfg, head, precision):
    return bc.ClickstreamTransformer(
        sequential_input_config={"items": ["s_items", "b_items"], "events": ["s_events", "b_events"]},
        feature_vocabs={"items": [str(v) for v in d["item_vocab"]],
                        "events": [str(v) for v in d["event_vocab"]]},
        embedding_dims={"items": cfg["d_items"], "events": cfg["d_events"]}, head_unit=head,
        segment_to_head=cfg["segment"], num_encoder_layers=cfg["layers"],
        num_attention_heads=cfg["heads"], dropout_rate=0.0, precision=precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_segment_binary_step_matches_the_reference_run(cuda_lib, tmp_path, precision):
    import bert4clickpath_b200 as bc
    d = load("segment")
    cfg = json.loads(str(d["config"]))
    model = _two_feature_model(bc, d, cfg, bc.BinaryClassificationHead(dense_layer_dims=cfg["head"]), precision)
    load_reference_weights(model, d, tmp_path)
    feats = {k[len("feature:"):]: d[k].astype(object) for k in d.files if k.startswith("feature:")}
    ids_list, B, S, starts, ends = model.prepare_inputs(feats)
    assert np.array_equal(ids_list[0].view(B, S).cpu().numpy(), d["ids:items"])
    assert np.array_equal(ids_list[1].view(B, S).cpu().numpy(), d["ids:events"])
    assert np.array_equal(np.asarray(starts), d["segment_starts"]) and np.array_equal(np.asarray(ends), d["segment_ends"])
    y = torch.from_numpy(d["labels"].astype(np.float32)).cuda()
    for tag, pw in (("pw3:", 3.0), ("pw_none:", None)):
        stats = model.binary_forward_backward(ids_list, y, B, S, (starts, ends), pos_weight=pw,
                                              training=False).cpu().numpy()
        loss = stats[0] / stats[1] / (((pw + 1.0) / 2) if pw is not None else 1.0)
        want_loss = float(d[tag + "loss"])
        want = reference_grads(d, tag, ["items", "events"])
        errs = tensor_errors(W.to_reference_layout(model.store.get_grads()), want, want)
        probs = model._last_probs.cpu().numpy().reshape(d[tag + "probs"].shape)
        judge(f"segment_{tag[:-1]}", precision, abs(loss - want_loss) / want_loss, errs,
              rel_pair(probs, d[tag + "probs"]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_multilabel_step_matches_the_reference_run(cuda_lib, tmp_path, precision):
    import bert4clickpath_b200 as bc
    d = load("multilabel")
    cfg = json.loads(str(d["config"]))
    head = bc.MultiLabel_MultiClass_classification(dense_layer_dims=cfg["head"], output_vocab_size=cfg["classes"])
    model = _two_feature_model(bc, d, cfg, head, precision)
    load_reference_weights(model, d, tmp_path)
    feats = {k[len("feature:"):]: d[k].astype(object) for k in d.files if k.startswith("feature:")}
    ids_list, B, S, starts, ends = model.prepare_inputs(feats)
    assert np.array_equal(ids_list[0].view(B, S).cpu().numpy(), d["ids:items"])
    y = torch.from_numpy(d["labels"].astype(np.float32)).cuda()
    pw = cfg["pos_weight"]
    stats = model.multilabel_forward_backward(ids_list, y, B, S, (starts, ends), pos_weight=pw,
                                              training=False).cpu().numpy()
    loss = stats[0] / stats[1] / ((pw + 1.0) / 2)
    want_loss = float(d["loss"])
    want = reference_grads(d, "", ["items", "events"])
    errs = tensor_errors(W.to_reference_layout(model.store.get_grads()), want, want)
    probs = model.call(feats, training=False).cpu().numpy().reshape(d["probs"].shape)
    judge("multilabel", precision, abs(loss - want_loss) / want_loss, errs, rel_pair(probs, d["probs"]))


def test_materialised_losses_in_the_clip_regime_match_the_reference_run(cuda_lib):
    """ClozeMaskedLoss / MaskedLoss of the drop-in on materialised PROBABILITIES, where TF 2.3's
    clip to [1e-7, 1 - 1e-7] is active (labels with probability far below 1e-7, saturated
    sigmoids): `b4cp_clip_log` + the row CE kernels and `b4cp_masked_bce` against the values the
    reference's own loss classes returned (tests/golden/reference_misc_f32.npz)."""
    import bert4clickpath_b200 as bc
    d = np.load(os.path.join(G, "reference_misc_f32.npz"))
    p, y = d["clip:probs"], d["clip:labels"]
    V = p.shape[-1]
    got = bc.ClozeMaskedLoss(bc.sparse_categorical_crossentropy)(y, p)
    assert abs(got - float(d["clip:cloze_loss"])) < 2e-5 * float(d["clip:cloze_loss"])
    got = bc.MaskedLoss(bc.sparse_categorical_crossentropy)(y.reshape(-1), p.reshape(-1, V))
    assert abs(got - float(d["clip:masked_scc_loss"])) < 2e-5 * float(d["clip:masked_scc_loss"])
    q, t = d["clip:sigmoid_probs"], d["clip:binary_labels"]
    got = bc.MaskedLoss(bc.binary_crossentropy)(t, q)
    assert abs(got - float(d["clip:masked_bce_loss"])) < 2e-5 * float(d["clip:masked_bce_loss"])
    got = bc.MaskedLoss(bc.binary_crossentropy, pos_weight=4.0)(t, q)
    assert abs(got - float(d["clip:masked_bce_loss_pw"])) < 2e-5 * float(d["clip:masked_bce_loss_pw"])
    assert bc.MaskedLoss(bc.binary_crossentropy)(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)) == 0.0
