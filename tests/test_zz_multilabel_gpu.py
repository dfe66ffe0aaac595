"""MultiLabel_MultiClass_classification (head.py:50-69) trained in segment mode: the item-wise
sigmoid-BCE gradient kernel and the whole head backward against the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import clickpath_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols,pw", [(5, 13, None), (33, 200, 2.5), (1, 1, 1.5)])
def test_sigmoid_bce_dz_kernel_matches_oracle(cuda_lib, rows, cols, pw):
    from bert4clickpath_b200 import ops
    rng = np.random.default_rng(rows)
    p = rng.random((rows, cols)).astype(np.float32)
    p.flat[0] = 1.0          # clipped: no gradient
    y = (rng.random((rows, cols)) < 0.3).astype(np.float32)
    y[rng.random((rows, cols)) < 0.2] = -1.0
    y.flat[0] = 1.0
    pd, yd = torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda()
    stats = ops.masked_bce(yd.view(-1), pd.view(-1), -1.0, pw)
    dz32 = torch.full((rows, cols), 7.0, dtype=torch.float32, device="cuda")
    dzb = torch.full((rows, ops.ld8(cols)), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.sigmoid_bce_dz(yd, pd, rows, cols, -1.0, pw, stats, dz_f32=dz32, dz_bf16=dzb)
    # oracle: d(loss)/dz with z = logit(p) on an identity "head" (no MLP, w = I)
    mask = y != -1.0
    n = mask.sum()
    eps = 1e-7
    p64, yy = p.astype(np.float64), np.where(mask, y, 0.0).astype(np.float64)
    w = np.where(yy == 1, pw, 1.0) if pw is not None else 1.0
    norm = (pw + 1) / 2 if pw is not None else 1.0
    inside = (p64 > eps) & (p64 < 1 - eps)
    want = np.where(mask, -(yy / (p64 + eps) - (1 - yy) / (1 - p64 + eps)) * w * inside
                    * p64 * (1 - p64), 0.0) / max(n, 1) / norm
    got = dz32.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=1e-9)
    assert got.flat[0] == 0.0 and (got[~mask] == 0).all()
    gb = dzb.float().cpu().numpy()
    np.testing.assert_allclose(gb[:, :cols], want, rtol=1e-2, atol=1e-9)
    assert (gb[:, cols:] == 0).all()                      # pad columns are zeroed
    s = stats.cpu().numpy()
    assert s[1] == n
    want_loss = O.masked_loss(y.astype(np.float64), p64, O.binary_crossentropy_probs, pos_weight=pw)
    assert abs(s[0] / s[1] / norm - want_loss) < 1e-4 * max(1.0, abs(want_loss))


@pytest.mark.parametrize("dims,pw", [([32, 16], 2.0), ([], None)])
def test_multilabel_head_training_matches_oracle(cuda_lib, dims, pw):
    import bert4clickpath_b200 as bc
    rng = np.random.default_rng(3 + len(dims))
    items_vocab = [f"it{j}" for j in range(60)]
    B, L1, V = 29, 9, 45
    a = rng.choice(items_vocab, size=(B, L1)).astype(object)
    for b in range(B):
        a[b, rng.integers(1, L1 + 1):] = "[PAD]"
    feats = {"s_items": a}
    head = bc.MultiLabel_MultiClass_classification(dense_layer_dims=dims, output_vocab_size=V)
    model = bc.ClickstreamTransformer(
        sequential_input_config={"items": ["s_items"]}, feature_vocabs={"items": items_vocab},
        embedding_dims={"items": 32}, head_unit=head, segment_to_head=0,
        num_encoder_layers=1, num_attention_heads=4, dropout_rate=0.0, seed=5)
    ids_list, Bq, S, starts, ends = model.prepare_inputs(feats)
    s0, s1 = int(starts[0]), int(ends[0])
    assert s1 - s0 == 1                                    # segment 0 = the [CLS] slot
    y = (rng.random((B, V)) < 0.15).astype(np.float32)
    y[rng.random((B, V)) < 0.1] = -1.0
    yd = torch.from_numpy(y).cuda()
    stats = model.multilabel_forward_backward(ids_list, yd, B, S, (starts, ends), pos_weight=pw,
                                              training=False).cpu().numpy()
    grads, Wts = model.store.get_grads(), model.store.get_weights()
    x = model._encode(ids_list, B, S, False, 0).cpu().numpy().reshape(B, S, 32)
    layers = [(Wts[f"head.{i}.w"].astype(np.float64), Wts[f"head.{i}.b"].astype(np.float64))
              for i in range(len(dims))]
    o_loss, _, o_lg, o_dWo, o_dbo = O.multilabel_head_loss_and_grads(
        x[:, s0:s1].astype(np.float64), layers, Wts["head.out.w"].astype(np.float64),
        Wts["head.out.b"].astype(np.float64), y, pos_weight=pw)
    norm = (pw + 1) / 2 if pw is not None else 1.0
    got_loss = stats[0] / stats[1] / norm
    assert stats[1] == (y != -1).sum()
    assert abs(got_loss - o_loss) < 2e-2 * abs(o_loss)
    want = {"head.out.w": o_dWo, "head.out.b": o_dbo}
    for i, (dw, db_) in enumerate(o_lg):
        want[f"head.{i}.w"], want[f"head.{i}.b"] = dw, db_
    for k, w in want.items():
        g = grads[k].astype(np.float64).reshape(w.shape)
        assert np.linalg.norm(g - w) / max(np.linalg.norm(w), 1e-12) < 6e-2, k   # bf16 operands
    enc = [k for k in grads if not k.startswith("head.")]
    assert all(np.isfinite(grads[k]).all() for k in enc) and any(np.abs(grads[k]).max() > 0 for k in enc)
    # the forward the reference-shaped call returns is what the loss was computed on
    probs = model.call(feats, training=False)
    assert tuple(probs.shape) == (B, V)
    assert torch.allclose(probs, model._last_probs, rtol=0, atol=1e-6)
    # Keras-style training through train_step: the loss goes down
    model.compile(optimizer=bc.Adam(1e-2), loss=bc.MaskedLoss(bc.binary_crossentropy, pos_weight=pw))
    losses = [model.train_step((feats, y))["loss"] for _ in range(12)]
    assert abs(losses[0] - o_loss) < 2e-2 * abs(o_loss) and losses[-1] < 0.9 * losses[0]
    # a longer segment cannot be squeezed
    feats2 = dict(feats)
    bad = bc.ClickstreamTransformer(
        sequential_input_config={"items": ["s_items"]}, feature_vocabs={"items": items_vocab},
        embedding_dims={"items": 32},
        head_unit=bc.MultiLabel_MultiClass_classification(dense_layer_dims=[], output_vocab_size=V),
        segment_to_head=1, num_encoder_layers=1, num_attention_heads=4, dropout_rate=0.0)
    i2, _, S2, st2, en2 = bad.prepare_inputs(feats2)
    with pytest.raises(ValueError):
        bad.multilabel_forward_backward(i2, yd, B, S2, (st2, en2))
