"""Model-level parity: the reference-named classes on libb4cp against the float64 oracle on the
same weights and inputs (loss, every gradient, probabilities, metrics, Adam trajectory)."""
import numpy as np
import pytest
import torch

from oracle import clickpath_oracle as O
from tests.test_oracle import make_tiny_problem

pytestmark = pytest.mark.gpu

# bf16 tensor-core operands with fp32 accumulation: gradients are compared on the max-norm of
# each tensor; 3e-2 covers two encoder layers + a 3-layer head at d_model = 8..64.
BF16_TOL = 3e-2


def build_model(P, dims, L, H, dff, head_dims, V, dropout=0.0, rows2=20):
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.weights import to_store_layout
    feats = ["items", "events"][:len(dims)]
    head = bc.SoftMaxHead(dense_layer_dims=list(head_dims), output_vocab_size=V)
    model = bc.ClickstreamTransformer(
        sequential_input_config={f: [f"seq_{f}"] for f in feats},
        feature_vocabs=dict([("items", V)] + ([("events", rows2 - 11)] if len(dims) > 1 else [])),
        embedding_dims={f: d for f, d in zip(feats, dims)},
        head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
        num_encoder_layers=L, num_attention_heads=H, dropout_rate=dropout, encoder_ff_dim=dff)
    model.store.set_weights(to_store_layout({k: v for k, v in P.items()}))
    return model


def rel_err(got, want, floor=0.0):
    """max-norm error relative to the tensor's own scale (or `floor` for tensors whose true
    gradient is ~0, e.g. the key bias, to which softmax attention is invariant)."""
    return np.abs(got - want).max() / max(np.abs(want).max(), floor, 1e-12)


def grad_floor(G):
    return 1e-2 * max(np.abs(v).max() for v in G.values())


@pytest.mark.parametrize("dims", [(8,), (8, 8)])
def test_cloze_forward_backward_matches_oracle(cuda_lib, dims):
    from bert4clickpath_b200.weights import to_reference_layout
    ids_list, labels, P, L, H, pe, _ = make_tiny_problem(dims=dims, dff=12, head=(16, 8))
    V = P["head.out.w"].shape[1]
    model = build_model(P, dims, L, H, 12, (16, 8), V, rows2=P["emb.1"].shape[0] if len(dims) > 1 else 20)
    B, S = ids_list[0].shape
    dev_ids = [torch.from_numpy(i.astype(np.int32)).cuda().view(-1) for i in ids_list]
    lab = torch.from_numpy(labels.astype(np.float32)).cuda()
    n_masked = int((labels >= 0).sum())
    stats = model.cloze_forward_backward(dev_ids, lab, B, S, n_masked=n_masked, training=False)
    torch.cuda.synchronize()
    loss, G, ex = O.cloze_train_step(ids_list, labels, P, L, H, pe, np.float64)
    s = stats.cpu().numpy()
    assert s[1] == ex["n_valid"]
    assert abs(s[0] / s[1] - loss) < 2e-2 * abs(loss)
    got = to_reference_layout(model.store.get_grads())
    for k in sorted(G):
        assert rel_err(got[k], G[k], grad_floor(G)) < BF16_TOL, (k, rel_err(got[k], G[k], grad_floor(G)))
    # PAD rows of the item table get exactly zero gradient (dead compute)
    assert not np.abs(got["emb.0"][0]).any() or (ids_list[0] == 0).any()


def test_materialize_matches_reference_layout(cuda_lib):
    ids_list, labels, P, L, H, pe, _ = make_tiny_problem(dims=(8,), dff=12, head=(16, 8))
    V = P["head.out.w"].shape[1]
    model = build_model(P, (8,), L, H, 12, (16, 8), V)
    B, S = ids_list[0].shape
    out = model.call({"seq_items": ids_list[0][:, 2:-1]}, training=False)
    probs = out.materialize().cpu().numpy()
    x, _ = O.encoder_fwd(ids_list, P, L, H, pe, np.float64)
    sel, _ = O.select_masked(ids_list[0], x)
    want, _, _ = O.softmax_head_fwd(sel, O.head_layers(P), P["head.out.w"], P["head.out.b"])
    assert probs.shape == want.shape
    np.testing.assert_allclose(probs, want, rtol=5e-2, atol=1e-4)
    np.testing.assert_allclose(probs.sum(-1), 1.0, atol=1e-4)
    # loss / metrics through the reference-named classes, lazy and materialised
    import bert4clickpath_b200 as bc
    loss_fn = bc.ClozeMaskedLoss(bc.sparse_categorical_crossentropy, label_pad=bc.LABEL_PAD)
    l_lazy = loss_fn(labels, out)
    l_mat = loss_fn(labels, out.materialize())
    want_loss = float(O.cloze_masked_loss(labels, want.astype(np.float32)))
    assert abs(l_lazy - want_loss) < 2e-2 * want_loss and abs(l_mat - l_lazy) < 1e-4 * want_loss
    for k in (1, 5):
        m_lazy, m_mat = bc.ClozeMaskedNDCG(k), bc.ClozeMaskedNDCG(k)
        m_lazy.update_state(labels, out)
        m_mat.update_state(labels, out.materialize())
        r = bc.ClozeMaskedRecall(k)
        r.update_state(labels, out.materialize())
        gs, gn = O.cloze_ndcg_update(labels, probs, k)   # oracle on the SAME fp32 scores
        rs, rn = O.cloze_recall_update(labels, probs, k)
        assert abs(m_mat.result() - gs / gn) < 1e-6 and abs(r.result() - rs / rn) < 1e-6
        assert abs(m_lazy.result() - m_mat.result()) < 1e-6


def test_dropout_training_step_matches_oracle_with_exported_masks(cuda_lib):
    from bert4clickpath_b200 import ops
    from bert4clickpath_b200.engine import SITE_INPUT, site
    from bert4clickpath_b200.weights import to_reference_layout
    ids_list, labels, P, L, H, pe, _ = make_tiny_problem(dims=(8,), dff=12, head=(16, 8))
    V = P["head.out.w"].shape[1]
    model = build_model(P, (8,), L, H, 12, (16, 8), V, dropout=0.25)
    B, S = ids_list[0].shape
    seed = 77
    mk = lambda st: ops.dropout_mask(B * S * 8, 0.25, seed, st).cpu().numpy().reshape(B, S, 8).astype(np.float64)
    masks = {"in": mk(SITE_INPUT)}
    for l in range(L):
        masks[(l, 1)] = mk(site(l, 1))
        masks[(l, 2)] = mk(site(l, 2))
    dev_ids = [torch.from_numpy(i.astype(np.int32)).cuda().view(-1) for i in ids_list]
    lab = torch.from_numpy(labels.astype(np.float32)).cuda()
    stats = model.cloze_forward_backward(dev_ids, lab, B, S, n_masked=int((labels >= 0).sum()),
                                         training=True, seed=seed)
    loss, G, ex = O.cloze_train_step(ids_list, labels, P, L, H, pe, np.float64, masks)
    s = stats.cpu().numpy()
    assert abs(s[0] / s[1] - loss) < 2e-2 * abs(loss)
    got = to_reference_layout(model.store.get_grads())
    for k in sorted(G):
        assert rel_err(got[k], G[k], grad_floor(G)) < BF16_TOL, (k, rel_err(got[k], G[k], grad_floor(G)))


def test_three_adam_steps_track_oracle(cuda_lib):
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.weights import to_reference_layout
    ids_list, labels, P, L, H, pe, _ = make_tiny_problem(dims=(8,), dff=12, head=(16, 8))
    V = P["head.out.w"].shape[1]
    model = build_model(P, (8,), L, H, 12, (16, 8), V)
    model.compile(optimizer=bc.Adam(1e-3, 0.9, 0.999, 1e-9))
    Pm = {k: v.copy() for k, v in P.items()}
    M_ = {k: np.zeros_like(v) for k, v in P.items()}
    V_ = {k: np.zeros_like(v) for k, v in P.items()}
    losses = []
    for t in range(1, 4):
        logs = model.train_step(({"seq_items": ids_list[0][:, 2:-1]}, labels))
        loss, G, _ = O.cloze_train_step(ids_list, labels, Pm, L, H, pe, np.float64)
        for k in Pm:
            Pm[k], M_[k], V_[k] = O.adam_step(Pm[k], G[k], M_[k], V_[k], t)
        losses.append((logs["loss"], loss))
    for got, want in losses:
        assert abs(got - want) < 2e-2 * want
    W = to_reference_layout(model.store.get_weights())
    for k in Pm:
        # every weight moved by ~lr per step in the same direction as the oracle's
        moved = W[k] - P[k]
        want = Pm[k] - P[k]
        big = np.abs(want) > 2e-3
        if big.any():
            assert (np.sign(moved[big]) == np.sign(want[big])).mean() > 0.97, k
