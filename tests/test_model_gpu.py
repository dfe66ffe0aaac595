"""Model-level parity: the reference-named classes on libb4cp against the float64 oracle on the
same weights and inputs (loss, every gradient, probabilities, metrics, Adam trajectory)."""
import numpy as np
import pytest
import torch

from oracle import clickpath_oracle as O
from oracle.mixed_precision import cloze_train_step_bf16
from tests.test_oracle import make_tiny_problem

pytestmark = pytest.mark.gpu

# Two bars (DESIGN.md "Numerics"):
#  * KERNEL_TOL — CUDA vs the oracle with bf16 rounding applied at the points where the pipeline
#    stores bf16 (oracle/mixed_precision.py): the kernels compute what they claim.  The only
#    differences are fp32-vs-float64 accumulation and rare 1-ulp bf16 rounding flips.
#  * BF16_TOL — CUDA vs the exact float64 oracle: the bf16-operand approximation itself, on the
#    max-norm of each gradient tensor (tiny batches at random init are the worst case because the
#    gradients are sums with heavy cancellation).
#    A 1-ulp bf16 flip upstream can flip a ReLU gate downstream, which changes single gradient
#    entries discontinuously in ANY two implementations; the kernel bar is therefore measured in
#    the Frobenius norm (kernel-by-kernel max-norm parity is in test_kernels_gpu.py).
KERNEL_TOL = 2e-2
BF16_TOL = 6e-2


def build_model(P, dims, L, H, dff, head_dims, V, dropout=0.0, rows2=20, precision="bf16"):
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.weights import to_store_layout
    feats = ["items", "events"][:len(dims)]
    head = bc.SoftMaxHead(dense_layer_dims=list(head_dims), output_vocab_size=V)
    model = bc.ClickstreamTransformer(
        sequential_input_config={f: [f"seq_{f}"] for f in feats},
        feature_vocabs=dict([("items", V)] + ([("events", rows2 - 11)] if len(dims) > 1 else [])),
        embedding_dims={f: d for f, d in zip(feats, dims)},
        head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
        num_encoder_layers=L, num_attention_heads=H, dropout_rate=dropout, encoder_ff_dim=dff,
        precision=precision)
    model.store.set_weights(to_store_layout({k: v for k, v in P.items()}))
    return model


def rel_err(got, want, scale=None):
    """Frobenius-norm error relative to the tensor's OWN norm.  `scale` replaces the denominator
    for the one tensor family whose true gradient is identically zero - the key bias, to which
    softmax attention is invariant - and is the query-bias gradient of the same layer."""
    ref = want if scale is None else scale
    return np.linalg.norm(got - want) / max(np.linalg.norm(ref), 1e-300)


def grad_floor(G):
    """{key-bias name: the same layer's query-bias gradient}: the only floored tensors."""
    return {k: G[k[:-2] + "bq"] for k in G if k.endswith(".bk")}


def all_errs(got, want, G):
    fl = grad_floor(G)
    return {k: rel_err(got[k], want[k], fl.get(k)) for k in sorted(want)}


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
    eloss, EG, _ = cloze_train_step_bf16(ids_list, labels, P, L, H, pe)
    s = stats.cpu().numpy()
    assert s[1] == ex["n_valid"]
    assert abs(s[0] / s[1] - loss) < 2e-2 * abs(loss)
    assert abs(s[0] / s[1] - eloss) < 1e-4 * abs(eloss)
    got = to_reference_layout(model.store.get_grads())
    ek, eb = all_errs(got, EG, G), all_errs(got, G, G)
    assert max(ek.values()) < KERNEL_TOL, max(ek.items(), key=lambda kv: kv[1])
    assert max(eb.values()) < BF16_TOL, max(eb.items(), key=lambda kv: kv[1])
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
    eloss, EG, _ = cloze_train_step_bf16(ids_list, labels, P, L, H, pe, masks)
    s = stats.cpu().numpy()
    assert abs(s[0] / s[1] - loss) < 2e-2 * abs(loss)
    assert abs(s[0] / s[1] - eloss) < 1e-4 * abs(eloss)
    got = to_reference_layout(model.store.get_grads())
    ek, eb = all_errs(got, EG, G), all_errs(got, G, G)
    assert max(ek.values()) < KERNEL_TOL, max(ek.items(), key=lambda kv: kv[1])
    assert max(eb.values()) < BF16_TOL, max(eb.items(), key=lambda kv: kv[1])


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_three_adam_steps_track_oracle(cuda_lib, precision):
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.weights import to_reference_layout
    ids_list, labels, P, L, H, pe, _ = make_tiny_problem(dims=(8,), dff=12, head=(16, 8))
    V = P["head.out.w"].shape[1]
    model = build_model(P, (8,), L, H, 12, (16, 8), V, precision=precision)
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
        assert abs(got - want) < (1e-4 if precision == "fp32" else 2e-2) * want
    W = to_reference_layout(model.store.get_weights())
    for k in Pm:
        moved = W[k] - P[k]
        want = Pm[k] - P[k]
        if precision == "fp32":
            # VALUES: three Keras-Adam updates of the oracle, weight by weight.  Adam divides by
            # sqrt(v): entries whose gradient is itself rounding noise (|g| << the tensor's
            # scale) move by +-lr in a direction no two implementations agree on, so the
            # comparison is over the entries the oracle moved by a full-size step.
            sure = np.abs(want) > 2.5e-3           # 3 steps of lr = 1e-3, consistent sign
            if sure.any():
                np.testing.assert_allclose(moved[sure], want[sure], rtol=2e-3, atol=2e-6, err_msg=k)
            assert np.abs(moved - want).max() < 6.1e-3, k   # nothing moves further than 2 * 3 * lr
        else:
            # bf16 operands: every weight moved by ~lr per step in the oracle's direction
            big = np.abs(want) > 2e-3
            if big.any():
                assert (np.sign(moved[big]) == np.sign(want[big])).mean() > 0.97, k


@pytest.mark.parametrize("cfg", [
    dict(V=300, d=32, L=2, H=2, dff=100, hd=[64, 32], B=16, max_len=20, lengths="beauty", mp=0.4),
    dict(V=1000, d=64, L=2, H=2, dff=100, hd=[128, 64], B=64, max_len=50, lengths="dense", mp=0.15),
    dict(V=500, d=128, L=1, H=4, dff=100, hd=[], B=32, max_len=30, lengths="beauty", mp=0.4),
    # C4-shaped: d_model 256, 4 heads of depth 64, head [] -> V (h = 256: fused TS-form vocabulary path)
    dict(V=2000, d=256, L=2, H=4, dff=100, hd=[], B=8, max_len=40, lengths="beauty", mp=0.15),
])
def test_kernels_match_bf16_emulation_on_c1_like_shapes(cuda_lib, cfg):
    """Ragged (beauty-shaped) and dense sessions, dff=100 (not a multiple of 8), heads with and
    without an MLP: CUDA == oracle-with-bf16-rounding to KERNEL_TOL on every gradient tensor."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.synthetic import make_cloze_batch
    from bert4clickpath_b200.weights import to_reference_layout
    V, d, L, H = cfg["V"], cfg["d"], cfg["L"], cfg["H"]
    head = bc.SoftMaxHead(dense_layer_dims=cfg["hd"], output_vocab_size=V)
    model = bc.ClickstreamTransformer(
        sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": V},
        embedding_dims={"items": d}, head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
        num_encoder_layers=L, num_attention_heads=H, dropout_rate=0.0, encoder_ff_dim=cfg["dff"])
    batch = make_cloze_batch(np.random.default_rng(0), cfg["B"], V, max_len=cfg["max_len"],
                             mode="train", masked_percentage=cfg["mp"], lengths=cfg["lengths"])
    ids = torch.from_numpy(batch["ids"]).cuda().view(-1)
    labels = torch.from_numpy(batch["labels"]).cuda()
    B, S = batch["ids"].shape
    stats = model.cloze_forward_backward([ids], labels, B, S, n_masked=batch["n_masked"],
                                         training=False).cpu().numpy()
    P = {k: v.astype(np.float64) for k, v in to_reference_layout(model.store.get_weights()).items()}
    pe = O.positional_encoding(10000, d)
    eloss, EG, _ = cloze_train_step_bf16([batch["ids"].astype(np.int64)], batch["labels"], P, L, H, pe)
    assert abs(stats[0] / stats[1] - eloss) < 1e-4 * abs(eloss)
    got = to_reference_layout(model.store.get_grads())
    ek = all_errs(got, EG, EG)
    assert max(ek.values()) < KERNEL_TOL, max(ek.items(), key=lambda kv: kv[1])


# ------------------------------------------------------------------ committed golden vectors
def test_gpu_matches_committed_golden_embed_and_topk(cuda_lib):
    import os
    from bert4clickpath_b200 import ops
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "embed_topk.npz"))
    pe = O.positional_encoding(10000, 24)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out, _ = ops.embed_fwd([dev(z["ids0"]).view(-1), dev(z["ids1"]).view(-1)],
                           [dev(z["t0"]), dev(z["t1"])], dev(pe), 3, 9)
    assert out.cpu().numpy().reshape(3, 9, 24).tobytes() == z["out"].tobytes()  # bit-exact
    ids, _ = ops.topk_rows(dev(z["scores"]), 500, 10)
    assert ids.cpu().numpy().tolist() == z["top10"].tolist()                      # bit-exact


@pytest.mark.parametrize("name,dims", [("cloze_tiny_1feat.npz", (8,)),
                                       ("cloze_tiny_2feat_dropout.npz", (8, 8))])
def test_gpu_matches_committed_golden_cloze_step(cuda_lib, name, dims):
    import os
    from bert4clickpath_b200 import ops
    from bert4clickpath_b200.engine import SITE_INPUT, site
    from bert4clickpath_b200.weights import to_reference_layout
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    P = {k[2:]: z[k] for k in z.files if k.startswith("P/")}
    G = {k[2:]: z[k] for k in z.files if k.startswith("G/")}
    ids_list = [z[k] for k in sorted(f for f in z.files if f.startswith("ids"))]
    L, H = (int(v) for v in z["meta"])
    V = P["head.out.w"].shape[1]
    has_drop = any(k.startswith("mask/") for k in z.files)
    model = build_model(P, dims, L, H, 12, (16, 8), V, rows2=P["emb.1"].shape[0] if len(dims) > 1 else 20)
    B, S = ids_list[0].shape
    dev_ids = [torch.from_numpy(i.astype(np.int32)).cuda().view(-1) for i in ids_list]
    lab = torch.from_numpy(z["labels"].astype(np.float32)).cuda()
    n_masked = int((z["labels"] >= 0).sum())
    if has_drop:
        # the fixture's step used explicit dropout masks drawn by NumPy, which the device's
        # counter-based masks cannot reproduce: on the fixture's WEIGHTS AND INPUTS compare (i) the
        # un-dropped forward with the oracle's un-dropped forward, value by value, and (ii) a
        # training step under the device's own exported masks with the oracle under those masks
        out = model.forward_ids(dev_ids, B, S, training=False, n_masked=n_masked)
        probs = out.materialize().cpu().numpy()
        P64 = {k: v.astype(np.float64) for k, v in P.items()}
        pe = O.positional_encoding(10000, sum(dims))
        x, _ = O.encoder_fwd(ids_list, P64, L, H, pe, np.float64)
        sel, _ = O.select_masked(ids_list[0], x)
        want, _, _ = O.softmax_head_fwd(sel, O.head_layers(P64), P64["head.out.w"], P64["head.out.b"])
        assert probs.shape == z["probs"].shape == want.shape
        np.testing.assert_allclose(probs, want, rtol=5e-2, atol=1e-4)
        # (fp32-class mode: a tiny batch under bf16 operands is dominated by ReLU-gate flips)
        dmodel = build_model(P, dims, L, H, 12, (16, 8), V, dropout=0.25, rows2=P["emb.1"].shape[0],
                             precision="fp32")
        seed, d = 5, sum(dims)
        mk = lambda st: ops.dropout_mask(B * S * d, 0.25, seed, st).cpu().numpy().reshape(B, S, d).astype(np.float64)
        masks = {"in": mk(SITE_INPUT)}
        for l in range(L):
            masks[(l, 1)], masks[(l, 2)] = mk(site(l, 1)), mk(site(l, 2))
        st = dmodel.cloze_forward_backward(dev_ids, lab, B, S, n_masked=n_masked, training=True,
                                           seed=seed).cpu().numpy()
        loss, G2, _ = O.cloze_train_step(ids_list, z["labels"], P64, L, H, pe, np.float64, masks)
        assert abs(st[0] / st[1] - loss) < 1e-4 * abs(loss)
        eb = all_errs(to_reference_layout(dmodel.store.get_grads()), G2, G2)
        assert max(eb.values()) < 1e-3, max(eb.items(), key=lambda kv: kv[1])
        return
    stats = model.cloze_forward_backward(dev_ids, lab, B, S, n_masked=n_masked, training=False)
    s = stats.cpu().numpy()
    assert abs(s[0] / s[1] - float(z["loss"])) < 2e-2 * float(z["loss"])
    got = to_reference_layout(model.store.get_grads())
    eb = all_errs(got, G, G)
    assert max(eb.values()) < BF16_TOL, max(eb.items(), key=lambda kv: kv[1])
    out = model.forward_ids(dev_ids, B, S, training=False, n_masked=n_masked)
    np.testing.assert_allclose(out.materialize().cpu().numpy(), z["probs"], rtol=5e-2, atol=1e-4)


# ------------------------------------------------------------------ C3: multi-variable, segment mode
@pytest.mark.parametrize("segment", [0, 2])
def test_multivariable_segment_head_forward_matches_oracle(cuda_lib, segment):
    """(action, item) click-path encoder with a sigmoid classification head fed from a segment
    slice (clickstream_transformer.py:317-322): segment 0 = [CLS] (purchase intention),
    segment 2 = the basket sequence (return prediction).  String inputs, two chained sequences."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.weights import to_reference_layout
    rng = np.random.default_rng(7)
    items_vocab = [f"it{j}" for j in range(60)]
    ev_vocab = [f"ev{j}" for j in range(7)]
    B, L1, L2 = 6, 9, 4
    def draw(vocab, L):
        a = rng.choice(vocab, size=(B, L)).astype(object)
        for b in range(B):
            n = rng.integers(1, L + 1)
            a[b, n:] = "[PAD]"
        return a
    feats = {"s_items": draw(items_vocab, L1), "b_items": draw(items_vocab, L2)}
    feats["s_ev"] = np.where(feats["s_items"] == "[PAD]", "[PAD]", rng.choice(ev_vocab, size=(B, L1)).astype(object))
    feats["b_ev"] = np.where(feats["b_items"] == "[PAD]", "[PAD]", rng.choice(ev_vocab, size=(B, L2)).astype(object))
    feats["s_items"][0, 1] = "never-seen-item"  # OOV bucket
    head = bc.BinaryClassificationHead(dense_layer_dims=[32, 16])
    model = bc.ClickstreamTransformer(
        sequential_input_config={"items": ["s_items", "b_items"], "events": ["s_ev", "b_ev"]},
        feature_vocabs={"items": items_vocab, "events": ev_vocab},
        embedding_dims={"items": 24, "events": 8}, head_unit=head, segment_to_head=segment,
        num_encoder_layers=2, num_attention_heads=4, dropout_rate=0.1)
    probs = model.call(feats, training=False).cpu().numpy()
    # oracle on the same weights
    P = {k: v.astype(np.float64) for k, v in to_reference_layout(model.store.get_weights()).items()}
    ids_items = O.chain_sequences([O.lookup_ids(feats["s_items"], items_vocab), O.lookup_ids(feats["b_items"], items_vocab)])
    ids_ev = O.chain_sequences([O.lookup_ids(feats["s_ev"], ev_vocab), O.lookup_ids(feats["b_ev"], ev_vocab)])
    assert ids_items[0, 3] == 10 + 60  # OOV id = len(reserved) + len(vocab)
    pe = O.positional_encoding(10000, 32)
    x, _ = O.encoder_fwd([ids_items, ids_ev], P, 2, 4, pe, np.float64)
    starts, ends = O.segment_bounds(ids_items[0])
    seg = O.select_segment(x, starts, ends, segment)
    want, _, _ = O.binary_head_fwd(seg, O.head_layers(P), P["head.out.w"], P["head.out.b"])
    assert probs.shape == want.shape == (B, 1 if segment == 0 else L2)
    np.testing.assert_allclose(probs, want, rtol=3e-2, atol=3e-3)
    # MaskedLoss with binary cross-entropy and pos_weight on the materialised probabilities
    y = rng.integers(0, 2, size=want.shape).astype(np.float32)
    if segment == 2:
        y[feats["b_items"] == "[PAD]"] = -1.0
    loss = bc.MaskedLoss(bc.binary_crossentropy, pos_weight=3.0)(y, probs)
    want_loss = O.masked_loss(y, probs.astype(np.float64), O.binary_crossentropy_probs, pos_weight=3.0)
    assert abs(loss - want_loss) < 1e-4 * abs(want_loss)


@pytest.mark.parametrize("segment,dims", [(0, [32, 16]), (2, [32, 16]), (2, [])])
def test_binary_head_training_matches_autograd(cuda_lib, segment, dims):
    """C3 training path (segment mode + BinaryClassificationHead + MaskedLoss(binary_crossentropy,
    pos_weight), head.py:4-26, losses.py:31-98): loss and every head gradient against torch
    autograd (float64, bf16 rounding at the points where the pipeline stores bf16) on the same
    encoder output; encoder gradients finite and non-zero; Adam steps reduce the loss."""
    import bert4clickpath_b200 as bc
    rng = np.random.default_rng(11 + segment)
    items_vocab = [f"it{j}" for j in range(60)]
    ev_vocab = [f"ev{j}" for j in range(7)]
    B, L1, L2, pw = 37, 9, 4, 3.0
    def draw(vocab, L):
        a = rng.choice(vocab, size=(B, L)).astype(object)
        for b in range(B):
            a[b, rng.integers(1, L + 1):] = "[PAD]"
        return a
    feats = {"s_items": draw(items_vocab, L1), "b_items": draw(items_vocab, L2)}
    feats["s_ev"] = np.where(feats["s_items"] == "[PAD]", "[PAD]", rng.choice(ev_vocab, size=(B, L1)).astype(object))
    feats["b_ev"] = np.where(feats["b_items"] == "[PAD]", "[PAD]", rng.choice(ev_vocab, size=(B, L2)).astype(object))
    head = bc.BinaryClassificationHead(dense_layer_dims=dims)
    model = bc.ClickstreamTransformer(
        sequential_input_config={"items": ["s_items", "b_items"], "events": ["s_ev", "b_ev"]},
        feature_vocabs={"items": items_vocab, "events": ev_vocab},
        embedding_dims={"items": 24, "events": 8}, head_unit=head, segment_to_head=segment,
        num_encoder_layers=1, num_attention_heads=4, dropout_rate=0.0, seed=5)
    ids_list, Bq, S, starts, ends = model.prepare_inputs(feats)
    s0, s1 = int(starts[segment]), int(ends[segment])
    Ls = s1 - s0
    y = rng.integers(0, 2, size=(B, Ls)).astype(np.float32)
    if segment == 2:
        y[feats["b_items"] == "[PAD]"] = -1.0
    yd = torch.from_numpy(y).cuda()
    stats = model.binary_forward_backward(ids_list, yd, B, S, (starts, ends), pos_weight=pw,
                                          training=False).cpu().numpy()
    grads = model.store.get_grads()
    Wts = model.store.get_weights()
    # ---- reference: torch autograd on the head, fed with the device's encoder output
    x = model._encode(ids_list, B, S, False, 0).cpu().numpy().reshape(B, S, 32)
    def q(t):  # straight-through bf16 rounding
        return t + (t.detach().to(torch.bfloat16).to(torch.float64) - t.detach())
    a = q(torch.from_numpy(x[:, s0:s1].reshape(B * Ls, 32).astype(np.float64)))
    params = {}
    for i in range(len(dims)):
        params[f"head.{i}.w"] = torch.tensor(Wts[f"head.{i}.w"], dtype=torch.float64, requires_grad=True)
        params[f"head.{i}.b"] = torch.tensor(Wts[f"head.{i}.b"], dtype=torch.float64, requires_grad=True)
        a = q(torch.relu(a @ q(params[f"head.{i}.w"]) + params[f"head.{i}.b"]))
    params["head.out.w"] = torch.tensor(Wts["head.out.w"], dtype=torch.float64, requires_grad=True)
    params["head.out.b"] = torch.tensor(Wts["head.out.b"], dtype=torch.float64, requires_grad=True)
    p = torch.sigmoid(a @ q(params["head.out.w"]) + params["head.out.b"]).reshape(-1)
    yt = torch.from_numpy(y.reshape(-1).astype(np.float64))
    valid = yt != -1.0
    eps = 1e-7
    pc = torch.clamp(p, eps, 1 - eps)
    item = -(yt * torch.log(pc + eps) + (1 - yt) * torch.log(1 - pc + eps))
    item = torch.where(yt == 1.0, item * pw, item)
    loss = (item * valid).sum() / valid.sum() / ((pw + 1) / 2)
    loss.backward()
    got_loss = stats[0] / stats[1] / ((pw + 1) / 2)
    assert stats[1] == float(valid.sum()) and abs(got_loss - loss.item()) < 2e-3 * abs(loss.item())
    for k, t in params.items():
        g, w = grads[k].astype(np.float64).reshape(t.shape), t.grad.numpy()
        e = np.linalg.norm(g - w) / max(np.linalg.norm(w), 1e-12)
        assert e < 3e-2, (k, e)
    # ---- and against the NumPy oracle's restatement (exact float64 math, no bf16 rounding):
    # the stated bf16 tolerance
    layers = [(Wts[f"head.{i}.w"].astype(np.float64), Wts[f"head.{i}.b"].astype(np.float64)) for i in range(len(dims))]
    o_loss, _, o_lg, o_dWo, o_dbo = O.binary_head_loss_and_grads(
        x[:, s0:s1].astype(np.float64), layers, Wts["head.out.w"].astype(np.float64),
        Wts["head.out.b"].astype(np.float64), y, pos_weight=pw)
    assert abs(got_loss - o_loss) < 2e-2 * abs(o_loss)
    want = {"head.out.w": o_dWo, "head.out.b": o_dbo}
    for i, (dw, db_) in enumerate(o_lg):
        want[f"head.{i}.w"], want[f"head.{i}.b"] = dw, db_
    for k, w in want.items():
        g = grads[k].astype(np.float64).reshape(w.shape)
        assert np.linalg.norm(g - w) / max(np.linalg.norm(w), 1e-12) < BF16_TOL, k
    enc = [k for k in grads if not k.startswith("head.")]
    assert all(np.isfinite(grads[k]).all() for k in enc) and any(np.abs(grads[k]).max() > 0 for k in enc)
    # ---- Keras-style training: the loss goes down
    model.compile(optimizer=bc.Adam(1e-2), loss=bc.MaskedLoss(bc.binary_crossentropy, pos_weight=pw))
    losses = [model.train_step((feats, y))["loss"] for _ in range(12)]
    assert losses[-1] < 0.8 * losses[0], losses


def test_h256_vocabulary_stage_fused_and_in_bounded_row_ranges(cuda_lib):
    """C4 / C5 head width (h = 256, head [] -> V).  Default: the fused TS-form kernels (logits never
    in HBM).  The materialised fallback (`force_materialized`) holds logits for a bounded row range
    at a time: forcing 128-row ranges must reproduce the single-range loss, gradients (dW
    accumulated over ranges) and top-k ids, and the fused path must agree with both."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.synthetic import make_cloze_batch
    V, d = 3001, 256

    def build():
        head = bc.SoftMaxHead(dense_layer_dims=[], output_vocab_size=V)
        return bc.ClickstreamTransformer(
            sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": V},
            embedding_dims={"items": d}, head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
            num_encoder_layers=1, num_attention_heads=4, dropout_rate=0.0, seed=3)

    batch = make_cloze_batch(np.random.default_rng(1), 64, V, max_len=50, mode="train",
                             masked_percentage=0.15)
    ids = torch.from_numpy(batch["ids"]).cuda().view(-1)
    labels = torch.from_numpy(batch["labels"]).cuda()
    B, S = batch["ids"].shape
    ev = make_cloze_batch(np.random.default_rng(2), 300, V, max_len=50, mode="eval")
    ev_ids = torch.from_numpy(ev["ids"]).cuda().view(-1)
    res = []
    for limit in ("fused", None, 1):
        m = build()
        assert m.head.vocab.fused
        if limit != "fused":
            m.head.vocab.force_materialized = True
            assert not m.head.vocab.fused
        if limit == 1:
            m.head.vocab.MATERIALIZE_LIMIT_BYTES = limit      # -> 128-row ranges
            assert len(m.head.vocab._row_chunks(batch["n_masked"])) == -(-batch["n_masked"] // 128) > 2
        st = m.cloze_forward_backward([ids], labels, B, S, n_masked=batch["n_masked"],
                                      training=False).cpu().numpy()
        top, _ = m.topk_ids([ev_ids], 300, ev["ids"].shape[1], 10, n_masked=300)
        res.append((st, m.store.get_grads(), top.cpu().numpy().copy()))
    (sf, gf, _), res = res[0], res[1:]
    np.testing.assert_allclose(sf, res[0][0], rtol=2e-5)
    floor_f = 1e-3 * max(np.linalg.norm(v) for v in res[0][1].values())
    for k in gf:
        g0k = res[0][1][k].astype(np.float64)
        e = np.linalg.norm(gf[k].astype(np.float64) - g0k) / max(np.linalg.norm(g0k), floor_f)
        assert e < 1e-2, ("fused vs materialised", k, e)  # bf16 P' / dZ tiles vs bf16 dZ matrix
    (s0, g0, t0), (s1, g1, t1) = res
    np.testing.assert_allclose(s1, s0, rtol=1e-6)
    assert (t0 == t1).all()
    floor = 1e-3 * max(np.linalg.norm(v) for v in g0.values())
    for k in g0:
        e = np.linalg.norm(g1[k].astype(np.float64) - g0[k]) / max(np.linalg.norm(g0[k]), floor)
        assert e < 2e-3, (k, e)  # split-K order differs per range -> bf16 rounding of dX differs


def test_cuda_graph_replay_is_bit_identical_to_eager_steps(cuda_lib):
    """ClozeTrainStep(use_graph=True): two eager steps, capture, replays.  With dropout seeds read
    on the device from the Adam step counter the replayed steps must reproduce the eagerly
    launched ones bit for bit (same kernels, same order, deterministic reductions)."""
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200 import ops
    from bert4clickpath_b200.synthetic import make_cloze_batch
    from bert4clickpath_b200.training import ClozeTrainStep
    V = 1237

    def build():
        head = bc.SoftMaxHead(dense_layer_dims=[64, 128], output_vocab_size=V)
        return bc.ClickstreamTransformer(
            sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": V},
            embedding_dims={"items": 64}, head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
            num_encoder_layers=2, num_attention_heads=2, dropout_rate=0.1, seed=5)

    rng = np.random.default_rng(0)
    batches = [make_cloze_batch(rng, 32, V, max_len=30, mode="train") for _ in range(3)]
    mg, me = build(), build()
    tg = ClozeTrainStep(mg, bc.Adam(1e-3), use_graph=True)
    te = ClozeTrainStep(me, bc.Adam(1e-3))
    losses = []
    for i in range(7):
        b = batches[i % 3]
        sg = tg.step_device(tg.to_device(b)).clone()
        se = te._eager(te.to_device(b), ops.device_seed(me.store.step_dev)).clone()
        torch.cuda.synchronize()
        assert torch.equal(sg, se), i
        losses.append(float(sg[0] / sg[1]))
    assert len(tg._graphs) == 1 and next(iter(tg._graphs.values()))["graph"] is not None
    wg, we = mg.store.get_weights(), me.store.get_weights()
    for k in wg:
        assert np.array_equal(wg[k], we[k]), k
    assert int(mg.store.step_dev.item()) == 8
    assert losses[-1] < losses[0]
    # dropout masks differ from step to step (the seed is the step counter)
    m1 = ops.dropout_mask(1000, 0.1, ops.device_seed(mg.store.step_dev), 1).clone()
    ops.step_increment(mg.store.step_dev)
    m2 = ops.dropout_mask(1000, 0.1, ops.device_seed(mg.store.step_dev), 1)
    assert not torch.equal(m1, m2)
    assert torch.equal(m2, ops.dropout_mask(1000, 0.1, 9, 1))  # device seed == same host seed
