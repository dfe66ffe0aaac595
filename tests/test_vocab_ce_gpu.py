"""Fused tcgen05 vocab-projection + softmax-CE kernels against the oracle's logits-mode Cloze
loss on bf16-rounded operands (the kernels' declared input precision)."""
import numpy as np
import pytest
import torch

from oracle import clickpath_oracle as O
from oracle.mixed_precision import bf16

pytestmark = pytest.mark.gpu


def make(M, h, V, seed, npad=3):
    rng = np.random.default_rng(seed)
    x = bf16(rng.normal(size=(M, h)) * 0.7)
    w = bf16(rng.normal(size=(h, V)) * 0.2)
    b = (rng.normal(size=V) * 0.3).astype(np.float32)
    labels = rng.integers(0, V, size=M).astype(np.int32)
    labels[rng.choice(M, size=min(npad, M), replace=False)] = -1
    return x, w, b, labels


def to_dev(x, w, b, labels):
    from bert4clickpath_b200 import ops
    xb = torch.from_numpy(x.astype(np.float32)).cuda().to(torch.bfloat16).contiguous()
    V = w.shape[1]
    wb = torch.zeros((w.shape[0], ops.ld8(V)), dtype=torch.bfloat16, device="cuda")
    wb[:, :V] = torch.from_numpy(w.astype(np.float32)).cuda().to(torch.bfloat16)
    return xb, wb, torch.from_numpy(b).cuda(), torch.from_numpy(labels).cuda()


@pytest.mark.parametrize("want_dx", [False, True])
@pytest.mark.parametrize("M,h,V", [(300, 128, 1237), (128, 128, 128), (77, 128, 54293),
                                   (1000, 64, 5000), (260, 256, 3001)])
def test_fused_forward_lse_and_target(cuda_lib, M, h, V, want_dx):
    want_dx = want_dx and h in (128, 256)
    from bert4clickpath_b200 import ops
    x, w, b, labels = make(M, h, V, M + V)
    z = x @ w + b.astype(np.float64)
    m = z.max(-1)
    lse_ref = m + np.log(np.exp(z - m[:, None]).sum(-1))
    tgt_ref = np.where(labels >= 0, z[np.arange(M), np.maximum(labels, 0)], 0.0)
    xb, wb, bd, ld = to_dev(x, w, b, labels)
    lse = torch.full((M,), float("nan"), device="cuda")
    tgt = torch.full((M,), float("nan"), device="cuda")
    ops.vocab_ce_fwd(xb, M, h, wb, bd, V, ld, lse, tgt, want_dx=want_dx)
    torch.cuda.synchronize()
    np.testing.assert_allclose(lse.cpu().numpy(), lse_ref, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(tgt.cpu().numpy(), tgt_ref, rtol=1e-5, atol=1e-5)
    stats = torch.empty(2, device="cuda")
    ops.ce_loss_reduce(lse, tgt, ld, stats)
    loss, _, n = O.cloze_ce_from_logits(z, labels)
    s = stats.cpu().numpy()
    assert s[1] == n and abs(s[0] / s[1] - loss) < 1e-5 * abs(loss)


@pytest.mark.parametrize("M,V,h", [(300, 1237, 128), (128, 128, 128), (77, 54293, 128),
                                   (1000, 5000, 128), (5, 300, 128), (300, 1237, 256),
                                   (129, 54293, 256), (1000, 5000, 256), (5, 300, 256)])
def test_fused_backward_gradients(cuda_lib, M, V, h):
    from bert4clickpath_b200 import ops
    x, w, b, labels = make(M, h, V, 7 * M + V)
    z = x @ w + b.astype(np.float64)
    loss, dz, n = O.cloze_ce_from_logits(z, labels)
    dzq = bf16(dz)  # the kernel feeds bf16 dZ to the tensor cores
    # db is a register sum of the unrounded fp32 dZ (lane = vocabulary entry): exact reference
    dX_ref, dW_ref, db_ref = dzq @ w.T, x.T @ dzq, dz.sum(0)
    xb, wb, bd, ld = to_dev(x, w, b, labels)
    lse = torch.empty(M, device="cuda")
    tgt = torch.empty(M, device="cuda")
    stats = torch.empty(2, device="cuda")
    ops.vocab_ce_fwd(xb, M, h, wb, bd, V, ld, lse, tgt, want_dx=True)
    ops.ce_loss_reduce(lse, tgt, ld, stats)
    dX = torch.full((M, h), float("nan"), device="cuda")
    dW = torch.full((h, V), float("nan"), device="cuda")
    db = torch.full((V,), float("nan"), device="cuda")
    ops.vocab_ce_dx(M, h, V, ld, stats, wb, None, dX, None)
    ops.vocab_ce_bwd(xb, M, h, wb, bd, V, ld, lse, stats, dW, db)
    torch.cuda.synchronize()
    # dX: the forward accumulates exp(z - max) as bf16 against W (flash-attention style), i.e.
    # (sum_v bf16(p'_v) W_v) / sum - W_label: compare with the exact softmax expectation
    dX_exact = dz @ w.T
    for name, got, want, tol in (("dX", dX, dX_exact, 4e-3), ("dW", dW, dW_ref, 2e-3),
                                 ("db", db, db_ref, 4e-3)):
        g = got.cpu().numpy()
        assert np.isfinite(g).all(), name
        # fp32 accumulation + bf16 rounding of the probabilities: fraction of the max-norm
        assert np.abs(g - want).max() <= tol * np.abs(want).max(), (name, np.abs(g - want).max() / np.abs(want).max())
    # padded rows get exactly zero gradient
    assert not dX.cpu().numpy()[labels < 0].any()
    dW2 = torch.empty((h, V), device="cuda")
    dX2 = torch.empty((M, h), device="cuda")
    ops.vocab_ce_fwd(xb, M, h, wb, bd, V, ld, lse, tgt, want_dx=True)
    ops.vocab_ce_dx(M, h, V, ld, stats, wb, None, dX2, None)
    ops.vocab_ce_bwd(xb, M, h, wb, bd, V, ld, lse, stats, dW2, db)
    assert torch.equal(dW, dW2) and torch.equal(dX, dX2)  # fixed accumulation order: reproducible


def test_fused_lazy_rescale_of_the_running_maximum(cuda_lib):
    """The forward keeps U = sum_v exp2(z - m_ref) W_v in TMEM and only rescales it when a row's
    running maximum jumps by more than 2^8: logits that grow along the vocabulary (so the maximum
    moves in many tiles, by more and by less than the threshold) must give the same dX."""
    from bert4clickpath_b200 import ops
    M, h, V = 200, 128, 4000
    x, w, b, labels = make(M, h, V, 99)
    b = (np.linspace(-60.0, 60.0, V) + np.random.default_rng(5).normal(size=V) * 4).astype(np.float32)
    b[V // 2] = 90.0    # one big jump in the middle of a chunk
    z = x @ w + b.astype(np.float64)
    loss, dz, n = O.cloze_ce_from_logits(z, labels)
    xb, wb, bd, ld = to_dev(x, w, b, labels)
    lse, tgt, stats = (torch.empty(M, device="cuda"), torch.empty(M, device="cuda"),
                       torch.empty(2, device="cuda"))
    ops.vocab_ce_fwd(xb, M, h, wb, bd, V, ld, lse, tgt, want_dx=True)
    ops.ce_loss_reduce(lse, tgt, ld, stats)
    dX = torch.full((M, h), float("nan"), device="cuda")
    ops.vocab_ce_dx(M, h, V, ld, stats, wb, None, dX, None)
    torch.cuda.synchronize()
    m = z.max(-1)
    np.testing.assert_allclose(lse.cpu().numpy(), m + np.log(np.exp(z - m[:, None]).sum(-1)), rtol=2e-5, atol=2e-5)
    want = dz @ w.T
    g = dX.cpu().numpy()
    assert np.isfinite(g).all()
    assert np.abs(g - want).max() <= 4e-3 * np.abs(want).max()


@pytest.mark.parametrize("h", [128, 256])
def test_fused_all_rows_padded(cuda_lib, h):
    from bert4clickpath_b200 import ops
    M, V = 130, 700
    x, w, b, labels = make(M, h, V, 1)
    labels[:] = -1
    xb, wb, bd, ld = to_dev(x, w, b, labels)
    lse, tgt, stats = (torch.empty(M, device="cuda"), torch.empty(M, device="cuda"),
                       torch.empty(2, device="cuda"))
    ops.vocab_ce_fwd(xb, M, h, wb, bd, V, ld, lse, tgt, want_dx=True)
    ops.ce_loss_reduce(lse, tgt, ld, stats)
    dX, dW, db = (torch.ones((M, h), device="cuda"), torch.ones((h, V), device="cuda"),
                  torch.ones(V, device="cuda"))
    ops.vocab_ce_dx(M, h, V, ld, stats, wb, None, dX, None)
    ops.vocab_ce_bwd(xb, M, h, wb, bd, V, ld, lse, stats, dW, db)
    torch.cuda.synchronize()
    assert stats.cpu().numpy().tolist() == [0.0, 0.0]
    assert not dX.any().item() and not dW.any().item() and not db.any().item()


# ------------------------------------------------------------------ fused scoring + top-k
@pytest.mark.parametrize("M,h,V,k", [(5, 128, 54293, 100), (300, 128, 5000, 10), (1, 64, 1237, 5),
                                     (130, 128, 200, 100), (700, 128, 100000, 100),
                                     (300, 256, 5000, 10), (131, 256, 100000, 100), (2, 256, 300, 104)])
def test_fused_score_topk_is_exact(cuda_lib, M, h, V, k):
    """ids bit-exact against the oracle's tf.math.top_k order applied to the same tensor-core
    scores, including exact ties across tiles and vocabulary chunks (duplicated W columns)."""
    from bert4clickpath_b200 import ops
    x, w, b, _ = make(M, h, V, 3 * M + V)
    w[:, V // 3:] = w[:, (np.arange(V - V // 3) % 53)]   # thousands of exactly tied scores
    b[V // 3:] = b[np.arange(V - V // 3) % 53]
    labels = np.zeros(M, dtype=np.int32)
    xb, wb, bd, _ = to_dev(x, w, b, labels)
    logits = torch.empty((M, ops.ld8(V)), device="cuda")
    ops.gemm(xb, 0, wb, 1, M, V, h, bias=bd, out_f32=logits)
    ids, scores = ops.score_topk(xb, M, h, wb, bd, V, k, out_scores=torch.empty((M, k), device="cuda"))
    torch.cuda.synchronize()
    z = logits.cpu().numpy()[:, :V]
    want = O.top_k_ids(z, k)
    kk = min(k, V)
    got = ids.cpu().numpy()
    assert got[:, :kk].tolist() == want[:, :kk].tolist()
    assert (got[:, kk:] == -1).all()
    np.testing.assert_array_equal(scores.cpu().numpy()[:, :kk], np.take_along_axis(z, want[:, :kk], 1))
    # and the materialised top-k kernel agrees
    ids2, _ = ops.topk_rows(logits, V, k)
    assert torch.equal(ids, ids2)


@pytest.mark.parametrize("M,h,V,k,adversarial", [(131, 256, 300000, 100, False), (300, 128, 400003, 10, False),
                                                 (70, 128, 280000, 100, True)])
def test_fused_score_topk_long_vocabulary(cuda_lib, M, h, V, k, adversarial):
    """V >= 262144: seed (first 65536 entries ranked exactly) + threshold filter in the tcgen05
    epilogue + exact merge.  `adversarial`: scores grow along the vocabulary, every later entry
    beats the seed threshold, the candidate lists overflow and the rows are redone by the heap
    kernel behind the device-side flag.  Ids are exact either way (ties included)."""
    from bert4clickpath_b200 import ops
    x, w, b, _ = make(M, h, V, 17 * M + V)
    w[:, V // 2:] = w[:, (np.arange(V - V // 2) % 101)]      # exact ties across seed / rest
    b[V // 2:] = b[np.arange(V - V // 2) % 101]
    if adversarial:
        b = (b + np.linspace(0.0, 40.0, V)).astype(np.float32)
    labels = np.zeros(M, dtype=np.int32)
    xb, wb, bd, _ = to_dev(x, w, b, labels)
    ids, scores = ops.score_topk(xb, M, h, wb, bd, V, k, out_scores=torch.empty((M, k), device="cuda"))
    torch.cuda.synchronize()
    got = ids.cpu().numpy()
    want = np.empty((M, k), dtype=np.int64)
    sc_want = np.empty((M, k), dtype=np.float32)
    for a in range(0, M, 64):                                  # bounded logits on the device
        rows = min(64, M - a)
        logits = torch.empty((rows, ops.ld8(V)), device="cuda")
        ops.gemm(xb[a:a + rows], 0, wb, 1, rows, V, h, bias=bd, out_f32=logits)
        z = logits.cpu().numpy()[:, :V]
        want[a:a + rows] = O.top_k_ids(z, k)
        sc_want[a:a + rows] = np.take_along_axis(z, want[a:a + rows], 1)
    assert got.tolist() == want.tolist()
    np.testing.assert_array_equal(scores.cpu().numpy(), sc_want)


# ------------------------------------------------------------------ vocabulary-parallel pieces
def _shards(V, R):
    per = (V + R - 1) // R
    per = (per + 7) // 8 * 8
    return [(min(r * per, V), min(r * per + per, V)) for r in range(R)]


@pytest.mark.parametrize("M,V,R", [(300, 1237, 2), (77, 54293, 3), (1000, 5000, 4)])
def test_vocab_shards_merge_to_the_unsharded_result(cuda_lib, M, V, R):
    """The vocabulary-parallel algebra on one GPU: per-shard fused kernels (labels remapped by
    b4cp_shard_labels), b4cp_lse_merge, dx with the global lse summed over shards, shard-local
    dW/db — against the unsharded kernels and the oracle.  (The NCCL exchanges of
    VocabParallelOutputEngine are replaced by sums / concatenation here.)"""
    from bert4clickpath_b200 import ops
    h = 128
    x, w, b, labels = make(M, h, V, 11 * M + V)
    z = x @ w + b.astype(np.float64)
    loss, dz, n = O.cloze_ce_from_logits(z, labels)
    xb, wb, bd, ld = to_dev(x, w, b, labels)
    # unsharded reference run of the same kernels
    lse0, tgt0, stats0 = (torch.empty(M, device="cuda"), torch.empty(M, device="cuda"),
                          torch.empty(2, device="cuda"))
    ops.vocab_ce_fwd(xb, M, h, wb, bd, V, ld, lse0, tgt0, want_dx=True)
    ops.ce_loss_reduce(lse0, tgt0, ld, stats0)
    dX0, dW0, db0 = (torch.empty((M, h), device="cuda"), torch.empty((h, V), device="cuda"),
                     torch.empty(V, device="cuda"))
    ops.vocab_ce_dx(M, h, V, ld, stats0, wb, None, dX0, None)
    ops.vocab_ce_bwd(xb, M, h, wb, bd, V, ld, lse0, stats0, dW0, db0)

    shards = _shards(V, R)
    lse_parts = torch.empty((R, M), device="cuda")
    tgt_sum = torch.zeros(M, device="cuda")
    lab_sh, w_sh, b_sh = [], [], []
    for r, (v0, v1) in enumerate(shards):
        xs, ws, bs, _ = to_dev(x, w[:, v0:v1], b[v0:v1], labels)
        ls = ops.shard_labels(ld, v0, v1 - v0)
        want = np.where(labels < 0, -1, np.where((labels >= v0) & (labels < v1), labels - v0, 0x3FFFFFFF))
        assert ls.cpu().numpy().tolist() == want.tolist()
        lab_sh.append(ls); w_sh.append(ws); b_sh.append(bs)
    stats = torch.empty(2, device="cuda")
    # forward of every shard, then the merge
    for r, (v0, v1) in enumerate(shards):
        tgt = torch.full((M,), float("nan"), device="cuda")
        ops.vocab_ce_fwd(xb, M, h, w_sh[r], b_sh[r], v1 - v0, lab_sh[r], lse_parts[r], tgt, want_dx=False)
        tgt_sum += tgt
    lse = ops.lse_merge(lse_parts)
    ops.ce_loss_reduce(lse, tgt_sum, ld, stats)
    torch.cuda.synchronize()
    m = z.max(-1)
    np.testing.assert_allclose(lse.cpu().numpy(), m + np.log(np.exp(z - m[:, None]).sum(-1)), rtol=2e-5, atol=2e-5)
    s = stats.cpu().numpy()
    assert s[1] == n and abs(s[0] / s[1] - loss) < 1e-5 * abs(loss)
    np.testing.assert_allclose(s, stats0.cpu().numpy(), rtol=2e-6)
    # backward: the workspace holds one shard's forward at a time
    dX = torch.zeros((M, h), device="cuda")
    for r, (v0, v1) in enumerate(shards):
        Vs = v1 - v0
        scratch_lse, scratch_tgt = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
        ops.vocab_ce_fwd(xb, M, h, w_sh[r], b_sh[r], Vs, lab_sh[r], scratch_lse, scratch_tgt, want_dx=True)
        part = torch.full((M, h), float("nan"), device="cuda")
        ops.vocab_ce_dx(M, h, Vs, lab_sh[r], stats, w_sh[r], None, part, None, lse_global=lse)
        dX += part
        dW = torch.full((h, Vs), float("nan"), device="cuda")
        db = torch.full((Vs,), float("nan"), device="cuda")
        ops.vocab_ce_bwd(xb, M, h, w_sh[r], b_sh[r], Vs, lab_sh[r], lse, stats, dW, db)
        torch.cuda.synchronize()
        for name, got, ref in (("dW", dW, dW0[:, v0:v1]), ("db", db, db0[v0:v1])):
            g, f = got.cpu().numpy(), ref.cpu().numpy()
            assert np.isfinite(g).all()
            assert np.abs(g - f).max() <= 2e-3 * max(np.abs(f).max(), 1e-12), (name, r)
    torch.cuda.synchronize()
    g = dX.cpu().numpy()
    dX_exact = dz @ w.T
    assert np.abs(g - dX_exact).max() <= 4e-3 * np.abs(dX_exact).max()
    assert np.abs(g - dX0.cpu().numpy()).max() <= 4e-3 * np.abs(dX_exact).max()
    assert not g[labels < 0].any()


@pytest.mark.parametrize("M,V,R,k", [(200, 5000, 2, 10), (64, 54293, 4, 100), (33, 300, 3, 100)])
def test_vocab_shards_topk_merge_is_exact(cuda_lib, M, V, R, k):
    """Per-shard fused top-k with global ids (id_base) + candidate merge == unsharded top-k,
    including exact score ties that straddle shard boundaries."""
    from bert4clickpath_b200 import ops
    h = 128
    x, w, b, _ = make(M, h, V, 5 * M + V)
    w[:, V // 4:] = w[:, (np.arange(V - V // 4) % 37)]
    b[V // 4:] = b[np.arange(V - V // 4) % 37]
    labels = np.zeros(M, dtype=np.int32)
    xb, wb, bd, _ = to_dev(x, w, b, labels)
    ids0, sc0 = ops.score_topk(xb, M, h, wb, bd, V, k, out_scores=torch.empty((M, k), device="cuda"))
    cand_i, cand_s = [], []
    for v0, v1 in _shards(V, R):
        _, ws, bs, _ = to_dev(x, w[:, v0:v1], b[v0:v1], labels)
        kk = min(k, v1 - v0)
        i = torch.full((M, k), -1, dtype=torch.int32, device="cuda")
        s = torch.full((M, k), float("-inf"), device="cuda")
        ops.score_topk(xb, M, h, ws, bs, v1 - v0, kk, out_ids=i, out_scores=s, id_base=v0, V_total=V)
        cand_i.append(i); cand_s.append(s)
    ci, cs = torch.cat(cand_i, 1).contiguous(), torch.cat(cand_s, 1).contiguous()
    ids, sc = ops.topk_candidates(cs, ci, V, k, out_scores=torch.empty((M, k), device="cuda"))
    torch.cuda.synchronize()
    assert torch.equal(ids, ids0)
    kk = min(k, V)
    assert torch.equal(sc[:, :kk], sc0[:, :kk])
