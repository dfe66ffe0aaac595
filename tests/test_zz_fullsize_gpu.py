"""Size-independent properties at BASELINE.json's full shapes (too large for the NumPy oracle):
C1/C2 output stage (M = 28,672 [MASK] rows, V = 54,293, h = 128), C4 output stage (V = 1,000,000,
h = 256) and C5 top-k (V = 1,000,000).  Checked: row independence (any subset of rows run alone
gives the same per-row results), conservation laws of softmax-CE gradients (every dZ row sums to
zero => sum_v db = 0 and every dW row sums to zero), run-to-run bit reproducibility, and for top-k:
sortedness, agreement of the reported scores with the score matrix, and subset consistency."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stage(ops, xb, wb, bias, labels, M, h, V):
    lse, tgt, stats = (torch.empty(M, device="cuda"), torch.empty(M, device="cuda"),
                       torch.empty(2, device="cuda"))
    ops.vocab_ce_fwd(xb, M, h, wb, bias, V, labels, lse, tgt, want_dx=True)
    ops.ce_loss_reduce(lse, tgt, labels, stats)
    dX = torch.empty(M, h, device="cuda")
    dW = torch.empty(h, V, device="cuda")
    db = torch.empty(V, device="cuda")
    ops.vocab_ce_dx(M, h, V, labels, stats, wb, None, dX, None)
    ops.vocab_ce_bwd(xb, M, h, wb, bias, V, labels, lse, stats, dW, db)
    return lse, tgt, stats, dX, dW, db


@pytest.mark.parametrize("M,h,V", [(28672, 128, 54293), (7424, 256, 1_000_000)])
def test_output_stage_properties_at_full_size(cuda_lib, M, h, V):
    from bert4clickpath_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + V)
    xb = (torch.randn(M, h, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    wb = torch.zeros(h, ops.ld8(V), device="cuda", dtype=torch.bfloat16)
    wb[:, :V] = (torch.randn(h, V, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    bias = torch.randn(V, device="cuda", generator=g) * 0.2
    labels = torch.randint(0, V, (M,), device="cuda", dtype=torch.int32, generator=g)
    labels[::97] = -1
    lse, tgt, stats, dX, dW, db = _stage(ops, xb, wb, bias, labels, M, h, V)
    torch.cuda.synchronize()
    n = int((labels >= 0).sum().item())
    assert stats[1].item() == n and torch.isfinite(stats).all()
    # conservation: each dZ row sums to zero  =>  sum_v db[v] = 0 and sum_v dW[j, v] = 0
    # (entries are O(1/n); the sums cancel to rounding level)
    assert abs(db.double().sum().item()) < 2e-4
    assert dW.double().sum(1).abs().max().item() < 2e-3 * dW.double().abs().sum(1).max().item()
    # padded rows: exactly zero dX
    assert not dX[labels < 0].any().item()
    # row independence: a subset of rows run alone gives the same lse / target logit
    idx = torch.arange(0, M, max(1, M // 200), device="cuda")[:160]
    m = idx.numel()
    lse_s, tgt_s = torch.empty(m, device="cuda"), torch.empty(m, device="cuda")
    ops.vocab_ce_fwd(xb[idx].contiguous(), m, h, wb, bias, V, labels[idx].contiguous(), lse_s, tgt_s, want_dx=False)
    torch.cuda.synchronize()
    np.testing.assert_allclose(lse_s.cpu().numpy(), lse[idx].cpu().numpy(), rtol=2e-6, atol=2e-6)
    valid = (labels[idx] >= 0).cpu().numpy()
    np.testing.assert_allclose(tgt_s.cpu().numpy()[valid], tgt[idx].cpu().numpy()[valid], rtol=1e-6, atol=1e-6)
    # the loss is the masked mean of lse - target logit
    want = ((lse - tgt)[labels >= 0]).double().sum().item()
    assert abs(stats[0].item() - want) < 1e-4 * abs(want)
    # bit reproducibility of the whole stage
    lse2, tgt2, stats2, dX2, dW2, db2 = _stage(ops, xb, wb, bias, labels, M, h, V)
    torch.cuda.synchronize()
    assert torch.equal(lse, lse2) and torch.equal(dX, dX2) and torch.equal(dW, dW2) and torch.equal(db, db2)


def test_topk_properties_at_catalogue_size(cuda_lib):
    """C5: V = 1,000,000, k = 100.  Scores come out sorted (descending, ties by ascending id), equal
    the entries of the score matrix at the reported ids, no score outside the list beats the k-th,
    and ranking a subset of rows alone gives the same ids."""
    from bert4clickpath_b200 import ops
    V, k, B = 1_000_000, 100, 96
    g = torch.Generator(device="cuda").manual_seed(5)
    sc = torch.randn(B, ops.ld8(V), device="cuda", generator=g)
    sc[1, :V] = torch.round(sc[1, :V] * 4) / 4                 # heavy ties
    sc[2, :V] = torch.arange(V, device="cuda").float()         # ascending: radix-select fallback
    ids, out = ops.topk_rows(sc, V, k, out_scores=torch.empty(B, k, device="cuda"))
    torch.cuda.synchronize()
    idl = ids.long()
    assert (idl >= 0).all() and (idl < V).all()
    assert torch.equal(out, torch.gather(sc, 1, idl))
    d = out[:, 1:] - out[:, :-1]
    assert (d <= 0).all()
    tie = d == 0
    assert (idl[:, 1:][tie] > idl[:, :-1][tie]).all()
    kth = out[:, -1:]
    masked = sc[:, :V].clone()
    masked.scatter_(1, idl, float("-inf"))
    assert (masked <= kth).all()                               # nothing better was left out
    beat = (masked == kth)                                     # equal scores left out must have higher ids
    rows_with = beat.any(1).nonzero().flatten().tolist()
    for r in rows_with[:8]:
        assert beat[r].nonzero().min().item() > idl[r, -1].item()
    sub = torch.tensor([0, 1, 2, 17, 95], device="cuda")
    ids_s, _ = ops.topk_rows(sc[sub].contiguous(), V, k)
    torch.cuda.synchronize()
    assert torch.equal(ids_s, ids[sub])
