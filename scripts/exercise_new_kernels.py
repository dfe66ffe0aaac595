"""Small invocations of the kernels added late in round 1 (TS vocabulary kernels at h = 128 / 256,
streaming top-k incl. its fallbacks, long-vocabulary fused top-k, tile-based embedding backward,
S = 202 attention): a quick on-device exercise (compute-sanitizer is closed on this pool)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
torch.manual_seed(0)
dev = "cuda"
for (M, h, V) in ((300, 128, 1237), (131, 256, 3001)):
    xb = (torch.randn(M, h, device=dev) * 0.5).to(torch.bfloat16)
    wb = torch.zeros(h, ops.ld8(V), device=dev, dtype=torch.bfloat16); wb[:, :V] = (torch.randn(h, V, device=dev) * 0.1).to(torch.bfloat16)
    bias = torch.randn(V, device=dev) * 0.1
    labels = torch.randint(0, V, (M,), device=dev, dtype=torch.int32); labels[:3] = -1
    lse, tgt, stats = torch.empty(M, device=dev), torch.empty(M, device=dev), torch.empty(2, device=dev)
    ops.vocab_ce_fwd(xb, M, h, wb, bias, V, labels, lse, tgt, want_dx=True)
    ops.ce_loss_reduce(lse, tgt, labels, stats)
    dX = torch.empty(M, h, device=dev); dW = torch.empty(h, V, device=dev); db = torch.empty(V, device=dev)
    ops.vocab_ce_dx(M, h, V, labels, stats, wb, None, dX, None)
    ops.vocab_ce_bwd(xb, M, h, wb, bias, V, labels, lse, stats, dW, db)
    ids, _ = ops.score_topk(xb, M, h, wb, bias, V, 10)
for V, k in ((20000, 100), (300001, 10), (16384, 256)):
    sc = torch.randn(5, ops.ld8(V), device=dev); sc[0, :V] = torch.arange(V, device=dev).float(); sc[1] = 0.25
    ops.topk_rows(sc, V, k)
M, h, V = 70, 128, 280000
xb = (torch.randn(M, h, device=dev) * 0.5).to(torch.bfloat16)
wb = torch.zeros(h, ops.ld8(V), device=dev, dtype=torch.bfloat16); wb[:, :V] = (torch.randn(h, V, device=dev) * 0.1).to(torch.bfloat16)
ops.score_topk(xb, M, h, wb, torch.zeros(V, device=dev), V, 100)
B, S, d, rows = 64, 52, 64, 5000
ids = torch.randint(10, rows, (B * S,), device=dev, dtype=torch.int32); ids[::3] = 1
ops.embed_bwd(torch.randn(B * S, d, device=dev), d, 0, d, ids, rows, torch.empty(rows, d, device=dev))
B, S, H, dh = 2, 202, 2, 64
dm = H * dh
qkv = (torch.randn(B * S, 3 * dm, device=dev) * 0.5).to(torch.bfloat16)
out = torch.empty(B * S, dm, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, S, device=dev)
idsa = torch.randint(0, 50, (B * S,), device=dev, dtype=torch.int32)
ops.attention_fwd(qkv, idsa, B, S, H, dh, out, lse)
ops.attention_bwd(qkv, out, lse, idsa, B, S, H, dh, torch.empty(B * S, 3 * dm, device=dev, dtype=torch.bfloat16), out=out)
torch.cuda.synchronize()
print("sanitize script ok")
