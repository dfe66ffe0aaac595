"""One pass of the fused vocabulary stage at the bench shape (for `ncu --set full`):
vocab_ce_fwd (with U for dX) -> merge -> loss reduce -> vocab_ce_dx -> vocab_ce_bwd, twice."""
import sys

import torch

sys.path.insert(0, ".")
from bert4clickpath_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 28672
V = int(sys.argv[2]) if len(sys.argv) > 2 else 54293
h = int(sys.argv[3]) if len(sys.argv) > 3 else 128
torch.manual_seed(0)
xb = (torch.randn(M, h, device="cuda") * 0.5).to(torch.bfloat16)
wb = torch.zeros(h, ops.ld8(V), device="cuda", dtype=torch.bfloat16)
wb[:, :V] = (torch.randn(h, V, device="cuda") * 0.1).to(torch.bfloat16)
bias = torch.zeros(V, device="cuda")
labels = torch.randint(0, V, (M,), device="cuda", dtype=torch.int32)
lse, tgt, stats = torch.empty(M, device="cuda"), torch.empty(M, device="cuda"), torch.empty(2, device="cuda")
dXb = torch.empty(M, h, device="cuda", dtype=torch.bfloat16)
dW, db = torch.empty(h, V, device="cuda"), torch.empty(V, device="cuda")
for _ in range(2):
    ops.vocab_ce_fwd(xb, M, h, wb, bias, V, labels, lse, tgt, want_dx=True)
    ops.ce_loss_reduce(lse, tgt, labels, stats)
    ops.vocab_ce_dx(M, h, V, labels, stats, wb, xb, None, dXb)
    ops.vocab_ce_bwd(xb, M, h, wb, bias, V, labels, lse, stats, dW, db)
torch.cuda.synchronize()
print("ok", stats.tolist())
