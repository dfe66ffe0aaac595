"""embed_bwd at the C1 and C4 roofline shapes, for an ncu launch list."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
g = np.random.default_rng(0)
for B, S, V, d in ((16384, 52, 54293, 64), (1024, 202, 1_000_000, 256)):
    rows = V + 11
    p = 1.0 / np.arange(1, V + 1) ** 0.8; p /= p.sum()
    ids = torch.from_numpy(g.choice(V, size=B * S, p=p).astype(np.int32) + 10).cuda()
    dout = torch.randn(B * S, d, device="cuda")
    tg = torch.empty(rows, d, device="cuda")
    for _ in range(2):
        ops.embed_bwd(dout, d, 0, d, ids, rows, tg)
    torch.cuda.synchronize()
print("ok")
