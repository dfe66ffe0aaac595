"""Two launches of the streaming top-k for ncu: python scripts/profile_topk.py"""
import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
for B, V, k in ((592, 1_000_000, 100), (4736, 54293, 100)):
    sc = torch.randn(B, ops.ld8(V), device="cuda")
    ids = torch.empty(B, k, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ops.topk_rows(sc, V, k, out_ids=ids)
    torch.cuda.synchronize()
    del sc
print("ok")
