import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
V, h, k = 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 256, 100
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wb = torch.zeros(h, ops.ld8(V), device="cuda", dtype=torch.bfloat16)
wb[:, :V] = (torch.randn(h, V, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.zeros(V, device="cuda")
xb = (torch.randn(B, h, device="cuda") * 0.5).to(torch.bfloat16)
ids = torch.empty(B, k, dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.score_topk(xb, B, h, wb, bias, V, k, out_ids=ids)
torch.cuda.synchronize()
