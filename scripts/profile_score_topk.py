import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
B = int(sys.argv[1]); h = int(sys.argv[2]); k = int(sys.argv[3]); V = int(sys.argv[4]) if len(sys.argv) > 4 else 1_000_000
wb = torch.zeros(h, ops.ld8(V), device="cuda", dtype=torch.bfloat16)
wb[:, :V] = (torch.randn(h, V, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.zeros(V, device="cuda")
xb = (torch.randn(B, h, device="cuda") * 0.5).to(torch.bfloat16)
ids = torch.empty(B, k, dtype=torch.int32, device="cuda")
for _ in range(2): ops.score_topk(xb, B, h, wb, bias, V, k, out_ids=ids)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): ops.score_topk(xb, B, h, wb, bias, V, k, out_ids=ids)
e1.record(); torch.cuda.synchronize()
print(f"B={B} h={h} k={k} V={V}: {e0.elapsed_time(e1)/3:.3f} ms")
