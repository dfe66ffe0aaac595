"""Why is vocab_ce_fwd slower inside the step than alone?  Times single launches of the C1-shape
forward under different preceding contexts: python scripts/probe_fwd_context.py"""
import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
M, h, V = 28672, 128, 54293
xb = (torch.randn(M, h, device="cuda") * 0.5).to(torch.bfloat16)
wb = torch.zeros(h, ops.ld8(V), device="cuda", dtype=torch.bfloat16); wb[:, :V] = (torch.randn(h, V, device="cuda") * 0.1).to(torch.bfloat16)
bias = torch.zeros(V, device="cuda"); labels = torch.randint(0, V, (M,), device="cuda", dtype=torch.int32)
lse = torch.empty(M, device="cuda"); tgt = torch.empty(M, device="cuda")
big = torch.empty(64 << 20, device="cuda")          # 256 MB
act = torch.empty(M, 1024, device="cuda", dtype=torch.bfloat16)
fwd = lambda: ops.vocab_ce_fwd(xb, M, h, wb, bias, V, labels, lse, tgt, want_dx=True)
def single(pre, n=12):
    ts = []
    for _ in range(3): fwd()
    for _ in range(n):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fwd(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
print("back-to-back (L2 warm)        %.3f ms" % single(lambda: None))
print("after 256 MB fill (L2 dirty)  %.3f ms" % single(lambda: big.fill_(1.0)))
print("after 58 MB bf16 fill         %.3f ms" % single(lambda: act.fill_(1.0)))
print("after 10 ms idle spin         %.3f ms" % single(lambda: torch.cuda._sleep(20_000_000)))
def pre_sync(): torch.cuda.synchronize(); 
print("after host sync (cold launch) %.3f ms" % single(pre_sync))
# zero-ish logits as at initialisation of the training step (no lazy rescales, flat softmax)
wb.mul_(0.01)
print("flat logits, back-to-back     %.3f ms" % single(lambda: None))
print("flat logits, after 256 MB fill %.3f ms" % single(lambda: big.fill_(1.0)))
