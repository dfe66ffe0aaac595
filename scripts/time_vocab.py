"""Times the fused vocabulary-stage kernels: python scripts/time_vocab.py M [h] [V].
B4CP_BWD_GROUPS=1 selects the single-group backward epilogue."""
import sys, os, torch, numpy as np
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
M = int(sys.argv[1]); h = int(sys.argv[2]) if len(sys.argv) > 2 else 128
V = int(sys.argv[3]) if len(sys.argv) > 3 else 54293
xb = (torch.randn(M, h, device="cuda") * 0.5).to(torch.bfloat16)
wb = torch.zeros(h, ops.ld8(V), device="cuda", dtype=torch.bfloat16); wb[:, :V] = (torch.randn(h, V, device="cuda") * 0.1).to(torch.bfloat16)
bias = torch.zeros(V, device="cuda"); labels = torch.randint(0, V, (M,), device="cuda", dtype=torch.int32)
lse = torch.empty(M, device="cuda"); tgt = torch.empty(M, device="cuda"); stats = torch.empty(2, device="cuda")
dX = torch.empty(M, h, device="cuda"); dW = torch.empty(h, V, device="cuda"); db = torch.empty(V, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
dxok = h in (128, 256)
tf0 = t(lambda: ops.vocab_ce_fwd(xb, M, h, wb, bias, V, labels, lse, tgt, want_dx=False))
fl = 2.0 * M * h * V
if not dxok:
    print(f"M={M} h={h} V={V}: fwd(no dx) {tf0:.3f} ms ({fl/tf0/1e9:.0f} TF/s)")
    sys.exit(0)
tf = t(lambda: ops.vocab_ce_fwd(xb, M, h, wb, bias, V, labels, lse, tgt, want_dx=True))
ops.ce_loss_reduce(lse, tgt, labels, stats)
tb = t(lambda: ops.vocab_ce_bwd(xb, M, h, wb, bias, V, labels, lse, stats, dW, db))
td = t(lambda: ops.vocab_ce_dx(M, h, V, labels, stats, wb, None, dX, None))
print(f"M={M} h={h} V={V}: fwd(no dx) {tf0:.3f} ms ({fl/tf0/1e9:.0f} TF/s), fwd+U {tf:.3f} ms ({2*fl/tf/1e9:.0f} TF/s of 2 MMAs), dx {td:.3f} ms, bwd {tb:.3f} ms ({2*fl/tb/1e9:.0f} TF/s of 2 MMAs; {fl/tb/1e9:.0f} algorithmic); total {tf+td+tb:.3f} ms = {3*fl/(tf+td+tb)/1e9:.0f} TF/s useful (6MhV)")
