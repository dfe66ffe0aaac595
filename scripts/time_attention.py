"""Attention fwd/bwd timing at the C4 shape: python scripts/time_attention.py [B S H dh]"""
import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
B, S, H, dh = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (256, 202, 4, 64)
d = H * dh
qkv = (torch.randn(B * S, 3 * d, device="cuda") * 0.5).to(torch.bfloat16)
do = (torch.randn(B * S, d, device="cuda") * 0.5).to(torch.bfloat16)
ids = torch.randint(10, 1000, (B * S,), device="cuda", dtype=torch.int32)
out = torch.empty(B * S, d, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device="cuda")
dqkv = torch.empty(B * S, 3 * d, device="cuda", dtype=torch.bfloat16)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tf = t(lambda: ops.attention_fwd(qkv, ids, B, S, H, dh, out, lse))
tb = t(lambda: ops.attention_bwd(qkv, do, lse, ids, B, S, H, dh, dqkv, out=out))
tbs = t(lambda: ops.attention_bwd(qkv, do, lse, ids, B, S, H, dh, dqkv))
fl = 4.0 * B * H * S * S * dh
by_f = B * S * d * 2 * 4.0; by_b = B * S * d * 2 * 7.0
print(f"  HBM bytes (algorithmic): fwd {by_f/1e6:.0f} MB -> {by_f/tf/1e6:.0f} GB/s, bwd {by_b/1e6:.0f} MB -> {by_b/tb/1e6:.0f} GB/s")
print(f"B={B} S={S} H={H} dh={dh}: fwd {tf:.3f} ms ({fl/tf/1e9:.1f} TF/s), bwd(tensor) {tb:.3f} ms ({2.5*fl/tb/1e9:.1f} TF/s), bwd(out=None -> {'SIMT' if S > 128 else 'tensor'}) {tbs:.3f} ms")
