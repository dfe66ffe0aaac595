"""Bias-gradient column sums at the bench step's shapes: python scripts/time_colsum.py
(B4CP_COLSUM_FIXED=1 selects the fixed 32-group x 512-row thread shape)."""
import sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
flush = torch.empty(64 << 20, device="cuda")
def t(fn, n=20):
    """graph replay (no host launch cost), L2 flushed before every call"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            flush.fill_(0.0)
            fn()
    g0 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g0):
        for _ in range(n):
            flush.fill_(0.0)
    def run(gr):
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    return (run(g) - run(g0)) / n * 1e3
tot = 0.0
for T, n in ((212992, 192), (212992, 100), (28672, 1024), (28672, 512), (28672, 256), (28672, 128)):
    x = torch.randn(T, ops.ld8(n), device="cuda").to(torch.bfloat16)
    out = torch.empty(n, device="cuda")
    us = t(lambda: ops.colsum_bf16(x, T, n, out))
    ref = x[:, :n].float().sum(0)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    tot += us
    print(f"T={T} n={n}: {us:.1f} us ({T*n*2/us/1e3:.0f} GB/s), rel err {err:.1e}")
print(f"total {tot:.0f} us")
