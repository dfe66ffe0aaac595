"""Residual + LayerNorm forward / backward alone against the HBM roofline (same method as
time_hbm_kernels.py: graph replay after an L2 flush, CUDA events, algorithmic bytes / median)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops  # noqa: E402

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=9):
    fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    graph.replay()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, shape, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "shape": shape, "algorithmic_bytes": int(nbytes), "ms": round(ms, 4),
                      "achieved_GBps": round(gbs, 1), "peak_GBps": PEAK, "frac": round(gbs / PEAK, 3)}), flush=True)


for rows_, d in ((4096 * 52, 64), (16384 * 52, 64), (256 * 202, 256), (1024 * 202, 256)):
    x, r, dy = (torch.randn(rows_, d, device="cuda") for _ in range(3))
    gam, bet = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
    y, yb = torch.empty(rows_, d, device="cuda"), torch.empty(rows_, d, device="cuda", dtype=torch.bfloat16)
    dx = torch.empty(rows_, d, device="cuda")
    dg, db_, dbias = (torch.zeros(d, device="cuda") for _ in range(3))
    report("residual_ln_fwd_kernel", f"rows={rows_} d={d}", rows_ * d * 14,
           timed(lambda: ops.residual_ln_fwd(x, r, gam, bet, y, yb)))
    report("residual_ln_bwd_kernel", f"rows={rows_} d={d}", rows_ * d * 18,
           timed(lambda: ops.residual_ln_bwd(dy, x, r, gam, dx, yb, dg, db_, dbias)))
