"""Embedding backward (segment sums over pre-sorted ids) alone: the roofline shapes of
time_hbm_kernels.py (Zipf ids) and the batch a C1 training step actually sees (4,096 Cloze rows
with their [CLS] / [SEP] / [MASK] / [PAD] tokens, input dropout on).  Graph replay after an L2
flush, CUDA events, algorithmic bytes = gradient rows + ids read, unique rows written."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops  # noqa: E402
from bert4clickpath_b200.synthetic import make_cloze_batch  # noqa: E402

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


def run(name, ids, rows, d, rate):
    n = ids.numel()
    dout = torch.randn(n, d, device="cuda")
    tg = torch.empty(rows, d, device="cuda")
    ops.embed_sort(ids, rows, d, "t_sort")
    ms = timed(lambda: ops.embed_bwd_sorted(dout, d, 0, d, n, rows, tg, "t_sort", dropout_rate=rate, seed=7, site=1))
    U = int(torch.unique(ids).numel())
    alg = n * d * 4 + n * 4 + U * d * 4 + U * 4
    print(json.dumps({"kernel": "embed_bwd_sorted", "shape": f"{name}: N={n} U={U} d={d} rows={rows} dropout={rate}",
                      "ms": round(ms, 4), "achieved_GBps": round(alg / ms / 1e6, 1),
                      "frac": round(alg / ms / 1e6 / PEAK, 3),
                      "frac_incl_zero_fill": round((alg + rows * d * 4) / ms / 1e6 / PEAK, 3),
                      "checksum": float(tg.double().sum().item())}), flush=True)


g = np.random.default_rng(0)
b = make_cloze_batch(g, 4096, 54293, 50, "train", 0.15, 10)
run("C1 training batch", torch.from_numpy(b["ids"]).cuda().view(-1), 54293 + 11, 64, 0.1)
for name, B, S, V, d in (("C1 zipf", 16384, 52, 54293, 64), ("C4 zipf", 1024, 202, 1_000_000, 256)):
    p = 1.0 / np.arange(1, V + 1) ** 0.8
    p /= p.sum()
    ids = torch.from_numpy(g.choice(V, size=B * S, p=p).astype(np.int32) + 10).cuda()
    run(name, ids, V + 11, d, 0.0)
