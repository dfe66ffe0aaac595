"""Top-k over materialised score rows against the HBM roofline (same method as
time_hbm_kernels.py: graph replay after an L2 flush, CUDA events, algorithmic bytes = the scores
read once + ids written).  B4CP_TOPK_SMALL=0 selects the 512-thread sampled kernel for rows of
16K..128K scores (the previous default) for an A/B in a second process."""
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


for B, V, k in ((9472, 54293, 100), (9472, 54293, 10), (4096, 54293, 100), (512, 54293, 100),
                (8192, 16384, 100), (4736, 100000, 100), (1184, 1000000, 100)):
    sc = torch.randn(B, ops.ld8(V), device="cuda")
    ids_o = torch.empty(B, k, dtype=torch.int32, device="cuda")
    ms = timed(lambda: ops.topk_rows(sc, V, k, out_ids=ids_o))
    nbytes = B * V * 4 + B * k * 4
    gbs = nbytes / (ms * 1e-3) / 1e9
    redo = int((ids_o[:, 0] < 0).sum().item())
    print(json.dumps({"kernel": "topk_rows", "variant": os.environ.get("B4CP_TOPK_SMALL", "1"),
                      "shape": f"B={B} V={V} k={k}", "ms": round(ms, 4), "achieved_GBps": round(gbs, 1),
                      "frac": round(gbs / PEAK, 3), "unfinished_rows": redo}), flush=True)
    del sc
