"""C5: next-item top-100 over a 1M-item catalogue, h = 256, batch sweep (scoring + top-k only).
python scripts/time_c5_topk.py"""
import json, sys, torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops
V = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
h = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = 100
wb = torch.zeros(h, ops.ld8(V), device="cuda", dtype=torch.bfloat16)
wb[:, :V] = (torch.randn(h, V, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.zeros(V, device="cuda")
for B in ((1, 16, 256, 1024, 4096, 16384) if len(sys.argv) < 4 else tuple(int(a) for a in sys.argv[3:])):
    xb = (torch.randn(B, h, device="cuda") * 0.5).to(torch.bfloat16)
    ids = torch.empty(B, k, dtype=torch.int32, device="cuda")
    fn = lambda: ops.score_topk(xb, B, h, wb, bias, V, k, out_ids=ids)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * h * V
    # materialised alternative: logits for <= 2048 rows at a time + streaming top-k
    RC = min(B, 2048)
    z = torch.empty(RC, ops.ld8(V), device="cuda")
    def mat():
        for a in range(0, B, RC):
            rows = min(RC, B - a)
            ops.gemm(xb[a:a + rows], 0, wb, 1, rows, V, h, bias=bias, out_f32=z[:rows])
            ops.topk_rows(z[:rows], V, k, out_ids=ids[a:a + rows])
    for _ in range(2): mat()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n): mat()
    e1.record(); torch.cuda.synchronize()
    ms_mat = e0.elapsed_time(e1) / n
    del z
    print(json.dumps({"B": B, "ms_materialised": round(ms_mat, 3), "queries_per_sec_materialised": round(B / ms_mat * 1e3), "V": V, "h": h, "k": k, "ms": round(ms, 3), "queries_per_sec": round(B / ms * 1e3),
                      "TFLOPs": round(fl / ms / 1e9, 1), "W_stream_GBps": round(V * h * 2 / ms / 1e6, 1)}), flush=True)
