"""Timing of the on-device Cloze batch builder (b4cp_cloze_build) against the HBM
roofline.  `sets` independent batches are built inside ONE CUDA graph (the kernel runs for a few
microseconds, less than a host launch), replayed after an L2 flush, timed with CUDA events.
Algorithmic bytes per batch: items read (4 B each) + offsets (16 B / row) + session index (4 B /
row) + the ids row (4 (L + 3) B) + the labels row (4 Mmax B).
Prints one JSON line per shape."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from bert4clickpath_b200 import ops  # noqa: E402
from bert4clickpath_b200.data import DeviceClozeBuilder  # noqa: E402
from bert4clickpath_b200.synthetic import BEAUTY_LEN_HIST  # noqa: E402

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def run(name, lens, B, L, Mmax, p, max_masked, sets=16, iters=7):
    rng = np.random.default_rng(0)
    n_sess = len(lens)
    sessions = [rng.integers(10, 54303, size=int(l)).astype(np.int32) for l in lens]
    builder = DeviceClozeBuilder(sessions)
    idx = [torch.from_numpy(rng.permutation(n_sess)[:B].astype(np.int32)).cuda() for _ in range(sets)]
    ids = [torch.empty((B, L + 3), dtype=torch.int32, device="cuda") for _ in range(sets)]
    lab = [torch.empty((B, Mmax), dtype=torch.float32, device="cuda") for _ in range(sets)]
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")

    def launch(j):
        ops.cloze_build(builder.items, builder.offsets, idx[j], B, L, Mmax, True, p, max_masked,
                        1234 + j, (3, 4, 1, 0, 10), -1.0, ids[j], lab[j], cnt, status)

    for j in range(sets):
        launch(j)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for j in range(sets):
            launch(j)
    g.replay()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / sets)
    ms = float(np.median(ts))
    items = float(np.mean([builder.lengths[i.cpu().numpy()].sum() for i in idx]))
    alg = items * 4 + B * 20 + B * (L + 3) * 4 + B * Mmax * 4
    print(json.dumps(dict(kernel="cloze_build_kernel", shape=name, B=B, L=L, Mmax=Mmax,
                          us_per_batch=round(ms * 1e3, 2), alg_bytes=int(alg),
                          gbs=round(alg / ms / 1e6, 1), frac=round(alg / ms / 1e6 / PEAK, 4),
                          sequences_per_s=round(B / ms * 1e3))), flush=True)


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    beauty = rng.choice(np.arange(5, 51), size=100000, p=BEAUTY_LEN_HIST / BEAUTY_LEN_HIST.sum())
    run("C1 beauty-shaped B4096", beauty, 4096, 49, 10, 0.15, 10)
    run("C1 dense B4096", np.full(20000, 50), 4096, 49, 10, 0.15, 10)
    run("C1 dense B32768", np.full(40000, 50), 32768, 49, 10, 0.15, 10, sets=4)
    run("C4 dense B256", np.full(4000, 200), 256, 199, 30, 0.15, 30)
    run("C4 dense B8192", np.full(10000, 200), 8192, 199, 30, 0.15, 30, sets=4)
