"""Vocabulary-parallel output stage vs the replicated one on the same weights and batches.

Run with one process per GPU:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 scripts/check_vocab_parallel.py [--vocab 54293] [--batch 512]
Each rank draws its own batch.  Model A replicates the output kernel (data parallel: gradients
all-reduced); model B shards it by vocabulary ranges.  Checked: identical global loss, gradients of
every replicated parameter, the shard's dW/db against the matching slice of A's, top-k ids, and the
loss trajectory over a few Adam steps.  Prints one JSON line on rank 0; exit code 1 on mismatch.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert4clickpath_b200 import Adam, ClickstreamTransformer, SoftMaxHead, INPUT_MASKING_TOKEN  # noqa: E402
from bert4clickpath_b200.synthetic import make_cloze_batch  # noqa: E402
from bert4clickpath_b200.training import ClozeTrainStep  # noqa: E402


def build(V, vocab_parallel, seed=7):
    head = SoftMaxHead(dense_layer_dims=[1024, 512, 256, 128], output_vocab_size=V,
                       vocab_parallel=vocab_parallel)
    m = ClickstreamTransformer(sequential_input_config={'items': ['asin']},
                               feature_vocabs={'items': V}, embedding_dims={'items': 64},
                               head_unit=head, value_to_head=INPUT_MASKING_TOKEN,
                               num_encoder_layers=2, num_attention_heads=2, dropout_rate=0.1,
                               seed=seed)
    m.compile(optimizer=Adam(1e-3, 0.9, 0.999, 1e-9))
    return m


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vocab", type=int, default=54293)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    V = args.vocab
    A, Bm = build(V, False), build(V, True)
    vp = Bm.head.vocab
    # identical weights: B's shard must be A's slice (same rng draw)
    wa = A.store.get_weights()
    wb = Bm.store.get_weights()
    name_w, name_b = f"{Bm.head.prefix}.out.w", f"{Bm.head.prefix}.out.b"
    assert np.array_equal(wa[name_w][:, vp.v_begin:vp.v_end], wb[name_w])
    for n in wa:
        if n not in (name_w, name_b):
            assert np.array_equal(wa[n], wb[n]), n

    rng = np.random.default_rng(100 + rank)
    batch = make_cloze_batch(rng, args.batch, V, max_len=50, mode="train")
    ta, tb = ClozeTrainStep(A, A.optimizer), ClozeTrainStep(Bm, Bm.optimizer)
    da, db_ = ta.to_device(batch), tb.to_device(batch)
    out = {"world": world, "vocab": V, "batch_per_rank": args.batch, "shard": [vp.v_begin, vp.v_end]}
    ok = True

    # ---- one forward/backward without dropout: loss + gradients
    sa = A.cloze_forward_backward(da.ids, da.labels, da.B, da.S, n_masked=da.n_masked, training=False)
    sb = Bm.cloze_forward_backward(db_.ids, db_.labels, db_.B, db_.S, n_masked=db_.n_masked, training=False)
    sa, sb = sa.cpu().numpy(), sb.cpu().numpy()
    out["loss_replicated"], out["loss_vocab_parallel"] = float(sa[0] / sa[1]), float(sb[0] / sb[1])
    ok &= sa[1] == sb[1] and abs(sa[0] - sb[0]) <= 2e-5 * abs(sa[0])
    ga, gb = A.store.get_grads(), Bm.store.get_grads()
    # parameters whose gradient is analytically zero (attention key biases) hold rounding noise
    # only: measure every error against at least 1e-3 of the largest gradient norm
    floor = 1e-3 * max(np.linalg.norm(ga[n]) for n in ga)
    worst, worst_name = 0.0, None
    for n in ga:
        if n in (name_w, name_b):
            continue
        e = float(np.linalg.norm(gb[n].astype(np.float64) - ga[n]) / max(np.linalg.norm(ga[n]), floor))
        if e > worst:
            worst, worst_name = e, n
    if os.environ.get("VP_VERBOSE"):
        for n in ga:
            if n not in (name_w, name_b):
                print(f"[rank {rank}] {n:28s} |A|={np.linalg.norm(ga[n]):.3e} |B|={np.linalg.norm(gb[n]):.3e} "
                      f"|B-A|={np.linalg.norm(gb[n].astype(np.float64) - ga[n]):.3e}", flush=True)
    out["grad_rel_replicated_params_max"] = worst
    out["grad_rel_replicated_params_argmax"] = worst_name
    out["grad_rel_out_w_shard"] = rel(gb[name_w], ga[name_w][:, vp.v_begin:vp.v_end])
    out["grad_rel_out_b_shard"] = rel(gb[name_b], ga[name_b][vp.v_begin:vp.v_end])
    ok &= worst < 5e-3 and out["grad_rel_out_w_shard"] < 5e-3 and out["grad_rel_out_b_shard"] < 5e-3

    # ---- top-k of the inference path
    tbatch = make_cloze_batch(rng, args.batch, V, max_len=50, mode="test")
    ids = torch.from_numpy(np.ascontiguousarray(tbatch["ids"])).cuda().view(-1)
    Bq, S = tbatch["ids"].shape
    ia, _ = A.topk_ids([ids], Bq, S, args.k, n_masked=Bq)
    ib, _ = Bm.topk_ids([ids], Bq, S, args.k, n_masked=Bq)
    ia, ib = ia.cpu().numpy(), ib.cpu().numpy()[:Bq]
    out["topk_mismatch_fraction"] = float((ia != ib).mean())
    ok &= out["topk_mismatch_fraction"] == 0.0

    # ---- short training run (dropout on, same seeds): loss trajectories
    la, lb = [], []
    for _ in range(args.steps):
        s1 = ta.step_device(da).cpu().numpy()
        s2 = tb.step_device(db_).cpu().numpy()
        la.append(float(s1[0] / s1[1]))
        lb.append(float(s2[0] / s2[1]))
    out["losses_replicated"], out["losses_vocab_parallel"] = la, lb
    ok &= max(abs(x - y) for x, y in zip(la, lb)) < 2e-2

    # ---- timing of the vocabulary-parallel step vs the replicated one
    def timed(t, d, n=10):
        for _ in range(3):
            t.step_device(d)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            t.step_device(d)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())
    out["ms_per_step_replicated"] = timed(ta, da)
    out["ms_per_step_vocab_parallel"] = timed(tb, db_)
    out["ok"] = bool(ok)
    flags = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["ok_all_ranks"] = bool(flags.item())
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    sys.exit(0 if flags.item() else 1)


if __name__ == "__main__":
    main()
