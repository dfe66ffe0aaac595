"""Data-parallel Cloze step against ONE process on the concatenated batch (SURVEY.md Appendix D:
"1/2/4/8 ranks give the same loss / gradients as a single-GPU run on the concatenated batch").

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29513 scripts/check_data_parallel.py

Every rank draws its own batch with a DIFFERENT number of [MASK] rows (so the global masked mean
of losses.py:80-91 differs from the mean of the ranks' means), runs the data-parallel step
(valid-row count all-reduced under the forward, output-kernel gradient all-reduced under the
backward, the rest after it), and then - with collectives disabled - the same model on the
concatenation of all ranks' batches.  Case 2 uses a table large enough for the row exchange of
table gradients (EncoderEngine.plan_table_exchange) and, with dropout on, checks the row exchange
against the dense all-reduce of the same step.  Prints one JSON line on rank 0; exit 1 on mismatch.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bert4clickpath_b200 import ClickstreamTransformer, SoftMaxHead, INPUT_MASKING_TOKEN  # noqa: E402
from bert4clickpath_b200.engine import EncoderEngine  # noqa: E402
from bert4clickpath_b200.synthetic import make_cloze_batch  # noqa: E402


def build(V, d, hd, dropout, seed=7):
    head = SoftMaxHead(dense_layer_dims=hd, output_vocab_size=V)
    return ClickstreamTransformer(sequential_input_config={'items': ['asin']},
                                  feature_vocabs={'items': V}, embedding_dims={'items': d},
                                  head_unit=head, value_to_head=INPUT_MASKING_TOKEN,
                                  num_encoder_layers=2, num_attention_heads=2, dropout_rate=dropout,
                                  seed=seed)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def batches(world, B, V):
    out = []
    for r in range(world):
        rng = np.random.default_rng(100 + r)
        out.append(make_cloze_batch(rng, B, V, max_len=50, mode="train",
                                    masked_percentage=0.15 if r % 2 == 0 else 0.4, max_masked=10))
    return out


def dev(batch):
    ids = torch.from_numpy(np.ascontiguousarray(batch["ids"])).cuda().view(-1)
    labels = torch.from_numpy(np.ascontiguousarray(batch["labels"])).cuda()
    return ids, labels


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    solo = None
    for r in range(world):                      # one singleton group per rank: "no data parallelism"
        g = dist.new_group([r])
        if r == rank:
            solo = g
    out, ok = {"world": world}, True
    for name, V, d, hd, B in (("c1", 54293, 64, [1024, 512, 256, 128], 256),
                              ("big_table", 400_000, 64, [128], 64)):
        bs = batches(world, B, V)
        S = bs[0]["ids"].shape[1]
        mmax = max(b["labels"].shape[1] for b in bs)
        cat_ids = np.concatenate([b["ids"] for b in bs], axis=0)
        cat_lab = np.full((world * B, mmax), -1.0, np.float32)
        for r, b in enumerate(bs):
            cat_lab[r * B:(r + 1) * B, :b["labels"].shape[1]] = b["labels"]
        n_cat = sum(b["n_masked"] for b in bs)
        # ---- data parallel, no dropout
        A = build(V, d, hd, 0.0)
        ids, lab = dev(bs[rank])
        sa = A.cloze_forward_backward([ids], lab, B, S, n_masked=bs[rank]["n_masked"], training=False).cpu().numpy()
        ga = A.store.get_grads()
        exchanged = [p.name for p in A.store.params.values() if p.grad_is_global]
        # ---- one process, concatenated batch, collectives off
        R = build(V, d, hd, 0.0)
        R.set_process_group(solo)
        ids_c, lab_c = dev(dict(ids=cat_ids, labels=cat_lab))
        sr = R.cloze_forward_backward([ids_c], lab_c, world * B, S, n_masked=n_cat, training=False).cpu().numpy()
        gr = R.store.get_grads()
        errs = {k: rel(ga[k], gr[k]) for k in ga if not k.endswith("bqkv")}   # (key-bias slice ~ 0: noise)
        worst = max(errs.items(), key=lambda kv: kv[1])
        res = {"loss_dp": float(sa[0] / sa[1]), "loss_single": float(sr[0] / sr[1]), "n_valid": [float(sa[1]), float(sr[1])],
               "worst_grad": worst, "tables_by_row_exchange": exchanged}
        ok &= sa[1] == sr[1] == n_cat and abs(sa[0] / sa[1] - sr[0] / sr[1]) < 2e-5 * abs(sr[0] / sr[1]) and worst[1] < 2e-3
        if name == "big_table":
            assert exchanged == ["emb.0"], exchanged
            # ---- dropout on: row exchange vs dense all-reduce of the same step (same seeds / masks)
            D1, D2 = build(V, d, hd, 0.1), build(V, d, hd, 0.1)
            s1 = D1.cloze_forward_backward([ids], lab, B, S, n_masked=bs[rank]["n_masked"], training=True, seed=11).cpu().numpy()
            ratio, EncoderEngine.ROW_EXCHANGE_RATIO = EncoderEngine.ROW_EXCHANGE_RATIO, 1e30   # never exchange
            s2 = D2.cloze_forward_backward([ids], lab, B, S, n_masked=bs[rank]["n_masked"], training=True, seed=11).cpu().numpy()
            EncoderEngine.ROW_EXCHANGE_RATIO = ratio
            assert not any(p.grad_is_global for p in D2.store.params.values())
            g1, g2 = D1.store.get_grads(), D2.store.get_grads()
            e = rel(g1["emb.0"], g2["emb.0"])
            res["dropout_row_exchange_vs_dense_allreduce"] = {"emb.0": e, "loss": [float(s1[0] / s1[1]), float(s2[0] / s2[1])]}
            ok &= e < 1e-5 and abs(s1[0] - s2[0]) < 1e-5 * abs(s2[0])
        out[name] = res
        del A, R
        torch.cuda.empty_cache()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


main()
