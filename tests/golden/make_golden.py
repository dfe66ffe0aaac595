"""Generates tests/golden/*.npz from the CPU oracle (seeded).  The reference itself cannot run
here (TensorFlow 2.3.1 is not installable offline), so these vectors pin the ORACLE's outputs:
any later edit of the oracle or of the kernels that changes results shows up as a diff against
them.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import clickpath_oracle as O  # noqa: E402
from tests.test_oracle import make_tiny_problem  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def tiny_cloze(name, dims, with_dropout):
    ids_list, labels, P, L, H, pe, masks = make_tiny_problem(seed=11, dims=dims, dff=12,
                                                             head=(16, 8), with_dropout=with_dropout)
    loss, G, ex = O.cloze_train_step(ids_list, labels, P, L, H, pe, np.float64, masks)
    x, _ = O.encoder_fwd(ids_list, P, L, H, pe, np.float64, masks)
    sel, _ = O.select_masked(ids_list[0], x)
    probs, _, _ = O.softmax_head_fwd(sel, O.head_layers(P), P["head.out.w"], P["head.out.b"])
    nd = O.cloze_ndcg_update(labels, probs.astype(np.float32), 5)
    rc = O.cloze_recall_update(labels, probs.astype(np.float32), 5)
    arrs = {f"ids{f}": a for f, a in enumerate(ids_list)}
    arrs.update({f"P/{k}": v for k, v in P.items()})
    arrs.update({f"G/{k}": v for k, v in G.items()})
    if masks:
        arrs.update({f"mask/{k if isinstance(k, str) else f'{k[0]}_{k[1]}'}": v for k, v in masks.items()})
    arrs.update(labels=labels, loss=np.float64(loss), enc_out=x, probs=probs,
                ndcg5=np.array(nd, dtype=np.float32), recall5=np.array(rc, dtype=np.float32),
                meta=np.array([L, H], dtype=np.int64))
    np.savez_compressed(os.path.join(OUT, name), **arrs)


def embed_and_topk():
    rng = np.random.default_rng(5)
    dims, rows = (20, 4), (111, 21)
    tables = [rng.uniform(-0.05, 0.05, size=(r, d)).astype(np.float32) for r, d in zip(rows, dims)]
    ids = [rng.integers(0, r, size=(3, 9)).astype(np.int32) for r in rows]
    pe = O.positional_encoding(10000, sum(dims))
    out = O.embed_fwd(ids, tables, pe, np.float32)
    scores = rng.normal(size=(4, 500)).astype(np.float32)
    scores[1] = np.round(scores[1] * 4) / 4
    scores[2] = 0.5
    np.savez_compressed(os.path.join(OUT, "embed_topk.npz"), t0=tables[0], t1=tables[1],
                        ids0=ids[0], ids1=ids[1], out=out, pe_rows=pe[:9], scores=scores,
                        top10=O.top_k_ids(scores, 10))


if __name__ == "__main__":
    tiny_cloze("cloze_tiny_1feat.npz", (8,), False)
    tiny_cloze("cloze_tiny_2feat_dropout.npz", (8, 8), True)
    embed_and_topk()
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
