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


def sigmoid_heads():
    """Loss and every gradient of the two sigmoid heads under MaskedLoss(binary_crossentropy,
    pos_weight) (head.py:4-26, :50-69; losses.py:31-98)."""
    rng = np.random.default_rng(21)
    B, Ls, d, V = 6, 3, 8, 11
    layers = [(rng.normal(size=(d, 6)) * 0.5, rng.normal(size=6) * 0.1)]
    xb = rng.normal(size=(B, Ls, d))
    yb = rng.integers(0, 2, size=(B, Ls)).astype(np.float64)
    yb[rng.random((B, Ls)) < 0.25] = -1.0
    wb, bb = rng.normal(size=(6, 1)) * 0.5, rng.normal(size=1) * 0.1
    lb = O.binary_head_loss_and_grads(xb, layers, wb, bb, yb, pos_weight=3.0)
    xm = rng.normal(size=(B, 1, d))
    ym = (rng.random((B, V)) < 0.2).astype(np.float64)
    ym[rng.random((B, V)) < 0.15] = -1.0
    wm, bm = rng.normal(size=(6, V)) * 0.5, rng.normal(size=V) * 0.1
    lm = O.multilabel_head_loss_and_grads(xm, layers, wm, bm, ym, pos_weight=None)
    np.savez_compressed(
        os.path.join(OUT, "sigmoid_heads.npz"), w0=layers[0][0], b0=layers[0][1],
        xb=xb, yb=yb, wb=wb, bb=bb, b_loss=np.float64(lb[0]), b_dx=lb[1], b_dw0=lb[2][0][0],
        b_db0=lb[2][0][1], b_dwo=lb[3], b_dbo=lb[4],
        xm=xm, ym=ym, wm=wm, bm=bm, m_loss=np.float64(lm[0]), m_dx=lm[1], m_dw0=lm[2][0][0],
        m_db0=lm[2][0][1], m_dwo=lm[3], m_dbo=lm[4])


def keyed_cloze():
    """Cloze batches with keyed mask positions (input_pipeline.py:59-133 + the position key of
    include/b4cp.h): what b4cp_cloze_build must reproduce bit for bit."""
    rng = np.random.default_rng(31)
    lens = [5, 1, 12, 50, 7, 2, 33, 9]
    sessions = [rng.integers(10, 5000, size=n).astype(np.int32) for n in lens]
    idx = np.array([3, 0, 5, 1, 2, 7, 6], dtype=np.int32)
    tr = O.keyed_cloze_batch(sessions, idx, "train", 1234, 0.4, 10)
    ev = O.keyed_cloze_batch(sessions, idx, "eval", 1234, 0.4, 10, L=52, Mmax=3)
    flat = np.concatenate(sessions)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    keys = np.array([O.splitmix64((O.splitmix64(1234 + 3) + i) & (2 ** 64 - 1)) for i in range(4)],
                    dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, "keyed_cloze.npz"), items=flat, offsets=offs, idx=idx,
                        train_ids=tr[0], train_labels=tr[1], train_n=np.int64(tr[2]),
                        eval_ids=ev[0], eval_labels=ev[1], eval_n=np.int64(ev[2]), keys_s3=keys)


if __name__ == "__main__":
    sigmoid_heads()
    keyed_cloze()
    tiny_cloze("cloze_tiny_1feat.npz", (8,), False)
    tiny_cloze("cloze_tiny_2feat_dropout.npz", (8, 8), True)
    embed_and_topk()
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
