"""CPU-only checks (`-m "not gpu"`): the oracle against the committed golden vectors, the C-ABI
library's exported symbols, host-side input preparation, synthetic data rules, and the
data-parallel loss/gradient normalisation on a 2-rank gloo group."""
import ctypes
import os
import socket

import numpy as np
import pytest

from oracle import clickpath_oracle as O
from oracle.mixed_precision import bf16, cloze_train_step_bf16

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name))
    P = {k[2:]: z[k] for k in z.files if k.startswith("P/")}
    G = {k[2:]: z[k] for k in z.files if k.startswith("G/")}
    masks = None
    mk = [k for k in z.files if k.startswith("mask/")]
    if mk:
        masks = {}
        for k in mk:
            n = k[5:]
            masks["in" if n == "in" else tuple(int(t) for t in n.split("_"))] = z[k]
    ids = [z[k] for k in sorted(f for f in z.files if f.startswith("ids"))]
    return z, ids, P, G, masks


@pytest.mark.parametrize("name", ["cloze_tiny_1feat.npz", "cloze_tiny_2feat_dropout.npz"])
def test_oracle_reproduces_golden_cloze_step(name):
    z, ids, P, G, masks = _load(name)
    L, H = (int(v) for v in z["meta"])
    pe = O.positional_encoding(10000, sum(P[f"emb.{f}"].shape[1] for f in range(len(ids))))
    loss, G2, ex = O.cloze_train_step(ids, z["labels"], P, L, H, pe, np.float64, masks)
    assert abs(loss - float(z["loss"])) < 1e-12
    np.testing.assert_allclose(ex["enc_out"], z["enc_out"], rtol=1e-12, atol=1e-14)
    for k in G:
        np.testing.assert_allclose(G2[k], G[k], rtol=1e-10, atol=1e-14, err_msg=k)
    x, _ = O.encoder_fwd(ids, P, L, H, pe, np.float64, masks)
    sel, _ = O.select_masked(ids[0], x)
    probs, _, _ = O.softmax_head_fwd(sel, O.head_layers(P), P["head.out.w"], P["head.out.b"])
    np.testing.assert_allclose(probs, z["probs"], rtol=1e-12)
    assert O.cloze_ndcg_update(z["labels"], probs.astype(np.float32), 5) == tuple(z["ndcg5"])
    assert O.cloze_recall_update(z["labels"], probs.astype(np.float32), 5) == tuple(z["recall5"])
    # the bf16-emulating variant stays within the stated bf16 tolerance of the exact oracle
    eloss, EG, _ = cloze_train_step_bf16(ids, z["labels"], P, L, H, pe, masks)
    assert abs(eloss - loss) < 2e-2 * abs(loss)


def test_oracle_reproduces_golden_embed_and_topk():
    z = np.load(os.path.join(GOLD, "embed_topk.npz"))
    pe = O.positional_encoding(10000, 24)
    assert pe[:9].tobytes() == z["pe_rows"].tobytes()
    out = O.embed_fwd([z["ids0"], z["ids1"]], [z["t0"], z["t1"]], pe, np.float32)
    assert out.tobytes() == z["out"].tobytes()  # bit-exact gather + scale + PE
    assert O.top_k_ids(z["scores"], 10).tolist() == z["top10"].tolist()
    assert z["top10"][2].tolist() == list(range(10))  # all-ties row: lowest ids first


def test_bf16_rounding_helper_matches_torch():
    import torch
    a = np.random.default_rng(0).normal(size=1000).astype(np.float32) * 37
    want = torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    np.testing.assert_array_equal(bf16(a).astype(np.float32), want)


# --------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    from bert4clickpath_b200 import _lib
    from bert4clickpath_b200.build import build_lib
    build_lib(verbose=False)
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _lib.declared_symbols()
    assert len(names) >= 38
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    L.b4cp_version.restype = ctypes.c_int
    assert L.b4cp_version() == 1
    L.b4cp_last_error.restype = ctypes.c_char_p
    assert isinstance(L.b4cp_last_error(), bytes)
    # pure host-side queries work without a GPU
    L.b4cp_embed_bwd_workspace_bytes.restype = ctypes.c_long
    assert L.b4cp_embed_bwd_workspace_bytes(ctypes.c_long(1000), ctypes.c_int(64)) > 0
    L.b4cp_vocab_ce_workspace_bytes.restype = ctypes.c_long
    assert L.b4cp_vocab_ce_workspace_bytes(ctypes.c_long(7168), ctypes.c_int(54293), ctypes.c_int(128)) > 0
    assert L.b4cp_gemm_splits_for(64, 100, 200000) > 1


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import bert4clickpath_b200 as bc
    with pytest.raises(Exception):
        bc.ClickstreamTransformer(
            sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": 50},
            embedding_dims={"items": 8}, head_unit=bc.SoftMaxHead([8], 50),
            value_to_head=bc.INPUT_MASKING_TOKEN)


# --------------------------------------------------------------------------- host logic
def test_input_prep_chains_strings_and_ids_like_the_reference():
    from bert4clickpath_b200.clickstream_transformer import StaticVocabularyTable, TransformerInputPrep
    from bert4clickpath_b200.constants import RESERVED_TOKENS
    prep = TransformerInputPrep({"items": ["s_items", "b_items"], "events": ["s_ev", "b_ev"]})
    feats = {
        "s_items": np.array([["a", "b", "[PAD]"], ["c", "[MASK]", "zzz"]], dtype=object),
        "b_items": np.array([["d"], ["[PAD]"]], dtype=object),
        "s_ev": np.array([["v", "v", "[PAD]"], ["v", "w", "v"]], dtype=object),
        "b_ev": np.array([["w"], ["[PAD]"]], dtype=object),
        "instance": np.array([1, 2]),
    }
    out, starts, ends = prep(feats)
    assert set(out) == {"items", "events", "instance"}
    assert out["items"][0].tolist() == ["[CLS]", "[SEP]", "a", "b", "[PAD]", "[SEP]", "d", "[SEP]"]
    assert starts.tolist() == [0, 2, 6] and ends.tolist() == [1, 5, 7]
    table = StaticVocabularyTable(RESERVED_TOKENS + ["a", "b", "c", "d"])
    assert table.size() == 15
    ids = table.lookup(out["items"])
    assert ids[1].tolist() == [3, 4, 12, 1, 14, 4, 0, 4]  # zzz -> OOV bucket 14, [MASK] -> 1
    want = O.chain_sequences([O.lookup_ids(feats["s_items"], ["a", "b", "c", "d"]),
                              O.lookup_ids(feats["b_items"], ["a", "b", "c", "d"])])
    assert ids.tolist() == want.tolist()


def test_model_constructor_contract():
    import bert4clickpath_b200 as bc
    with pytest.raises(AssertionError):
        bc.ClickstreamTransformer({"items": ["a"]}, {"items": 10}, {"items": 8}, bc.SoftMaxHead([], 10))
    with pytest.raises(AssertionError):
        bc.ClickstreamTransformer({"items": ["a"]}, {"items": 10}, {"items": 8}, bc.SoftMaxHead([], 10),
                                  segment_to_head=0, value_to_head="[MASK]")
    with pytest.raises(AssertionError):
        bc.MaskedLoss(bc.binary_crossentropy, label_pad=1.0)
    assert bc.ClozeMaskedNDCG(5).name == "NDCG_at_5" and bc.ClozeMaskedRecall(10).name == "Recall_at_10"


def test_synthetic_cloze_batches_follow_the_reference_rules():
    from bert4clickpath_b200.synthetic import make_cloze_batch, n_masked_for
    assert n_masked_for(49, 0.15, 10) == 7 and n_masked_for(49, 0.4, 10) == 10
    rng = np.random.default_rng(0)
    b = make_cloze_batch(rng, 64, 1000, max_len=50, mode="train", masked_percentage=0.15)
    assert b["ids"].shape == (64, 52) and b["labels"].shape == (64, 7) and b["n_masked"] == 64 * 7
    assert (b["ids"][:, 0] == 3).all() and (b["ids"][:, 1] == 4).all() and (b["ids"][:, -1] == 4).all()
    assert ((b["ids"] == 1).sum(1) == 7).all()
    r = make_cloze_batch(rng, 200, 1000, max_len=50, mode="train", masked_percentage=0.4, lengths="beauty")
    nmask = (r["ids"] == 1).sum(1)
    assert ((r["labels"] >= 0).sum(1) == nmask).all() and nmask.max() <= 10
    assert (r["ids"] == 0).any()  # interior pads before the trailing [SEP]
    assert (r["ids"][np.arange(200), -1] == 4).all()
    e = make_cloze_batch(rng, 32, 1000, max_len=50, mode="eval", lengths="beauty")
    assert ((e["ids"] == 1).sum(1) == 1).all() and e["labels"].shape == (32, 1)


# --------------------------------------------------------------------------- data parallel
def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.test_oracle import make_tiny_problem
    ids_list, labels, P, L, H, pe, _ = make_tiny_problem(seed=3, B=6)
    sl = slice(rank * 3, rank * 3 + 3)  # uneven numbers of masked positions per rank
    ids_r, lab_r = [i[sl] for i in ids_list], labels[sl]
    # local statistics, then the exchange ClickstreamTransformer.cloze_forward_backward performs:
    # all-reduce (loss_sum, n_valid) BEFORE the backward so gradients are of the GLOBAL masked mean
    loss_r, G_r, ex = O.cloze_train_step(ids_r, lab_r, P, L, H, pe, np.float64)
    n_r = ex["n_valid"]
    stats = torch.tensor([loss_r * n_r, float(n_r)], dtype=torch.float64)
    dist.all_reduce(stats)
    n_glob = stats[1].item()
    flat = torch.from_numpy(np.concatenate([(G_r[k] * n_r / n_glob).reshape(-1) for k in sorted(G_r)]))
    dist.all_reduce(flat)
    if rank == 0:
        q.put((stats[0].item() / n_glob, flat.numpy()))
    dist.destroy_process_group()


def test_two_rank_gloo_data_parallel_equals_single_process():
    import torch.multiprocessing as mp
    from tests.test_oracle import make_tiny_problem
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss_dp, flat_dp = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids_list, labels, P, L, H, pe, _ = make_tiny_problem(seed=3, B=6)
    loss, G, _ = O.cloze_train_step(ids_list, labels, P, L, H, pe, np.float64)
    flat = np.concatenate([G[k].reshape(-1) for k in sorted(G)])
    assert abs(loss_dp - loss) < 1e-12
    np.testing.assert_allclose(flat_dp, flat, rtol=1e-9, atol=1e-13)


# --------------------------------------------------------------------------- vocabulary parallel
def _vp_problem(world, M=7, h=16, V=53, seed=5):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(world, M, h))
    W = rng.normal(size=(h, V)) * 0.4
    b = rng.normal(size=V) * 0.2
    labels = rng.integers(0, V, size=(world, M)).astype(np.int32)
    labels[0, 2] = -1
    labels[world - 1, 0] = -1
    return X, W, b, labels


def _vp_worker(rank, world, port, q):
    """The exchange schedule of engine.VocabParallelOutputEngine with NumPy arithmetic."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X, W, b, labels = _vp_problem(world)
    M, h = X.shape[1:]
    V = W.shape[1]
    per = -(-(-(-V // world)) // 8) * 8
    v0, v1 = min(rank * per, V), min(rank * per + per, V)
    Ws, bs = W[:, v0:v1], b[v0:v1]
    # all-gather rows and labels
    x_all = [torch.empty(M, h, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(x_all, torch.from_numpy(X[rank]))
    l_all = [torch.empty(M, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(l_all, torch.from_numpy(labels[rank]))
    xa, la = torch.cat(x_all).numpy(), torch.cat(l_all).numpy()
    lsh = np.where(la < 0, -1, np.where((la >= v0) & (la < v1), la - v0, 0x3FFFFFFF))  # b4cp_shard_labels
    z = xa @ Ws + bs
    m = z.max(-1)
    lse_loc = m + np.log(np.exp(z - m[:, None]).sum(-1))
    owned = (lsh >= 0) & (lsh < v1 - v0)
    tgt = np.where(owned, z[np.arange(len(la)), np.where(owned, lsh, 0)], 0.0)
    parts = [torch.empty(len(la), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(lse_loc))
    P = torch.stack(parts).numpy()
    pm = P.max(0)
    lse = pm + np.log(np.exp(P - pm).sum(0))                       # b4cp_lse_merge
    tg = torch.from_numpy(tgt)
    dist.all_reduce(tg)
    valid = la >= 0
    n = valid.sum()
    loss = float(((lse - tg.numpy()) * valid).sum() / n)
    # backward: shard-local dZ with the global lse; dX partial -> reduce-scatter; dW local
    dz = np.exp(z - lse[:, None]) * valid[:, None] / n
    dz[np.nonzero(owned)[0], lsh[owned]] -= 1.0 / n
    dx_all = torch.from_numpy(dz @ Ws.T)
    dx_loc = torch.empty(M, h, dtype=torch.float64)
    dist.reduce_scatter(dx_loc, list(dx_all.view(world, M, h).unbind(0)))
    dW, db = xa.T @ dz, dz.sum(0)
    # top-k: per-shard candidates with global ids, all-to-all so each rank merges its own rows
    k = 5
    order = np.lexsort((np.broadcast_to(np.arange(v1 - v0), z.shape), -z), axis=-1)[:, :k]
    cs = torch.from_numpy(np.take_along_axis(z, order, 1).copy())
    ci = torch.from_numpy((order + v0).astype(np.int64))
    cs_x, ci_x = torch.empty_like(cs), torch.empty_like(ci)
    dist.all_to_all_single(cs_x, cs)
    dist.all_to_all_single(ci_x, ci)
    cs_m = cs_x.view(world, M, k).permute(1, 0, 2).reshape(M, world * k).numpy()
    ci_m = ci_x.view(world, M, k).permute(1, 0, 2).reshape(M, world * k).numpy()
    sel = np.lexsort((ci_m, -cs_m), axis=-1)[:, :k]
    top = np.take_along_axis(ci_m, sel, 1)
    q.put((rank, loss, dx_loc.numpy(), (v0, v1), dW, db, top))
    dist.destroy_process_group()


def test_two_rank_gloo_vocab_parallel_equals_single_process():
    import torch.multiprocessing as mp
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_vp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X, W, b, labels = _vp_problem(world)
    M, h = X.shape[1:]
    xa, la = X.reshape(-1, h), labels.reshape(-1)
    z = xa @ W + b
    loss, dz, n = O.cloze_ce_from_logits(z, la)
    dX, dW, db = dz @ W.T, xa.T @ dz, dz.sum(0)
    want_top = O.top_k_ids(z, 5)
    for rank, loss_r, dx_loc, (v0, v1), dW_r, db_r, top in res:
        assert abs(loss_r - loss) < 1e-12
        np.testing.assert_allclose(dx_loc, dX[rank * M:(rank + 1) * M], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(dW_r, dW[:, v0:v1], rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(db_r, db[v0:v1], rtol=1e-10, atol=1e-14)
        assert top.tolist() == want_top[rank * M:(rank + 1) * M].tolist()


# ------------------------------------------------------------------ real-data input side (N2)
def test_bert4rec_text_reader_and_cloze_batches_follow_the_reference_rules():
    """data.py against the reference's rules: first-50 cut and vocabulary order
    (data_prep/main.py:62-76), TRAIN drop-last + n = clip(int(len*p),0,max) distinct sorted
    positions + label ids without the +10 offset, EVAL masks the last item only
    (input_pipeline.py:21-32,:59-133), and the chained layout with the trailing [SEP] after the
    pad run (clickstream_transformer.py:54-61)."""
    import os
    from bert4clickpath_b200 import data as D
    from bert4clickpath_b200.constants import CLS, SEP, MASK_ID
    from oracle import clickpath_oracle as O
    path = os.path.join(os.path.dirname(__file__), "golden", "tiny_bert4rec.txt")
    ds = D.ClozeDataset(path)
    users, items = D.read_bert4rec_text_data(path)
    assert len(ds) == 12 and max(len(s) for s in ds.sessions) == 50        # user 7: 60 -> 50
    # vocabulary = order of first appearance over the KEPT rows
    kept, cnt = [], {}
    for u, it in zip(users, items):
        cnt[u] = cnt.get(u, 0) + 1
        if cnt[u] <= 50:
            kept.append(it)
    assert ds.vocab == list(dict.fromkeys(kept))
    rng = np.random.default_rng(0)
    tr = D.cloze_batch(ds.session_ids, "train", rng, 0.4, 10)
    B, S = tr["ids"].shape
    assert (tr["ids"][:, 0] == CLS).all() and (tr["ids"][:, 1] == SEP).all() and (tr["ids"][:, -1] == SEP).all()
    tot = 0
    for b in range(B):
        full = ds.session_ids[b]
        n_in = len(full) - 1                                               # last item held out
        row = tr["ids"][b, 2:2 + n_in]
        assert (tr["ids"][b, 2 + n_in:-1] == 0).all()                      # pad run BEFORE the final [SEP]
        pos = np.nonzero(row == MASK_ID)[0]
        assert len(pos) == min(int(np.float32(n_in) * np.float32(0.4)), 10)
        lab = tr["labels"][b]
        assert (lab[:len(pos)] == full[:-1][pos] - 10).all() and (lab[len(pos):] == -1).all()
        assert (np.delete(row, pos) == np.delete(full[:-1], pos)).all()    # everything else untouched
        tot += len(pos)
    assert tr["n_masked"] == tot
    ev = D.cloze_batch(ds.session_ids, "eval", rng)
    for b in range(B):
        full = ds.session_ids[b]
        assert ev["ids"][b, 2 + len(full) - 1] == MASK_ID and ev["labels"][b, 0] == full[-1] - 10
        assert (ev["ids"][b, 2:2 + len(full) - 1] == full[:-1]).all()
    assert ev["labels"].shape == (B, 1) and ev["n_masked"] == B
    # the oracle's chaining of the same padded item matrix gives the same ids
    assert (O.chain_sequences([tr["items"]]) == tr["ids"]).all()
    # batches(): every session exactly once per epoch
    seen = sum(b["ids"].shape[0] for b in ds.batches(5, "train", rng))
    assert seen == len(ds)


def test_keyed_mask_positions_oracle_matches_library_key_and_reference_rules():
    """N2: the oracle's position keys are the library's (host export, no GPU), the selected
    positions are n distinct sorted positions, uniformly spread, and the keyed batch obeys the
    reference's rules (drop-last, n = clip(int(len p), 0, max), labels = id - 10 in position order,
    pad-before-chain layout) exactly like data.cloze_batch does for its own draw."""
    from bert4clickpath_b200 import ops
    from bert4clickpath_b200.build import build_lib
    from bert4clickpath_b200.data import cloze_batch
    build_lib(verbose=False)
    for seed, sess, pos in [(0, 0, 0), (1234, 17, 5), (2 ** 64 - 1, 3, 9), (99, 2 ** 40, 2047)]:
        want = O.splitmix64((O.splitmix64((seed + sess) & (2 ** 64 - 1)) + pos) & (2 ** 64 - 1))
        assert ops.cloze_position_key(seed, sess, pos) == want
    assert O.splitmix64(0) == 0xE220A8397B1DCDAF          # published SplitMix64 first output
    # subsets: distinct, sorted, right size, and every position equally likely
    hits = np.zeros(20)
    for s in range(4000):
        p = O.keyed_mask_positions(7, s, 20, 5)
        assert len(p) == 5 and len(set(p)) == 5 and p == sorted(p) and 0 <= p[0] and p[-1] < 20
        hits[p] += 1
    assert np.abs(hits / 4000 - 0.25).max() < 0.03        # 4.4 sigma of a binomial(4000, .25)
    assert O.keyed_mask_positions(7, 1, 20, 0) == [] and O.keyed_mask_positions(7, 1, 3, 3) == [0, 1, 2]
    assert O.keyed_mask_positions(7, 1, 20, 5) != O.keyed_mask_positions(8, 1, 20, 5)
    # batch rules
    rng = np.random.default_rng(0)
    sessions = [rng.integers(10, 500, size=n).astype(np.int32) for n in (5, 1, 12, 50, 7, 2)]
    idx = [3, 0, 5, 1, 2]
    for mode in ("train", "eval"):
        use = idx if mode == "train" else [i for i in idx if len(sessions[i]) >= 1]
        ids, lab, n = O.keyed_cloze_batch(sessions, use, mode, 42, 0.4, 10)
        ref = cloze_batch([sessions[i] for i in use], mode, np.random.default_rng(1), 0.4, 10)
        assert ids.shape == ref["ids"].shape and lab.shape == ref["labels"].shape and n == ref["n_masked"]
        for b, s in enumerate(use):
            src = sessions[s][:-1] if mode == "train" else sessions[s]
            row = ids[b]
            assert row[0] == 3 and row[1] == 4 and row[-1] == 4
            body = row[2:2 + len(src)]
            masked = np.nonzero(body == 1)[0]
            k = int((lab[b] != -1).sum())
            assert len(masked) == k == int((ref["labels"][b] != -1).sum())
            assert (body[body != 1] == src[body != 1]).all() and (row[2 + len(src):-1] == 0).all()
            assert (lab[b, :k] == src[masked] - 10).all() and (lab[b, k:] == -1).all()
            if mode == "eval":
                assert list(masked) == [len(src) - 1]
    # the mask count is the reference's float32 product (90 * 0.7 -> 63, not float64's 62)
    from bert4clickpath_b200.synthetic import n_masked_for
    long = [np.arange(10, 10 + n, dtype=np.int32) for n in (91, 171, 181)]
    _, lab7, n7 = O.keyed_cloze_batch(long, [0, 1, 2], "train", 1, 0.7, 1000)
    assert [int((r != -1).sum()) for r in lab7] == [63, 119, 126] == [n_masked_for(n, 0.7, 1000) for n in (90, 170, 180)]
    assert n7 == 63 + 119 + 126
    # fixed shapes pad further
    ids2, lab2, _ = O.keyed_cloze_batch(sessions, idx, "train", 42, 0.4, 10, L=60, Mmax=12)
    assert ids2.shape == (5, 63) and lab2.shape == (5, 12) and (ids2[:, -1] == 4).all()


def test_oracle_reproduces_golden_sigmoid_heads_and_keyed_cloze():
    z = np.load(os.path.join(GOLD, "sigmoid_heads.npz"))
    layers = [(z["w0"], z["b0"])]
    loss, dx, lg, dwo, dbo = O.binary_head_loss_and_grads(z["xb"], layers, z["wb"], z["bb"], z["yb"],
                                                          pos_weight=3.0)
    assert abs(loss - float(z["b_loss"])) < 1e-13
    for got, key in [(dx, "b_dx"), (lg[0][0], "b_dw0"), (lg[0][1], "b_db0"), (dwo, "b_dwo"), (dbo, "b_dbo")]:
        np.testing.assert_allclose(got, z[key], rtol=1e-12, atol=1e-15, err_msg=key)
    loss, dx, lg, dwo, dbo = O.multilabel_head_loss_and_grads(z["xm"], layers, z["wm"], z["bm"], z["ym"])
    assert abs(loss - float(z["m_loss"])) < 1e-13
    for got, key in [(dx, "m_dx"), (lg[0][0], "m_dw0"), (lg[0][1], "m_db0"), (dwo, "m_dwo"), (dbo, "m_dbo")]:
        np.testing.assert_allclose(got, z[key], rtol=1e-12, atol=1e-15, err_msg=key)
    k = np.load(os.path.join(GOLD, "keyed_cloze.npz"))
    offs = k["offsets"]
    sessions = [k["items"][offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    ids, lab, n = O.keyed_cloze_batch(sessions, k["idx"], "train", 1234, 0.4, 10)
    assert ids.tobytes() == k["train_ids"].tobytes() and lab.tobytes() == k["train_labels"].tobytes()
    assert n == int(k["train_n"])
    ids, lab, n = O.keyed_cloze_batch(sessions, k["idx"], "eval", 1234, 0.4, 10, L=52, Mmax=3)
    assert ids.tobytes() == k["eval_ids"].tobytes() and lab.tobytes() == k["eval_labels"].tobytes()
    assert n == int(k["eval_n"]) == len(k["idx"])
    # the committed keys are also what the library's host export computes
    from bert4clickpath_b200 import ops
    from bert4clickpath_b200.build import build_lib
    build_lib(verbose=False)
    assert [ops.cloze_position_key(1234, 3, i) for i in range(4)] == [int(v) for v in k["keys_s3"]]


def test_vocab_forward_planner_invariants():
    """b4cp_vocab_ce_plan (host-only): under the persistent-range schedule every (row tile,
    vocabulary tile) is covered exactly once, a CTA's pieces of a row tile land in distinct slots
    below the reported slot count, and the lock-step grid is kept when W does not fit L2."""
    import ctypes
    from bert4clickpath_b200 import _lib
    L = _lib.lib()
    out = (ctypes.c_long * 4)()
    for M, V, h in [(28672, 54293, 128), (448, 54293, 128), (7168, 54293, 128), (232, 20000, 256),
                    (1, 1, 128), (129, 257, 64), (100000, 54293, 128), (5000, 130000, 128)]:
        assert L.b4cp_vocab_ce_plan(ctypes.c_long(M), V, h, out) == 0
        sched, ctas, slots, q = list(out)
        n_m, n_v = -(-M // 128), -(-V // 128)
        assert 1 <= slots <= 24
        if sched == 0:
            assert ctas == n_m * slots and q * slots >= n_v
            continue
        total = n_m * n_v
        assert ctas == -(-total // q) <= 148
        seen = {}
        for k in range(ctas):
            lo, hi = k * q, min((k + 1) * q, total)
            while lo < hi:
                m, v0 = divmod(lo, n_v)
                ln = min(n_v - v0, hi - lo)
                slot = k - (m * n_v) // q
                assert 0 <= slot < slots and (m, slot) not in seen
                seen[(m, slot)] = (v0, ln)
                lo += ln
        for m in range(n_m):     # the slots of a row tile tile its vocabulary range, in order
            pieces = sorted(v for (mm, s), v in seen.items() if mm == m)
            assert pieces[0][0] == 0 and sum(p[1] for p in pieces) == n_v
            n_slots_m = ((m + 1) * n_v - 1) // q - (m * n_v) // q + 1
            assert len(pieces) == n_slots_m
    assert L.b4cp_vocab_ce_plan(ctypes.c_long(7424), 1_000_000, 256, out) == 0 and out[0] == 0


def test_table_exchange_plan_is_by_size(monkeypatch):
    """EncoderEngine.plan_table_exchange (host logic): a table goes by row exchange only when a
    step's gathered token rows are much smaller than the table, and the store then leaves it out
    of the data-parallel all-reduce."""
    import types
    import torch.distributed as dist
    from bert4clickpath_b200 import engine as E

    class FakeParam:
        def __init__(self):
            self.grad_is_global = False

    eng = E.EncoderEngine.__new__(E.EncoderEngine)
    eng.rows, eng.dims, eng.d = [1_000_011, 61], [240, 16], 256
    eng.store = {"emb.0": FakeParam(), "emb.1": FakeParam()}
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_world_size", lambda group=None: 8)
    plan, world = eng.plan_table_exchange(256 * 202)           # C4: 424 MB of rows vs a 960 MB table
    assert world == 8 and plan == [True, False]
    assert eng.store["emb.0"].grad_is_global and not eng.store["emb.1"].grad_is_global
    eng.rows, eng.dims, eng.d = [54304], [64], 64
    eng.store = {"emb.0": FakeParam()}
    assert eng.plan_table_exchange(4096 * 52)[0] == [False]    # C1: 436 MB of rows vs a 13.9 MB table
    monkeypatch.setattr(dist, "get_world_size", lambda group=None: 1)
    eng.rows, eng.dims, eng.d = [1_000_011], [256], 256
    assert eng.plan_table_exchange(256 * 202)[0] == [False]    # one rank: nothing to exchange


def test_native_vocabulary_table_equals_a_dict_lookup():
    """b4cp_vocab_table_* (host code in libb4cp, csrc/vocab_table.cu) against the Python dict
    semantics of tf.lookup.StaticVocabularyTable with one OOV bucket
    (clickstream_transformer.py:247-258): first occurrence of a duplicated key, every unknown
    string -> len(keys), any string width, unicode, empty strings, bytes, ids passed through."""
    from bert4clickpath_b200.clickstream_transformer import StaticVocabularyTable
    from bert4clickpath_b200.constants import RESERVED_TOKENS
    rng = np.random.default_rng(5)
    alphabet = list("abcXYZ019_-[]") + ["é", "ü", "項", "𝄞"]
    keys = list(RESERVED_TOKENS)
    for _ in range(3000):
        keys.append("".join(rng.choice(alphabet, size=int(rng.integers(0, 14)))))
    keys += keys[100:120]                                    # duplicates: first occurrence wins
    table = StaticVocabularyTable(keys)
    want = {}
    for i, k in enumerate(keys):
        want.setdefault(k, i)
    assert table.size() == len(keys) + 1
    probes = list(rng.choice(keys, size=5000)) + [k + "?" for k in rng.choice(keys, size=500)] + \
        ["", "a-much-longer-token-than-any-key-in-the-table", "[MASK]", "[PAD]"]
    for dtype in (np.str_, object):
        arr = np.asarray(probes, dtype=dtype).reshape(-1, 4)
        got = table.lookup(arr)
        assert got.dtype == np.int32 and got.shape == arr.shape
        assert got.reshape(-1).tolist() == [want.get(p, len(keys)) for p in probes]
    ascii_probes = [p for p in probes if p.isascii()]
    got = table.lookup(np.asarray([p.encode() for p in ascii_probes]))
    assert got.tolist() == [want.get(p, len(keys)) for p in ascii_probes]
    assert table.lookup(np.array([[3, 7]])).tolist() == [[3, 7]]          # ids pass through
    assert table.lookup(np.zeros((0, 5), dtype="<U3")).shape == (0, 5)
    out = np.full((2, 2), -7, dtype=np.int32)
    assert table.lookup(np.array([["[CLS]", "[SEP]"], ["nope", "[NA]"]]), out=out) is out
    assert out.tolist() == [[3, 4], [len(keys), 5]]
    # many tokens: the multi-threaded split covers every element exactly once
    big = np.asarray(rng.choice(keys[:2000], size=200_000), dtype=np.str_)
    assert np.array_equal(table.lookup(big), np.array([want[p] for p in big.tolist()], dtype=np.int32))


def test_chaining_fixed_width_strings_stays_fixed_width():
    from bert4clickpath_b200.clickstream_transformer import TransformerInputPrep
    a = np.array([["x1", "[PAD]"], ["a-long-token", "y"]])
    b = np.array([["q"], ["r"]])
    raw, starts, ends = TransformerInputPrep({"items": ["a", "b"]})(features={"a": a, "b": b})
    assert raw["items"].dtype.kind == "U"
    assert raw["items"].tolist() == [["[CLS]", "[SEP]", "x1", "[PAD]", "[SEP]", "q", "[SEP]"],
                                     ["[CLS]", "[SEP]", "a-long-token", "y", "[SEP]", "r", "[SEP]"]]
    assert starts.tolist() == [0, 2, 5] and ends.tolist() == [1, 4, 6]
