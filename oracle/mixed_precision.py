"""ORACLE add-on (test infrastructure): the same Cloze training step as
clickpath_oracle.cloze_train_step, with operands rounded to bfloat16 at exactly the points where
the CUDA pipeline stores bf16 (GEMM operands, attention inputs/outputs, ReLU activations, the
softmax gradient).  All arithmetic between those points is float64.

Purpose: separates "the kernels compute what they claim" (CUDA vs this emulation: tight, ~1e-3)
from "bf16 tensor-core operands are an acceptable approximation of the reference's fp32 math"
(this emulation vs the exact oracle: the stated bf16 tolerance).
"""
import numpy as np

from . import clickpath_oracle as O


def bf16(a):
    """Round-to-nearest-even to bfloat16, returned as float64."""
    a32 = np.ascontiguousarray(a, dtype=np.float32)
    u = a32.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return rounded.view(np.float32).astype(np.float64).reshape(np.shape(a))


def _tensor_core_attention(S, dh):
    """The CUDA attention uses mma.sync (bf16 P / dZ fragments) for S <= 128 and head depth
    32 or 64, and fp32 SIMT math otherwise (csrc/attention_mma.cu vs csrc/encoder.cu)."""
    return S <= 128 and dh in (32, 64)


def _umma_attention_fwd(S, H, dh):
    """csrc/attention_umma.cu (tcgen05): head depth 32, an even number of heads, S <= 128.  Its
    probabilities go to the tensor core UN-normalised (exp(z - max) in bf16) and the output is
    divided by the fp32 row sum afterwards."""
    return dh == 32 and H % 2 == 0 and S <= 128


def _umma_attention_bwd(S, H, dh):
    """... and its backward: dZ is rounded to bf16 AFTER the 1/sqrt(dh) scale."""
    return dh == 32 and H % 2 == 0 and S <= 128


def _mha_fwd(qm, km, vm, pad, H):
    o, att = O.mha_core_fwd(qm, km, vm, pad, H)
    B, S, d = qm.shape
    dh = d // H
    if _umma_attention_fwd(S, H, dh):
        e = att["a"] / att["a"].max(-1, keepdims=True)      # exp(z - max)
        o = ((bf16(e) @ att["vh"]) / e.sum(-1, keepdims=True)).transpose(0, 2, 1, 3).reshape(B, S, d)
    elif _tensor_core_attention(S, dh):
        o = (bf16(att["a"]) @ att["vh"]).transpose(0, 2, 1, 3).reshape(B, S, d)
    return o, att


def _mha_bwd(do_merged, att):
    a, qh, kh, vh, H = att["a"], att["qh"], att["kh"], att["vh"], att["H"]
    B, _, S, dh = qh.shape
    if not _tensor_core_attention(S, dh):
        return O.mha_core_bwd(do_merged, att)
    d = H * dh
    do = do_merged.reshape(B, S, H, dh).transpose(0, 2, 1, 3)
    dv = bf16(a).transpose(0, 1, 3, 2) @ do
    da = do @ vh.transpose(0, 1, 3, 2)
    inv = 1.0 / np.sqrt(np.float32(dh)).astype(np.float64)
    if _umma_attention_bwd(S, H, dh):
        dz = bf16(a * (da - (da * a).sum(-1, keepdims=True)) * inv)
        dq = dz @ kh
        dk = dz.transpose(0, 1, 3, 2) @ qh
    else:
        dz = bf16(a * (da - (da * a).sum(-1, keepdims=True)))
        dq = (dz @ kh) * inv
        dk = (dz.transpose(0, 1, 3, 2) @ qh) * inv
    merge = lambda t: t.transpose(0, 2, 1, 3).reshape(B, S, d)
    return merge(dq), merge(dk), merge(dv)


def _layer_fwd(x, pad, p, H, drop1, drop2):
    q = bf16
    B, S, d = x.shape
    xb = q(x)
    wq = np.concatenate([p["wq"], p["wk"], p["wv"]], axis=1)
    bq = np.concatenate([p["bq"], p["bk"], p["bv"]])
    qkv = q(xb @ q(wq) + bq)
    o, att = _mha_fwd(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], pad, H)
    ob = q(o)
    y1 = ob @ q(p["wo"]) + p["bo"]
    r1 = x + (y1 * drop1 if drop1 is not None else y1)
    x1, ln1 = O.layer_norm_fwd(r1, p["ln1_g"], p["ln1_b"])
    x1b = q(x1)
    hb = q(np.maximum(x1b @ q(p["w1"]) + p["b1"], 0))
    y2 = hb @ q(p["w2"]) + p["b2"]
    r2 = x1 + (y2 * drop2 if drop2 is not None else y2)
    x2, ln2 = O.layer_norm_fwd(r2, p["ln2_g"], p["ln2_b"])
    return x2, dict(xb=xb, att=att, ob=ob, ln1=ln1, x1b=x1b, hb=hb, ln2=ln2, drop1=drop1,
                    drop2=drop2, d=d)


def _layer_bwd(dx2, c, p):
    q = bf16
    d = c["d"]
    g = {}
    f2 = lambda t: t.reshape(-1, t.shape[-1])
    dr2, g["ln2_g"], g["ln2_b"] = O.layer_norm_bwd(dx2, c["ln2"], p["ln2_g"])
    dy2 = dr2 * c["drop2"] if c["drop2"] is not None else dr2
    g["b2"] = f2(dy2).sum(0)
    dy2b = q(dy2)
    g["w2"] = f2(c["hb"]).T @ f2(dy2b)
    dhb = q((dy2b @ q(p["w2"]).T) * (c["hb"] > 0))
    g["w1"] = f2(c["x1b"]).T @ f2(dhb)
    g["b1"] = f2(dhb).sum(0)
    dx1 = dr2 + dhb @ q(p["w1"]).T
    dr1, g["ln1_g"], g["ln1_b"] = O.layer_norm_bwd(dx1, c["ln1"], p["ln1_g"])
    dy1 = dr1 * c["drop1"] if c["drop1"] is not None else dr1
    g["bo"] = f2(dy1).sum(0)
    dy1b = q(dy1)
    g["wo"] = f2(c["ob"]).T @ f2(dy1b)
    dob = q(dy1b @ q(p["wo"]).T)
    dq, dk, dv = _mha_bwd(dob, c["att"])
    dqkvb = q(np.concatenate([dq, dk, dv], axis=-1))
    wqkv = np.concatenate([p["wq"], p["wk"], p["wv"]], axis=1)
    gw = f2(c["xb"]).T @ f2(dqkvb)
    gb = f2(dqkvb).sum(0)
    for i, nm in enumerate("qkv"):
        g["w" + nm] = gw[:, i * d:(i + 1) * d]
        g["b" + nm] = gb[i * d:(i + 1) * d]
    dx = dr1 + dqkvb @ q(wqkv).T
    return dx, g


def cloze_train_step_bf16(ids_list, labels, P, num_layers, num_heads, pe, masks=None):
    """Same signature / outputs as clickpath_oracle.cloze_train_step (dtype is float64)."""
    q = bf16
    F = len(ids_list)
    P = {k: np.asarray(v, dtype=np.float64) for k, v in P.items()}
    tables = [P[f"emb.{f}"] for f in range(F)]
    x = O.embed_fwd(ids_list, tables, pe, dtype=np.float32).astype(np.float64)
    if masks and masks.get("in") is not None:
        x = (x.astype(np.float32) * masks["in"].astype(np.float32)).astype(np.float64)
    pad = O.create_padding_mask(ids_list[0])
    caches = []
    for l in range(num_layers):
        p = O.layer_params(P, l)
        x, c = _layer_fwd(x, pad, p, num_heads, masks.get((l, 1)) if masks else None,
                          masks.get((l, 2)) if masks else None)
        caches.append((c, p))
    B, S, d = x.shape
    sel, index = O.select_masked(ids_list[0], x)
    Mmax = sel.shape[1]
    acts = [q(sel.reshape(B * Mmax, d))]
    layers = O.head_layers(P)
    for w, b in layers:
        acts.append(q(np.maximum(acts[-1] @ q(w) + b, 0)))
    w_out, b_out = P["head.out.w"], P["head.out.b"]
    logits = acts[-1] @ q(w_out) + b_out
    loss, dz, n = O.cloze_ce_from_logits(logits, np.asarray(labels).reshape(-1))
    dzb = q(dz)
    G = {"head.out.w": acts[-1].T @ dzb, "head.out.b": dzb.sum(0)}
    dbg = dict(dz=dzb, acts=acts)
    if layers:
        dzl = q((dzb @ q(w_out).T) * (acts[-1] > 0))
        for i in reversed(range(len(layers))):
            w, b = layers[i]
            dbg[f"dzl{i}"] = dzl
            G[f"head.{i}.w"] = acts[i].T @ dzl
            G[f"head.{i}.b"] = dzl.sum(0)
            if i > 0:
                dzl = q((dzl @ q(w).T) * (acts[i] > 0))
            else:
                dflat = dzl @ q(w).T
    else:
        dflat = dzb @ q(w_out).T
    dsel = dflat.reshape(B, Mmax, d)
    dx = np.zeros_like(x)
    cnt = np.zeros(B, dtype=np.int64)
    for (b, s) in index:
        dx[b, s] = dsel[b, cnt[b]]
        cnt[b] += 1
    for l in reversed(range(num_layers)):
        c, p = caches[l]
        dx, g = _layer_bwd(dx, c, p)
        for k, v in g.items():
            G[f"enc.{l}.{k}"] = v
    if masks and masks.get("in") is not None:
        dx = dx * masks["in"]
    dims = [P[f"emb.{f}"].shape[1] for f in range(F)]
    rows = [P[f"emb.{f}"].shape[0] for f in range(F)]
    for f, g in enumerate(O.embed_bwd(dx, ids_list, dims, rows, np.float64)):
        G[f"emb.{f}"] = g
    dbg["dsel"] = dflat
    return loss, G, dict(logits=logits, n_valid=n, enc_out=x, dbg=dbg)
