"""CPU ORACLE (test infrastructure, NOT a product path) for the clickstream-transformer hot path.

A NumPy restatement of MiladShahidi/BERT4ClickPath's algorithm, function by function, with the
reference file:line each block follows (paths relative to the reference repository).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module, and only as the checker or the timed CPU baseline.

PARITY PIN: the reference's arithmetic lives in TensorFlow 2.3.1 (requirements.txt:33), which
cannot be installed here, and the reference ships no tests or golden vectors.  The oracle is
pinned by OUTPUTS OF THE REFERENCE'S OWN SOURCE RUN IN THIS CONTAINER: its modules are imported
unmodified from /root/reference on top of tests/golden/tf_shim/tensorflow (a small eager
implementation of the TensorFlow calls they make, torch autograd as the tape) by
tests/golden/make_reference_golden.py, and the results are committed as
tests/golden/reference_*.npz.  tests/test_reference_golden.py holds this file to them: loss,
probabilities and every gradient of the Cloze model (inference and with recorded dropout masks),
of the two-feature segment / BinaryClassificationHead model and of the multi-label head within
1e-12 in float64 (measured 2e-15) and 2e-5 in float32 (measured 1e-6); NDCG / recall / the binary
metrics; the masking pipeline under keyed permutations; the float32 mask-count rule.  What that
does NOT pin is TensorFlow's own kernels (summation order, its softmax / rsqrt implementations):
the float32 run is "float32 with torch's summation order".  Also held: (i) the two known answers
in the reference's `__main__` blocks (losses.py:101-123 -> 2.9957323;
examples/BERT4Rec/source/utils.py:262-272 -> 0.81546488), (ii) the docstring example of
transformer.py:8-19, (iii) an independent torch-autograd restatement in tests/test_oracle.py.
TensorFlow-library behaviours assumed are listed in SURVEY.md Appendix C and selectable here
where they matter (`ce_mode`).

All functions take a `dtype` (np.float64 = "truth", np.float32 = TF-like rounding).
"""
import math

import numpy as np

# ----------------------------------------------------------------------------- constants.py:1-31
LABEL_PAD = -1.0
NUM_RESERVED_TOKENS = 10
RESERVED_TOKENS = ["[PAD]", "[MASK]", "[UNK]", "[CLS]", "[SEP]", "[NA]"] + [
    f"[RESERVED_{i}]" for i in range(6, NUM_RESERVED_TOKENS)]
INPUT_PAD = 0
MASK_ID = 1
UNK_ID = 2
CLS = 3
SEP = 4


# ------------------------------------------------------------- clickstream_transformer.py:38-103
def chain_sequences(sequences):
    """[CLS] [SEP] seq_1 [SEP] seq_2 [SEP] ... on integer ids (the reference chains strings and
    looks them up afterwards, :307-308; the id of a token does not depend on its position)."""
    B = sequences[0].shape[0]
    cls = np.full((B, 1), CLS, dtype=sequences[0].dtype)
    sep = np.full((B, 1), SEP, dtype=sequences[0].dtype)
    parts = [cls, sep]
    for s in sequences:
        parts += [s, sep]
    return np.concatenate(parts, axis=1)


def segment_bounds(chained_row0):
    """segment_ends = SEP positions of sample 0; starts = [0, ends[:-1]+1]  (:86-94)."""
    ends = np.nonzero(chained_row0 == SEP)[0]
    starts = np.concatenate([[0], ends[:-1] + 1])
    return starts, ends


def create_segment_markers(seq, sep=SEP):
    """transformer.py:6-35 (unused by the model; kept for its docstring known answer)."""
    return np.cumsum((seq == sep).astype(np.int32), axis=1)


def lookup_ids(tokens, vocab):
    """clickstream_transformer.py:247-258: reserved tokens 0..9, vocab line j -> 10+j,
    anything else -> the single OOV bucket = len(reserved)+len(vocab)."""
    table = {t: i for i, t in enumerate(RESERVED_TOKENS)}
    for j, t in enumerate(vocab):
        table.setdefault(t, NUM_RESERVED_TOKENS + j)
    oov = NUM_RESERVED_TOKENS + len(vocab)
    flat = [table.get(t, oov) for t in np.asarray(tokens).reshape(-1)]
    return np.asarray(flat, dtype=np.int64).reshape(np.asarray(tokens).shape)


# ----------------------------------------------------------------------- transformer.py:38-61
def create_padding_mask(ids_first_feature):
    return (ids_first_feature == INPUT_PAD)


def positional_encoding(position, d_model):
    """float64 angles (np.float32(d_model) only enters the exponent divisor), cast to f32."""
    pos = np.arange(position)[:, np.newaxis]
    i = np.arange(d_model)[np.newaxis, :]
    angle_rates = 1 / np.power(10000, (2 * (i // 2)) / np.float32(d_model))
    angle_rads = pos * angle_rates
    angle_rads[:, 0::2] = np.sin(angle_rads[:, 0::2])
    angle_rads[:, 1::2] = np.cos(angle_rads[:, 1::2])
    return angle_rads.astype(np.float32)


# --------------------------------------------------------------------- transformer.py:376-398
def embed_fwd(ids_list, tables, pe, dtype=np.float32):
    """Per-feature gather, concat, * sqrt(d_model) (separately rounded), + PE[:S]."""
    emb = np.concatenate([t.astype(dtype)[ids] for ids, t in zip(ids_list, tables)], axis=-1)
    d_model = emb.shape[-1]
    S = emb.shape[1]
    scale = dtype(np.sqrt(np.float32(d_model)))
    out = (emb * scale).astype(dtype)
    out = (out + pe[:S].astype(dtype)[None]).astype(dtype)
    return out


def embed_bwd(dout, ids_list, dims, rows, dtype=np.float64):
    """dE_f[r] = sqrt(d_model) * sum over tokens with id r of dout[..., off_f:off_f+d_f]."""
    d_model = dout.shape[-1]
    scale = dtype(np.sqrt(np.float32(d_model)))
    grads, off = [], 0
    for ids, d_f, R in zip(ids_list, dims, rows):
        g = np.zeros((R, d_f), dtype=dtype)
        np.add.at(g, ids.reshape(-1), dout.reshape(-1, d_model)[:, off:off + d_f].astype(dtype))
        grads.append(g * scale)
        off += d_f
    return grads


# ----------------------------------------------------------------------------- building blocks
def layer_norm_fwd(r, gamma, beta, eps=1e-6):
    """tf.keras LayerNormalization(epsilon=1e-6), non-fused path: biased variance."""
    mu = r.mean(axis=-1, keepdims=True)
    var = ((r - mu) ** 2).mean(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + r.dtype.type(eps))
    xhat = (r - mu) * rstd
    return xhat * gamma + beta, (xhat, rstd)


def layer_norm_bwd(dy, cache, gamma):
    xhat, rstd = cache
    dgamma = (dy * xhat).reshape(-1, xhat.shape[-1]).sum(0)
    dbeta = dy.reshape(-1, xhat.shape[-1]).sum(0)
    g = dy * gamma
    dr = rstd * (g - g.mean(-1, keepdims=True) - xhat * (g * xhat).mean(-1, keepdims=True))
    return dr, dgamma, dbeta


def _softmax(z):
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m)
    return e / e.sum(axis=-1, keepdims=True)


def mha_core_fwd(q, k, v, pad, H):
    """split_heads + scaled_dot_product_attention + merge (transformer.py:130-156, :64-97).
    q,k,v (B,S,d) after the Dense projections; pad (B,S) bool; returns merged (B,S,d)."""
    B, S, d = q.shape
    dh = d // H
    dt = q.dtype.type

    def split(t):
        return t.reshape(B, S, H, dh).transpose(0, 2, 1, 3)

    qh, kh, vh = split(q), split(k), split(v)
    z = (qh @ kh.transpose(0, 1, 3, 2)) / dt(np.sqrt(np.float32(dh)))
    z = z + pad[:, None, None, :].astype(q.dtype) * dt(-1e9)
    a = _softmax(z)
    o = (a @ vh).transpose(0, 2, 1, 3).reshape(B, S, d)
    m = z.max(-1)
    lse = m + np.log(np.exp(z - m[..., None]).sum(-1))
    return o, dict(qh=qh, kh=kh, vh=vh, a=a, H=H, lse=lse)


def mha_core_bwd(do_merged, att):
    a, qh, kh, vh, H = att["a"], att["qh"], att["kh"], att["vh"], att["H"]
    B, _, S, dh = qh.shape
    d = H * dh
    dt = qh.dtype.type
    do = do_merged.reshape(B, S, H, dh).transpose(0, 2, 1, 3)
    dv = a.transpose(0, 1, 3, 2) @ do
    da = do @ vh.transpose(0, 1, 3, 2)
    dz = a * (da - (da * a).sum(-1, keepdims=True))
    inv = dt(1.0) / dt(np.sqrt(np.float32(dh)))
    dq = (dz @ kh) * inv
    dk = (dz.transpose(0, 1, 3, 2) @ qh) * inv

    def merge(t):
        return t.transpose(0, 2, 1, 3).reshape(B, S, d)

    return merge(dq), merge(dk), merge(dv)


def encoder_layer_fwd(x, pad, p, num_heads, drop1=None, drop2=None):
    """transformer.py:202-213 (EncoderLayer), :137-160 (MHA), :64-97 (SDPA), :163-167 (FFN).

    x (B,S,d); pad (B,S) bool (True = key is padding); p: dict of this layer's parameters with
    Keras (in,out) kernels.  drop1/drop2: optional multiplicative masks already scaled by
    1/(1-rate) (Dropout is identity when they are None)."""
    B, S, d = x.shape
    H = num_heads
    dh = d // H
    dt = x.dtype.type
    q = x @ p["wq"] + p["bq"]
    k = x @ p["wk"] + p["bk"]
    v = x @ p["wv"] + p["bv"]
    o, att = mha_core_fwd(q, k, v, pad, H)
    y = o @ p["wo"] + p["bo"]
    if drop1 is not None:
        y_d = y * drop1
    else:
        y_d = y
    r1 = x + y_d
    x1, ln1 = layer_norm_fwd(r1, p["ln1_g"], p["ln1_b"])
    pre = x1 @ p["w1"] + p["b1"]
    hdn = np.maximum(pre, 0)
    g = hdn @ p["w2"] + p["b2"]
    g_d = g * drop2 if drop2 is not None else g
    r2 = x1 + g_d
    x2, ln2 = layer_norm_fwd(r2, p["ln2_g"], p["ln2_b"])
    cache = dict(x=x, att=att, a=att["a"], o=o, ln1=ln1, x1=x1, pre=pre, hdn=hdn, ln2=ln2,
                 drop1=drop1, drop2=drop2, H=H)
    return x2, cache


def encoder_layer_bwd(dx2, c, p):
    """Hand-derived backward of encoder_layer_fwd."""
    x = c["x"]
    B, S, d = x.shape
    H = c["H"]
    dh = d // H
    dt = x.dtype.type
    g = {}
    dr2, g["ln2_g"], g["ln2_b"] = layer_norm_bwd(dx2, c["ln2"], p["ln2_g"])
    dx1 = dr2.copy()
    dg = dr2 * c["drop2"] if c["drop2"] is not None else dr2
    g["w2"] = c["hdn"].reshape(-1, c["hdn"].shape[-1]).T @ dg.reshape(-1, d)
    g["b2"] = dg.reshape(-1, d).sum(0)
    dh_ = (dg @ p["w2"].T) * (c["pre"] > 0)
    g["w1"] = c["x1"].reshape(-1, d).T @ dh_.reshape(-1, dh_.shape[-1])
    g["b1"] = dh_.reshape(-1, dh_.shape[-1]).sum(0)
    dx1 = dx1 + dh_ @ p["w1"].T
    dr1, g["ln1_g"], g["ln1_b"] = layer_norm_bwd(dx1, c["ln1"], p["ln1_g"])
    dx = dr1.copy()
    dy = dr1 * c["drop1"] if c["drop1"] is not None else dr1
    g["wo"] = c["o"].reshape(-1, d).T @ dy.reshape(-1, d)
    g["bo"] = dy.reshape(-1, d).sum(0)
    dq, dk, dv = mha_core_bwd(dy @ p["wo"].T, c["att"])
    x2d = x.reshape(-1, d)
    for nm, dd in (("q", dq), ("k", dk), ("v", dv)):
        g["w" + nm] = x2d.T @ dd.reshape(-1, d)
        g["b" + nm] = dd.reshape(-1, d).sum(0)
        dx = dx + dd @ p["w" + nm].T
    return dx, g


# ------------------------------------------------------ clickstream_transformer.py:260-297, :322
def select_masked(ids_first_feature, x, value_id=MASK_ID):
    """Rows (b,s) whose first-feature token == value, ordered (b,s), per-example right-padded
    with zero vectors to max_b M_b.  Returns (B, Mmax, d) and the flat (b,s) index list."""
    B, S, d = x.shape
    hit = ids_first_feature == value_id
    counts = hit.sum(1)
    mmax = int(counts.max()) if B > 0 else 0
    out = np.zeros((B, mmax, d), dtype=x.dtype)
    index = []
    for b in range(B):
        pos = np.nonzero(hit[b])[0]
        out[b, :len(pos)] = x[b, pos]
        index += [(b, int(s)) for s in pos]
    return out, index


def select_segment(x, starts, ends, k):
    return x[:, starts[k]:ends[k], :]


# --------------------------------------------------------------------------------- head.py:4-69
def mlp_fwd(x, layers):
    """ReLU Dense stack (head.py:16-19, :41-43). layers: list of (W (in,out), b)."""
    acts = [x]
    for w, b in layers:
        x = np.maximum(x @ w + b, 0)
        acts.append(x)
    return x, acts


def mlp_bwd(dout, acts, layers):
    grads = []
    for (w, b), a_in, a_out in zip(reversed(layers), reversed(acts[:-1]), reversed(acts[1:])):
        dz = dout * (a_out > 0)
        dw = a_in.reshape(-1, a_in.shape[-1]).T @ dz.reshape(-1, dz.shape[-1])
        db = dz.reshape(-1, dz.shape[-1]).sum(0)
        dout = dz @ w.T
        grads.append((dw, db))
    return dout, list(reversed(grads))


def softmax_head_fwd(x, layers, w_out, b_out):
    """SoftMaxHead.call (head.py:38-47): probabilities over the output vocabulary."""
    h, acts = mlp_fwd(x, layers)
    logits = h @ w_out + b_out
    return _softmax(logits), logits, acts


def binary_head_fwd(x, layers, w_out, b_out):
    """BinaryClassificationHead.call (head.py:13-26): sigmoid(Dense(1)), squeeze(-1)."""
    h, acts = mlp_fwd(x, layers)
    z = (h @ w_out + b_out)[..., 0]
    return 1.0 / (1.0 + np.exp(-z)), z, acts


def binary_head_loss_and_grads(x, layers, w_out, b_out, y_true, pos_weight=None, label_pad=LABEL_PAD):
    """Loss and gradients of BinaryClassificationHead (head.py:13-26) under
    MaskedLoss(K.binary_crossentropy, pos_weight) (losses.py:31-98) - what TF autodiff derives:
      l_i = -w_i (y log(pc + e) + (1 - y) log(1 - pc + e)), pc = clip(p, e, 1 - e), e = 1e-7,
      loss = sum_i mask_i l_i / sum_i mask_i  [/ ((pos_weight + 1) / 2) when pos_weight is given]
    (the clip passes a gradient only strictly inside (e, 1 - e)).
    x: (..., in) head input rows; y_true: (...) labels padded with label_pad.
    Returns loss, dx (same shape as x), [(dW, db) per MLP layer], dW_out, db_out."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y_true, dtype=np.float64)
    p, z, acts = binary_head_fwd(x, layers, w_out, b_out)
    mask = (y != label_pad)
    n = mask.sum()
    eps = 1e-7
    pc = np.clip(p, eps, 1 - eps)
    yy = np.where(mask, y, 0.0)
    item = -(yy * np.log(pc + eps) + (1 - yy) * np.log(1 - pc + eps))
    w = np.where(yy == 1, pos_weight, 1.0) if pos_weight is not None else np.ones_like(item)
    norm = ((pos_weight + 1.0) / 2) if pos_weight is not None else 1.0
    loss = (item * w * mask).sum() / n / norm if n > 0 else 0.0
    inside = (p > eps) & (p < 1 - eps)
    dl_dp = -(yy / (pc + eps) - (1 - yy) / (1 - pc + eps)) * w * inside
    dz = np.where(mask, dl_dp * p * (1 - p), 0.0) / max(n, 1) / norm       # (...)
    h = acts[-1]
    dz2 = dz.reshape(-1, 1)
    h2 = h.reshape(-1, h.shape[-1])
    dW_out = h2.T @ dz2
    db_out = dz2.sum(0)
    dh = (dz2 @ np.asarray(w_out, dtype=np.float64).T).reshape(h.shape)
    dx, layer_grads = mlp_bwd(dh, acts, layers)
    return loss, dx, layer_grads, dW_out, db_out


def multilabel_head_fwd(x, layers, w_out, b_out):
    """MultiLabel_MultiClass_classification.call (head.py:59-69): ReLU MLP, sigmoid(Dense(V)),
    tf.squeeze(axis=1).  x: (B, 1, in) -> probabilities (B, V), logits (B, 1, V), activations."""
    assert x.ndim == 3 and x.shape[1] == 1, "tf.squeeze(axis=1) needs a length-1 segment"
    h, acts = mlp_fwd(x, layers)
    z = h @ w_out + b_out
    return (1.0 / (1.0 + np.exp(-z)))[:, 0, :], z, acts


def multilabel_head_loss_and_grads(x, layers, w_out, b_out, y_true, pos_weight=None,
                                   label_pad=LABEL_PAD):
    """Loss and gradients of MultiLabel_MultiClass_classification (head.py:50-69) under
    MaskedLoss(K.binary_crossentropy, pos_weight) (losses.py:31-98): every (row, class) cell with
    y_true != label_pad is one item of the masked mean; same item formula as the binary head.
    x: (B, 1, in); y_true: (B, V).  Returns loss, dx, [(dW, db) per MLP layer], dW_out, db_out."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y_true, dtype=np.float64)
    p, z, acts = multilabel_head_fwd(x, layers, w_out, b_out)
    mask = (y != label_pad)
    n = mask.sum()
    eps = 1e-7
    pc = np.clip(p, eps, 1 - eps)
    yy = np.where(mask, y, 0.0)
    item = -(yy * np.log(pc + eps) + (1 - yy) * np.log(1 - pc + eps))
    w = np.where(yy == 1, pos_weight, 1.0) if pos_weight is not None else np.ones_like(item)
    norm = ((pos_weight + 1.0) / 2) if pos_weight is not None else 1.0
    loss = (item * w * mask).sum() / n / norm if n > 0 else 0.0
    inside = (p > eps) & (p < 1 - eps)
    dl_dp = -(yy / (pc + eps) - (1 - yy) / (1 - pc + eps)) * w * inside
    dz = np.where(mask, dl_dp * p * (1 - p), 0.0) / max(n, 1) / norm       # (B, V)
    h2 = acts[-1].reshape(-1, acts[-1].shape[-1])                           # (B, h)
    dW_out = h2.T @ dz
    db_out = dz.sum(0)
    dh = (dz @ np.asarray(w_out, dtype=np.float64).T).reshape(acts[-1].shape)
    dx, layer_grads = mlp_bwd(dh, acts, layers)
    return loss, dx, layer_grads, dW_out, db_out


# ------------------------------------------- input_pipeline.py:21-32, :59-133 with keyed positions
_M64 = (1 << 64) - 1


def splitmix64(x):
    """The 64-bit mixer the device builder keys mask positions with (plain Python ints)."""
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def keyed_mask_positions(seed, session, length, n):
    """The n positions of range(length) with the smallest keys splitmix64(splitmix64(seed +
    session) + pos), ties by position, ascending - the stand-in for
    `tf.random.shuffle(tf.range(length))[:n]` followed by the sort of input_pipeline.py:77-78."""
    return sorted(keyed_mask_positions_order(seed, session, length)[:n])


def keyed_mask_positions_order(seed, session, length):
    """All positions of range(length) by ascending key (ties by position): the permutation that
    stands in for tf.random.shuffle(tf.range(length)) (input_pipeline.py:28)."""
    base = splitmix64((seed + session) & _M64)
    keys = [(splitmix64((base + i) & _M64), i) for i in range(length)]
    return [i for _, i in sorted(keys)]


def keyed_cloze_batch(sessions_ids, session_idx, mode, seed, masked_percentage, max_masked, L=None,
                      Mmax=None, num_reserved=10, cls=3, sep=4, mask=1):
    """Reference batch (input_pipeline.py:59-133, :198-214; clickstream_transformer.py:38-63) for
    the sessions `session_idx`, with mask positions from keyed_mask_positions.  Pure-Python loops:
    small cases only.  Returns ids (B, L + 3) int32, labels (B, Mmax) float32, n_masked."""
    rows, labs = [], []
    for s in session_idx:
        ids = [int(v) for v in sessions_ids[s]]
        if mode == "train":
            ids = ids[:-1]                                         # :101-104
            # :68-70 - a float32 product, truncated (90 * 0.7 gives 63 there, 62 in float64)
            n = max(0, min(int(np.float32(len(ids)) * np.float32(masked_percentage)), max_masked))
            pos = keyed_mask_positions(seed, int(s), len(ids), n)
        else:
            pos = [len(ids) - 1]                                   # :118-121
        labs.append([float(ids[p] - num_reserved) for p in pos])
        for p in pos:
            ids[p] = mask
        rows.append(ids)
    L = max(1, max(len(r) for r in rows)) if L is None else L
    Mmax = max(1, max(len(l) for l in labs)) if Mmax is None else Mmax
    out = np.zeros((len(rows), L + 3), dtype=np.int32)
    lab = np.full((len(rows), Mmax), LABEL_PAD, dtype=np.float32)
    for b, (r, l) in enumerate(zip(rows, labs)):
        out[b, 0], out[b, 1], out[b, L + 2] = cls, sep, sep
        out[b, 2:2 + len(r)] = r
        lab[b, :len(l)] = l
    return out, lab, sum(len(l) for l in labs)


# ------------------------------------------- examples/BERT4Rec/source/utils.py:56-113 (adaptor)
def cloze_output_adaptor(y_true, y_pred):
    y_pred = y_pred.reshape(-1, y_pred.shape[-1])
    y_true = y_true.reshape(-1, 1)
    keep = (y_true[:, 0] != LABEL_PAD)
    return y_true[keep], y_pred[keep]


def sparse_categorical_crossentropy_probs(labels, probs, ce_mode="exact_tf23"):
    """K.sparse_categorical_crossentropy(from_logits=False) on a plain tensor (TF 2.3):
    clip to [1e-7, 1-1e-7], log, then softmax-CE on those log-probabilities.
    ce_mode='logits' is plain -log p_t (identical unless clipping is active)."""
    labels = labels.astype(np.int64).reshape(-1)
    rows = np.arange(len(labels))
    if ce_mode == "logits":
        return -np.log(probs[rows, labels])
    eps = probs.dtype.type(1e-7)
    pc = np.clip(probs, eps, probs.dtype.type(1.0) - eps)
    lg = np.log(pc)
    m = lg.max(-1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(lg - m).sum(-1))
    return lse - lg[rows, labels]


def binary_crossentropy_probs(y, p):
    """K.binary_crossentropy(from_logits=False) (TF 2.3): clip, then
    -(y log(p+eps) + (1-y) log(1-p+eps)), eps = 1e-7  (SURVEY.md A12)."""
    eps = p.dtype.type(1e-7)
    pc = np.clip(p, eps, p.dtype.type(1.0) - eps)
    return -(y * np.log(pc + eps) + (1 - y) * np.log(1 - pc + eps))


def masked_loss(y_true, y_pred, item_wise_loss_fn, pos_weight=None, label_pad=LABEL_PAD):
    """MaskedLoss.call (losses.py:31-98)."""
    y_true = np.asarray(y_true)
    dt = y_pred.dtype.type
    mask = (y_true != label_pad).astype(y_pred.dtype)
    y = y_true.astype(y_pred.dtype) - (1 - mask) * dt(label_pad)
    item_loss = item_wise_loss_fn(y, y_pred).reshape(y.shape)
    item_loss = item_loss * mask
    if pos_weight is not None:
        item_loss = np.where(y == 1, dt(pos_weight), dt(1.0)) * item_loss
    total, n = item_loss.sum(), mask.sum()
    if y_true.size == 0:
        mean = dt(0.0)
    else:
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = total / n
    if pos_weight is not None:
        mean = mean / ((dt(pos_weight) + dt(1.0)) / 2)
    return mean


def cloze_masked_loss(y_true, y_pred, ce_mode="exact_tf23"):
    """ClozeMaskedLoss.call (utils.py:130-134) with sparse_categorical_crossentropy (main.py:89)."""
    yt, yp = cloze_output_adaptor(np.asarray(y_true, dtype=np.float32), y_pred)
    return masked_loss(
        yt, yp, lambda y, p: sparse_categorical_crossentropy_probs(y[:, 0], p, ce_mode))


def cloze_ce_from_logits(logits, labels):
    """Logits-mode Cloze loss on flat rows: mean over rows with label != -1 of lse(z) - z_t.
    Returns (loss, dlogits, n_valid); empty / all-pad input gives loss 0 (losses.py:89-91)."""
    labels = np.asarray(labels).reshape(-1).astype(np.int64)
    valid = labels >= 0
    n = int(valid.sum())
    dz = np.zeros_like(logits)
    if n == 0:
        return logits.dtype.type(0.0), dz, 0
    z = logits[valid]
    m = z.max(-1, keepdims=True)
    e = np.exp(z - m)
    s = e.sum(-1, keepdims=True)
    lse = (m + np.log(s))[:, 0]
    t = labels[valid]
    rows = np.arange(n)
    loss = (lse - z[rows, t]).sum() / n
    p = e / s
    p[rows, t] -= 1.0
    dz[valid] = p / n
    return loss, dz, n


# --------------------------------------------------------------------- utils.py:137-259 metrics
def top_k_ids(scores, k):
    """tf.math.top_k: descending score, ties -> lower index first."""
    order = np.argsort(-scores, axis=-1, kind="stable")
    return order[:, :k]


def log_base_2_f32(x):
    return (np.log(x.astype(np.float32)) / np.log(np.float32(2.0))).astype(np.float32)


def cloze_recall_update(y_true, y_pred, k):
    """ClozeMaskedRecall.update_state (utils.py:161-187): returns (sum of hits, n_examples)."""
    yt, yp = cloze_output_adaptor(np.asarray(y_true, dtype=np.float32), y_pred)
    if len(yt) == 0:
        return np.float32(0), np.float32(0)
    ids = top_k_ids(yp, k).astype(np.float32)
    rel = (ids == yt).astype(np.float32)
    return np.float32(rel.sum(1).sum()), np.float32(len(yt))


def cloze_ndcg_update(y_true, y_pred, k):
    """ClozeMaskedNDCG.update_state (utils.py:235-252); ideal DCG = 1 (weights[:1])."""
    yt, yp = cloze_output_adaptor(np.asarray(y_true, dtype=np.float32), y_pred)
    if len(yt) == 0:
        return np.float32(0), np.float32(0)
    w = (1.0 / log_base_2_f32(np.arange(2, k + 2))).astype(np.float32)
    ids = top_k_ids(yp, k).astype(np.float32)
    gains = (ids == yt).astype(np.float32)
    dcg = (gains * w[:ids.shape[1]]).sum(1)
    ideal = (np.ones_like(yt) * w[:1]).sum(1)
    return np.float32((dcg / ideal).sum()), np.float32(len(yt))


def rank_metrics_from_topk(topk_ids, labels, k):
    """(hits, ndcg_sum, n) from already-computed top-k ids and int labels (-1 = pad)."""
    labels = np.asarray(labels).reshape(-1)
    valid = labels >= 0
    w = (1.0 / log_base_2_f32(np.arange(2, k + 2))).astype(np.float32)
    eq = (topk_ids[valid] == labels[valid, None])
    return (np.float32(eq.sum()), np.float32((eq * w[None, :topk_ids.shape[1]]).sum()),
            int(valid.sum()))


# ----------------------------------------------------------------------- optimizer (main.py:87)
def adam_step(theta, grad, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-9):
    """Keras Adam (TF 2.3): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t * m / (sqrt(v)+eps).
    t is the 1-based step count.  Dense-equivalent for embedding tables."""
    dt = theta.dtype.type
    m = dt(b1) * m + dt(1 - b1) * grad
    v = dt(b2) * v + dt(1 - b2) * grad * grad
    lr_t = dt(lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t))
    theta = theta - lr_t * m / (np.sqrt(v) + dt(eps))
    return theta, m, v


# -------------------------------------------------------- input_pipeline.py:59-133 Cloze masking
def n_masked_for(length, masked_percentage, max_masked):
    """n = clip(int(float32(len) * p), 0, max)  (input_pipeline.py:68-70)."""
    return int(min(max(int(np.float32(length) * np.float32(masked_percentage)), 0), max_masked))


# ======================================================================= whole-model reference
def init_params(rng, rows, dims, num_layers, dff, head_dims, out_vocab, head_kind="softmax",
                dtype=np.float32):
    """Keras default initialisers: Embedding U(-0.05,0.05); Dense glorot-uniform
    kernels, zero biases; LayerNorm gamma=1, beta=0."""
    def glorot(i, o):
        lim = math.sqrt(6.0 / (i + o))
        return rng.uniform(-lim, lim, size=(i, o)).astype(dtype)

    d = int(sum(dims))
    P = {}
    for f, (R, df) in enumerate(zip(rows, dims)):
        P[f"emb.{f}"] = rng.uniform(-0.05, 0.05, size=(R, df)).astype(dtype)
    for l in range(num_layers):
        for nm in ("q", "k", "v", "o"):
            P[f"enc.{l}.w{nm}"] = glorot(d, d)
            P[f"enc.{l}.b{nm}"] = np.zeros(d, dtype)
        P[f"enc.{l}.w1"] = glorot(d, dff)
        P[f"enc.{l}.b1"] = np.zeros(dff, dtype)
        P[f"enc.{l}.w2"] = glorot(dff, d)
        P[f"enc.{l}.b2"] = np.zeros(d, dtype)
        for k in ("ln1", "ln2"):
            P[f"enc.{l}.{k}_g"] = np.ones(d, dtype)
            P[f"enc.{l}.{k}_b"] = np.zeros(d, dtype)
    prev = d
    for i, hd in enumerate(head_dims):
        P[f"head.{i}.w"] = glorot(prev, hd)
        P[f"head.{i}.b"] = np.zeros(hd, dtype)
        prev = hd
    n_out = out_vocab if head_kind != "binary" else 1
    P["head.out.w"] = glorot(prev, n_out)
    P["head.out.b"] = np.zeros(n_out, dtype)
    return P


def layer_params(P, l):
    pre = f"enc.{l}."
    return {k[len(pre):]: v for k, v in P.items() if k.startswith(pre)}


def head_layers(P):
    out, i = [], 0
    while f"head.{i}.w" in P:
        out.append((P[f"head.{i}.w"], P[f"head.{i}.b"]))
        i += 1
    return out


def encoder_fwd(ids_list, P, num_layers, num_heads, pe, dtype=np.float64, masks=None,
                embed_dtype=np.float32):
    """Transformer.call (transformer.py:376-402): embed, (dropout), encoder stack.
    masks: optional dict {'in': m, (l,1): m, (l,2): m} of scaled dropout masks.
    embed_dtype: the scale-and-add of the embedding is a float32 computation in the reference and
    stays one by default even when the rest runs in float64 (the device kernel is byte-exact to
    it); np.float64 evaluates that step in double as well (the all-float64 graph of
    tests/golden/reference_*_f64.npz)."""
    F = len(ids_list)
    tables = [P[f"emb.{f}"] for f in range(F)]
    x0 = embed_fwd(ids_list, tables, pe, dtype=embed_dtype).astype(dtype)
    x = x0 * masks["in"] if masks and masks.get("in") is not None else x0
    pad = create_padding_mask(ids_list[0])
    caches = []
    for l in range(num_layers):
        p = {k: v.astype(dtype) for k, v in layer_params(P, l).items()}
        x, c = encoder_layer_fwd(x, pad, p, num_heads,
                                 masks.get((l, 1)) if masks else None,
                                 masks.get((l, 2)) if masks else None)
        caches.append((c, p))
    return x, caches


def cloze_train_step(ids_list, labels, P, num_layers, num_heads, pe, dtype=np.float64,
                     masks=None, embed_dtype=np.float32):
    """Forward + backward of the BERT4Rec Cloze model (value_to_head='[MASK]', SoftMaxHead,
    ClozeMaskedLoss) in logits mode.  labels: (B, Mmax) float/ints padded with -1, aligned with
    the (b,s)-ordered [MASK] positions.  Returns loss, grads dict (same keys as P), extras."""
    F = len(ids_list)
    x, caches = encoder_fwd(ids_list, P, num_layers, num_heads, pe, dtype, masks, embed_dtype)
    B, S, d = x.shape
    sel, index = select_masked(ids_list[0], x)
    Mmax = sel.shape[1]
    layers = [(w.astype(dtype), b.astype(dtype)) for w, b in head_layers(P)]
    w_out, b_out = P["head.out.w"].astype(dtype), P["head.out.b"].astype(dtype)
    flat = sel.reshape(B * Mmax, d)
    h, acts = mlp_fwd(flat, layers)
    logits = h @ w_out + b_out
    lab = np.asarray(labels).reshape(-1)
    loss, dz, n = cloze_ce_from_logits(logits, lab)
    G = {}
    G["head.out.w"] = h.T @ dz
    G["head.out.b"] = dz.sum(0)
    dh = dz @ w_out.T
    dflat, hg = mlp_bwd(dh, acts, layers)
    for i, (dw, db) in enumerate(hg):
        G[f"head.{i}.w"], G[f"head.{i}.b"] = dw, db
    dsel = dflat.reshape(B, Mmax, d)
    dx = np.zeros_like(x)
    cnt = np.zeros(B, dtype=np.int64)
    for (b, s) in index:
        dx[b, s] = dsel[b, cnt[b]]
        cnt[b] += 1
    for l in reversed(range(num_layers)):
        c, p = caches[l]
        dx, g = encoder_layer_bwd(dx, c, p)
        for k, v in g.items():
            G[f"enc.{l}.{k}"] = v
    if masks and masks.get("in") is not None:
        dx = dx * masks["in"]
    dims = [P[f"emb.{f}"].shape[1] for f in range(F)]
    rows = [P[f"emb.{f}"].shape[0] for f in range(F)]
    for f, g in enumerate(embed_bwd(dx, ids_list, dims, rows, dtype)):
        G[f"emb.{f}"] = g
    extras = dict(logits=logits, hidden=h, enc_out=x, index=index, n_valid=n)
    return loss, G, extras


def segment_binary_train_step(ids_list, y_true, P, num_layers, num_heads, pe, segment,
                              pos_weight=None, dtype=np.float64, embed_dtype=np.float32,
                              head_kind="binary"):
    """Forward + backward of the multi-variable click-path classifier: encoder
    (transformer.py:376-402) -> rows of segment `segment` (clickstream_transformer.py:317-322)
    -> BinaryClassificationHead (head.py:4-26) -> MaskedLoss(K.binary_crossentropy, pos_weight)
    (losses.py:31-98), and what TF autodiff derives from it.  y_true: (B, segment length) padded
    with LABEL_PAD.  Returns loss, grads dict (same keys as P), extras (probs).
    head_kind='multilabel': MultiLabel_MultiClass_classification (head.py:50-69) on a length-1
    segment instead, y_true (B, classes)."""
    F = len(ids_list)
    x, caches = encoder_fwd(ids_list, P, num_layers, num_heads, pe, dtype, None, embed_dtype)
    starts, ends = segment_bounds(ids_list[0][0])
    s0, s1 = int(starts[segment]), int(ends[segment])
    seg = x[:, s0:s1, :]
    layers = [(w.astype(dtype), b.astype(dtype)) for w, b in head_layers(P)]
    w_out, b_out = P["head.out.w"].astype(dtype), P["head.out.b"].astype(dtype)
    loss_and_grads, fwd = ((binary_head_loss_and_grads, binary_head_fwd) if head_kind == "binary"
                           else (multilabel_head_loss_and_grads, multilabel_head_fwd))
    loss, dseg, lg, dWo, dbo = loss_and_grads(seg, layers, w_out, b_out, y_true, pos_weight=pos_weight)
    probs, _, _ = fwd(seg, layers, w_out, b_out)
    G = {"head.out.w": dWo, "head.out.b": dbo}
    for i, (dw, db) in enumerate(lg):
        G[f"head.{i}.w"], G[f"head.{i}.b"] = dw, db
    dx = np.zeros_like(x)
    dx[:, s0:s1, :] = dseg
    for l in reversed(range(num_layers)):
        c, p = caches[l]
        dx, g = encoder_layer_bwd(dx, c, p)
        for k, v in g.items():
            G[f"enc.{l}.{k}"] = v
    dims = [P[f"emb.{f}"].shape[1] for f in range(F)]
    rows = [P[f"emb.{f}"].shape[0] for f in range(F)]
    for f, g in enumerate(embed_bwd(dx, ids_list, dims, rows, dtype)):
        G[f"emb.{f}"] = g
    return loss, G, dict(probs=probs, enc_out=x)


# ------------------------------------------------------------------ binary-task metrics
def binary_metric_counts(y_true, y_pred, label_pad=-1.0):
    """Accumulators of clickstream_transformer/metrics.py as one vector:
    (sum mask, sum y*mask [PositiveRate :12-20], sum round(p)*mask [PredictedPositives :36-45],
    tp, condition_true, predicted_true [F1Score :63-78: int32 casts, no mask]).  tf.round rounds
    half to even, as np.rint does."""
    yt = np.asarray(y_true, dtype=np.float32).reshape(-1)
    yp = np.asarray(y_pred, dtype=np.float32).reshape(-1)
    mask = (yt != np.float32(label_pad)).astype(np.float32)
    pr = np.rint(yp)
    ct = yt.astype(np.int32) == 1
    pt = pr.astype(np.int32) == 1
    return np.array([mask.sum(), (yt * mask).sum(), (pr * mask).sum(), (ct & pt).sum(), ct.sum(), pt.sum()],
                    dtype=np.float64)
