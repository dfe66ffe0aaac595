"""Thin Python wrappers over the libb4cp C ABI.

torch is used only for device memory and streams: every function here launches hand-written
sm_100a kernels on torch's current stream and returns torch tensors that alias plain device
buffers.  Nothing in this module computes with torch operators.
"""
import ctypes

import torch

from . import _lib as L

BF16 = torch.bfloat16
F32 = torch.float32
I32 = torch.int32


def ld8(n):
    """Leading dimension of a bf16 matrix: TMA needs 16-byte row strides."""
    return (int(n) + 7) // 8 * 8


def _dev(t):
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return t


def empty(shape, dtype=F32):
    return torch.empty(shape, dtype=dtype, device="cuda")


def zeros(shape, dtype=F32):
    return torch.zeros(shape, dtype=dtype, device="cuda")


class Workspace:
    """Grow-only byte buffers keyed by name (the C ABI never allocates)."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, nbytes):
        nbytes = max(int(nbytes), 256)
        buf = self._bufs.get(name)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            self._bufs[name] = buf
        return buf


WS = Workspace()


class KernelTimer:
    """CUDA-event brackets around selected launches on the current stream (bench.py roofline)."""

    def __init__(self):
        self.enabled = False
        self.events = {}

    def begin(self, tag):
        if not self.enabled:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def end(self, tag, e0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.events.setdefault(tag, []).append((e0, e1))

    def totals_ms(self):
        """{tag: (total ms, launches)} — call after a synchronize."""
        return {t: (sum(a.elapsed_time(b) for a, b in ev), len(ev)) for t, ev in self.events.items()}

    def reset(self):
        self.events = {}


TIMER = KernelTimer()


# ------------------------------------------------------------------------------------ GEMM
def split3(x, mn, rows_mn, K, order, ws_name):
    """fp32 operand -> its bf16 (hi | hi | lo) [order 0] / (hi | lo | hi) [order 1] split along the
    contraction axis (b4cp_split_bf16x3).  mn = 0: x is [rows_mn][K] (K contiguous) -> bf16
    [rows_mn][3*ld8(K)]; mn = 1: x is [K][rows_mn] -> bf16 [3*ld8(K)][ld8(rows_mn)]."""
    kp = ld8(K)
    if mn == 0:
        out = WS.get(ws_name, rows_mn * 3 * kp * 2).view(BF16)[: rows_mn * 3 * kp].view(rows_mn, 3 * kp)
        L.call("b4cp_split_bf16x3", L.ptr(x), L.c_long(rows_mn), L.c_int(K), L.c_long(x.stride(0)),
               L.ptr(out), L.c_long(3 * kp), L.c_int(0), L.c_int(order), L.stream_ptr())
    else:
        ldo = ld8(rows_mn)
        out = WS.get(ws_name, 3 * kp * ldo * 2).view(BF16)[: 3 * kp * ldo].view(3 * kp, ldo)
        L.call("b4cp_split_bf16x3", L.ptr(x), L.c_long(K), L.c_int(rows_mn), L.c_long(x.stride(0)),
               L.ptr(out), L.c_long(ldo), L.c_int(1), L.c_int(order), L.stream_ptr())
    return out


def x3_operands(A, a_mn, B, b_mn, M, N, K):
    """fp32-class mode: both operands fp32 -> (A3, B3, K3) for ONE bf16 GEMM over K3 = 3*ld8(K)
    that accumulates hi*hi + hi*lo + lo*hi in fp32."""
    assert A.dtype == F32 and B.dtype == F32, "fp32-class GEMM needs both operands in fp32"
    return (split3(A, a_mn, M, K, 0, "x3A"), split3(B, b_mn, N, K, 1, "x3B"), 3 * ld8(K))


def _gate_bf16(gate, N):
    """ReLU-backward gates are tested for > 0 only: a bf16 copy of a fp32 gate is sign-exact."""
    if gate is None or gate.dtype == BF16:
        return gate
    rows = gate.shape[0]
    buf = WS.get("x3gate", rows * ld8(N) * 2).view(BF16)[: rows * ld8(N)].view(rows, ld8(N))
    return cast_bf16(gate, N, out=buf)


def gemm(A, a_mn, B, b_mn, M, N, K, *, bias=None, relu=False, gate=None, addend=None,
         out_f32=None, out_bf16=None, alpha=1.0, splits=1, lda=None, ldb=None):
    """C[M,N] = epilogue(alpha * sum_k A(m,k) B(n,k)); see include/b4cp.h.  fp32 operands select
    the fp32-class product (bf16 x 3 split along K, same tcgen05 kernel); an fp32 tensor passed as
    `out_bf16` is then simply the fp32 output."""
    if A.dtype == F32 or B.dtype == F32:
        A, B, K = x3_operands(A, a_mn, B, b_mn, M, N, K)
        lda = ldb = None
    gate = _gate_bf16(gate, N)
    if out_bf16 is not None and out_bf16.dtype == F32:
        assert out_f32 is None or out_f32.data_ptr() == out_bf16.data_ptr()
        out_f32, out_bf16 = out_bf16, None
    ep = L.GemmEpilogue()
    ep.alpha = alpha
    ep.bias = bias.data_ptr() if bias is not None else None
    ep.relu = 1 if relu else 0
    if gate is not None:
        ep.gate = gate.data_ptr()
        ep.ld_gate = gate.stride(0)
    if addend is not None:
        ep.addend = addend.data_ptr()
        ep.ld_addend = addend.stride(0)
    if out_f32 is not None:
        ep.out_f32 = out_f32.data_ptr()
        ep.ld_f32 = out_f32.stride(-2)
        ep.split_stride = out_f32.stride(0) if out_f32.dim() == 3 else 0
    if out_bf16 is not None:
        ep.out_bf16 = out_bf16.data_ptr()
        ep.ld_bf16 = out_bf16.stride(0)
    lda = A.stride(0) if lda is None else lda
    ldb = B.stride(0) if ldb is None else ldb
    L.call("b4cp_gemm_bf16", L.ptr(A), L.c_int(a_mn), L.c_long(lda), L.ptr(B), L.c_int(b_mn),
           L.c_long(ldb), L.c_int(M), L.c_int(N), L.c_int(K), L.c_int(splits), ctypes.byref(ep),
           L.stream_ptr())


def gemm_splits_for(M, N, K):
    return L.lib().b4cp_gemm_splits_for(int(M), int(N), int(K))


def reduce_splits(partials, out):
    splits = partials.shape[0]
    n = out.numel()
    L.call("b4cp_reduce_splits", L.ptr(partials), L.c_int(splits), L.c_long(n),
           L.c_long(partials.stride(0)), L.ptr(out), L.stream_ptr())


def gemm_splitk(A, a_mn, B, b_mn, M, N, K, out_f32, ws_name="splitk"):
    """Deterministic split-K product into out_f32 [M,N] (used for weight gradients)."""
    if A.dtype == F32 or B.dtype == F32:
        A, B, K = x3_operands(A, a_mn, B, b_mn, M, N, K)
    splits = gemm_splits_for(M, N, K)
    if splits <= 1:
        gemm(A, a_mn, B, b_mn, M, N, K, out_f32=out_f32)
        return
    buf = WS.get(ws_name, splits * M * N * 4).view(F32)[: splits * M * N].view(splits, M, N)
    gemm(A, a_mn, B, b_mn, M, N, K, out_f32=buf, splits=splits)
    reduce_splits(buf, out_f32)


def gemm_splitk_ex(A, a_mn, B, b_mn, M, N, K, gate=None, out_f32=None, out_bf16=None,
                   ws_name="splitk_dx"):
    """Split-K product whose reduction applies a ReLU-backward gate and writes fp32 and/or bf16
    (dx = dz W^T with K = V long and M x N small)."""
    if A.dtype == F32 or B.dtype == F32:
        A, B, K = x3_operands(A, a_mn, B, b_mn, M, N, K)
    splits = gemm_splits_for(M, N, K)
    part = WS.get(ws_name, splits * M * N * 4).view(F32)[: splits * M * N].view(splits, M, N)
    gemm(A, a_mn, B, b_mn, M, N, K, out_f32=part, splits=splits)
    reduce_splits_ex(part, M, N, gate, out_f32, out_bf16)


def reduce_splits_ex(partials, M, N, gate=None, out_f32=None, out_bf16=None):
    gate = _gate_bf16(gate, N)
    if out_bf16 is not None and out_bf16.dtype == F32:
        assert out_f32 is None or out_f32.data_ptr() == out_bf16.data_ptr()
        out_f32, out_bf16 = out_bf16, None
    splits = partials.shape[0]
    L.call("b4cp_reduce_splits_ex", L.ptr(partials), L.c_int(splits), L.c_long(M), L.c_int(N),
           L.c_long(partials.stride(0)), L.ptr(gate),
           L.c_long(gate.stride(0) if gate is not None else 0), L.ptr(out_f32),
           L.c_long(out_f32.stride(0) if out_f32 is not None and out_f32.dim() == 2 else N),
           L.ptr(out_bf16), L.c_long(out_bf16.stride(0) if out_bf16 is not None else 0),
           L.stream_ptr())


def cast_bf16(x_f32, cols=None, out=None):
    """fp32 [rows, cols] -> bf16 [rows, ld8(cols)] (zero padded)."""
    rows = x_f32.shape[0]
    cols = x_f32.shape[1] if cols is None else cols
    if out is None:
        out = empty((rows, ld8(cols)), BF16)
    L.call("b4cp_cast_f32_bf16", L.ptr(x_f32), L.c_long(rows), L.c_int(cols),
           L.c_long(x_f32.stride(0)), L.ptr(out), L.c_long(out.stride(0)), L.stream_ptr())
    return out


# ------------------------------------------------------------------------------- embedding
def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def _int_array(vals):
    arr = (ctypes.c_int * len(vals))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def embed_fwd(ids_list, tables, pe, B, S, *, dropout_rate=0.0, seed=0, site=0, out_f32=None,
              out_bf16=None):
    F = len(ids_list)
    dims = [t.shape[1] for t in tables]
    rows = [t.shape[0] for t in tables]
    d = sum(dims)
    assert pe.shape[0] >= S and pe.shape[1] == d
    if out_f32 is None and out_bf16 is None:
        out_f32 = empty((B * S, d))
    L.call("b4cp_embed_fwd", _ptr_array(ids_list), _ptr_array(tables), _int_array(dims),
           _int_array(rows), L.c_int(F), L.ptr(pe), L.c_int(B), L.c_int(S),
           L.c_float(dropout_rate), L.c_u64(seed), ctypes.c_uint32(site), L.ptr(out_f32),
           L.ptr(out_bf16), L.stream_ptr())
    return out_f32, out_bf16


def embed_bwd(dout, d_model, col_offset, dim, ids, rows, table_grad, *, dropout_rate=0.0, seed=0,
              site=0, uniq_ids=None, n_unique=None):
    tokens = ids.numel()
    nbytes = L.lib().b4cp_embed_bwd_workspace_bytes
    nbytes.restype = ctypes.c_long
    need = nbytes(ctypes.c_long(tokens), ctypes.c_int(dim))
    ws = WS.get("embed_bwd", need)
    L.call("b4cp_embed_bwd", L.ptr(dout), L.c_int(d_model), L.c_int(col_offset), L.c_int(dim),
           L.ptr(ids), L.c_long(tokens), L.c_int(rows), L.c_float(dropout_rate), L.c_u64(seed),
           ctypes.c_uint32(site), L.ptr(table_grad), L.ptr(uniq_ids), L.ptr(n_unique), L.ptr(ws),
           L.c_long(ws.numel()), L.stream_ptr())


def _embed_ws(tokens, dim, name):
    nbytes = L.lib().b4cp_embed_bwd_workspace_bytes
    nbytes.restype = ctypes.c_long
    return WS.get(name, nbytes(ctypes.c_long(tokens), ctypes.c_int(dim)))


def embed_sort(ids, rows, dim, ws_name):
    """First half of embed_bwd: stable sort of (id, token) into the named workspace."""
    ws = _embed_ws(ids.numel(), dim, ws_name)
    L.call("b4cp_embed_sort", L.ptr(ids), L.c_long(ids.numel()), L.c_int(rows), L.c_int(dim),
           L.ptr(ws), L.c_long(ws.numel()), L.stream_ptr())


def embed_bwd_sorted(dout, d_model, col_offset, dim, tokens, rows, table_grad, ws_name, *,
                     dropout_rate=0.0, seed=0, site=0):
    """Second half: segment sums over the order `embed_sort` left in the named workspace."""
    ws = _embed_ws(tokens, dim, ws_name)
    L.call("b4cp_embed_bwd_sorted", L.ptr(dout), L.c_int(d_model), L.c_int(col_offset), L.c_int(dim),
           L.c_long(tokens), L.c_int(rows), L.c_float(dropout_rate), L.c_u64(seed),
           ctypes.c_uint32(site), L.ptr(table_grad), L.ptr(ws), L.c_long(ws.numel()), L.stream_ptr())


# --------------------------------------------------------------------------------- encoder
def attention_fwd(qkv, ids_first, B, S, H, dh, out, lse):
    if qkv.dtype == F32:   # fp32-class mode
        L.call("b4cp_attention_f32_fwd", L.ptr(qkv), L.ptr(ids_first), L.c_int(B), L.c_int(S),
               L.c_int(H), L.c_int(dh), L.ptr(out), L.ptr(lse), L.stream_ptr())
        return
    L.call("b4cp_attention_fwd", L.ptr(qkv), L.ptr(ids_first), L.c_int(B), L.c_int(S), L.c_int(H),
           L.c_int(dh), L.ptr(out), L.ptr(lse), L.stream_ptr())


def attention_bwd(qkv, dout, lse, ids_first, B, S, H, dh, dqkv, out=None):
    """`out`: the forward output; needed by the tensor-core path for 128 < S <= 256."""
    if qkv.dtype == F32:
        L.call("b4cp_attention_f32_bwd", L.ptr(qkv), L.ptr(dout), L.ptr(lse), L.ptr(ids_first),
               L.c_int(B), L.c_int(S), L.c_int(H), L.c_int(dh), L.ptr(dqkv), L.stream_ptr())
        return
    L.call("b4cp_attention_bwd", L.ptr(qkv), L.ptr(out), L.ptr(dout), L.ptr(lse), L.ptr(ids_first),
           L.c_int(B), L.c_int(S), L.c_int(H), L.c_int(dh), L.ptr(dqkv), L.stream_ptr())


def residual_ln_fwd(x, r, gamma, beta, y_f32, y_bf16, *, dropout_rate=0.0, seed=0, site=0):
    T, d = x.shape
    if y_bf16 is not None and y_bf16.dtype == F32:   # fp32-class mode: one fp32 output serves both
        assert y_f32 is None or y_f32.data_ptr() == y_bf16.data_ptr()
        y_f32, y_bf16 = y_bf16, None
    L.call("b4cp_residual_ln_fwd", L.ptr(x), L.ptr(r), L.c_long(T), L.c_int(d), L.ptr(gamma),
           L.ptr(beta), L.c_float(dropout_rate), L.c_u64(seed), ctypes.c_uint32(site),
           L.ptr(y_f32), L.ptr(y_bf16), L.c_long(y_bf16.stride(0) if y_bf16 is not None else 0),
           L.stream_ptr())


def residual_ln_bwd(dy, x, r, gamma, dx, dr_bf16, dgamma, dbeta, dbias, *, dropout_rate=0.0,
                    seed=0, site=0):
    T, d = x.shape
    fn = L.lib().b4cp_residual_ln_bwd_workspace_bytes
    fn.restype = ctypes.c_long
    ws = WS.get("ln_bwd", fn(ctypes.c_int(d)))
    dr_f32 = None
    if dr_bf16 is not None and dr_bf16.dtype == F32:   # fp32-class mode
        dr_f32, dr_bf16 = dr_bf16, None
    L.call("b4cp_residual_ln_bwd", L.ptr(dy), L.ptr(x), L.ptr(r), L.c_long(T), L.c_int(d),
           L.ptr(gamma), L.c_float(dropout_rate), L.c_u64(seed), ctypes.c_uint32(site), L.ptr(dx),
           L.ptr(dr_bf16), L.c_long(dr_bf16.stride(0) if dr_bf16 is not None else 0),
           L.ptr(dr_f32), L.ptr(dgamma), L.ptr(dbeta), L.ptr(dbias), L.ptr(ws), L.stream_ptr())


def colsum_bf16(x_bf16, T, n, out, ws_name="colsum"):
    fn = L.lib().b4cp_colsum_workspace_bytes
    fn.restype = ctypes.c_long
    ws = WS.get(ws_name, fn(ctypes.c_long(T), ctypes.c_int(n)))
    if x_bf16.dtype == F32:
        L.call("b4cp_colsum_f32", L.ptr(x_bf16), L.c_long(T), L.c_int(n), L.c_long(x_bf16.stride(0)),
               L.ptr(out), L.ptr(ws), L.stream_ptr())
        return
    L.call("b4cp_colsum_bf16", L.ptr(x_bf16), L.c_long(T), L.c_int(n),
           L.c_long(x_bf16.stride(0)), L.ptr(out), L.ptr(ws), L.stream_ptr())


def device_seed(counter_i32):
    """Seed argument that makes kernels read the seed from a device int32 at run time
    (B4CP_SEED_FROM_DEVICE): used by CUDA-graph replays."""
    return (1 << 63) | int(counter_i32.data_ptr())


def dropout_mask(n, rate, seed, site):
    out = empty((n,))
    L.call("b4cp_dropout_mask", L.ptr(out), L.c_long(n), L.c_float(rate), L.c_u64(seed),
           ctypes.c_uint32(site), L.stream_ptr())
    return out


def dropout_apply(x, rate, seed, site):
    L.call("b4cp_dropout_apply", L.ptr(x), L.c_long(x.numel()), L.c_float(rate), L.c_u64(seed),
           ctypes.c_uint32(site), L.stream_ptr())


# ------------------------------------------------------------------------------- selection
def select_masked(ids_first, value, capacity, row_index=None, count=None):
    """`row_index` (int32 [>= max(capacity, 1)]) / `count` (int32 [1]): optional caller-owned
    outputs (a caller that runs this on a second stream must not let it allocate there)."""
    tokens = ids_first.numel()
    fn = L.lib().b4cp_select_workspace_bytes
    fn.restype = ctypes.c_long
    ws = WS.get("select", fn(ctypes.c_long(tokens)))
    row_index = empty((max(capacity, 1),), I32) if row_index is None else row_index
    count = empty((1,), I32) if count is None else count
    L.call("b4cp_select_masked", L.ptr(ids_first), L.c_long(tokens), L.c_int(value),
           L.ptr(row_index), L.c_long(capacity), L.ptr(count), L.ptr(ws), L.stream_ptr())
    return row_index[:capacity], count


def compact_labels(labels_f32, capacity, label_pad=-1.0, out=None, count=None):
    """(B, Mmax) float32 labels padded with -1 -> int32 [capacity] valid labels, -1 padded.
    `out` / `count`: optional caller-owned outputs (see select_masked)."""
    n = labels_f32.numel()
    fn = L.lib().b4cp_select_workspace_bytes
    fn.restype = ctypes.c_long
    ws = WS.get("select", fn(ctypes.c_long(max(n, 1))))
    out = empty((max(capacity, 1),), I32) if out is None else out
    count = empty((1,), I32) if count is None else count
    L.call("b4cp_compact_labels", L.ptr(labels_f32), L.c_long(n), L.c_float(label_pad), L.ptr(out),
           L.c_long(capacity), L.ptr(count), L.ptr(ws), L.stream_ptr())
    return out[:capacity], count


def gather_rows(x, row_index, out_f32=None, out_bf16=None):
    M = row_index.numel()
    if out_bf16 is not None and out_bf16.dtype == F32:
        assert out_f32 is None and out_bf16.stride(0) == x.shape[1]
        out_f32, out_bf16 = out_bf16, None
    d = x.shape[1]
    L.call("b4cp_gather_rows", L.ptr(x), L.c_int(d), L.ptr(row_index), L.c_long(M), L.ptr(out_f32),
           L.ptr(out_bf16), L.c_long(out_bf16.stride(0) if out_bf16 is not None else 0),
           L.stream_ptr())


def scatter_rows(src, row_index, dst):
    M = row_index.numel()
    L.call("b4cp_scatter_rows", L.ptr(src), L.c_int(src.shape[1]), L.ptr(row_index), L.c_long(M),
           L.ptr(dst), L.stream_ptr())


# ------------------------------------------------------------------------- loss and metrics
def ce_rows_stats(logits, V, labels, lse, tgt):
    M = logits.shape[0]
    L.call("b4cp_ce_rows_stats", L.ptr(logits), L.c_long(logits.stride(0)), L.c_long(M),
           L.c_int(V), L.ptr(labels), L.ptr(lse), L.ptr(tgt), L.stream_ptr())


def ce_loss_reduce(lse, tgt, labels, loss_stats, n_global=None):
    """n_global: int32 device scalar holding the valid-row count over all ranks (see b4cp.h)."""
    if n_global is not None:
        L.call("b4cp_ce_loss_reduce_n", L.ptr(lse), L.ptr(tgt), L.ptr(labels),
               L.c_long(labels.numel()), L.ptr(n_global), L.ptr(loss_stats), L.stream_ptr())
        return
    L.call("b4cp_ce_loss_reduce", L.ptr(lse), L.ptr(tgt), L.ptr(labels), L.c_long(labels.numel()),
           L.ptr(loss_stats), L.stream_ptr())


def ce_rows_grad(logits, V, labels, lse, loss_stats, dz_bf16=None, probs=None):
    M = logits.shape[0]
    L.call("b4cp_ce_rows_grad", L.ptr(logits), L.c_long(logits.stride(0)), L.c_long(M), L.c_int(V),
           L.ptr(labels), L.ptr(lse), L.ptr(loss_stats), L.ptr(dz_bf16),
           L.c_long(dz_bf16.stride(0) if dz_bf16 is not None else 0), L.ptr(probs),
           L.c_long(probs.stride(0) if probs is not None else 0), L.stream_ptr())


def ce_rows_grad_f32(logits, V, labels, lse, loss_stats):
    """fp32-class mode: dz = (softmax - onehot) / n in place over the fp32 logits [M][ld]."""
    M = logits.shape[0]
    L.call("b4cp_ce_rows_grad_f32", L.ptr(logits), L.c_long(logits.stride(0)), L.c_long(M),
           L.c_int(V), L.ptr(labels), L.ptr(lse), L.ptr(loss_stats), L.stream_ptr())
    return logits


def topk_rows(scores, V, k, out_ids=None, out_scores=None):
    rows = scores.shape[0]
    if out_ids is None:
        out_ids = empty((rows, k), I32)
    L.call("b4cp_topk_rows", L.ptr(scores), L.c_long(scores.stride(0)), L.c_long(rows), L.c_int(V),
           L.c_int(k), L.ptr(out_ids), L.ptr(out_scores), L.c_long(out_ids.stride(0)),
           L.stream_ptr())
    return out_ids, out_scores


def rank_metrics(topk_ids, k, labels, counters):
    L.call("b4cp_rank_metrics", L.ptr(topk_ids), L.c_long(topk_ids.shape[0]), L.c_int(k),
           L.c_long(topk_ids.stride(0)), L.ptr(labels), L.ptr(counters), L.stream_ptr())


# ------------------------------------------------------------------------------- optimizer
def adam_step(theta, grad, m, v, *, lr, beta1=0.9, beta2=0.999, eps=1e-9, step_dev=None,
              step_host=0, grad_scale=1.0, shadow=None, cols=0):
    n = theta.numel()
    L.call("b4cp_adam_step", L.ptr(theta), L.ptr(grad), L.ptr(m), L.ptr(v), L.c_long(n),
           L.c_float(lr), L.c_float(beta1), L.c_float(beta2), L.c_float(eps), L.ptr(step_dev),
           L.c_int(step_host), L.c_float(grad_scale), L.ptr(shadow), L.c_int(cols),
           L.c_long(shadow.stride(0) if shadow is not None else 0), L.stream_ptr())


def adam_flat(theta, grad, m, v, segs, n_segs, *, lr, beta1=0.9, beta2=0.999, eps=1e-9, step_dev=None,
              step_host=0, grad_scale=1.0):
    """One Adam sweep over the whole flat parameter buffer (+ the bf16 shadows listed in segs)."""
    L.call("b4cp_adam_flat", L.ptr(theta), L.ptr(grad), L.ptr(m), L.ptr(v), L.c_long(theta.numel()),
           L.c_float(lr), L.c_float(beta1), L.c_float(beta2), L.c_float(eps), L.ptr(step_dev),
           L.c_int(step_host), L.c_float(grad_scale), segs, L.c_int(n_segs), L.stream_ptr())


def zero_(t):
    """Stream-ordered zero fill of a contiguous device tensor (no torch kernel in the hot path)."""
    L.call("b4cp_zero", L.ptr(t), L.c_long(t.numel() * t.element_size()), L.stream_ptr())
    return t


def step_increment(step_dev):
    L.call("b4cp_step_increment", L.ptr(step_dev), L.stream_ptr())


def sigmoid(z, out=None):
    if out is None:
        out = torch.empty_like(z)
    L.call("b4cp_sigmoid", L.ptr(z), L.ptr(out), L.c_long(z.numel()), L.stream_ptr())
    return out


def clip_log(p, lo=1e-7, hi=1.0 - 1e-7):
    out = torch.empty_like(p)
    L.call("b4cp_clip_log", L.ptr(p), L.ptr(out), L.c_long(p.numel()), L.c_float(lo), L.c_float(hi),
           L.stream_ptr())
    return out


def masked_bce(y_true, probs, label_pad, pos_weight=None):
    stats = empty((2,))
    L.call("b4cp_masked_bce", L.ptr(y_true), L.ptr(probs), L.c_long(y_true.numel()),
           L.c_float(label_pad), L.c_float(pos_weight if pos_weight is not None else 1.0),
           L.c_int(0 if pos_weight is None else 1), L.ptr(stats), L.stream_ptr())
    return stats


def binary_metric_counts(y_true, probs, label_pad, counters):
    L.call("b4cp_binary_metric_counts", L.ptr(y_true), L.ptr(probs), L.c_long(y_true.numel()),
           L.c_float(label_pad), L.ptr(counters), L.stream_ptr())


def binary_head_bwd(y_true, probs, label_pad, pos_weight, stats, ab, h, w_out, gated, dz,
                    dab_f32=None, dab_bf16=None, dw=None, db=None):
    """Backward of Dense(1, sigmoid) + MaskedLoss(binary_crossentropy) (head.py:11, losses.py:31-98)."""
    M = y_true.numel()
    L.call("b4cp_binary_head_bwd", L.ptr(y_true), L.ptr(probs), L.c_long(M), L.c_float(label_pad),
           L.c_float(pos_weight if pos_weight is not None else 1.0),
           L.c_int(0 if pos_weight is None else 1), L.ptr(stats), L.ptr(ab), L.c_long(ab.stride(0)),
           L.c_int(h), L.ptr(w_out), L.c_int(1 if gated else 0), L.ptr(dz), L.ptr(dab_f32),
           L.ptr(dab_bf16), L.c_long(dab_bf16.stride(0) if dab_bf16 is not None else 0), L.ptr(dw),
           L.ptr(db), L.stream_ptr())


def sigmoid_bce_dz(y_true, probs, rows, cols, label_pad, pos_weight, stats, dz_f32=None,
                   dz_bf16=None):
    """Item-wise gradient of MaskedLoss(binary_crossentropy) through a (rows, cols) sigmoid output."""
    L.call("b4cp_sigmoid_bce_dz", L.ptr(y_true), L.ptr(probs), L.c_long(rows), L.c_int(cols),
           L.c_float(label_pad), L.c_float(pos_weight if pos_weight is not None else 1.0),
           L.c_int(0 if pos_weight is None else 1), L.ptr(stats), L.ptr(dz_f32), L.ptr(dz_bf16),
           L.c_long(dz_bf16.stride(0) if dz_bf16 is not None else 0), L.stream_ptr())


def cloze_build(items, offsets, session_idx, B, max_items, Mmax, train, masked_percentage,
                max_masked, seed, special, label_pad, ids, labels, n_masked, status):
    """On-device Cloze batch builder (include/b4cp.h); special = (cls, sep, mask, pad, label_offset)."""
    L.call("b4cp_cloze_build", L.ptr(items), L.ptr(offsets), L.ptr(session_idx), L.c_int(B),
           L.c_int(max_items), L.c_int(Mmax), L.c_int(1 if train else 0),
           ctypes.c_double(float(masked_percentage)), L.c_int(max_masked), L.c_u64(seed),
           *[L.c_int(v) for v in special], L.c_float(label_pad), L.ptr(ids), L.ptr(labels),
           L.ptr(n_masked), L.ptr(status), L.stream_ptr())


def cloze_position_key(seed, session, pos):
    """Host helper: the 64-bit key the builder ranks mask positions by (no GPU needed)."""
    fn = L.lib().b4cp_cloze_position_key
    fn.restype = ctypes.c_uint64
    return int(fn(L.c_u64(seed), L.c_u64(session), L.c_u64(pos)))


def _vocab_ws(M, V, h):
    fn = L.lib().b4cp_vocab_ce_workspace_bytes
    fn.restype = ctypes.c_long
    return WS.get("vocab_ce", fn(ctypes.c_long(M), ctypes.c_int(V), ctypes.c_int(h)))


def vocab_ce_fwd(xb, M, h, wb, bias, V, labels, lse, tgt, want_dx=False):
    ws = _vocab_ws(M, V, h)
    t0 = TIMER.begin("k:vocab_ce_fwd")
    L.call("b4cp_vocab_ce_fwd", L.ptr(xb), L.c_long(xb.stride(0)), L.c_long(M), L.c_int(h),
           L.ptr(wb), L.c_long(wb.stride(0)), L.ptr(bias), L.c_int(V), L.ptr(labels),
           L.c_int(1 if want_dx else 0), L.ptr(lse), L.ptr(tgt), L.ptr(ws), L.stream_ptr())
    TIMER.end("k:vocab_ce_fwd", t0)


def vocab_ce_dx(M, h, V, labels, loss_stats, wb, gate=None, out_f32=None, out_bf16=None,
                lse_global=None):
    ws = _vocab_ws(M, V, h)
    L.call("b4cp_vocab_ce_dx", L.c_long(M), L.c_int(h), L.c_int(V), L.ptr(labels),
           L.ptr(loss_stats), L.ptr(lse_global), L.ptr(wb), L.c_long(wb.stride(0)), L.ptr(gate),
           L.c_long(gate.stride(0) if gate is not None else 0), L.ptr(out_f32), L.ptr(out_bf16),
           L.c_long(out_bf16.stride(0) if out_bf16 is not None else 0), L.ptr(ws), L.stream_ptr())


def vocab_ce_bwd(xb, M, h, wb, bias, V, labels, lse, loss_stats, dW, db):
    t0 = TIMER.begin("k:vocab_ce_bwd")
    L.call("b4cp_vocab_ce_bwd", L.ptr(xb), L.c_long(xb.stride(0)), L.c_long(M), L.c_int(h),
           L.ptr(wb), L.c_long(wb.stride(0)), L.ptr(bias), L.c_int(V), L.ptr(labels), L.ptr(lse),
           L.ptr(loss_stats), L.ptr(dW), L.ptr(db), L.stream_ptr())
    TIMER.end("k:vocab_ce_bwd", t0)


def shard_labels(labels, v_begin, v_count, out=None):
    M = labels.numel()
    if out is None:
        out = empty((M,), I32)
    L.call("b4cp_shard_labels", L.ptr(labels), L.c_long(M), L.c_int(v_begin), L.c_int(v_count),
           L.ptr(out), L.stream_ptr())
    return out


def lse_merge(parts, out=None):
    R, M = parts.shape
    if out is None:
        out = empty((M,))
    L.call("b4cp_lse_merge", L.ptr(parts), L.c_int(R), L.c_long(M), L.ptr(out), L.stream_ptr())
    return out


def score_topk(xb, M, h, wb, bias, V, k, out_ids=None, out_scores=None, id_base=0, V_total=0):
    """Fused scoring + exact top-k (scores never written to HBM)."""
    fn = L.lib().b4cp_score_topk_workspace_bytes
    fn.restype = ctypes.c_long
    ws = WS.get("score_topk", fn(ctypes.c_long(M), ctypes.c_int(V), ctypes.c_int(k)))
    if out_ids is None:
        out_ids = empty((M, k), I32)
    L.call("b4cp_score_topk", L.ptr(xb), L.c_long(xb.stride(0)), L.c_long(M), L.c_int(h), L.ptr(wb),
           L.c_long(wb.stride(0)), L.ptr(bias), L.c_int(V), L.c_int(k), L.c_int(id_base),
           L.c_int(V_total), L.ptr(out_ids), L.ptr(out_scores), L.c_long(out_ids.stride(0)),
           L.ptr(ws), L.stream_ptr())
    return out_ids, out_scores


def topk_candidates(cand_scores, cand_ids, V, k, out_ids=None, out_scores=None):
    """Exact top-k over explicit (score, id) candidate rows [rows, n_cand] (ids < V, -1 = empty)."""
    rows, n_cand = cand_scores.shape
    if out_ids is None:
        out_ids = empty((rows, k), I32)
    L.call("b4cp_topk_candidates", L.ptr(cand_scores), L.ptr(cand_ids),
           L.c_long(cand_scores.stride(0)), L.c_long(rows), L.c_int(n_cand), L.c_int(V), L.c_int(k),
           L.ptr(out_ids), L.ptr(out_scores), L.c_long(out_ids.stride(0)), L.stream_ptr())
    return out_ids, out_scores
