"""Explicit forward / backward / optimizer schedule of the clickstream transformer on libb4cp.

This is the runtime under the reference-named classes (transformer.py, head.py,
clickstream_transformer.py): parameters live in flat fp32 buffers (one all-reduce, one Adam
sweep), every activation buffer is allocated once per shape and reused, and each step is a fixed
sequence of kernel launches on one stream (CUDA-graph capturable: no host sync, no allocation).

Numerics: fp32 master weights and residual stream; Dense layers run on tcgen05 tensor cores with
bf16 operands and fp32 accumulation (bf16 shadows of the weights are refreshed by the Adam
kernel); attention / LayerNorm / softmax math is fp32.

precision="fp32" (the parity mode) keeps every activation in fp32: Dense layers
still run on the tcgen05 GEMM, with both operands split into bf16 hi + lo parts along the
contraction axis (`ops.x3_operands`: hi*hi + hi*lo + lo*hi accumulated in fp32), attention runs in
the fp32 SIMT kernel and the Cloze output stage on materialised fp32 logits.  The same engine code
serves both modes: activation buffers are allocated in `act` (bf16 or fp32) and the `ops`
wrappers dispatch on the dtype they are handed.
"""
import math
import os

import numpy as np
import torch

from . import ops
from .ops import BF16, F32, I32, ld8

SITE_INPUT = 1  # dropout site ids (site = 16*layer + k for encoder layers)


def site(layer, k):
    return 16 * (layer + 1) + k


class Param:
    __slots__ = ("name", "shape", "w", "g", "m", "v", "wb", "offset", "numel", "replicated",
                 "grad_is_global")

    def __init__(self, name, shape):
        self.name, self.shape = name, tuple(shape)
        self.numel = int(np.prod(shape))
        self.w = self.g = self.m = self.v = self.wb = None
        self.offset = 0
        self.replicated = True  # False: a per-rank shard, excluded from the data-parallel all-reduce
        self.grad_is_global = False  # True: .g already holds the sum over ranks (row exchange)


class ParamStore:
    """Flat fp32 parameter / gradient / Adam-moment buffers with named views."""

    def __init__(self):
        self.params = {}
        self._init = {}
        self._shadow = set()
        self.flat_w = self.flat_g = self.flat_m = self.flat_v = None
        self.step_dev = None

    def add(self, name, init, shadow=False, replicated=True):
        init = np.ascontiguousarray(init, dtype=np.float32)
        assert name not in self.params, name
        self.params[name] = Param(name, init.shape)
        self.params[name].replicated = replicated
        self._init[name] = init
        if shadow:
            self._shadow.add(name)
        return self.params[name]

    def finalize(self):
        total = 0
        for p in self.params.values():
            p.offset = total
            total += (p.numel + 63) // 64 * 64  # keep every view 256-byte aligned
        self.flat_w = torch.zeros(total, dtype=F32, device="cuda")
        self.flat_g = torch.zeros(total, dtype=F32, device="cuda")
        self.flat_m = torch.zeros(total, dtype=F32, device="cuda")
        self.flat_v = torch.zeros(total, dtype=F32, device="cuda")
        self.step_dev = torch.ones(1, dtype=I32, device="cuda")
        for p in self.params.values():
            sl = slice(p.offset, p.offset + p.numel)
            p.w = self.flat_w[sl].view(p.shape)
            p.g = self.flat_g[sl].view(p.shape)
            p.m = self.flat_m[sl].view(p.shape)
            p.v = self.flat_v[sl].view(p.shape)
            p.w.copy_(torch.from_numpy(self._init[p.name]))
            if p.name in self._shadow:
                rows, cols = p.shape
                p.wb = torch.zeros((rows, ld8(cols)), dtype=BF16, device="cuda")
                ops.cast_bf16(p.w, cols, out=p.wb)
        self._init = None
        torch.cuda.synchronize()

    def __getitem__(self, name):
        return self.params[name]

    def set_weights(self, arrays):
        """Load {name: ndarray} (Keras layouts) and refresh the bf16 shadows."""
        for name, arr in arrays.items():
            p = self.params[name]
            p.w.copy_(torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float32)).view(p.shape))
            if p.wb is not None:
                ops.cast_bf16(p.w, p.shape[1], out=p.wb)
        torch.cuda.synchronize()

    def get_weights(self):
        return {n: p.w.detach().cpu().numpy().copy() for n, p in self.params.items()}

    def replicated_grad_runs(self, only=None, exclude=()):
        """Maximal contiguous slices of flat_g that hold replicated parameters (what data
        parallelism all-reduces; vocabulary-parallel shards are already complete per rank, and so
        are tables whose gradient rows were exchanged before the segment sums).  `only` /
        `exclude`: parameter names, to all-reduce a group as soon as its gradients are final."""
        runs = []
        for p in sorted(self.params.values(), key=lambda q: q.offset):
            if not p.replicated or p.grad_is_global:
                continue
            if (only is not None and p.name not in only) or p.name in exclude:
                continue
            end = p.offset + (p.numel + 63) // 64 * 64
            if runs and runs[-1][1] == p.offset:
                runs[-1][1] = end
            else:
                runs.append([p.offset, end])
        return [self.flat_g[a:b] for a, b in runs]

    def get_grads(self):
        WSTREAM.join()
        return {n: p.g.detach().cpu().numpy().copy() for n, p in self.params.items()}

    def adam(self, lr, beta1=0.9, beta2=0.999, eps=1e-9, grad_scale=1.0):
        """Keras-semantics Adam over every parameter (tables dense-equivalent), then t += 1: one
        sweep over the flat buffers (`b4cp_adam_flat`), which also refreshes the bf16 shadows of
        the Dense kernels; one launch per shadowed kernel only when there are more kernels than
        the segment table holds."""
        from . import _lib
        WSTREAM.join()
        shadowed = sorted((p for p in self.params.values() if p.wb is not None), key=lambda q: q.offset)
        if len(shadowed) <= _lib.ADAM_MAX_SEGS:
            if getattr(self, "_adam_segs", None) is None:
                segs = (_lib.AdamSegment * max(len(shadowed), 1))()
                for i, p in enumerate(shadowed):
                    segs[i].begin, segs[i].numel = p.offset, p.numel
                    segs[i].cols, segs[i].ld_shadow = p.shape[1], p.wb.stride(0)
                    segs[i].shadow_bf16 = p.wb.data_ptr()
                self._adam_segs = (segs, len(shadowed))
            segs, n = self._adam_segs
            ops.adam_flat(self.flat_w, self.flat_g, self.flat_m, self.flat_v, segs, n, lr=lr,
                          beta1=beta1, beta2=beta2, eps=eps, step_dev=self.step_dev,
                          grad_scale=grad_scale)
            ops.step_increment(self.step_dev)
            return
        plain = [p for p in self.params.values() if p.wb is None]
        # parameters without shadows are swept in maximal contiguous runs of the flat buffers
        runs = []
        for p in sorted(plain, key=lambda q: q.offset):
            end = p.offset + (p.numel + 63) // 64 * 64
            if runs and runs[-1][1] == p.offset:
                runs[-1][1] = end
            else:
                runs.append([p.offset, end])
        for a, b in runs:
            ops.adam_step(self.flat_w[a:b], self.flat_g[a:b], self.flat_m[a:b], self.flat_v[a:b],
                          lr=lr, beta1=beta1, beta2=beta2, eps=eps, step_dev=self.step_dev,
                          grad_scale=grad_scale)
        for p in self.params.values():
            if p.wb is not None:
                ops.adam_step(p.w, p.g, p.m, p.v, lr=lr, beta1=beta1, beta2=beta2, eps=eps,
                              step_dev=self.step_dev, grad_scale=grad_scale, shadow=p.wb,
                              cols=p.shape[1])
        ops.step_increment(self.step_dev)


class BufferPool:
    """Named device buffers allocated once per (name, shape, dtype)."""

    def __init__(self):
        self._b = {}

    def get(self, name, shape, dtype=F32, zero=False):
        key = (name, tuple(shape), dtype)
        t = self._b.get(key)
        if t is None:
            t = torch.zeros(shape, dtype=dtype, device="cuda")
            self._b[key] = t
        elif zero:
            ops.zero_(t)
        return t


def glorot_uniform(rng, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


# =============================================================================== weight stream
class WeightStream:
    """Weight gradients leave the critical path.  dW = x^T dy (split-K GEMM + reduction) and
    db = column sums of dy are needed by the optimizer alone, while the backward's critical path
    is the chain of INPUT gradients: every Dense layer used to queue four small launches between
    one dX product and the next (~0.3 ms per C1 step).  They now run on a second stream, forked
    after the kernel that produces dy and joined once before the gradient all-reduce / Adam, and
    fill whatever the dX chain leaves idle.  Same kernels, same summation order: results are
    bit-identical to the single-stream schedule (B4CP_WEIGHT_STREAM=0).

    The side work READS buffers the main stream will overwrite later (the per-layer dy buffers):
    `run` remembers an event per buffer, `before_write` makes the main stream wait for it."""

    def __init__(self):
        self.stream = None
        self.readers = {}
        self.forked = False     # side work queued since the last join
        self.enabled = os.environ.get("B4CP_WEIGHT_STREAM", "1") != "0"

    def run(self, fn, *reads):
        """Run fn on the weight stream after everything queued on the current stream so far.
        Returns the event that marks its completion (None when it ran inline)."""
        if not self.enabled:
            fn()
            return None
        cur = torch.cuda.current_stream()
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        self.stream.wait_stream(cur)
        self.forked = True
        with torch.cuda.stream(self.stream):
            fn()
            done = torch.cuda.Event()
            done.record(self.stream)
        for t in reads:
            self.readers[t.data_ptr()] = done
        return done

    @staticmethod
    def wait(done):
        """The current stream waits for side work started with run()."""
        if done is not None:
            torch.cuda.current_stream().wait_event(done)

    def before_write(self, *tensors):
        for t in tensors:
            ev = self.readers.pop(t.data_ptr(), None)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def join(self):
        # only a stream that was forked from the current one since the last join: waiting on an
        # idle side stream inside a graph capture is a dependency on uncaptured work
        if self.forked:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.forked = False
        self.readers.clear()


WSTREAM = WeightStream()


# =============================================================================== dense layer
def kernel_of(W, like):
    """The Dense kernel as a GEMM operand: the bf16 shadow next to bf16 activations, the fp32
    master weights next to fp32 activations (fp32-class mode)."""
    return W.w if like.dtype == F32 else W.wb


def dense_fwd(xb, K, W, bias, M, *, relu=False, out_f32=None, out_bf16=None):
    """y = act(x W + b).  xb bf16 [M, ld8(K)], W: Param with Keras (K, N) kernel + shadow."""
    N = W.shape[1]
    ops.gemm(xb, 0, kernel_of(W, xb), 1, M, N, K, bias=bias.w if bias is not None else None,
             relu=relu, out_f32=out_f32, out_bf16=out_bf16)


def dense_bwd_weights(xb, dyb, W, bias, M, *, db_from=None):
    """dW = x^T dy (split-K, deterministic), db = column sums of dy (unless already reduced).
    bf16 operands: on the weight stream (own workspaces); the caller joins it (WSTREAM.join) before
    the gradients are used and guards later writes of `dyb` with WSTREAM.before_write."""
    K, N = W.shape

    def work(ws_k="splitk", ws_c="colsum"):
        ops.gemm_splitk(xb, 1, dyb, 1, K, N, M, W.g, ws_name=ws_k)
        if bias is not None and db_from is None:
            ops.colsum_bf16(dyb, M, N, bias.g, ws_name=ws_c)

    if dyb.dtype == F32 or xb.dtype == F32:   # fp32-class mode: operand splits share workspaces
        work()
    else:
        WSTREAM.run(lambda: work("splitk_w", "colsum_w"), dyb)


def dense_bwd_input(dyb, W, M, *, gate=None, addend=None, out_f32=None, out_bf16=None):
    """dx = dy W^T, optionally gated by the ReLU of the layer that produced x / accumulated."""
    K, N = W.shape
    ops.gemm(dyb, 0, kernel_of(W, dyb), 0, M, K, N, gate=gate, addend=addend, out_f32=out_f32,
             out_bf16=out_bf16)


# =============================================================================== encoder
class EncoderEngine:
    """Embedding + encoder stack (clickstream_transformer/transformer.py:271-402)."""

    def __init__(self, store, embedding_sizes, embedding_dims, num_layers, num_heads, dff,
                 dropout_rate, rng, max_pos=10000, precision="bf16"):
        from .transformer import positional_encoding
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' (fast path) or 'fp32' (parity mode)")
        self.precision = precision
        self.act = F32 if precision == "fp32" else BF16
        self.store = store
        self.features = list(embedding_dims.keys())
        self.rows = [int(embedding_sizes[f]) for f in self.features]
        self.dims = [int(embedding_dims[f]) for f in self.features]
        self.d = sum(self.dims)
        assert self.d % num_heads == 0
        assert self.d % 8 == 0, "d_model must be a multiple of 8 (16-byte bf16 rows for TMA)"
        self.L, self.H, self.dff, self.rate = num_layers, num_heads, dff, float(dropout_rate)
        self.dh = self.d // num_heads
        d = self.d
        for f, (R, df) in enumerate(zip(self.rows, self.dims)):
            store.add(f"emb.{f}", rng.uniform(-0.05, 0.05, size=(R, df)))
        for l in range(num_layers):
            wq, wk, wv = (glorot_uniform(rng, d, d) for _ in range(3))
            store.add(f"enc.{l}.wqkv", np.concatenate([wq, wk, wv], axis=1), shadow=True)
            store.add(f"enc.{l}.bqkv", np.zeros(3 * d))
            store.add(f"enc.{l}.wo", glorot_uniform(rng, d, d), shadow=True)
            store.add(f"enc.{l}.bo", np.zeros(d))
            store.add(f"enc.{l}.w1", glorot_uniform(rng, d, dff), shadow=True)
            store.add(f"enc.{l}.b1", np.zeros(dff))
            store.add(f"enc.{l}.w2", glorot_uniform(rng, dff, d), shadow=True)
            store.add(f"enc.{l}.b2", np.zeros(d))
            for k in ("ln1", "ln2"):
                store.add(f"enc.{l}.{k}_g", np.ones(d))
                store.add(f"enc.{l}.{k}_b", np.zeros(d))
        self.pe_host = positional_encoding(max_pos, d)
        self.pe = None
        self.pool = BufferPool()
        self.saved = None
        self._side = None

    def _pe(self):
        if self.pe is None:
            self.pe = torch.from_numpy(self.pe_host).cuda()
        return self.pe

    def forward(self, ids_list, B, S, training, seed=0):
        """ids_list: per-feature int32 [B*S] device tensors.  Returns fp32 [B*S, d] (+ bf16)."""
        st, pool, d, T = self.store, self.pool, self.d, B * S
        rate = self.rate if training else 0.0
        act, lowp = self.act, self.act == BF16
        tables = [st[f"emb.{f}"].w for f in range(len(self.features))]
        x = pool.get("x0", (T, d))
        xb = pool.get("x0b", (T, d), BF16) if lowp else x   # fp32 mode: the fp32 buffer IS the operand
        ops.embed_fwd(ids_list, tables, self._pe(), B, S, dropout_rate=rate, seed=seed,
                      site=SITE_INPUT, out_f32=x, out_bf16=xb if lowp else None)
        acts = []
        dffp = ld8(self.dff)
        for l in range(self.L):
            g = lambda n: st[f"enc.{l}.{n}"]
            a = dict(x=x, xb=xb)
            a["qkvb"] = pool.get(f"qkvb{l}", (T, 3 * d), act)
            dense_fwd(xb, d, g("wqkv"), g("bqkv"), T, out_bf16=a["qkvb"])
            a["ob"] = pool.get(f"ob{l}", (T, d), act)
            a["lse"] = pool.get(f"lse{l}", (B, self.H, S))
            ops.attention_fwd(a["qkvb"], ids_list[0], B, S, self.H, self.dh, a["ob"], a["lse"])
            a["y1"] = pool.get(f"y1_{l}", (T, d))
            dense_fwd(a["ob"], d, g("wo"), g("bo"), T, out_f32=a["y1"])
            a["x1"] = pool.get(f"x1_{l}", (T, d))
            a["x1b"] = pool.get(f"x1b{l}", (T, d), BF16) if lowp else a["x1"]
            ops.residual_ln_fwd(x, a["y1"], g("ln1_g").w, g("ln1_b").w, a["x1"], a["x1b"],
                                dropout_rate=rate, seed=seed, site=site(l, 1))
            a["hb"] = pool.get(f"hb{l}", (T, dffp), act)
            dense_fwd(a["x1b"], d, g("w1"), g("b1"), T, relu=True, out_bf16=a["hb"])
            a["y2"] = pool.get(f"y2_{l}", (T, d))
            dense_fwd(a["hb"], self.dff, g("w2"), g("b2"), T, out_f32=a["y2"])
            x2 = pool.get(f"x2_{l}", (T, d))
            x2b = pool.get(f"x2b{l}", (T, d), BF16) if lowp else x2
            ops.residual_ln_fwd(a["x1"], a["y2"], g("ln2_g").w, g("ln2_b").w, x2, x2b,
                                dropout_rate=rate, seed=seed, site=site(l, 2))
            acts.append(a)
            x, xb = x2, x2b
        self.saved = dict(acts=acts, ids=ids_list, B=B, S=S, rate=rate, seed=seed)
        return x, xb

    # Data-parallel table gradients.  The reference's MirroredStrategy reduces the IndexedSlices
    # gradient of tf.gather (examples/BERT4Rec/source/main.py:52).  Here a table's gradient is
    # either all-reduced dense with the rest of the flat buffer, or - when the table is much
    # larger than a step's token rows (C3 / C4: 10^5..10^6-row tables) - every rank all-gathers
    # the (ids, dX rows) of all ranks and runs the same deterministic segment sum over them, so
    # the 1 GB dense gradient never crosses NVLink and every rank holds the global gradient.
    ROW_EXCHANGE_RATIO = 2.0   # exchange rows when world * T * d * 4 B < table bytes / ratio

    def plan_table_exchange(self, T, group=None):
        """Decide (per table) between the dense all-reduce and the row exchange; marks the
        parameters so that ParamStore.replicated_grad_runs() skips exchanged tables."""
        import torch.distributed as dist
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        plan = []
        for f, (R, df) in enumerate(zip(self.rows, self.dims)):
            rows_bytes = world * T * self.d * 4
            ex = world > 1 and rows_bytes * self.ROW_EXCHANGE_RATIO < R * df * 4
            self.store[f"emb.{f}"].grad_is_global = ex
            plan.append(ex)
        return plan, world

    def backward(self, dx, group=None):
        """dx: fp32 [T, d] gradient of the encoder output.  Fills every encoder / table .g
        (`group`: the data-parallel process group, for the table-gradient row exchange)."""
        sv, st, pool, d = self.saved, self.store, self.pool, self.d
        B, S, rate, seed = sv["B"], sv["S"], sv["rate"], sv["seed"]
        T = B * S
        dffp = ld8(self.dff)
        act = self.act
        plan, world = self.plan_table_exchange(T, group)
        # The sort of (id, token) that the table gradients need depends on the ids alone: it runs
        # NOW, on a side stream, under the encoder layers' backward (ten small latency-bound
        # launches that used to sit at the end of the step's critical path).
        presorted = not any(plan)
        if presorted:
            cur = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            for f, df in enumerate(self.dims):      # workspaces belong to the main stream
                ops._embed_ws(T, df, f"embed_bwd{f}")
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                for f, (R, df) in enumerate(zip(self.rows, self.dims)):
                    ops.embed_sort(sv["ids"][f], R, df, f"embed_bwd{f}")
        for l in reversed(range(self.L)):
            g = lambda n: st[f"enc.{l}.{n}"]
            a = sv["acts"][l]
            par = l & 1   # dy buffers alternate between layers: the weight stream may lag a layer
            # LN2: dx -> dx1 (residual), dy2 (FFN output grad, bf16) + dgamma/dbeta/db2
            dx1 = pool.get("dxa", (T, d))
            dy2b = pool.get(f"dyb2.{par}", (T, d), act)
            WSTREAM.before_write(dy2b)
            ops.residual_ln_bwd(dx, a["x1"], a["y2"], g("ln2_g").w, dx1, dy2b, g("ln2_g").g,
                                g("ln2_b").g, g("b2").g, dropout_rate=rate, seed=seed,
                                site=site(l, 2))
            dense_bwd_weights(a["hb"], dy2b, g("w2"), None, T)
            # (columns dff..ld8(dff) are zero from allocation and never written: no per-step fill)
            dhb = pool.get(f"dhb.{par}", (T, dffp), act)
            WSTREAM.before_write(dhb)
            dense_bwd_input(dy2b, g("w2"), T, gate=a["hb"], out_bf16=dhb)
            dense_bwd_weights(a["x1b"], dhb, g("w1"), g("b1"), T)
            dx1b = pool.get("dxb", (T, d))
            dense_bwd_input(dhb, g("w1"), T, addend=dx1, out_f32=dx1b)
            # LN1
            dxr = pool.get("dxa", (T, d))
            dy1b = pool.get(f"dyb1.{par}", (T, d), act)
            WSTREAM.before_write(dy1b)
            ops.residual_ln_bwd(dx1b, a["x"], a["y1"], g("ln1_g").w, dxr, dy1b, g("ln1_g").g,
                                g("ln1_b").g, g("bo").g, dropout_rate=rate, seed=seed,
                                site=site(l, 1))
            dense_bwd_weights(a["ob"], dy1b, g("wo"), None, T)
            dob = pool.get("dob", (T, d), act)
            dense_bwd_input(dy1b, g("wo"), T, out_bf16=dob)
            dqkvb = pool.get(f"dqkvb.{par}", (T, 3 * d), act)
            WSTREAM.before_write(dqkvb)
            ops.attention_bwd(a["qkvb"], dob, a["lse"], sv["ids"][0], B, S, self.H, self.dh, dqkvb,
                              out=a["ob"])
            dense_bwd_weights(a["xb"], dqkvb, g("wqkv"), g("bqkv"), T)
            dx = pool.get("dxb", (T, d))
            dense_bwd_input(dqkvb, g("wqkv"), T, addend=dxr, out_f32=dx)
        WSTREAM.join()   # every dW / db of the head MLP and the encoder is complete from here on
        dx_all = None
        if presorted:
            torch.cuda.current_stream().wait_stream(self._side)
            off = 0
            for f, (R, df) in enumerate(zip(self.rows, self.dims)):
                ops.embed_bwd_sorted(dx, d, off, df, T, R, st[f"emb.{f}"].g, f"embed_bwd{f}",
                                     dropout_rate=rate, seed=seed, site=SITE_INPUT)
                off += df
            return dx
        if any(plan):
            import torch.distributed as dist
            if rate > 0.0:   # the input-dropout mask is indexed by the LOCAL token: apply it here
                ops.dropout_apply(dx, rate, seed, SITE_INPUT)
            dx_all = pool.get("dx_all_ranks", (world * T, d))
            dist.all_gather_into_tensor(dx_all.view(-1), dx.view(-1), group=group)
        off = 0
        for f, (R, df) in enumerate(zip(self.rows, self.dims)):
            if plan[f]:
                ids_all = pool.get(f"ids_all_ranks{f}", (world * T,), I32)
                dist.all_gather_into_tensor(ids_all, sv["ids"][f], group=group)
                ops.embed_bwd(dx_all, d, off, df, ids_all, R, st[f"emb.{f}"].g)
            else:
                ops.embed_bwd(dx, d, off, df, sv["ids"][f], R, st[f"emb.{f}"].g,
                              dropout_rate=0.0 if dx_all is not None else rate, seed=seed,
                              site=SITE_INPUT)
            off += df
        return dx


# =============================================================================== head MLP
class MlpEngine:
    """ReLU Dense stack shared by every Head unit (clickstream_transformer/head.py:10,35,56)."""

    def __init__(self, store, prefix, in_dim, dims, rng):
        self.store, self.prefix, self.in_dim, self.dims = store, prefix, int(in_dim), list(dims)
        prev = self.in_dim
        for i, hd in enumerate(self.dims):
            store.add(f"{prefix}.{i}.w", glorot_uniform(rng, prev, hd), shadow=True)
            store.add(f"{prefix}.{i}.b", np.zeros(hd))
            prev = hd
        self.out_dim = prev
        self.pool = BufferPool()
        self.saved = None

    def forward(self, xb, M):
        """xb bf16 [M, ld8(in)] -> bf16 [M, ld8(out)] (post-ReLU of the last layer)."""
        acts = [xb]
        prev = self.in_dim
        for i, hd in enumerate(self.dims):
            W, b = self.store[f"{self.prefix}.{i}.w"], self.store[f"{self.prefix}.{i}.b"]
            out = self.pool.get(f"a{i}", (M, ld8(hd)), xb.dtype)
            dense_fwd(acts[-1], prev, W, b, M, relu=True, out_bf16=out)
            acts.append(out)
            prev = hd
        self.saved = dict(acts=acts, M=M)
        return acts[-1]

    def backward(self, dzb, out_f32):
        """dzb: bf16 [M, ld8(out)] gradient w.r.t. the last layer's PRE-activation (already
        gated).  Writes the gradient w.r.t. the MLP input (fp32 [M, in]) into out_f32."""
        acts, M = self.saved["acts"], self.saved["M"]
        for i in reversed(range(len(self.dims))):
            W, b = self.store[f"{self.prefix}.{i}.w"], self.store[f"{self.prefix}.{i}.b"]
            dense_bwd_weights(acts[i], dzb, W, b, M)
            if i > 0:
                prev = self.pool.get(f"da{i}", (M, ld8(self.dims[i - 1])), dzb.dtype)
                dense_bwd_input(dzb, W, M, gate=acts[i], out_bf16=prev)
                dzb = prev
            else:
                dense_bwd_input(dzb, W, M, out_f32=out_f32)


# =============================================================================== vocab output
class VocabOutputEngine:
    """Dense(V) output layer of SoftMaxHead (head.py:36,45) fused with the Cloze loss
    (examples/BERT4Rec/source/utils.py:116-134, losses.py:31-98) and the ranking metrics."""

    MATERIALIZE_LIMIT_BYTES = 12 << 30

    def __init__(self, store, prefix, in_dim, vocab, rng):
        self.store, self.prefix, self.h, self.V = store, prefix, int(in_dim), int(vocab)
        store.add(f"{prefix}.out.w", glorot_uniform(rng, self.h, self.V), shadow=True)
        store.add(f"{prefix}.out.b", np.zeros(self.V))
        self.pool = BufferPool()
        self.saved = None

    @property
    def W(self):
        return self.store[f"{self.prefix}.out.w"]

    @property
    def b(self):
        return self.store[f"{self.prefix}.out.b"]

    def _row_chunks(self, M):
        """Row ranges whose fp32 logits fit MATERIALIZE_LIMIT_BYTES (multiples of 128 rows)."""
        rows = max(128, self.MATERIALIZE_LIMIT_BYTES // (ld8(self.V) * 4) // 128 * 128)
        return [(a, min(a + rows, M)) for a in range(0, M, rows)]

    def logits(self, ab, M, chunk=None):
        """fp32 [rows, ld8(V)] logits of rows `chunk` = (a, b) of ab (default: all M rows, which
        must fit the limit).  Materialised; validation / small-V / h != 128 path."""
        Vp = ld8(self.V)
        a, b = chunk if chunk is not None else (0, M)
        if (b - a) * Vp * 4 > max(self.MATERIALIZE_LIMIT_BYTES, 128 * Vp * 4):
            raise MemoryError(f"refusing to materialise {b - a}x{self.V} logits; use the fused path")
        cap = min(M, self._row_chunks(M)[0][1])
        z = self.pool.get("logits", (cap, Vp))
        t0 = ops.TIMER.begin("vocab_gemm")
        dense_fwd(ab[a:b], self.h, self.W, self.b, b - a, out_f32=z)
        ops.TIMER.end("vocab_gemm", t0)
        return z[: b - a]

    def probabilities(self, ab, M):
        z = self.logits(ab, M)
        lse = self.pool.get("lse", (M,))
        tgt = self.pool.get("tgt", (M,))
        nolabel = self.pool.get("nolabel", (M,), I32, zero=True)
        ops.ce_rows_stats(z, self.V, nolabel, lse, tgt)
        probs = torch.empty((M, self.V), dtype=F32, device="cuda")
        ops.ce_rows_grad(z, self.V, None, lse, None, None, probs)
        return probs

    @property
    def fused(self):
        """The tcgen05 fused projection + CE kernels cover h = 128 and h = 256 (forward and
        backward; csrc/vocab_ce_ts.cu); other head widths take the materialised path."""
        return self.h in (128, 256) and not self.force_materialized

    force_materialized = False

    def loss_forward(self, ab, M, labels, loss_stats, need_grad=True, n_global=None):
        """loss_stats <- (sum over valid rows of lse - z_t, number of valid rows).  n_global
        (int32 device scalar): the valid-row count over all data-parallel ranks, reduced while
        the forward ran; loss_stats[1] is then that count and loss_stats[0] stays local."""
        lse = self.pool.get("lse", (M,))
        tgt = self.pool.get("tgt", (M,))
        z = None
        if self.fused and ab.dtype == BF16:   # fp32-class mode (fp32 rows) is materialised
            t0 = ops.TIMER.begin("vocab_ce")
            ops.vocab_ce_fwd(ab, M, self.h, self.W.wb, self.b.w, self.V, labels, lse, tgt,
                             want_dx=need_grad)
            ops.TIMER.end("vocab_ce", t0)
            chunks = None
        else:
            # materialised path (head widths the fused kernels do not cover): logits exist for one
            # bounded row range at a time; with several ranges the backward recomputes them
            chunks = self._row_chunks(M)
            for a, b in chunks:
                z = self.logits(ab, M, (a, b))
                ops.ce_rows_stats(z, self.V, labels[a:b], lse[a:b], tgt[a:b])
            if len(chunks) > 1:
                z = None
        ops.ce_loss_reduce(lse, tgt, labels, loss_stats, n_global)
        self.saved = dict(ab=ab, M=M, labels=labels, z=z, lse=lse, chunks=chunks)

    def loss_backward(self, loss_stats, gate, out_f32=None, out_bf16=None):
        """Gradients of mean CE (normalised by loss_stats[1], which may already be the global
        count): fills W.g, b.g and returns d(loss)/d(ab) gated by `gate` (the ReLU output that
        produced ab, or None) as fp32 and/or bf16."""
        sv = self.saved
        ab, M, V, h = sv["ab"], sv["M"], self.V, self.h
        if sv["chunks"] is None:  # fused: logits are recomputed tile by tile, never materialised
            t0 = ops.TIMER.begin("vocab_ce")
            ops.vocab_ce_dx(M, h, V, sv["labels"], loss_stats, self.W.wb, gate, out_f32, out_bf16)
            ops.vocab_ce_bwd(ab, M, h, self.W.wb, self.b.w, V, sv["labels"], sv["lse"], loss_stats,
                             self.W.g, self.b.g)
            ops.TIMER.end("vocab_ce", t0)
            return
        chunks = sv["chunks"]
        rows_cap = chunks[0][1] - chunks[0][0]
        fp32 = ab.dtype == F32
        dz_all = None if fp32 else self.pool.get("dz", (rows_cap, ld8(V)), BF16)
        db_parts = self.pool.get("db_parts", (len(chunks), V)) if len(chunks) > 1 else None
        for ci, (a, b) in enumerate(chunks):
            rows = b - a
            z = sv["z"] if sv["z"] is not None else self.logits(ab, M, (a, b))
            if fp32:   # dZ in fp32, in place over the logits
                dz = ops.ce_rows_grad_f32(z, V, sv["labels"][a:b], sv["lse"][a:b], loss_stats)
            else:
                dz = dz_all[:rows]
                ops.ce_rows_grad(z, V, sv["labels"][a:b], sv["lse"][a:b], loss_stats, dz, None)
            t0 = ops.TIMER.begin("vocab_gemm")
            if len(chunks) == 1:
                ops.gemm_splitk(ab, 1, dz, 1, h, V, rows, self.W.g, ws_name="splitk_vocab")
            else:  # dW accumulates over the row ranges in a fixed order (in-place addend)
                ops.gemm(ab[a:b], 1, dz, 1, h, V, rows, addend=self.W.g if ci else None,
                         out_f32=self.W.g)   # fp32 rows -> fp32-class product (ops.gemm dispatch)
            ops.TIMER.end("vocab_gemm", t0)
            ops.colsum_bf16(dz, rows, V, self.b.g if db_parts is None else db_parts[ci])
            # dx = dz W^T : K = V is long and rows x h is small -> split-K with a gated reduce
            t0 = ops.TIMER.begin("vocab_gemm")
            ops.gemm_splitk_ex(dz, 0, kernel_of(self.W, dz), 0, rows, h, V,
                               gate[a:b] if gate is not None else None,
                               out_f32[a:b] if out_f32 is not None else None,
                               out_bf16[a:b] if out_bf16 is not None else None)
            ops.TIMER.end("vocab_gemm", t0)
        if db_parts is not None:
            ops.reduce_splits(db_parts, self.b.g)

    FUSED_TOPK_MIN_V = 262144

    def topk(self, ab, M, k):
        """(M, k) int32 ids of the k highest scores per row, ties -> lower id.

        V >= 262,144 (C5: the 1M-item catalogue): the FUSED kernels of `b4cp_score_topk` - the
        scores never reach HBM.  A seed range is scored and ranked exactly, which gives every row
        a threshold; one tcgen05 sweep over the rest of the vocabulary appends the scores above
        it to per-row candidate lists; an exact merge finishes.  1.16 M queries/s at V = 1M,
        h = 256, k = 100, 4,096 queries per call, against 0.54 M for the materialised path.

        Shorter vocabularies (C1: 54,293 entries): logits materialised for a bounded row range at
        a time (tcgen05 GEMM, fp32) + the single-pass streaming top-k (`b4cp_topk_rows`): 3.7 M
        queries/s at h = 128; the fused alternative there is a per-row heap kernel, which is
        slower (0.67 M).  `prefer_fused_topk` = True / False overrides the choice."""
        ids = self.pool.get(f"topk{k}", (M, k), I32)
        fused_ok = (self.h in (64, 128, 256) and k <= 104 and not self.force_materialized
                    and ab.dtype == BF16)
        prefer = self.prefer_fused_topk
        if prefer is None:
            prefer = self.V >= self.FUSED_TOPK_MIN_V
        if fused_ok and prefer:
            t0 = ops.TIMER.begin("score_topk")
            ops.score_topk(ab, M, self.h, self.W.wb, self.b.w, self.V, k, out_ids=ids)
            ops.TIMER.end("score_topk", t0)
            return ids
        for a, b in self._row_chunks(M):
            z = self.logits(ab, M, (a, b))
            ops.topk_rows(z, self.V, k, out_ids=ids[a:b])
        return ids

    prefer_fused_topk = None   # None: by vocabulary size (see topk)


class VocabParallelOutputEngine(VocabOutputEngine):
    """The Dense(V) output layer sharded by contiguous vocabulary ranges over a process group
    (the reference only replicates head.py:36's kernel under MirroredStrategy).

    Rank r owns columns [v_begin, v_end) of the kernel and bias.  Per step: all-gather the head
    hidden rows and labels of every rank, run the fused projection + online-softmax kernels on the
    local shard for ALL rows, all-gather the per-shard log-sum-exps and merge them, all-reduce the
    target logits, and - in backward - reduce-scatter the per-shard dX partials back to the rank
    that owns each row.  dW / db of the shard are complete locally (they saw every row) and are
    excluded from the data-parallel gradient all-reduce.  Loss statistics are global by
    construction.  Every rank must present the same row capacity M."""

    stats_are_global = True

    def __init__(self, store, prefix, in_dim, vocab, rng, group=None):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("vocabulary-parallel output layer needs torch.distributed initialised")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.store, self.prefix, self.h, self.V_total = store, prefix, int(in_dim), int(vocab)
        if self.h not in (128, 256):
            raise ValueError("vocabulary-parallel output layer needs a 128- or 256-wide head (fused kernels)")
        # contiguous shards whose boundaries are multiples of 8 (bf16 leading-dimension rule)
        per = ld8((self.V_total + self.world - 1) // self.world)
        self.v_begin = min(self.rank * per, self.V_total)
        self.v_end = min(self.v_begin + per, self.V_total)
        self.V = self.v_end - self.v_begin
        if self.V <= 0:
            raise ValueError(f"rank {self.rank} would own an empty vocabulary shard")
        full = glorot_uniform(rng, self.h, self.V_total)  # same draw as the replicated layer
        store.add(f"{prefix}.out.w", full[:, self.v_begin:self.v_end], shadow=True, replicated=False)
        store.add(f"{prefix}.out.b", np.zeros(self.V), replicated=False)
        self.pool = BufferPool()
        self.saved = None

    def gather_full_weights(self):
        """{name: full (h, V_total) kernel / (V_total,) bias} assembled from every rank's shard
        (collective: all ranks call it).  Used by save_weights so that one file holds the model."""
        import torch.distributed as dist
        per = ld8((self.V_total + self.world - 1) // self.world)
        wt = torch.zeros((per, self.h), dtype=F32, device="cuda")
        wt[: self.V] = self.W.w.t()
        bt = torch.zeros((per,), dtype=F32, device="cuda")
        bt[: self.V] = self.b.w
        w_all = torch.empty((self.world * per, self.h), dtype=F32, device="cuda")
        b_all = torch.empty((self.world * per,), dtype=F32, device="cuda")
        dist.all_gather_into_tensor(w_all, wt, group=self.group)
        dist.all_gather_into_tensor(b_all, bt, group=self.group)
        return {f"{self.prefix}.out.w": w_all[: self.V_total].t().contiguous().cpu().numpy(),
                f"{self.prefix}.out.b": b_all[: self.V_total].cpu().numpy()}

    def shard_full_weights(self, arrays):
        """Inverse: replace full output-layer arrays by this rank's columns."""
        out = dict(arrays)
        for suffix, cut in (("w", lambda a: a[:, self.v_begin:self.v_end]),
                            ("b", lambda a: a[self.v_begin:self.v_end])):
            k = f"{self.prefix}.out.{suffix}"
            if k in out and np.asarray(out[k]).shape[-1] == self.V_total:
                out[k] = np.ascontiguousarray(cut(np.asarray(out[k])))
        return out

    def common_rows(self, M):
        """Row capacity shared by all ranks (max over the group); extra rows are padding."""
        import torch.distributed as dist
        t = torch.tensor([int(M)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def _gather_rows(self, ab, M, labels):
        import torch.distributed as dist
        W_ = self.world
        ab_all = self.pool.get("ab_all", (W_ * M, ab.shape[1]), BF16)
        dist.all_gather_into_tensor(ab_all, ab[:M], group=self.group)
        lab_all = None
        if labels is not None:
            lab_all = self.pool.get("labels_all", (W_ * M,), I32)
            dist.all_gather_into_tensor(lab_all, labels[:M], group=self.group)
        return ab_all, lab_all

    def loss_forward(self, ab, M, labels, loss_stats, need_grad=True, n_global=None):
        import torch.distributed as dist
        W_ = self.world
        ab_all, lab_all = self._gather_rows(ab, M, labels)
        Ma = W_ * M
        lab_sh = ops.shard_labels(lab_all, self.v_begin, self.V, out=self.pool.get("labels_shard", (Ma,), I32))
        lse_parts = self.pool.get("lse_parts", (W_, Ma))
        lse_loc = self.pool.get("lse_local", (Ma,))
        tgt = self.pool.get("tgt", (Ma,))
        t0 = ops.TIMER.begin("vocab_ce")
        ops.vocab_ce_fwd(ab_all, Ma, self.h, self.W.wb, self.b.w, self.V, lab_sh, lse_loc, tgt,
                         want_dx=need_grad)
        ops.TIMER.end("vocab_ce", t0)
        dist.all_gather_into_tensor(lse_parts.view(-1), lse_loc, group=self.group)
        dist.all_reduce(tgt, op=dist.ReduceOp.SUM, group=self.group)
        lse = ops.lse_merge(lse_parts, out=self.pool.get("lse", (Ma,)))
        ops.ce_loss_reduce(lse, tgt, lab_all, loss_stats)  # global sum and count, same on all ranks
        self.saved = dict(ab=ab_all, M=M, Ma=Ma, labels=lab_sh, lse=lse, z=None)

    def loss_backward(self, loss_stats, gate, out_f32=None, out_bf16=None):
        import torch.distributed as dist
        sv = self.saved
        M, Ma, h = sv["M"], sv["Ma"], self.h
        dx_all = self.pool.get("dx_all", (Ma, h))
        t0 = ops.TIMER.begin("vocab_ce")
        ops.vocab_ce_dx(Ma, h, self.V, sv["labels"], loss_stats, self.W.wb, None, out_f32=dx_all,
                        lse_global=sv["lse"])
        ops.vocab_ce_bwd(sv["ab"], Ma, h, self.W.wb, self.b.w, self.V, sv["labels"], sv["lse"],
                         loss_stats, self.W.g, self.b.g)
        ops.TIMER.end("vocab_ce", t0)
        dx_loc = self.pool.get("dx_local", (1, M, h))
        dist.reduce_scatter_tensor(dx_loc.view(-1), dx_all.view(-1), op=dist.ReduceOp.SUM,
                                   group=self.group)
        ops.reduce_splits_ex(dx_loc, M, h, gate, out_f32, out_bf16)  # ReLU gate + casts

    def topk(self, ab, M, k):
        """Global (M, k) top-k: per-shard fused scoring + top-k with global ids, all-gather of the
        world * k candidates per row, exact merge (ties -> lower id, as on one GPU)."""
        import torch.distributed as dist
        if k > 104:
            raise ValueError("vocabulary-parallel top-k supports k <= 104")
        W_ = self.world
        ab_all, _ = self._gather_rows(ab, M, None)
        Ma = W_ * M
        kk = min(k, self.V)
        ids_l = self.pool.get(f"vp_ids{k}", (Ma, k), I32)
        sc_l = self.pool.get(f"vp_sc{k}", (Ma, k))
        if kk < k:
            ids_l.fill_(-1)
            sc_l.fill_(float("-inf"))
        t0 = ops.TIMER.begin("score_topk")
        ops.score_topk(ab_all, Ma, self.h, self.W.wb, self.b.w, self.V, kk, out_ids=ids_l,
                       out_scores=sc_l, id_base=self.v_begin, V_total=self.V_total)
        ops.TIMER.end("score_topk", t0)
        # exchange: rank r needs the candidates of ITS rows from every shard -> all-to-all
        ids_x = self.pool.get(f"vp_idsx{k}", (W_, M, k), I32)
        sc_x = self.pool.get(f"vp_scx{k}", (W_, M, k))
        dist.all_to_all_single(ids_x.view(-1), ids_l.view(-1), group=self.group)
        dist.all_to_all_single(sc_x.view(-1), sc_l.view(-1), group=self.group)
        cand_ids = ids_x.permute(1, 0, 2).contiguous().view(M, W_ * k)
        cand_sc = sc_x.permute(1, 0, 2).contiguous().view(M, W_ * k)
        ids = self.pool.get(f"topk{k}", (M, k), I32)
        ops.topk_candidates(cand_sc, cand_ids, self.V_total, k, out_ids=ids)
        return ids

    def logits(self, ab, M):
        raise NotImplementedError("vocabulary-parallel output layer never materialises logits")

    probabilities = logits
