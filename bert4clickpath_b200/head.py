"""Head units with the reference's names and constructor arguments
(clickstream_transformer/head.py:4-69).  A head is bound to a model's parameter store by
`build` (the Keras layers build on first call); the Dense stacks run on the tcgen05 GEMM."""
import numpy as np
import torch

from . import ops
from .engine import (MlpEngine, VocabOutputEngine, VocabParallelOutputEngine, glorot_uniform,
                     kernel_of)
from .ops import BF16, F32, I32, ld8


class _Head:
    def __init__(self, dense_layer_dims, **kwargs):
        self.dense_layer_dims = list(dense_layer_dims)
        self.built = False

    def build(self, store, in_dim, rng, prefix="head", precision="bf16"):
        assert not self.built, "a Head unit can be attached to one model only"
        self.store, self.in_dim, self.prefix = store, int(in_dim), prefix
        self.precision = precision
        self.mlp = MlpEngine(store, prefix, in_dim, self.dense_layer_dims, rng)
        self._build_output(store, self.mlp.out_dim, rng)
        self.built = True

    def hidden(self, xb, M):
        """bf16 [M, ld8(in)] -> bf16 [M, ld8(h)]: the ReLU Dense stack (head.py:16-19,41-43).
        (fp32 in -> fp32 out in the fp32-class mode.)"""
        return self.mlp.forward(xb, M)

    def _operand(self, x_f32):
        """A fp32 (rows, in_dim) input as the Dense stack's operand in this head's precision."""
        return x_f32 if self.precision == "fp32" else ops.cast_bf16(x_f32)


class ClozeOutput:
    """Lazy SoftMaxHead output: the (M, V) probabilities are never written to HBM on the
    training path.  Rows are the [MASK] positions in (b, s) order (clickstream_transformer.py:
    260-297); `materialize()` produces the reference-shaped (B, max_n_masked, V) tensor."""

    def __init__(self, head, ab, M, row_index, count, model_state):
        self.head, self.ab, self.M = head, ab, M
        self.row_index, self.count = row_index, count
        self.state = model_state

    def probabilities(self):
        """fp32 (M, V) softmax rows in compact (b, s) order (small-V only)."""
        return self.head.vocab.probabilities(self.ab, self.M)

    def materialize(self):
        """(B, max_n_masked, V) probabilities incl. the rows the reference computes from its
        zero-vector padding (head.py:38-47 applied to clickstream_transformer.py:295)."""
        st = self.state
        ids = st["ids_first"].view(st["B"], st["S"]).cpu().numpy()
        hit = ids == st["value_id"]
        counts = hit.sum(1)
        mmax = int(counts.max()) if len(counts) else 0
        B, S = st["B"], st["S"]
        V = self.head.output_vocab_size
        if mmax == 0:
            return torch.zeros((B, 0, V), dtype=F32, device="cuda")
        idx = np.full((B, mmax), -1, dtype=np.int32)
        for b in range(B):
            pos = np.nonzero(hit[b])[0]
            idx[b, :len(pos)] = b * S + pos
        row_index = torch.from_numpy(idx.reshape(-1)).cuda()
        M = B * mmax
        hsel = torch.empty((M, ld8(self.head.in_dim)), device="cuda",
                           dtype=F32 if self.head.precision == "fp32" else BF16)
        ops.gather_rows(st["x"], row_index, None, hsel)
        ab = self.head.hidden(hsel, M)
        return self.head.vocab.probabilities(ab, M).view(B, mmax, V)


class SoftMaxHead(_Head):
    """SoftMaxHead(dense_layer_dims, output_vocab_size) — head.py:29-47."""

    def __init__(self, dense_layer_dims, output_vocab_size, vocab_parallel=False,
                 vocab_parallel_group=None, **kwargs):
        """vocab_parallel=True (an extension; the reference only replicates) shards the output
        kernel by vocabulary ranges over `vocab_parallel_group` (default: the world)."""
        super().__init__(dense_layer_dims, **kwargs)
        self.output_vocab_size = int(output_vocab_size)
        self.vocab_parallel = bool(vocab_parallel)
        self.vocab_parallel_group = vocab_parallel_group

    def _build_output(self, store, h, rng):
        if self.vocab_parallel:
            self.vocab = VocabParallelOutputEngine(store, self.prefix, h, self.output_vocab_size,
                                                   rng, group=self.vocab_parallel_group)
        else:
            self.vocab = VocabOutputEngine(store, self.prefix, h, self.output_vocab_size, rng)

    def call(self, inputs, **kwargs):
        """inputs: fp32 (..., in_dim) device tensor -> (..., V) probabilities (materialised)."""
        lead = inputs.shape[:-1]
        x = inputs.reshape(-1, self.in_dim).contiguous()
        M = x.shape[0]
        ab = self.hidden(self._operand(x), M)
        return self.vocab.probabilities(ab, M).view(*lead, self.output_vocab_size)

    __call__ = call


class BinaryClassificationHead(_Head):
    """BinaryClassificationHead(dense_layer_dims) — head.py:4-26: ReLU MLP, Dense(1, sigmoid),
    squeeze(-1)."""

    def _build_output(self, store, h, rng):
        self.h = h
        store.add(f"{self.prefix}.out.w", glorot_uniform(rng, h, 1), shadow=True)
        store.add(f"{self.prefix}.out.b", np.zeros(1))

    def logits(self, xb, M):
        ab = self.hidden(xb, M)
        z = torch.empty((M, 1), dtype=F32, device="cuda")
        W, b = self.store[f"{self.prefix}.out.w"], self.store[f"{self.prefix}.out.b"]
        ops.gemm(ab, 0, kernel_of(W, ab), 1, M, 1, self.h, bias=b.w, out_f32=z)
        return z, ab

    def call(self, inputs, **kwargs):
        lead = inputs.shape[:-1]
        x = inputs.reshape(-1, self.in_dim).contiguous()
        z, _ = self.logits(self._operand(x), x.shape[0])
        return ops.sigmoid(z.view(-1)).view(*lead)

    __call__ = call


class MultiLabel_MultiClass_classification(_Head):
    """MultiLabel_MultiClass_classification(dense_layer_dims, output_vocab_size) —
    head.py:50-69: ReLU MLP, Dense(V, sigmoid), squeeze(axis=1)."""

    def __init__(self, dense_layer_dims, output_vocab_size, **kwargs):
        super().__init__(dense_layer_dims, **kwargs)
        self.output_vocab_size = int(output_vocab_size)

    def _build_output(self, store, h, rng):
        self.h = h
        store.add(f"{self.prefix}.out.w", glorot_uniform(rng, h, self.output_vocab_size), shadow=True)
        store.add(f"{self.prefix}.out.b", np.zeros(self.output_vocab_size))

    def call(self, inputs, **kwargs):
        lead = inputs.shape[:-1]
        x = inputs.reshape(-1, self.in_dim).contiguous()
        M, V = x.shape[0], self.output_vocab_size
        ab = self.hidden(self._operand(x), M)
        z = torch.empty((M, V), dtype=F32, device="cuda")
        W, b = self.store[f"{self.prefix}.out.w"], self.store[f"{self.prefix}.out.b"]
        ops.gemm(ab, 0, kernel_of(W, ab), 1, M, V, self.h, bias=b.w, out_f32=z)
        out = ops.sigmoid(z.view(-1)).view(*lead, V)
        if out.dim() >= 2 and out.shape[1] == 1:
            out = out.squeeze(1)  # tf.squeeze(logits, axis=1)
        return out

    __call__ = call
