"""ClickstreamTransformer with the reference's construction surface
(clickstream_transformer/clickstream_transformer.py:8-375) on the libb4cp hot path.

Host side (this file): token chaining `[CLS] [SEP] seq_1 [SEP] seq_2 [SEP] ...`, vocabulary
lookup (10 reserved tokens + vocab + 1 OOV bucket), segment bookkeeping.  Device side: everything
from int ids to the head output / loss / gradients / metrics.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib as L
from . import ops
from .constants import (CLASSIFICATION_TOKEN, CLS, INPUT_MASKING_TOKEN, LABEL_PAD, RESERVED_TOKENS,
                        SEP, SEPARATOR_TOKEN)
from .engine import WSTREAM, BufferPool, ParamStore
from .head import ClozeOutput, SoftMaxHead
from .ops import BF16, F32, I32, ld8
from .transformer import Transformer


def load_vocabulary(vocab_file):
    """One token per line, stripped (clickstream_transformer/training_utils.py:5-12)."""
    with open(vocab_file, 'r') as f:
        return [line.strip() for line in f if line.strip() != '']


class StaticVocabularyTable:
    """tf.lookup.StaticVocabularyTable(KeyValueTensorInitializer(keys, range), num_oov_buckets=1)
    as used at clickstream_transformer.py:247-258: known keys -> their index (the first occurrence
    of a duplicated key), anything else -> len(keys); size() counts the OOV bucket.

    The table lives in libb4cp (`b4cp_vocab_table_*`, csrc/vocab_table.cu: open addressing over
    UCS4 code points, a few host threads): the reference runs this lookup inside its TensorFlow
    graph on string tensors, and per token in Python it costs ten times the training step it
    feeds.  String arrays are passed in NumPy's '<U' layout; byte strings are decoded and object
    arrays converted first (pass '<U' arrays to skip that)."""

    THREADS = max(1, min(8, os.cpu_count() or 1))

    def __init__(self, keys):
        self.keys = list(keys)
        self.oov = len(self.keys)
        arr = (np.asarray(self.keys, dtype=np.str_) if self.keys else np.zeros((0,), dtype="<U1"))
        arr = np.ascontiguousarray(arr)
        width = max(1, arr.dtype.itemsize // 4)
        if arr.dtype.itemsize == 0:
            arr = arr.astype("<U1")
        fn = L.lib().b4cp_vocab_table_create
        fn.restype = ctypes.c_void_p
        self._handle = fn(arr.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(self.keys)),
                          ctypes.c_int(width))
        if not self._handle:
            raise L.B4cpError(f"b4cp_vocab_table_create failed: {L.lib().b4cp_last_error().decode()}")

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                L.lib().b4cp_vocab_table_destroy(ctypes.c_void_p(h))
            except Exception:      # interpreter shutdown
                pass

    def size(self):
        return len(self.keys) + 1

    def lookup(self, tokens, out=None):
        """ids (int32, same shape) of a string array; integer arrays pass through as ids.
        `out`: optional C-contiguous int32 array of the same shape to fill (e.g. pinned memory)."""
        arr = np.asarray(tokens)
        if arr.dtype.kind in "iu":
            ids = arr.astype(np.int32)  # already ids
            if out is not None:
                out[...] = ids
                return out
            return ids
        if arr.dtype.kind == "S":
            arr = np.char.decode(arr, "utf-8")
        elif arr.dtype.kind != "U":
            arr = arr.astype(np.str_)
        if arr.dtype.itemsize == 0:
            arr = arr.astype("<U1")
        arr = np.ascontiguousarray(arr)
        if not arr.dtype.isnative:
            arr = arr.astype(arr.dtype.newbyteorder("="))
        if out is None:
            out = np.empty(arr.shape, dtype=np.int32)
        assert out.dtype == np.int32 and out.shape == arr.shape and out.flags.c_contiguous
        L.call("b4cp_vocab_table_lookup", ctypes.c_void_p(self._handle),
               arr.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(arr.size),
               ctypes.c_int(arr.dtype.itemsize // 4), out.ctypes.data_as(ctypes.c_void_p),
               ctypes.c_int(self.THREADS))
        return out


class TransformerInputPrep:
    """Chains raw sequences into `[CLS] [SEP] seq_1 [SEP] seq_2 [SEP] ...` and reports segment
    starts / ends from the SEP positions of sample 0 (clickstream_transformer.py:8-103).
    Works on string arrays (reference behaviour) or on integer id arrays."""

    def __init__(self, seq_chain_mapping):
        self.seq_chain_mapping = seq_chain_mapping

    @staticmethod
    def _chain_sequences(sequences):
        first = np.asarray(sequences[0])
        is_str = first.dtype.kind in "USO"
        cls_v = CLASSIFICATION_TOKEN if is_str else CLS
        sep_v = SEPARATOR_TOKEN if is_str else SEP
        shape = (first.shape[0], 1) + tuple(first.shape[2:])
        dtype = object if is_str else first.dtype
        if is_str and all(np.asarray(q).dtype.kind == "U" for q in sequences):
            # fixed-width unicode stays fixed-width (no per-token Python objects on the way to
            # the native lookup): wide enough for the longest input and for the two tokens
            dtype = np.result_type(*[np.asarray(q).dtype for q in sequences], np.dtype("<U5"))
        cls_token = np.full(shape, cls_v, dtype=dtype)
        sep_token = np.full(shape, sep_v, dtype=dtype)
        seqs = [cls_token] + [np.asarray(s).astype(dtype) for s in sequences]
        concat_list = [seqs[i // 2] if i % 2 == 0 else sep_token for i in range(2 * len(seqs))]
        return np.concatenate(concat_list, axis=1)

    def __call__(self, features, keep_features=False):
        features = dict(features)
        for new_feature, seq_pair in self.seq_chain_mapping.items():
            features[new_feature] = self._chain_sequences([features[name] for name in seq_pair])
        some = features[list(self.seq_chain_mapping.keys())[0]]
        sample = some[0, :] if some.shape[0] > 0 else np.asarray([])
        sep_v = SEPARATOR_TOKEN if np.asarray(some).dtype.kind in "USO" else SEP
        segment_ends = np.nonzero(sample == sep_v)[0]
        segment_starts = np.concatenate([[0], segment_ends[:-1] + 1]).astype(np.int64)
        if not keep_features:
            drop = set()
            for seq_pair in self.seq_chain_mapping.values():
                drop |= set(seq_pair)
            drop -= set(self.seq_chain_mapping.keys())
            features = {k: v for k, v in features.items() if k not in drop}
        return features, segment_starts, segment_ends


class Adam:
    """tf.keras.optimizers.Adam(learning_rate, beta_1, beta_2, epsilon) hyper-parameters
    (examples/BERT4Rec/source/main.py:87); the update itself is b4cp_adam_step."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-9):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = (learning_rate, beta_1, beta_2,
                                                                      epsilon)


class ClickstreamTransformer:
    """ClickstreamTransformer(sequential_input_config, feature_vocabs, embedding_dims, head_unit,
    segment_to_head=None, value_to_head=None, num_encoder_layers=1, num_attention_heads=1,
    dropout_rate=0.1) — clickstream_transformer.py:160-227.

    feature_vocabs values may be a vocabulary file path (reference), a list of tokens, or an int
    vocabulary size (inputs are then integer ids 10..V+9 already).  `encoder_ff_dim` defaults to
    the reference's hard-coded 100 (clickstream_transformer.py:225).

    precision="bf16" (default) feeds the tensor cores bf16 operands (fp32 accumulation, fp32 master
    weights); precision="fp32" is the parity mode: fp32 activations end to end, Dense layers as
    bf16 x 3 split products on the same tcgen05 GEMM, fp32 attention, materialised fp32 logits -
    loss, logits and gradients then match the fp32 reference to ~1e-5.
    """

    def __init__(self, sequential_input_config, feature_vocabs, embedding_dims, head_unit,
                 segment_to_head=None, value_to_head=None, num_encoder_layers=1,
                 num_attention_heads=1, dropout_rate=0.1, *, encoder_ff_dim=100, seed=0,
                 precision="bf16", **kwargs):
        self.sequential_input_config = sequential_input_config
        self.feature_vocabs = feature_vocabs
        self.embedding_dims = embedding_dims
        self.head = head_unit
        self.num_encoder_layers = num_encoder_layers
        self.num_attention_heads = num_attention_heads
        self.dropout_rate = dropout_rate
        assert (segment_to_head is not None or value_to_head is not None) and \
               (segment_to_head is None or value_to_head is None), \
            "Exactly one of segment_to_head and value_to_head must be provided."
        self.segment_to_head = segment_to_head
        self.value_to_head = value_to_head
        self.transformer_input_prep = TransformerInputPrep(self.sequential_input_config)
        self.vocab_lookup_tables = self._create_lookup_tables(self.feature_vocabs, RESERVED_TOKENS)
        self.embedding_sizes = {f: self.vocab_lookup_tables[f].size() for f in self.feature_vocabs}
        seq_keys = list(self.sequential_input_config.keys())
        self.store = ParamStore()
        self.transformer = Transformer(
            embedding_sizes={k: self.embedding_sizes[k] for k in seq_keys},
            embedding_dims={k: self.embedding_dims[k] for k in seq_keys},
            num_layers=num_encoder_layers, num_attention_heads=num_attention_heads,
            encoder_ff_dim=encoder_ff_dim, dropout_rate=dropout_rate, store=self.store, seed=seed,
            precision=precision)
        self.precision = precision
        self.act = self.transformer.engine.act
        self.d_model = self.transformer.d_model
        self.head.build(self.store, self.d_model, np.random.default_rng(seed + 1), precision=precision)
        self.store.finalize()
        if self.value_to_head is not None:
            v = self.value_to_head
            self._value_id = int(v) if isinstance(v, (int, np.integer)) else \
                int(self.vocab_lookup_tables[seq_keys[0]].lookup(np.asarray([v]))[0])
        self.pool = BufferPool()
        self.optimizer = self.loss = None
        self.metrics = []
        self.process_group = None
        self._step_seed = 0
        self._opt_iterations = 0
        self.stop_training = False

    # ------------------------------------------------------------------ construction helpers
    @staticmethod
    def _create_lookup_tables(vocabularies, tokens_to_prepend=None):
        tables = {}
        for name, vocab in vocabularies.items():
            if isinstance(vocab, (int, np.integer)):
                keys = [f"item_{j}" for j in range(int(vocab))]
            elif isinstance(vocab, str):
                keys = load_vocabulary(vocab)
            else:
                keys = list(vocab)
            if tokens_to_prepend is not None:
                keys = list(tokens_to_prepend) + keys
            tables[name] = StaticVocabularyTable(keys)
        return tables

    def get_config(self):
        return {
            'sequential_input_config': self.sequential_input_config,
            'feature_vocabs': self.feature_vocabs,
            'embedding_dims': self.embedding_dims,
            'head_unit': self.head,
            'segment_to_head': self.segment_to_head,
            'value_to_head': self.value_to_head,
            'num_encoder_layers': self.num_encoder_layers,
            'num_attention_heads': self.num_attention_heads,
            'dropout_rate': self.dropout_rate,
        }

    def get_serving_signature(self):
        """{raw sequence feature: ([None, None], 'string')} (clickstream_transformer.py:354-375)."""
        feats = []
        for chain in self.sequential_input_config.values():
            feats.extend(chain)
        return {f: ([None, None], 'string') for f in feats}

    # ------------------------------------------------------------------ input preparation
    def prepare_host(self, inputs, out=None):
        """dict of raw (B, L_i) string / id arrays -> ({sequential feature: chained int32 (B, S)
        ids}, segment starts, segment ends), all on the host.

        The reference chains the STRING sequences and looks the chained tensor up
        (clickstream_transformer.py:38-63, :307-308); the id of a token does not depend on where
        it stands, so here every raw sequence is looked up first (native table, '<U' arrays are
        never turned into Python objects) and the int32 pieces are chained with the ids of
        [CLS] / [SEP] - the same tensor.  `out`: optional {feature: int32 (B, S) array} to fill
        (pinned staging memory)."""
        chained = {}
        for name, chain in self.sequential_input_config.items():
            table = self.vocab_lookup_tables[name]
            pieces = []
            for raw_name in chain:
                v = inputs[raw_name]
                pieces.append(table.lookup(v.cpu().numpy() if torch.is_tensor(v) else np.asarray(v)))
            ids = TransformerInputPrep._chain_sequences(pieces)
            if out is not None and name in out:
                out[name][...] = ids
                ids = out[name]
            chained[name] = ids
        first = chained[list(self.sequential_input_config.keys())[0]]
        sample = first[0, :] if first.shape[0] > 0 else np.asarray([], dtype=np.int32)
        ends = np.nonzero(sample == SEP)[0]
        starts = np.concatenate([[0], ends[:-1] + 1]).astype(np.int64)
        return chained, starts, ends

    def prepare_inputs(self, inputs):
        """dict of raw (B, L_i) string / id arrays -> chained device ids per sequential feature."""
        chained, starts, ends = self.prepare_host(inputs)
        ids_list, shape = [], None
        for name in self.sequential_input_config.keys():
            ids = chained[name]
            if shape is None:
                self._host_ids_first = ids
            shape = ids.shape
            ids_list.append(torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int32)).cuda().view(-1))
        B, S = shape
        return ids_list, B, S, starts, ends

    # ------------------------------------------------------------------ forward
    def _encode(self, ids_list, B, S, training, seed):
        x, _ = self.transformer.engine.forward(ids_list, B, S, training, seed)
        return x

    def forward_ids(self, ids_list, B, S, training=False, seed=0, n_masked=None,
                    segment_bounds=None, rows_are_common=False):
        """Hot-path entry on already-chained device ids (int32 [B*S] per feature)."""
        if self.segment_to_head is not None:
            x = self._encode(ids_list, B, S, training, seed)
            starts, ends = segment_bounds
            s0, s1 = int(starts[self.segment_to_head]), int(ends[self.segment_to_head])
            head_input = x.view(B, S, self.d_model)[:, s0:s1, :]
            return self.head(head_input)
        cap = int(n_masked) if n_masked is not None else B * S
        vocab = getattr(self.head, "vocab", None)
        if hasattr(vocab, "common_rows") and not rows_are_common:
            cap = vocab.common_rows(cap)  # vocabulary-parallel: every rank presents the same rows
        # which rows go to the head depends on the ids alone: the five small launches of the
        # selection run on the second stream under the encoder
        row_buf = self.pool.get("row_index", (max(cap, 1),), I32)
        count = self.pool.get("row_count", (1,), I32)
        row_index = row_buf[:cap]
        selected = WSTREAM.run(lambda: ops.select_masked(ids_list[0], self._value_id, cap, row_buf, count))
        x = self._encode(ids_list, B, S, training, seed)
        state = dict(x=x, ids_first=ids_list[0], B=B, S=S)
        WSTREAM.wait(selected)
        hsel = self.pool.get("hsel", (cap, ld8(self.d_model)), self.act)
        ops.gather_rows(x, row_index, None, hsel)
        state["value_id"] = self._value_id
        if isinstance(self.head, SoftMaxHead):
            ab = self.head.hidden(hsel, cap)
            return ClozeOutput(self.head, ab, cap, row_index, count, state)
        # other heads consume the reference's padded (B, max_n_masked, d) tensor
        idx = row_index.cpu().numpy()
        n = int(count.item())
        b_of = idx[:n] // S
        counts = np.bincount(b_of, minlength=B)
        mmax = int(counts.max()) if n else 0
        padded = torch.zeros((B, mmax, self.d_model), dtype=F32, device="cuda")
        if n:
            j = np.arange(n) - np.repeat(np.cumsum(counts) - counts, counts)
            padded[torch.from_numpy(b_of).cuda(), torch.from_numpy(j).cuda()] = x[row_index[:n].long()]
        return self.head(padded)

    def call(self, inputs, training=None, mask=None):
        """inputs: dict of raw features -> head output (a lazy ClozeOutput for SoftMaxHead in
        value_to_head mode), or {'instance_id', 'logits'} when 'instance_id' is given."""
        feats = {k: v for k, v in inputs.items() if k != 'instance_id'}
        ids_list, B, S, starts, ends = self.prepare_inputs(feats)
        if self.segment_to_head is None and self.value_to_head is None:
            raise ValueError("One of value_to_head and segment_to_head must be provided.")
        n_masked = None
        if self.value_to_head is not None:
            n_masked = int((self._host_ids_first == self._value_id).sum())
        out = self.forward_ids(ids_list, B, S, bool(training), self._next_seed() if training else 0,
                               n_masked=n_masked, segment_bounds=(starts, ends))
        if 'instance_id' in inputs:
            return {'instance_id': inputs['instance_id'], 'logits': out}
        return out

    __call__ = call

    def topk_ids(self, ids_list, B, S, k, n_masked=None):
        """Next-item inference on chained device ids: the k best label-vocabulary ids for every
        [MASK] position, in (b, s) order -> int32 (n_masked, k).  Scores never reach HBM."""
        out = self.forward_ids(ids_list, B, S, training=False, n_masked=n_masked)
        return out.head.vocab.topk(out.ab, out.M, k), out

    def _next_seed(self):
        self._step_seed += 1
        return self._step_seed

    def _next_lr(self, opt):
        """Learning rate of this update: a float, or a schedule evaluated at the optimizer's
        iteration count (Keras passes `iterations`, which is 0 on the first update)."""
        from .training_utils import current_learning_rate
        lr = current_learning_rate(opt, self._opt_iterations)
        self._opt_iterations += 1
        return lr

    # ------------------------------------------------------------------ training (Keras-like)
    def compile(self, optimizer=None, loss=None, metrics=None):
        self.optimizer = optimizer if optimizer is not None else Adam()
        self._opt_iterations = 0
        self.loss = loss
        self.metrics = list(metrics or [])

    def set_process_group(self, group):
        """Data-parallel training: gradients / loss statistics are all-reduced over `group`."""
        self.process_group = group

    def _allreduce(self, t, async_op=False):
        """Sum over the data-parallel group.  async_op=True returns the pending work (or None on
        one rank): the collective runs on NCCL's stream while this stream carries on with the
        rest of the backward, and `_finish_reduce` makes this stream wait for it."""
        import os
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.process_group) > 1:
            if os.environ.get("B4CP_DP_OVERLAP") == "0":   # developer switch: A/B the overlap
                async_op = False
            return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.process_group,
                                   async_op=async_op)
        return None

    def _dp_world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def _reduce_gradients(self, early=()):
        """All-reduce every replicated gradient that is not reduced yet (`early`: names whose
        all-reduce was launched asynchronously when they became final)."""
        for run in self.store.replicated_grad_runs(exclude=set(early)):
            self._allreduce(run)

    @staticmethod
    def _finish_reduce(works):
        for w in works:
            if w is not None:
                w.wait()

    def cloze_forward_backward(self, ids_list, labels_f32, B, S, n_masked=None, training=True,
                               seed=0, rows_are_common=False):
        """One Cloze forward + backward on device-resident inputs.  labels_f32: (B, Mmax) float32
        padded with -1 (the reference contract).  Gradients land in store.flat_g; returns the
        device tensor loss_stats = (sum of per-position losses, valid positions) — already
        all-reduced when a process group is set, so loss = stats[0] / stats[1] is the GLOBAL
        masked mean."""
        assert isinstance(self.head, SoftMaxHead) and self.value_to_head is not None
        vocab, mlp = self.head.vocab, self.head.mlp
        dp = self._dp_world() > 1 and not getattr(vocab, "stats_are_global", False)
        labels = n_global = count_work = None
        if dp and n_masked is not None:
            # data parallel: the backward needs the GLOBAL number of valid rows and nothing else
            # from the other ranks, and that count depends on the labels alone - reduce it now, on
            # NCCL's stream, while the forward runs (it used to be a 2-float all-reduce between
            # the loss and the backward: ~25 us of exposed latency per step)
            labels, n_global = ops.compact_labels(labels_f32, int(n_masked))
            count_work = self._allreduce(n_global, async_op=True)
        compacted = None
        if (not dp and n_masked is not None
                and (rows_are_common or not hasattr(vocab, "common_rows"))):
            # single process: the label compaction (a function of the labels alone) runs on the
            # second stream under the encoder forward
            cap0 = int(n_masked)
            lab_buf = self.pool.get("labels_compact", (max(cap0, 1),), I32)
            n_global = self.pool.get("labels_count", (1,), I32)
            labels = lab_buf[:cap0]
            compacted = WSTREAM.run(lambda: ops.compact_labels(labels_f32, cap0, out=lab_buf, count=n_global))
        out = self.forward_ids(ids_list, B, S, training, seed, n_masked=n_masked,
                               rows_are_common=rows_are_common)
        cap = out.M
        WSTREAM.wait(compacted)
        if labels is None or labels.numel() != cap:
            labels, n_global = ops.compact_labels(labels_f32, cap)
            count_work = self._allreduce(n_global, async_op=True) if dp else None
        stats = self.pool.get("loss_stats", (2,))
        self._finish_reduce([count_work])
        vocab.loss_forward(out.ab, cap, labels, stats, n_global=n_global if dp else None)
        # the loss SUM is only reported: its all-reduce rides behind the backward
        sum_work = self._allreduce(stats[0:1], async_op=True) if dp else None
        d = self.d_model
        dsel = self.pool.get("dsel", (cap, d))
        # the output kernel's gradient (the largest tensor: 28 of 45 MB at C1) is final as soon as
        # the vocabulary backward has run: its all-reduce overlaps the head-MLP / encoder backward
        early = (f"{self.head.prefix}.out.w", f"{self.head.prefix}.out.b")
        if mlp.dims:
            dzb = self.pool.get("dz_head", (cap, ld8(mlp.out_dim)), self.act)
            vocab.loss_backward(stats, out.ab, out_bf16=dzb)
            works = [self._allreduce(r, async_op=True)
                     for r in self.store.replicated_grad_runs(only=set(early))]
            mlp.backward(dzb, dsel)
        else:
            vocab.loss_backward(stats, None, out_f32=dsel)
            works = [self._allreduce(r, async_op=True)
                     for r in self.store.replicated_grad_runs(only=set(early))]
        dx = self.pool.get("dx_top", (B * S, d), zero=True)
        ops.scatter_rows(dsel, out.row_index, dx)
        self.transformer.engine.backward(dx, self.process_group)
        self._reduce_gradients(early)
        self._finish_reduce(works + [sum_work])
        self._last_output = out
        self._last_labels = labels
        return stats

    def binary_forward_backward(self, ids_list, y_f32, B, S, segment_bounds, pos_weight=None,
                                label_pad=LABEL_PAD, training=True, seed=0):
        """Segment mode with a BinaryClassificationHead (purchase intention / return prediction,
        SURVEY.md C3): encoder -> rows of segment `segment_to_head` (clickstream_transformer.py:
        317-322) -> ReLU MLP -> Dense(1, sigmoid) (head.py:4-26) -> MaskedLoss(binary_crossentropy,
        pos_weight) (losses.py:31-98), and the whole backward.  y_f32: (B, segment length) float32
        padded with label_pad.  Gradients land in store.flat_g; returns loss_stats = (sum of item
        losses, valid items), all-reduced when a process group is set."""
        from .head import BinaryClassificationHead
        assert isinstance(self.head, BinaryClassificationHead) and self.segment_to_head is not None
        x = self._encode(ids_list, B, S, training, seed)
        starts, ends = segment_bounds
        s0, s1 = int(starts[self.segment_to_head]), int(ends[self.segment_to_head])
        Ls = s1 - s0
        M = B * Ls
        key = (B, S, s0, s1)
        if getattr(self, "_seg_rows_key", None) != key:
            idx = (np.arange(B, dtype=np.int64)[:, None] * S + np.arange(s0, s1)[None, :]).reshape(-1)
            self._seg_rows = torch.from_numpy(idx.astype(np.int32)).cuda()
            self._seg_rows_key = key
        row_index = self._seg_rows
        d = self.d_model
        hsel = self.pool.get("hsel_bin", (M, ld8(d)), self.act)
        ops.gather_rows(x, row_index, None, hsel)
        head = self.head
        z, ab = head.logits(hsel, M)
        probs = ops.sigmoid(z.view(-1))
        y = y_f32.reshape(-1).contiguous()
        assert y.numel() == M, "labels must be (B, segment length)"
        stats = ops.masked_bce(y, probs, label_pad, pos_weight)
        self._allreduce(stats)
        W, b = self.store[f"{head.prefix}.out.w"], self.store[f"{head.prefix}.out.b"]
        dz = self.pool.get("dz_bin", (M,))
        dsel = self.pool.get("dsel_bin", (M, d))
        mlp = head.mlp
        if self.act == F32:
            # fp32-class mode: the same backward from generic pieces (item-wise dz, then the three
            # Dense(1) products as fp32-class GEMMs)
            h = head.h
            ops.sigmoid_bce_dz(y, probs, M, 1, label_pad, pos_weight, stats, dz_f32=dz)
            dz2 = dz.view(M, 1)
            ops.gemm_splitk(ab, 1, dz2, 1, h, 1, M, W.g.view(h, 1))
            ops.colsum_bf16(dz2, M, 1, b.g)
            if mlp.dims:
                dab = self.pool.get("dab_bin", (M, ld8(mlp.out_dim)), F32)
                ops.gemm(dz2, 0, W.w, 0, M, h, 1, gate=ab, out_f32=dab)
                mlp.backward(dab, dsel)
            else:
                ops.gemm(dz2, 0, W.w, 0, M, h, 1, out_f32=dsel)
        elif mlp.dims:
            dab = self.pool.get("dab_bin", (M, ld8(mlp.out_dim)), BF16)
            ops.binary_head_bwd(y, probs, label_pad, pos_weight, stats, ab, head.h, W.w, True, dz,
                                dab_bf16=dab, dw=W.g, db=b.g)
            mlp.backward(dab, dsel)
        else:
            ops.binary_head_bwd(y, probs, label_pad, pos_weight, stats, ab, head.h, W.w, False, dz,
                                dab_f32=dsel, dw=W.g, db=b.g)
        dx = self.pool.get("dx_top", (B * S, d), zero=True)
        ops.scatter_rows(dsel, row_index, dx)
        self.transformer.engine.backward(dx, self.process_group)
        self._reduce_gradients()
        self._last_probs = probs.view(B, Ls)
        return stats

    def multilabel_forward_backward(self, ids_list, y_f32, B, S, segment_bounds, pos_weight=None,
                                    label_pad=LABEL_PAD, training=True, seed=0):
        """Segment mode with MultiLabel_MultiClass_classification (head.py:50-69): encoder -> the
        single row of segment `segment_to_head` (tf.squeeze(axis=1) needs length 1, e.g. [CLS]) ->
        ReLU MLP -> Dense(V, sigmoid) -> MaskedLoss(binary_crossentropy, pos_weight) over every
        (row, class) cell, and the whole backward.  y_f32: (B, V) float32, label_pad = ignored
        cell.  Same contract as binary_forward_backward."""
        from .head import MultiLabel_MultiClass_classification
        head = self.head
        assert isinstance(head, MultiLabel_MultiClass_classification) and self.segment_to_head is not None
        x = self._encode(ids_list, B, S, training, seed)
        starts, ends = segment_bounds
        s0, s1 = int(starts[self.segment_to_head]), int(ends[self.segment_to_head])
        if s1 - s0 != 1:
            raise ValueError(f"MultiLabel_MultiClass_classification squeezes axis 1: segment "
                             f"{self.segment_to_head} has length {s1 - s0}, not 1")
        M, V, d, h = B, head.output_vocab_size, self.d_model, head.h
        key = (B, S, s0, s1)
        if getattr(self, "_seg_rows_key", None) != key:
            idx = (np.arange(B, dtype=np.int64) * S + s0)
            self._seg_rows = torch.from_numpy(idx.astype(np.int32)).cuda()
            self._seg_rows_key = key
        row_index = self._seg_rows
        from .engine import kernel_of
        fp32 = self.act == F32
        hsel = self.pool.get("hsel_bin", (M, ld8(d)), self.act)
        ops.gather_rows(x, row_index, None, hsel)
        ab = head.hidden(hsel, M)
        W, b = self.store[f"{head.prefix}.out.w"], self.store[f"{head.prefix}.out.b"]
        z = self.pool.get("z_ml", (M, V))
        ops.gemm(ab, 0, kernel_of(W, ab), 1, M, V, h, bias=b.w, out_f32=z)
        probs = ops.sigmoid(z.view(-1), out=self.pool.get("p_ml", (M * V,)))
        y = y_f32.reshape(-1).contiguous()
        assert y.numel() == M * V, "labels must be (B, output_vocab_size)"
        stats = ops.masked_bce(y, probs, label_pad, pos_weight)
        self._allreduce(stats)
        if fp32:   # fp32-class mode: dz stays fp32, the GEMMs below split it (ops.gemm dispatch)
            dzb = self.pool.get("dz_ml32", (M, V))
            ops.sigmoid_bce_dz(y, probs, M, V, label_pad, pos_weight, stats, dz_f32=dzb)
        else:
            dzb = self.pool.get("dz_ml", (M, ld8(V)), BF16)
            ops.sigmoid_bce_dz(y, probs, M, V, label_pad, pos_weight, stats, dz_bf16=dzb)
        ops.gemm_splitk(ab, 1, dzb, 1, h, V, M, W.g, ws_name="splitk_vocab")
        ops.colsum_bf16(dzb, M, V, b.g)
        # d(ab) = dz W^T: K = V is long and M x h small -> split-K with a (ReLU-gated) reduce
        dsel = self.pool.get("dsel_bin", (M, d))
        mlp = head.mlp
        if mlp.dims:
            dab = self.pool.get("dab_bin", (M, ld8(mlp.out_dim)), self.act)
            ops.gemm_splitk_ex(dzb, 0, kernel_of(W, dzb), 0, M, h, V, ab, None, dab)
            mlp.backward(dab, dsel)
        else:
            ops.gemm_splitk_ex(dzb, 0, kernel_of(W, dzb), 0, M, h, V, None, dsel, None)
        dx = self.pool.get("dx_top", (B * S, d), zero=True)
        ops.scatter_rows(dsel, row_index, dx)
        self.transformer.engine.backward(dx, self.process_group)
        self._reduce_gradients()
        self._last_probs = probs.view(B, V)
        return stats

    def train_step(self, data, n_masked=None):
        """Keras Model.train_step: data = (inputs dict, labels (B, max_n_masked) float32).
        Returns {'loss': float} (+ metric results)."""
        inputs, y = data
        ids_list, B, S, starts, ends = self.prepare_inputs(inputs)
        y = torch.as_tensor(np.ascontiguousarray(y, dtype=np.float32)) if not torch.is_tensor(y) else y
        n_host = int((y != LABEL_PAD).sum().item())
        y = y.to(device="cuda", dtype=F32).contiguous()
        from .head import BinaryClassificationHead, MultiLabel_MultiClass_classification
        sigmoid_head = isinstance(self.head, (BinaryClassificationHead,
                                              MultiLabel_MultiClass_classification))
        if sigmoid_head and self.segment_to_head is not None:
            pw = getattr(self.loss, "pos_weight", None)
            step = (self.binary_forward_backward if isinstance(self.head, BinaryClassificationHead)
                    else self.multilabel_forward_backward)
            stats = step(ids_list, y, B, S, (starts, ends), pos_weight=pw, seed=self._next_seed())
            opt = self.optimizer or Adam()
            self.store.adam(self._next_lr(opt), opt.beta_1, opt.beta_2, opt.epsilon)
            s = stats.cpu().numpy()
            mean = float(s[0] / s[1]) if s[1] > 0 else 0.0
            return {'loss': mean / ((pw + 1.0) / 2) if pw is not None else mean}
        stats = self.cloze_forward_backward(ids_list, y, B, S, n_masked=n_host,
                                            seed=self._next_seed())
        # metrics first: they score the forward pass's hidden rows against the weights that
        # produced them (Keras computes train metrics on the forward predictions), then Adam
        for m in self.metrics:
            m.update_state(y, self._last_output)
        opt = self.optimizer or Adam()
        self.store.adam(self._next_lr(opt), opt.beta_1, opt.beta_2, opt.epsilon)
        s = stats.cpu().numpy()
        logs = {'loss': float(s[0] / s[1]) if s[1] > 0 else 0.0}
        for m in self.metrics:
            logs[m.name] = float(m.result())
        return logs

    def test_step(self, data):
        inputs, y = data
        out = self.call(inputs, training=False)
        logs = {}
        if self.loss is not None:
            logs['loss'] = float(self.loss(y, out))
        for m in self.metrics:
            m.update_state(y, out)
            logs[m.name] = float(m.result())
        return logs

    def fit(self, dataset, steps_per_epoch, epochs=1, verbose=0, validation_data=None,
            validation_steps=None, callbacks=()):
        """Keras Model.fit over an iterator of (inputs, labels) (examples/BERT4Rec/source/
        main.py:159-165): see training_utils.run_fit for the epoch / validation / callback order."""
        from .training_utils import run_fit
        return run_fit(self, dataset, steps_per_epoch, epochs=epochs, verbose=verbose,
                       validation_data=validation_data, validation_steps=validation_steps,
                       callbacks=callbacks)

    # ------------------------------------------------------------------ weights (N4)
    def get_weights(self):
        """{name: float32 ndarray} in the reference's per-layer layout (separate wq / wk / wv)."""
        from .weights import to_reference_layout
        return to_reference_layout(self.store.get_weights())

    def set_weights(self, weights):
        from .weights import to_store_layout
        self.store.set_weights(to_store_layout(weights))

    def save_weights(self, path):
        """One .npz keyed by the TF object-checkpoint keys of the reference's variables
        (weights.tf_checkpoint_key), so the same file can be produced on a TensorFlow box from a
        reference checkpoint and loaded here, and the other way round."""
        from .weights import export_reference_variables
        import torch.distributed as dist
        weights = self.store.get_weights()
        vocab = getattr(self.head, "vocab", None)
        if hasattr(vocab, "gather_full_weights"):   # vocabulary-parallel: every rank takes part
            weights.update(vocab.gather_full_weights())
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            return   # replicas are identical: one writer
        np.savez(path, **export_reference_variables(weights, self.transformer.engine.features))

    def load_weights(self, path):
        from .weights import import_reference_variables
        with np.load(path) as z:
            arrays = import_reference_variables({k: z[k] for k in z.files},
                                                self.transformer.engine.features)
        vocab = getattr(self.head, "vocab", None)
        if hasattr(vocab, "shard_full_weights"):
            arrays = vocab.shard_full_weights(arrays)
        self.store.set_weights(arrays)
