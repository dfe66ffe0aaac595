"""Encoder-only Transformer with the reference's construction surface
(clickstream_transformer/transformer.py:271-402), running on libb4cp kernels."""
import numpy as np
import torch

from . import ops
from .constants import INPUT_PAD, SEP
from .engine import EncoderEngine, ParamStore


def create_segment_markers(seq, sep=SEP):
    """Running count of SEP tokens per row (transformer.py:6-35; unused by the model)."""
    seq = np.asarray(seq)
    if seq.ndim != 2:
        raise ValueError('Expected 2-D tensor')
    return np.cumsum((seq == sep).astype(np.int32), axis=1)


def create_padding_mask(seq):
    """(batch, 1, 1, seq_len) float mask of INPUT_PAD positions (transformer.py:38-41).  The
    CUDA attention kernel derives the same mask from the ids; this host version is for callers."""
    seq = np.asarray(seq)
    return (seq == INPUT_PAD).astype(np.float32)[:, np.newaxis, np.newaxis, :]


def get_angles(pos, i, d_model):
    return pos * (1 / np.power(10000, (2 * (i // 2)) / np.float32(d_model)))


def positional_encoding(position, d_model):
    """Sinusoid table (transformer.py:44-61): float64 angles, sin on even / cos on odd columns,
    cast to float32.  Returns (position, d_model)."""
    ang = get_angles(np.arange(position)[:, np.newaxis], np.arange(d_model)[np.newaxis, :], d_model)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return ang.astype(np.float32)


class Transformer:
    """Transformer(num_layers, num_attention_heads, embedding_sizes, embedding_dims,
    encoder_ff_dim, dropout_rate, item_embedding_weights=None) — transformer.py:292-357.

    call(inputs: dict feature -> (B, S) int ids, training, mask) -> (B, S, d_model) fp32 tensor.
    """

    def __init__(self, num_layers, num_attention_heads, embedding_sizes, embedding_dims,
                 encoder_ff_dim, dropout_rate, item_embedding_weights=None, *, store=None,
                 seed=0, precision="bf16", **kwargs):
        assert set(embedding_sizes.keys()) == set(embedding_dims.keys()), \
            "embedding_sizes and embedding_dims must have the same set of keys."
        self.num_layers = num_layers
        self.num_attention_heads = num_attention_heads
        self.embedding_sizes = embedding_sizes
        self.embedding_dims = embedding_dims
        self.encoder_ff_dim = encoder_ff_dim
        self.dropout_rate = dropout_rate
        self.item_embedding_weights = item_embedding_weights
        self.maximum_position_encoding = 10000
        self.d_model = sum(embedding_dims.values())
        assert self.d_model % num_attention_heads == 0
        self._own_store = store is None
        self.store = ParamStore() if store is None else store
        self.engine = EncoderEngine(self.store, embedding_sizes, embedding_dims, num_layers,
                                    num_attention_heads, encoder_ff_dim, dropout_rate,
                                    np.random.default_rng(seed), self.maximum_position_encoding,
                                    precision=precision)
        self.pos_encoding = self.engine.pe_host[np.newaxis, ...]
        if self._own_store:
            self.store.finalize()

    def get_config(self):
        return {
            'num_layers': self.num_layers,
            'num_attention_heads': self.num_attention_heads,
            'embedding_sizes': self.embedding_sizes,
            'embedding_dims': self.embedding_dims,
            'encoder_ff_dim': self.encoder_ff_dim,
            'dropout_rate': self.dropout_rate,
            'item_embedding_weights': self.item_embedding_weights,
        }

    def _device_ids(self, inputs):
        ids_list, shape = [], None
        for name in self.engine.features:
            t = inputs[name]
            if not torch.is_tensor(t):
                t = torch.as_tensor(np.ascontiguousarray(t))
            if t.dim() != 2:
                raise ValueError('features must be (batch_size, seq_len)')
            shape = tuple(t.shape) if shape is None else shape
            assert tuple(t.shape) == shape, "all sequential features must share one shape"
            ids_list.append(t.to(device="cuda", dtype=torch.int32).contiguous().view(-1))
        return ids_list, shape

    def call(self, inputs, training=None, mask=None, seed=0):
        assert set(inputs.keys()) >= set(self.engine.features)
        ids_list, (B, S) = self._device_ids(inputs)
        x, _ = self.engine.forward(ids_list, B, S, bool(training), seed)
        return x.view(B, S, self.d_model)

    __call__ = call
