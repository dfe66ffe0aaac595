"""Weight import / export between the reference's per-layer Keras variables (separate wq / wk /
wv Dense kernels in (in, out) layout, transformer.py:112-116) and the fused buffers used here."""
import numpy as np


def to_store_layout(ref_params):
    """{'enc.l.wq', 'enc.l.wk', 'enc.l.wv', ...} -> {'enc.l.wqkv', ...}; other names unchanged."""
    out = {}
    for k, v in ref_params.items():
        parts = k.split(".")
        if parts[0] == "enc" and parts[2] in ("wq", "wk", "wv", "bq", "bk", "bv"):
            continue
        out[k] = np.asarray(v)
    layers = sorted({int(k.split(".")[1]) for k in ref_params if k.startswith("enc.")})
    for l in layers:
        out[f"enc.{l}.wqkv"] = np.concatenate(
            [ref_params[f"enc.{l}.w{n}"] for n in "qkv"], axis=1)
        out[f"enc.{l}.bqkv"] = np.concatenate([ref_params[f"enc.{l}.b{n}"] for n in "qkv"])
    return out


def to_reference_layout(store_params):
    out = {}
    for k, v in store_params.items():
        if k.endswith(".wqkv"):
            d = v.shape[0]
            for i, n in enumerate("qkv"):
                out[k[:-4] + "w" + n] = v[:, i * d:(i + 1) * d]
        elif k.endswith(".bqkv"):
            d = v.shape[0] // 3
            for i, n in enumerate("qkv"):
                out[k[:-4] + "b" + n] = v[i * d:(i + 1) * d]
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------- N4: name map
_TF_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def tf_checkpoint_key(name):
    """Reference-layout parameter name -> the key the same variable has in a TF2 object-based
    checkpoint written by the reference (`model.save_weights` / ModelCheckpoint,
    examples/BERT4Rec/source/main.py:112-118, :137-142).

    Such keys are the chain of Python attribute names from the model object to the variable, so
    they follow from the reference's source alone: `ClickstreamTransformer.transformer`
    (clickstream_transformer.py:220), `.head` (:196), `Transformer.encoder` / `.embedding_layers`
    (a dict keyed by feature; transformer.py:338, :346), `Encoder.enc_layers` (a list; :245),
    `EncoderLayer.mha / .ffn / .layernorm1 / .layernorm2` (:181-184), `MultiHeadAttention.wq /
    .wk / .wv / .dense` (:112-116), the two Dense layers of the `Sequential` ffn
    (`layer_with_weights-0/1`, :163-167), and the heads' `intermediate_layers` list and
    `output_layer` (head.py:10-11).  Derived, not verified: TensorFlow cannot be run here."""
    parts = name.split(".")
    if parts[0] == "emb":
        return f"transformer/embedding_layers/{'.'.join(parts[1:])}/embeddings" + _TF_SUFFIX
    if parts[0] == "enc":
        base = f"transformer/encoder/enc_layers/{int(parts[1])}/"
        leaf = {
            "wq": "mha/wq/kernel", "bq": "mha/wq/bias", "wk": "mha/wk/kernel", "bk": "mha/wk/bias",
            "wv": "mha/wv/kernel", "bv": "mha/wv/bias", "wo": "mha/dense/kernel",
            "bo": "mha/dense/bias",
            "w1": "ffn/layer_with_weights-0/kernel", "b1": "ffn/layer_with_weights-0/bias",
            "w2": "ffn/layer_with_weights-1/kernel", "b2": "ffn/layer_with_weights-1/bias",
            "ln1_g": "layernorm1/gamma", "ln1_b": "layernorm1/beta",
            "ln2_g": "layernorm2/gamma", "ln2_b": "layernorm2/beta",
        }[parts[2]]
        return base + leaf + _TF_SUFFIX
    if parts[0] == "head":
        kind = {"w": "kernel", "b": "bias"}[parts[2]]
        if parts[1] == "out":
            return f"head/output_layer/{kind}" + _TF_SUFFIX
        return f"head/intermediate_layers/{int(parts[1])}/{kind}" + _TF_SUFFIX
    raise KeyError(name)


def _table_name(name, features, to_reference):
    """The store registers embedding tables by position (`emb.0`, `emb.1`, ... in the key order of
    sequential_input_config); the reference's `embedding_layers` is a dict keyed by FEATURE NAME
    (transformer.py:346-355).  `features` (the engine's ordered feature list) converts between
    the two; without it names pass through unchanged."""
    parts = name.split(".")
    if parts[0] != "emb" or features is None:
        return name
    tail = ".".join(parts[1:])
    if to_reference:
        return f"emb.{features[int(tail)]}" if tail.isdigit() else name
    if tail in features:
        return f"emb.{list(features).index(tail)}"
    if tail.isdigit() and int(tail) < len(features):
        return name
    raise KeyError(f"embedding table {tail!r} is not one of the model's features {list(features)}")


def export_reference_variables(store_params, features=None):
    """{TF checkpoint key: float32 array in the Keras layout} for every parameter.  `features`:
    the model's ordered sequential features (`model.transformer.engine.features`), so that table
    `emb.<i>` is written under `.../embedding_layers/<feature>/embeddings` as the reference does."""
    return {tf_checkpoint_key(_table_name(k, features, True)): np.asarray(v, dtype=np.float32)
            for k, v in to_reference_layout(store_params).items()}


def import_reference_variables(variables, features=None, dtype=np.float32):
    """Inverse of export_reference_variables: a {TF checkpoint key: array} mapping (what
    `tf.train.load_checkpoint(...).get_tensor` yields on a TF box; optimizer slots and
    bookkeeping keys are ignored) -> the fused store layout, ready for `store.set_weights`.
    `features` maps the reference's feature-named tables onto the store's positional names.
    `dtype=None` keeps the arrays' own type (float64 gradients keyed like the variables)."""
    import re
    pats = [
        (re.compile(r"^transformer/embedding_layers/(.+)/embeddings$"), lambda m: f"emb.{m[1]}"),
        (re.compile(r"^transformer/encoder/enc_layers/(\d+)/mha/w([qkv])/(kernel|bias)$"),
         lambda m: f"enc.{m[1]}.{'w' if m[3] == 'kernel' else 'b'}{m[2]}"),
        (re.compile(r"^transformer/encoder/enc_layers/(\d+)/mha/dense/(kernel|bias)$"),
         lambda m: f"enc.{m[1]}.{'wo' if m[2] == 'kernel' else 'bo'}"),
        (re.compile(r"^transformer/encoder/enc_layers/(\d+)/ffn/layer_with_weights-([01])/(kernel|bias)$"),
         lambda m: f"enc.{m[1]}.{'w' if m[3] == 'kernel' else 'b'}{int(m[2]) + 1}"),
        (re.compile(r"^transformer/encoder/enc_layers/(\d+)/layernorm([12])/(gamma|beta)$"),
         lambda m: f"enc.{m[1]}.ln{m[2]}_{'g' if m[3] == 'gamma' else 'b'}"),
        (re.compile(r"^head/intermediate_layers/(\d+)/(kernel|bias)$"),
         lambda m: f"head.{m[1]}.{'w' if m[2] == 'kernel' else 'b'}"),
        (re.compile(r"^head/output_layer/(kernel|bias)$"),
         lambda m: f"head.out.{'w' if m[1] == 'kernel' else 'b'}"),
    ]
    ref = {}
    for key, arr in variables.items():
        if not key.endswith(_TF_SUFFIX) or key.startswith("optimizer/"):
            continue
        stem = key[:-len(_TF_SUFFIX)]
        for pat, name in pats:
            m = pat.match(stem)
            if m:
                ref[_table_name(name(m), features, False)] = np.asarray(arr, dtype=dtype)
                break
    return to_store_layout(ref)
