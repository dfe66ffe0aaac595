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
