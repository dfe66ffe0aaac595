import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bert4clickpath_b200 as bc
from bert4clickpath_b200.synthetic import make_cloze_batch
from bert4clickpath_b200.weights import to_reference_layout
from oracle import clickpath_oracle as O

def run(V, d, L, H, dff, hd, B, max_len, lengths, mp):
    head = bc.SoftMaxHead(dense_layer_dims=hd, output_vocab_size=V)
    model = bc.ClickstreamTransformer(
        sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": V},
        embedding_dims={"items": d}, head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
        num_encoder_layers=L, num_attention_heads=H, dropout_rate=0.0, encoder_ff_dim=dff)
    rng = np.random.default_rng(0)
    batch = make_cloze_batch(rng, B, V, max_len=max_len, mode="train", masked_percentage=mp, lengths=lengths)
    ids = torch.from_numpy(batch["ids"]).cuda().view(-1)
    labels = torch.from_numpy(batch["labels"]).cuda()
    B, S = batch["ids"].shape
    stats = model.cloze_forward_backward([ids], labels, B, S, n_masked=batch["n_masked"], training=False).cpu().numpy()
    P = {k: v.astype(np.float64) for k, v in to_reference_layout(model.store.get_weights()).items()}
    pe = O.positional_encoding(10000, d)
    loss, G, _ = O.cloze_train_step([batch["ids"].astype(np.int64)], batch["labels"], P, L, H, pe)
    grads = to_reference_layout(model.store.get_grads())
    gmax = max(np.abs(v).max() for v in G.values())
    print(f"V={V} d={d} dff={dff} hd={hd} B={B} S={S} M={batch['n_masked']} {lengths}: loss {stats[0]/stats[1]:.5f} vs {loss:.5f}")
    for k, w in G.items():
        err = np.abs(grads[k] - w).max() / max(np.abs(w).max(), 1e-2 * gmax)
        if k in ("head.out.w", "head.0.w", "head.1.w", "enc.0.wo", "emb.0"):
            print(f"   {k:14s} err={err:.4f} wmax={np.abs(w).max():.3e}")


run(300, 32, 1, 2, 96, [64, 32], 64, 20, "dense", 0.4)   # a: 2-layer head, M=448
run(300, 32, 1, 2, 96, [64], 16, 20, "dense", 0.4)       # b: 1-layer head, M=112
run(300, 32, 1, 2, 96, [64, 64], 64, 20, "dense", 0.4)   # c: h=64
run(300, 32, 1, 2, 96, [32], 64, 20, "dense", 0.4)       # d: 1-layer head h=32
run(300, 32, 1, 2, 96, [128, 32], 256, 20, "dense", 0.4) # e
