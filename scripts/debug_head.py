import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bert4clickpath_b200 as bc
from bert4clickpath_b200.synthetic import make_cloze_batch
from bert4clickpath_b200.weights import to_reference_layout
from oracle import clickpath_oracle as O
from oracle.mixed_precision import cloze_train_step_bf16

V, d, L, H, dff, hd = 300, 32, 2, 2, 100, [64, 32]
head = bc.SoftMaxHead(dense_layer_dims=hd, output_vocab_size=V)
model = bc.ClickstreamTransformer(
    sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": V},
    embedding_dims={"items": d}, head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
    num_encoder_layers=L, num_attention_heads=H, dropout_rate=0.0, encoder_ff_dim=dff)
batch = make_cloze_batch(np.random.default_rng(0), 64, V, max_len=20, mode="train", masked_percentage=0.4, lengths="beauty")
ids = torch.from_numpy(batch["ids"]).cuda().view(-1)
labels = torch.from_numpy(batch["labels"]).cuda()
B, S = batch["ids"].shape
M = batch["n_masked"]
stats = model.cloze_forward_backward([ids], labels, B, S, n_masked=M, training=False).cpu().numpy()
P = {k: v.astype(np.float64) for k, v in to_reference_layout(model.store.get_weights()).items()}
pe = O.positional_encoding(10000, d)
eloss, EG, ex = cloze_train_step_bf16([batch["ids"].astype(np.int64)], batch["labels"], P, L, H, pe)
dbg = ex["dbg"]
# emulation rows are in padded (B, Mmax) layout; map to compact
lab = batch["labels"].reshape(-1)
keep = lab >= 0
def cmp(name, got, want):
    want = want[keep]
    e = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    nz = (np.abs(got - want) > 1e-2 * np.abs(want).max()).sum()
    print(f"{name:10s} shape={got.shape} relerr={e:.4e} n_bad={nz} maxwant={np.abs(want).max():.3e}")
mlp, vocab = model.head.mlp, model.head.vocab
get = lambda pool, name: [v for (k, v) in pool._b.items() if k[0] == name][0]
acts = mlp.saved["acts"]
for i, a in enumerate(acts):
    cmp(f"act{i}", a.float().cpu().numpy()[:, :dbg["acts"][i].shape[1]], dbg["acts"][i])
cmp("dz", get(vocab.pool, "dz").float().cpu().numpy()[:, :V], dbg["dz"])
cmp("dzl1", get(model.pool, "dz_head").float().cpu().numpy()[:, :32], dbg["dzl1"])
cmp("dzl0", get(mlp.pool, "da1").float().cpu().numpy()[:, :64], dbg["dzl0"])
cmp("dsel", get(model.pool, "dsel").cpu().numpy(), dbg["dsel"])
g = to_reference_layout(model.store.get_grads())
for k in ("head.out.w", "head.1.w", "head.1.b", "head.0.w", "head.0.b"):
    e = np.abs(g[k] - EG[k]).max() / np.abs(EG[k]).max()
    print(k, f"{e:.4e}")
# direct check of the dW GEMM for head.0.w from the GPU's own operands
a0 = acts[0].float().cpu().numpy()[:, :32].astype(np.float64)
dz0 = get(mlp.pool, "da1").float().cpu().numpy()[:, :64].astype(np.float64)
ref = a0.T @ dz0
print("head.0.w vs its own GPU operands:", np.abs(g["head.0.w"] - ref).max() / np.abs(ref).max())
