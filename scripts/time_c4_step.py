"""Full Cloze training step at SURVEY.md C4 (scaled BERT4Rec): V = 1,000,000, d_model = 256,
max_len 200 (TRAIN S = 202), 4 layers, 4 heads, head [] -> V (h = 256), 29 masks / sequence.
python scripts/time_c4_step.py [B] [dff]"""
import json, sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bert4clickpath_b200 as bc
from bert4clickpath_b200 import ops
from bert4clickpath_b200.synthetic import make_cloze_batch
from bert4clickpath_b200.training import ClozeTrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dff = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
V, d = 1_000_000, 256
head = bc.SoftMaxHead(dense_layer_dims=[], output_vocab_size=V)
model = bc.ClickstreamTransformer({"items": ["asin"]}, {"items": V}, {"items": d}, head,
                                  value_to_head=bc.INPUT_MASKING_TOKEN, num_encoder_layers=4,
                                  num_attention_heads=4, dropout_rate=0.1, encoder_ff_dim=dff)
tr = ClozeTrainStep(model, use_graph=True)
rng = np.random.default_rng(0)
batches = [tr.to_device(make_cloze_batch(rng, B, V, 200, "train", 0.15, 30)) for _ in range(2)]
for i in range(5):
    tr.step_device(batches[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for i in range(n):
    st = tr.step_device(batches[i % 2])
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
# stage breakdown from eagerly launched steps with event brackets
tr.use_graph = False
tr.step_device(batches[0]); torch.cuda.synchronize()
ops.TIMER.reset(); ops.TIMER.enabled = True
for i in range(4):
    tr.step_device(batches[i % 2])
torch.cuda.synchronize()
tot = {k: v[0] / 4 for k, v in ops.TIMER.totals_ms().items()}
s = st.cpu().numpy()
print(json.dumps({"config": f"C4: V={V} d={d} S={batches[0].S} layers=4 heads=4 dff={dff} head=[]->V B={B} masks/seq=29",
                  "ms_per_step": ms, "seqs_per_sec": B / ms * 1e3, "loss": float(s[0] / max(s[1], 1)),
                  "n_masked": int(batches[0].n_masked), "stage_ms_eager": tot,
                  "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
