import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import bert4clickpath_b200 as bc
from bert4clickpath_b200.synthetic import make_cloze_batch
from bert4clickpath_b200.training import ClozeTrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
V, d = 54293, 64
head = bc.SoftMaxHead(dense_layer_dims=[1024, 512, 256, 128], output_vocab_size=V)
model = bc.ClickstreamTransformer({"items": ["asin"]}, {"items": V}, {"items": d}, head,
                                  value_to_head=bc.INPUT_MASKING_TOKEN, num_encoder_layers=2,
                                  num_attention_heads=2, dropout_rate=0.1)
tr = ClozeTrainStep(model)
rng = np.random.default_rng(0)
db = tr.to_device(make_cloze_batch(rng, B, V, 50, "train", 0.15, 10))
for _ in range(3): tr.step_device(db)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): tr.step_device(db)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B}: host enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step")
