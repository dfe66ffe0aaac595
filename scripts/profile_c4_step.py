"""Two eager C4 training steps for an ncu launch list (the second step's launches are the tail of
the list; the count is printed)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bert4clickpath_b200 as bc
from bert4clickpath_b200 import _lib
from bert4clickpath_b200.synthetic import make_cloze_batch
from bert4clickpath_b200.training import ClozeTrainStep
B, V, d = 256, 1_000_000, 256
head = bc.SoftMaxHead(dense_layer_dims=[], output_vocab_size=V)
model = bc.ClickstreamTransformer({"items": ["asin"]}, {"items": V}, {"items": d}, head,
                                  value_to_head=bc.INPUT_MASKING_TOKEN, num_encoder_layers=4,
                                  num_attention_heads=4, dropout_rate=0.1, encoder_ff_dim=1024)
tr = ClozeTrainStep(model, use_graph=False)
db = tr.to_device(make_cloze_batch(np.random.default_rng(0), B, V, 200, "train", 0.15, 30))
tr.step_device(db); torch.cuda.synchronize()
n0 = _lib.lib().b4cp_launch_count()
tr.step_device(db); torch.cuda.synchronize()
print("launches_last_step", _lib.lib().b4cp_launch_count() - n0)
