"""On-device Cloze batch builder against the oracle's keyed restatement: bit-exact
ids, labels and counts; error reporting; a built batch drives a training step."""
import numpy as np
import pytest
import torch

from oracle import clickpath_oracle as O

pytestmark = pytest.mark.gpu


def _sessions(rng, n, max_len, vocab):
    lens = rng.integers(1, max_len + 1, size=n)
    lens[:3] = [1, 2, max_len]
    lens[3:5] = [min(91, max_len), min(171, max_len)]   # 90 * 0.7 and 170 * 0.7: float32 != float64
    return [rng.integers(10, vocab + 10, size=int(l)).astype(np.int32) for l in lens]


@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("p,max_masked,max_len", [(0.4, 10, 50), (0.15, 30, 200), (1.0, 300, 300),
                                                (0.7, 300, 181)])
def test_device_builder_is_bit_exact_against_the_oracle(cuda_lib, mode, p, max_masked, max_len):
    from bert4clickpath_b200.data import DeviceClozeBuilder
    rng = np.random.default_rng(max_len)
    sessions = _sessions(rng, 300, max_len, 5000)
    builder = DeviceClozeBuilder(sessions)
    idx = np.concatenate([np.arange(5), 5 + rng.permutation(295)[:92]])
    for seed, (L, Mmax) in [(5, (None, None)), (2 ** 63 + 11, (max_len + 3, max_masked + 2))]:
        if Mmax is not None and mode == "train":
            Mmax = max(Mmax, builder.shapes(idx, mode, p, max_masked)[1])
        out = builder.build(idx, mode, seed, p, max_masked, L=L, Mmax=Mmax, check=True)
        ids, lab, n = O.keyed_cloze_batch(sessions, idx, mode, seed, p, max_masked, L=L, Mmax=Mmax)
        assert out["n_masked"] == n and out["S"] == ids.shape[1]
        assert np.array_equal(out["ids"].cpu().numpy(), ids)
        assert np.array_equal(out["labels"].cpu().numpy(), lab)
    # a different seed moves the masks (train), never the eval position
    a = builder.build(idx, mode, 1, p, max_masked)["ids"]
    b = builder.build(idx, mode, 2, p, max_masked)["ids"]
    assert torch.equal(a, b) == (mode == "eval" or p == 1.0)


def test_device_builder_reports_rows_that_do_not_fit(cuda_lib):
    from bert4clickpath_b200 import ops
    from bert4clickpath_b200.data import DeviceClozeBuilder
    sessions = [np.arange(10, 30, dtype=np.int32), np.arange(10, 14, dtype=np.int32)]
    builder = DeviceClozeBuilder(sessions)
    with pytest.raises(ValueError):
        builder.build([0, 1], "train", 0, L=5)
    with pytest.raises(ValueError):
        builder.build([0, 1], "sideways", 0)
    # bypass the host check: the kernel flags the row, writes pads, and leaves the others intact
    ids = torch.full((2, 8), 77, dtype=torch.int32, device="cuda")
    labels = torch.full((2, 4), 77.0, device="cuda")
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    idx = torch.tensor([0, 1], dtype=torch.int32, device="cuda")
    ops.cloze_build(builder.items, builder.offsets, idx, 2, 5, 4, True, 0.4, 10, 0,
                    (3, 4, 1, 0, 10), -1.0, ids, labels, count, status)
    assert int(status.item()) == 1
    assert (ids[0] == 0).all() and (labels[0] == -1).all()
    want, lab, n = O.keyed_cloze_batch(sessions, [1], "train", 0, 0.4, 10, L=5, Mmax=4)
    assert np.array_equal(ids[1].cpu().numpy(), want[0]) and np.array_equal(labels[1].cpu().numpy(), lab[0])
    assert int(count.item()) == n


def test_built_batches_train_the_model(cuda_lib):
    import bert4clickpath_b200 as bc
    from bert4clickpath_b200.data import DeviceClozeBuilder
    from bert4clickpath_b200.training import ClozeTrainStep, DeviceBatch
    V = 700
    rng = np.random.default_rng(0)
    # popularity-skewed items, so that there is something to learn within a few steps
    sessions = [(np.minimum(rng.geometric(0.02, size=int(l)), V) + 9).astype(np.int32)
                for l in rng.integers(5, 31, size=256)]
    builder = DeviceClozeBuilder(sessions)
    head = bc.SoftMaxHead(dense_layer_dims=[64, 128], output_vocab_size=V)
    model = bc.ClickstreamTransformer(
        sequential_input_config={"items": ["asin"]}, feature_vocabs={"items": V},
        embedding_dims={"items": 64}, head_unit=head, value_to_head=bc.INPUT_MASKING_TOKEN,
        num_encoder_layers=1, num_attention_heads=2, dropout_rate=0.0, seed=5)
    step = ClozeTrainStep(model, bc.Adam(1e-2))
    losses = []
    for t in range(12):
        idx = rng.permutation(256)[:64]
        out = builder.build(idx, "train", seed=t, L=30, Mmax=10)
        db = DeviceBatch([out["ids"].view(-1)], out["labels"], out["B"], out["S"], out["n_masked"])
        s = step.step_device(db).cpu().numpy()
        assert s[1] == out["n_masked"]
        losses.append(s[0] / s[1])
    assert losses[-1] < losses[0]


def test_device_builder_matches_committed_golden_batches(cuda_lib):
    import os
    from bert4clickpath_b200.data import DeviceClozeBuilder
    k = np.load(os.path.join(os.path.dirname(__file__), "golden", "keyed_cloze.npz"))
    offs = k["offsets"]
    builder = DeviceClozeBuilder([k["items"][offs[i]:offs[i + 1]] for i in range(len(offs) - 1)])
    tr = builder.build(k["idx"], "train", 1234, 0.4, 10, check=True)
    assert tr["ids"].cpu().numpy().tobytes() == k["train_ids"].tobytes()
    assert tr["labels"].cpu().numpy().tobytes() == k["train_labels"].tobytes()
    assert tr["n_masked"] == int(k["train_n"])
    ev = builder.build(k["idx"], "eval", 1234, 0.4, 10, L=52, Mmax=3, check=True)
    assert ev["ids"].cpu().numpy().tobytes() == k["eval_ids"].tobytes()
    assert ev["labels"].cpu().numpy().tobytes() == k["eval_labels"].tobytes()
    assert ev["n_masked"] == int(k["eval_n"])
