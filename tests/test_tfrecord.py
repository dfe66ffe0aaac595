"""TFRecord / tf.train.Example reader without TensorFlow: CRC32C known answers, the
wire format cross-checked in both directions against the protobuf runtime (descriptors of
tensorflow/core/example/{example,feature}.proto rebuilt at run time), framing errors, and the
session reader feeding ClozeDataset exactly like the text reader."""
import os
import struct

import numpy as np
import pytest

from bert4clickpath_b200 import tfrecord as R
from bert4clickpath_b200.data import ClozeDataset, prepare_sessions, read_bert4rec_text_data

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _example_class():
    from google.protobuf import descriptor_pb2 as D, descriptor_pool, message_factory
    F = D.FieldDescriptorProto
    f = D.FileDescriptorProto(name="b4cp_example_test.proto", package="b4cp_test", syntax="proto3")
    for name, typ in (("BytesList", F.TYPE_BYTES), ("FloatList", F.TYPE_FLOAT), ("Int64List", F.TYPE_INT64)):
        m = f.message_type.add(name=name)
        m.field.add(name="value", number=1, type=typ, label=F.LABEL_REPEATED)
    feat = f.message_type.add(name="Feature")
    feat.oneof_decl.add(name="kind")
    for i, (n, t) in enumerate((("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List"))):
        feat.field.add(name=n, number=i + 1, type=F.TYPE_MESSAGE, type_name=f".b4cp_test.{t}",
                       label=F.LABEL_OPTIONAL, oneof_index=0)
    feats = f.message_type.add(name="Features")
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, type_name=".b4cp_test.Feature",
                    label=F.LABEL_OPTIONAL)
    feats.field.add(name="feature", number=1, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED,
                    type_name=".b4cp_test.Features.FeatureEntry")
    ex = f.message_type.add(name="Example")
    ex.field.add(name="features", number=1, type=F.TYPE_MESSAGE, type_name=".b4cp_test.Features",
                 label=F.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("b4cp_test.Example"))


def test_crc32c_known_answers_and_mask():
    assert R.crc32c(b"123456789") == 0xE3069283                     # the standard check value
    assert R.crc32c(b"\x00" * 32) == 0x8A9136AA                      # RFC 3720 B.4
    assert R.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert R.crc32c(bytes(range(32))) == 0x46DD794E
    assert R.crc32c(b"") == 0
    crc = R.crc32c(b"abc")
    assert R.masked_crc32c(b"abc") == (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_wire_format_matches_protobuf_runtime_both_ways():
    Example = _example_class()
    feats = {"reviewerID": "A1YJEY40YUW4SE", "asin": ["B004756YJA", "B004ZT0SSG", "7806397051"],
             "unixReviewTime": [1391040000, -5, 0, 2 ** 40], "score": [0.5, -2.25, 3.0],
             "empty": [], "one_int": 7, "raw": b"\x00\xff"}
    mine = R.encode_example(feats)
    msg = Example()
    msg.ParseFromString(mine)                                        # their parser reads my bytes
    fm = msg.features.feature
    assert list(fm["asin"].bytes_list.value) == [s.encode() for s in feats["asin"]]
    assert list(fm["reviewerID"].bytes_list.value) == [b"A1YJEY40YUW4SE"]
    assert list(fm["unixReviewTime"].int64_list.value) == feats["unixReviewTime"]
    assert list(fm["score"].float_list.value) == feats["score"]
    assert list(fm["one_int"].int64_list.value) == [7] and list(fm["raw"].bytes_list.value) == [b"\x00\xff"]
    assert fm["empty"].WhichOneof("kind") == "bytes_list" and len(fm["empty"].bytes_list.value) == 0
    theirs = msg.SerializeToString(deterministic=True)               # my parser reads their bytes
    got = R.decode_example(theirs)
    want = {"reviewerID": [b"A1YJEY40YUW4SE"], "asin": [s.encode() for s in feats["asin"]],
            "unixReviewTime": feats["unixReviewTime"], "score": feats["score"], "empty": [],
            "one_int": [7], "raw": [b"\x00\xff"]}
    assert got == want and R.decode_example(mine) == want
    assert theirs == mine                                            # same canonical bytes, even
    # unpacked repeated scalars (proto2-style writers) are accepted too
    unpacked = R._len_field(1, R._len_field(1, R._len_field(1, b"k") + R._len_field(2, R._len_field(
        3, R._varint((1 << 3) | 0) + R._varint(5) + R._varint((1 << 3) | 0) + R._varint(2 ** 64 - 1)))))
    assert R.decode_example(unpacked) == {"k": [5, -1]}
    m2 = Example()
    m2.ParseFromString(unpacked)
    assert list(m2.features.feature["k"].int64_list.value) == [5, -1]
    with pytest.raises(TypeError):
        R.encode_example({"x": [object()]})


def test_record_framing_round_trip_and_corruption(tmp_path):
    p = str(tmp_path / "a.tfrecord")
    payloads = [b"", b"x", os.urandom(1000)]
    R.write_records(p, payloads)
    assert list(R.read_records(p)) == payloads
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[:8])[0] == 0 and len(raw) == sum(len(x) + 16 for x in payloads)
    bad = bytearray(raw)
    bad[-10] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        list(R.read_records(p))
    assert len(list(R.read_records(p, verify_crc=False))) == 3
    open(p, "wb").write(raw[:-3])
    with pytest.raises(ValueError):
        list(R.read_records(p))
    open(p, "wb").write(raw[:5])
    with pytest.raises(ValueError):
        list(R.read_records(p))


def test_tfrecord_sessions_feed_the_cloze_dataset_like_the_text_reader(tmp_path):
    text = os.path.join(GOLD, "tiny_bert4rec.txt")
    users, items = read_bert4rec_text_data(text)
    sessions, vocab, order = prepare_sessions(users, items)
    a, b = str(tmp_path / "part-0.tfrecord"), str(tmp_path / "part-1.tfrecord")
    half = len(sessions) // 2
    R.write_sessions(a, order[:half], sessions[:half])
    R.write_sessions(b, order[half:], sessions[half:])
    u2, s2 = R.read_sessions([a, b])
    assert u2 == order and s2 == sessions
    vocab_file = str(tmp_path / "item_vocab.txt")
    open(vocab_file, "w").writelines("\n".join(vocab))              # data_prep/main.py:79-80
    ds_text = ClozeDataset(text)
    for ds in (ClozeDataset.from_tfrecord([a, b]), ClozeDataset.from_tfrecord([a, b], vocab=vocab_file)):
        assert ds.users == ds_text.users and ds.vocab == ds_text.vocab
        assert all(np.array_equal(x, y) for x, y in zip(ds.session_ids, ds_text.session_ids))
        b1 = next(ds.batches(4, "train", np.random.default_rng(3)))
        b2 = next(ds_text.batches(4, "train", np.random.default_rng(3)))
        assert np.array_equal(b1["ids"], b2["ids"]) and np.array_equal(b1["labels"], b2["labels"])
    # an unknown item falls into the single OOV bucket (StaticVocabularyTable)
    small = ClozeDataset.from_sessions([["a", "b", "zzz"]], vocab=["a", "b"])
    assert small.session_ids[0].tolist() == [10, 11, 12]
    # a record without the group feature is rejected
    R.write_records(a, [R.encode_example({"asin": ["x"]})])
    with pytest.raises(ValueError):
        R.read_sessions(a)


def test_create_cloze_dataset_yields_the_reference_contract(tmp_path):
    """input_pipeline.py:136-232: endless (features, labels) batches from TFRecord files or a
    generator; masked strings, '[PAD]' padding, float32 label indices padded with -1."""
    from bert4clickpath_b200.clickstream_transformer import StaticVocabularyTable
    from bert4clickpath_b200.constants import RESERVED_TOKENS
    from bert4clickpath_b200.data import create_cloze_dataset
    text = os.path.join(GOLD, "tiny_bert4rec.txt")
    users, items = read_bert4rec_text_data(text)
    sessions, vocab, order = prepare_sessions(users, items)
    by_user = dict(zip(order, sessions))
    rec = str(tmp_path / "amazon_beauty-0.tfrecord")
    R.write_sessions(rec, order, sessions)
    vocab_file = str(tmp_path / "item_vocab.txt")
    open(vocab_file, "w").writelines("\n".join(vocab))
    index = {t: i for i, t in enumerate(vocab)}
    table = StaticVocabularyTable(RESERVED_TOKENS + vocab)
    B = 4
    for mode, source in (("train", str(tmp_path / "*.tfrecord")), ("eval", str(tmp_path / "*.tfrecord")),
                         ("train", lambda: ({"reviewerID": u, "asin": s} for u, s in zip(order, sessions)))):
        ds = create_cloze_dataset(source, mode, B, vocab_file, rng=np.random.default_rng(1), shuffle_buffer=3)
        seen = []
        for _ in range(2 * len(order) // B + 1):          # more than one pass: the dataset repeats
            feats, labels = next(ds)
            assert set(feats) == {"reviewerID", "asin"} and labels.dtype == np.float32
            assert feats["asin"].shape[0] == B == labels.shape[0] == len(feats["reviewerID"])
            for b in range(B):
                u = feats["reviewerID"][b]
                seen.append(u)
                src = by_user[u][:-1] if mode == "train" else by_user[u]
                row = list(feats["asin"][b])
                body, pad = row[:len(src)], row[len(src):]
                assert all(t == "[PAD]" for t in pad)
                masked = [i for i, t in enumerate(body) if t == "[MASK]"]
                assert all(t == s for i, (t, s) in enumerate(zip(body, src)) if i not in masked)
                k = int((labels[b] != -1).sum())
                want_k = 1 if mode == "eval" else max(0, min(int(np.float32(len(src)) * np.float32(0.4)), 10))
                assert len(masked) == k == want_k and (labels[b, k:] == -1).all()
                assert [index[src[i]] for i in masked] == labels[b, :k].astype(int).tolist()
                if mode == "eval":
                    assert masked == [len(src) - 1]
            # the model's lookup turns the strings into the id layout of the ClozeDataset path
            ids = table.lookup(feats["asin"])
            assert ids.dtype == np.int32 and ((ids == 1) == (feats["asin"] == "[MASK]")).all()
            assert ((ids == 0) == (feats["asin"] == "[PAD]")).all() and ids.max() < len(vocab) + 10
        assert set(seen) == set(order)                     # every session shows up
        assert seen[:len(order)] != order                  # shuffled
    # labels outside the target vocabulary go to the OOV bucket
    small_vocab = str(tmp_path / "small.txt")
    open(small_vocab, "w").write(vocab[0] + "\n")
    _, lab = next(create_cloze_dataset(str(tmp_path / "*.tfrecord"), "eval", B, small_vocab,
                                       rng=np.random.default_rng(0)))
    assert set(lab.reshape(-1).tolist()) <= {0.0, 1.0}
    with pytest.raises(ValueError):
        next(create_cloze_dataset(rec, "predict", B, vocab_file))
    with pytest.raises(TypeError):
        next(create_cloze_dataset(123, "train", B, vocab_file))
    with pytest.raises(FileNotFoundError):
        next(create_cloze_dataset(str(tmp_path / "nope*"), "train", B, vocab_file))
