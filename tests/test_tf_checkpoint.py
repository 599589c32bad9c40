"""Row N1: TensorFlow checkpoint (tensor bundle) reader / writer without TensorFlow.
PARITY UNPINNED: no TensorFlow-written checkpoint exists offline, so the two directions are validated against each
other, against hand-built table blocks / snappy streams, and against the published crc32c check value."""
import struct

import numpy as np
import pytest

from alignnet_b200 import tf_checkpoint as T


def _tensors():
    rng = np.random.default_rng(0)
    t = {f"siamese/transformer1/embedding/conv{i}/weights": rng.normal(size=(1, 3 if i == 1 else 1, 4, 8)).astype(np.float32)
         for i in (1, 2, 3)}
    t.update({f"siamese/transformer1/mlp/fc{i}/biases": rng.normal(size=(5 + i,)).astype(np.float32) for i in range(40)})
    t["Variable"] = np.array(1234, np.int32)
    t["beta1_power"] = np.array(0.5, np.float32)
    t["counts"] = np.arange(7, dtype=np.int64).reshape(7, 1)
    t["flags"] = np.array([True, False, True])
    t["empty"] = np.zeros((0, 3), np.float32)
    return t


@pytest.mark.parametrize("block_size,restart", [(4096, 16), (64, 1), (200, 4), (1 << 20, 1000)])
def test_round_trip(tmp_path, block_size, restart):
    t = _tensors()
    prefix = str(tmp_path / "model.ckpt")
    T.write_checkpoint(prefix, t, block_size=block_size, restart_interval=restart)
    back = T.read_checkpoint(prefix)
    assert set(back) == set(t)
    for k in t:
        assert back[k].dtype == t[k].dtype and back[k].shape == t[k].shape, k
        np.testing.assert_array_equal(back[k], t[k])
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and raw[-8:] == bytes.fromhex("57fb808b247547db")
    keys = [k for k, _ in T.read_table(prefix + ".index")]
    assert keys == sorted(keys) and keys[0] == b""                      # header entry first, names in byte order


def test_crc32c_and_mask_known_answers():
    assert T.crc32c(b"123456789") == 0xE3069283                         # the standard CRC-32C check value
    assert T.crc32c(b"") == 0
    assert T.mask_crc(0) == 0xA282EAD8
    # the chunk-parallel path (large tensors) agrees with the byte-serial definition, ragged tail included
    rng = np.random.default_rng(1)
    for n in (64 * 1024, 64 * 1024 + 1, 200_003):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.crc32c(data) == T._crc32c_scalar(data), n


def test_snappy_blocks_and_prefix_compressed_keys():
    # literal 'abc' + copy(offset 3, length 6): "abcabcabc"
    assert T.snappy_decompress(bytes([9, (3 - 1) << 2]) + b"abc" + bytes([((6 - 4) << 2) | 1, 3])) == b"abcabcabc"
    # long literal (one extra length byte: tag 60)
    payload = bytes(range(200))
    assert T.snappy_decompress(T._put_varint(len(payload)) + bytes([60 << 2, len(payload) - 1]) + payload) == payload
    # overlapping copy (run-length): 'a' + copy(offset 1, length 7)
    assert T.snappy_decompress(bytes([8, 0]) + b"a" + bytes([((7 - 4) << 2) | 1, 1])) == b"a" * 8
    block = T._build_block([(b"siamese/a", b"1"), (b"siamese/ab", b"22"), (b"siamese_1/a", b"333")], restart_interval=16)
    assert T._block_entries(block) == [(b"siamese/a", b"1"), (b"siamese/ab", b"22"), (b"siamese_1/a", b"333")]
    assert len(block) < sum(len(k) + len(v) for k, v in T._block_entries(block)) + 3 * 3 + 8   # prefixes were shared
    # a snappy-compressed block (literal-only stream) is read like an uncompressed one
    comp = T._put_varint(len(block)) + bytes([(len(block) - 1) << 2]) + block if len(block) <= 60 else None
    if comp is not None:
        assert T._read_block(comp + b"\x01" + b"\0\0\0\0", 0, len(comp)) == block


def test_doubled_scope_ema_names_are_found():
    name = "siamese/transformer1/embedding/conv1/bn/moments/Squeeze/ExponentialMovingAverage"
    doubled = "siamese/transformer1/embedding/conv1/bn/" + name
    v = np.ones(4, np.float32)
    assert T._lookup({name: v}, name) is v
    assert T._lookup({doubled: v}, name) is v
    assert T._lookup({"other": v}, name) is None


def test_ema_shadow_keys_follow_tf_slot_creator_rule():
    """[TF-sem] ExponentialMovingAverage.apply -> slot_creator.create_zeros_slot names the shadow
    `<current variable scope>/<op name of the averaged tensor>/ExponentialMovingAverage`.  The second siamese branch
    re-enters variable scope `siamese` (tp8.py:142, AUTO_REUSE) -- only its name scope becomes `siamese_1`."""
    b1 = "siamese/transformer1/embedding/conv1/bn/moments/Squeeze/ExponentialMovingAverage"
    b2 = "siamese_1/transformer2/mlp/fc2/bn/moments/Squeeze_1/ExponentialMovingAverage"
    hd = "fc1/bn/moments/Squeeze/ExponentialMovingAverage"
    assert T.tf_ema_key(b1) == "siamese/transformer1/embedding/conv1/bn/" + b1
    assert T.tf_ema_key(b2) == "siamese/transformer2/mlp/fc2/bn/" + b2          # variable scope stays `siamese/`
    assert T.tf_ema_key(hd) == "fc1/bn/" + hd
    assert T.tf_ema_key("siamese_1/embedding/conv1/bn/gamma") == "siamese_1/embedding/conv1/bn/gamma"   # not a shadow
    v = np.ones(4, np.float32)
    assert T._lookup({T.tf_ema_key(b2): v}, b2) is v                 # a reference checkpoint's key for a branch-2 shadow
    assert T._lookup({"siamese_1/transformer2/mlp/fc2/bn/" + b2: v}, b2) is v   # files of this module's first version
    assert T._lookup({"x/" + b2: v, "y/" + b2: v}, b2) is None       # ambiguous suffix matches are refused


def test_bad_files_are_rejected(tmp_path):
    p = tmp_path / "x.index"
    p.write_bytes(b"\0" * 100)
    with pytest.raises(ValueError):
        T.read_table(str(p))


@pytest.mark.gpu
def test_engine_export_import_round_trip(tmp_path):
    import torch
    import __graft_entry__ as ge
    ge.build()
    from alignnet_b200 import engine, synth
    a = engine.Engine(engine.shipped_arch(), "cuda:0", "fp32", seed=1)
    batch = {k: torch.from_numpy(v).cuda() for k, v in synth.make_batch_fast(8, 32, seed=2).items()}
    for i in range(2):
        a.train_step(batch, lr=0.01, bn_decay=0.5, seed=i)
    prefix = str(tmp_path / "model.ckpt")
    T.save_from_engine(a, prefix)
    ck = T.read_checkpoint(prefix)
    assert ck["siamese/transformer1/embedding/conv1/weights"].shape == (1, 3, 1, 64)      # TF kernel shape
    assert ck["fc3/weights"].shape == (256, 103) and int(ck["Variable"]) == 2
    # shadows are written under TensorFlow's keys (both branches under variable scope `siamese/`)
    assert "siamese/embedding/conv3/bn/siamese_1/embedding/conv3/bn/moments/Squeeze/ExponentialMovingAverage" in ck
    assert "fc2/bn/fc2/bn/moments/Squeeze_1/ExponentialMovingAverage" in ck
    assert not any(k.startswith("siamese_1/") and k.endswith("ExponentialMovingAverage") for k in ck)
    b = engine.Engine(engine.shipped_arch(), "cuda:0", "fp32", seed=99)
    info = T.load_into_engine(b, prefix)
    assert info["missing"] == [] and info["unused"] == []
    assert b.step == 2
    torch.testing.assert_close(b.params, a.params, rtol=0, atol=0)
    torch.testing.assert_close(b.bn_state, a.bn_state, rtol=0, atol=0)
    torch.testing.assert_close(b.adam_m, a.adam_m, rtol=0, atol=0)
    torch.testing.assert_close(b.adam_v, a.adam_v, rtol=0, atol=0)
    ea, eb = a.forward(batch["pcs1"], batch["pcs2"], False), b.forward(batch["pcs1"], batch["pcs2"], False)
    for k in ea:
        torch.testing.assert_close(eb[k], ea[k], rtol=0, atol=1e-6)


def test_proto_layer_and_crc_against_tensorflows_own_definitions():
    """Third-party pin of the parts of the checkpoint format that CAN be pinned offline.  TensorBoard (installed here) ships
    TensorFlow's own generated protos for the sub-messages of a bundle entry (TensorShapeProto, DataType, VersionDef) and its
    own crc32c / crc mask used for TensorFlow record files.  The two bundle messages themselves
    (tensorflow/core/protobuf/tensor_bundle.proto) are declared here over those definitions and run through Google's protobuf
    runtime: what this module writes must parse into them field for field, and the runtime's serialisation of the same
    values must be byte-identical to this module's encoder.  (The LevelDB table container around the entries stays
    unpinned: no independent reader of that format is installed.)"""
    pytest.importorskip("tensorboard")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2
    from tensorboard.compat.tensorflow_stub import dtypes as tb_dtypes, pywrap_tensorflow as pw
    from alignnet_b200 import tf_checkpoint as T

    # crc32c and the LevelDB / TensorFlow mask
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 64, 4097, 300000):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.crc32c(data) == pw.crc32c(data), n
        assert T.mask_crc(T.crc32c(data)) == pw.masked_crc32c(data), n

    # DataType numbers -> numpy dtypes
    for num, np_t in T.DT.items():
        assert tb_dtypes.as_dtype(num).as_numpy_dtype == np_t, num
    assert T.DT_INV[np.dtype(np.float32).str] == types_pb2.DT_FLOAT and T.DT_INV[np.dtype(np.int64).str] == types_pb2.DT_INT64

    # tensor_bundle.proto over TensorFlow's own sub-message definitions
    pkg = tensor_shape_pb2.DESCRIPTOR.package
    fd = descriptor_pb2.FileDescriptorProto(name="an3d_test/tensor_bundle.proto", package="an3d_test", syntax="proto3")
    fd.dependency.extend([tensor_shape_pb2.DESCRIPTOR.name, types_pb2.DESCRIPTOR.name, versions_pb2.DESCRIPTOR.name])
    F = descriptor_pb2.FieldDescriptorProto
    hdr = fd.message_type.add(name="BundleHeaderProto")
    hdr.field.add(name="num_shards", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    hdr.field.add(name="endianness", number=2, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)      # enum {LITTLE = 0, BIG = 1}
    hdr.field.add(name="version", number=3, type=F.TYPE_MESSAGE, type_name=f".{pkg}.VersionDef", label=F.LABEL_OPTIONAL)
    ent = fd.message_type.add(name="BundleEntryProto")
    ent.field.add(name="dtype", number=1, type=F.TYPE_ENUM, type_name=f".{pkg}.DataType", label=F.LABEL_OPTIONAL)
    ent.field.add(name="shape", number=2, type=F.TYPE_MESSAGE, type_name=f".{pkg}.TensorShapeProto", label=F.LABEL_OPTIONAL)
    ent.field.add(name="shard_id", number=3, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    ent.field.add(name="offset", number=4, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
    ent.field.add(name="size", number=5, type=F.TYPE_INT64, label=F.LABEL_OPTIONAL)
    ent.field.add(name="crc32c", number=6, type=F.TYPE_FIXED32, label=F.LABEL_OPTIONAL)
    pool = descriptor_pool.Default()
    try:
        file_desc = pool.Add(fd) if hasattr(pool, "Add") else pool.AddSerializedFile(fd.SerializeToString())
    except TypeError:
        file_desc = pool.AddSerializedFile(fd.SerializeToString())
    file_desc = pool.FindFileByName("an3d_test/tensor_bundle.proto")
    Header = message_factory.GetMessageClass(file_desc.message_types_by_name["BundleHeaderProto"])
    Entry = message_factory.GetMessageClass(file_desc.message_types_by_name["BundleEntryProto"])

    cases = [(types_pb2.DT_FLOAT, (1, 3, 1, 64), 0, 768, 0xDEADBEEF), (types_pb2.DT_FLOAT, (2048, 512), 123456789012, 4194304, 1),
             (types_pb2.DT_INT64, (), 17, 8, 0x80000000), (types_pb2.DT_INT32, (0,), 5, 0, 0), (types_pb2.DT_FLOAT, (1024,), 4096, 4096, 77)]
    for dtype, shape, offset, size, crc in cases:
        mine = T._encode_entry(dtype, shape, offset, size, crc)
        e = Entry()
        e.ParseFromString(mine)                                    # Google's runtime reads what this module writes ...
        assert e.dtype == dtype and tuple(d.size for d in e.shape.dim) == tuple(shape)
        assert (e.shard_id, e.offset, e.size, e.crc32c) == (0, offset, size, crc)
        theirs = Entry(dtype=dtype, shape=tensor_shape_pb2.TensorShapeProto(dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=s) for s in shape]),
                       offset=offset, size=size, crc32c=crc).SerializeToString(deterministic=True)
        if size and crc:
            assert theirs == mine, (shape, theirs.hex(), mine.hex())    # ... and writes the same bytes (proto3 skips zero fields)
        d = T._decode_entry(theirs)                                # and this module reads what the runtime writes
        assert (d["dtype"], tuple(d["shape"]), d["offset"], d["size"], d["crc32c"]) == (dtype, tuple(shape), offset, size, crc)
    h = Header()
    h.ParseFromString(b"\x08\x01" + b"\x1a\x02\x08\x01")           # the header record write_checkpoint emits
    assert h.num_shards == 1 and h.endianness == 0 and h.version.producer == 1
    assert Header(num_shards=1, version=versions_pb2.VersionDef(producer=1)).SerializeToString(deterministic=True) == b"\x08\x01\x1a\x02\x08\x01"


def test_snappy_decoder_against_pyarrows_snappy():
    """The block decompressor (LevelDB tables may carry snappy-compressed blocks) against the snappy library bundled
    with pyarrow: literals, short and long copies, incompressible input."""
    pa = pytest.importorskip("pyarrow")
    if not pa.Codec.is_available("snappy"):
        pytest.skip("pyarrow built without snappy")
    from alignnet_b200 import tf_checkpoint as T
    rng = np.random.default_rng(0)
    for data in (b"", b"a", b"abcabcabcabc" * 100, rng.integers(0, 4, 100000, dtype=np.uint8).tobytes(),
                 rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(), b"siamese/transformer1/embedding/conv1/weights" * 50):
        assert T.snappy_decompress(pa.compress(data, codec="snappy", asbytes=True)) == data
