"""TensorFlow checkpoint (tensor-bundle, "V2") import / export without TensorFlow.

The reference saves and restores its models with `tf.train.Saver` (train.py:220,250-293,316-322), i.e. as
`<prefix>.index` + `<prefix>.data-00000-of-00001`.  TensorFlow is not installable here, so the format is restated from
its specification:

  <prefix>.index   a LevelDB-format table (tensorflow/core/lib/io/table*, format.h): data blocks of prefix-compressed
                   key/value entries with a restart array, each block followed by a 5-byte trailer (compression type,
                   masked crc32c), an index block of BlockHandles, and a 48-byte footer ending in the magic
                   0xdb4775248b80fb57.  Key "" holds a BundleHeaderProto, every other key is a tensor name whose value
                   is a BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}
                   (tensorflow/core/protobuf/tensor_bundle.proto).
  <prefix>.data-*  the tensors' raw little-endian bytes at the recorded offsets.

PARITY PARTLY PINNED for this row: no TensorFlow-written checkpoint is available offline (the reference ships none; its
released weights are an external download, README.md:92).  What can be pinned to third-party code is
(tests/test_tf_checkpoint.py::test_proto_layer_and_crc_against_tensorflows_own_definitions): the entry / header records
parse under Google's protobuf runtime over TensorFlow's own generated TensorShapeProto / DataType / VersionDef (shipped
with TensorBoard) and that runtime serialises the same values to identical bytes; crc32c and its mask equal TensorBoard's
implementation for TensorFlow record files; DataType numbers map to the same numpy dtypes.  The LevelDB table container
around the records stays UNPINNED (no independent reader of that format is installed): reader and writer are validated
against each other and against hand-built blocks, not against a file produced by TensorFlow.  The reader
accepts what a real Saver emits beyond what the writer produces: several data blocks, shared key prefixes, snappy-
compressed blocks, unknown proto fields, and the EMA shadow names TF's slot creator produces
(`<variable scope>/bn/<op name>/ExponentialMovingAverage`, see `tf_ema_key`: the second siamese branch reads
`siamese/X/bn/siamese_1/X/bn/moments/Squeeze/ExponentialMovingAverage`).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

MAGIC = 0xDB4775248B80FB57
DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_, 4: np.uint8, 6: np.int8, 5: np.int16}
DT_INV = {np.dtype(v).str: k for k, v in DT.items()}

# ------------------------------------------------------------------------------------------------------------------
# varints, protobuf wire format (just enough for the two bundle messages)
# ------------------------------------------------------------------------------------------------------------------


def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """[(field number, wire type, value)]: varint -> int, fixed32/64 -> int, length-delimited -> bytes."""
    out, pos = [], 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.append((field, wt, v))
    return out


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= 1 << 63 else v


def _decode_entry(buf: bytes) -> Dict:
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=0, sliced=False)
    for field, _, v in _parse_fields(buf):
        if field == 1:
            e["dtype"] = v
        elif field == 2:                                   # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
            for f2, _, v2 in _parse_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _parse_fields(v2):
                        if f3 == 1:
                            size = _signed(v3)
                    e["shape"].append(size)
        elif field == 3:
            e["shard_id"] = v
        elif field == 4:
            e["offset"] = _signed(v)
        elif field == 5:
            e["size"] = _signed(v)
        elif field == 6:
            e["crc32c"] = v
        elif field == 7:
            e["sliced"] = True
    return e


def _encode_entry(dtype: int, shape: Iterable[int], offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


# ------------------------------------------------------------------------------------------------------------------
# crc32c (Castagnoli), masked as in LevelDB / TensorFlow
# ------------------------------------------------------------------------------------------------------------------
_CRC_TABLE = None
_POLY = 0x82F63B78


def _table() -> np.ndarray:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ _POLY if c & 1 else c >> 1
            tbl.append(c)
        _CRC_TABLE = np.array(tbl, dtype=np.uint32)
    return _CRC_TABLE


def _crc32c_scalar(data: bytes) -> int:
    tbl = _table()
    crc = 0xFFFFFFFF
    for b in data:
        crc = int(tbl[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _gf2_times(mat: List[int], vec: int) -> int:
    out, i = 0, 0
    while vec:
        if vec & 1:
            out ^= mat[i]
        vec >>= 1
        i += 1
    return out


def _gf2_square(mat: List[int]) -> List[int]:
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def _zeros_operator(nbytes: int) -> List[int]:
    """GF(2) matrix that advances a CRC register over `nbytes` zero bytes (zlib's crc32_combine construction)."""
    odd = [_POLY] + [1 << n for n in range(31)]            # operator for one zero bit
    even = _gf2_square(odd)                                 # two bits
    odd = _gf2_square(even)                                 # four bits
    op = [1 << n for n in range(32)]                        # identity
    n = nbytes
    while n:
        even = _gf2_square(odd)                             # first pass: one byte
        if n & 1:
            op = [_gf2_times(even, c) for c in op]
        n >>= 1
        if not n:
            break
        odd = _gf2_square(even)
        if n & 1:
            op = [_gf2_times(odd, c) for c in op]
        n >>= 1
    return op


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli).  Large buffers (checkpoint tensors) are cut into equal chunks whose CRCs are computed
    simultaneously with numpy (one table look-up per byte POSITION, vectorised over the chunks) and merged with the
    zero-extension operator of zlib's crc32_combine."""
    n = len(data)
    chunk = 1024
    if n < 64 * chunk:
        return _crc32c_scalar(data)
    tbl = _table()
    nfull = n // chunk
    arr = np.frombuffer(data, dtype=np.uint8, count=nfull * chunk).reshape(nfull, chunk)
    state = np.full(nfull, 0xFFFFFFFF, dtype=np.uint32)
    for j in range(chunk):
        state = tbl[(state ^ arr[:, j]) & np.uint32(0xFF)] ^ (state >> np.uint32(8))
    crcs = (state ^ np.uint32(0xFFFFFFFF)).tolist()
    op = _zeros_operator(chunk)
    crc = crcs[0]
    for c in crcs[1:]:
        crc = _gf2_times(op, crc) ^ c
    tail = data[nfull * chunk:]
    if tail:
        crc = _gf2_times(_zeros_operator(len(tail)), crc) ^ _crc32c_scalar(tail)
    return crc


def mask_crc(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------------------------
# snappy (raw format) decoder -- TensorFlow's table builder may compress index blocks
# ------------------------------------------------------------------------------------------------------------------


def snappy_decompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                       # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(ln):                                 # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------------------------------------------------
# LevelDB-format table
# ------------------------------------------------------------------------------------------------------------------


def _read_block(data: bytes, offset: int, size: int) -> bytes:
    raw, ctype = data[offset:offset + size], data[offset + size]
    if ctype == 0:
        return raw
    if ctype == 1:
        return snappy_decompress(raw)
    raise ValueError(f"unknown block compression type {ctype}")


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    out, pos, key = [], 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path: str) -> List[Tuple[bytes, bytes]]:
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != MAGIC:
        raise ValueError(f"{path}: not a TensorFlow checkpoint index (bad table magic)")
    footer = data[-48:]
    _, p = _get_varint(footer, 0)            # metaindex handle (unused)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    entries: List[Tuple[bytes, bytes]] = []
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        off, q = _get_varint(handle, 0)
        size, _ = _get_varint(handle, q)
        entries += _block_entries(_read_block(data, off, size))
    return entries


def _build_block(items: List[Tuple[bytes, bytes]], restart_interval: int) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path: str, items: List[Tuple[bytes, bytes]], block_size: int = 4096, restart_interval: int = 16) -> None:
    items = sorted(items)
    blob, index = bytearray(), []

    def emit(block: bytes) -> Tuple[int, int]:
        off = len(blob)
        blob.extend(block + b"\x00" + struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return off, len(block)

    cur: List[Tuple[bytes, bytes]] = []
    size = 0
    for kv in items:
        cur.append(kv)
        size += len(kv[0]) + len(kv[1]) + 3
        if size >= block_size:
            off, n = emit(_build_block(cur, restart_interval))
            index.append((cur[-1][0], _put_varint(off) + _put_varint(n)))
            cur, size = [], 0
    if cur or not index:
        off, n = emit(_build_block(cur, restart_interval))
        index.append((cur[-1][0] if cur else b"", _put_varint(off) + _put_varint(n)))
    moff, mn = emit(_build_block([], restart_interval))
    ioff, isz = emit(_build_block(index, 1))
    footer = _put_varint(moff) + _put_varint(mn) + _put_varint(ioff) + _put_varint(isz)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    with open(path, "wb") as fh:
        fh.write(bytes(blob) + footer)


# ------------------------------------------------------------------------------------------------------------------
# checkpoints
# ------------------------------------------------------------------------------------------------------------------


def read_checkpoint(prefix: str) -> Dict[str, np.ndarray]:
    """All tensors of `<prefix>.index` / `<prefix>.data-*` by variable name (what tf.train.load_checkpoint exposes)."""
    entries = read_table(prefix + ".index")
    num_shards = 1
    out: Dict[str, np.ndarray] = {}
    shards: Dict[int, bytes] = {}
    for key, value in entries:
        if key == b"":
            for field, _, v in _parse_fields(value):
                if field == 1:
                    num_shards = v
                if field == 2 and v != 0:
                    raise ValueError("big-endian checkpoints are not supported")
            continue
        e = _decode_entry(value)
        if e["sliced"]:
            raise ValueError(f"{key.decode()}: partitioned (sliced) variables are not supported")
        if e["dtype"] not in DT:
            continue                                        # e.g. DT_STRING bookkeeping entries
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", "rb").read()
        dt = np.dtype(DT[e["dtype"]])
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        arr = np.frombuffer(raw, dtype=dt.newbyteorder("<")).astype(dt)
        out[key.decode()] = arr.reshape(e["shape"])
    return out


def write_checkpoint(prefix: str, tensors: Dict[str, np.ndarray], block_size: int = 4096, restart_interval: int = 16) -> None:
    """Writes a single-shard tensor bundle readable by `tf.train.Saver.restore` / `tf.train.load_checkpoint`."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items: List[Tuple[bytes, bytes]] = [(b"", b"\x08\x01" + b"\x1a\x02\x08\x01")]   # num_shards=1, version{producer=1}
    data = bytearray()
    for name in sorted(tensors):
        a = np.asarray(tensors[name], order="C")       # (ascontiguousarray would turn a scalar into shape (1,))
        if a.dtype.str.replace(">", "<") not in DT_INV and a.dtype.str not in DT_INV:
            raise ValueError(f"{name}: dtype {a.dtype} has no TensorFlow equivalent here")
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        items.append((name.encode(), _encode_entry(DT_INV[a.dtype.str], a.shape, len(data), len(raw), mask_crc(crc32c(raw)))))
        data += raw
    write_table(prefix + ".index", items, block_size, restart_interval)
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))


def tf_ema_key(name: str) -> str:
    """The key TensorFlow gives the EMA shadow the engine calls `name` [TF-sem, slot_creator.create_zeros_slot inside
    tf.train.ExponentialMovingAverage.apply]: `<current VARIABLE scope>/<primary tensor's OP name>/ExponentialMovingAverage`.
    The variable scope of the second siamese branch is still `siamese/...` (tp8.py:142 re-enters it with AUTO_REUSE),
    only its NAME scope -- hence the op name of the batch moment -- is uniquified to `siamese_1/...`:
        branch 1: siamese/X/bn/siamese/X/bn/moments/Squeeze/ExponentialMovingAverage
        branch 2: siamese/X/bn/siamese_1/X/bn/moments/Squeeze/ExponentialMovingAverage
        head    : fc1/bn/fc1/bn/moments/Squeeze/ExponentialMovingAverage"""
    marker = "/bn/moments/"
    if marker not in name:
        return name
    scope = name[:name.index(marker)] + "/bn/"
    if scope.startswith("siamese_1/"):
        scope = "siamese/" + scope[len("siamese_1/"):]
    return scope + name


def _lookup(ckpt: Dict[str, np.ndarray], name: str) -> Optional[np.ndarray]:
    """Variable by the reference graph's name; EMA shadows under TF's spelling (`tf_ema_key`), under the engine's own
    (files written by earlier versions of this module), or under any unique key that ends with the op-name part."""
    if name in ckpt:
        return ckpt[name]
    if "/bn/moments/" in name:
        key = tf_ema_key(name)
        if key in ckpt:
            return ckpt[key]
        hits = [k for k in ckpt if k.endswith("/" + name)]
        if len(hits) == 1:
            return ckpt[hits[0]]
    return None


def load_into_engine(engine, prefix: str, strict: bool = True) -> Dict[str, List[str]]:
    """saver.restore (train.py:250-293): parameters, BN shadows, Adam slots and the global step of a TF checkpoint of
    the tp8 graph into an Engine.  Conv kernels [1, kw, Cin, Cout] are flattened to the engine's [kw*Cin, Cout]."""
    import torch
    ckpt = read_checkpoint(prefix)
    missing: List[str] = []
    params, state = engine.get_params(), engine.get_state()
    for name in list(params):
        v = _lookup(ckpt, name)
        if v is None:
            missing.append(name)
            continue
        params[name] = np.asarray(v, np.float32).reshape(params[name].shape)
    for name in list(state):
        v = _lookup(ckpt, name)
        if v is None:
            missing.append(name)
            continue
        state[name] = np.asarray(v, np.float32).reshape(state[name].shape)
    if strict and missing:
        raise KeyError(f"{prefix}: {len(missing)} variables of the tp8 graph are missing, e.g. {missing[:3]}")
    engine.set_params(params)
    engine.set_state(state)
    m, v = engine._unflatten(engine.params_layout, engine.adam_m.cpu().numpy()), \
        engine._unflatten(engine.params_layout, engine.adam_v.cpu().numpy())
    have_adam = False
    for name in list(m):
        a, b = ckpt.get(name + "/Adam"), ckpt.get(name + "/Adam_1")
        if a is not None and b is not None:
            m[name], v[name] = np.asarray(a, np.float32).reshape(m[name].shape), np.asarray(b, np.float32).reshape(v[name].shape)
            have_adam = True
    if have_adam:
        engine.adam_m.copy_(torch.from_numpy(engine._flatten(engine.params_layout, m)))
        engine.adam_v.copy_(torch.from_numpy(engine._flatten(engine.params_layout, v)))
    if getattr(engine, "mom_accum", None) is not None:       # MomentumOptimizer's slot, `<var>/Momentum` (train.py:211-212)
        acc = engine._unflatten(engine.params_layout, engine.mom_accum.cpu().numpy())
        have_mom = False
        for name in list(acc):
            a = ckpt.get(name + "/Momentum")
            if a is not None:
                acc[name], have_mom = np.asarray(a, np.float32).reshape(acc[name].shape), True
        if have_mom:
            engine.mom_accum.copy_(torch.from_numpy(engine._flatten(engine.params_layout, acc)))
    if "Variable" in ckpt:                                   # global step (train.py:195)
        engine.step = int(ckpt["Variable"])
    used = set(params) | set(state) | {tf_ema_key(k) for k in state}
    return dict(missing=missing, unused=[k for k in ckpt if k not in used and not k.endswith(("/Adam", "/Adam_1", "/Momentum"))
                                         and k not in ("Variable", "beta1_power", "beta2_power")])


def save_from_engine(engine, prefix: str, beta1: float = 0.9, beta2: float = 0.999) -> None:
    """saver.save (train.py:316-322): the engine's variables under the reference graph's names, conv kernels in
    TF's [1, kw, Cin, Cout] shape, the optimiser's slots (Adam: `<var>/Adam`, `<var>/Adam_1`, `beta{1,2}_power`; momentum:
    `<var>/Momentum`) and the global step."""
    tensors: Dict[str, np.ndarray] = {}
    for name, a in engine.get_params().items():
        if "/conv" in name and name.endswith("/weights"):
            cin, cout = a.shape
            kw = 3 if name.endswith("conv1/weights") else 1
            a = a.reshape(1, kw, cin // kw, cout)
        tensors[name] = a.astype(np.float32)
    tensors.update({tf_ema_key(k): v.astype(np.float32) for k, v in engine.get_state().items()})
    tensors["Variable"] = np.array(engine.step, np.int32)
    if getattr(engine, "optimizer", "adam") == "momentum":   # a MomentumOptimizer graph holds one slot per variable
        acc = engine._unflatten(engine.params_layout, engine.mom_accum.cpu().numpy())
        for name in acc:
            tensors[name + "/Momentum"] = acc[name].reshape(tensors[name].shape)
    else:
        m = engine._unflatten(engine.params_layout, engine.adam_m.cpu().numpy())
        v = engine._unflatten(engine.params_layout, engine.adam_v.cpu().numpy())
        for name in m:
            shape = tensors[name].shape
            tensors[name + "/Adam"], tensors[name + "/Adam_1"] = m[name].reshape(shape), v[name].reshape(shape)
        tensors["beta1_power"] = np.array(beta1 ** max(engine.step, 0) * beta1, np.float32)
        tensors["beta2_power"] = np.array(beta2 ** max(engine.step, 0) * beta2, np.float32)
    write_checkpoint(prefix, tensors)
