"""TensorFlow-1.x checkpoint files (``tf.train.Saver``, the "tensor bundle" V2 format) without TensorFlow.

The reference saves and restores its models with ``tf.train.Saver`` (cbfssm/model/cbfssm.py:276;
cbfssm/training/trainer.py:31,59,63; cbfssm/outputs/outputs.py:41): ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``.
This module reads and writes that pair so that weights trained with the reference can be loaded into
``cbf_ssm_b200.model.CBFSSM`` (and back), on a machine that has no TensorFlow.

Format, as published in the TensorFlow sources (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*):

* ``.index`` is an SSTable in LevelDB's table format: data blocks of prefix-compressed entries
  ``varint shared | varint non_shared | varint value_len | key delta | value`` followed by the restart offsets
  (uint32 LE) and their count; every block is followed by a 5-byte trailer (compression type, masked CRC32C of block
  + type); an index block maps separator keys to data-block handles; the 48-byte footer holds the metaindex and
  index handles (varint64 offset, size; zero-padded to 40 bytes) and the magic 0xdb4775248b80fb57.
* key ``""`` -> ``BundleHeaderProto`` (num_shards = 1, endianness = LITTLE, version);
  key ``<variable name>`` -> ``BundleEntryProto`` (dtype, shape, shard_id, offset, size, masked crc32c of the bytes).
* ``.data-00000-of-00001`` holds the raw little-endian tensor bytes at those offsets.

NOT VERIFIED AGAINST A TENSORFLOW BINARY: TensorFlow cannot be installed in the build image, so the reader and writer
are tested against each other and against the format description only (tests/test_host.py).  The variable names the
reference's graph gets (``reference_variable_names``) are TensorFlow's default uniquified names for unnamed
``tf.Variable`` s created in that order; pass ``name_map`` if a checkpoint was written by a modified script.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 10: np.dtype("bool")}   # DataType enum
_DT_INV = {v: k for k, v in _DT.items()}


# ------------------------------------------------------------------------------------------------------------------
# CRC32C (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h
# ------------------------------------------------------------------------------------------------------------------
def _make_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return np.asarray(tbl, dtype=np.uint32)


_TABLE = _make_table()


def crc32c(data: bytes) -> int:
    crc = 0xFFFFFFFF
    tbl = _TABLE
    for b in data:
        crc = int(tbl[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------------------------
# varints and the two small protobuf messages (hand-encoded: no generated code, no protobuf dependency)
# ------------------------------------------------------------------------------------------------------------------
def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _get_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _pb_fields(buf: bytes) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) of one protobuf message."""
    pos = 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fn, wt, v


def _encode_shape(shape) -> bytes:       # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
    out = b""
    for s in shape:
        dim = b"\x08" + _put_varint(int(s))
        out += b"\x12" + _put_varint(len(dim)) + dim
    return out


def _decode_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for fn, wt, v in _pb_fields(buf):
        if fn == 2 and wt == 2:
            size = 0
            for f2, w2, v2 in _pb_fields(v):
                if f2 == 1 and w2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _encode_entry(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    # BundleEntryProto: dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, crc32c = 6 (fixed32)
    sh = _encode_shape(shape)
    out = b"\x08" + _put_varint(dtype) + b"\x12" + _put_varint(len(sh)) + sh
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size) + b"\x35" + struct.pack("<I", crc)
    return out


def _decode_entry(buf: bytes):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for fn, wt, v in _pb_fields(buf):
        if fn == 1: e["dtype"] = v
        elif fn == 2: e["shape"] = _decode_shape(v)
        elif fn == 3: e["shard_id"] = v
        elif fn == 4: e["offset"] = v
        elif fn == 5: e["size"] = v
        elif fn == 6: e["crc32c"] = struct.unpack("<I", v)[0]
        elif fn == 7: e["sliced"] = True
    return e


_HEADER = b"\x08\x01" + b"\x1a\x02\x08\x01"      # BundleHeaderProto: num_shards = 1, (endianness = LITTLE default), version { producer = 1 }


# ------------------------------------------------------------------------------------------------------------------
# SSTable
# ------------------------------------------------------------------------------------------------------------------
def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _read_block(buf: bytes, offset: int, size: int, verify=True) -> bytes:
    block, trailer = buf[offset:offset + size], buf[offset + size:offset + size + 5]
    if len(trailer) != 5:
        raise ValueError("truncated table block")
    if trailer[0] != 0:
        raise ValueError("compressed table blocks (type %d) are not supported; tf.train.Saver writes them uncompressed" % trailer[0])
    if verify and _mask(crc32c(block + trailer[:1])) != struct.unpack("<I", trailer[1:])[0]:
        raise ValueError("table block checksum mismatch (corrupt .index file)")
    return block


def _build_block(entries: List[Tuple[bytes, bytes]], restart_interval=16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
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


def _handle(offset: int, size: int) -> bytes:
    return _put_varint(offset) + _put_varint(size)


def read_tf_checkpoint(prefix: str, verify=True) -> Dict[str, np.ndarray]:
    """All tensors of the checkpoint ``<prefix>.index`` / ``<prefix>.data-00000-of-00001`` by variable name."""
    idx = open(prefix + ".index", "rb").read()
    if len(idx) < 48 or struct.unpack("<Q", idx[-8:])[0] != _MAGIC:
        raise ValueError("%s.index is not a TensorFlow V2 checkpoint index (bad magic)" % prefix)
    footer = idx[-48:]
    _, p = _get_varint(footer, 0)
    _, p = _get_varint(footer, p)                       # metaindex handle (unused)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    entries = []
    for _, hv in _block_entries(_read_block(idx, ioff, isize, verify)):
        off, q = _get_varint(hv, 0)
        size, _ = _get_varint(hv, q)
        entries += _block_entries(_read_block(idx, off, size, verify))
    if not entries or entries[0][0] != b"":
        raise ValueError("checkpoint index has no header entry")
    num_shards = 1
    for fn, wt, v in _pb_fields(entries[0][1]):
        if fn == 1: num_shards = v
        if fn == 2 and v != 0: raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    for key, val in entries[1:]:
        e = _decode_entry(val)
        if e["sliced"]:
            raise ValueError("partitioned variable %r: sliced entries are not supported" % key.decode())
        if e["dtype"] not in _DT:
            continue                                      # e.g. string tensors: not part of this model
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), "rb").read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = _DT[e["dtype"]]
        if len(raw) != e["size"] or e["size"] != int(np.prod(e["shape"], dtype=np.int64)) * dt.itemsize:
            raise ValueError("tensor %r: size mismatch / truncated data file" % key.decode())
        if verify and e["crc32c"] is not None and _mask(crc32c(raw)) != e["crc32c"]:
            raise ValueError("tensor %r: checksum mismatch" % key.decode())
        out[key.decode()] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    return out


def write_tf_checkpoint(prefix: str, tensors: Dict[str, np.ndarray], block_size=4096) -> None:
    """Write ``tensors`` as a single-shard V2 checkpoint (what ``tf.train.Saver().save(sess, prefix)`` produces)."""
    names = sorted(tensors, key=lambda s: s.encode())
    data, kv = bytearray(), [(b"", _HEADER)]
    for n in names:
        a = np.asarray(tensors[n])
        if a.ndim and not a.flags.c_contiguous:       # (ascontiguousarray would turn a scalar into a [1] tensor)
            a = np.ascontiguousarray(a)
        dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
        if np.dtype(dt) not in _DT_INV:
            raise ValueError("unsupported dtype %s for %r" % (a.dtype, n))
        raw = a.astype(dt, copy=False).tobytes()
        kv.append((n.encode(), _encode_entry(_DT_INV[np.dtype(dt)], a.shape, len(data), len(raw), _mask(crc32c(raw)))))
        data += raw
    out, index_entries, cur, cur_bytes = bytearray(), [], [], 0

    def flush():
        nonlocal cur, cur_bytes
        if not cur:
            return
        block = _build_block(cur)
        index_entries.append((cur[-1][0], _handle(len(out), len(block))))       # separator = last key of the block
        out.extend(block + b"\x00" + struct.pack("<I", _mask(crc32c(block + b"\x00"))))
        cur, cur_bytes = [], 0
    for k, v in kv:
        cur.append((k, v))
        cur_bytes += len(k) + len(v)
        if cur_bytes >= block_size:
            flush()
    flush()
    meta = _build_block([])
    meta_h = _handle(len(out), len(meta))
    out.extend(meta + b"\x00" + struct.pack("<I", _mask(crc32c(meta + b"\x00"))))
    iblock = _build_block(index_entries, restart_interval=1)
    index_h = _handle(len(out), len(iblock))
    out.extend(iblock + b"\x00" + struct.pack("<I", _mask(crc32c(iblock + b"\x00"))))
    footer = meta_h + index_h
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


# ------------------------------------------------------------------------------------------------------------------
# the reference graph's variable names
# ------------------------------------------------------------------------------------------------------------------
def reference_variable_names(half=False) -> Dict[str, str]:
    """Our tensor name -> the name TensorFlow gives the reference's variable.

    The reference never names its variables, so they get TensorFlow's defaults in creation order: inside
    ``GPModel.__init__`` (gp_tf.py:112-127) ``Variable``, ``Variable_1``, ``Variable_2`` and, under
    ``tf.name_scope('kern')`` (gp_tf.py:24-31), ``kern/Variable``, ``kern/Variable_1``; the second GP continues the
    counters (``Variable_3`` ..., scope ``kern_1``); then ``var_x_unc`` and ``var_y_unc`` (cbfssm.py:51-54)."""
    names, n = {}, 0
    for gi, tag in enumerate(("f",) if half else ("f", "b")):
        for field in ("zeta_pos", "zeta_mean", "zeta_var_unc"):
            names[f"{tag}.{field}"] = "Variable" if n == 0 else "Variable_%d" % n
            n += 1
        scope = "kern" if gi == 0 else "kern_%d" % gi
        names[f"{tag}.variance_unc"] = scope + "/Variable"
        names[f"{tag}.lengthscales_unc"] = scope + "/Variable_1"
    names["var_x_unc"] = "Variable_%d" % n
    names["var_y_unc"] = "Variable_%d" % (n + 1)
    return names


def import_reference_checkpoint(model, prefix: str, name_map: Optional[Dict[str, str]] = None) -> List[str]:
    """Load the 12 (CBFSSMHALF: 7) trainable tensors -- and the Adam slots ``<name>/Adam``, ``<name>/Adam_1`` and the
    step count from ``beta1_power`` if present -- of a reference checkpoint into ``model``.  Returns the names loaded."""
    import torch
    ck = read_tf_checkpoint(prefix)
    eng = model.engine
    names = name_map or reference_variable_names(getattr(model.dims, "half", False))
    loaded = []
    for ours in eng.names:
        theirs = names[ours]
        if theirs not in ck:
            raise KeyError("checkpoint has no variable %r (for %s); it holds %s" % (theirs, ours, sorted(ck)[:20]))
        view = eng.view(ours)
        val = np.asarray(ck[theirs], dtype=np.float64)
        if val.size != view.numel():
            raise ValueError("%s: checkpoint shape %s does not match %s" % (theirs, val.shape, tuple(view.shape)))
        view.copy_(torch.as_tensor(val.reshape(tuple(view.shape))))
        loaded.append(theirs)
        for slot, dst in (("/Adam", eng.adam_m), ("/Adam_1", eng.adam_v)):
            if theirs + slot in ck:
                eng.view(ours, dst).copy_(torch.as_tensor(np.asarray(ck[theirs + slot], dtype=np.float64).reshape(tuple(view.shape))))
                loaded.append(theirs + slot)
    if "beta1_power" in ck:          # beta1^t after t steps
        b1p = float(np.asarray(ck["beta1_power"]).reshape(-1)[0])
        eng.adam_t = int(round(np.log(b1p) / np.log(0.9))) - 1 if 0 < b1p < 1 else 0
        eng.adam_t = max(eng.adam_t, 0)
        loaded.append("beta1_power")
    return loaded


def export_reference_checkpoint(model, prefix: str, name_map: Optional[Dict[str, str]] = None) -> None:
    """Write ``model``'s tensors and Adam slots under the reference graph's variable names."""
    eng = model.engine
    names = name_map or reference_variable_names(getattr(model.dims, "half", False))
    out = {}
    for ours in eng.names:
        theirs = names[ours]
        shape = tuple(eng.view(ours).shape) or (1,)                   # kern variance is a [1] variable (tf_transform.py:15)
        out[theirs] = eng.view(ours).detach().cpu().numpy().reshape(shape)
        out[theirs + "/Adam"] = eng.view(ours, eng.adam_m).detach().cpu().numpy().reshape(shape)
        out[theirs + "/Adam_1"] = eng.view(ours, eng.adam_v).detach().cpu().numpy().reshape(shape)
    out["beta1_power"] = np.asarray(0.9 ** (eng.adam_t + 1), dtype=np.float64)
    out["beta2_power"] = np.asarray(0.999 ** (eng.adam_t + 1), dtype=np.float64)
    write_tf_checkpoint(prefix, out)
