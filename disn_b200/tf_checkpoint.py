"""Pure-Python reader (and minimal writer) for TensorFlow "tensor bundle" checkpoints
(``model.ckpt.index`` + ``model.ckpt.data-00000-of-00001``), so the reference's trained variables
(``tf.train.Saver`` files, test/create_sdf.py:180-192, README.md:31-37) can be loaded by name without TensorFlow:

    from disn_b200.tf_checkpoint import load_checkpoint
    weights = load_checkpoint("checkpoint/SDF_DISN/model.ckpt", prefixes=("vgg_16/", "sdfprediction"))
    engine.load_weights(weights)

Format (restated from TensorFlow's tensor_bundle / leveldb-table sources, which are not vendored in the
reference; UNPINNED against a real checkpoint -- the authors' files are Dropbox downloads -- and therefore
covered by a writer/reader round trip only):
  * .index is an SSTable: data blocks of prefix-compressed (key, value) entries + restart array, a 5-byte block
    trailer (compression type, masked crc32c), an index block mapping last-keys to block handles, and a 48-byte
    footer (metaindex handle, index handle, padding, magic 0xdb4775248b80fb57);
  * key "" holds a BundleHeaderProto, every other key a BundleEntryProto
    {1: dtype, 2: shape{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size, 6: crc32c};
  * .data-* files hold the raw little-endian tensor bytes.
"""
from __future__ import annotations

import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 6: np.int8, 9: np.int64, 10: np.bool_}
DTYPE_CODES = {np.dtype(v): k for k, v in DTYPES.items()}


def _varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf):
    """-> {field: [values]} with varints as int, length-delimited as bytes, fixed32/64 as int."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _block_entries(block):
    """iterate (key, value) of one table block (contents without the 5-byte trailer)"""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _read_block(data, offset, size):
    ctype = data[offset + size]
    if ctype != 0:
        raise NotImplementedError("compressed table blocks (type %d) are not supported; re-save the checkpoint "
                                  "uncompressed" % ctype)
    return data[offset:offset + size]


def read_index(index_path):
    """-> {tensor_name: dict(dtype, shape, shard_id, offset, size)}"""
    data = open(index_path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad table magic)" % index_path)
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex offset
    _, pos = _varint(footer, pos)          # metaindex size
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, value in _block_entries(_read_block(data, boff, bsize)):
            if key == b"":
                continue                    # BundleHeaderProto
            e = _parse_proto(value)
            shape = []
            if 2 in e:
                for dim in _parse_proto(e[2][0]).get(2, []):
                    shape.append(_parse_proto(dim).get(1, [0])[0])
            entries[key.decode()] = dict(dtype=e.get(1, [0])[0], shape=tuple(shape), shard_id=e.get(3, [0])[0],
                                         offset=e.get(4, [0])[0], size=e.get(5, [0])[0])
    return entries


def load_checkpoint(prefix, prefixes=None):
    """Read every variable (optionally only names starting with one of `prefixes`) as numpy arrays."""
    index = read_index(prefix + ".index")
    shards = {}
    out = {}
    num_shards = max(e["shard_id"] for e in index.values()) + 1 if index else 1
    for name, e in index.items():
        if prefixes and not name.startswith(tuple(prefixes)):
            continue
        if e["dtype"] not in DTYPES:
            continue                        # non-numeric bookkeeping entries
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        raw = np.asarray(shards[sid][e["offset"]:e["offset"] + e["size"]])
        out[name] = raw.view(DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


# ---------------------------------------------------------------------------------------------------------
# minimal writer (single data block per 4 KB, no compression) -- used by the tests and for exporting synthetic weights
# ---------------------------------------------------------------------------------------------------------
def _crc32c_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TBL = _crc32c_table()


def _masked_crc32c(data):
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TBL[(c ^ b) & 0xFF] ^ (c >> 8)
    c ^= 0xFFFFFFFF
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _entry_proto(dtype, shape, offset, size):
    dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(s)) for s in shape))
    return (_field(1, 0, _put_varint(dtype)) + _field(2, 2, _put_varint(len(dims)) + dims) +
            _field(4, 0, _put_varint(offset)) + _field(5, 0, _put_varint(size)))


def _build_block_multi(items):
    body, restarts = bytearray(), []
    for k, v in items:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(k)) + _put_varint(len(v)) + k + v
    return bytes(body) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def save_checkpoint(prefix, tensors):
    """Write {name: array} as <prefix>.index + <prefix>.data-00000-of-00001 (uncompressed, one shard)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = [(b"", _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1))))]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors):
            a = np.ascontiguousarray(tensors[name])
            raw = a.tobytes()
            f.write(raw)
            items.append((name.encode(), _entry_proto(DTYPE_CODES[a.dtype], a.shape, offset, len(raw))))
            offset += len(raw)
    out = bytearray()
    index_items = []
    chunk, chunk_bytes = [], 0
    def flush():
        nonlocal chunk, chunk_bytes
        if not chunk:
            return
        blk = _build_block_multi(chunk)
        index_items.append((chunk[-1][0], _put_varint(len(out)) + _put_varint(len(blk))))
        out.extend(blk + b"\x00" + struct.pack("<I", _masked_crc32c(blk + b"\x00")))
        chunk, chunk_bytes = [], 0
    for kv in items:
        chunk.append(kv)
        chunk_bytes += len(kv[0]) + len(kv[1])
        if chunk_bytes > 4096:
            flush()
    flush()
    meta = _build_block_multi([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta))
    out.extend(meta + b"\x00" + struct.pack("<I", _masked_crc32c(meta + b"\x00")))
    idx = _build_block_multi(index_items)
    idx_handle = _put_varint(len(out)) + _put_varint(len(idx))
    out.extend(idx + b"\x00" + struct.pack("<I", _masked_crc32c(idx + b"\x00")))
    footer = meta_handle + idx_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
