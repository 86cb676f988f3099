"""Test infrastructure: an INDEPENDENT assembler of a TensorFlow tensor-bundle checkpoint, written from the documented
on-disk layout (leveldb table format: table_format.md; tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc};
tensorflow/core/protobuf/tensor_bundle.proto) -- it shares no code with disn_b200/tf_checkpoint.py, so that a symmetric
misunderstanding in that module's reader + writer cannot hide.  Everything a tf.train.Saver-written .index contains and
the module's own writer does NOT produce is exercised:
  * keys prefix-compressed against the previous key (shared > 0), restart points every 16 entries;
  * several data blocks (block_size 4096 as in TF's table builder), index keys shortened like leveldb's
    FindShortestSeparator (so they are NOT keys of the table);
  * BundleHeaderProto {num_shards, endianness = LITTLE, version{producer}}; BundleEntryProto with shard_id and the masked
    crc32c of the tensor bytes; optimizer-slot entries and an int64 global_step next to the model variables;
  * crc32c as specified by RFC 3720 B.4, implemented bitwise (no table), masked as leveldb does.
"""
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
DT = {np.dtype(np.float32): 1, np.dtype(np.int64): 9, np.dtype(np.int32): 3}


def crc32c_bitwise(data: bytes) -> int:
    crc = 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def varint(v: int) -> bytes:
    b = bytearray()
    while v >= 0x80:
        b.append((v & 0x7F) | 0x80)
        v >>= 7
    b.append(v)
    return bytes(b)


def pb_varint(field, v):
    return varint(field << 3) + varint(v)


def pb_bytes(field, payload):
    return varint((field << 3) | 2) + varint(len(payload)) + payload


def pb_fixed32(field, v):
    return varint((field << 3) | 5) + struct.pack("<I", v)


def entry_proto(arr, shard_id, offset):
    raw = arr.tobytes()
    shape = b"".join(pb_bytes(2, pb_varint(1, int(d))) for d in arr.shape)
    msg = pb_varint(1, DT[arr.dtype]) + pb_bytes(2, shape)
    if shard_id:
        msg += pb_varint(3, shard_id)
    if offset:
        msg += pb_varint(4, offset)
    msg += pb_varint(5, len(raw)) + pb_fixed32(6, mask(crc32c_bitwise(raw)))
    return msg


class BlockBuilder:
    def __init__(self, restart_interval):
        self.ri, self.buf, self.restarts, self.n, self.last = restart_interval, bytearray(), [0], 0, b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.n % self.ri == 0 and self.n:
            self.restarts.append(len(self.buf))
        elif self.n:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last, self.n = key, self.n + 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def shortest_separator(a: bytes, b: bytes) -> bytes:
    """leveldb BytewiseComparator::FindShortestSeparator: a <= result < b, as short as possible."""
    n = 0
    while n < min(len(a), len(b)) and a[n] == b[n]:
        n += 1
    if n < min(len(a), len(b)) and a[n] < 0xFF and a[n] + 1 < b[n]:
        return a[:n] + bytes([a[n] + 1])
    return a


def build_bundle(tensors: dict, num_shards=1, block_size=4096, compression_type=0, shard_of=lambda name: 0):
    """-> (index_bytes, {shard_id: data_bytes}).  tensors: name -> ndarray (written in sorted key order, like TF)."""
    shard_data = {i: bytearray() for i in range(num_shards)}
    kv = [(b"", pb_varint(1, num_shards) + pb_varint(2, 0) + pb_bytes(3, pb_varint(1, 1)))]
    for name in sorted(tensors):
        a = np.ascontiguousarray(tensors[name])
        sid = shard_of(name)
        off = len(shard_data[sid])
        shard_data[sid] += a.tobytes()
        kv.append((name.encode(), entry_proto(a, sid, off)))
    out = bytearray()

    def write_block(contents: bytes) -> bytes:
        handle = varint(len(out)) + varint(len(contents))
        trailer_type = bytes([compression_type])
        out.extend(contents + trailer_type + struct.pack("<I", mask(crc32c_bitwise(contents + trailer_type))))
        return handle

    index = BlockBuilder(1)
    blk = BlockBuilder(16)
    pending = None          # (last key of the finished block, its handle)
    for key, value in kv:
        if pending is not None:
            index.add(shortest_separator(pending[0], key), pending[1])
            pending = None
        blk.add(key, value)
        if blk.size() >= block_size:
            pending = (key, write_block(blk.finish()))
            blk = BlockBuilder(16)
    if blk.n:
        pending = (blk.last, write_block(blk.finish()))
    if pending is not None:
        index.add(pending[0] + b"\x00" if False else pending[0], pending[1])      # FindShortSuccessor is optional
    meta_handle = write_block(BlockBuilder(1).finish())
    index_handle = write_block(index.finish())
    footer = meta_handle + index_handle
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))
    return bytes(out), {i: bytes(d) for i, d in shard_data.items()}


def write_bundle(prefix: str, tensors: dict, **kw):
    idx, shards = build_bundle(tensors, **kw)
    with open(prefix + ".index", "wb") as f:
        f.write(idx)
    n = len(shards)
    for i, d in shards.items():
        with open("%s.data-%05d-of-%05d" % (prefix, i, n), "wb") as f:
            f.write(d)
    return idx
