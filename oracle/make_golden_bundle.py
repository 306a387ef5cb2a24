"""Generates tests/golden/bundle_* -- TensorFlow tensor-bundle ("V2") checkpoint fixtures that are NOT written by
tensorflow_yolo_b200/checkpoint.py (test infrastructure only; the product never imports oracle/).

    python oracle/make_golden_bundle.py            # rewrites tests/golden/bundle_tf_layout.* and bundle_expected.npz

TensorFlow cannot be installed here, so the fixture is produced by an independent restatement of what
tf.train.Saver writes on the reference's path (net/yolo.py:71, net/v2.py:201-205), sharing no code with the reader
under test:

  * the protocol buffers are encoded by Google's protobuf runtime from message descriptors that restate
    tensorflow/core/protobuf/tensor_bundle.proto, framework/tensor_shape.proto, framework/versions.proto and
    framework/tensor_slice.proto field by field (numbers, types, the int64/fixed32/enum wire types);
  * the .index file is a LevelDB-format table built the way tensorflow/core/lib/io/table_builder.cc does it:
    prefix-compressed entries with a restart point every 16 keys, data blocks cut at block_size, an index block with
    restart interval 1 whose keys are SHORTENED separators (BytewiseComparator::FindShortestSeparator between
    blocks, FindShortSuccessor after the last), an empty metaindex block, 5-byte block trailers (type 0 + masked
    CRC-32C over contents and type) and the 48-byte footer with the magic 0xdb4775248b80fb57;
  * CRC-32C is computed bit by bit from the Castagnoli polynomial (no tables, no SSE4.2), cross-checked against the
    RFC 3720 B.4 vectors before anything is written;
  * the data files follow BundleWriter / MergeBundles: two shards (.data-00000-of-00002, .data-00001-of-00002), tensors
    appended back to back in key order inside their shard, entries carrying shard_id, offset, size and the masked CRC
    of the tensor bytes.

Content: the variables of a three-conv network with the reference's names (yolo/conv2d_bn_act_<i>/..., kernels HWIO,
net/layers.py:53-63) plus what a TRAIN checkpoint adds and the TEST path must ignore: Adam slot variables
(<var>/Adam, <var>/Adam_1), beta1_power, beta2_power (net/v2.py:205) and an int64 global_step.
"""
import os
import struct
import sys

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PREFIX = os.path.join(GOLDEN, "bundle_tf_layout.ckpt")


# ---- CRC-32C, bitwise (reflected polynomial 0x82F63B78), and LevelDB's mask ----
def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
    return crc ^ 0xFFFFFFFF


def masked(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---- the four .proto files, restated as descriptors ----
def build_messages():
    T = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="tf_bundle_restated.proto", package="tensorflow", syntax="proto3")

    def field(msg, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None):
        f = msg.field.add(name=name, number=number, type=ftype, label=label)
        if type_name:
            f.type_name = type_name
        return f

    shape = fd.message_type.add(name="TensorShapeProto")                 # tensor_shape.proto
    dim = shape.nested_type.add(name="Dim")
    field(dim, "size", 1, T.TYPE_INT64)
    field(dim, "name", 2, T.TYPE_STRING)
    field(shape, "dim", 2, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow.TensorShapeProto.Dim")
    field(shape, "unknown_rank", 3, T.TYPE_BOOL)
    ver = fd.message_type.add(name="VersionDef")                         # versions.proto
    field(ver, "producer", 1, T.TYPE_INT32)
    field(ver, "min_consumer", 2, T.TYPE_INT32)
    field(ver, "bad_consumers", 3, T.TYPE_INT32, T.LABEL_REPEATED)
    sl = fd.message_type.add(name="TensorSliceProto")                    # tensor_slice.proto
    ext = sl.nested_type.add(name="Extent")
    field(ext, "start", 1, T.TYPE_INT64)
    field(ext, "length", 2, T.TYPE_INT64)
    field(sl, "extent", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow.TensorSliceProto.Extent")
    hdr = fd.message_type.add(name="BundleHeaderProto")                  # tensor_bundle.proto
    field(hdr, "num_shards", 1, T.TYPE_INT32)
    field(hdr, "endianness", 2, T.TYPE_INT32)                            # enum Endianness { LITTLE = 0; BIG = 1; }: varint
    field(hdr, "version", 3, T.TYPE_MESSAGE, type_name=".tensorflow.VersionDef")
    ent = fd.message_type.add(name="BundleEntryProto")
    field(ent, "dtype", 1, T.TYPE_INT32)                                 # enum DataType (types.proto): varint
    field(ent, "shape", 2, T.TYPE_MESSAGE, type_name=".tensorflow.TensorShapeProto")
    field(ent, "shard_id", 3, T.TYPE_INT32)
    field(ent, "offset", 4, T.TYPE_INT64)
    field(ent, "size", 5, T.TYPE_INT64)
    field(ent, "crc32c", 6, T.TYPE_FIXED32)
    field(ent, "slices", 7, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow.TensorSliceProto")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow." + n))
    return get("BundleHeaderProto"), get("BundleEntryProto")


DT = {np.dtype("float32"): 1, np.dtype("int64"): 9, np.dtype("int32"): 3}      # types.proto: DT_FLOAT, DT_INT64, DT_INT32


# ---- LevelDB table builder (table_builder.cc, block_builder.cc, format.cc, comparator.cc) ----
def varint(v):
    out = bytearray()
    while v >= 128:
        out.append((v & 127) | 128)
        v >>= 7
    out.append(v)
    return bytes(out)


class BlockBuilder(object):
    def __init__(self, restart_interval):
        self.interval, self.buf, self.restarts, self.counter, self.last_key = restart_interval, bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            limit = min(len(self.last_key), len(key))
            while shared < limit and self.last_key[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.counter += 1

    def estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return len(self.buf) == 0

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def shortest_separator(start, limit):
    """BytewiseComparatorImpl::FindShortestSeparator"""
    n = min(len(start), len(limit))
    d = 0
    while d < n and start[d] == limit[d]:
        d += 1
    if d >= n:
        return start                       # one is a prefix of the other: not shortened
    b = start[d]
    if b < 0xff and b + 1 < limit[d]:
        return start[:d] + bytes([b + 1])
    return start


def short_successor(key):
    """BytewiseComparatorImpl::FindShortSuccessor"""
    for i, b in enumerate(key):
        if b != 0xff:
            return key[:i] + bytes([b + 1])
    return key


def build_table(items, block_size):
    out = bytearray()

    def write_block(contents):
        handle = varint(len(out)) + varint(len(contents))
        out.extend(contents + b"\x00" + struct.pack("<I", masked(crc32c(contents + b"\x00"))))
        return handle

    data, index = BlockBuilder(16), BlockBuilder(1)
    pending, last_key = None, b""
    for key, value in items:
        assert key > last_key or (key == b"" and last_key == b""), "keys must be added in increasing order"
        if pending is not None:
            index.add(shortest_separator(last_key, key), pending)
            pending = None
        data.add(key, value)
        last_key = key
        if data.estimate() >= block_size:
            pending = write_block(data.finish())
            data = BlockBuilder(16)
    if not data.empty():
        pending = write_block(data.finish())
    if pending is not None:
        index.add(short_successor(last_key), pending)
    meta = write_block(BlockBuilder(16).finish())          # no filter policy: empty metaindex block
    idx = write_block(index.finish())
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<II", 0x8b80fb57, 0xdb477524))
    return bytes(out)


def fixture_tensors():
    """name -> (array, shard).  A three-conv network: conv 3x3 3->8 (BN), conv 3x3 s2 8->16 (BN), conv 1x1 16->4 (bias)."""
    rs = np.random.RandomState(7)
    t = {}
    for i, (cin, cout, k, bn) in enumerate([(3, 8, 3, True), (8, 16, 3, True), (16, 4, 1, False)]):
        stem = "yolo/conv2d_bn_act_%d/" % i
        t[stem + "kernel"] = (rs.standard_normal((k, k, cin, cout)).astype(np.float32), i % 2)
        if bn:
            t[stem + "beta"] = (rs.standard_normal(cout).astype(np.float32), 0)
            t[stem + "gamma"] = (rs.uniform(0.5, 1.5, cout).astype(np.float32), 0)
            t[stem + "moving_mean"] = (rs.standard_normal(cout).astype(np.float32), 1)
            t[stem + "moving_variance"] = (rs.uniform(0.5, 2.0, cout).astype(np.float32), 0)
        else:
            t[stem + "bias"] = (rs.standard_normal(cout).astype(np.float32), 0)
        # what AdamOptimizer adds per trainable variable (TRAIN checkpoints, net/v2.py:205): slots the TEST path skips
        for name in ([stem + "kernel"] + ([stem + "beta", stem + "gamma"] if bn else [stem + "bias"])):
            for slot in ("/Adam", "/Adam_1"):
                t[name + slot] = (rs.standard_normal(t[name][0].shape).astype(np.float32) * 1e-3, 1)
    t["beta1_power"] = (np.asarray(0.9 ** 40, dtype=np.float32), 1)
    t["beta2_power"] = (np.asarray(0.999 ** 40, dtype=np.float32), 1)
    t["global_step"] = (np.asarray(40, dtype=np.int64), 0)
    return t


def main():
    assert crc32c(b"123456789") == 0xe3069283 and crc32c(bytes(32)) == 0x8a9136aa and crc32c(bytes([0xff] * 32)) == 0x62a8ab43
    assert crc32c(bytes(range(32))) == 0x46dd794e                            # RFC 3720 B.4
    Header, Entry = build_messages()
    tensors = fixture_tensors()
    n_shards = 2
    shards = [bytearray() for _ in range(n_shards)]
    items = []
    header = Header(num_shards=n_shards)
    header.version.producer = 1
    items.append((b"", header.SerializeToString()))
    for name in sorted(tensors):
        arr, shard = tensors[name]
        raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes(order="C")
        e = Entry(dtype=DT[arr.dtype], shard_id=shard, offset=len(shards[shard]), size=len(raw), crc32c=masked(crc32c(raw)))
        for d in arr.shape:
            e.shape.dim.add(size=d)
        shards[shard] += raw
        items.append((name.encode("utf-8"), e.SerializeToString()))
    os.makedirs(GOLDEN, exist_ok=True)
    with open(PREFIX + ".index", "wb") as f:
        f.write(build_table(items, block_size=400))         # small blocks: several data blocks, shortened index keys
    for i, blob in enumerate(shards):
        with open("%s.data-%05d-of-%05d" % (PREFIX, i, n_shards), "wb") as f:
            f.write(bytes(blob))
    np.savez(os.path.join(GOLDEN, "bundle_expected.npz"), **{k.replace("/", "|"): v[0] for k, v in tensors.items()})
    print("wrote", PREFIX + ".index", os.path.getsize(PREFIX + ".index"), "bytes;", [len(b) for b in shards], "data bytes;", len(items) - 1, "tensors")


if __name__ == "__main__":
    sys.exit(main())
