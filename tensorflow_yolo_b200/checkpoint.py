"""TensorFlow checkpoint (tensor-bundle, "V2") reader and writer without TensorFlow.

Replaces ``tf.train.Saver().restore(sess, checkpoint_path)`` on the reference's TEST path
(net/yolo.py:71-78, net/base.py:55-61): ``checkpoint_path`` is the first weights source ``Yolo.test`` tries.
A checkpoint ``<prefix>`` is two kinds of files:

  <prefix>.index                    a LevelDB-format sorted string table (tensorflow/core/lib/io/table*):
                                    key "" -> BundleHeaderProto, key <variable name> -> BundleEntryProto
  <prefix>.data-0000i-of-0000N      raw little-endian tensor bytes, addressed by (shard_id, offset, size)

Table format: data blocks of prefix-compressed entries (varint32 shared, non_shared, value_len; key delta; value)
followed by a restart array and its count; every block is followed by a 5-byte trailer (compression type, masked
CRC-32C of block + type); the index block maps separator keys to block handles (varint64 offset, size); the 48-byte
footer holds the metaindex and index handles and the magic number 0xdb4775248b80fb57.
Protos (tensorflow/core/protobuf/tensor_bundle.proto): BundleHeaderProto {1: num_shards, 2: endianness, 3: version};
BundleEntryProto {1: dtype, 2: shape {2: dim {1: size}}, 3: shard_id, 4: offset, 5: size, 6: crc32c (fixed32, masked),
7: slices}.

Variables of the reference's graph are named ``yolo/conv2d_bn_act_<n>/{kernel,bias,beta,gamma,moving_mean,
moving_variance}`` (net/layers.py:25,53-63); kernels are HWIO.  ``stream_from_checkpoint`` lays them out as the
darknet float stream ``base.load_weights`` consumes (net/base.py:26-46: OIHW kernels), so both weight sources feed the
engine through the same entry point (yb_engine_load_weights).

Parity note: TensorFlow is not installable in this environment, so no file written by TensorFlow itself exists to
pin against.  The reader is checked against (a) tests/golden/bundle_tf_layout.ckpt.*, a two-shard checkpoint with Adam
slot variables produced by oracle/make_golden_bundle.py -- an independent restatement of BundleWriter + the LevelDB
table builder (shortened index separators, restart interval 16) whose protos are encoded by Google's protobuf
runtime and whose CRCs are computed bit by bit; it shares no code with this module -- (b) literal bytes of that file
(footer, index block, one data block) asserted in tests/test_checkpoint.py, (c) the CRC-32C / masking / varint
known-answer values of the LevelDB and TensorFlow test suites, (d) round trips through the writer below.
"""
import os
import struct

import numpy as np

from . import _lib

TABLE_MAGIC = 0xdb4775248b80fb57
FOOTER_LEN = 48
BLOCK_TRAILER_LEN = 5
MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_INT16, DT_INT8, DT_INT64, DT_BOOL, DT_BFLOAT16, DT_HALF = 1, 2, 3, 4, 5, 6, 9, 10, 14, 19
_NP_OF_DT = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_UINT8: np.dtype("u1"),
             DT_INT16: np.dtype("<i2"), DT_INT8: np.dtype("i1"), DT_INT64: np.dtype("<i8"), DT_BOOL: np.dtype("?"),
             DT_HALF: np.dtype("<f2")}
_DT_OF_NP = {v: k for k, v in _NP_OF_DT.items()}


class CheckpointError(Exception):
    pass


_crc_table = None


def crc32c_py(data, seed=0):
    """CRC-32C in plain Python (byte-wise table): what _crc32c falls back to when libyolo_b200.so cannot be loaded, so
    that a checkpoint can be listed / verified / converted on a machine without the CUDA library."""
    global _crc_table
    if _crc_table is None:
        table = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
            table.append(c)
        _crc_table = table
    crc = seed ^ 0xFFFFFFFF
    for b in bytes(data):
        crc = (crc >> 8) ^ _crc_table[(crc ^ b) & 0xFF]
    return crc ^ 0xFFFFFFFF


def _crc32c(data):
    try:
        return _lib.crc32c(data)
    except (ImportError, OSError, AttributeError):
        return crc32c_py(data if isinstance(data, (bytes, bytearray)) else np.ascontiguousarray(data).tobytes())


def mask_crc(crc):
    """crc32c::Mask: rotate right by 15 bits and add a constant (LevelDB/TensorFlow store CRCs masked)."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------
# varints / minimal protobuf wire format
# ---------------------------------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise CheckpointError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise CheckpointError("varint too long")


def _put_varint(v):
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _parse_message(buf):
    """{field number: [values]}: varints as int, fixed32/64 as int, length-delimited as bytes."""
    fields, pos = {}, 0
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            if len(v) != ln:
                raise CheckpointError("truncated length-delimited field")
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise CheckpointError("unsupported protobuf wire type {}".format(wt))
        fields.setdefault(num, []).append(v)
    return fields


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


class BundleEntry(object):
    __slots__ = ("dtype", "shape", "shard_id", "offset", "size", "crc32c", "has_slices")

    def __init__(self, dtype, shape, shard_id, offset, size, crc32c, has_slices=False):
        self.dtype, self.shape, self.shard_id, self.offset, self.size, self.crc32c = dtype, tuple(shape), shard_id, offset, size, crc32c
        self.has_slices = has_slices

    @staticmethod
    def parse(buf):
        f = _parse_message(buf)
        shape = []
        for sh in f.get(2, []):
            for dim in _parse_message(sh).get(2, []):
                shape.append(_signed64(_parse_message(dim).get(1, [0])[0]))
        return BundleEntry(f.get(1, [0])[0], shape, f.get(3, [0])[0], _signed64(f.get(4, [0])[0]), _signed64(f.get(5, [0])[0]),
                           f.get(6, [0])[0], has_slices=7 in f)

    def serialize(self):
        dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(s) for s in self.shape))
        out = b"\x08" + _put_varint(self.dtype) + b"\x12" + _put_varint(len(dims)) + dims
        if self.shard_id:
            out += b"\x18" + _put_varint(self.shard_id)
        if self.offset:
            out += b"\x20" + _put_varint(self.offset)
        out += b"\x28" + _put_varint(self.size)
        out += b"\x35" + struct.pack("<I", self.crc32c)
        return out


# ---------------------------------------------------------------------------------------------------------------
# table (sorted string table) reading
# ---------------------------------------------------------------------------------------------------------------
def _read_block(data, offset, size, verify):
    end = offset + size + BLOCK_TRAILER_LEN
    if offset < 0 or end > len(data):
        raise CheckpointError("block handle [{}, +{}] outside the index file".format(offset, size))
    block = data[offset:offset + size]
    ctype = data[offset + size]
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        actual = _crc32c(bytes(data[offset:offset + size + 1]))
        if unmask_crc(stored) != actual:
            raise CheckpointError("index block at {} fails its CRC-32C".format(offset))
    if ctype != 0:
        raise CheckpointError("compressed index blocks (type {}) are not supported; TensorFlow writes bundles uncompressed".format(ctype))
    return block


def _block_entries(block):
    if len(block) < 4:
        raise CheckpointError("index block too small")
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * n_restarts
    if limit < 0:
        raise CheckpointError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > limit:
            raise CheckpointError("corrupt index entry")
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def read_index(index_path, verify=True):
    """-> (header dict, {name: BundleEntry}) from <prefix>.index."""
    with open(index_path, "rb") as f:
        data = f.read()
    if len(data) < FOOTER_LEN:
        raise CheckpointError("{} is too short to be a tensor-bundle index".format(index_path))
    footer = data[-FOOTER_LEN:]
    if struct.unpack("<Q", footer[40:])[0] != TABLE_MAGIC:
        raise CheckpointError("{} is not a TensorFlow tensor-bundle index (bad magic number)".format(index_path))
    _, p = _get_varint(footer, 0)            # metaindex handle (unused: no filter policy)
    _, p = _get_varint(footer, p)
    idx_off, p = _get_varint(footer, p)
    idx_size, p = _get_varint(footer, p)
    entries, header = {}, None
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size, verify)):
        b_off, q = _get_varint(handle, 0)
        b_size, q = _get_varint(handle, q)
        for key, value in _block_entries(_read_block(data, b_off, b_size, verify)):
            if key == b"":
                h = _parse_message(value)
                header = {"num_shards": h.get(1, [0])[0], "endianness": h.get(2, [0])[0]}
            else:
                entries[key.decode("utf-8")] = BundleEntry.parse(value)
    if header is None:
        raise CheckpointError("{} has no bundle header entry".format(index_path))
    if header["endianness"] != 0:
        raise CheckpointError("big-endian bundles are not supported")
    return header, entries


class BundleReader(object):
    """Random access to the tensors of a checkpoint prefix."""

    def __init__(self, prefix, verify=True):
        self.prefix, self.verify = prefix, verify
        index_path = prefix + ".index"
        if not os.path.exists(index_path):
            raise CheckpointError("{} does not exist".format(index_path))
        self.header, self.entries = read_index(index_path, verify)
        self._shards = {}

    def names(self):
        return sorted(self.entries)

    def has_tensor(self, name):
        return name in self.entries

    def shape_dtype(self, name):
        e = self.entries[name]
        return e.shape, _NP_OF_DT.get(e.dtype)

    def _shard(self, shard_id):
        if shard_id not in self._shards:
            path = "{}.data-{:05d}-of-{:05d}".format(self.prefix, shard_id, max(self.header["num_shards"], 1))
            if not os.path.exists(path):
                raise CheckpointError("{} does not exist".format(path))
            self._shards[shard_id] = np.memmap(path, dtype=np.uint8, mode="r")
        return self._shards[shard_id]

    def get_tensor(self, name):
        if name not in self.entries:
            raise CheckpointError("Key {} not found in checkpoint".format(name))
        e = self.entries[name]
        if e.has_slices:
            raise CheckpointError("{}: partitioned (sliced) variables are not supported".format(name))
        if e.dtype not in _NP_OF_DT:
            raise CheckpointError("{}: unsupported dtype enum {}".format(name, e.dtype))
        dt = _NP_OF_DT[e.dtype]
        count = int(np.prod(e.shape, dtype=np.int64)) if e.shape else 1
        if count * dt.itemsize != e.size:
            raise CheckpointError("{}: entry size {} does not match shape {} of {}".format(name, e.size, e.shape, dt))
        shard = self._shard(e.shard_id)
        if e.offset < 0 or e.offset + e.size > shard.size:
            raise CheckpointError("{}: bytes [{}, +{}] outside its data shard".format(name, e.offset, e.size))
        raw = np.array(shard[e.offset:e.offset + e.size])        # copy out of the mapping
        if self.verify and unmask_crc(e.crc32c) != _crc32c(raw):
            raise CheckpointError("{}: tensor bytes fail their CRC-32C".format(name))
        return raw.view(dt).reshape(e.shape)


# ---------------------------------------------------------------------------------------------------------------
# writer (fixtures, and exporting weights in the reference's checkpoint layout)
# ---------------------------------------------------------------------------------------------------------------
class _BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last_key, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            n = min(len(key), len(self.last_key))
            while shared < n and key[shared] == self.last_key[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.count += 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4


def _write_block(out, contents):
    offset = len(out)
    trailer_crc = mask_crc(_crc32c(contents + b"\x00"))
    out += contents + b"\x00" + struct.pack("<I", trailer_crc)
    return offset, len(contents)


def write_bundle(prefix, tensors, block_size=262144):
    """Writes {name: numpy array} as <prefix>.index + <prefix>.data-00000-of-00001 (one shard, uncompressed,
    keys sorted, every tensor and block with its masked CRC-32C), the layout tf.train.Saver produces."""
    names = sorted(tensors)
    entries = []
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        offset = 0
        for name in names:
            arr = np.asarray(tensors[name])                  # (ascontiguousarray would turn a scalar into shape (1,))
            dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
            if np.dtype(dt) not in _DT_OF_NP:
                raise CheckpointError("{}: dtype {} cannot be stored".format(name, arr.dtype))
            raw = arr.astype(dt, copy=False).tobytes(order="C")
            f.write(raw)
            entries.append((name.encode("utf-8"), BundleEntry(_DT_OF_NP[np.dtype(dt)], arr.shape, 0, offset, len(raw),
                                                              mask_crc(_crc32c(raw)))))
            offset += len(raw)
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"            # num_shards = 1, version { producer: 1 }
    out = bytearray()
    index = _BlockBuilder(restart_interval=1)
    block = _BlockBuilder()
    pending = [(b"", header)] + [(k, e.serialize()) for k, e in entries]
    for i, (key, value) in enumerate(pending):
        block.add(key, value)
        if block.size_estimate() >= block_size or i == len(pending) - 1:
            off, size = _write_block(out, block.finish())
            index.add(key, _put_varint(off) + _put_varint(size))        # separator: the block's last key
            block = _BlockBuilder()
    meta_off, meta_size = _write_block(out, _BlockBuilder().finish())
    idx_off, idx_size = _write_block(out, index.finish())
    footer = _put_varint(meta_off) + _put_varint(meta_size) + _put_varint(idx_off) + _put_varint(idx_size)
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))
    return names


# ---------------------------------------------------------------------------------------------------------------
# the reference's variables <-> the darknet float stream
# ---------------------------------------------------------------------------------------------------------------
def _expected_shape(name, spec, cin):
    leaf = name.rsplit("/", 1)[-1]
    return (spec.ksize, spec.ksize, cin, spec.filters) if leaf == "kernel" else (spec.filters,)


def stream_from_checkpoint(layers, prefix, verify=True):
    """Reads every variable the layer list names (layer.variable_names, net/layers.py:53-63) from the checkpoint and
    returns the float32 stream base.load_weights would have consumed from a darknet .weights file: per conv
    beta, gamma, moving_mean, moving_variance (or bias), then the kernel as [O][I][kh][kw] (net/base.py:36-40).
    Raises CheckpointError when a variable is missing or has another shape/dtype -- tf.train.Saver.restore fails in
    the same situations."""
    state = layers[0]._yb_state
    specs = state.graph.specs
    reader = BundleReader(prefix, verify=verify)
    chunks = []
    for i, layer in enumerate(layers):
        names = getattr(layer, "variable_names", [])
        if not names:
            continue
        spec = specs[i]
        cin = specs[spec.src[0]].shape[2]
        for name in names:
            if not reader.has_tensor(name):
                raise CheckpointError("Key {} not found in checkpoint".format(name))
            value = reader.get_tensor(name)
            want = _expected_shape(name, spec, cin)
            if value.dtype != np.float32 or tuple(value.shape) != want:
                raise CheckpointError("{}: checkpoint holds {} {}, the graph needs float32 {}".format(
                    name, value.dtype, tuple(value.shape), want))
            if name.endswith("/kernel"):
                value = np.transpose(value, (3, 2, 0, 1))            # HWIO -> OIHW, the inverse of net/base.py:40
            chunks.append(np.ascontiguousarray(value, dtype=np.float32).reshape(-1))
    return np.concatenate(chunks) if chunks else np.zeros(0, np.float32)


def checkpoint_from_stream(layers, stream, prefix, extra=None):
    """The inverse: writes a checkpoint holding the network's variables (as tf.train.Saver would name and lay them
    out) from a darknet-order float stream.  extra: additional {name: array} entries (e.g. a global_step)."""
    state = layers[0]._yb_state
    specs = state.graph.specs
    stream = np.asarray(stream, dtype=np.float32)
    tensors, read = {}, 0
    for i, layer in enumerate(layers):
        names = getattr(layer, "variable_names", [])
        if not names:
            continue
        spec = specs[i]
        cin = specs[spec.src[0]].shape[2]
        for name in names:
            shape = _expected_shape(name, spec, cin)
            size = int(np.prod(shape))
            if read + size > stream.size:
                raise ValueError("weight stream too short for {}".format(name))
            chunk = stream[read:read + size]
            read += size
            if name.endswith("/kernel"):
                chunk = np.transpose(chunk.reshape(spec.filters, cin, spec.ksize, spec.ksize), (2, 3, 1, 0))
            tensors[name] = np.ascontiguousarray(chunk.reshape(shape))
    if extra:
        tensors.update(extra)
    write_bundle(prefix, tensors)
    return read
