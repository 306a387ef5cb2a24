"""TensorFlow tensor-bundle checkpoints without TensorFlow: known-answer values of the pieces the
format is made of, writer/reader round trips, corruption detection, and the reference's restore-or-fall-back
behaviour (net/yolo.py:71-78, net/base.py:55-61)."""
import os
import struct

import numpy as np
import pytest

import helpers
from tensorflow_yolo_b200 import _lib, checkpoint as ckpt, plan as P, synth
from tensorflow_yolo_b200.net import base as pbase


def test_crc32c_known_answers(lib_built):
    # RFC 3720 B.4 / LevelDB crc32c_test.cc / TensorFlow crc32c_test.cc
    assert _lib.crc32c(b"123456789") == 0xe3069283
    assert _lib.crc32c(bytes(32)) == 0x8a9136aa
    assert _lib.crc32c(bytes([0xff] * 32)) == 0x62a8ab43
    assert _lib.crc32c(bytes(range(32))) == 0x46dd794e
    assert _lib.crc32c(bytes(range(31, -1, -1))) == 0x113fdb5c
    assert _lib.crc32c(b"") == 0
    data = np.random.RandomState(0).bytes(100003)
    assert _lib.crc32c(data, impl=0) == _lib.crc32c(data, impl=1)                       # SSE4.2 vs table
    assert _lib.crc32c(data[777:], seed=_lib.crc32c(data[:777])) == _lib.crc32c(data)   # Extend
    # masking (crc32c.h): Mask(crc) = rotr(crc, 15) + 0xa282ead8; Unmask inverts it
    crc = _lib.crc32c(b"foo")
    assert ckpt.mask_crc(crc) != crc and ckpt.unmask_crc(ckpt.mask_crc(crc)) == crc
    assert ckpt.mask_crc(0) == 0xa282ead8
    assert ckpt.unmask_crc(ckpt.unmask_crc(ckpt.mask_crc(ckpt.mask_crc(crc)))) == crc


def test_varint_and_entry_proto_wire_format():
    assert ckpt._put_varint(0) == b"\x00" and ckpt._put_varint(300) == b"\xac\x02"      # protobuf docs example
    for v in (0, 1, 127, 128, 16383, 16384, 2 ** 31 - 1, 2 ** 40 + 3):
        assert ckpt._get_varint(ckpt._put_varint(v), 0) == (v, len(ckpt._put_varint(v)))
    e = ckpt.BundleEntry(ckpt.DT_FLOAT, (3, 3, 32, 64), 0, 1234567, 73728, 0xdeadbeef)
    raw = e.serialize()
    # dtype=1 | shape{dim{3} dim{3} dim{32} dim{64}} | offset | size | crc32c(fixed32)
    assert raw == (b"\x08\x01" + b"\x12\x10" + b"\x12\x02\x08\x03" * 2 + b"\x12\x02\x08\x20" + b"\x12\x02\x08\x40" +
                   b"\x20\x87\xad\x4b" + b"\x28\x80\xc0\x04" + b"\x35" + struct.pack("<I", 0xdeadbeef))
    back = ckpt.BundleEntry.parse(raw)
    assert (back.dtype, back.shape, back.shard_id, back.offset, back.size, back.crc32c) == (1, (3, 3, 32, 64), 0, 1234567, 73728, 0xdeadbeef)


def test_bundle_round_trip_many_keys_and_blocks(tmp_path, lib_built):
    rng = np.random.RandomState(3)
    tensors = {"yolo/conv2d_bn_act_%d/kernel" % i: rng.standard_normal((3, 3, 4, 8)).astype(np.float32) for i in range(70)}
    tensors.update({"yolo/conv2d_bn_act_%d/beta" % i: rng.standard_normal(8).astype(np.float32) for i in range(70)})
    tensors["global_step"] = np.asarray(12345, dtype=np.int64)
    tensors["train/beta1_power"] = np.asarray(0.9, dtype=np.float32)
    prefix = str(tmp_path / "model.ckpt")
    ckpt.write_bundle(prefix, tensors, block_size=512)             # forces dozens of data blocks + restarts
    with open(prefix + ".index", "rb") as f:
        idx = f.read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xdb4775248b80fb57   # table magic (table/format.h)
    r = ckpt.BundleReader(prefix)
    assert r.header == {"num_shards": 1, "endianness": 0}
    assert r.names() == sorted(tensors)
    for k, v in tensors.items():
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    with pytest.raises(ckpt.CheckpointError, match="not found"):
        r.get_tensor("yolo/missing")
    # the default (single 256 KB block) layout reads back too
    ckpt.write_bundle(prefix, tensors)
    assert np.array_equal(ckpt.BundleReader(prefix).get_tensor("global_step"), tensors["global_step"])


def test_corruption_is_detected(tmp_path, lib_built):
    prefix = str(tmp_path / "m")
    ckpt.write_bundle(prefix, {"a/kernel": np.arange(64, dtype=np.float32).reshape(1, 1, 8, 8), "a/bias": np.ones(8, np.float32)})
    data_path = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data_path, "rb").read())
    raw[40] ^= 0x01
    open(data_path, "wb").write(bytes(raw))
    r = ckpt.BundleReader(prefix)
    with pytest.raises(ckpt.CheckpointError, match="CRC"):
        [r.get_tensor(n) for n in r.names()]
    assert ckpt.BundleReader(prefix, verify=False).get_tensor("a/bias").shape == (8,)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 0x40
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ckpt.CheckpointError):
        ckpt.BundleReader(prefix)
    open(prefix + ".index", "wb").write(b"not a table")
    with pytest.raises(ckpt.CheckpointError):
        ckpt.BundleReader(prefix)


def test_stream_round_trip_in_reference_variable_layout(tmp_path, lib_built):
    shape = (64, 96, 3)
    for build, nc in ((helpers.build_v3, 80), (helpers.build_v2, 20)):
        net, _, stream = build(shape, nc)
        prefix = str(tmp_path / ("net%d.ckpt" % nc))
        used = ckpt.checkpoint_from_stream(net, stream, prefix, extra={"global_step": np.asarray(7, np.int64)})
        assert used == stream.size == P.weight_count(net[0]._yb_state.graph.specs)
        r = ckpt.BundleReader(prefix)
        # names and shapes are the reference graph's: yolo/conv2d_bn_act_N/{beta,gamma,moving_mean,moving_variance|bias,kernel}, HWIO
        first = [l for l in net if getattr(l, "variable_names", None)][0]
        assert first.variable_names[-1] == "yolo/conv2d_bn_act_0/kernel"
        assert r.shape_dtype("yolo/conv2d_bn_act_0/kernel") == ((3, 3, 3, 32), np.dtype("<f4"))
        assert r.shape_dtype("yolo/conv2d_bn_act_0/moving_variance") == ((32,), np.dtype("<f4"))
        # kernel element [kh,kw,i,o] of the checkpoint is element [o,i,kh,kw] of the darknet stream (net/base.py:36-40)
        k = r.get_tensor("yolo/conv2d_bn_act_0/kernel")
        dark = stream[4 * 32:4 * 32 + 32 * 3 * 3 * 3].reshape(32, 3, 3, 3)
        assert k[1, 2, 0, 5] == dark[5, 0, 1, 2]
        back = ckpt.stream_from_checkpoint(net, prefix)
        assert back.dtype == np.float32 and np.array_equal(back, stream)


def test_restore_or_fall_back_like_the_reference(tmp_path, capsys, lib_built):
    shape = (64, 64, 3)
    net, _, stream = helpers.build_v3(shape, 80)
    sess = pbase.Session(net)
    # missing checkpoint -> reported, False (the caller then loads the .weights file, net/yolo.py:72-76)
    assert pbase.load_checkpoint_by_path(pbase.Saver(), sess, str(tmp_path / "nope")) is False
    assert "Failed to load" in capsys.readouterr().out
    # a checkpoint of another network (v2 variables only cover part of the names) -> False as well
    net2, _, stream2 = helpers.build_v2(shape, 20)
    ckpt.checkpoint_from_stream(net2, stream2, str(tmp_path / "v2.ckpt"))
    assert pbase.load_checkpoint_by_path(pbase.Saver(), sess, str(tmp_path / "v2.ckpt")) is False
    assert "Failed to load" in capsys.readouterr().out
    # the right checkpoint -> True, and the engine-side stream is exactly the darknet stream
    ckpt.checkpoint_from_stream(net, stream, str(tmp_path / "v3.ckpt"))
    assert pbase.load_checkpoint_by_path(pbase.Saver(), sess, str(tmp_path / "v3.ckpt")) is True
    assert np.array_equal(net[0]._yb_state.pending_stream, stream)
    # Saver.save writes what restore reads
    pbase.Saver().save(sess, str(tmp_path / "again.ckpt"))
    assert np.array_equal(ckpt.stream_from_checkpoint(net, str(tmp_path / "again.ckpt")), stream)


# ---------------------------------------------------------------------------------------------------------------
# fixtures that this package's writer did NOT produce (oracle/make_golden_bundle.py: protobuf runtime + an independent
# LevelDB table builder with shortened separators; two data shards; Adam slot variables)
# ---------------------------------------------------------------------------------------------------------------
FIXTURE = os.path.join(helpers.GOLDEN, "bundle_tf_layout.ckpt")


def _three_conv_net():
    from tensorflow_yolo_b200.net import layers as L
    L.conv2d_bn_act.reset()
    graph = L.reset_default_graph()
    net = [L.input_layer([None, 32, 32, 3], "input")]
    net.append(L.conv2d_bn_act(net[-1].out, 8, 3))
    net.append(L.conv2d_bn_act(net[-1].out, 16, 3, 2))
    net.append(L.conv2d_bn_act(net[-1].out, 4, 1, use_batch_normalization=False, activation_fn="linear"))
    state = pbase.NetworkState(graph, "v3", 1, input_shape=(32, 32, 3))
    graph._yb_state = state
    net[0]._yb_state = state
    return net


def test_reads_checkpoint_not_written_by_own_writer(lib_built):
    exp = {k.replace("|", "/"): v for k, v in np.load(os.path.join(helpers.GOLDEN, "bundle_expected.npz")).items()}
    r = ckpt.BundleReader(FIXTURE)
    assert r.header == {"num_shards": 2, "endianness": 0}
    assert r.names() == sorted(exp) and len(exp) == 31
    for name, want in exp.items():
        got = r.get_tensor(name)
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), name
    assert {r.entries[n].shard_id for n in exp} == {0, 1}            # both .data-0000i-of-00002 files are used
    # the TEST path (net/yolo.py:71-78): only layer.variable_names are read; Adam slots, beta*_power, global_step are ignored
    net = _three_conv_net()
    stream = ckpt.stream_from_checkpoint(net, FIXTURE)
    parts = []
    for i, bn in ((0, True), (1, True), (2, False)):
        stem = "yolo/conv2d_bn_act_%d/" % i
        for leaf in (["beta", "gamma", "moving_mean", "moving_variance"] if bn else ["bias"]):
            parts.append(exp[stem + leaf].reshape(-1))
        parts.append(np.transpose(exp[stem + "kernel"], (3, 2, 0, 1)).reshape(-1))       # HWIO -> OIHW (net/base.py:36-40)
    assert np.array_equal(stream, np.concatenate(parts)) and stream.size == P.weight_count(net[0]._yb_state.graph.specs)


def test_fixture_bytes_are_the_documented_layout():
    """Literal bytes of the independently produced .index, field by field (LevelDB table_format.md, tensor_bundle.proto)."""
    data = open(FIXTURE + ".index", "rb").read()
    # footer: metaindex handle (offset 1061, size 8), index handle (offset 1074, size 102) as varint64s, zero padding to
    # 40 bytes, magic 0xdb4775248b80fb57 little-endian
    assert data[-48:].hex() == "a50808b20866" + "00" * 34 + "57fb808b247547db"
    # index block: restart interval 1; keys are SHORTENED separators -- "yolo/conv2d_bn_act_0/moving_n" sits between
    # .../moving_mean and .../moving_variance, "z" is the short successor of the last key -- values are block handles
    idx = data[1074:1074 + 102]
    keys = [k for k, _ in ckpt._block_entries(idx)]
    assert keys == [b"yolo/conv2d_bn_act_0/moving_n", b"yolo/conv2d_bn_act_1/moving_variance", b"z"]
    assert struct.unpack("<I", idx[-4:])[0] == 3                                  # three restart points
    # first data block, first entries: shared=0 non_shared=0 value_len=6 | key "" | BundleHeaderProto
    #   08 02 = num_shards: 2;   1a 02 08 01 = version { producer: 1 }   (endianness LITTLE = 0 is not serialised)
    assert data[:9].hex() == "000006" + "08021a020801"
    # second entry: shared=0 non_shared=11 value_len=11 | "beta1_power" | BundleEntryProto
    #   08 01 dtype DT_FLOAT | (shape: scalar, empty message omitted) | 18 01 shard_id 1 | (offset 0 omitted) | 28 04 size 4 |
    #   35 <fixed32> masked crc32c
    assert data[9:12].hex() == "000b0b" and data[12:23] == b"beta1_power"
    assert data[23:30].hex() == "08011801280435"
    value = np.asarray(0.9 ** 40, dtype=np.float32).tobytes()
    assert struct.unpack("<I", data[30:34])[0] == ckpt.mask_crc(ckpt.crc32c_py(value))
    # third entry shares the 4-byte prefix "beta" with its predecessor: shared=4 non_shared=7 ("2_power")
    assert data[34:37].hex() == "04070d" and data[37:44] == b"2_power"
    # block trailer: compression type 0 + masked CRC-32C over contents and type
    handles = [v for _, v in ckpt._block_entries(idx)]
    off, q = ckpt._get_varint(handles[0], 0)
    size, _ = ckpt._get_varint(handles[0], q)
    assert (off, data[off + size]) == (0, 0)
    assert struct.unpack("<I", data[off + size + 1:off + size + 5])[0] == ckpt.mask_crc(ckpt.crc32c_py(data[off:off + size + 1]))


def test_fixture_corruption_and_missing_shard_are_detected(tmp_path, lib_built):
    import shutil
    for ext in (".index", ".data-00000-of-00002", ".data-00001-of-00002"):
        shutil.copy(FIXTURE + ext, str(tmp_path / ("c.ckpt" + ext)))
    prefix = str(tmp_path / "c.ckpt")
    blob = bytearray(open(prefix + ".data-00001-of-00002", "rb").read())
    blob[100] ^= 0x40
    open(prefix + ".data-00001-of-00002", "wb").write(bytes(blob))
    r = ckpt.BundleReader(prefix)
    bad = [n for n in r.names() if r.entries[n].shard_id == 1 and r.entries[n].offset <= 100 < r.entries[n].offset + r.entries[n].size]
    assert len(bad) == 1
    with pytest.raises(ckpt.CheckpointError, match="CRC"):
        r.get_tensor(bad[0])
    os.remove(prefix + ".data-00000-of-00002")
    with pytest.raises(ckpt.CheckpointError, match="does not exist"):
        ckpt.BundleReader(prefix).get_tensor("global_step")


def test_pure_python_crc32c_matches_native(lib_built):
    data = np.random.RandomState(1).bytes(4099)
    assert ckpt.crc32c_py(b"123456789") == 0xe3069283 and ckpt.crc32c_py(bytes(32)) == 0x8a9136aa
    assert ckpt.crc32c_py(data) == _lib.crc32c(data)
    assert ckpt.crc32c_py(data[100:], seed=ckpt.crc32c_py(data[:100])) == ckpt.crc32c_py(data)
