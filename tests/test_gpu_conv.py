"""GPU: the conv stack (tcgen05/TMA implicit GEMM + fused epilogues) through the C ABI vs the fp32 oracle.
Bar (north_star): ||gpu - ref|| / ||ref|| <= 2e-2 per output tensor and per materialised layer
(bf16 storage with fp32 accumulation vs fp32)."""
import os

import numpy as np
import pytest

import helpers
from oracle import convstack, make_golden
from tensorflow_yolo_b200 import engine, plan as P, synth
from tensorflow_yolo_b200.net import base as pbase

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _engine(net, shape, nc, mode, n, stream):
    eng = engine.Engine(net[0]._yb_state.plan(), shape, nc, mode, max_batch=n)
    assert eng.load_weights(stream) == stream.size
    return eng


def _per_layer(eng, topo, stream, x):
    _, outs = convstack.forward(topo, stream, x, return_all=True)
    worst = 0.0
    checked = 0
    for i, spec in enumerate(eng.plan[:len(topo)]):
        if spec.kind in (P.KIND_YOLO, P.KIND_DETECTION):
            continue
        try:
            v = eng.read_layer(i)
        except Exception as ex:
            assert "fused" in str(ex), str(ex)
            continue
        e = helpers.rel_err(v, outs[i].permute(0, 2, 3, 1).numpy())
        assert e <= TOL, (i, spec.as_dict(), e)
        worst = max(worst, e)
        checked += 1
    return worst, checked


@pytest.fixture
def keep_all(monkeypatch):
    monkeypatch.setenv("YB_KEEP_ALL", "1")


def test_v3_small_golden_and_layers(keep_all):
    g = helpers.golden("conv_v3.npz")
    shape = make_golden.CONV_V3_SHAPE                      # 96 x 64: H != W on purpose
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    x = synth.images(2, shape[0], shape[1], seed=1)
    eng = _engine(net, shape, 80, engine.YB_DECODE_V3, 2, stream)
    eng.forward(x)
    y = eng.read_output()
    assert y.shape == g["net_out"].shape
    assert helpers.rel_err(y, g["net_out"]) <= TOL          # reference builder + loader over the TF stub
    worst, checked = _per_layer(eng, topo, stream, x)
    assert checked >= 60
    # the CUDA-core cross-check kernel computes the same thing from the same packed operands
    eng.set_conv_impl(1)
    eng.forward(x)
    # (two bf16-storage pipelines with different fp32 summation orders differ by flipped bf16 roundings)
    assert helpers.rel_err(eng.read_output(), y) < 1.5e-2
    eng.close()


def test_v2_small_golden_and_layers(keep_all):
    g = helpers.golden("conv_v2.npz")
    shape = make_golden.CONV_V2_SHAPE
    net, topo, stream = helpers.build_v2(shape, 20, seed=3)
    x = synth.images(2, shape[0], shape[1], seed=4)
    eng = _engine(net, shape, 20, engine.YB_DECODE_V2, 2, stream)
    eng.forward(x)
    y = eng.read_output().reshape(g["net_out"].shape)
    assert helpers.rel_err(y, g["net_out"]) <= TOL
    _per_layer(eng, topo, stream, x)
    eng.close()


def test_v3_416_full_size_layers(keep_all):
    """Every GEMM shape of Appendix B at its real size (batch 2), layer by layer."""
    shape = (416, 416, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    x = synth.images(2, 416, 416, seed=1)
    eng = _engine(net, shape, 80, engine.YB_DECODE_V3, 2, stream)
    eng.forward(x)
    y = eng.read_output()
    ref = convstack.forward(topo, stream, x)
    assert y.shape == (2, 10647, 85)
    assert helpers.rel_err(y, ref) <= TOL
    _per_layer(eng, topo, stream, x)
    eng.close()


def test_v3_608_config4_output_and_batch_consistency():
    """BASELINE config 4: the 608 x 608 network (R = 22,743 rows).  Output vs the fp32 oracle for one image, and a
    batch of 6 through an autotuned engine must reproduce that image bit for bit at every batch position."""
    shape = (608, 608, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    x = synth.images(1, 608, 608, seed=11)
    eng = _engine(net, shape, 80, engine.YB_DECODE_V3, 6, stream)
    eng.forward(x)
    y = eng.read_output()
    assert y.shape == (1, 22743, 85)
    assert helpers.rel_err(y, convstack.forward(topo, stream, x)) <= TOL
    eng.autotune(6, reps=2)
    xb = np.concatenate([x] * 6, 0)
    eng.forward(xb)
    yb_ = eng.read_output()
    for i in range(6):
        assert np.array_equal(yb_[i], y[0]), i
    dets = eng.detect(0.5, 0.6)
    assert all(np.array_equal(d, dets[0]) for d in dets)
    eng.close()


def test_v2_416_coco_and_voc():
    for nc, anchors in ((80, helpers.V2_ANCHORS_COCO), (20, helpers.V2_ANCHORS_VOC)):
        shape = (416, 416, 3)
        net, topo, stream = helpers.build_v2(shape, nc, seed=3, anchors=anchors)
        x = synth.images(1, 416, 416, seed=6)
        eng = _engine(net, shape, nc, engine.YB_DECODE_V2, 1, stream)
        eng.forward(x)
        y = eng.read_output().reshape(1, 13, 13, 5 * (5 + nc))
        assert helpers.rel_err(y, convstack.forward(topo, stream, x)) <= TOL
        eng.close()


def test_arena_reuse_partial_batch_u8_and_device_input():
    """Default engine (arena recycling on): results must not depend on max_batch, on the batch position of an
    image, on uint8 vs float input, or on host vs device residency of the input."""
    import torch
    shape = (128, 160, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=4)
    x8 = np.random.RandomState(3).randint(0, 256, size=(5, 128, 160, 3)).astype(np.uint8)
    xf = (x8 / 255.).astype(np.float32)
    eng = _engine(net, shape, 80, engine.YB_DECODE_V3, 8, stream)
    eng.forward(xf)
    y = eng.read_output()
    assert helpers.rel_err(y, convstack.forward(topo, stream, xf)) <= TOL
    eng.forward(xf[3:4])
    assert np.array_equal(eng.read_output()[0], y[3])                 # per-image independence, bit-exact
    eng.forward(x8)
    assert np.array_equal(eng.read_output(), y)                       # u8 path == float path
    eng.forward(torch.from_numpy(xf).cuda())
    assert np.array_equal(eng.read_output(), y)
    # a device tensor that does not start on a 16-byte boundary (a view into a larger buffer): the TMA-fed first layers
    # cannot address it, the engine stages it -- same bits
    big8 = torch.zeros(x8.size + 64, dtype=torch.uint8, device="cuda")
    odd8 = big8[3:3 + x8.size].view(x8.shape)
    odd8.copy_(torch.from_numpy(x8))
    assert odd8.data_ptr() % 16 != 0
    eng.forward(odd8)
    assert np.array_equal(eng.read_output(), y)
    bigf = torch.zeros(xf.size + 16, dtype=torch.float32, device="cuda")
    oddf = bigf[1:1 + xf.size].view(xf.shape)
    oddf.copy_(torch.from_numpy(xf))
    assert oddf.data_ptr() % 16 != 0
    eng.forward(oddf)
    assert np.array_equal(eng.read_output(), y)
    with pytest.raises(Exception):
        eng.forward(np.zeros((9, 128, 160, 3), np.float32))           # > max_batch
    with pytest.raises(Exception):
        eng.read_layer(5)                                             # recycled intermediate without YB_KEEP_ALL
    eng.close()


def test_error_paths():
    shape = (64, 64, 3)
    net, _, stream = helpers.build_v3(shape, 80, seed=2)
    eng = engine.Engine(net[0]._yb_state.plan(), shape, 80, engine.YB_DECODE_V3, max_batch=1)
    with pytest.raises(Exception) as ei:
        eng.forward(np.zeros((1, 64, 64, 3), np.float32))              # before load_weights
    assert "load_weights" in str(ei.value)
    with pytest.raises(Exception) as ei:
        eng.load_weights(stream[:-1])
    assert ei.value.code == -3                                         # YB_ERR_SHORT_WEIGHTS
    assert eng.load_weights(np.concatenate([stream, np.zeros(7, np.float32)])) == stream.size   # surplus only reported
    eng.close()


def test_session_run_matches_reference_protocol():
    """sess.run(ops); sess.run(net[-1].out, {net[0].out: x}) -- net/yolo.py:74-83."""
    shape = (64, 64, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    x = synth.images(2, 64, 64, seed=8).astype(np.float64)             # the reference feeds float64
    with pbase.Session(net) as sess:
        sess.run(pbase.load_weights(net, stream))
        out = sess.run(net[-1].out, feed_dict={net[0].out: x})
    assert out.shape == (2, 252, 85) and out.dtype == np.float32
    assert helpers.rel_err(out, convstack.forward(topo, stream, x)) <= TOL


def test_autotune_keeps_results_bit_identical():
    """Every launch configuration walks K in the same order, so tuning must not change a single output bit;
    the same holds for programmatic dependent launch on/off."""
    shape = (96, 64, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    x = synth.images(3, shape[0], shape[1], seed=5)
    eng = _engine(net, shape, 80, engine.YB_DECODE_V3, 3, stream)
    eng.forward(x)
    y0 = eng.read_output()
    report = eng.autotune(3, reps=2)
    assert len(report["ops"]) >= 10 and all(len(o["candidates"]) >= 2 for o in report["ops"])
    with pytest.raises(Exception):
        eng.read_output()                     # activations were clobbered: a forward has to come first
    eng.forward(x)
    assert np.array_equal(eng.read_output(), y0)
    eng.set_option("pdl", 0)
    eng.forward(x)
    assert np.array_equal(eng.read_output(), y0)
    eng.set_option("pdl", 1)
    # forced configurations: every N tile / pair / stationary / epilogue variant of one mid-network conv
    n_ops = len(eng.profile(x))
    tc_ops = [i for i in range(n_ops) if eng.op_info(i)["path"] == 0]
    for op_i in tc_ops[3:40:6]:
        for bn in (32, 64, 128, 256):
            for pair in (0, 1):
                for bstat in (0, 1):
                    for te, ksub in ((0, 0), (1, 0), (1, 1), (0, 2), (1, 3), (1, 9)):
                        try:
                            eng.set_conv_cfg(op_i, bn, pair, bstat, te, ksub)
                        except Exception:
                            continue          # not available for this layer (N tile wider than Cout, pair without BN=256)
                        eng.forward(x)
                        assert np.array_equal(eng.read_output(), y0), (op_i, bn, pair, bstat, te, ksub)
        eng.set_conv_cfg(op_i, 0)
    with pytest.raises(Exception):
        eng.set_option("no_such_option", 1)
    assert eng.time_op(tc_ops[0], 3, 2) > 0.0
    eng.close()


def test_cuda_graph_forward_is_bit_identical_and_replayed():
    """Small batches replay a captured CUDA graph of the 75 launches (eager the first time a (batch, input buffer, dtype)
    is seen, captured the second time): the outputs and the detections must equal the eager path's bit for bit, for host
    inputs (two staging buffers -> two graphs), device inputs, uint8 inputs and a different batch size in between."""
    import torch
    shape = (96, 64, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2, obj_bias=-1.0)
    x = synth.images(3, shape[0], shape[1], seed=5)
    x8 = (x * 255).astype(np.uint8)
    eng = _engine(net, shape, 80, engine.YB_DECODE_V3, 3, stream)
    eng.set_option("graph", 0)
    eng.forward(x)
    y0 = eng.read_output()
    d0 = eng.detect(0.5, 0.6)
    eng.forward(x8)
    y8 = eng.read_output()
    eng.forward(x[:2])
    y2 = eng.read_output()
    assert eng.graph_replays() == 0
    eng.set_option("graph", -1)                # automatic: batch 3 <= 32
    xd = torch.from_numpy(x).cuda()
    for it in range(6):
        eng.forward(x)                         # host input: staging buffers alternate
        assert np.array_equal(eng.read_output(), y0), it
        d = eng.detect(0.5, 0.6)
        assert all(np.array_equal(a["row"], b["row"]) for a, b in zip(d, d0))
        eng.forward(xd)                        # device input: one more key
        assert np.array_equal(eng.read_output(), y0), it
        eng.forward(x[:2])
        assert np.array_equal(eng.read_output(), y2), it
        eng.forward(x8)
        assert np.array_equal(eng.read_output(), y8), it
    assert eng.graph_replays() >= 12
    fwd, _ = eng.launch_count()
    assert fwd == 73          # 75 convs; the first conv and the 1x1 of the first block run inside their consumers (conv_fused.cuh)
    # a configuration change drops the captured graphs; results stay the same
    n_before = eng.graph_replays()
    eng.set_option("pdl", 0)
    eng.forward(xd)
    assert np.array_equal(eng.read_output(), y0) and eng.graph_replays() == n_before
    eng.close()


def test_pixel_pair_view_is_bit_identical(monkeypatch):
    """The Cin=32 3x3 convs can be launched over pairs of pixels with zero-padded weights (YB_PIXEL_PAIRS): layer 4 (stride 1,
    +residual) with both sides paired (1, the default), layer 2 (stride 2) with the input side paired and 3 x 2 taps (2,
    measured slower and therefore off).  Same bytes in and out, and -- the extra products being exact zeros, the K walk
    visiting the real ones in the same order -- the same bits."""
    shape = (128, 160, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=4)
    x = synth.images(3, shape[0], shape[1], seed=5)
    got = {}
    monkeypatch.setenv("YB_KEEP_ALL", "1")
    for flag in ("2", "1", "0"):
        monkeypatch.setenv("YB_PIXEL_PAIRS", flag)
        eng = _engine(net, shape, 80, engine.YB_DECODE_V3, 3, stream)
        eng.forward(x)
        # layer 2 = the stride-2 conv, layer 5 = conv 4 + shortcut
        got[flag] = (eng.read_layer(2)[0], eng.read_layer(5)[0], eng.read_output(), eng.op_cfg(3)["bn"], eng.op_info(1)["bk"])
        eng.close()
    assert [got[f][3] for f in "210"] == [128, 128, 64]         # the paired problem has twice the output channels
    assert [got[f][4] for f in "210"] == [64, 32, 32]           # ... and 64 input channels per K block
    for flag in ("2", "1"):
        for a, b in zip(got[flag][:3], got["0"][:3]):
            assert np.array_equal(a, b)
    assert helpers.rel_err(got["1"][2], convstack.forward(topo, stream, x)) <= TOL


def test_split_k_latency_mode_matches_the_plain_kernels():
    """Option "split_k": convs whose grid fills a fraction of the SMs are computed by 2-4 work units per tile along K and
    reduced in a fixed order.  Same products, a different fp32 summation order: the network output stays within bf16
    rounding of the plain path and within TOL of the fp32 oracle, repeated runs are bit-identical (deterministic
    reduction), some convs really run split, and switching the option off restores the plain bits."""
    shape = (416, 416, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    x = synth.images(2, shape[0], shape[1], seed=1)
    eng = _engine(net, shape, 80, engine.YB_DECODE_V3, 2, stream)
    eng.set_option("graph", 0)
    eng.forward(x)
    y_plain = eng.read_output()
    eng.set_option("split_k", 1)
    eng.forward(x)
    y_split = eng.read_output()
    factors = []
    for i in range(len(eng.plan)):
        try:
            if eng.op_info(i)["path"] == 0:
                factors.append(eng.op_cfg(i)["splitk"])
        except Exception:
            break                                                       # past the last launched op
    assert max(factors) >= 2 and min(factors) >= 1, factors
    eng.forward(x)
    assert np.array_equal(eng.read_output(), y_split)                   # fixed summation order, whichever part arrives last
    eng.set_option("graph", 1)                                          # and inside a captured graph
    for _ in range(3):
        eng.forward(x)
        assert np.array_equal(eng.read_output(), y_split)
    assert helpers.rel_err(y_split, y_plain) <= 1.5e-2, helpers.rel_err(y_split, y_plain)   # bf16 rounding flips through 75 layers
    ref = convstack.forward(topo, stream, x)
    assert helpers.rel_err(y_split, ref) <= TOL
    eng.set_option("split_k", 0)
    eng.set_option("graph", 0)
    eng.forward(x)
    assert np.array_equal(eng.read_output(), y_plain)
    eng.close()
