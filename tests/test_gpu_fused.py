"""GPU: the two-layer fused kernels of csrc/conv_fused.cuh (first conv + stride-2 conv; 1x1 + 3x3 + shortcut of the
64-channel residual block, net/v3.py:16-19,25-27) against (a) the fp32 oracle, 2e-2 per materialised layer, and
(b) the unfused tcgen05 kernels on the same inputs (same bf16 storage points, so the two agree to bf16 rounding flips)."""
import numpy as np
import pytest

import helpers
from oracle import convstack
from tensorflow_yolo_b200 import engine, synth

pytestmark = pytest.mark.gpu
TOL = 2e-2
STEM_OUT, BLOCK_OUT = 2, 5          # plan layers: conv 3x3 s2 (L2) and the first shortcut (L5)


def _layers(monkeypatch, fuse, net, shape, stream, x):
    monkeypatch.setenv("YB_KEEP_ALL", "1")
    monkeypatch.setenv("YB_FUSE_STEM", "2" if fuse else "0")
    monkeypatch.setenv("YB_FUSE_BLOCK", "2" if fuse else "0")
    eng = engine.Engine(net[0]._yb_state.plan(), shape, 80, engine.YB_DECODE_V3, max_batch=x.shape[0])
    eng.load_weights(stream)
    eng.forward(x)
    out = {i: eng.read_layer(i) for i in (STEM_OUT, BLOCK_OUT)}
    fused_away = 0
    for i in (1, 3):
        try:
            eng.read_layer(i)
        except Exception as ex:
            assert "fused" in str(ex)
            fused_away += 1
    assert fused_away == (2 if fuse else 0)
    y = eng.read_output()
    launches = eng.launch_count()[0]
    eng.close()
    return out, y, launches


@pytest.mark.parametrize("shape,n,u8", [((96, 64, 3), 3, False), ((64, 96, 3), 2, True), ((416, 416, 3), 2, False),
                                        ((416, 416, 3), 1, True)])
def test_fused_pairs_match_oracle_and_unfused(monkeypatch, shape, n, u8):
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    if u8:
        x = np.random.RandomState(5).randint(0, 256, (n,) + shape).astype(np.uint8)
        xf = (x / 255.).astype(np.float32)
    else:
        x = synth.images(n, shape[0], shape[1], seed=1)
        xf = x
    fused, y_f, l_f = _layers(monkeypatch, True, net, shape, stream, x)
    plain, y_p, l_p = _layers(monkeypatch, False, net, shape, stream, x)
    assert l_f == l_p - 2                                   # two launches fewer: L1 and L3 run inside their consumers
    _, outs = convstack.forward(topo, stream, xf, return_all=True)
    for i in (STEM_OUT, BLOCK_OUT):
        ref = outs[i].permute(0, 2, 3, 1).numpy()
        assert helpers.rel_err(fused[i], ref) <= TOL, (i, helpers.rel_err(fused[i], ref))
        # fused vs unfused: the same arithmetic up to fp32 summation order inside the producer conv
        assert helpers.rel_err(fused[i], plain[i]) <= 3e-3, (i, helpers.rel_err(fused[i], plain[i]))
    # the stem's producer runs the very same mma.sync sequence as the stand-alone first conv
    assert np.array_equal(fused[STEM_OUT], plain[STEM_OUT])
    assert helpers.rel_err(y_f, convstack.forward(topo, stream, xf)) <= TOL
    assert helpers.rel_err(y_f, y_p) <= 1.5e-2


def test_fused_is_batch_position_invariant(monkeypatch):
    """A tile never mixes images: the same image at every batch position gives the same bits."""
    shape = (128, 128, 3)
    net, topo, stream = helpers.build_v3(shape, 80, seed=2)
    x1 = synth.images(1, 128, 128, seed=9)
    x = np.concatenate([x1] * 5, 0)
    fused, y, _ = _layers(monkeypatch, True, net, shape, stream, x)
    for i in (STEM_OUT, BLOCK_OUT):
        for b in range(1, 5):
            assert np.array_equal(fused[i][b], fused[i][0]), (i, b)


@pytest.mark.parametrize("shape,n,u8", [((96, 64, 3), 3, False), ((416, 416, 3), 2, True)])
def test_convs_with_max_pool_in_registers_are_bit_identical(monkeypatch, shape, n, u8):
    """Darknet-19's conv 3->32 + 2x2 max pool and conv 32->64 + 2x2 max pool (net/v2.py:20-24), each in one kernel: the
    pooled maps equal conv -> bf16 -> maxpool2 bit for bit (max commutes with the monotone bf16 rounding; the second conv
    walks K like the stand-alone tcgen05 kernel)."""
    net, topo, stream = helpers.build_v2(shape, 20, seed=3)
    if u8:
        x = np.random.RandomState(6).randint(0, 256, (n,) + shape).astype(np.uint8)
        xf = (x / 255.).astype(np.float32)
    else:
        x = synth.images(n, shape[0], shape[1], seed=4)
        xf = x
    out = {}
    for fuse in ("2", "0"):
        monkeypatch.setenv("YB_KEEP_ALL", "1")
        monkeypatch.setenv("YB_FUSE_POOL", fuse)
        eng = engine.Engine(net[0]._yb_state.plan(), shape, 20, engine.YB_DECODE_V2, max_batch=n)
        eng.load_weights(stream)
        eng.forward(x)
        out[fuse] = (eng.read_layer(2), eng.read_layer(4), eng.read_output(), eng.launch_count()[0])
        if fuse == "2":
            for conv in (1, 3):
                with pytest.raises(Exception, match="fused"):
                    eng.read_layer(conv)
        eng.close()
    assert out["2"][3] == out["0"][3] - 2                          # two launches fewer
    assert np.array_equal(out["2"][0], out["0"][0])                # first pooled activation, bit for bit
    assert np.array_equal(out["2"][1], out["0"][1])                # second pooled activation
    assert np.array_equal(out["2"][2], out["0"][2])                # and therefore the whole network output
    _, outs = convstack.forward(topo, stream, xf, return_all=True)
    assert helpers.rel_err(out["2"][0], outs[2].permute(0, 2, 3, 1).numpy()) <= TOL
    assert helpers.rel_err(out["2"][1], outs[4].permute(0, 2, 3, 1).numpy()) <= TOL
