"""Deterministic synthetic inputs for tests and benches: darknet ``.weights`` streams and images.

The reference ships no pretrained weights (``bin/`` is git-ignored) and there is no network, so
benches and parity tests run on random-init weights written in the reference's own on-disk layout
(net/base.py:26-46 order; headers net/v3.py:101-104 and net/v2.py:68-77) -- the same file is read
by the CUDA engine and by the CPU oracle.

The initialisation is variance-preserving so that activations stay O(1) through all 75 (v3) /
23 (v2) convs without any data-dependent calibration (a naive He init with random BN statistics
overflows the 23-block residual trunk; TF's default glorot/identity-BN init collapses to 0):
  * conv kernel ~ N(0, 1/fan_in);
  * BN: gamma ~ U(0.9,1.1) (x0.5 on convs that feed a shortcut), beta ~ N(0,0.1),
        moving_mean ~ N(0,0.05), moving_variance ~ 0.505*U(0.9,1.1)
        (0.505 = E[leaky_0.1(z)^2] for z~N(0,1): the second moment arriving at the next conv);
  * head convs (linear + bias): kernel ~ N(0, head_std^2/(0.505*fan_in)), bias ~ N(0,0.5), with the
    objectness channels shifted by ``obj_bias`` to set the candidate density after thresholding.
"""
import numpy as np

from . import plan as _plan


def weight_stream(plan, seed=2, num_classes=80, head_std=1.5, obj_bias=-4.0):
    """float32 stream in darknet order for ``plan`` (list of LayerSpec)."""
    rng = np.random.RandomState(seed)
    chunks = []
    for _, cin, cout, k, bn, feeds_shortcut, is_head in _plan.conv_specs(plan):
        fan_in = cin * k * k
        if bn:
            gamma = rng.uniform(0.9, 1.1, cout) * (0.5 if feeds_shortcut else 1.0)
            beta = rng.normal(0.0, 0.1, cout)
            mean = rng.normal(0.0, 0.05, cout)
            var = 0.505 * rng.uniform(0.9, 1.1, cout)
            chunks += [beta, gamma, mean, var]
            kernel = rng.standard_normal(cout * fan_in) * np.sqrt(1.0 / fan_in)
        else:
            bias = rng.normal(0.0, 0.5, cout)
            per_anchor = 5 + num_classes
            if cout % per_anchor == 0:
                bias[4::per_anchor] += obj_bias
            chunks.append(bias)
            kernel = rng.standard_normal(cout * fan_in) * (head_std / np.sqrt(0.505 * fan_in))
        chunks.append(kernel)
    return np.concatenate(chunks).astype(np.float32)


def write_weights_v3(path, stream, header=(0, 2, 0, 0, 0)):
    """5 x int32 header then the float32 stream (net/v3.py:101-104)."""
    with open(path, "wb") as f:
        np.asarray(header, dtype=np.int32).tofile(f)
        np.asarray(stream, dtype=np.float32).tofile(f)


def write_weights_v2(path, stream, major=0, minor=1, revision=0, seen=0):
    """3 x int32 + one 4-byte ``seen`` then the float32 stream (net/v2.py:68-77)."""
    with open(path, "wb") as f:
        np.asarray([major, minor, revision], dtype=np.int32).tofile(f)
        if (major * 10 + minor) >= 2 and major < 1000 and minor < 1000:
            np.asarray([seen], dtype=np.float32).tofile(f)
        else:
            np.asarray([seen], dtype=np.int32).tofile(f)
        np.asarray(stream, dtype=np.float32).tofile(f)


def images(n, h, w, c=3, seed=1):
    """Uniform [0,1) RGB NHWC float32 (the reference feeds RGB/255, net/base.py:122,153)."""
    return np.random.RandomState(seed).random_sample((n, h, w, c)).astype(np.float32)


def head_tensor(n, rows, cols, seed=0, obj_shift=0.0):
    """N(0,1) logits in the reference's ``net[-1].out`` layout [n, rows, cols] (BASELINE config 5)."""
    t = np.random.RandomState(seed).standard_normal((n, rows, cols)).astype(np.float32)
    if obj_shift:
        t[..., 4] += np.float32(obj_shift)
    return t
