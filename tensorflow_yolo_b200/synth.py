"""Deterministic synthetic inputs for tests and benches: darknet ``.weights`` streams and images.

The reference ships no pretrained weights (``bin/`` is git-ignored) and there is no network, so
benches and parity tests run on random-init weights written in the reference's own on-disk layout
(net/base.py:26-46 order; headers net/v3.py:101-104 and net/v2.py:68-77) -- the same file is read
by the CUDA engine and by the CPU oracle.

Initialisation (kernels first, then statistics that keep every activation O(1)):
  * conv kernel ~ N(0, 1/fan_in);
  * BN: gamma ~ U(0.9,1.1) (x0.5 on convs that feed a shortcut), beta ~ N(0,0.1),
        moving_mean = m + N(0,0.05)*sqrt(v), moving_variance = v*U(0.9,1.1), where (m, v) is the
        measured mean/variance of that conv's pre-BN output over all channels and pixels.  A naive
        init with unit statistics lets the 23-block residual trunk grow to rms ~40 and saturates the
        heads; TF's default (glorot, identity BN) collapses to 0.  The (m, v) pairs are ~100 scalars
        per network measured once with the fp32 CPU oracle on seed-2 kernels
        (oracle/calibrate_synth.py -> synth_calib.json, keyed by the BN-conv shape list); networks
        without a table fall back to (0, 0.505);
  * head convs (linear + bias): kernel ~ N(0, head_std^2/(m2*fan_in)) with m2 the measured second
    moment of the head's input, bias ~ N(0,0.5), objectness channels shifted by ``obj_bias`` to set
    the candidate density after thresholding.
"""
import json
import os

import numpy as np

from . import plan as _plan


def _signature(plan):
    return "|".join("{}-{}-{}".format(cin, cout, k) for _, cin, cout, k, bn, _, _ in _plan.conv_specs(plan) if bn)


_calib_cache = None


def load_calibration(plan):
    """{'bn': [[mean, var], ...] per BN conv, 'head_m2': [...] per head conv} or None."""
    global _calib_cache
    if _calib_cache is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_calib.json")
        _calib_cache = {}
        if os.path.exists(path):
            with open(path) as f:
                for entry in json.load(f)["tables"]:
                    _calib_cache[entry["signature"]] = entry
    return _calib_cache.get(_signature(plan))


def weight_stream(plan, seed=2, num_classes=80, head_std=1.5, obj_bias=-4.0, calib="auto"):
    """float32 stream in darknet order for ``plan`` (list of LayerSpec).  ``calib``: "auto" (table lookup),
    None (unit statistics) or an explicit table dict."""
    if isinstance(calib, str):
        calib = load_calibration(plan)
    rng = np.random.RandomState(seed)
    chunks = []
    i_bn = i_head = 0
    for _, cin, cout, k, bn, feeds_shortcut, is_head in _plan.conv_specs(plan):
        fan_in = cin * k * k
        if bn:
            m, v = (calib["bn"][i_bn] if calib else (0.0, 0.505))
            i_bn += 1
            gamma = rng.uniform(0.9, 1.1, cout) * (0.5 if feeds_shortcut else 1.0)
            beta = rng.normal(0.0, 0.1, cout)
            mean = m + rng.normal(0.0, 0.05, cout) * np.sqrt(v)
            var = v * rng.uniform(0.9, 1.1, cout)
            chunks += [beta, gamma, mean, var]
            kernel = rng.standard_normal(cout * fan_in) * np.sqrt(1.0 / fan_in)
        else:
            m2 = (calib["head_m2"][i_head] if calib else 0.505)
            i_head += 1
            bias = rng.normal(0.0, 0.5, cout)
            per_anchor = 5 + num_classes
            if cout % per_anchor == 0:
                bias[4::per_anchor] += obj_bias
            chunks.append(bias)
            kernel = rng.standard_normal(cout * fan_in) * (head_std / np.sqrt(m2 * fan_in))
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
