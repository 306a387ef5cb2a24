"""Import the *unmodified* reference package from /root/reference -- TEST INFRASTRUCTURE.

Only usable in the build container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the CPU tests that
validate the oracle restatements, which skip when the reference is absent.

``net/base.py:5-9`` and ``net/layers.py:1`` import tensorflow / imgaug at module top level
(and ``base.py:15-23`` calls ``iaa.*`` at import time), neither of which is installed.
We pre-seed ``sys.modules`` with
  * ``tensorflow``  -> ``oracle.tfstub`` (a small TF1 emulator over torch-CPU fp32), so that the
    reference's graph builders and weight loader really execute;
  * ``imgaug`` / ``imgaug.augmenters`` -> MagicMock (training-only, never executed).
"""
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("YOLO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "net", "v3.py"))


_cached = None


def load():
    """Returns a namespace with the reference modules: base, layers, v2, v3, yolo, tf."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise ImportError("reference checkout not found at " + REFERENCE_ROOT)
    from oracle import tfstub
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "imgaug", "imgaug.augmenters", "net")}
    saved_net = {k: v for k, v in sys.modules.items() if k == "net" or k.startswith("net.")}
    for k in saved_net:
        del sys.modules[k]
    sys.modules["tensorflow"] = tfstub
    ia = MagicMock()
    sys.modules["imgaug"] = ia
    sys.modules["imgaug.augmenters"] = ia.augmenters
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import net.base as base
        import net.layers as layers
        import net.v2 as v2
        import net.v3 as v3
        import net.yolo as yolo
    finally:
        sys.path.remove(REFERENCE_ROOT)
        # keep the reference's modules private to this namespace: our own product package
        # also has a sub-package called ``net`` (tensorflow_yolo_b200.net), never top-level.
        for k in [k for k in sys.modules if k == "net" or k.startswith("net.")]:
            del sys.modules[k]
        sys.modules.update(saved_net)
        for k, v in saved.items():
            if k == "net":
                continue
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    class _Ref(object):
        pass

    ref = _Ref()
    ref.base, ref.layers, ref.v2, ref.v3, ref.yolo, ref.tf = base, layers, v2, v3, yolo, tfstub
    _cached = ref
    return ref
