"""CPU oracle for the YOLO TEST hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of wns349/tensorflow-yolo's
inference path (conv stack -> head decode -> greedy NMS).  It is the checker the
CUDA path is compared against.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``tensorflow_yolo_b200`` never does.

Pinning status (see DESIGN.md "Oracle"):
  * decode + NMS (``oracle.postprocess``): pinned.  Checked bit-for-bit (NMS) and to
    1 ulp (decode transcendental) against the reference's own, unmodified functions
    ``net.v3._find_bounding_boxes``, ``net.v2._find_bounding_boxes`` and
    ``net.base.non_maximum_suppression`` executed in the build container
    (``oracle/refimport.py`` + ``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
  * conv stack (``oracle.convstack``): PARITY UNPINNED.  The arithmetic lives in
    un-vendored TensorFlow 1.x (requirements.txt:5,7) which is not installable here and
    the reference ships no tests or golden tensors.  The restatement follows
    net/layers.py line by line and its topology is pinned against the reference's
    builders executed under a recording TensorFlow stub (``oracle/refimport.py``).
"""
