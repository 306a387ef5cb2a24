"""Generates tests/golden/preprocess.npz -- TEST INFRASTRUCTURE, run in the build container only.

    python -m oracle.make_golden_pre

Source images (seeded noise and a smooth ramp, written to PNG and read back with cv2.imread like the reference does) and
what the UNMODIFIED reference `net.base.preprocess_image` (net/base.py:115-155, i.e. cv2.resize + BGR->RGB + /255.)
returns for them, stored as the uint8 value before the division (the division by 255. is exact to recover:
out * 255 rounds back to the integer).  Covers up- and down-scaling, the exact-2x INTER_AREA reroute and a non-square
target (the dsize quirk: shape (input_w, input_h, 3)).
"""
import os
import tempfile

import numpy as np

from oracle import refimport

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [((37, 50), (64, 64, 3)), ((90, 120), (64, 64, 3)), ((128, 128), (64, 64, 3)), ((100, 75), (96, 96, 3)),
         ((33, 200), (48, 80, 3)), ((240, 31), (80, 48, 3))]


def main():
    import cv2
    ref_base = refimport.load().base
    rs = np.random.RandomState(21)
    out = {"n_cases": np.int64(len(CASES))}
    with tempfile.TemporaryDirectory() as d:
        for i, ((h, w), shape) in enumerate(CASES):
            if i % 2 == 0:
                img = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
            else:
                yy, xx = np.mgrid[0:h, 0:w]
                img = np.stack([xx * 255 // max(w - 1, 1), yy * 255 // max(h - 1, 1), (xx * 3 + yy * 5) % 256], 2).astype(np.uint8)
            path = os.path.join(d, "im%d.png" % i)
            cv2.imwrite(path, img)
            want, _ = ref_base.preprocess_image(path, shape)
            u8 = np.rint(want * 255.).astype(np.uint8)
            assert np.array_equal(u8 / 255., want)                 # the float64 result is exactly value / 255.
            out["src%d" % i] = cv2.imread(path)
            out["shape%d" % i] = np.asarray(shape, dtype=np.int64)
            out["rgb%d" % i] = u8
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "preprocess.npz"), **out)
    print("preprocess:", [(out["src%d" % i].shape, out["rgb%d" % i].shape) for i in range(len(CASES))])


if __name__ == "__main__":
    main()
