"""Oracle (test infrastructure only): CPU restatement of the reference's image preprocessing.

``preprocess_image`` (net/base.py:115-155, TEST path: no augmentation) is
    image = cv2.imread(path); net_image = cv2.resize(image, tuple(new_shape[0:2])); net_image[:, :, ::-1] / 255.
The arithmetic lives in OpenCV (requirements.txt: opencv-python), whose INTER_LINEAR for 8-bit images is restated here
from modules/imgproc/src/resize.cpp (cv::resize -> resizeGeneric_ with HResizeLinear / VResizeLinear for uchar,
INTER_RESIZE_COEF_BITS = 11, and the INTER_AREA reroute of exact 2x decimations).  cv2 itself IS installed in this image,
so the restatement is pinned bit for bit against cv2.resize (tests/test_preprocess.py) -- not only against fixtures.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _table(src, dst, clamp):
    """(offsets int32 [dst], weights int32 [dst, 2]) in cv2's types: double scale = 1/(dst/src), float32 fraction."""
    inv_scale = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:                                  # columns (resize.cpp: "if( sx < 0 ) fx = 0, sx = 0; if( sx >= ssize.width-1 ) ...")
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= src - 1
        f[hi], s[hi] = 0, src - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)     # cvRound: half to even
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    return s, np.stack([w0, w1], 1)


def resize_linear_u8(img, dst_w, dst_h):
    """cv2.resize(img, (dst_w, dst_h)) for uint8 [h, w, c] (interpolation = INTER_LINEAR, the default)."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    sh, sw = img.shape[:2]
    s = img.astype(np.int32)
    if sw == 2 * dst_w and sh == 2 * dst_h:    # resize(): INTER_LINEAR with iscale_x == iscale_y == 2 -> INTER_AREA fast path
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    xo, xa = _table(sw, dst_w, True)
    yo, yb = _table(sh, dst_h, False)
    x1 = np.minimum(xo + 1, sw - 1)
    hor = s[:, xo] * xa[:, 0][None, :, None] + s[:, x1] * xa[:, 1][None, :, None]          # HResizeLinear, int
    r0, r1 = np.clip(yo, 0, sh - 1), np.clip(yo + 1, 0, sh - 1)                            # clip(sy0 + k, 0, ssize.height)
    out = (((yb[:, 0][:, None, None] * (hor[r0] >> 4)) >> 16) + ((yb[:, 1][:, None, None] * (hor[r1] >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def preprocess_u8(image_bgr, new_shape):
    """The uint8 RGB array the reference divides by 255: resize to dsize = (new_shape[0], new_shape[1]) -- i.e. width =
    input_h, height = input_w, the reference's quirk (net/base.py:121) -- then BGR -> RGB."""
    return resize_linear_u8(image_bgr, int(new_shape[0]), int(new_shape[1]))[:, :, ::-1]


def preprocess_image(image_bgr, new_shape):
    """float64 RGB in [0,1], exactly what preprocess_image returns for a decoded image."""
    return preprocess_u8(image_bgr, new_shape) / 255.
