"""TEST INFRASTRUCTURE ONLY (oracle): numpy restatement of `cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)` for
uint8 images, as used by the reference's `letterbox` (utils/imgproc_utils.py:86-117, called from preprocess_img,
inference.py:72-83) and by the mask back-projection (inference.py:164-168).

OpenCV 4.x (modules/imgproc/src/resize.cpp) is an un-vendored binary dependency here, so this restates its published
8-bit bilinear algorithm and is PINNED against the installed cv2 itself (tests/test_cpu_resize.py):
  * source coordinate fx = (dx + 0.5) * scale - 0.5 computed in float32 from the double scale = src / dst,
    sx = floor(fx), weights 1-fx / fx; COLUMN taps left of the image clamp to (pixel 0, weight 1), taps at or beyond
    the last column use the last pixel with weight 1; ROW taps keep their fractional weights and only the row
    indices are clipped;
  * weights are quantised to 11 bits: cvRound(w * 2048) as int16 (round-half-to-even);
  * horizontal pass in int32: S = a0*p[sx] + a1*p[sx+1];
  * vertical pass: dst = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
  * an exact 2x2 decimation (dst*2 == src in both axes) is routed to INTER_AREA (mean of the 2x2 block, +2 >> 2).
"""
import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def _axis_taps(ssize, dsize, clamp_weights):
    """-> (first tap index, second tap index, weights int16 [dsize, 2]).
    Columns (clamp_weights=True): a tap left of the image becomes (pixel 0, weight 1), a tap at or beyond the last
    column becomes (last pixel, weight 1) -- resize.cpp's xofs/alpha loop.
    Rows (clamp_weights=False): the weights keep their fractional values and only the ROW INDICES are clipped
    (resizeGeneric_Invoker: `clip(sy0 + k, 0, ssize.height)`), so a border row is blended with itself through two
    separately truncated products."""
    scale = np.float64(ssize) / np.float64(dsize)
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= ssize - 1
        f[hi] = 0.0
        s[hi] = ssize - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int16)   # cvRound = round half to even
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int16)
    i0 = np.clip(s, 0, ssize - 1)
    i1 = np.clip(s + 1, 0, ssize - 1)
    return i0, i1, np.stack([w0, w1], 1)


def resize_linear_u8(src, dsize_wh):
    """src: uint8 [H,W] or [H,W,C]; dsize_wh = (width, height) like cv2.resize."""
    src = np.asarray(src)
    assert src.dtype == np.uint8
    squeeze = src.ndim == 2
    if squeeze:
        src = src[:, :, None]
    sh, sw, _c = src.shape
    dw, dh = int(dsize_wh[0]), int(dsize_wh[1])
    if dw == sw and dh == sh:
        out = src.copy()
    elif dw * 2 == sw and dh * 2 == sh:
        # resize.cpp: INTER_LINEAR with an integer 2x2 decimation is computed as INTER_AREA
        s = src.astype(np.int32)
        out = ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    else:
        x0, x1, xa = _axis_taps(sw, dw, True)
        y0, y1, ya = _axis_taps(sh, dh, False)
        s = src.astype(np.int32)
        hrow = s[:, x0, :] * xa[None, :, 0, None].astype(np.int32) + s[:, x1, :] * xa[None, :, 1, None].astype(np.int32)
        s0, s1 = hrow[y0], hrow[y1]
        b0 = ya[:, 0].astype(np.int32)[:, None, None]
        b1 = ya[:, 1].astype(np.int32)[:, None, None]
        out = ((((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)
    return out[:, :, 0] if squeeze else out
