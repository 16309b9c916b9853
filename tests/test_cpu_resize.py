"""not-gpu: oracle/resize_ref.py (numpy restatement of OpenCV's 8-bit INTER_LINEAR resize) is pinned against the installed
cv2 itself -- the third-party binary the reference calls in `letterbox` (imgproc_utils.py:86-117) and for the mask
back-projection (inference.py:164-168)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle.resize_ref import resize_linear_u8

CASES = [((360, 250), (178, 256)), ((1654, 1170), (724, 1024)), ((1024, 724), (1170, 1654)), ((100, 100), (50, 50)),
         ((100, 100), (200, 200)), ((77, 33), (100, 211)), ((512, 512), (511, 513)), ((1024, 1024), (512, 512)),
         ((300, 500), (150, 250)), ((256, 178), (250, 360)), ((17, 5), (3, 9)), ((2, 2), (7, 5)), ((1, 9), (4, 4)),
         ((640, 480), (320, 480)), ((640, 480), (640, 240)), ((9, 1), (1, 30)), ((1, 1), (5, 5)), ((33, 40), (100, 80)),
         ((700, 500), (731, 1024)), ((64, 64), (256, 256)), ((500, 400), (1024, 819)), ((64, 64), (64, 64))]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%d_to_%dx%d" % (c[0][0], c[0][1], c[1][1], c[1][0]))
@pytest.mark.parametrize("channels", [1, 3])
def test_oracle_resize_equals_cv2(case, channels):
    (sh, sw), (dw, dh) = case
    rng = np.random.default_rng(sh * 131 + sw * 7 + dw)
    src = rng.integers(0, 256, (sh, sw, channels), dtype=np.uint8)
    if channels == 1:
        src = src[:, :, 0]
    ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    got = resize_linear_u8(src, (dw, dh))
    assert got.shape == ref.shape
    assert np.array_equal(got, ref), int((got != ref).sum())


def test_letterbox_geometry_matches_host_letterbox():
    import ctd_b200
    from ctd_b200 import inference
    for shape in [(1654, 1170), (700, 500), (360, 250), (1024, 1024), (512, 2048), (333, 777)]:
        img = np.zeros((shape[0], shape[1], 3), np.uint8)
        out, (r, _), (dw, dh) = inference.letterbox(img, (1024, 1024))
        r2, unpad, dw2, dh2 = inference.letterbox_geometry(shape, (1024, 1024))
        assert out.shape == (1024, 1024, 3) and (dw, dh) == (dw2, dh2) and r == r2
        assert unpad == (1024 - dw, 1024 - dh)


def test_oracle_resize_equals_committed_cv2_goldens():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resize_cv2.npz"))
    n = 0
    for k in g.files:
        if not k.startswith("src_"):
            continue
        want = g["dst_" + k[4:]]
        got = resize_linear_u8(g[k], (want.shape[1], want.shape[0]))
        assert np.array_equal(got, want), k
        n += 1
    assert n == 12
