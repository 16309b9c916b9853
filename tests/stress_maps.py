"""Probability maps fed directly at the stage boundary (SURVEY 8d iii)."""
import cv2
import numpy as np


def blobs(seed, h=512, w=512, n=40, holes=True):
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), np.float32)
    for _ in range(n):
        c = (int(rng.integers(0, w)), int(rng.integers(0, h)))
        ax = (int(rng.integers(3, 70)), int(rng.integers(3, 30)))
        cv2.ellipse(m, c, ax, float(rng.uniform(0, 180)), 0, 360, float(rng.uniform(0.5, 1.0)), -1)
    for _ in range(n // 2):
        x0, y0 = int(rng.integers(0, w - 60)), int(rng.integers(0, h - 30))
        cv2.rectangle(m, (x0, y0), (x0 + int(rng.integers(4, 60)), y0 + int(rng.integers(3, 30))), float(rng.uniform(0.5, 1.0)), -1)
    if holes:
        for _ in range(n // 2):
            c = (int(rng.integers(0, w)), int(rng.integers(0, h)))
            cv2.circle(m, c, int(rng.integers(1, 8)), 0.0, -1)
    m = cv2.GaussianBlur(m, (0, 0), 1.2)
    return np.ascontiguousarray(m + rng.uniform(-0.02, 0.02, m.shape).astype(np.float32))


def nested(seed, h=256, w=256):
    """rings inside rings: holes containing islands containing holes"""
    m = np.zeros((h, w), np.float32)
    for k, r in enumerate(range(110, 5, -14)):
        cv2.circle(m, (w // 2, h // 2), r, 0.9 if k % 2 == 0 else 0.05, -1)
    cv2.rectangle(m, (5, 5), (60, 40), 0.8, 3)
    cv2.rectangle(m, (20, 15), (40, 30), 0.7, -1)
    return m


def checkerboard(h=128, w=128, cell=1):
    yy, xx = np.indices((h, w))
    return ((((yy // cell) + (xx // cell)) % 2) * 0.9).astype(np.float32)


def tiny_components(h=96, w=96):
    m = np.zeros((h, w), np.float32)
    m[10, 10] = 1      # 1 px
    m[20, 20:22] = 1   # 2 px
    m[30:33, 30:33] = 1
    m[40:42, 40:50] = 1
    m[60, 5:90] = 1    # 1-px line
    m[0, 0:7] = 1      # touching the frame
    m[70:96, 80:96] = 1
    m[80:85, 85:90] = 0  # hole at the frame corner block
    return m


CASES = {
    "blobs0": lambda: blobs(0), "blobs1": lambda: blobs(1, 384, 640, 60), "blobs_noholes": lambda: blobs(2, holes=False),
    "nested": lambda: nested(0), "checker1": lambda: checkerboard(96, 96, 1), "checker3": lambda: checkerboard(128, 128, 3),
    "tiny": tiny_components, "empty": lambda: np.zeros((64, 64), np.float32), "full": lambda: np.ones((64, 96), np.float32),
}
