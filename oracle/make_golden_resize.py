"""TEST INFRASTRUCTURE ONLY (oracle): golden vectors of `cv2.resize(.., INTER_LINEAR)` produced by the OpenCV installed in the
build container (the same binary dependency the reference calls), committed as tests/golden/resize_cv2.npz so that
the GPU box can check the resize kernel against OpenCV's own output even where cv2 is not importable.
    python -m oracle.make_golden_resize
"""
import os

import cv2
import numpy as np

CASES = [((37, 53), (20, 29)), ((37, 53), (111, 64)), ((64, 48), (32, 24)), ((5, 9), (31, 17)), ((90, 60), (64, 64)),
         ((33, 21), (21, 33))]


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    rng = np.random.default_rng(2024)
    for i, ((sh, sw), (dw, dh)) in enumerate(CASES):
        for c in (1, 3):
            src = rng.integers(0, 256, (sh, sw, c), dtype=np.uint8)
            if c == 1:
                src = src[:, :, 0]
            out["src_%d_%d" % (i, c)] = src
            out["dst_%d_%d" % (i, c)] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "resize_cv2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
