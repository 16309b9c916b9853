"""-m gpu: refine_mask (utils/textmask.py:159-169) on the GPU, bit-exact against the oracle restatement
(which equals the unmodified reference except for the documented stable tie order of np.argsort).  Both device
implementations are tested: the phase-synchronous kernels of csrc/refine_mk.cu (default) and the cooperative
one-CTA / one-cluster-per-window kernels of csrc/refine.cu (CTD_REFINE=coop)."""
import os

import cv2
import numpy as np
import pytest

import ctd_b200
from ctd_b200 import compiler as cc
from oracle import postproc_ref, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    P = cc.Program()
    P.nc = 2
    P.newbuf(8, 1)
    e = ctd_b200.Engine(P, max_batch=1, max_h=1024, max_w=1024, skip_postproc=True)
    yield e
    e.close()


def make_case(seed, size=512, nblk=10):
    rng = np.random.default_rng(seed)
    img = synth.structured_page(2000 + seed, size, size)
    m = np.zeros((size, size), np.float32)
    wins = []
    for _ in range(nblk):
        x0, y0 = int(rng.integers(0, size - 112)), int(rng.integers(0, size - 92))
        w, h = int(rng.integers(30, 110)), int(rng.integers(20, 90))
        txt = "Ab%d" % rng.integers(0, 99)
        cv2.putText(m, txt, (x0 + 4, y0 + h - 6), cv2.FONT_HERSHEY_SIMPLEX, h / 40, 1.0, 3)
        cv2.putText(img, txt, (x0 + 4, y0 + h - 6), cv2.FONT_HERSHEY_SIMPLEX, h / 40, (10, 10, 10), 2)
        wins.append([x0, y0, min(size - 1, x0 + w), min(size - 1, y0 + h)])
    mask = (cv2.GaussianBlur(m, (0, 0), 1.5) * 255).clip(0, 255).astype(np.uint8)
    return img, mask, wins


@pytest.fixture(params=["mk", "coop"])
def impl(request):
    old = os.environ.get("CTD_REFINE")
    os.environ["CTD_REFINE"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("CTD_REFINE", None)
    else:
        os.environ["CTD_REFINE"] = old


@pytest.mark.parametrize("mode", [0, 1], ids=["inpaint", "annotation"])
@pytest.mark.parametrize("seed", range(6))
def test_refine_mask_matches_oracle(eng, impl, seed, mode):
    img, mask, wins = make_case(seed)
    ref = postproc_ref.refine_mask(img, mask.copy(), wins, mode)
    ex = [postproc_ref.expand_textwindow(img.shape, w, expand_r=16) for w in wins]
    got = eng.refine_mask(img, mask, ex, mode)
    assert np.array_equal(got, ref), int((got != ref).sum())


def test_refine_mask_edge_cases(eng, impl):
    img, mask, _ = make_case(3, 256, 4)
    # window covering the whole page, an empty-mask window, a 1-pixel-high window, no windows at all
    wins = [[0, 0, 255, 255], [200, 200, 240, 240], [10, 10, 60, 11]]
    mask[190:256, 190:256] = 0
    for mode in (0, 1):
        ref = postproc_ref.refine_mask(img, mask.copy(), wins, mode)
        ex = [postproc_ref.expand_textwindow(img.shape, w, expand_r=16) for w in wins]
        assert np.array_equal(eng.refine_mask(img, mask, ex, mode), ref)
    assert not eng.refine_mask(img, mask, np.zeros((0, 4), np.int32), 0).any()


def test_refine_large_windows_both_implementations_agree(eng):
    """a 1024x1024 page with overlapping page-sized windows (cluster kernel / many chunks per window): the two device
    implementations must agree bit for bit (the oracle takes minutes at this size)"""
    img, mask, wins = make_case(11, 1024, 30)
    wins += [[0, 0, 1023, 1023], [100, 50, 900, 1000], [0, 300, 1023, 700]]
    ex = [postproc_ref.expand_textwindow(img.shape, w, expand_r=16) for w in wins]
    out = {}
    for name in ("mk", "coop"):
        os.environ["CTD_REFINE"] = name
        out[name] = eng.refine_mask(img, mask, ex, 0)
    os.environ.pop("CTD_REFINE", None)
    assert out["mk"].any() and np.array_equal(out["mk"], out["coop"]), int((out["mk"] != out["coop"]).sum())
