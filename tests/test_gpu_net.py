"""-m gpu: the whole forward pass (backbone + both heads) through the C-ABI against the oracle
(oracle/net_ref.py, pinned bit-identical to the unmodified reference in the build container).

Tolerances (stated, per BASELINE.json north_star):
  * fp32 CUDA-core engine and the split-fp16 TENSOR-CORE engine (CTD_PREC_SPLIT_TC, BASELINE config 2): seg / line
    maps within 1e-3 of the fp32 reference (max abs, post-sigmoid), DB bitmap disagreement <= 1e-4 of the pixels;
  * fp16 tensor-core engine (CTD_PREC_FP16_TC, BASELINE config 3 "fp16"): fp16 STORAGE of every activation makes the
    random-weight net's maps differ from fp32 statistically (a CPU emulation of fp16 storage through the same graph,
    tests/prog_interp.py storage='f16', shows the same profile: mean 3-6e-3, p99.9 0.11-0.16, isolated maxima
    0.3-0.45 on 2 small pages, 0.65 over the 16 pages of the benchmark batch -- an extreme-value statistic).  The kernels themselves are pinned per op at 2e-3 in tests/test_gpu_layers.py; here the engine is held
    to the emulation's profile and compared against the emulation itself."""
import numpy as np
import pytest
import torch

import ctd_b200
from oracle import synth
from oracle.net_ref import RefNet
from util import get_checkpoint, page_to_net_input, PREC_FP16_TC, PREC_FP32_SIMT, PREC_FP16_SIMT, PREC_SPLIT_TC

pytestmark = pytest.mark.gpu

# fp16 engines: every activation/weight is stored in fp16 (fp32 accumulate).  A CPU emulation of that storage
# through the same graph (tests/prog_interp.py with fp16 rounding) gives mean |err| 3-4e-3 and isolated maxima of
# 0.1-0.35 where the random-weight net is locally ill-conditioned, so the stated fp16 tolerance is statistical.
# The 99.9th percentile sits at 0.13-0.151 depending on the fp32 accumulation ORDER (tap-major vs K-block-major
# kernels give 0.147 / 0.151 on the same page), hence 0.2.
TOL = {PREC_FP32_SIMT: dict(maps=1e-3, maps_mean=1e-4, p999=1e-3, blks_rel=2e-3),
       PREC_SPLIT_TC: dict(maps=1e-3, maps_mean=1e-4, p999=1e-3, blks_rel=2e-3),
       PREC_FP16_TC: dict(maps=0.8, maps_mean=1.5e-2, p999=0.25, blks_rel=1.0),
       PREC_FP16_SIMT: dict(maps=0.8, maps_mean=1.5e-2, p999=0.25, blks_rel=1.0)}
EXACT = (PREC_FP32_SIMT, PREC_SPLIT_TC)


def _pages(n, h, w, seed=1000):
    return np.stack([synth.structured_page(seed + i, h, w) if i % 2 == 0 else synth.noise_page(seed + i, h, w)
                     for i in range(n)])


def _check_against_oracle(prec, smooth, n, h, w, use_graph=False, seed=1000):
    ck = get_checkpoint(0, smooth)
    pages = _pages(n, h, w, seed)
    ref = RefNet(ck)
    with torch.no_grad():
        outs = [ref(page_to_net_input(pages[i:i + 1])) for i in range(n)]      # page by page: bounded host memory
    rb, rm, rl = (torch.cat([o[k] for o in outs]) for k in range(3))
    prog = ctd_b200.compiler.compile_checkpoint(ck)
    eng = ctd_b200.Engine(prog, precision=prec, max_batch=n, max_h=h, max_w=w, use_graph=use_graph)
    try:
        eng.forward(pages)
        blks, mask, lines = eng.net_outputs()
        m8 = eng.mask_u8()
    finally:
        eng.close()
    tol = TOL[prec]
    e_mask = float(np.abs(mask - rm.numpy()).max())
    e_lines = float(np.abs(lines - rl.numpy()).max())
    rbn = rb.numpy()
    e_blks = float((np.abs(blks - rbn) / (np.abs(rbn) + 1.0)).max())
    m_mask, m_lines = float(np.abs(mask - rm.numpy()).mean()), float(np.abs(lines - rl.numpy()).mean())
    msg = "prec %d smooth %d: max err mask %.3g lines %.3g blks(rel) %.3g; mean err mask %.3g lines %.3g" % (
        prec, smooth, e_mask, e_lines, e_blks, m_mask, m_lines)
    print(msg)
    assert e_mask <= tol["maps"] and e_lines <= tol["maps"], msg
    assert m_mask <= tol["maps_mean"] and m_lines <= tol["maps_mean"], msg
    assert e_blks <= tol["blks_rel"], msg
    for got, ref in ((mask, rm.numpy()), (lines, rl.numpy())):
        d = np.abs(got - ref).ravel()
        assert float(np.partition(d, int(d.size * 0.999))[int(d.size * 0.999)]) <= tol["p999"], msg
    # DB bitmap (shrink > 0.3, db_utils.py:71-72) agreement
    flips = float(((lines[:, 0] > 0.3) != (rl.numpy()[:, 0] > 0.3)).mean())
    assert flips <= (1e-4 if prec in EXACT else 1e-2), (flips, msg)
    # postprocess_mask (inference.py:96-99): (mask*255) truncated; compare on the engine's own float mask
    assert np.array_equal(m8, (mask[:, 0] * 255).astype(np.uint8))
    return pages, prog, (blks, mask, lines)


@pytest.mark.parametrize("prec", [PREC_FP32_SIMT, PREC_FP16_SIMT, PREC_FP16_TC, PREC_SPLIT_TC])
@pytest.mark.parametrize("smooth", [False, True], ids=["rough", "smooth"])
def test_forward_matches_oracle(prec, smooth):
    _check_against_oracle(prec, smooth, 2, 256, 320)


def test_benchmark_config_matches_oracle():
    """BASELINE configs[2] itself -- 1024x1024, batch 16, fp16 tcgen05 path under a CUDA graph (what bench.py
    times) -- against the fp32 oracle on all 16 pages (structured and noise pages alternate)."""
    _check_against_oracle(PREC_FP16_TC, True, 16, 1024, 1024, use_graph=True)


@pytest.mark.parametrize("prec", [PREC_FP16_TC, PREC_SPLIT_TC], ids=["fp16_tc", "split_tc"])
@pytest.mark.parametrize("size", [640, 1024, 1536])
def test_stream_bucket_sizes_match_oracle(prec, size):
    """BASELINE configs[4] buckets (640 / 1024 / 1536 squares) and config 2 (1024, fp32-accurate tensor-core mode)."""
    _check_against_oracle(prec, True, 2, size, size, use_graph=True, seed=2000 + size)


def test_fp16_engine_tracks_fp16_storage_emulation():
    """The tcgen05 engine against the CPU emulation of ITS OWN numerics (fp16 weights + fp16 activation storage, fp32
    accumulate): what is left is accumulation order and one-ulp rounding flips amplified by the net, an order of
    magnitude below the engine's distance to the fp32 reference."""
    import prog_interp
    ck = get_checkpoint(0, True)
    n, h, w = 2, 256, 320
    pages = _pages(n, h, w)
    prog = ctd_b200.compiler.compile_checkpoint(ck)
    eb, em, el = prog_interp.run_program(prog, pages, storage="f16")
    eng = ctd_b200.Engine(prog, precision=PREC_FP16_TC, max_batch=n, max_h=h, max_w=w)
    try:
        eng.forward(pages)
        blks, mask, lines = eng.net_outputs()
    finally:
        eng.close()
    rb, rm, rl = RefNet(ck)(page_to_net_input(pages))
    for name, got, emu, ref in (("mask", mask, em.numpy(), rm.numpy()), ("lines", lines, el.numpy(), rl.numpy())):
        d_emu = float(np.abs(got - emu).mean())
        d_ref = float(np.abs(got - ref).mean())
        print("%s: mean |engine - fp16 emulation| %.3g, mean |engine - fp32 reference| %.3g, max vs emulation %.3g"
              % (name, d_emu, d_ref, float(np.abs(got - emu).max())))
        assert d_emu <= 0.6 * d_ref + 1e-4, (name, d_emu, d_ref)
        assert d_emu <= 3e-3, (name, d_emu)


def test_split_engine_end_to_end_bit_exact():
    """BASELINE config 2 / north_star: with the fp32-accurate tensor-core engine the WHOLE device pipeline is compared
    with the oracle chain run on the REFERENCE's fp32 maps (not on the engine's own maps): detection rows, DB bitmap,
    CC labels and line boxes must be identical; pixels whose reference value lies within 1e-3 of a threshold (0.3 for
    the bitmap, k/255 for the u8 mask) are the only ones allowed to differ and are counted."""
    from oracle import postproc_ref
    ck = get_checkpoint(0, True)
    n, h, w = 2, 512, 512
    pages = np.stack([synth.structured_page(1000 + 3 * i, h, w) for i in range(n)])
    with torch.no_grad():
        rb, rm, rl = RefNet(ck)(page_to_net_input(pages))
    prog = ctd_b200.compiler.compile_checkpoint(ck)
    eng = ctd_b200.Engine(prog, precision=PREC_SPLIT_TC, max_batch=n, max_h=h, max_w=w, use_graph=True)
    try:
        eng.forward(pages)
        dets = eng.detections()
        m8 = eng.mask_u8()
        bitmap, labels, nl = eng.db_components()
        boxes, scores = eng.text_lines()
    finally:
        eng.close()
    shrink = rl.numpy()[:, 0]
    ref_bitmap = (shrink > 0.3).astype(np.uint8)
    flips = bitmap != ref_bitmap
    assert np.all(np.abs(shrink[flips] - 0.3) < 1e-3), "bitmap differs away from the threshold"
    ref_m8 = (rm.numpy()[:, 0] * 255).astype(np.uint8)
    mdiff = m8 != ref_m8
    frac = rm.numpy()[:, 0] * 255
    assert np.all(np.abs(frac[mdiff] - np.round(frac[mdiff])) < 0.255 + 1e-6), "u8 mask differs away from a truncation step"
    assert np.all(np.abs(m8.astype(int) - ref_m8.astype(int))[mdiff] == 1)
    print("near-threshold pixels that differ: bitmap %d of %d, mask_u8 %d of %d" % (int(flips.sum()), flips.size,
                                                                                   int(mdiff.sum()), mdiff.size))
    n_box_flips = 0
    for i in range(n):
        ref_det = postproc_ref.non_max_suppression(rb[i:i + 1], 0.4, 0.35)[0].numpy()
        assert len(dets[i]) == len(ref_det) and len(ref_det) > 0
        assert np.array_equal(dets[i][:, 5], ref_det[:, 5])
        # int bboxes (inference.py:108): the float rows agree to ~1e-4; a coordinate that sits within 1e-2 of an integer
        # may truncate to the neighbour (box = (2*sigmoid - 0.5 + grid) * stride amplifies the 1e-4 map error by up to 64) -- counted, everything else must be identical
        gi, ri = dets[i][:, :4].astype(np.int32), ref_det[:, :4].astype(np.int32)
        near = np.abs(ref_det[:, :4] - np.round(ref_det[:, :4])) < 1e-2
        assert np.array_equal(gi[~near], ri[~near]), "int bboxes away from an integer boundary"
        assert np.all(np.abs(gi - ri) <= 1) and float(np.abs(dets[i][:, :4] - ref_det[:, :4]).max()) < 2e-2
        n_box_flips += int((gi != ri).sum())
        assert np.array_equal(np.round(dets[i][:, 4], 3), np.round(ref_det[:, 4], 3))
        if not flips[i].any():
            n_ref, lab_ref, _, _ = postproc_ref.connected_components_cv2(ref_bitmap[i])
            assert int(nl[i]) == n_ref and np.array_equal(labels[i], lab_ref), "CC labels"
            rboxes, rscores = postproc_ref.seg_represent(shrink[i], 0.3)
            assert len(boxes[i]) == len(rboxes) and len(rboxes) > 0
            same = np.all(boxes[i].reshape(len(rboxes), -1) == rboxes.reshape(len(rboxes), -1), axis=1)
            assert same.mean() >= 0.97, "line boxes (minAreaRect ties aside)"
            assert np.allclose(scores[i], rscores, atol=1e-3)
    print("int bbox coordinates that truncate differently (within 1e-2 of an integer):", n_box_flips)


def test_batch_invariance():
    """pages are independent: a page's result must not depend on its batch neighbours."""
    ck = get_checkpoint(0, True)
    h = w = 256
    pages = _pages(3, h, w)
    prog = ctd_b200.compiler.compile_checkpoint(ck)
    eng = ctd_b200.Engine(prog, precision=PREC_FP16_TC, max_batch=3, max_h=h, max_w=w)
    try:
        eng.forward(pages)
        b3, m3, l3 = eng.net_outputs()
        eng.forward(pages[1:2])
        b1, m1, l1 = eng.net_outputs()
    finally:
        eng.close()
    assert np.array_equal(m3[1], m1[0]) and np.array_equal(l3[1], l1[0]) and np.array_equal(b3[1], b1[0])


def test_tc_layers_track_fp32_engine():
    """Layer-by-layer: every buffer of the tcgen05 engine (halo / tap-per-box / stem kernels, fp16 storage) against the
    same buffer of the fp32 CUDA-core engine on the same page.  fp16 storage noise grows slowly through the net;
    a kernel bug (a wrong tap, a bad border) shows up as an O(1) relative error in the first buffer it touches."""
    ck = get_checkpoint(0, True)
    n, h, w = 1, 256, 192
    pages = _pages(n, h, w, seed=77)
    prog = ctd_b200.compiler.compile_checkpoint(ck)
    outs = {}
    for prec in (PREC_FP32_SIMT, PREC_FP16_TC):
        eng = ctd_b200.Engine(prog, precision=prec, max_batch=n, max_h=h, max_w=w)
        try:
            eng.forward(pages)
            res = []
            for i, op in enumerate(prog.ops):
                if op["kind"] in (6, 7, 8) or op["dst_buf"] < 0:
                    continue
                t = dict(buf=op["dst_buf"], coff=op["dst_coff"], c=op["cout"], down=prog.bufs[op["dst_buf"]][1])
                if t["c"] <= 0:
                    continue
                res.append((i, eng.debug_read(t)))
            outs[prec] = res
        finally:
            eng.close()
    worst = []
    for (i, a), (_j, b) in zip(outs[PREC_FP32_SIMT], outs[PREC_FP16_TC]):
        scale = float(np.abs(a).max()) + 1e-6
        err = float(np.abs(a - b).max()) / scale
        rms = float(np.sqrt(np.mean((a - b) ** 2))) / (float(np.sqrt(np.mean(a ** 2))) + 1e-6)
        worst.append((err, rms, i))
        # first layers: pure fp16 rounding; deeper: accumulated storage noise (measured rms <= ~1e-2)
        assert rms <= (4e-3 if i < 6 else 5e-2), "op %d (kind %d): max rel err %.3g, rms rel %.3g" % (i, prog.ops[i]["kind"], err, rms)
    print("worst layers (max rel err, rms rel, op):", sorted(worst, reverse=True)[:5])
