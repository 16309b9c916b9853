"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's post-processing stages
that the GPU kernels replace, each citing the reference lines it follows.  cv2 / torchvision are
the same library builds on the GPU box (same image), so they are used directly where the
reference itself calls them.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs import this.
"""
import numpy as np
import torch
import torchvision


def xywh2xyxy(x):
    """utils/yolov5_utils.py:220-227"""
    y = x.clone()
    y[:, 0] = x[:, 0] - x[:, 2] / 2
    y[:, 1] = x[:, 1] - x[:, 3] / 2
    y[:, 2] = x[:, 0] + x[:, 2] / 2
    y[:, 3] = x[:, 1] + x[:, 3] / 2
    return y


def non_max_suppression(prediction, conf_thres=0.4, iou_thres=0.35, max_det=300, max_wh=4096, max_nms=30000):
    """utils/yolov5_utils.py:124-218 with its defaults on the inference path (single label,
    class-aware, no merge).  prediction: (N, A, 5+nc) float32 tensor -> list of (n,6) tensors."""
    prediction = torch.as_tensor(prediction).clone()
    xc = prediction[..., 4] > conf_thres  # :136
    out = [torch.zeros((0, 6))] * prediction.shape[0]
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]  # :152
        if not x.shape[0]:
            continue
        x[:, 5:] *= x[:, 4:5]  # :169
        box = xywh2xyxy(x[:, :4])  # :172
        conf, j = x[:, 5:].max(1, keepdim=True)  # :179
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]  # :180
        n = x.shape[0]
        if not n:
            continue
        elif n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]  # :194-195
        c = x[:, 5:6] * max_wh  # :198
        boxes, scores = x[:, :4] + c, x[:, 4]
        i = torchvision.ops.nms(boxes, scores, iou_thres)  # :200
        if i.shape[0] > max_det:
            i = i[:max_det]
        out[xi] = x[i]
    return out


def connected_components_cv2(img_u8):
    """The call the reference effectively makes (utils/textmask.py:93,113,138):
    cv2.connectedComponentsWithStats(img) with the defaults connectivity=8, ltype=CV_32S."""
    import cv2
    n, labels, stats, centroids = cv2.connectedComponentsWithStats(img_u8)
    return n, labels, stats, centroids


# ---------------------------------------------------------------------------------------------------
# SegDetectorRepresenter (utils/db_utils.py:32-211), non-polygon path used by inference.py:158


def _get_mini_boxes(contour):
    """db_utils.py:176-195"""
    import cv2
    bounding_box = cv2.minAreaRect(contour)
    points = sorted(list(cv2.boxPoints(bounding_box)), key=lambda x: x[0])
    if points[1][1] > points[0][1]:
        i1, i4 = 0, 1
    else:
        i1, i4 = 1, 0
    if points[3][1] > points[2][1]:
        i2, i3 = 2, 3
    else:
        i2, i3 = 3, 2
    return [points[i1], points[i2], points[i3], points[i4]], min(bounding_box[1])


def _box_score_fast(bitmap, _box):
    """db_utils.py:197-211"""
    import cv2
    h, w = bitmap.shape[:2]
    box = _box.copy()
    xmin = np.clip(np.floor(box[:, 0].min()).astype(np.int64), 0, w - 1)
    xmax = np.clip(np.ceil(box[:, 0].max()).astype(np.int64), 0, w - 1)
    ymin = np.clip(np.floor(box[:, 1].min()).astype(np.int64), 0, h - 1)
    ymax = np.clip(np.ceil(box[:, 1].max()).astype(np.int64), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    box[:, 0] = box[:, 0] - xmin
    box[:, 1] = box[:, 1] - ymin
    cv2.fillPoly(mask, box.reshape(1, -1, 2).astype(np.int32), 1)
    return cv2.mean(bitmap[ymin:ymax + 1, xmin:xmax + 1], mask)[0]


def _unclip(box, unclip_ratio=1.5):
    """db_utils.py:168-174 with the third-party calls restated in oracle/geom_ref.py (parity unpinned)."""
    from oracle import geom_ref
    distance = geom_ref.geos_ring_area(box) * unclip_ratio / geom_ref.geos_ring_length(box)
    pts = geom_ref.clipper_offset_closed_polygon(np.asarray(box).tolist(), distance)
    return np.array([pts])


def seg_represent(pred, thresh=0.3, max_candidates=1000, unclip_ratio=1.5):
    """SegDetectorRepresenter(thresh)(batch, pred[None,None]) for ONE map: (boxes int16 [k,4,2], scores f32 [k])
    -- db_utils.py:40-72 (binarize) and 123-166 (boxes_from_bitmap)."""
    import cv2
    pred = np.asarray(pred, np.float32)
    bitmap = pred > thresh
    height, width = bitmap.shape
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    n = min(len(contours), max_candidates)
    boxes = np.zeros((n, 4, 2), dtype=np.int16)
    scores = np.zeros((n,), dtype=np.float32)
    for index in range(n):
        contour = contours[index].squeeze(1)
        points, sside = _get_mini_boxes(contour)
        if sside < 2:
            continue
        points = np.array(points)
        score = _box_score_fast(pred, contour)
        box = _unclip(points, unclip_ratio).reshape(-1, 1, 2)
        box, sside = _get_mini_boxes(box)
        box = np.array(box)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * width), 0, width)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * height), 0, height)
        boxes[index, :, :] = box.astype(np.int16)
        scores[index] = score
    return boxes, scores


# ---------------------------------------------------------------------------------------------------
# refine_mask (utils/textmask.py:159-169) and callees, restated with the same cv2/numpy calls.
# ONE deliberate normalisation: `np.argsort(bins * -1)` in get_topk_color (textmask.py:17) is an unstable
# sort whose tie order depends on numpy's SIMD sort build (SURVEY 8c "known non-determinism"); here and in
# the CUDA path ties are broken by ascending bin index (kind="stable").
REFINEMASK_INPAINT, REFINEMASK_ANNOTATION = 0, 1


def expand_textwindow(img_size, xyxy, expand_r=8):
    """utils/imgproc_utils.py:151-161"""
    im_h, im_w = img_size[:2]
    x1, y1, x2, y2 = xyxy
    w, h = x2 - x1, y2 - y1
    paddings = int(round((max(h, w) * 0.25 + min(h, w) * 0.75) / expand_r))
    x1, y1 = max(0, x1 - paddings), max(0, y1 - paddings)
    x2, y2 = min(im_w - 1, x2 + paddings), min(im_h - 1, y2 + paddings)
    return [x1, y1, x2, y2]


def _topk_color(color_list, bins, k=3, color_var=10, bin_tol=0.001, tie_order="stable"):
    """textmask.py:16-27.  tie_order='stable': ties of the histogram sort broken by ascending bin index (the documented
    normalisation the CUDA kernel follows); 'numpy': `np.argsort(bins * -1)` exactly as the reference calls it (unstable,
    its tie order depends on numpy's SIMD build) -- used to pin the oracle against the unmodified reference."""
    idx = np.argsort(bins * -1, kind="stable") if tie_order == "stable" else np.argsort(bins * -1)
    color_list, bins = color_list[idx], bins[idx]
    top = [color_list[0]]
    tol = np.sum(bins) * bin_tol
    if len(color_list) > 1:
        for color, b in zip(color_list[1:], bins[1:]):
            if np.abs(np.array(top) - color).min() > color_var:
                top.append(color)
            if len(top) >= k or b < tol:
                break
    return top


def _minxor(threshed, mask):
    """textmask.py:29-41 (dilate=False)"""
    import cv2
    neg = 255 - threshed
    neg_sum = cv2.bitwise_xor(neg, mask).sum()
    pos_sum = cv2.bitwise_xor(threshed, mask).sum()
    return (neg, neg_sum) if neg_sum < pos_sum else (threshed, pos_sum)


def candidate_masks(im, msk, tie_order="stable"):
    """get_topk_masklist + get_otsuthresh_masklist (textmask.py:43-71) -> list of [mask, xor_sum]"""
    import cv2
    grey = cv2.cvtColor(im, cv2.COLOR_BGR2GRAY)
    msk = np.ascontiguousarray(msk)
    px = grey[np.where(cv2.erode(msk, np.ones((3, 3), np.uint8), iterations=1) > 127)]
    counts, edges = np.histogram(px, bins=255)
    out = []
    for color in _topk_color(edges, counts, color_var=10, k=3, tie_order=tie_order):
        c_top = min(color + 30, 255)
        c_bottom = c_top - 60
        t, s = _minxor(cv2.inRange(grey, c_bottom, c_top), msk)
        out.append([t, s])
    per_ch = []
    for c in (im[..., 0], im[..., 1], im[..., 2]):
        _, t = cv2.threshold(c, 1, 255, cv2.THRESH_OTSU + cv2.THRESH_BINARY)
        t, s = _minxor(t, msk)
        per_ch.append([t, s])
    per_ch.sort(key=lambda x: x[1])
    return out + [per_ch[0]]


def merge_masks(mask_list, pred_mask, refine_mode=REFINEMASK_INPAINT):
    """merge_mask_list (textmask.py:73-132) with the defaults refine_mask uses (no line filter)."""
    import cv2
    mask_list.sort(key=lambda x: x[1])
    element = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3), (1, 1))
    pred = cv2.erode(pred_mask, element, iterations=1)
    _, pred = cv2.threshold(pred, 60, 255, cv2.THRESH_BINARY)
    merged = np.zeros_like(pred)

    def try_labels(labels, stats, n, skip_bg, area_thresh=None):
        for li in range(n):
            if skip_bg and li == 0:
                continue
            x, y, w, h, area = stats[li]
            if area_thresh is None:
                if w * h < 3:
                    continue
            elif not area < area_thresh:
                continue
            sl = (slice(y, y + h), slice(x, x + w))
            tmp = np.zeros((h, w), np.uint8)
            tmp[labels[sl] == li] = 255
            tmp = cv2.bitwise_or(merged[sl], tmp)
            if cv2.bitwise_xor(tmp, pred[sl]).sum() < cv2.bitwise_xor(merged[sl], pred[sl]).sum():
                merged[sl] = tmp

    for cand, _ in mask_list:
        n, labels, stats, _c = cv2.connectedComponentsWithStats(cand, 8, cv2.CV_16U)  # effectively (cand): 8-conn, CV_32S
        try_labels(labels, stats, n, True)
    if refine_mode == REFINEMASK_INPAINT:
        merged[...] = cv2.dilate(merged, np.ones((3, 3), np.uint8), iterations=1)
    n, labels, stats, _c = cv2.connectedComponentsWithStats(255 - merged, 8, cv2.CV_16U)
    areas = np.sort(stats[:, -1])
    try_labels(labels, stats, n, False, areas[-2] if len(areas) > 1 else areas[-1])
    return merged


def refine_mask(img, pred_mask, windows_xyxy, refine_mode=REFINEMASK_INPAINT, tie_order="stable"):
    """refine_mask (textmask.py:159-169); `windows_xyxy` = [blk.xyxy for blk in blk_list]."""
    import cv2
    out = np.zeros_like(pred_mask)
    for xyxy in windows_xyxy:
        bx1, by1, bx2, by2 = expand_textwindow(img.shape, xyxy, expand_r=16)
        im = np.ascontiguousarray(img[by1:by2, bx1:bx2])
        msk = np.ascontiguousarray(pred_mask[by1:by2, bx1:bx2])
        merged = merge_masks(candidate_masks(im, msk, tie_order), msk, refine_mode)
        out[by1:by2, bx1:bx2] = cv2.bitwise_or(out[by1:by2, bx1:bx2], merged)
    return out
