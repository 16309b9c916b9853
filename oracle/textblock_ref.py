"""TEST INFRASTRUCTURE ONLY (oracle): python/numpy restatement of the reference's line -> block grouping
(`group_output`, utils/textblock.py:421-508, and its callees 302-419, 267-300; `TextBlock` fields 12-108;
`union_area` / `xywh2xyxypoly`, utils/imgproc_utils.py:13-37) with the same numpy calls in the same order.
tests/test_cpu_textblock.py pins it against the UNMODIFIED reference (24 random cases + committed goldens); the
product's implementation is the native `ctd_group_output` (comic-text-detector_b200/csrc/group.cpp), which the tests
compare against this file on the GPU box, where /root/reference does not exist.  shapely's `Polygon.intersects`
is replaced by an exact integer test (`quads_intersect`).  Nothing under comic-text-detector_b200/ imports this.
"""
import copy
import math
from typing import List

import numpy as np

LANG_LIST = ["eng", "ja", "unknown"]
LANGCLS2IDX = {"eng": 0, "ja": 1, "unknown": 2}


# ---------------------------------------------------------------------------------------------------
# exact closed-set intersection of two simple polygons with integer vertices (shapely intersects)
def _orient(a, b, c):
    v = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
    return (v > 0) - (v < 0)


def _between(a, b, p):
    return min(a[0], b[0]) <= p[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= p[1] <= max(a[1], b[1])


def _segments_touch(a, b, c, d):
    o1, o2, o3, o4 = _orient(a, b, c), _orient(a, b, d), _orient(c, d, a), _orient(c, d, b)
    if o1 != o2 and o3 != o4:
        return True
    return ((o1 == 0 and _between(a, b, c)) or (o2 == 0 and _between(a, b, d)) or
            (o3 == 0 and _between(c, d, a)) or (o4 == 0 and _between(c, d, b)))


def _inside_or_on(p, poly):
    n = len(poly)
    inside = False
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        if _orient(a, b, p) == 0 and _between(a, b, p):
            return True
        if (a[1] > p[1]) != (b[1] > p[1]):
            t = (b[0] - a[0]) * (p[1] - a[1]) - (p[0] - a[0]) * (b[1] - a[1])
            if t != 0 and ((t > 0) == (b[1] > a[1])):
                inside = not inside
    return inside


def quads_intersect(pa, pb) -> bool:
    pa = [(int(x), int(y)) for x, y in pa]
    pb = [(int(x), int(y)) for x, y in pb]
    # closed sets with disjoint bounding boxes cannot touch (exact, and most line pairs on a page are far apart)
    if (max(p[0] for p in pa) < min(p[0] for p in pb) or max(p[0] for p in pb) < min(p[0] for p in pa) or
            max(p[1] for p in pa) < min(p[1] for p in pb) or max(p[1] for p in pb) < min(p[1] for p in pa)):
        return False
    for i in range(len(pa)):
        for j in range(len(pb)):
            if _segments_touch(pa[i], pa[(i + 1) % len(pa)], pb[j], pb[(j + 1) % len(pb)]):
                return True
    return _inside_or_on(pa[0], pb) or _inside_or_on(pb[0], pa)


# ---------------------------------------------------------------------------------------------------
class TextBlock(object):
    """Result record with the reference's field names (textblock.py:12-85); the UI/OCR helpers of the
    reference class (min_rect, get_transformed_region, colours ...) are out of scope."""

    def __init__(self, xyxy, lines=None, language="unknown", vertical=False, font_size=-1, distance=None, angle=0,
                 vec=None, norm=-1, merged=False, weight=-1, text=None, translation="", fg_r=0, fg_g=0, fg_b=0,
                 bg_r=0, bg_g=0, bg_b=0, line_spacing=1., font_family="", bold=False, underline=False, italic=False,
                 alignment=-1, alpha=255, rich_text="", _bounding_rect=None, accumulate_color=True,
                 default_stroke_width=0.2, target_lang="", **kwargs):
        self.xyxy = [int(v) for v in xyxy]
        self.lines = [] if lines is None else lines
        self.vertical = vertical
        self.language = language
        self.font_size = font_size
        self.distance = None if distance is None else np.array(distance, np.float64)
        self.angle = angle
        self.vec = None if vec is None else np.array(vec, np.float64)
        self.norm = norm
        self.merged = merged
        self.weight = weight
        self.text = text if text is not None else []
        self.prob = 1
        self.translation = translation
        self.fg_r, self.fg_g, self.fg_b = fg_r, fg_g, fg_b
        self.bg_r, self.bg_g, self.bg_b = bg_r, bg_g, bg_b
        self.font_family = font_family
        self.bold, self.underline, self.italic = bold, underline, italic
        self.alpha = alpha
        self.rich_text = rich_text
        self.line_spacing = line_spacing
        self._alignment = alignment
        self._target_lang = target_lang
        self._bounding_rect = _bounding_rect
        self.default_stroke_width = default_stroke_width
        self.accumulate_color = accumulate_color

    def lines_array(self, dtype=np.float64):
        return np.array(self.lines, dtype=dtype)

    def adjust_bbox(self, with_bbox=False):
        pts = self.lines_array().astype(np.int32)
        lo_x, lo_y, hi_x, hi_y = pts[..., 0].min(), pts[..., 1].min(), pts[..., 0].max(), pts[..., 1].max()
        if with_bbox:
            self.xyxy = [min(lo_x, self.xyxy[0]), min(lo_y, self.xyxy[1]), max(hi_x, self.xyxy[2]), max(hi_y, self.xyxy[3])]
        else:
            self.xyxy = [lo_x, lo_y, hi_x, hi_y]

    def sort_lines(self):
        if self.distance is not None:
            order = np.argsort(self.distance)
            self.distance = self.distance[order]
            self.lines = np.array(self.lines, dtype=np.int32)[order].tolist()

    def __len__(self):
        return len(self.lines)

    def __getitem__(self, idx):
        return self.lines[idx]

    def to_dict(self):
        return copy.deepcopy(vars(self))


def overlap_area(a, b):
    """`union_area` of the reference (imgproc_utils.py:13-20): the INTERSECTION area, -1 when disjoint."""
    x1, y1 = max(a[0], b[0]), max(a[1], b[1])
    x2, y2 = min(a[2], b[2]), min(a[3], b[3])
    if y2 < y1 or x2 < x1:
        return -1
    return (y2 - y1) * (x2 - x1)


def _orientation_and_metrics(blk: TextBlock, im_w: int, im_h: int, sort: bool = False):
    """`examine_textblk` (textblock.py:302-342): reading direction, angle, font size, per-line distance."""
    quad = blk.lines_array()
    mid = (quad[:, [1, 2, 3, 0]] + quad) / 2
    vec_v = mid[:, 2] - mid[:, 0]
    vec_h = mid[:, 1] - mid[:, 3]
    centers = (quad[:, 0] + quad[:, 2]) / 2
    v = np.sum(vec_v, axis=0)
    hvec = np.sum(vec_h, axis=0)
    norm_v, norm_h = np.linalg.norm(v), np.linalg.norm(hvec)
    vertical = norm_v > norm_h if blk.language == "ja" else norm_v > norm_h * 2
    if vertical:
        primary, primary_norm = v, norm_v
        rel = centers - np.array([[im_w, 0]], dtype=np.float64)  # vertical text reads right-to-left
        font_size = int(round(norm_h / len(quad)))
    else:
        primary, primary_norm = hvec, norm_h
        rel = centers - np.array([[0, 0]], dtype=np.float64)
        font_size = int(round(norm_v / len(quad)))
    angle = int(math.atan2(primary[1], primary[0]) / math.pi * 180)
    dist = np.linalg.norm(rel, axis=1)
    rad = np.arccos(np.einsum('ij, j->i', rel, primary) / (dist * primary_norm))
    dist = np.abs(np.sin(rad) * dist)
    blk.lines = quad.astype(np.int32).tolist()
    blk.distance = dist
    blk.angle = angle - 90 if vertical else angle
    if abs(blk.angle) < 3:
        blk.angle = 0
    blk.font_size = font_size
    blk.vertical = vertical
    blk.vec = primary
    blk.norm = primary_norm
    if sort:
        blk.sort_lines()


def _try_merge(blk: TextBlock, other: TextBlock, fntsize_tol=1.3, distance_tol=2) -> bool:
    """`try_merge_textline` (textblock.py:344-373)"""
    if other.merged:
        return False
    ratio = blk.font_size / other.font_size
    n1, n2 = len(blk), len(other)
    avg = (blk.font_size * n1 + other.font_size * n2) / (n1 + n2)
    dot = blk.vec @ other.vec
    vsum = blk.vec + other.vec
    cosv = dot / blk.norm / other.norm
    gap = other.distance[-1] - blk.distance[-1]
    gap_p1 = np.linalg.norm(np.array(other.lines[-1][0]) - np.array(blk.lines[-1][0]))
    if not quads_intersect(blk.lines[-1], other.lines[-1]):
        if ratio > fntsize_tol or 1 / ratio > fntsize_tol:
            return False
        if abs(cosv) < 0.866:
            return False
        if gap > distance_tol * avg or gap_p1 > avg * 2.5:
            return False
    blk.lines.append(other.lines[0])
    blk.vec = vsum
    blk.angle = int(round(np.rad2deg(math.atan2(vsum[1], vsum[0]))))
    if blk.vertical:
        blk.angle -= 90
    blk.norm = np.linalg.norm(vsum)
    blk.distance = np.append(blk.distance, other.distance[-1])
    blk.font_size = avg
    other.merged = True
    return True


def _merge_scattered(blks: List[TextBlock]) -> List[TextBlock]:
    """`merge_textlines` (textblock.py:375-388)"""
    if len(blks) < 2:
        return blks
    blks.sort(key=lambda b: b.distance[0])
    out = []
    for i, cur in enumerate(blks):
        if cur.merged:
            continue
        for nxt in blks[i + 1:]:
            _try_merge(cur, nxt)
        out.append(cur)
    for b in out:
        b.adjust_bbox(with_bbox=False)
    return out


def _split_block(blk: TextBlock):
    """`split_textblk` (textblock.py:390-419)"""
    font_size, distance, lines = blk.font_size, blk.distance, blk.lines
    first = np.array(blk.lines[0])
    lines.sort(key=lambda ln: np.linalg.norm(np.array(ln[0]) - first[0]))
    tol = font_size * 2
    cur = copy.deepcopy(blk)
    cur.lines = [first]
    parts = [cur]
    for j, line in enumerate(lines[1:]):
        split = False
        if not quads_intersect(lines[j], line):
            d = abs(distance[j + 1] - distance[j])
            if d > tol:
                split = True
            elif blk.vertical and abs(blk.angle) < 15:
                if len(cur.lines) > 1 or d > font_size:
                    split = abs(lines[j][0][1] - line[0][1]) > font_size
        if split:
            cur = copy.deepcopy(cur)
            cur.lines = [line]
            parts.append(cur)
        else:
            cur.lines.append(line)
    if len(parts) > 1:
        for c in parts:
            c.adjust_bbox(with_bbox=False)
        return True, parts
    return False, parts


def _reading_order(blks: List[TextBlock], im_w: int, im_h: int) -> List[TextBlock]:
    """`sort_textblk_list` (textblock.py:267-300): 4x3 grid reading order, right-to-left when ja dominates."""
    if len(blks) == 0:
        return blks
    n_ja = sum(1 for b in blks if b.language == "ja")
    xyxy = np.array([b.xyxy for b in blks])
    flip = n_ja > len(blks) / 2
    full_w = im_w
    if im_w > im_h:
        im_w /= 2
    gy, gx = 4, 3
    area = im_h * im_w
    cx = (xyxy[:, 0] + xyxy[:, 2]) / 2
    if flip:
        cx = (full_w - cx) if im_w != full_w else (im_w - cx)
    col = (cx / im_w * gx).astype(np.int32)
    cy = (xyxy[:, 1] + xyxy[:, 3]) / 2
    row = (cy / im_h * gy).astype(np.int32)
    cell = row * gx + col
    weights = cell * area + 1.2 * (cx - col * im_w / gx) + (cy - row * im_h / gy)
    if im_w != full_w:
        weights[np.where(col >= gx)] += area * gy * gx
    for b, wt in zip(blks, weights):
        b.weight = wt
    blks.sort(key=lambda b: b.weight)
    return blks


def group_output(blks, lines, im_w, im_h, mask=None, sort_blklist=True) -> List[TextBlock]:
    """`group_output` (textblock.py:421-508).  blks = (boxes int32 [n,4], cls int32 [n], conf f32 [n]),
    lines = int32 [m,4,2] (or []), mask = u8 page mask."""
    blk_list = [TextBlock(bbox, language=LANG_LIST[cls]) for bbox, cls, conf in zip(*blks)]
    loose = {"ver": [], "hor": []}
    assign_thresh, mask_thresh = 0.4, 0.1
    # 1. lines -> blocks by overlap / line area
    bxy = np.array([b.xyxy for b in blk_list], np.int64).reshape(-1, 4)
    for line in lines:
        bx1, bx2 = line[:, 0].min(), line[:, 0].max()
        by1, by2 = line[:, 1].min(), line[:, 1].max()
        best, best_i = -1, -1
        line_area = (by2 - by1) * (bx2 - bx1)
        if line_area != 0 and len(blk_list) > 0:
            # all blocks at once; first maximum wins, like the reference's strict `best < score` scan
            ix1, iy1 = np.maximum(bxy[:, 0], int(bx1)), np.maximum(bxy[:, 1], int(by1))
            ix2, iy2 = np.minimum(bxy[:, 2], int(bx2)), np.minimum(bxy[:, 3], int(by2))
            inter = np.where((iy2 < iy1) | (ix2 < ix1), -1, (iy2 - iy1) * (ix2 - ix1))
            score = inter / float(line_area)
            j = int(np.argmax(score))
            if best < score[j]:
                best, best_i = score[j], j
        else:
            for j, blk in enumerate(blk_list):
                score = overlap_area(blk.xyxy, [bx1, by1, bx2, by2]) / line_area
                if best < score:
                    best, best_i = score, j
        if best > assign_thresh:
            blk_list[best_i].lines.append(line)
            continue
        if mask is not None and mask[by1: by2, bx1: bx2].mean() / 255 < mask_thresh:
            continue
        single = TextBlock([bx1, by1, bx2, by2], [line])
        _orientation_and_metrics(single, im_w, im_h, sort=False)
        loose["ver" if single.vertical else "hor"].append(single)
    # 2. per block: drop empty low-mask blocks, measure, split manga columns
    final = []
    for blk in blk_list:
        if len(blk.lines) == 0:
            bx1, by1, bx2, by2 = blk.xyxy
            if mask is not None and mask[by1: by2, bx1: bx2].mean() / 255 < mask_thresh:
                continue
            blk.lines = np.array([[bx1, by1, bx2, by1, bx2, by2, bx1, by2]]).astype(np.int64).reshape(-1, 4, 2).tolist()
        _orientation_and_metrics(blk, im_w, im_h, sort=True)
        want_split = len(blk.lines) > 1 and (blk.language == "ja" or blk.vertical)
        did_split, parts = _split_block(blk) if want_split else (False, [blk])
        if not did_split:
            for b in parts:
                b.adjust_bbox(with_bbox=True)
        final += parts
    # 3. merge the loose lines, order the page
    final += _merge_scattered(loose["hor"])
    final += _merge_scattered(loose["ver"])
    if sort_blklist:
        final = _reading_order(final, im_w, im_h)
    for blk in final:
        if blk.language == "eng" and not blk.vertical:
            if len(blk.lines) == 0:
                continue
            grow = max(int(blk.font_size * 0.1), 2)
            rad = np.deg2rad(blk.angle)
            shift = np.array([[[-1, -1], [1, -1], [1, 1], [-1, 1]]]) * np.array([[[np.sin(rad), np.cos(rad)]]]) * grow
            pts = blk.lines_array() + shift
            pts[..., 0] = np.clip(pts[..., 0], 0, im_w - 1)
            pts[..., 1] = np.clip(pts[..., 1], 0, im_h - 1)
            blk.lines = pts.astype(np.int64).tolist()
            blk.font_size += grow
    return final
