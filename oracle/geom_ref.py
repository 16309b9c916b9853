"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatements of the two third-party geometry
libraries the reference's post-processing calls but that are NOT in /root/reference and
NOT installed in this image (they are not even listed in requirements.txt:1-12):

  * pyclipper (any 1.x; Cython wrapper of Angus Johnson's Clipper 6.4.2) --
    call site utils/db_utils.py:168-174 (`unclip`):
        offset = pyclipper.PyclipperOffset(); offset.AddPath(box, JT_ROUND, ET_CLOSEDPOLYGON)
        expanded = np.array(offset.Execute(distance))
  * shapely (GEOS) -- call sites utils/db_utils.py:169-170 (`Polygon(box).area/.length`),
    utils/textblock.py:355-356,400-402 (`Polygon(line).intersects`).

PARITY UNPINNED: the reference holds no test or golden vector for these results, and the
libraries cannot be run here.  What follows restates their *published* algorithms:
Clipper 6.4.2 `ClipperOffset::{AddPath,FixOrientations,DoOffset,OffsetPoint,DoRound}`
(clipper.cpp) and GEOS `algorithm::Area::ofRingSigned`, `algorithm::Length::ofLine`.
"""
import math
import types

import numpy as np

JT_SQUARE, JT_ROUND, JT_MITER = 0, 1, 2
ET_CLOSEDPOLYGON, ET_CLOSEDLINE, ET_OPENBUTT, ET_OPENSQUARE, ET_OPENROUND = 0, 1, 2, 3, 4

_PI = 3.141592653589793238
_TWO_PI = _PI * 2
_DEF_ARC_TOLERANCE = 0.25


def _cround(val: float) -> int:
    # clipper.cpp `Round`: (val < 0) ? (cInt)(val - 0.5) : (cInt)(val + 0.5)  (C cast truncates)
    return int(val - 0.5) if val < 0 else int(val + 0.5)


def _clipper_area(poly):
    # clipper.cpp `Area(const Path&)`
    n = len(poly)
    if n < 3:
        return 0.0
    a = 0.0
    j = n - 1
    for i in range(n):
        a += (float(poly[j][0]) + float(poly[i][0])) * (float(poly[j][1]) - float(poly[i][1]))
        j = i
    return -a * 0.5


def _unit_normal(p1, p2):
    # clipper.cpp `GetUnitNormal`
    if p2[0] == p1[0] and p2[1] == p1[1]:
        return (0.0, 0.0)
    dx = float(p2[0] - p1[0])
    dy = float(p2[1] - p1[1])
    f = 1 * 1.0 / math.sqrt(dx * dx + dy * dy)
    dx *= f
    dy *= f
    return (dy, -dx)


def to_cint_path(path):
    """pyclipper `_to_clipper_path`: every coordinate is converted to Clipper's cInt
    (long long) by Cython's object->C-integer conversion, which for float objects goes
    through nb_int, i.e. truncation toward zero."""
    return [(int(p[0]), int(p[1])) for p in path]


def clipper_offset_closed_polygon(path, delta, jointype=JT_ROUND,
                                  miter_limit=2.0, arc_tolerance=0.25):
    """Raw `m_destPoly` of ClipperOffset for ONE closed polygon path, i.e. the vertex
    list before the trailing `Clipper::Execute(ctUnion, pftPositive)` clean-up.  For the
    reference's use (convex quad, delta > 0, consumer = cv2.minAreaRect which only looks
    at the convex hull of the returned points) the union only drops duplicate/collinear
    vertices and rotates the start vertex, none of which changes the hull.
    Returns [] when Clipper would add nothing (degenerate path)."""
    assert jointype == JT_ROUND, "only JT_ROUND is on the reference's path"
    path = to_cint_path(path)
    # ---- AddPath: strip duplicates -------------------------------------------------
    high = len(path) - 1
    if high < 0:
        return []
    while high > 0 and path[0] == path[high]:
        high -= 1
    contour = [path[0]]
    for i in range(1, high + 1):
        if contour[-1] != path[i]:
            contour.append(path[i])
    if len(contour) < 3:
        return []
    # ---- FixOrientations (single path => it holds the lowest vertex) ---------------
    if not (_clipper_area(contour) >= 0):
        contour = contour[::-1]
    # ---- DoOffset ---------------------------------------------------------------------
    if abs(delta) < 1e-20:  # NEAR_ZERO
        return list(contour)
    if arc_tolerance <= 0.0:
        y = _DEF_ARC_TOLERANCE
    elif arc_tolerance > abs(delta) * _DEF_ARC_TOLERANCE:
        y = abs(delta) * _DEF_ARC_TOLERANCE
    else:
        y = arc_tolerance
    steps = _PI / math.acos(1 - y / abs(delta))
    if steps > abs(delta) * _PI:
        steps = abs(delta) * _PI
    m_sin = math.sin(_TWO_PI / steps)
    m_cos = math.cos(_TWO_PI / steps)
    steps_per_rad = steps / _TWO_PI
    if delta < 0.0:
        m_sin = -m_sin
    n = len(contour)
    normals = [_unit_normal(contour[j], contour[j + 1]) for j in range(n - 1)]
    normals.append(_unit_normal(contour[n - 1], contour[0]))
    dest = []
    k = n - 1
    for j in range(n):
        # ---- OffsetPoint(j, k, jtRound) --------------------------------------------
        sin_a = normals[k][0] * normals[j][1] - normals[j][0] * normals[k][1]
        done = False
        if abs(sin_a * delta) < 1.0:
            cos_a = normals[k][0] * normals[j][0] + normals[j][1] * normals[k][1]
            if cos_a > 0:
                dest.append((_cround(contour[j][0] + normals[k][0] * delta),
                             _cround(contour[j][1] + normals[k][1] * delta)))
                done = True
        elif sin_a > 1.0:
            sin_a = 1.0
        elif sin_a < -1.0:
            sin_a = -1.0
        if not done:
            if sin_a * delta < 0:
                dest.append((_cround(contour[j][0] + normals[k][0] * delta),
                             _cround(contour[j][1] + normals[k][1] * delta)))
                dest.append(contour[j])
                dest.append((_cround(contour[j][0] + normals[j][0] * delta),
                             _cround(contour[j][1] + normals[j][1] * delta)))
            else:
                # ---- DoRound(j, k) ---------------------------------------------------
                a = math.atan2(sin_a, normals[k][0] * normals[j][0] + normals[k][1] * normals[j][1])
                nsteps = max(int(_cround(steps_per_rad * abs(a))), 1)
                X, Y = normals[k]
                for _ in range(nsteps):
                    dest.append((_cround(contour[j][0] + X * delta),
                                 _cround(contour[j][1] + Y * delta)))
                    X2 = X
                    X = X * m_cos - m_sin * Y
                    Y = X2 * m_sin + Y * m_cos
                dest.append((_cround(contour[j][0] + normals[j][0] * delta),
                             _cround(contour[j][1] + normals[j][1] * delta)))
        k = j
    return dest


# ------------------------------------------------------------------------------------
# GEOS restatements


def geos_ring_area(pts) -> float:
    """|GEOS algorithm::Area::ofRingSigned| of the closed ring pts+[pts[0]] (doubles)."""
    ring = [(float(p[0]), float(p[1])) for p in pts]
    ring.append(ring[0])
    n = len(ring)
    if n < 3:
        return 0.0
    x0 = ring[0][0]
    p1x, p1y = ring[0]
    p2x, p2y = ring[1][0] - x0, ring[1][1]
    s = 0.0
    for i in range(1, n - 1):
        p0y = p1y
        p1x, p1y = p2x, p2y
        p2x, p2y = ring[i + 1][0] - x0, ring[i + 1][1]
        s += p1x * (p0y - p2y)
    return abs(s / 2.0)


def geos_ring_length(pts) -> float:
    """GEOS algorithm::Length::ofLine of the closed ring (doubles)."""
    ring = [(float(p[0]), float(p[1])) for p in pts]
    ring.append(ring[0])
    x0, y0 = ring[0]
    ln = 0.0
    for i in range(1, len(ring)):
        x1, y1 = ring[i]
        dx, dy = x1 - x0, y1 - y0
        ln += math.sqrt(dx * dx + dy * dy)
        x0, y0 = x1, y1
    return ln


def _orient(a, b, c):
    v = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
    return (v > 0) - (v < 0)


def _on_seg(a, b, p):
    return (min(a[0], b[0]) <= p[0] <= max(a[0], b[0]) and
            min(a[1], b[1]) <= p[1] <= max(a[1], b[1]))


def _seg_intersect(a, b, c, d):
    o1, o2, o3, o4 = _orient(a, b, c), _orient(a, b, d), _orient(c, d, a), _orient(c, d, b)
    if o1 != o2 and o3 != o4:
        return True
    if o1 == 0 and _on_seg(a, b, c):
        return True
    if o2 == 0 and _on_seg(a, b, d):
        return True
    if o3 == 0 and _on_seg(c, d, a):
        return True
    if o4 == 0 and _on_seg(c, d, b):
        return True
    return False


def _pt_in_poly(p, poly):
    """closed-set point-in-polygon (boundary counts), exact for integer input."""
    n = len(poly)
    inside = False
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        if _orient(a, b, p) == 0 and _on_seg(a, b, p):
            return True
        if (a[1] > p[1]) != (b[1] > p[1]):
            # x of the edge at p.y compared to p.x, exact with integers via cross product
            t = (b[0] - a[0]) * (p[1] - a[1]) - (p[0] - a[0]) * (b[1] - a[1])
            if (t > 0) == (b[1] > a[1]) and t != 0:
                inside = not inside
    return inside


def polygons_intersect(pa, pb) -> bool:
    """shapely `Polygon(pa).intersects(Polygon(pb))` for simple polygons: closed point-set
    intersection (touching counts) = some edges cross/touch or one ring holds a vertex of
    the other."""
    pa = [(int(p[0]), int(p[1])) if float(p[0]).is_integer() and float(p[1]).is_integer()
          else (float(p[0]), float(p[1])) for p in pa]
    pb = [(int(p[0]), int(p[1])) if float(p[0]).is_integer() and float(p[1]).is_integer()
          else (float(p[0]), float(p[1])) for p in pb]
    na, nb = len(pa), len(pb)
    for i in range(na):
        for j in range(nb):
            if _seg_intersect(pa[i], pa[(i + 1) % na], pb[j], pb[(j + 1) % nb]):
                return True
    return _pt_in_poly(pa[0], pb) or _pt_in_poly(pb[0], pa)


# ------------------------------------------------------------------------------------
# stand-in modules so that the UNMODIFIED reference files import (see oracle/ref_shim.py)


class _PyclipperOffset:
    def __init__(self, miter_limit=2.0, arc_tolerance=0.25):
        self.MiterLimit = miter_limit
        self.ArcTolerance = arc_tolerance
        self._paths = []

    def AddPath(self, path, join_type, end_type):
        assert end_type == ET_CLOSEDPOLYGON
        self._paths.append((np.asarray(path).reshape(-1, 2).tolist(), join_type))

    def Execute(self, delta):
        out = []
        for path, jt in self._paths:
            pts = clipper_offset_closed_polygon(path, float(delta), jt, self.MiterLimit, self.ArcTolerance)
            if pts:
                out.append([[int(x), int(y)] for x, y in pts])
        return out


def make_pyclipper_module():
    m = types.ModuleType("pyclipper")
    m.PyclipperOffset = _PyclipperOffset
    m.JT_ROUND, m.JT_SQUARE, m.JT_MITER = JT_ROUND, JT_SQUARE, JT_MITER
    m.ET_CLOSEDPOLYGON = ET_CLOSEDPOLYGON
    return m


class _Polygon:
    def __init__(self, pts):
        self.pts = np.asarray(pts).reshape(-1, 2)

    @property
    def area(self):
        return geos_ring_area(self.pts)

    @property
    def length(self):
        return geos_ring_length(self.pts)

    def intersects(self, other):
        return polygons_intersect(self.pts.tolist(), other.pts.tolist())


def make_shapely_modules():
    shp = types.ModuleType("shapely")
    geo = types.ModuleType("shapely.geometry")
    geo.Polygon = _Polygon
    shp.geometry = geo
    return shp, geo
