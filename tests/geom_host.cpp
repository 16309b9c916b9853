// Host build of comic-text-detector_b200/csrc/geom.h for the CPU unit tests (tests/test_cpu_geom.py):
// g++ -O2 -ffp-contract=off -shared -fPIC.  No CUDA involved.
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include "../comic-text-detector_b200/csrc/geom.h"

using namespace ctdgeom;

static int build_hull(const int* xy, int n, std::vector<IPt>& hull) {
  std::vector<IPt> p(n);
  for (int i = 0; i < n; ++i) { p[i].x = xy[2 * i]; p[i].y = xy[2 * i + 1]; }
  std::sort(p.begin(), p.end(), [](const IPt& a, const IPt& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  hull.resize(kMaxHull);
  return hull_sorted(p.data(), n, hull.data(), kMaxHull);
}

extern "C" int geom_min_area_rect(const int* xy, int n, float* out5) {
  std::vector<IPt> hull, tmp(kMaxHull);
  const int nh = build_hull(xy, n, hull);
  if (nh < 3) return nh;
  hull_start_maxx(hull.data(), nh, tmp.data());
  std::vector<float> f0(kMaxHull), f1(kMaxHull), f2(kMaxHull);
  const RRect r = min_area_rect(hull.data(), nh, f0.data(), f1.data(), f2.data());
  out5[0] = r.cx; out5[1] = r.cy; out5[2] = r.w; out5[3] = r.h; out5[4] = r.angle;
  return nh;
}

extern "C" int geom_contour_box(const int* xy, int n, int map_w, int map_h, int dst_w, int dst_h, double ratio, int16_t* box8) {
  std::vector<IPt> hull, tmp(kMaxHull), off(kMaxOffsetPts);
  const int nh = build_hull(xy, n, hull);
  std::vector<float> f0(kMaxHull), f1(kMaxHull), f2(kMaxHull);
  return contour_to_box(hull.data(), nh, tmp.data(), off.data(), f0.data(), f1.data(), f2.data(), map_w, map_h, dst_w, dst_h,
                        ratio, box8) ? 1 : 0;
}

extern "C" int geom_unclip(const float* bx, const float* by, double ratio, int* out_xy, int cap) {
  std::vector<IPt> off(cap);
  const int m = unclip_offset(bx, by, ratio, off.data(), cap);
  for (int i = 0; i < m && i < cap; ++i) { out_xy[2 * i] = off[i].x; out_xy[2 * i + 1] = off[i].y; }
  return m;
}
