// refine_mask on the GPU (reference utils/textmask.py:159-169 and callees 16-132):
// per text block window: grey conversion, eroded-mask histogram -> top-3 colours -> inRange candidates,
// per-channel Otsu candidate, polarity by min xor-sum, candidate-by-candidate connected-component merge
// against the eroded/thresholded mask, 3x3 dilation (inpaint mode), hole filling, OR into the page mask.
//
// One GROUP of CTAs per block window: a single CTA for small windows, a thread-block CLUSTER of kRefCluster CTAs for
// large ones (a 1024x1024 window took 78 ms on one CTA: the phases are latency-bound sweeps over the window).  The
// phases are separated by group barriers (__syncthreads / barrier.cluster); every per-window plane lives in a
// global scratch arena (L2 resident), the per-window histograms / sums are reduced across the cluster through
// distributed shared memory and every CTA then takes the (deterministic) scalar decisions redundantly.  Windows of
// ALL pages of a batch are processed by one launch per group size (RefineWin::page selects the image / mask / output
// planes).  Integer arithmetic throughout except the float64 numpy/OpenCV formulas that are replicated literally
// (np.histogram bin mapping, np.linspace edges, cv2 Otsu, cvRound of inRange bounds).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace ctd {

constexpr int kRefThreads = 512;         // single-CTA windows
constexpr int kRefThreadsCluster = 1024; // cluster windows: the sweeps are latency-bound, one pixel per thread-iteration
constexpr int ref_threads(int cl) { return cl > 1 ? kRefThreadsCluster : kRefThreads; }
constexpr int kRefCluster = 8;          // CTAs per large window (portable cluster size)
constexpr int kRefLargePx = 24 * 1024;  // windows with more pixels go to the cluster kernel

// one window group: CL CTAs (cluster) of kRefThreads threads
template <int CL>
struct Grp {
  __device__ static __forceinline__ int rank() { return CL == 1 ? 0 : int(cg::this_cluster().block_rank()); }
  static constexpr int threads = ref_threads(CL);
  __device__ static __forceinline__ int tid() { return rank() * threads + int(threadIdx.x); }
  static constexpr int size = CL * threads;
  __device__ static __forceinline__ void sync() {
    if constexpr (CL == 1) __syncthreads();
    else cg::this_cluster().sync();
  }
  // every CTA ends up with the cluster-wide sum of its `n`-entry shared array (n <= 1024); `tmp` is shared scratch
  template <typename T>
  __device__ static __forceinline__ void allreduce(T* arr, int n, T* tmp) {
    if constexpr (CL == 1) {
      __syncthreads();
    } else {
      cg::cluster_group cl = cg::this_cluster();
      cl.sync();                       // local accumulation finished everywhere
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        T sum = 0;
        for (int r = 0; r < CL; ++r) sum += *cl.map_shared_rank(arr + i, r);
        tmp[i] = sum;
      }
      cl.sync();                       // everybody has read everybody
      for (int i = threadIdx.x; i < n; i += blockDim.x) arr[i] = tmp[i];
      __syncthreads();
    }
  }
};

struct RefineWin {
  int x1, y1, x2, y2;     // window (python slice semantics: rows y1..y2-1, cols x1..x2-1)
  long long off;          // pixel offset of this window's planes inside each scratch plane
  int page;               // page of the batch the window belongs to
  int pad;
};

struct RefinePlanes {
  uint8_t* grey;          // [total]
  uint8_t* cand;          // [total] current candidate (0/255)
  uint8_t* predm;         // [total] erode(cross)+thr 60 of the mask crop (0/255)
  uint8_t* merged;        // [total]
  uint8_t* tmp;           // [total] dilation target / inverse
  int* L;                 // [total] union-find
  int* acc;               // [4*total] per-root: area, gain, loss, maxidx
};

// XSM: the union-find array is updated by CTAs on other SMs (cluster groups): parent reads bypass L1 (a stale parent
// would still be an ancestor -- parents only decrease -- but the chase would take longer and the final flatten must
// see the final tree)
template <bool XSM>
__device__ __forceinline__ int rf_load(const int* p) {
  if constexpr (XSM) return __ldcg(p);
  else return *p;
}
template <bool XSM>
__device__ __forceinline__ int rf_find(const int* L, int a) {
  int p = rf_load<XSM>(L + a);
  while (p != a) {
    a = p;
    p = rf_load<XSM>(L + a);
  }
  return a;
}
// find + full path compression: every node on the path is pointed at the root (atomicMin: a concurrent union may
// already have lowered it further, and a root found here may meanwhile have become a child -- it is still an ancestor)
template <bool XSM>
__device__ __forceinline__ int rf_find_compress(int* L, int a) {
  const int r = rf_find<XSM>(L, a);
  while (a != r) {
    const int p = rf_load<XSM>(L + a);
    if (p <= r) break;
    atomicMin(&L[a], r);
    a = p;
  }
  return r;
}
template <bool XSM>
__device__ __forceinline__ void rf_union(int* L, int a, int b) {
  bool done;
  do {
    a = rf_find<XSM>(L, a);
    b = rf_find<XSM>(L, b);
    if (a < b) {
      const int old = atomicMin(&L[b], a);
      done = old == b;
      b = old;
    } else if (b < a) {
      const int old = atomicMin(&L[a], b);
      done = old == a;
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// 8-connectivity labelling of plane `src` (non-zero = foreground) inside one window by the whole group.
// Afterwards L[i] = root (smallest index of the component) or -1.
template <int CL>
__device__ void grp_ccl(const uint8_t* __restrict__ src, int rw, int rh, int* __restrict__ L) {
  using G = Grp<CL>;
  // run starts: one thread per 64-pixel row segment (sequential inside the segment), then the segment seams
  constexpr int kSegW = 64;
  const int segs = (rw + kSegW - 1) / kSegW;
  for (int t = G::tid(); t < rh * segs; t += G::size) {
    const int y = t / segs, x0 = (t - y * segs) * kSegW;
    const int x1 = min(rw, x0 + kSegW);
    int start = -1;
    const int base = y * rw;
    for (int x = x0; x < x1; ++x) {
      if (src[base + x]) {
        if (start < 0) start = base + x;
        L[base + x] = start;
      } else {
        start = -1;
        L[base + x] = -1;
      }
    }
  }
  G::sync();
  const int n = rw * rh;
  if (segs > 1) {
    for (int t = G::tid(); t < rh * (segs - 1); t += G::size) {
      const int y = t / (segs - 1), x = ((t - y * (segs - 1)) + 1) * kSegW;
      const int i = y * rw + x;
      if (src[i] && src[i - 1]) rf_union<(CL > 1)>(L, i, i - 1);
    }
    G::sync();
  }
  // contacts with the row above; all predicates read `src` (L is being rewritten by the unions)
  for (int i = G::tid(); i < n; i += G::size) {
    if (!src[i] || i < rw) continue;
    const int x = i % rw;
    const int up = i - rw;
    if (src[up]) {
      // only the first pixel of each (current run x upper run) overlap issues the union
      const bool first = x == 0 || !src[i - 1] || !src[up - 1];
      if (first) rf_union<(CL > 1)>(L, i, up);
    } else {
      if (x > 0 && src[up - 1]) rf_union<(CL > 1)>(L, i, up - 1);
      if (x + 1 < rw && src[up + 1]) rf_union<(CL > 1)>(L, i, up + 1);
    }
  }
  G::sync();
  // flatten in two steps.  The chain nodes of the forest are the segment-run starts (every other pixel points at its
  // run start): on a window that is mostly one component the unions above can leave a chain as long as the window is
  // high, and letting each of its ~10^6 pixels walk it alone was the single largest cost of the kernel.  Step 1 walks
  // from the run starts only and compresses the path for everybody; step 2 is then one or two hops per pixel.
  for (int i = G::tid(); i < n; i += G::size) {
    if (!src[i]) continue;
    const int x = i % rw;
    if ((x % kSegW) == 0 || !src[i - 1]) rf_find_compress<(CL > 1)>(L, i);
  }
  G::sync();
  for (int i = G::tid(); i < n; i += G::size) {
    const int p = L[i];
    if (p >= 0) L[i] = rf_find<(CL > 1)>(L, p);
  }
  G::sync();
}

// shared-memory histogram increment, aggregated over the lanes of the (fully converged) warp that hit the same bin;
// bin < 0 = this lane has nothing to add.  A page window is mostly one colour, so per-lane atomics serialise.
__device__ __forceinline__ void hist_add(int* hist, int bin) {
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (bin >= 0 && int(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
}

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* sm) {
  // sm: one shared slot, zeroed by the caller before a __syncthreads
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(sm, v);
  return 0;
}

// merge step shared by the candidate loop and the hole filling (textmask.py:92-108 / 118-131):
// a label is OR-ed into `merged` iff that lowers xor(merged, pred) inside the label's bounding box, i.e.
// iff among the label's pixels not yet in `merged` more have pred == 255 than pred == 0.
template <int CL>
__device__ void grp_merge_labels(const int* __restrict__ L, const uint8_t* __restrict__ predm, uint8_t* __restrict__ merged,
                                 int* __restrict__ acc, int n, int rw, bool small_bbox_rule, int area_thresh) {
  using G = Grp<CL>;
  int* area = acc;
  int* gain = acc + n;
  int* loss = acc + 2 * n;
  int* maxi = acc + 3 * n;
  for (int i = G::tid(); i < n; i += G::size)
    if (L[i] == i) { area[i] = 0; gain[i] = 0; loss[i] = 0; maxi[i] = -1; }
  G::sync();
  // per-label sums, warp-aggregated: the 32 consecutive pixels of a warp mostly share a label (page background,
  // strokes), and one global atomic per pixel on the SAME counter serialises in L2 (measured: 12 of the 15.8 ms of a
  // 1024x1024 window).  Lanes with equal roots elect a leader that adds the group's totals.
  for (int base = G::tid() - int(threadIdx.x & 31); base < n; base += G::size) {
    const int lane = threadIdx.x & 31;
    const int i = base + lane;
    const int r = i < n ? L[i] : -1;
    const bool un = r >= 0 && merged[i] == 0;
    const bool pg = un && predm[i] != 0;
    const unsigned peers = __match_any_sync(0xffffffffu, r);
    const unsigned bg = __ballot_sync(0xffffffffu, pg), bl = __ballot_sync(0xffffffffu, un && !pg);
    if (r >= 0 && lane == __ffs(peers) - 1) {
      atomicAdd(&area[r], __popc(peers));
      atomicMax(&maxi[r], base + 31 - __clz(peers));
      const int g_ = __popc(peers & bg), l_ = __popc(peers & bl);
      if (g_) atomicAdd(&gain[r], g_);
      if (l_) atomicAdd(&loss[r], l_);
    }
  }
  G::sync();
  for (int i = G::tid(); i < n; i += G::size) {
    const int r = L[i];
    if (r < 0) continue;
    bool ok;
    if (small_bbox_rule) {
      // `if w * h < 3: continue` (textmask.py:97): bounding boxes 1x1, 1x2, 2x1
      const int a = area[r];
      const bool tiny = a == 1 || (a == 2 && (maxi[r] == r + 1 || maxi[r] == r + rw));
      ok = !tiny;
    } else {
      ok = area[r] < area_thresh;  // textmask.py:120
    }
    if (ok && gain[r] > loss[r]) merged[i] = 255;
  }
  G::sync();
}

template <int CL>
__global__ void __launch_bounds__(ref_threads(CL)) refine_kernel(const uint8_t* __restrict__ img_all, const uint8_t* __restrict__ mask_all,
                                                             int H, int W, const RefineWin* __restrict__ wins,
                                                             const int* __restrict__ win_idx, RefinePlanes P,
                                                             int refine_mode, uint32_t* __restrict__ out_all) {
  using G = Grp<CL>;
  const RefineWin win = wins[win_idx[blockIdx.x / CL]];
  const int rw = win.x2 - win.x1, rh = win.y2 - win.y1;
  if (rw <= 0 || rh <= 0) return;     // uniform over the group
  const int n = rw * rh;
  const uint8_t* img = img_all + size_t(win.page) * H * W * 3;
  const uint8_t* mask = mask_all + size_t(win.page) * H * W;
  uint32_t* out_words = out_all + size_t(win.page) * H * W / 4;
  uint8_t* grey = P.grey + win.off;
  uint8_t* cand = P.cand + win.off;
  uint8_t* predm = P.predm + win.off;
  uint8_t* merged = P.merged + win.off;
  uint8_t* tmp = P.tmp + win.off;
  int* L = P.L + win.off;
  int* acc = P.acc + 4 * win.off;

  __shared__ int hist_all[1024];       // [0,256): grey histogram of the eroded-mask pixels; then the per-channel
  int* hist_g = hist_all;              // histograms of the whole window (Otsu), B | G | R
  int (*hist_c)[256] = reinterpret_cast<int (*)[256]>(hist_all + 256);
  __shared__ int red_tmp[1024];        // staging of the cluster all-reduce
  __shared__ unsigned long long red_tmp64[12];
  __shared__ int cnt255[256];          // np.histogram(bins=255) counts
  __shared__ int order[256];
  __shared__ double edges[256];
  __shared__ int s_first, s_last, s_ncol, s_total;
  __shared__ int lo[3], hi[3], otsu_t[3];
  __shared__ unsigned long long xs[12];  // xor sums: [k][pos/neg] for 3 colours, then 3 channels
  __shared__ int proc_kind[4], proc_neg[4], s_nproc;
  __shared__ int top2[2], s_area0, ctop[2];

  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    hist_g[i] = 0; hist_c[0][i] = 0; hist_c[1][i] = 0; hist_c[2][i] = 0; cnt255[i] = 0;
  }
  if (threadIdx.x < 12) xs[threadIdx.x] = 0ull;
  __syncthreads();

  // ---- phase 0: grey, eroded candidates, pred mask, histograms ---------------------------------
  // (warp-uniform trip count: hist_add aggregates equal bins across the warp before touching shared memory)
  for (int i0 = G::tid() - int(threadIdx.x & 31); i0 < n; i0 += G::size) {
    const int i = i0 + int(threadIdx.x & 31);
    const bool in = i < n;
    int b = -1, g = -1, r = -1, gr = 0, m3 = 255, mc = 255;
    if (in) {
      const int y = i / rw, x = i - y * rw;
      const size_t gp = size_t(win.y1 + y) * W + win.x1 + x;
      b = img[gp * 3]; g = img[gp * 3 + 1]; r = img[gp * 3 + 2];
      gr = (b * 1868 + g * 9617 + r * 4899 + 8192) >> 14;  // cv2.COLOR_BGR2GRAY, 8u fixed point
      grey[i] = (uint8_t)gr;
      // erosions of the mask CROP (window borders ignore the outside: BORDER_CONSTANT with +inf)
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= rh) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= rw) continue;
          const int mv = mask[size_t(win.y1 + yy) * W + win.x1 + xx];
          m3 = min(m3, mv);
          if (dx == 0 || dy == 0) mc = min(mc, mv);
        }
      }
      predm[i] = mc > 60 ? 255 : 0;                  // textmask.py:86-89
      merged[i] = 0;
    }
    hist_add(&hist_c[0][0], b);                      // all 32 lanes take part (out-of-range lanes add nothing)
    hist_add(&hist_c[1][0], g);
    hist_add(&hist_c[2][0], r);
    hist_add(hist_g, (in && m3 > 127) ? gr : -1);    // textmask.py:60
  }
  G::allreduce(hist_all, 1024, red_tmp);     // also orders the plane writes above before every later phase

  // ---- phase 1: np.histogram(bins=255) of the candidate grey values, top-k colours, Otsu ---------
  if (threadIdx.x == 0) {
    int first = -1, last = -1, total = 0;
    for (int v = 0; v < 256; ++v)
      if (hist_g[v]) { if (first < 0) first = v; last = v; total += hist_g[v]; }
    s_first = first; s_last = last; s_total = total;
  }
  __syncthreads();
  {
    // outer edges (numpy _get_outer_edges): empty -> (0,1); equal -> (v-0.5, v+0.5)
    double fe, le;
    if (s_total == 0) { fe = 0.0; le = 1.0; }
    else if (s_first == s_last) { fe = s_first - 0.5; le = s_last + 0.5; }
    else { fe = s_first; le = s_last; }
    const double step = (le - fe) / 255.0;  // np.linspace(fe, le, 256): arange * step + start, last = stop
    for (int i = threadIdx.x; i < 256; i += blockDim.x) edges[i] = (i == 255) ? le : __dadd_rn(__dmul_rn((double)i, step), fe);
    __syncthreads();
    for (int v = threadIdx.x; v < 256; v += blockDim.x) {
      if (!hist_g[v]) continue;
      const double a = (double)v;
      const double f = ((a - fe) / (le - fe)) * 255.0;  // numpy fast path: (tmp_a - first_edge) / norm_denom * n_bins
      int idx = (int)f;
      if (idx == 255) idx = 254;
      if (a < edges[idx]) --idx;
      if (a >= edges[idx + 1] && idx != 254) ++idx;
      atomicAdd(&cnt255[idx], hist_g[v]);
    }
  }
  __syncthreads();
  // stable descending order of the 255 bins (documented normalisation of np.argsort's tie order)
  for (int b = threadIdx.x; b < 255; b += blockDim.x) {
    int rank = 0;
    const int cb = cnt255[b];
    for (int c = 0; c < 255; ++c) rank += (cnt255[c] > cb) || (cnt255[c] == cb && c < b);
    order[rank] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // get_topk_color (textmask.py:16-27): colour = LEFT EDGE of the bin (textmask.py:61-62 swaps the names)
    double top[3];
    int nt = 1;
    top[0] = edges[order[0]];
    const double tol = (double)s_total * 0.001;
    for (int j = 1; j < 255; ++j) {
      const double col = edges[order[j]];
      double dmin = 1e300;
      for (int t = 0; t < nt; ++t) dmin = fmin(dmin, fabs(top[t] - col));
      if (dmin > 10.0) top[nt++] = col;
      if (nt >= 3 || (double)cnt255[order[j]] < tol) break;
    }
    s_ncol = nt;
    for (int t = 0; t < nt; ++t) {
      const double c_top = fmin(top[t] + 30.0, 255.0);
      const double c_bot = c_top - 60.0;
      // cv2.inRange with float bounds on 8u data: cvRound (half to even) + saturate
      lo[t] = (int)fmin(fmax(rint(c_bot), 0.0), 255.0);
      hi[t] = (int)fmin(fmax(rint(c_top), 0.0), 255.0);
    }
  }
  if (threadIdx.x >= 32 && threadIdx.x < 35) {
    // cv2.threshold(..., THRESH_OTSU): getThreshVal_Otsu_8u
    const int* hh = hist_c[threadIdx.x - 32];
    const double scale = 1.0 / (double)n;
    double mu = 0;
    for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)hh[i]));
    mu = __dmul_rn(mu, scale);
    double mu1 = 0, q1 = 0, max_sigma = 0;
    int max_val = 0;
    for (int i = 0; i < 256; ++i) {
      const double p_i = __dmul_rn((double)hh[i], scale);
      mu1 = __dmul_rn(mu1, q1);
      q1 = __dadd_rn(q1, p_i);
      const double q2 = 1.0 - q1;
      if (fmin(q1, q2) < 1.1920929e-07 || fmax(q1, q2) > 1.0 - 1.1920929e-07) continue;
      // explicit roundings: the x86 build of OpenCV has no FMA contraction here
      mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
      const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
      const double dm = __dsub_rn(mu1, mu2);
      const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), dm), dm);
      if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    otsu_t[threadIdx.x - 32] = max_val;
  }
  __syncthreads();

  // ---- phase 2: xor sums of every candidate and of its negative against the mask crop -------------
  {
    unsigned long long loc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) loc[k] = 0ull;
    const int ncol = s_ncol;
    for (int i = G::tid(); i < n; i += G::size) {
      const int y = i / rw, x = i - y * rw;
      const size_t gp = size_t(win.y1 + y) * W + win.x1 + x;
      const int mk = mask[gp];
      const int gr = grey[i];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (k < ncol) {
          const int t = (gr >= lo[k] && gr <= hi[k]) ? 255 : 0;
          loc[2 * k] += (unsigned)(t ^ mk);
          loc[2 * k + 1] += (unsigned)((255 - t) ^ mk);
        }
        const int ch = img[gp * 3 + k];
        const int t2 = ch > otsu_t[k] ? 255 : 0;
        loc[6 + 2 * k] += (unsigned)(t2 ^ mk);
        loc[6 + 2 * k + 1] += (unsigned)((255 - t2) ^ mk);
      }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) block_sum_u64(loc[k], &xs[k]);
  }
  G::allreduce(xs, 12, red_tmp64);
  if (threadIdx.x == 0) {
    // minxor_thresh (textmask.py:29-41): negative wins only if strictly smaller
    unsigned long long best[4];
    int kind[4], neg[4], np_ = 0;
    for (int k = 0; k < s_ncol; ++k) {
      const bool ng = xs[2 * k + 1] < xs[2 * k];
      best[np_] = ng ? xs[2 * k + 1] : xs[2 * k];
      kind[np_] = k; neg[np_] = ng; ++np_;
    }
    // Otsu: best channel (stable sort by xor sum -> first minimum in B,G,R order)  (textmask.py:43-54)
    int bc = 0, bneg = 0;
    unsigned long long bv = ~0ull;
    for (int c = 0; c < 3; ++c) {
      const bool ng = xs[6 + 2 * c + 1] < xs[6 + 2 * c];
      const unsigned long long v = ng ? xs[6 + 2 * c + 1] : xs[6 + 2 * c];
      if (v < bv) { bv = v; bc = c; bneg = ng; }
    }
    best[np_] = bv; kind[np_] = 3 + bc; neg[np_] = bneg; ++np_;
    // mask_list.sort(key=xor_sum) (textmask.py:74): stable insertion sort
    for (int i = 1; i < np_; ++i) {
      const unsigned long long v = best[i];
      const int kk = kind[i], nn = neg[i];
      int j = i - 1;
      while (j >= 0 && best[j] > v) { best[j + 1] = best[j]; kind[j + 1] = kind[j]; neg[j + 1] = neg[j]; --j; }
      best[j + 1] = v; kind[j + 1] = kk; neg[j + 1] = nn;
    }
    for (int i = 0; i < np_; ++i) { proc_kind[i] = kind[i]; proc_neg[i] = neg[i]; }
    s_nproc = np_;
  }
  __syncthreads();

  // ---- phase 3: candidates in order: label, test every label, merge -------------------------------
  for (int c = 0; c < s_nproc; ++c) {
    const int kind = proc_kind[c], neg = proc_neg[c];
    for (int i = G::tid(); i < n; i += G::size) {
      int t;
      if (kind < 3) {
        const int gr = grey[i];
        t = (gr >= lo[kind] && gr <= hi[kind]) ? 255 : 0;
      } else {
        const int y = i / rw, x = i - y * rw;
        const size_t gp = size_t(win.y1 + y) * W + win.x1 + x;
        t = img[gp * 3 + (kind - 3)] > otsu_t[kind - 3] ? 255 : 0;
      }
      cand[i] = (uint8_t)(neg ? 255 - t : t);
    }
    G::sync();
    grp_ccl<CL>(cand, rw, rh, L);
    grp_merge_labels<CL>(L, predm, merged, acc, n, rw, true, 0);
  }

  // ---- phase 4: dilate 3x3 (inpaint mode) ------------------------------------------------------------
  if (refine_mode == 0) {
    for (int i = G::tid(); i < n; i += G::size) {
      const int y = i / rw, x = i - y * rw;
      int m = 0;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= rh) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= rw) continue;
          m = max(m, (int)merged[yy * rw + xx]);
        }
      }
      tmp[i] = (uint8_t)m;
    }
    G::sync();
    for (int i = G::tid(); i < n; i += G::size) merged[i] = tmp[i];
    G::sync();
  }

  // ---- phase 5: fill holes: components of the inverse smaller than the 2nd largest area -------------
  if (threadIdx.x == 0) { top2[0] = -1; top2[1] = -1; s_area0 = 0; }
  for (int i = G::tid(); i < n; i += G::size) tmp[i] = merged[i] ? 0 : 255;
  G::sync();
  grp_ccl<CL>(tmp, rw, rh, L);
  {
    int* area = acc;
    for (int i = G::tid(); i < n; i += G::size)
      if (L[i] == i) area[i] = 0;
    G::sync();
    int a0 = 0;
    for (int base = G::tid() - int(threadIdx.x & 31); base < n; base += G::size) {
      const int lane = threadIdx.x & 31;
      const int i = base + lane;
      const int r = i < n ? L[i] : -2;
      const unsigned peers = __match_any_sync(0xffffffffu, r);
      if (r >= 0) {
        if (lane == __ffs(peers) - 1) atomicAdd(&area[r], __popc(peers));
      } else if (r == -1) {
        ++a0;
      }
    }
    if (a0) atomicAdd(&s_area0, a0);
    G::allreduce(&s_area0, 1, red_tmp);
    // two largest areas over all labels INCLUDING label 0 (the pixels where the inverse is 0), as a multiset
    int m1 = -1, m2 = -1;
    auto push = [&](int a) { if (a > m1) { m2 = m1; m1 = a; } else if (a > m2) m2 = a; };
    for (int i = G::tid(); i < n; i += G::size)
      if (L[i] == i) push(area[i]);
    if (G::tid() == 0) push(s_area0);
    for (int o = 16; o > 0; o >>= 1) {
      const int a1 = __shfl_down_sync(0xffffffffu, m1, o), a2 = __shfl_down_sync(0xffffffffu, m2, o);
      push(a1);
      push(a2);
    }
    __shared__ int wtop[G::threads / 32][2];
    if ((threadIdx.x & 31) == 0) { wtop[threadIdx.x >> 5][0] = m1; wtop[threadIdx.x >> 5][1] = m2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t1 = -1, t2 = -1;
      for (int wv = 0; wv < G::threads / 32; ++wv)
        for (int e = 0; e < 2; ++e) {
          const int a = wtop[wv][e];
          if (a > t1) { t2 = t1; t1 = a; } else if (a > t2) t2 = a;
        }
      ctop[0] = t1; ctop[1] = t2;
    }
    G::sync();
    if (threadIdx.x == 0) {   // merge the per-CTA pairs (multiset union of the two largest)
      int t1 = -1, t2 = -1;
      if constexpr (CL == 1) {
        t1 = ctop[0]; t2 = ctop[1];
      } else {
        cg::cluster_group cl = cg::this_cluster();
        for (int r = 0; r < CL; ++r) {
          const int* rc = cl.map_shared_rank(ctop, r);
          for (int e = 0; e < 2; ++e) {
            const int a = rc[e];
            if (a > t1) { t2 = t1; t1 = a; } else if (a > t2) t2 = a;
          }
        }
      }
      top2[0] = t1; top2[1] = t2;
    }
    G::sync();   // also keeps ctop alive until every CTA has read it
  }
  {
    // sorted_area[-2] if more than one label else sorted_area[-1] (textmask.py:114-118); label 0 always exists
    const int thresh = top2[1] >= 0 ? top2[1] : top2[0];
    grp_merge_labels<CL>(L, predm, merged, acc, n, rw, false, thresh);
  }

  // ---- phase 6: mask_refined[window] |= merged (textmask.py:168); windows may overlap -> atomic OR ----
  for (int i = G::tid(); i < n; i += G::size) {
    if (!merged[i]) continue;
    const int y = i / rw, x = i - y * rw;
    const size_t gp = size_t(win.y1 + y) * W + win.x1 + x;
    atomicOr(&out_words[gp >> 2], 0xffu << (8 * (gp & 3)));
  }
}

size_t refine_scratch_bytes(size_t total_px) { return total_px * (5 + 4 + 16) + 4096; }

// wins: n_wins RefineWin records on the device; idx_small / idx_large: indices into `wins` (device), split by the
// host at kRefLargePx pixels.  img / mask / out hold `H*W`-pixel planes per page (H*W % 4 == 0).
cudaError_t refine_launch(const uint8_t* d_img, const uint8_t* d_mask, int H, int W, const void* d_wins, const int* d_idx_small,
                          int n_small, const int* d_idx_large, int n_large, size_t total_px, void* scratch, int refine_mode,
                          uint8_t* d_out, cudaStream_t s) {
  if (n_small <= 0 && n_large <= 0) return cudaSuccess;
  char* p = static_cast<char*>(scratch);
  RefinePlanes P;
  P.L = reinterpret_cast<int*>(p); p += total_px * 4;
  P.acc = reinterpret_cast<int*>(p); p += total_px * 16;
  P.grey = reinterpret_cast<uint8_t*>(p); p += total_px;
  P.cand = reinterpret_cast<uint8_t*>(p); p += total_px;
  P.predm = reinterpret_cast<uint8_t*>(p); p += total_px;
  P.merged = reinterpret_cast<uint8_t*>(p); p += total_px;
  P.tmp = reinterpret_cast<uint8_t*>(p);
  const RefineWin* wins = static_cast<const RefineWin*>(d_wins);
  uint32_t* out = reinterpret_cast<uint32_t*>(d_out);
  if (n_large > 0) {
    // large windows first: they are the critical path, the single-CTA windows fill in around them
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(n_large) * kRefCluster, 1, 1);
    cfg.blockDim = dim3(kRefThreadsCluster, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kRefCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, refine_kernel<kRefCluster>, d_img, d_mask, H, W, wins, d_idx_large, P, refine_mode, out);
    if (e != cudaSuccess) return e;
  }
  if (n_small > 0)
    refine_kernel<1><<<n_small, kRefThreads, 0, s>>>(d_img, d_mask, H, W, wins, d_idx_small, P, refine_mode, out);
  return cudaGetLastError();
}
int refine_large_px() { return kRefLargePx; }
size_t refine_win_bytes() { return sizeof(RefineWin); }

}  // namespace ctd
