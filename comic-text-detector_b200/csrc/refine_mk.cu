// refine_mask on the GPU, phase-synchronous form (reference utils/textmask.py:159-169 and callees 16-132).
//
// csrc/refine.cu runs ONE cooperative kernel with a CTA (or an 8-CTA cluster) per block window and barriers between
// the phases.  Measured on the synthetic 1024^2 pages: every phase is a latency-bound sweep with one pixel per thread
// iteration, a 1 Mpx window keeps 8 SMs busy for 15 ms while the other 140 idle, and a batch of 16 pages (~50 Mpx of
// overlapping windows) cost 29 ms -- 4.5x the network.  Here every phase is its own kernel over ALL window pixels of
// the batch: the windows are cut into chunks of whole rows (<= kChunkPx pixels, table built by the host, which knows
// the window sizes), one CTA per chunk, so a giant window is spread over the whole GPU and the kernel boundary is the
// barrier.  Per-window reductions (histograms, xor sums, the two largest hole areas) go through a small per-window
// state record in global memory; per-window scalar decisions are one-CTA-per-window kernels.  Results are bit-identical to
// refine.cu and to the oracle (tests/test_gpu_refine.py runs all three).
//
// The sweeps are instruction-issue bound, not memory bound (ncu, profiles/r02_refine_full_summary.txt), so the binary planes
// are handled as BIT masks wherever a neighbourhood is involved: warp ballots pack 32 pixels per shared-memory word and one
// thread per (half) word does the work of 16 - 32 pixels with funnel shifts, ANDs and popcounts -- the erosions of phase 0,
// the dilation, the run contacts of the labelling and the per-label sums (one update per RUN, not per pixel).
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>

#include "kernels.h"

namespace ctd {

namespace {

constexpr int kThreads = 256;

struct RefineWin {
  int x1, y1, x2, y2;
  long long off;
  int page, pad;
};
struct Chunk { int win, y0, rows, pad; };

struct WinState {
  int hist[4][256];               // [0] grey of the eroded-mask pixels, [1..3] B, G, R of the whole window
  unsigned long long xs[12];      // xor sums: [k][pos/neg] for 3 colours, then 3 channels
  int lo[3], hi[3], otsu_t[3], ncol;
  int nproc, proc_kind[4], proc_neg[4];
  int area0, max1, cnt1, max2;
  int pad[2];
};

struct Ctx {
  const uint8_t* img_all;
  const uint8_t* mask_all;
  uint32_t* out_all;
  const RefineWin* wins;
  const Chunk* chunks;
  WinState* st;
  int H, W, mode;
  // planes (window-pixel indexed)
  int* L;
  int* acc;
  uint8_t *grey, *cand, *predm, *merged, *tmp;   // cand: unused here (kept: the scratch layout is shared with refine.cu)
};

struct View {
  RefineWin win;
  int w, rw, rh, y0, rows, i0, cnt, aligned;   // aligned: the chunk starts on a 4-byte boundary of the window planes
  const uint8_t* img;
  const uint8_t* mask;
};

__device__ __forceinline__ View view_of(const Ctx& c, int chunk) {
  View v;
  const Chunk ch = c.chunks[chunk];
  v.w = ch.win;
  v.win = c.wins[ch.win];
  v.rw = v.win.x2 - v.win.x1;
  v.rh = v.win.y2 - v.win.y1;
  v.y0 = ch.y0;
  v.rows = ch.rows;
  v.i0 = ch.y0 * v.rw;
  v.cnt = ch.rows * v.rw;
  v.aligned = ch.pad & 1;
  v.img = c.img_all + size_t(v.win.page) * c.H * c.W * 3;
  v.mask = c.mask_all + size_t(v.win.page) * c.H * c.W;
  return v;
}

// ---- union-find on a window's L plane (other CTAs update it: parent reads bypass L1) -------------------------------
__device__ __forceinline__ int uf_find(const int* L, int a) {
  int p = __ldcg(L + a);
  while (p != a) {
    a = p;
    p = __ldcg(L + a);
  }
  return a;
}
__device__ __forceinline__ int uf_find_compress(int* L, int a) {
  const int r = uf_find(L, a);
  while (a != r) {
    const int p = __ldcg(L + a);
    if (p <= r) break;
    atomicMin(&L[a], r);
    a = p;
  }
  return r;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  bool done;
  do {
    a = uf_find(L, a);
    b = uf_find(L, b);
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

__device__ __forceinline__ void hist_add(int* hist, int bin) {
  // flat regions (page background, solid strokes): the whole warp hits one bin -> one atomic, no match.any
  const int b0 = __shfl_sync(0xffffffffu, bin, 0);
  if (__all_sync(0xffffffffu, bin == b0)) {
    if ((threadIdx.x & 31) == 0 && bin >= 0) atomicAdd(&hist[bin], 32);
    return;
  }
  // mixed warp: plain shared-memory atomics (the hardware serialises equal bins; measured faster than a match.any vote)
  if (bin >= 0) atomicAdd(&hist[bin], 1);
}

// i = q * d + r for 0 <= i < 2^24 (exact in float; |q error| <= 1 before the correction), else integer division.
// Every sweep below turns window-local pixel indices into (y, x): a runtime integer division per pixel was a
// quarter of the instructions of the lean kernels.
struct DivW { int d; float inv; bool fast; };
__device__ __forceinline__ DivW make_div(int d, int n) { return DivW{d, __frcp_rn(float(d)), n < (1 << 24)}; }
__device__ __forceinline__ void divmod(int i, const DivW& dv, int& q, int& r) {
  if (dv.fast) {
    q = __float2int_rz(__int2float_rn(i) * dv.inv);
    r = i - q * dv.d;
    if (r < 0) { --q; r += dv.d; } else if (r >= dv.d) { ++q; r -= dv.d; }
  } else {
    q = i / dv.d;
    r = i - q * dv.d;
  }
}

// ---- phase 0: grey, pred mask (cross erosion > 60), merged = 0, histograms ------------------------------------------
constexpr int kU = 4;   // pixels per thread and outer iteration: the loads of all kU pixels are issued before the first use
constexpr int kChunkPxFwd = 8192;                       // = kChunkPx (defined with the labelling kernels below)
constexpr int kExtWords = (3 * kChunkPxFwd) / 32 + 4;   // chunk (<= kChunkPx) + one halo row above and below
__device__ __forceinline__ unsigned bits_from(const unsigned* M, int pos) {   // 32 bits starting at pixel `pos` (< 0 reads 0)
  if (pos <= -32) return 0u;
  if (pos < 0) return M[0] << (-pos);
  const int w = pos >> 5, sft = pos & 31;
  return __funnelshift_r(M[w], M[w + 1], sft);
}
// The two erosions of the mask crop are threshold tests of a minimum: min over the cross > 60 <=> NO pixel of the cross is
// <= 60.  So the chunk's rows plus one row above and below are packed into two "bad pixel" bit masks (mask <= 60,
// mask <= 127) by warp ballots -- ONE mask load per pixel instead of nine bounds-checked ones -- and one thread per
// 32-pixel word ORs the 5 / 9 shifted views (row ends masked; outside the window reads 0 = not bad = BORDER_CONSTANT +inf).
__global__ void __launch_bounds__(kThreads) k_phase0(Ctx c) {
  __shared__ int sh[4][256];
  __shared__ unsigned N60[kExtWords], N127[kExtWords];
  __shared__ unsigned F60[kChunkPxFwd / 32 + 1], F127[kChunkPxFwd / 32 + 1];
  const View v = view_of(c, blockIdx.x);
  for (int i = threadIdx.x; i < 1024; i += kThreads) (&sh[0][0])[i] = 0;
  for (int i = threadIdx.x; i < kExtWords; i += kThreads) { N60[i] = 0u; N127[i] = 0u; }
  __syncthreads();
  uint8_t* grey = c.grey + v.win.off;
  uint8_t* predm = c.predm + v.win.off;
  uint8_t* merged = c.merged + v.win.off;
  const DivW dv = make_div(v.rw, 3 * kChunkPxFwd + 1);
  const int ystart = v.y0 > 0 ? v.y0 - 1 : 0;
  const int yend = min(v.y0 + v.rows + 1, v.rh);
  const int ext = (yend - ystart) * v.rw;
  const int off = (v.y0 - ystart) * v.rw;
  for (int e0 = 0; e0 < ext; e0 += kThreads) {
    const int e = e0 + threadIdx.x;
    int mv = 255;
    if (e < ext) {
      int ye, xe;
      divmod(e, dv, ye, xe);
      mv = v.mask[size_t(v.win.y1 + ystart + ye) * c.W + v.win.x1 + xe];
    }
    const unsigned b60 = __ballot_sync(0xffffffffu, mv <= 60), b127 = __ballot_sync(0xffffffffu, mv <= 127);
    if ((threadIdx.x & 31) == 0) { N60[e >> 5] = b60; N127[e >> 5] = b127; }
  }
  __syncthreads();
  for (int w = threadIdx.x; w * 32 < v.cnt; w += kThreads) {
    const int k0 = w * 32;
    int yl, x0;
    divmod(k0, dv, yl, x0);
    unsigned rs = 0u;
    for (int j = x0 == 0 ? 0 : v.rw - x0; j < 32; j += v.rw) rs |= 1u << j;
    int xe = x0 + 32;
    if (xe >= v.rw) xe %= v.rw;
    const unsigned re = (rs >> 1) | (xe == 0 ? 0x80000000u : 0u);
    const int p = k0 + off;
    // cross (textmask.py:86-89, MORPH_CROSS 3x3) on the <= 60 mask
    F60[w] = bits_from(N60, p) | bits_from(N60, p - v.rw) | bits_from(N60, p + v.rw) | (bits_from(N60, p - 1) & ~rs) |
             (bits_from(N60, p + 1) & ~re);
    // full 3x3 (textmask.py:60, the eroded mask of get_topk_color) on the <= 127 mask
    F127[w] = bits_from(N127, p) | bits_from(N127, p - v.rw) | bits_from(N127, p + v.rw) |
              ((bits_from(N127, p - 1) | bits_from(N127, p - v.rw - 1) | bits_from(N127, p + v.rw - 1)) & ~rs) |
              ((bits_from(N127, p + 1) | bits_from(N127, p - v.rw + 1) | bits_from(N127, p + v.rw + 1)) & ~re);
  }
  __syncthreads();
  const DivW dvw = make_div(v.rw, v.rw * v.rh);
  for (int k0 = 0; k0 < v.cnt; k0 += kThreads * kU) {
    int b[kU], g[kU], r[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int k = k0 + u * kThreads + threadIdx.x;
      b[u] = -1; g[u] = -1; r[u] = -1;
      if (k < v.cnt) {
        int y, x;
        divmod(v.i0 + k, dvw, y, x);
        const size_t gp = size_t(v.win.y1 + y) * c.W + v.win.x1 + x;
        b[u] = v.img[gp * 3]; g[u] = v.img[gp * 3 + 1]; r[u] = v.img[gp * 3 + 2];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int k = k0 + u * kThreads + threadIdx.x;
      const bool in = k < v.cnt;
      int gr = 0;
      bool core = false;
      if (in) {
        const int i = v.i0 + k;
        gr = (b[u] * 1868 + g[u] * 9617 + r[u] * 4899 + 8192) >> 14;  // cv2.COLOR_BGR2GRAY, 8u fixed point
        grey[i] = (uint8_t)gr;
        predm[i] = ((F60[k >> 5] >> (k & 31)) & 1u) ? 0 : 255;          // cross erosion > 60 (textmask.py:86-89)
        merged[i] = 0;
        core = !((F127[k >> 5] >> (k & 31)) & 1u);                       // 3x3 erosion > 127 (textmask.py:60)
      }
      if (k0 + u * kThreads < v.cnt) {   // CTA-uniform: skip the histogram votes of iterations past the chunk
        hist_add(sh[1], b[u]);
        hist_add(sh[2], g[u]);
        hist_add(sh[3], r[u]);
        hist_add(sh[0], core ? gr : -1);
      }
    }
  }
  __syncthreads();
  int* gh = &c.st[v.w].hist[0][0];
  for (int i = threadIdx.x; i < 1024; i += kThreads) {
    const int val = (&sh[0][0])[i];
    if (val) atomicAdd(&gh[i], val);
  }
}

// ---- phase 1 (one CTA per window): np.histogram(bins=255), top-k colours, Otsu ---------------------------------------
constexpr int kDecideThreads = 64;   // mostly serial per-window work: small CTAs so that many windows are resident per SM
__global__ void __launch_bounds__(kDecideThreads) k_decide1(Ctx c) {
  __shared__ int cnt255[256];
  __shared__ int order[256];
  __shared__ double edges[256];
  __shared__ int s_first, s_last, s_total;
  __shared__ int hist_s[4][256];   // the serial loops below read every bin: from shared memory, not one global load per step
  WinState& st = c.st[blockIdx.x];
  const RefineWin win = c.wins[blockIdx.x];
  const int n = (win.x2 - win.x1) * (win.y2 - win.y1);
  for (int i = threadIdx.x; i < 1024; i += kDecideThreads) (&hist_s[0][0])[i] = (&st.hist[0][0])[i];
  const int* hist_g = hist_s[0];
  for (int i = threadIdx.x; i < 256; i += kDecideThreads) cnt255[i] = 0;
  __syncthreads();
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
    for (int i = threadIdx.x; i < 256; i += kDecideThreads) edges[i] = (i == 255) ? le : __dadd_rn(__dmul_rn((double)i, step), fe);
    __syncthreads();
    for (int v = threadIdx.x; v < 256; v += kDecideThreads) {
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
  for (int b = threadIdx.x; b < 255; b += kDecideThreads) {
    int rank = 0;
    const int cb = cnt255[b];
    for (int q = 0; q < 255; ++q) rank += (cnt255[q] > cb) || (cnt255[q] == cb && q < b);
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
    st.ncol = nt;
    for (int t = 0; t < nt; ++t) {
      const double c_top = fmin(top[t] + 30.0, 255.0);
      const double c_bot = c_top - 60.0;
      // cv2.inRange with float bounds on 8u data: cvRound (half to even) + saturate
      st.lo[t] = (int)fmin(fmax(rint(c_bot), 0.0), 255.0);
      st.hi[t] = (int)fmin(fmax(rint(c_top), 0.0), 255.0);
    }
  }
  if (threadIdx.x >= 32 && threadIdx.x < 35) {
    // cv2.threshold(..., THRESH_OTSU): getThreshVal_Otsu_8u
    const int* hh = hist_s[1 + threadIdx.x - 32];
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
    st.otsu_t[threadIdx.x - 32] = max_val;
  }
}

// ---- phase 2: xor sums of every candidate and of its negative against the mask crop -----------------------------------
__global__ void __launch_bounds__(kThreads) k_xor(Ctx c) {
  __shared__ unsigned long long sx[12];
  const View v = view_of(c, blockIdx.x);
  const WinState& st = c.st[v.w];
  if (threadIdx.x < 12) sx[threadIdx.x] = 0ull;
  __syncthreads();
  const uint8_t* grey = c.grey + v.win.off;
  const int ncol = st.ncol;
  int lo[3], hi[3], ot[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { lo[k] = st.lo[k]; hi[k] = st.hi[k]; ot[k] = st.otsu_t[k]; }
  // per thread <= kChunkPx / kThreads pixels x 255: 32-bit partial sums.  Only the POSITIVE candidates are summed: for
  // t in {0, 255}, (255 - t) ^ m == 255 - (t ^ m), so the negative's sum is 255 * pixels - the positive's sum.
  unsigned pos[6], npx = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) pos[k] = 0u;
  const DivW dv = make_div(v.rw, v.rw * v.rh);
  for (int k0 = 0; k0 < v.cnt; k0 += kThreads * kU) {
    int mk[kU], gr[kU], ch[kU][3];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int k = k0 + u * kThreads + threadIdx.x;
      mk[u] = -1;
      if (k < v.cnt) {
        int y, x;
        divmod(v.i0 + k, dv, y, x);
        const size_t gp = size_t(v.win.y1 + y) * c.W + v.win.x1 + x;
        mk[u] = v.mask[gp];
        gr[u] = grey[v.i0 + k];
#pragma unroll
        for (int q = 0; q < 3; ++q) ch[u][q] = v.img[gp * 3 + q];
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (mk[u] < 0) continue;
      ++npx;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int t = (gr[u] >= lo[k] && gr[u] <= hi[k]) ? 255 : 0;     // only read for k < ncol
        pos[k] += (unsigned)(t ^ mk[u]);
        const int t2 = ch[u][k] > ot[k] ? 255 : 0;
        pos[3 + k] += (unsigned)(t2 ^ mk[u]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const bool used = k >= 3 || k < ncol;
    unsigned long long sp = used ? pos[k] : 0ull, sn = used ? 255ull * npx - pos[k] : 0ull;
    for (int o = 16; o > 0; o >>= 1) {
      sp += __shfl_down_sync(0xffffffffu, sp, o);
      sn += __shfl_down_sync(0xffffffffu, sn, o);
    }
    // xs layout: [2k] positive, [2k+1] negative for the 3 colours, then the 3 channels
    if ((threadIdx.x & 31) == 0) {
      if (sp) atomicAdd(&sx[2 * k], sp);
      if (sn) atomicAdd(&sx[2 * k + 1], sn);
    }
  }
  __syncthreads();
  if (threadIdx.x < 12 && sx[threadIdx.x]) atomicAdd(&c.st[v.w].xs[threadIdx.x], sx[threadIdx.x]);
}

// candidate order (one thread per window): minxor_thresh + sort
__global__ void k_decide2(Ctx c, int n_wins) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_wins) return;
  WinState& st = c.st[w];
  const unsigned long long* xs = st.xs;
  // minxor_thresh (textmask.py:29-41): negative wins only if strictly smaller
  unsigned long long best[4];
  int kind[4], neg[4], np_ = 0;
  for (int k = 0; k < st.ncol; ++k) {
    const bool ng = xs[2 * k + 1] < xs[2 * k];
    best[np_] = ng ? xs[2 * k + 1] : xs[2 * k];
    kind[np_] = k; neg[np_] = ng; ++np_;
  }
  // Otsu: best channel (stable sort by xor sum -> first minimum in B,G,R order)  (textmask.py:43-54)
  int bc = 0, bneg = 0;
  unsigned long long bv = ~0ull;
  for (int ch = 0; ch < 3; ++ch) {
    const bool ng = xs[6 + 2 * ch + 1] < xs[6 + 2 * ch];
    const unsigned long long val = ng ? xs[6 + 2 * ch + 1] : xs[6 + 2 * ch];
    if (val < bv) { bv = val; bc = ch; bneg = ng; }
  }
  best[np_] = bv; kind[np_] = 3 + bc; neg[np_] = bneg; ++np_;
  // mask_list.sort(key=xor_sum) (textmask.py:74): stable insertion sort
  for (int i = 1; i < np_; ++i) {
    const unsigned long long val = best[i];
    const int kk = kind[i], nn = neg[i];
    int j = i - 1;
    while (j >= 0 && best[j] > val) { best[j + 1] = best[j]; kind[j + 1] = kind[j]; neg[j + 1] = neg[j]; --j; }
    best[j + 1] = val; kind[j + 1] = kk; neg[j + 1] = nn;
  }
  for (int i = 0; i < np_; ++i) { st.proc_kind[i] = kind[i]; st.proc_neg[i] = neg[i]; }
  st.nproc = np_;
  st.area0 = 0; st.max1 = -1; st.cnt1 = 0; st.max2 = -1;
}

// ---- labelling of a source plane: candidate `round` (0..3) or, round == 4, the inverse of `merged` (hole filling) -----
// Level 1, one CTA per chunk (whole rows, <= kChunkPx pixels), everything in SHARED memory: source pixels (coalesced),
// run starts by warp ballot, seams between warps and the contacts between the rows of the chunk united in a shared
// union-find, one root per chunk-local component written to L (tmp = 1 marks those roots: they are the chain nodes of
// the global forest).  Level 2: only the first row of every chunk issues global unions with the row above it.
// Level 3: compress from the chain nodes, then every pixel takes its (chunk-local) parent's root.
constexpr int kChunkPx = 8192;
static_assert(kChunkPx == kChunkPxFwd, "kChunkPxFwd mirrors kChunkPx");
constexpr int kLabelThreads = 512;

__device__ __forceinline__ int suf_find(const int* L, int a) {
  int p = L[a];
  while (p != a) {
    a = p;
    p = L[a];
  }
  return a;
}
// find with path halving: every visited node is re-pointed at its grandparent (atomicMin: parents have smaller indices
// than their children, and concurrent unions only ever lower a root's entry).  The walks of the union phase were ~10 hops.
__device__ __forceinline__ int suf_find_halve(int* L, int a) {
  int p = L[a];
  while (p != a) {
    const int gp = L[p];
    if (gp != p) atomicMin(&L[a], gp);
    a = p;
    p = gp;
  }
  return a;
}
__device__ __forceinline__ void suf_union(int* L, int a, int b) {
  bool done;
  do {
    a = suf_find_halve(L, a);
    b = suf_find_halve(L, b);
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

__global__ void __launch_bounds__(kLabelThreads, 3) k_label_local(Ctx c, int round) {
  __shared__ int Ls[kChunkPx];
  __shared__ unsigned Mw[kChunkPx / 32 + 2];      // foreground bits, 32 pixels per word (+ zero padding)
  __shared__ unsigned Sw[kChunkPx / 32];          // run-start bits (the only nodes of the chunk-local forest)
  __shared__ unsigned Gw[kChunkPx / 32];          // foreground & not yet merged & predicted     ("gain" pixels)
  __shared__ unsigned Bw[kChunkPx / 32];          // foreground & not yet merged & not predicted ("loss" pixels)
  __shared__ int s_fg;
  const View v = view_of(c, blockIdx.x);
  WinState& st = c.st[v.w];
  if (round < 4 && round >= st.nproc) return;
  for (int i = threadIdx.x; i < kChunkPx / 32 + 2; i += kLabelThreads) Mw[i] = 0u;
  if (threadIdx.x == 0) s_fg = 0;
  __syncthreads();
  uint8_t* rootflag = c.tmp + v.win.off;
  int* L = c.L + v.win.off;
  int* acc = c.acc + 4 * v.win.off;
  const int n = v.rw * v.rh;
  int* area = acc; int* gain = acc + n; int* loss = acc + 2 * n; int* maxi = acc + 3 * n;
  const uint8_t* grey = c.grey + v.win.off;
  const uint8_t* merged = c.merged + v.win.off;
  const uint8_t* predm = c.predm + v.win.off;
  int kind = 0, neg = 0, lo = 0, hi = 0, ot = 0;
  if (round < 4) {
    kind = st.proc_kind[round]; neg = st.proc_neg[round];
    if (kind < 3) { lo = st.lo[kind]; hi = st.hi[kind]; } else ot = st.otsu_t[kind - 3];
  }
  const int lane = threadIdx.x & 31;
  const DivW dv = make_div(v.rw, kChunkPx + 1);   // chunk-local indices: k = row * rw + x, k < kChunkPx
  constexpr int kIt = kChunkPx / kLabelThreads;    // 16 pixels per thread
  // pass 1: source value, `merged` and `pred` of every pixel (coalesced; the loads of a batch are issued before their
  // first use) -> foreground / gain / loss BIT masks and the run starts inside each warp's 32 consecutive pixels
  if (v.aligned) {
    // pass 1, vector form (the chunk starts on a 4-byte boundary of the byte planes): FOUR consecutive pixels per thread from
    // one 32-bit load per plane; the three 4-bit results of 8 lanes are OR-reduced into the 32-pixel words (redux.sync) --
    // a quarter of the loads / address arithmetic and no per-pixel ballots or shared-memory stores.
    const uint32_t* mg32 = reinterpret_cast<const uint32_t*>(merged + v.i0);
    const uint32_t* pd32 = reinterpret_cast<const uint32_t*>(predm + v.i0);
    const uint32_t* gr32 = reinterpret_cast<const uint32_t*>(grey + v.i0);
    const int sh = 4 * (lane & 7);
    const unsigned grp = 0xffu << (8 * (lane >> 3));
#pragma unroll 1
    for (int it = 0; it < kChunkPx / (4 * kLabelThreads); ++it) {
      if (it * 4 * kLabelThreads >= v.cnt) break;      // CTA-uniform
      const int k = 4 * (it * kLabelThreads + int(threadIdx.x));
      unsigned fg4 = 0u, g4 = 0u, b4 = 0u;
      if (k < v.cnt) {
        const unsigned m4 = mg32[k >> 2], p4 = pd32[k >> 2];
        unsigned s4 = m4;
        if (round < 4) {
          if (kind < 3) {
            s4 = gr32[k >> 2];
          } else {
            s4 = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (k + j < v.cnt) {
                int yl, x;
                divmod(k + j, dv, yl, x);
                s4 |= unsigned(v.img[(size_t(v.win.y1 + v.y0 + yl) * c.W + v.win.x1 + x) * 3 + (kind - 3)]) << (8 * j);
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k + j < v.cnt) {
            const int sbv = int((s4 >> (8 * j)) & 0xffu), mb = int((m4 >> (8 * j)) & 0xffu), pb = int((p4 >> (8 * j)) & 0xffu);
            bool fg;
            if (round == 4) {
              fg = sbv == 0;
            } else {
              const bool t = kind < 3 ? (sbv >= lo && sbv <= hi) : (sbv > ot);
              fg = neg ? !t : t;
            }
            const bool un = fg && mb == 0;
            fg4 |= unsigned(fg) << j;
            g4 |= unsigned(un && pb != 0) << j;
            b4 |= unsigned(un && pb == 0) << j;
          }
        }
      }
      const unsigned M = __reduce_or_sync(grp, fg4 << sh), G = __reduce_or_sync(grp, g4 << sh), B = __reduce_or_sync(grp, b4 << sh);
      if ((lane & 7) == 0) { const int w = k >> 5; Mw[w] = M; Gw[w] = G; Bw[w] = B; }
    }
    __syncthreads();
    // run starts: a foreground pixel whose left neighbour in the same row AND word is not foreground
    for (int w = threadIdx.x; w * 32 < v.cnt; w += kLabelThreads) {
      const unsigned M = Mw[w];
      int yl, x0;
      divmod(w * 32, dv, yl, x0);
      unsigned rs = 0u;
      for (int j = x0 == 0 ? 0 : v.rw - x0; j < 32; j += v.rw) rs |= 1u << j;
      const unsigned S = M & (~(M << 1) | rs);
      Sw[w] = S;
      unsigned f = S;
      while (f) {
        const int k = w * 32 + __ffs(f) - 1;
        f &= f - 1u;
        Ls[k] = k;
      }
    }
  } else {
    constexpr int kB = 4;
    int xrun, xstep;                                 // x = k % rw without a division per pixel: k advances by kLabelThreads
    {
      int q;
      divmod(int(threadIdx.x), dv, q, xrun);
      divmod(kLabelThreads, dv, q, xstep);
    }
  #pragma unroll 1
    for (int ub = 0; ub < kIt; ub += kB) {
      if (ub * kLabelThreads >= v.cnt) break;        // CTA-uniform
      int raw[kB], mg[kB], pd[kB];
  #pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int k = (ub + u) * kLabelThreads + threadIdx.x;
        raw[u] = 0; mg[u] = 1; pd[u] = 0;
        if (k < v.cnt) {
          const int i = v.i0 + k;
          mg[u] = merged[i];
          pd[u] = predm[i];
          if (round == 4) raw[u] = mg[u];
          else if (kind < 3) raw[u] = grey[i];
          else {
            int yl, x;
            divmod(k, dv, yl, x);
            raw[u] = v.img[(size_t(v.win.y1 + v.y0 + yl) * c.W + v.win.x1 + x) * 3 + (kind - 3)];
          }
        }
      }
  #pragma unroll
      for (int u = 0; u < kB; ++u) {
        const int k0 = (ub + u) * kLabelThreads;
        if (k0 >= v.cnt) break;                       // CTA-uniform
        const int k = k0 + threadIdx.x;
        const bool in = k < v.cnt;
        int sv = 0;
        const int x = xrun;                           // x of pixel k, carried from iteration to iteration
        xrun += xstep;
        if (xrun >= v.rw) xrun -= v.rw;
        if (in) {
          if (round == 4) {
            sv = raw[u] ? 0 : 255;
          } else {
            const int tv = kind < 3 ? ((raw[u] >= lo && raw[u] <= hi) ? 255 : 0) : (raw[u] > ot ? 255 : 0);
            sv = neg ? 255 - tv : tv;
          }
        }
        const bool fg = in && sv != 0;
        const bool un = fg && mg[u] == 0;
        const unsigned m = __ballot_sync(0xffffffffu, fg);
        const unsigned gb = __ballot_sync(0xffffffffu, un && pd[u] != 0);
        const unsigned lb = __ballot_sync(0xffffffffu, un && pd[u] == 0);
        // a run starts at a foreground pixel whose left neighbour (same row, same warp) is not foreground
        const bool starts = fg && (lane == 0 || x == 0 || !((m >> (lane - 1)) & 1u));
        const unsigned sb = __ballot_sync(0xffffffffu, starts);
        if (lane == 0) { Mw[k >> 5] = m; Sw[k >> 5] = sb; Gw[k >> 5] = gb; Bw[k >> 5] = lb; }
        if (starts) Ls[k] = k;    // forest nodes = run starts; any other foreground pixel maps to its run start via start_of()
      }
    }
  }
  __syncthreads();
  // pass 2: seams between warps, contacts with the row above inside the chunk -- on the foreground BIT masks, two
  // threads per 32-pixel word: the neighbour tests of 32 pixels are a handful of shifts and ANDs, and only the pixels
  // that really start a (run x upper run) contact walk the union-find (first version: every foreground pixel tested
  // its four neighbours with byte loads; half of the kernel's instructions, ncu).
  // run start of foreground pixel p: the highest run-start bit at or below p in its word (runs restart at every word)
  auto start_of = [&](int pos) -> int {
    return (pos & ~31) + 31 - __clz(Sw[pos >> 5] & (0xffffffffu >> (31 - (pos & 31))));
  };
  auto bits_at = [&](int pos) -> unsigned {        // 32 foreground bits starting at pixel `pos` (pixels < 0 read as 0)
    if (pos <= -32) return 0u;
    if (pos < 0) return Mw[0] << (-pos);
    const int w = pos >> 5, sft = pos & 31;
    return __funnelshift_r(Mw[w], Mw[w + 1], sft);
  };
  for (int hw = threadIdx.x; hw * 16 < v.cnt; hw += kLabelThreads) {
    const int w = hw >> 1;
    const unsigned half = (hw & 1) ? 0xffff0000u : 0x0000ffffu;
    const unsigned cur = Mw[w];
    if (!(cur & half)) continue;
    const int k0 = w * 32;
    int yl, x0;
    divmod(k0, dv, yl, x0);
    unsigned rs = 0u;                               // bits whose pixel is the FIRST of its row (x == 0)
    for (int j = x0 == 0 ? 0 : v.rw - x0; j < 32; j += v.rw) rs |= 1u << j;
    int xe = x0 + 32;                               // x of the pixel after this word
    if (xe >= v.rw) xe %= v.rw;
    const unsigned re = (rs >> 1) | (xe == 0 ? 0x80000000u : 0u);   // bits whose pixel is the LAST of its row
    const unsigned lft = bits_at(k0 - 1);           // fg(k - 1)
    // seam: the run labelling of pass 1 restarts at every word
    if (!(hw & 1) && (cur & 1u) && !(rs & 1u) && (lft & 1u)) suf_union(Ls, k0, start_of(k0 - 1));
    if (k0 + 32 <= v.rw) continue;                  // the whole word lies in the first row of the chunk
    const unsigned vup = (k0 >= v.rw ? 0xffffffffu : (0xffffffffu << (v.rw - k0))) & half;   // pixels that have a row above
    const unsigned up = bits_at(k0 - v.rw), upl = bits_at(k0 - v.rw - 1), upr = bits_at(k0 - v.rw + 1);
    // pixel and the pixel above are foreground: only the first pixel of each (current run x upper run) overlap unions
    unsigned f = cur & up & vup & (rs | ~lft | ~upl);
    while (f) {
      const int j = __ffs(f) - 1;
      f &= f - 1u;
      suf_union(Ls, start_of(k0 + j), start_of(k0 + j - v.rw));
    }
    const unsigned nb = cur & ~up & vup;
    f = nb & upl & ~rs;                             // diagonal contacts when the pixel above is background
    while (f) {
      const int j = __ffs(f) - 1;
      f &= f - 1u;
      suf_union(Ls, start_of(k0 + j), start_of(k0 + j - v.rw - 1));
    }
    f = nb & upr & ~re;
    while (f) {
      const int j = __ffs(f) - 1;
      f &= f - 1u;
      suf_union(Ls, start_of(k0 + j), start_of(k0 + j - v.rw + 1));
    }
  }
  __syncthreads();
  // pass 3a: flatten the forest.  Its nodes are the run starts only (every other foreground pixel points at its run
  // start and is never re-parented): after this pass every run start points straight at its root, so the root of ANY
  // foreground pixel is Ls[Ls[k]] -- two loads instead of a walk (the walks were 10 hops on average, ncu).  The roots
  // zero their per-label sums here (area, gain, loss, max pixel index).
  int fgc = 0;
  for (int hw = threadIdx.x; hw * 16 < v.cnt; hw += kLabelThreads) {
    const unsigned half = (hw & 1) ? 0xffff0000u : 0x0000ffffu;
    unsigned f = Sw[hw >> 1] & half;
    fgc += __popc(Mw[hw >> 1] & half);
    const int k0 = (hw >> 1) * 32;
    while (f) {
      const int s0 = k0 + __ffs(f) - 1;
      f &= f - 1u;
      const int r = suf_find(Ls, s0);
      if (r != s0) {
        atomicMin(&Ls[s0], r);
      } else {
        const int gi = v.i0 + s0;
        area[gi] = 0; gain[gi] = 0; loss[gi] = 0; maxi[gi] = -1;
      }
    }
  }
  if (round == 4) {   // label 0 of the inverse = the pixels already in `merged`
    for (int o = 16; o > 0; o >>= 1) fgc += __shfl_down_sync(0xffffffffu, fgc, o);
    if (lane == 0 && fgc) atomicAdd(&s_fg, fgc);
  }
  __syncthreads();
  if (round == 4 && threadIdx.x == 0) {
    const int a0 = v.cnt - s_fg;
    if (a0) atomicAdd(&st.area0, a0);
  }
  // pass 3b: chunk-local root of every pixel to global memory (window-local pixel indices; -1 = background)
  for (int k = threadIdx.x; k < v.cnt; k += kLabelThreads) {
    const int r = ((Mw[k >> 5] >> (k & 31)) & 1u) ? Ls[start_of(k)] : -1;
    L[v.i0 + k] = r < 0 ? -1 : v.i0 + r;
    rootflag[v.i0 + k] = (r == k) ? 1 : 0;
  }
  // pass 3c: per-label sums, one update per RUN (popcounts of the run's bits in the foreground / gain / loss masks) into
  // the sums of its chunk-local root.  k_flat1 adds the sums of the chunk roots of a multi-chunk window to their global
  // root; there is no per-pixel accumulation sweep any more.
  for (int hw = threadIdx.x; hw * 16 < v.cnt; hw += kLabelThreads) {
    const int w = hw >> 1;
    const unsigned M = Mw[w], S = Sw[w];
    unsigned f = S & ((hw & 1) ? 0xffff0000u : 0x0000ffffu);
    if (!f) continue;
    const unsigned G = Gw[w], B = Bw[w];
    const int k0 = w * 32;
    while (f) {
      const int sbit = __ffs(f) - 1;
      f &= f - 1u;
      const unsigned stop = (~M | S) & (0xfffffffeu << sbit);          // first position after the run
      const int e = stop ? __ffs(stop) - 2 : 31;                       // last pixel of the run
      const unsigned rm = (e == 31 ? 0xffffffffu : ((2u << e) - 1u)) & ~((1u << sbit) - 1u);
      const int gi = v.i0 + Ls[k0 + sbit];
      atomicAdd(&area[gi], __popc(rm));
      atomicMax(&maxi[gi], v.i0 + k0 + e);
      const int g_ = __popc(rm & G), l_ = __popc(rm & B);
      if (g_) atomicAdd(&gain[gi], g_);
      if (l_) atomicAdd(&loss[gi], l_);
    }
  }
}
// level 2: the first row of every chunk against the last row of the chunk above
constexpr int kBorderThreads = 256;  // (64-thread CTAs measured 3x slower: a window row is up to thousands of pixels)
__global__ void __launch_bounds__(kBorderThreads) k_union_border(Ctx c, int round) {
  const View v = view_of(c, blockIdx.x);
  if (round < 4 && round >= c.st[v.w].nproc) return;
  if (v.y0 == 0) return;
  int* L = c.L + v.win.off;   // foreground <=> L >= 0 (written for every pixel by k_label_local)
  for (int x = threadIdx.x; x < v.rw; x += int(blockDim.x)) {
    const int i = v.i0 + x;
    if (__ldcg(L + i) < 0) continue;
    const int up = i - v.rw;
    if (__ldcg(L + up) >= 0) {
      const bool first = x == 0 || __ldcg(L + i - 1) < 0 || __ldcg(L + up - 1) < 0;
      if (first) uf_union(L, i, up);
    } else {
      if (x > 0 && __ldcg(L + up - 1) >= 0) uf_union(L, i, up - 1);
      if (x + 1 < v.rw && __ldcg(L + up + 1) >= 0) uf_union(L, i, up + 1);
    }
  }
}
// level 3a: compress from the chain nodes (chunk-local roots); 3b: every pixel takes its parent's root
__global__ void __launch_bounds__(kThreads) k_flat1(Ctx c, int round) {
  const View v = view_of(c, blockIdx.x);
  if (round < 4 && round >= c.st[v.w].nproc) return;
  const uint8_t* rootflag = c.tmp + v.win.off;
  int* L = c.L + v.win.off;
  int* acc = c.acc + 4 * v.win.off;
  const int n = v.rw * v.rh;
  int* area = acc; int* gain = acc + n; int* loss = acc + 2 * n; int* maxi = acc + 3 * n;
  constexpr int U = 8;
  for (int k0 = 0; k0 < v.cnt; k0 += kThreads * U) {
    uint8_t rf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * kThreads + threadIdx.x;
      rf[u] = k < v.cnt ? rootflag[v.i0 + k] : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!rf[u]) continue;
      // chunk-local root: point it straight at its global root and hand its sums (complete since k_label_local) over
      const int cr = v.i0 + k0 + u * kThreads + threadIdx.x;
      const int g = uf_find_compress(L, cr);
      if (g != cr) {
        atomicAdd(&area[g], area[cr]);
        atomicMax(&maxi[g], maxi[cr]);
        const int g_ = gain[cr], l_ = loss[cr];
        if (g_) atomicAdd(&gain[g], g_);
        if (l_) atomicAdd(&loss[g], l_);
      }
    }
  }
}
// ---- merge step (textmask.py:92-108 / 118-131) -------------------------------------------------------------------------
// hole filling only: the two largest areas over all labels incl. label 0, as a multiset (max1 with its multiplicity, max2)
__global__ void __launch_bounds__(kThreads) k_top_a(Ctx c) {
  const View v = view_of(c, blockIdx.x);
  WinState& st = c.st[v.w];
  const int* L = c.L + v.win.off;
  const int* area = c.acc + 4 * v.win.off;
  const uint8_t* rootflag = c.tmp + v.win.off;
  int m = -1;
  constexpr int U = 8;
  for (int k0 = 0; k0 < v.cnt; k0 += kThreads * U) {
    uint8_t rf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * kThreads + threadIdx.x;
      rf[u] = k < v.cnt ? rootflag[v.i0 + k] : 0;     // chunk-local roots: the only candidates for a global root
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = v.i0 + k0 + u * kThreads + threadIdx.x;
      if (rf[u] && L[i] == i) m = max(m, area[i]);
    }
  }
  if (v.y0 == 0 && threadIdx.x == 0) m = max(m, st.area0);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(&st.max1, m);
}
__global__ void __launch_bounds__(kThreads) k_top_b(Ctx c) {
  const View v = view_of(c, blockIdx.x);
  WinState& st = c.st[v.w];
  const int* L = c.L + v.win.off;
  const int* area = c.acc + 4 * v.win.off;
  const uint8_t* rootflag = c.tmp + v.win.off;
  const int m1 = st.max1;
  int m2 = -1, c1 = 0;
  auto push = [&](int a) { if (a == m1) ++c1; else m2 = max(m2, a); };
  constexpr int U = 8;
  for (int k0 = 0; k0 < v.cnt; k0 += kThreads * U) {
    uint8_t rf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * kThreads + threadIdx.x;
      rf[u] = k < v.cnt ? rootflag[v.i0 + k] : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = v.i0 + k0 + u * kThreads + threadIdx.x;
      if (rf[u] && L[i] == i) push(area[i]);
    }
  }
  if (v.y0 == 0 && threadIdx.x == 0) push(st.area0);
  for (int o = 16; o > 0; o >>= 1) {
    m2 = max(m2, __shfl_down_sync(0xffffffffu, m2, o));
    c1 += __shfl_down_sync(0xffffffffu, c1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c1) atomicAdd(&st.cnt1, c1);
    if (m2 >= 0) atomicMax(&st.max2, m2);
  }
}
__global__ void __launch_bounds__(kThreads) k_mapply(Ctx c, int round) {
  const View v = view_of(c, blockIdx.x);
  const WinState& st = c.st[v.w];
  if (round < 4 && round >= st.nproc) return;
  const int* L = c.L + v.win.off;
  uint8_t* merged = c.merged + v.win.off;
  const int* acc = c.acc + 4 * v.win.off;
  const int n = v.rw * v.rh;
  const int* area = acc; const int* gain = acc + n; const int* loss = acc + 2 * n; const int* maxi = acc + 3 * n;
  // sorted_area[-2] if more than one label else sorted_area[-1] (textmask.py:114-118); label 0 always exists
  const int second = st.cnt1 >= 2 ? st.max1 : st.max2;
  const int thresh = second >= 0 ? second : st.max1;
  for (int k0 = 0; k0 < v.cnt; k0 += kThreads * kU) {
    int r[kU], a[kU], g[kU], l[kU], mx[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int k = k0 + u * kThreads + threadIdx.x;
      r[u] = k < v.cnt ? L[v.i0 + k] : -1;      // chunk-local root of the pixel ...
    }
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (r[u] >= 0) r[u] = L[r[u]];             // ... which points straight at the global root (k_flat1)
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      a[u] = 0; g[u] = 0; l[u] = 0; mx[u] = 0;
      if (r[u] >= 0) { a[u] = area[r[u]]; g[u] = gain[r[u]]; l[u] = loss[r[u]]; if (round < 4) mx[u] = maxi[r[u]]; }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (r[u] < 0) continue;
      bool ok;
      if (round < 4) {
        // `if w * h < 3: continue` (textmask.py:97): bounding boxes 1x1, 1x2, 2x1
        const bool tiny = a[u] == 1 || (a[u] == 2 && (mx[u] == r[u] + 1 || mx[u] == r[u] + v.rw));
        ok = !tiny;
      } else {
        ok = a[u] < thresh;  // textmask.py:120
      }
      if (ok && g[u] > l[u]) merged[v.i0 + k0 + u * kThreads + threadIdx.x] = 255;
    }
  }
}

// ---- dilate 3x3 (inpaint mode): merged -> tmp; the caller swaps the two planes afterwards -----------------------------
// `merged` is binary (0 / 255): the 3x3 maximum is an OR of nine shifted copies of the foreground BIT mask.  The chunk's
// rows plus one row above and below (inside the window) are packed into shared-memory words by warp ballots, one thread
// per 32-pixel word ORs the nine views (row ends masked), and the bytes are written back coalesced: ~25 instructions per
// pixel instead of ~120 (nine bounds-checked byte loads).
constexpr int kDilWords = (3 * kChunkPx) / 32 + 4;   // chunk (<= kChunkPx) + two halo rows (a row is <= kChunkPx pixels)
__global__ void __launch_bounds__(kThreads) k_dilate(Ctx c) {
  __shared__ unsigned Mw[kDilWords];
  __shared__ unsigned Ow[kChunkPx / 32 + 1];
  const View v = view_of(c, blockIdx.x);
  const uint8_t* merged = c.merged + v.win.off;
  uint8_t* tmp = c.tmp + v.win.off;
  for (int i = threadIdx.x; i < kDilWords; i += kThreads) Mw[i] = 0u;
  __syncthreads();
  const int ystart = v.y0 > 0 ? v.y0 - 1 : 0;
  const int yend = min(v.y0 + v.rows + 1, v.rh);
  const int ext = (yend - ystart) * v.rw;            // pixels of the chunk + halo rows
  const int off = (v.y0 - ystart) * v.rw;            // chunk pixel k sits at extended position k + off
  const uint8_t* src = merged + size_t(ystart) * v.rw;
  for (int e0 = 0; e0 < ext; e0 += kThreads) {
    const int e = e0 + threadIdx.x;
    const unsigned m = __ballot_sync(0xffffffffu, e < ext && src[e] != 0);
    if ((threadIdx.x & 31) == 0) Mw[e >> 5] = m;
  }
  __syncthreads();
  const DivW dv = make_div(v.rw, kChunkPx + 1);
  for (int w = threadIdx.x; w * 32 < v.cnt; w += kThreads) {
    const int k0 = w * 32;
    int yl, x0;
    divmod(k0, dv, yl, x0);
    unsigned rs = 0u;                                 // bits whose pixel is the first of its row
    for (int j = x0 == 0 ? 0 : v.rw - x0; j < 32; j += v.rw) rs |= 1u << j;
    int xe = x0 + 32;
    if (xe >= v.rw) xe %= v.rw;
    const unsigned re = (rs >> 1) | (xe == 0 ? 0x80000000u : 0u);   // ... the last of its row
    const int p = k0 + off;
    unsigned o = bits_from(Mw, p) | bits_from(Mw, p - v.rw) | bits_from(Mw, p + v.rw);
    o |= (bits_from(Mw, p - 1) | bits_from(Mw, p - v.rw - 1) | bits_from(Mw, p + v.rw - 1)) & ~rs;
    o |= (bits_from(Mw, p + 1) | bits_from(Mw, p - v.rw + 1) | bits_from(Mw, p + v.rw + 1)) & ~re;
    Ow[w] = o;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < v.cnt; k += kThreads) tmp[v.i0 + k] = ((Ow[k >> 5] >> (k & 31)) & 1u) ? 255 : 0;
}
// mask_refined[window] |= merged (textmask.py:168); windows may overlap -> atomic OR
__global__ void __launch_bounds__(kThreads) k_or(Ctx c) {
  const View v = view_of(c, blockIdx.x);
  const uint8_t* merged = c.merged + v.win.off;
  uint32_t* out_words = c.out_all + size_t(v.win.page) * c.H * c.W / 4;
  const DivW dv = make_div(v.rw, v.rw * v.rh);
  constexpr int U = 8;
  for (int k0 = 0; k0 < v.cnt; k0 += kThreads * U) {
    uint8_t mg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = k0 + u * kThreads + threadIdx.x;
      mg[u] = k < v.cnt ? merged[v.i0 + k] : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!mg[u]) continue;
      int y, x;
      divmod(v.i0 + k0 + u * kThreads + threadIdx.x, dv, y, x);
      const size_t gp = size_t(v.win.y1 + y) * c.W + v.win.x1 + x;
      atomicOr(&out_words[gp >> 2], 0xffu << (8 * (gp & 3)));
    }
  }
}

}  // namespace

size_t refine_mk_state_bytes(int n_wins) { return (size_t(n_wins) * sizeof(WinState) + 255) / 256 * 256; }
size_t refine_mk_chunk_bytes() { return sizeof(Chunk); }
int refine_mk_chunk_px() { return kChunkPx; }

// d_wins: n_wins RefineWin records; d_chunks: n_chunks {win, y0, rows, pad} (whole rows, <= refine_mk_chunk_px() pixels
// each unless a single row is longer); d_state: refine_mk_state_bytes(n_wins) bytes (zeroed here); scratch planes as in
// refine_launch.  img / mask / out hold H*W-pixel planes per page.
cudaError_t refine_mk_launch(const uint8_t* d_img, const uint8_t* d_mask, int H, int W, const void* d_wins, int n_wins,
                             const void* d_chunks, int n_chunks, int n_multi_chunks, void* d_state, size_t total_px,
                             void* scratch, int refine_mode, uint8_t* d_out, cudaStream_t s) {
  if (n_wins <= 0 || n_chunks <= 0) return cudaSuccess;
  Ctx c;
  c.img_all = d_img; c.mask_all = d_mask; c.out_all = reinterpret_cast<uint32_t*>(d_out);
  c.wins = static_cast<const RefineWin*>(d_wins);
  c.chunks = static_cast<const Chunk*>(d_chunks);
  c.st = static_cast<WinState*>(d_state);
  c.H = H; c.W = W; c.mode = refine_mode;
  char* p = static_cast<char*>(scratch);
  c.L = reinterpret_cast<int*>(p); p += total_px * 4;
  c.acc = reinterpret_cast<int*>(p); p += total_px * 16;
  c.grey = reinterpret_cast<uint8_t*>(p); p += total_px;
  c.cand = reinterpret_cast<uint8_t*>(p); p += total_px;
  c.predm = reinterpret_cast<uint8_t*>(p); p += total_px;
  c.merged = reinterpret_cast<uint8_t*>(p); p += total_px;
  c.tmp = reinterpret_cast<uint8_t*>(p);
  cudaError_t e = cudaMemsetAsync(d_state, 0, refine_mk_state_bytes(n_wins), s);
  if (e != cudaSuccess) return e;
  const unsigned g = unsigned(n_chunks);
  k_phase0<<<g, kThreads, 0, s>>>(c);
  k_decide1<<<unsigned(n_wins), kDecideThreads, 0, s>>>(c);
  k_xor<<<g, kThreads, 0, s>>>(c);
  k_decide2<<<unsigned((n_wins + 127) / 128), 128, 0, s>>>(c, n_wins);
  for (int round = 0; round < 5; ++round) {
    if (round == 4 && refine_mode == 0) {
      k_dilate<<<g, kThreads, 0, s>>>(c);
      uint8_t* t = c.merged; c.merged = c.tmp; c.tmp = t;   // the dilated plane IS `merged` from here on (no copy back)
    }
    k_label_local<<<g, kLabelThreads, 0, s>>>(c, round);
    if (n_multi_chunks > 0) {   // single-chunk windows: the chunk-local roots ARE the global roots
      k_union_border<<<unsigned(n_multi_chunks), kBorderThreads, 0, s>>>(c, round);
      k_flat1<<<unsigned(n_multi_chunks), kThreads, 0, s>>>(c, round);
    }
    if (round == 4) {
      k_top_a<<<g, kThreads, 0, s>>>(c);
      k_top_b<<<g, kThreads, 0, s>>>(c);
    }
    k_mapply<<<g, kThreads, 0, s>>>(c, round);
  }
  k_or<<<g, kThreads, 0, s>>>(c);
  return cudaGetLastError();
}

}  // namespace ctd
