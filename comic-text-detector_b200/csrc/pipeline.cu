// Full-page pipeline of the engine: everything `TextDetector.__call__` does after the network (reference
// inference.py:148-178) behind the C ABI, with the page and its masks staying in HBM:
//
//   phase A (device)  forward + NMS + mask u8 + DB threshold + CCL + line boxes           (engine.cu)
//   phase B (host)    postprocess_yolo casts, box_thresh filter, `group_output` (group.cpp), expand_textwindow
//   phase C (device)  `refine_mask` on the resident page + mask (refine.cu), optionally refine_undetected_mask
//
// ctd_submit_full / ctd_collect run batches of net-sized pages through A -> B -> C with two batches in flight per
// engine: the caller's thread enqueues phase A, a per-engine worker thread waits for A's small results, runs phase B
// for the pages of the batch on a few host threads and enqueues phase C on a second stream, so the host stage and
// the refine kernels of batch i overlap the forward of batch i+1.  ctd_detect_page is the blocking single-page form
// for pages of any size (letterbox + back-projection on the GPU), the call behind the drop-in TextDetector.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "engine.h"

using namespace ctd;

#define CK(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return ctd_fail(h, CTD_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// ---- refine_mask plumbing ------------------------------------------------------------------------------------
namespace {
// python slice normalisation of one window bound pair on an axis of length n (negative indices wrap, then clamp)
inline void norm_slice(int& lo, int& hi, int n) {
  if (lo < 0) lo = std::max(lo + n, 0);
  if (hi < 0) hi = std::max(hi + n, 0);
  lo = std::min(lo, n);
  hi = std::min(hi, n);
  if (hi < lo) hi = lo;
}
}  // namespace

void RefineJob::add(int x1, int y1, int x2, int y2, int page, int iw, int ih) {
  norm_slice(x1, x2, iw);
  norm_slice(y1, y2, ih);
  const size_t a = (x2 > x1 && y2 > y1) ? size_t(x2 - x1) * (y2 - y1) : 0;
  if (a == 0) return;                                    // empty slice: the reference's loop body is a no-op
  HostWin w{x1, y1, x2, y2, (long long)total_px, page, 0};
  const int wi = int(wins.size());
  (a > size_t(refine_large_px()) ? idx_large : idx_small).push_back(wi);
  const int rw = x2 - x1, rh = y2 - y1;
  int rows_per = std::max(1, refine_mk_chunk_px() / rw);
  if (rows_per >= 8) rows_per &= ~3;   // chunk starts on multiples of 4 rows -> 4-byte aligned in the window planes
  for (int y0 = 0; y0 < rh; y0 += rows_per)   // pad bit 0: aligned start (the labelling kernel then loads 4 pixels per thread)
    chunks.push_back(HostChunk{wi, y0, std::min(rows_per, rh - y0), ((long long)y0 * rw) % 4 == 0 ? 1 : 0});
  wins.push_back(w);
  total_px = (total_px + a + 3) / 4 * 4;
}
size_t RefineJob::table_bytes() const {
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  return al(wins.size() * sizeof(HostWin)) + al((idx_small.size() + idx_large.size()) * 4) + al(chunks.size() * sizeof(HostChunk));
}

// uploads the window tables of `job` into the (grown on demand) refine scratch and launches the refine kernels:
// d_img / d_mask / d_out are device planes of `ih*iw` pixels per page.  Stream-ordered on `st`; `pinned` (optional)
// is a host staging buffer of >= job.table_bytes() bytes that stays valid until the copy has executed.
int launch_refine(ctd_handle* h, const RefineJob& job, const uint8_t* d_img, const uint8_t* d_mask, int ih, int iw,
                  int refine_mode, uint8_t* d_out, cudaStream_t st, char* pinned) {
  if (job.wins.empty()) return CTD_OK;
  static_assert(sizeof(HostWin) == 32 && sizeof(HostChunk) == 16, "RefineWin / Chunk layout");
  if (refine_win_bytes() != sizeof(HostWin) || refine_mk_chunk_bytes() != sizeof(HostChunk))
    return ctd_fail(h, CTD_E_INVALID, "RefineWin / Chunk layout mismatch");
  const char* rf_env = getenv("CTD_REFINE");                   // CTD_REFINE=coop: the cooperative kernels of refine.cu
  bool coop = rf_env && rf_env[0] == 'c';
  for (const HostWin& w : job.wins)                            // a window row must fit one chunk of the phase kernels
    if (w.x2 - w.x1 > refine_mk_chunk_px()) coop = true;
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t tb = job.table_bytes();   // a multiple of 256
  const size_t sb = refine_mk_state_bytes(int(job.wins.size()));
  const size_t need = tb + sb + refine_scratch_bytes(job.total_px);
  if (need > h->refine_scratch_cap) {
    CK(cudaDeviceSynchronize());         // the old scratch may still be in use by an earlier launch
    cudaFree(h->d_refine_scratch);
    h->d_refine_scratch = nullptr;
    h->refine_scratch_cap = 0;
    CK(cudaMalloc(&h->d_refine_scratch, need + need / 2));
    h->refine_scratch_cap = need + need / 2;
  }
  char* base = static_cast<char*>(h->d_refine_scratch);
  const size_t wb = al(job.wins.size() * sizeof(HostWin));
  const size_t ib = al((job.idx_small.size() + job.idx_large.size()) * 4);
  std::vector<char> local;
  char* stage = pinned;
  if (!stage) { local.resize(tb); stage = local.data(); }
  memcpy(stage, job.wins.data(), job.wins.size() * sizeof(HostWin));
  int* hidx = reinterpret_cast<int*>(stage + wb);
  if (!job.idx_small.empty()) memcpy(hidx, job.idx_small.data(), job.idx_small.size() * 4);
  if (!job.idx_large.empty()) memcpy(hidx + job.idx_small.size(), job.idx_large.data(), job.idx_large.size() * 4);
  // chunk table: the chunks of windows that span SEVERAL chunks first -- only those have chunk borders to unite and
  // chunk-local roots to re-point (k_union_border / k_flat1 run on that prefix); the kernels are order-agnostic
  int n_multi = 0;
  {
    HostChunk* hc = reinterpret_cast<HostChunk*>(stage + wb + ib);
    const size_t nc = job.chunks.size();
    size_t tail = nc;
    for (size_t i = 0; i < nc;) {
      size_t j = i;
      while (j < nc && job.chunks[j].win == job.chunks[i].win) ++j;
      if (j - i > 1) {
        for (size_t k = i; k < j; ++k) hc[n_multi++] = job.chunks[k];
      } else {
        hc[--tail] = job.chunks[i];
      }
      i = j;
    }
  }
  CK(cudaMemcpyAsync(base, stage, tb, cudaMemcpyHostToDevice, st));
  if (!pinned) CK(cudaStreamSynchronize(st));   // pageable staging dies with this frame
  const int* d_idx = reinterpret_cast<const int*>(base + wb);
  if (coop)
    CK(refine_launch(d_img, d_mask, ih, iw, base, d_idx, int(job.idx_small.size()), d_idx + job.idx_small.size(),
                     int(job.idx_large.size()), job.total_px, base + tb + sb, refine_mode, d_out, st));
  else
    CK(refine_mk_launch(d_img, d_mask, ih, iw, base, int(job.wins.size()), base + wb + ib, int(job.chunks.size()), n_multi,
                        base + tb, job.total_px, base + tb + sb, refine_mode, d_out, st));
  return CTD_OK;
}

extern "C" int ctd_refine_mask(ctd_handle* h, const uint8_t* img, const uint8_t* mask, int32_t ih, int32_t iw,
                               const int32_t* windows, int32_t n_win, int32_t refine_mode, uint8_t* out) {
  if (!h || !img || !mask || !out || (n_win > 0 && !windows)) return CTD_E_INVALID;
  if (ih < 1 || iw < 1) return ctd_fail(h, CTD_E_SHAPE, "bad image size");
  CK(cudaSetDevice(h->cfg.device));
  RefineJob job;
  for (int i = 0; i < n_win; ++i) job.add(windows[4 * i], windows[4 * i + 1], windows[4 * i + 2], windows[4 * i + 3], 0, iw, ih);
  const size_t px = size_t(ih) * iw, pxa = (px + 255) / 256 * 256;
  if (int rc = ensure_io_scratch(h, pxa * 5 + 1024)) return rc;
  uint8_t* d_img = h->d_io_scratch;
  uint8_t* d_mask = d_img + pxa * 3;
  uint8_t* d_out = d_mask + pxa;
  CK(cudaMemcpyAsync(d_img, img, px * 3, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_mask, mask, px, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(d_out, 0, pxa, h->stream));
  if (int rc = launch_refine(h, job, d_img, d_mask, ih, iw, refine_mode, d_out, h->stream, nullptr)) return rc;
  CK(cudaMemcpyAsync(out, d_out, px, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

// ---- phase B for one page ------------------------------------------------------------------------------------------
namespace {
// Results of phase A for one page (host pointers into the result arena), scale of the page (1,1 for net-sized pages)
struct PageIn {
  const float* det; int n_det;                 // [n_det][6] x1,y1,x2,y2,conf,cls in net coordinates
  const int16_t* line_boxes; const float* line_scores; int n_lines;
  const uint8_t* mask; int im_w, im_h;          // page-sized mask
  float ratio_x, ratio_y;                      // resize_ratio (inference.py:148)
};

// inference.py:101-114 (postprocess_yolo casts), 158-172 (box_thresh, line rescale), textblock.group_output,
// expand_textwindow(.., 16): fills the page's block section and appends its refine windows
int host_group_page(const PageIn& in, char* section, const ArenaLayout& L, std::vector<int32_t>& win_out) {
  std::vector<int32_t> bxy(size_t(in.n_det) * 4), bcls(size_t(in.n_det));
  for (int i = 0; i < in.n_det; ++i) {
    const float* d = in.det + 6 * i;
    // det[..., [0, 2]] * ratio in float32 (numpy: float32 array * python float), then astype(int32)
    bxy[4 * i + 0] = int32_t(d[0] * in.ratio_x);
    bxy[4 * i + 1] = int32_t(d[1] * in.ratio_y);
    bxy[4 * i + 2] = int32_t(d[2] * in.ratio_x);
    bxy[4 * i + 3] = int32_t(d[3] * in.ratio_y);
    bcls[i] = int32_t(d[5]);
  }
  std::vector<int32_t> lines;
  lines.reserve(size_t(in.n_lines) * 8);
  for (int i = 0; i < in.n_lines; ++i) {
    if (!(in.line_scores[i] > 0.6f)) continue;           // box_thresh (inference.py:159-161), float32 compare
    const int16_t* b = in.line_boxes + 8 * i;
    for (int k = 0; k < 4; ++k) {                        // astype(float64) * ratio -> astype(int32)
      lines.push_back(int32_t(double(b[2 * k]) * double(in.ratio_x)));
      lines.push_back(int32_t(double(b[2 * k + 1]) * double(in.ratio_y)));
    }
  }
  ctd_page_blocks* hdr = reinterpret_cast<ctd_page_blocks*>(section);
  ctd_block* rec = reinterpret_cast<ctd_block*>(section + L.rec_off);
  int32_t* lout = reinterpret_cast<int32_t*>(section + L.lines_off);
  double* dout = reinterpret_cast<double*>(section + L.dist_off);
  int32_t nb = 0;
  int rc = ctd_group_output(bxy.data(), bcls.data(), in.n_det, lines.data(), int32_t(lines.size() / 8), in.im_w, in.im_h, in.mask,
                            1, rec, CTD_MAX_BLOCKS, lout, CTD_MAX_BLOCKS, dout, CTD_MAX_BLOCK_DIST, &nb);
  hdr->flags = 0;
  if (rc == CTD_E_CAPACITY) {
    // more distance values than the section holds (split blocks copy their parent's array): drop the distances,
    // keep blocks and lines (flag bit 0)
    std::vector<double> big(size_t(CTD_MAX_BLOCKS) * CTD_MAX_BLOCKS / 4);
    rc = ctd_group_output(bxy.data(), bcls.data(), in.n_det, lines.data(), int32_t(lines.size() / 8), in.im_w, in.im_h, in.mask, 1,
                          rec, CTD_MAX_BLOCKS, lout, CTD_MAX_BLOCKS, big.data(), int32_t(big.size()), &nb);
    if (rc == CTD_OK)
      for (int i = 0; i < nb; ++i) { rec[i].n_dist = 0; rec[i].dist_off = 0; }
    hdr->flags |= 1;
  }
  if (rc != CTD_OK) { hdr->n_blocks = 0; hdr->n_lines = 0; hdr->n_dist = 0; hdr->flags |= 2; return rc; }
  hdr->n_blocks = nb;
  int32_t tl = 0, td = 0;
  for (int i = 0; i < nb; ++i) { tl += rec[i].n_lines; td += rec[i].n_dist; }
  hdr->n_lines = tl;
  hdr->n_dist = td;
  win_out.resize(size_t(nb) * 4);
  for (int i = 0; i < nb; ++i) {
    // expand_textwindow without the slice normalisation (RefineJob::add applies it)
    const int32_t* xy = rec[i].xyxy;
    const int64_t w = int64_t(xy[2]) - xy[0], hh = int64_t(xy[3]) - xy[1];
    const int64_t pad = int64_t(nearbyint((double(std::max(hh, w)) * 0.25 + double(std::min(hh, w)) * 0.75) / 16.0));
    win_out[4 * i + 0] = int32_t(std::max<int64_t>(0, xy[0] - pad));
    win_out[4 * i + 1] = int32_t(std::max<int64_t>(0, xy[1] - pad));
    win_out[4 * i + 2] = int32_t(std::min<int64_t>(in.im_w - 1, xy[2] + pad));
    win_out[4 * i + 3] = int32_t(std::min<int64_t>(in.im_h - 1, xy[3] + pad));
  }
  return CTD_OK;
}
}  // namespace

extern "C" int ctd_results_layout(ctd_handle* h, ctd_results_layout_t* out) {
  if (!h || !out) return CTD_E_INVALID;
  const ArenaLayout& L = h->layout;
  out->max_batch = h->cfg.max_batch; out->max_h = h->cfg.max_h; out->max_w = h->cfg.max_w; out->reserved = 0;
  out->total_bytes = L.total;
  out->mask_u8 = 0; out->det = L.det; out->det_count = L.cnt; out->n_labels = L.nl;
  out->line_boxes = L.lb; out->line_scores = L.ls; out->line_count = L.lc;
  out->phase_a_bytes = L.a_bytes;
  out->mask_refined = L.refined; out->blocks = L.blocks; out->blocks_stride = L.blocks_stride;
  out->blk_records_off = L.rec_off; out->blk_lines_off = L.lines_off; out->blk_dist_off = L.dist_off;
  return CTD_OK;
}

// ---- batch pipeline -------------------------------------------------------------------------------------------------
static void pipe_worker(ctd_handle* h) {
  cudaSetDevice(h->cfg.device);
  for (;;) {
    PipeJob job;
    {
      std::unique_lock<std::mutex> lk(h->pipe_mu);
      h->pipe_cv.wait(lk, [&] { return h->pipe_quit || !h->pipe_queue.empty(); });
      if (h->pipe_queue.empty()) return;    // quit requested and nothing left
      job = h->pipe_queue.front();
      h->pipe_queue.pop_front();
    }
    int rc = CTD_OK;
    std::string err;
    auto run = [&]() -> int {
      const ArenaLayout& L = h->layout;
      const int slot = job.slot, n = job.n, ph = job.ph, pw = job.pw;
      CK(cudaEventSynchronize(h->ev_out_done[slot]));          // phase A results are in results_host
      char* res = static_cast<char*>(job.results_host);
      const size_t px = size_t(ph) * pw;
      const int32_t* det_cnt = reinterpret_cast<const int32_t*>(res + L.cnt);
      const int32_t* line_cnt = reinterpret_cast<const int32_t*>(res + L.lc);
      std::vector<std::vector<int32_t>> wins;
      wins.resize(size_t(n));
      std::vector<int> prc(size_t(n), CTD_OK);
      auto one = [&](int i) {
        PageIn in;
        in.det = reinterpret_cast<const float*>(res + L.det) + size_t(i) * 300 * 6;
        in.n_det = std::min(std::max(det_cnt[i], 0), 300);
        in.line_boxes = reinterpret_cast<const int16_t*>(res + L.lb) + size_t(i) * 1000 * 8;
        in.line_scores = reinterpret_cast<const float*>(res + L.ls) + size_t(i) * 1000;
        in.n_lines = std::min(std::max(line_cnt[i], 0), 1000);
        in.mask = reinterpret_cast<const uint8_t*>(res) + size_t(i) * px;
        in.im_w = pw; in.im_h = ph; in.ratio_x = 1.f; in.ratio_y = 1.f;
        prc[size_t(i)] = host_group_page(in, res + L.blocks + size_t(i) * L.blocks_stride, L, wins[size_t(i)]);
      };
      const int nthreads = std::max(1, std::min(n, h->host_threads));
      if (nthreads == 1) {
        for (int i = 0; i < n; ++i) one(i);
      } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
          th.emplace_back([&, t] { for (int i = t; i < n; i += nthreads) one(i); });
        for (auto& x : th) x.join();
      }
      for (int i = 0; i < n; ++i)
        if (prc[size_t(i)] != CTD_OK) return ctd_fail(h, prc[size_t(i)], "group_output failed on page %d of the batch", i);
      RefineJob rj;
      for (int i = 0; i < n; ++i)
        for (size_t k = 0; k + 3 < wins[size_t(i)].size(); k += 4)
          rj.add(wins[size_t(i)][k], wins[size_t(i)][k + 1], wins[size_t(i)][k + 2], wins[size_t(i)][k + 3], i, pw, ph);
      // phase C on the post stream: block sections to the device arena copy (one gather then moves everything),
      // refine on the resident pages + mask, mask_refined back to the host
      cudaStream_t st = h->post;
      uint8_t* d_arena = h->d_stage_out[slot];
      CK(cudaMemcpyAsync(d_arena + L.blocks, res + L.blocks, size_t(n) * L.blocks_stride, cudaMemcpyHostToDevice, st));
      CK(cudaMemsetAsync(d_arena + L.refined, 0, size_t(n) * px, st));
      char* stage_tab = rj.table_bytes() <= h->pipe_pinned_cap ? h->pipe_pinned[slot] : nullptr;   // else: pageable + sync
      const uint8_t* d_pages = job.pages_dev ? job.pages_dev : h->d_stage_in[slot];
      if (int r2 = launch_refine(h, rj, d_pages, d_arena, ph, pw, job.refine_mode, d_arena + L.refined, st, stage_tab))
        return r2;
      CK(cudaMemcpyAsync(res + L.refined, d_arena + L.refined, size_t(n) * px, cudaMemcpyDeviceToHost, st));
      CK(cudaEventRecord(h->ev_post_done[slot], st));
      return CTD_OK;
    };
    rc = run();
    if (rc != CTD_OK) err = h->err;
    {
      std::lock_guard<std::mutex> lk(h->pipe_mu);
      h->pipe_state[job.slot] = rc == CTD_OK ? 2 : 3;
      h->pipe_rc[job.slot] = rc;
      h->pipe_err[job.slot] = err;
    }
    h->pipe_done_cv.notify_all();
  }
}

static int ensure_full_pipeline(ctd_handle* h) {
  if (int rc = ensure_pipeline(h)) return rc;
  if (h->pipe_thread.joinable()) return CTD_OK;
  int lo = 0, hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CK(cudaStreamCreateWithPriority(&h->post, cudaStreamNonBlocking, hi));
  // window tables: <= CTD_MAX_BLOCKS windows per page, 32-byte record + 4-byte index
  // <= CTD_MAX_BLOCKS windows per page (32 + 4 bytes each) and <= area / chunk + rows chunks per window (16 bytes)
  h->pipe_pinned_cap = size_t(h->cfg.max_batch) * (size_t(CTD_MAX_BLOCKS) * 36 + size_t(CTD_MAX_BLOCKS) * 16 * 8) + (size_t(8) << 20);
  for (int i = 0; i < 2; ++i) {
    CK(cudaHostAlloc(reinterpret_cast<void**>(&h->pipe_pinned[i]), h->pipe_pinned_cap, cudaHostAllocDefault));
    CK(cudaEventCreateWithFlags(&h->ev_post_done[i], cudaEventDisableTiming));
  }
  const char* ht = getenv("CTD_HOST_THREADS");
  const unsigned hc = std::thread::hardware_concurrency();
  h->host_threads = ht ? atoi(ht) : int(std::max(1u, std::min(8u, hc ? hc / 2 : 4u)));
  h->pipe_quit = false;
  h->pipe_thread = std::thread(pipe_worker, h);
  return CTD_OK;
}

void ctd_pipeline_shutdown(ctd_handle* h) {
  if (h->pipe_thread.joinable()) {
    {
      std::lock_guard<std::mutex> lk(h->pipe_mu);
      h->pipe_quit = true;
    }
    h->pipe_cv.notify_all();
    h->pipe_thread.join();
  }
  for (int i = 0; i < 2; ++i) {
    if (h->pipe_pinned[i]) cudaFreeHost(h->pipe_pinned[i]);
    if (h->ev_post_done[i]) cudaEventDestroy(h->ev_post_done[i]);
    h->pipe_pinned[i] = nullptr;
    h->ev_post_done[i] = nullptr;
  }
  if (h->post) cudaStreamDestroy(h->post);
  h->post = nullptr;
}

extern "C" int ctd_submit_full(ctd_handle* h, int32_t slot, const uint8_t* pages, int32_t n, int32_t ph, int32_t pw,
                               int32_t pages_on_device, int32_t refine_mode, void* results_host) {
  if (!h || !pages || !results_host || slot < 0 || slot > 1) return CTD_E_INVALID;
  if (h->cfg.debug_skip_postproc) return ctd_fail(h, CTD_E_INVALID, "ctd_submit_full needs the full pipeline");
  if (h->slot_busy[slot]) return ctd_fail(h, CTD_E_INVALID, "slot %d has an uncollected submission", slot);
  ShapePlan* sp = nullptr;
  if (int rc = prepare_forward(h, n, ph, pw, &sp)) return rc;
  if (int rc = ensure_full_pipeline(h)) return rc;
  const ArenaLayout& L = h->layout;
  const size_t bytes = size_t(n) * ph * pw * 3;
  // the previous use of this slot's staging (refine reads stage_in / stage_out) ended with its collect
  if (!pages_on_device) {
    CK(cudaStreamWaitEvent(h->copy_in, h->ev_in_free[slot], 0));
    CK(cudaMemcpyAsync(h->d_stage_in[slot], pages, bytes, cudaMemcpyHostToDevice, h->copy_in));
    CK(cudaEventRecord(h->ev_in_done[slot], h->copy_in));
    CK(cudaEventRecord(h->ev0, h->stream));
    CK(cudaStreamWaitEvent(h->stream, h->ev_in_done[slot], 0));
    CK(cudaMemcpyAsync(h->d_pages, h->d_stage_in[slot], bytes, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaEventRecord(h->ev_in_free[slot], h->stream));
  } else {
    CK(cudaEventRecord(h->ev0, h->stream));
    CK(cudaMemcpyAsync(h->d_pages, pages, bytes, cudaMemcpyDeviceToDevice, h->stream));
  }
  if (int rc = enqueue_forward(h, n, ph, pw, *sp)) return rc;
  CK(cudaMemcpyAsync(h->d_stage_out[slot], h->d_mask_u8, L.a_bytes, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaEventRecord(h->ev_out_ready[slot], h->stream));
  CK(cudaStreamWaitEvent(h->copy_out, h->ev_out_ready[slot], 0));
  CK(cudaMemcpyAsync(results_host, h->d_stage_out[slot], L.a_bytes, cudaMemcpyDeviceToHost, h->copy_out));
  CK(cudaEventRecord(h->ev_out_done[slot], h->copy_out));
  CK(cudaStreamWaitEvent(h->post, h->ev_out_ready[slot], 0));   // phase C never starts before its phase A copy
  {
    std::lock_guard<std::mutex> lk(h->pipe_mu);
    PipeJob job;
    job.slot = slot; job.n = n; job.ph = ph; job.pw = pw; job.refine_mode = refine_mode;
    job.results_host = results_host;
    job.pages_dev = pages_on_device ? pages : nullptr;
    h->pipe_state[slot] = 1;
    h->pipe_queue.push_back(job);
  }
  h->pipe_cv.notify_one();
  h->slot_busy[slot] = true;
  h->slot_full[slot] = true;
  return CTD_OK;
}

// called by ctd_collect for slots submitted with ctd_submit_full
int ctd_collect_full(ctd_handle* h, int slot) {
  int rc;
  {
    std::unique_lock<std::mutex> lk(h->pipe_mu);
    h->pipe_done_cv.wait(lk, [&] { return h->pipe_state[slot] >= 2; });
    rc = h->pipe_rc[slot];
    if (rc != CTD_OK) h->err = h->pipe_err[slot];
    h->pipe_state[slot] = 0;
  }
  h->slot_full[slot] = false;
  if (rc != CTD_OK) return rc;
  CK(cudaEventSynchronize(h->ev_post_done[slot]));
  return CTD_OK;
}

extern "C" int ctd_device_arena(ctd_handle* h, int32_t slot, void** base, void** post_stream) {
  if (!h || slot < 0 || slot > 1 || !base) return CTD_E_INVALID;
  if (!h->d_stage_out[slot]) return ctd_fail(h, CTD_E_INVALID, "no pipelined submission has been made on this handle");
  *base = h->d_stage_out[slot];
  if (post_stream) *post_stream = h->post;
  return CTD_OK;
}

// ---- single page, any size --------------------------------------------------------------------------------------------
namespace {
__global__ void undetected_prep_kernel(uint8_t* __restrict__ mask, const uint8_t* __restrict__ refined,
                                       uint8_t* __restrict__ thr, size_t n) {
  // mask_pred[mask_refined > 30] = 0; cv2.threshold(mask_pred, 30, 255, THRESH_BINARY)  (textmask.py:136-137)
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint8_t m = mask[i];
  if (refined[i] > 30) { m = 0; mask[i] = 0; }
  thr[i] = m > 30 ? 255 : 0;
}
__global__ void or_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t n) {
  const size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] |= src[i];
}
}  // namespace

extern "C" int ctd_detect_page(ctd_handle* h, const uint8_t* page, int32_t ih, int32_t iw, int32_t net_h, int32_t net_w,
                               int32_t refine_mode, int32_t keep_undetected, uint8_t* mask_out, uint8_t* mask_refined_out,
                               ctd_block* blocks, int32_t blocks_cap, int32_t* lines_out, int32_t lines_cap, double* dist_out,
                               int32_t dist_cap, int32_t* n_blocks) {
  if (!h || !page || !mask_out || !mask_refined_out || !n_blocks) return CTD_E_INVALID;
  if (ih < 1 || iw < 1) return ctd_fail(h, CTD_E_SHAPE, "bad page size %dx%d", ih, iw);
  *n_blocks = 0;
  // letterbox geometry (imgproc_utils.py:86-117 with auto=False; python round = half to even)
  const double r = std::min(double(net_h) / ih, double(net_w) / iw);
  const int unpad_w = int(nearbyint(iw * r)), unpad_h = int(nearbyint(ih * r));
  const int dw = net_w - unpad_w, dh = net_h - unpad_h;
  if (unpad_w < 1 || unpad_h < 1 || dw < 0 || dh < 0) return ctd_fail(h, CTD_E_SHAPE, "page does not letterbox into the net input");
  ShapePlan* sp = nullptr;
  if (int rc = prepare_forward(h, 1, net_h, net_w, &sp)) return rc;
  const size_t px = size_t(ih) * iw, pxa = (px + 255) / 256 * 256;
  // io scratch: page (3 px) | page-sized mask | mask_refined | second refine output | thresholded mask
  if (int rc = ensure_io_scratch(h, pxa * 7 + 1024)) return rc;
  uint8_t* d_page = h->d_io_scratch;
  uint8_t* d_mask = d_page + pxa * 3;
  uint8_t* d_ref = d_mask + pxa;
  uint8_t* d_ref2 = d_ref + pxa;
  uint8_t* d_thr = d_ref2 + pxa;
  cudaStream_t st = h->stream;
  CK(cudaEventRecord(h->ev0, st));
  CK(cudaMemcpyAsync(d_page, page, px * 3, cudaMemcpyHostToDevice, st));
  const bool same = ih == net_h && iw == net_w;
  if (same) CK(cudaMemcpyAsync(h->d_pages, d_page, px * 3, cudaMemcpyDeviceToDevice, st));
  else CK(resize_linear_u8_launch(d_page, ih, iw, size_t(iw) * 3, 3, h->d_pages, unpad_h, unpad_w, net_h, net_w, st));
  if (int rc = enqueue_forward(h, 1, net_h, net_w, *sp)) return rc;
  // mask back-projection (inference.py:164-168)
  if (same) CK(cudaMemcpyAsync(d_mask, h->d_mask_u8, px, cudaMemcpyDeviceToDevice, st));
  else CK(resize_linear_u8_launch(h->d_mask_u8, net_h - dh, net_w - dw, size_t(net_w), 1, d_mask, ih, iw, ih, iw, st));
  std::vector<float> det(300 * 6);
  std::vector<int16_t> lb(1000 * 8);
  std::vector<float> ls(1000);
  int32_t n_det = 0, n_lines = 0;
  CK(cudaMemcpyAsync(mask_out, d_mask, px, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(det.data(), h->d_det, det.size() * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&n_det, h->d_det_count, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(lb.data(), h->d_line_boxes, lb.size() * 2, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(ls.data(), h->d_line_scores, ls.size() * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&n_lines, h->d_line_count, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemsetAsync(d_ref, 0, pxa, st));
  CK(cudaStreamSynchronize(st));
  // phase B
  const ArenaLayout& L = h->layout;
  std::vector<char> section(L.blocks_stride);
  PageIn in;
  in.det = det.data(); in.n_det = std::min(std::max(n_det, 0), 300);
  in.line_boxes = lb.data(); in.line_scores = ls.data(); in.n_lines = std::min(std::max(n_lines, 0), 1000);
  in.mask = mask_out; in.im_w = iw; in.im_h = ih;
  in.ratio_x = float(double(iw) / double(net_w - dw));     // resize_ratio (inference.py:148)
  in.ratio_y = float(double(ih) / double(net_h - dh));
  std::vector<int32_t> wins;
  if (int rc = host_group_page(in, section.data(), L, wins)) return ctd_fail(h, rc, "group_output failed");
  const ctd_page_blocks* hdr = reinterpret_cast<const ctd_page_blocks*>(section.data());
  const ctd_block* rec = reinterpret_cast<const ctd_block*>(section.data() + L.rec_off);
  const int nb = hdr->n_blocks;
  *n_blocks = nb;
  if (nb > blocks_cap || hdr->n_lines > lines_cap || hdr->n_dist > dist_cap || (nb > 0 && (!blocks || !lines_out)) ||
      (hdr->n_dist > 0 && !dist_out))
    return ctd_fail(h, CTD_E_CAPACITY, "%d blocks / %d lines / %d distances do not fit the output arrays", nb, hdr->n_lines, hdr->n_dist);
  if (nb > 0) {
    memcpy(blocks, rec, size_t(nb) * sizeof(ctd_block));
    memcpy(lines_out, section.data() + L.lines_off, size_t(hdr->n_lines) * 32);
    if (hdr->n_dist > 0) memcpy(dist_out, section.data() + L.dist_off, size_t(hdr->n_dist) * 8);
  }
  // phase C
  RefineJob rj;
  for (int i = 0; i < nb; ++i) rj.add(wins[4 * i], wins[4 * i + 1], wins[4 * i + 2], wins[4 * i + 3], 0, iw, ih);
  if (int rc = launch_refine(h, rj, d_page, d_mask, ih, iw, refine_mode, d_ref, st, nullptr)) return rc;
  if (keep_undetected) {
    // refine_undetected_mask (textmask.py:135-156); the page mask is modified in place and returned, as in the reference
    undetected_prep_kernel<<<unsigned((px + 255) / 256), 256, 0, st>>>(d_mask, d_ref, d_thr, px);
    CK(cudaGetLastError());
    int32_t* d_stats = nullptr;
    int32_t n_lab = 0;
    const int stats_cap = ((ih + 1) / 2) * ((iw + 1) / 2) + 2;   // worst case of 8-connected components + background
    if (int rc = cc_device(h, d_thr, ih, iw, stats_cap, &d_stats, &n_lab)) return rc;
    std::vector<int32_t> stats(size_t(std::max(n_lab, 0)) * 5);
    if (n_lab > 0) CK(cudaMemcpyAsync(stats.data(), d_stats, stats.size() * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    RefineJob rj2;
    bool first_valid = true;
    for (int li = 0; li < n_lab; ++li) {
      const int32_t* s5 = &stats[size_t(li) * 5];
      if (!(s5[4] > 50)) continue;
      if (first_valid) { first_valid = false; continue; }        // valid_labels[1:]
      const int64_t bb[4] = {s5[0], s5[1], int64_t(s5[0]) + s5[2], int64_t(s5[1]) + s5[3]};
      int64_t score = -1;
      for (int b = 0; b < nb; ++b) {
        const int32_t* q = rec[b].xyxy;
        const int64_t x1 = std::max<int64_t>(q[0], bb[0]), y1 = std::max<int64_t>(q[1], bb[1]);
        const int64_t x2 = std::min<int64_t>(q[2], bb[2]), y2 = std::min<int64_t>(q[3], bb[3]);
        const int64_t a = (y2 < y1 || x2 < x1) ? -1 : (y2 - y1) * (x2 - x1);
        if (a > score) score = a;
      }
      if (double(score) / double(s5[2]) / double(s5[3]) < 0.5) {
        int32_t xy[4] = {int32_t(bb[0]), int32_t(bb[1]), int32_t(bb[2]), int32_t(bb[3])}, w4[4];
        const int64_t w = bb[2] - bb[0], hh = bb[3] - bb[1];
        const int64_t pad = int64_t(nearbyint((double(std::max(hh, w)) * 0.25 + double(std::min(hh, w)) * 0.75) / 16.0));
        w4[0] = int32_t(std::max<int64_t>(0, xy[0] - pad)); w4[1] = int32_t(std::max<int64_t>(0, xy[1] - pad));
        w4[2] = int32_t(std::min<int64_t>(iw - 1, xy[2] + pad)); w4[3] = int32_t(std::min<int64_t>(ih - 1, xy[3] + pad));
        rj2.add(w4[0], w4[1], w4[2], w4[3], 0, iw, ih);
      }
    }
    if (!rj2.wins.empty()) {
      CK(cudaMemsetAsync(d_ref2, 0, pxa, st));
      if (int rc = launch_refine(h, rj2, d_page, d_mask, ih, iw, refine_mode, d_ref2, st, nullptr)) return rc;
      or_kernel<<<unsigned((px + 255) / 256), 256, 0, st>>>(d_ref, d_ref2, px);
      CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(mask_out, d_mask, px, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaMemcpyAsync(mask_refined_out, d_ref, px, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CTD_OK;
}
