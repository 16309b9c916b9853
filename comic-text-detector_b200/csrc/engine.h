// Private header of libctd_b200.so: the handle and the helpers shared by engine.cu (network executor, C ABI of the
// forward pass) and pipeline.cu (group_output / refine_mask pipeline around it).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "kernels.h"

using ctd::BneckPlan;
using ctd::SegTailPlan;
using ctd::ConvTcPlan;
using ctd::NmsWorkspace;
using ctd::PFN_encodeTiled;

// byte offsets inside the result arena (one allocation per engine, sized for max_batch pages of max_h x max_w):
// phase-A section (mask_u8 | det | det_count | n_labels | line_boxes | line_scores | line_count) = [0, a_bytes),
// then mask_refined u8 planes, then one fixed-stride block section per page (ctd_page_blocks header, ctd_block
// records, line quads, distances)
struct ArenaLayout {
  size_t det, cnt, nl, lb, ls, lc, a_bytes, refined, blocks, blocks_stride, rec_off, lines_off, dist_off, total;
};

struct PipeJob {
  int slot = 0, n = 0, ph = 0, pw = 0, refine_mode = 0;
  void* results_host = nullptr;
  const uint8_t* pages_dev = nullptr;   // caller's device pages (pages_on_device) or null: the slot's staging copy
};

// refine windows of one launch (all pages of a batch): RefineWin records + small / large index lists
struct HostWin { int x1, y1, x2, y2; long long off; int page, pad; };   // = RefineWin (refine.cu / refine_mk.cu)
struct HostChunk { int win, y0, rows, pad; };                            // = Chunk (refine_mk.cu)
struct RefineJob {
  std::vector<HostWin> wins;
  std::vector<int> idx_small, idx_large;     // cooperative kernels (refine.cu)
  std::vector<HostChunk> chunks;             // phase-synchronous kernels (refine_mk.cu)
  size_t total_px = 0;
  void add(int x1, int y1, int x2, int y2, int page, int iw, int ih);   // python slice semantics; empty windows dropped
  size_t table_bytes() const;
};

struct ShapePlan {
  std::vector<ConvTcPlan> tc;  // index = op index (unused entries default)
  std::vector<char> has_tc;
  std::vector<BneckPlan> bn;   // fused Bottleneck ops (CTD_OP_BNECK), index = op index
  SegTailPlan seg;             // seg tail as GEMM + col2im (conv_fuse.cu) when seg_op >= 0
  int seg_op = -1;
  cudaGraphExec_t graph = nullptr;
  int launches = 0;
};
struct ctd_handle {
  ctd_config cfg{};
  std::vector<ctd_op> ops;
  std::vector<ctd_bufdesc> bufs;
  std::string err;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, tev0 = nullptr, tev1 = nullptr;
  std::vector<cudaEvent_t> op_events;
  PFN_encodeTiled enc = nullptr;
  char* d_blob = nullptr;
  size_t blob_bytes = 0;
  std::vector<void*> d_buf;
  // split-fp16 mode (CTD_PREC_SPLIT_TC): d_buf holds the FP32 master copy of every activation; d_buf16[i] holds its
  // fp16 hi | lo planes ([2*n][h][w][C], refreshed after every op that writes the buffer) = the MMA operands;
  // d_wsplit holds per GEMM op the fp16 weight rows hi then lo (wsplit_off[op], bytes).
  std::vector<void*> d_buf16;
  char* d_wsplit = nullptr;
  std::vector<size_t> wsplit_off;
  int elem = 2;  // bytes per activation element
  uint8_t* d_pages = nullptr;
  float* d_blks = nullptr;
  float* d_mask = nullptr;
  uint8_t* d_mask_u8 = nullptr;   // start of the contiguous result arena: mask_u8 | det | det_count | n_labels
  size_t results_bytes = 0;
  float* d_lines = nullptr;
  uint8_t* d_bitmap = nullptr;
  float* d_det = nullptr;
  int* d_det_count = nullptr;
  int32_t* d_labels = nullptr;
  int32_t* d_nlabels = nullptr;
  int32_t* d_ccl_scratch = nullptr;
  void* d_segrep_scratch = nullptr;
  void* d_refine_scratch = nullptr;
  size_t refine_scratch_cap = 0;
  void* d_cc_scratch = nullptr;      // ctd_connected_components: grow-on-demand, any image size
  size_t cc_scratch_cap = 0;
  uint8_t* d_io_scratch = nullptr;   // page upload / resized mask staging of the resize entry points
  size_t io_scratch_cap = 0;
  int16_t* d_line_boxes = nullptr;
  float* d_line_scores = nullptr;
  int32_t* d_line_count = nullptr;
  void* d_nms_ws = nullptr;
  NmsWorkspace nms{};
  std::map<std::tuple<int, int, int>, ShapePlan> plans;
  // pipelined host path (ctd_submit / ctd_collect): two staging slots, copy streams either side of compute
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  uint8_t* d_stage_in[2] = {nullptr, nullptr};
  uint8_t* d_stage_out[2] = {nullptr, nullptr};
  cudaEvent_t ev_in_done[2] = {nullptr, nullptr}, ev_in_free[2] = {nullptr, nullptr};
  cudaEvent_t ev_out_ready[2] = {nullptr, nullptr}, ev_out_done[2] = {nullptr, nullptr};
  bool slot_busy[2] = {false, false};
  // overlapped schedule: post-processing of the DB maps / the Detect rows runs on side streams under the
  // remaining network ops (see run_ops)
  int halo_mode = 15;  // CTD_HALO bit mask (0 routes every conv through conv_tc_kernel, for A/B measurements)
  bool overlap = false;
  cudaStream_t side = nullptr, side2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork2 = nullptr, ev_join2 = nullptr;
  cudaEvent_t ev_xjoin = nullptr;   // ctd_join (never part of a captured graph)
  std::vector<char> db_ancestor;   // op feeds the DB tail (computed once in ctd_create)
  // result arena layout (ctd_results_layout) and the full pipeline (pipeline.cu): worker thread + post stream
  ArenaLayout layout{};
  cudaStream_t post = nullptr;
  cudaEvent_t ev_post_done[2] = {nullptr, nullptr};
  char* pipe_pinned[2] = {nullptr, nullptr};     // pinned staging of the refine window tables, one per slot
  size_t pipe_pinned_cap = 0;
  bool slot_full[2] = {false, false};            // slot was submitted with ctd_submit_full
  std::thread pipe_thread;
  std::mutex pipe_mu;
  std::condition_variable pipe_cv, pipe_done_cv;
  std::deque<PipeJob> pipe_queue;
  bool pipe_quit = false;
  int pipe_state[2] = {0, 0};                    // 0 idle, 1 queued / running, 2 phase C enqueued, 3 failed
  int pipe_rc[2] = {0, 0};
  std::string pipe_err[2];
  int host_threads = 4;
  // last forward
  int n = 0, ph = 0, pw = 0;
  int last_launches = 0;
  bool have_forward = false;
};


int ctd_fail(ctd_handle* h, int code, const char* fmt, ...);
int prepare_forward(ctd_handle* h, int32_t n, int32_t ph, int32_t pw, ShapePlan** out);
int enqueue_forward(ctd_handle* h, int32_t n, int32_t ph, int32_t pw, ShapePlan& sp);
int ensure_pipeline(ctd_handle* h);
int ensure_io_scratch(ctd_handle* h, size_t bytes);
// connected components + stats of a DEVICE u8 image on the engine stream (grow-on-demand scratch): *d_stats points at
// [stats_cap][5] ints on the device, *n_labels is read back (synchronises the stream)
int cc_device(ctd_handle* h, const uint8_t* d_img, int ih, int iw, int stats_cap, int32_t** d_stats, int32_t* n_labels);
int launch_refine(ctd_handle* h, const RefineJob& job, const uint8_t* d_img, const uint8_t* d_mask, int ih, int iw,
                  int refine_mode, uint8_t* d_out, cudaStream_t st, char* pinned);
int ctd_collect_full(ctd_handle* h, int slot);
void ctd_pipeline_shutdown(ctd_handle* h);
