// libctd_b200.so: C-ABI engine (see include/ctd_b200.h).  Owns device buffers, the weight blob,
// per-shape launch plans (tensor maps) and the optional CUDA graph; runs the op list emitted by
// the Python host "compiler".  No CPU fallback: every entry point fails without an sm_100 GPU.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "engine.h"

using namespace ctd;

namespace {
thread_local std::string g_create_error;
}  // namespace

int ctd_fail(ctd_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else g_create_error = buf;
  return code;
}

#define CK(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return ctd_fail(h, CTD_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static int rows_per_image(int ph, int pw) { return 3 * ((ph / 8) * (pw / 8) + (ph / 16) * (pw / 16) + (ph / 32) * (pw / 32)); }

extern "C" const char* ctd_last_error(const ctd_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" void ctd_destroy(ctd_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  ctd_pipeline_shutdown(h);
  for (auto& kv : h->plans)
    if (kv.second.graph) cudaGraphExecDestroy(kv.second.graph);
  for (void* p : h->d_buf) cudaFree(p);
  for (void* p : h->d_buf16) cudaFree(p);
  cudaFree(h->d_wsplit);
  cudaFree(h->d_blob); cudaFree(h->d_pages); cudaFree(h->d_blks); cudaFree(h->d_mask); cudaFree(h->d_mask_u8);
  cudaFree(h->d_lines); cudaFree(h->d_bitmap); cudaFree(h->d_labels);
  cudaFree(h->d_ccl_scratch); cudaFree(h->d_nms_ws); cudaFree(h->d_segrep_scratch); cudaFree(h->d_refine_scratch); cudaFree(h->d_cc_scratch); cudaFree(h->d_io_scratch);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->tev0) cudaEventDestroy(h->tev0);
  if (h->tev1) cudaEventDestroy(h->tev1);
  for (auto e : h->op_events) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->d_stage_in[i]); cudaFree(h->d_stage_out[i]);
    if (h->ev_in_done[i]) cudaEventDestroy(h->ev_in_done[i]);
    if (h->ev_in_free[i]) cudaEventDestroy(h->ev_in_free[i]);
    if (h->ev_out_ready[i]) cudaEventDestroy(h->ev_out_ready[i]);
    if (h->ev_out_done[i]) cudaEventDestroy(h->ev_out_done[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_fork2) cudaEventDestroy(h->ev_fork2);
  if (h->ev_xjoin) cudaEventDestroy(h->ev_xjoin);
  if (h->ev_join2) cudaEventDestroy(h->ev_join2);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->side2) cudaStreamDestroy(h->side2);
  if (h->copy_in) cudaStreamDestroy(h->copy_in);
  if (h->copy_out) cudaStreamDestroy(h->copy_out);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int ctd_create(ctd_handle** out, const ctd_config* cfg, const ctd_op* ops, int32_t n_ops,
                          const ctd_bufdesc* bufs, int32_t n_bufs, const void* blob, size_t blob_bytes) {
  ctd_handle* h = nullptr;
  if (!out || !cfg || !ops || !bufs || !blob) return ctd_fail(nullptr, CTD_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != CTD_ABI_VERSION) return ctd_fail(nullptr, CTD_E_INVALID, "ABI version mismatch");
  if (cfg->max_h % 64 || cfg->max_w % 64 || cfg->max_batch < 1) return ctd_fail(nullptr, CTD_E_SHAPE, "bad max shape");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device)
    return ctd_fail(nullptr, CTD_E_NO_DEVICE, "no CUDA device %d (this engine has no CPU fallback)", cfg->device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
    return ctd_fail(nullptr, CTD_E_NO_DEVICE, "device %d is not sm_100 (compute %d.%d)", cfg->device, prop.major, prop.minor);
  h = new ctd_handle();
  h->cfg = *cfg;
  h->ops.assign(ops, ops + n_ops);
  h->bufs.assign(bufs, bufs + n_bufs);
  {
    // ancestors of the DB tail at buffer granularity (a buffer, once needed, stays needed: several ops may
    // write disjoint channel ranges of it); one backward pass is enough because producers precede consumers
    h->db_ancestor.assign(size_t(n_ops), 0);
    std::vector<char> needed(size_t(n_bufs), 0);
    bool have_db = false;
    for (int i = n_ops - 1; i >= 0; --i) {
      const ctd_op& op = ops[i];
      bool anc = false;
      if (op.kind == CTD_OP_DB_TAIL && !have_db) { anc = true; have_db = true; }
      else if (have_db && op.kind != CTD_OP_DETECT && op.kind != CTD_OP_SEG_TAIL && op.kind != CTD_OP_DB_TAIL) {
        // SPPF_POOL appends its pooled channels to its own source buffer (no dst_buf)
        const int wbuf = op.kind == CTD_OP_SPPF_POOL ? op.src_buf[0] : op.dst_buf;
        anc = wbuf >= 0 && wbuf < n_bufs && needed[wbuf];
      }
      if (!anc) continue;
      h->db_ancestor[size_t(i)] = 1;
      for (int k = 0; k < op.n_src && k < 3; ++k)
        if (op.src_buf[k] >= 0 && op.src_buf[k] < n_bufs) needed[op.src_buf[k]] = 1;
      if (op.residual && op.dst_buf >= 0 && op.dst_buf < n_bufs) needed[op.dst_buf] = 1;
    }
    const char* hm = getenv("CTD_HALO");
    // bit 0: conv_halo_kernel (resident weights), 1: conv_hs_kernel (streamed), 2: conv_sw_kernel, 3: seg tail as GEMM + col2im
    h->halo_mode = hm ? atoi(hm) : 15;
    const char* ov = getenv("CTD_OVERLAP");
    h->overlap = have_db && !(ov && ov[0] == '0');
  }
  h->elem = (cfg->precision == CTD_PREC_FP32_SIMT || cfg->precision == CTD_PREC_SPLIT_TC) ? 4 : 2;
  auto bail = [&](int code) { std::string e = h->err; ctd_destroy(h); g_create_error = e; return code; };
#define CKC(expr)                                                                                        \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      ctd_fail(h, CTD_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);          \
      return bail(CTD_E_CUDA);                                                                           \
    }                                                                                                    \
  } while (0)
  CKC(cudaSetDevice(cfg->device));
  {
    int lo = 0, hi = 0;
    CKC(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CKC(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, lo));
    // side streams get the HIGHER priority: their small blocks slot in whenever a persistent conv CTA retires
    CKC(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
    CKC(cudaStreamCreateWithPriority(&h->side2, cudaStreamNonBlocking, hi));
  }
  CKC(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&h->ev_xjoin, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming));
  CKC(cudaEventCreate(&h->ev0));
  CKC(cudaEventCreate(&h->ev1));
  CKC(cudaEventCreate(&h->tev0));
  CKC(cudaEventCreate(&h->tev1));
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CKC(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      ctd_fail(h, CTD_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
      return bail(CTD_E_CUDA);
    }
    h->enc = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  CKC(conv_tc_init());
  CKC(ctd::conv_bneck_init());
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == CTD_OP_BNECK && (cfg->precision != CTD_PREC_FP16_TC || !ctd::conv_bneck_supported(ops[i].cout))) {
      ctd_fail(h, CTD_E_INVALID, "op %d: fused Bottleneck needs CTD_PREC_FP16_TC and 32 or 64 channels", i);
      return bail(CTD_E_INVALID);
    }
  h->blob_bytes = blob_bytes;
  CKC(cudaMalloc(&h->d_blob, blob_bytes));
  CKC(cudaMemcpy(h->d_blob, blob, blob_bytes, cudaMemcpyHostToDevice));
  const size_t nb = size_t(cfg->max_batch), mh = cfg->max_h, mw = cfg->max_w;
  h->d_buf.assign(n_bufs, nullptr);
  for (int i = 0; i < n_bufs; ++i) {
    const size_t bytes = nb * (mh / bufs[i].down) * (mw / bufs[i].down + 4) * bufs[i].channels * h->elem;
    CKC(cudaMalloc(&h->d_buf[i], bytes));
    CKC(cudaMemset(h->d_buf[i], 0, bytes));
  }
  if (cfg->precision == CTD_PREC_SPLIT_TC) {
    h->d_buf16.assign(n_bufs, nullptr);
    for (int i = 0; i < n_bufs; ++i) {
      const size_t bytes = 2 * nb * (mh / bufs[i].down) * (mw / bufs[i].down + 4) * bufs[i].channels * 2;
      CKC(cudaMalloc(&h->d_buf16[i], bytes));
      CKC(cudaMemset(h->d_buf16[i], 0, bytes));
    }
    // weights: hi rows (the blob's fp16 copy = fp16(w32)) followed by lo rows fp16(w32 - hi), per GEMM op
    h->wsplit_off.assign(size_t(n_ops), 0);
    std::vector<__half> ws;
    const char* hb = static_cast<const char*>(blob);
    for (int i = 0; i < n_ops; ++i) {
      const ctd_op& op = ops[i];
      if (op.kind != CTD_OP_CONV && op.kind != CTD_OP_DECONV4 && op.kind != CTD_OP_DETECT) continue;
      int cin = 0;
      for (int k = 0; k < op.n_src; ++k) cin += op.src_c[k];
      const int taps = op.kind == CTD_OP_DECONV4 ? 4 : op.ksize * op.ksize;
      const int nph = op.kind == CTD_OP_DECONV4 ? 4 : 1;
      const size_t cnt = size_t(nph) * op.cout_pad * taps * cin;
      if (size_t(op.w32_off) + cnt * 4 > blob_bytes || size_t(op.w16_off) + cnt * 2 > blob_bytes) {
        ctd_fail(h, CTD_E_INVALID, "op %d: weights outside the blob", i);
        return bail(CTD_E_INVALID);
      }
      while (ws.size() % 128) ws.push_back(__float2half(0.f));   // 256-byte aligned rows for the tensor map
      h->wsplit_off[size_t(i)] = ws.size() * 2;
      const float* w32 = reinterpret_cast<const float*>(hb + op.w32_off);
      const size_t base = ws.size();
      ws.resize(base + 2 * cnt);
      for (size_t k = 0; k < cnt; ++k) {
        const __half hi = __float2half_rn(w32[k]);
        ws[base + k] = hi;
        ws[base + cnt + k] = __float2half_rn(w32[k] - __half2float(hi));
      }
    }
    CKC(cudaMalloc(&h->d_wsplit, ws.size() * 2 + 256));
    CKC(cudaMemcpy(h->d_wsplit, ws.data(), ws.size() * 2, cudaMemcpyHostToDevice));
  }
  const size_t px = nb * mh * mw;
  const int no = 5 + cfg->nc;
  CKC(cudaMalloc(&h->d_pages, px * 3));
  CKC(cudaMalloc(&h->d_blks, nb * rows_per_image(mh, mw) * no * sizeof(float)));
  CKC(cudaMalloc(&h->d_mask, px * 4));
  CKC(cudaMalloc(&h->d_lines, px * 2 * 4));
  CKC(cudaMalloc(&h->d_bitmap, px));
  CKC(cudaMalloc(&h->d_labels, px * 4));
  {
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    ArenaLayout& L = h->layout;
    L.det = al(px); L.cnt = L.det + al(nb * 300 * 6 * 4); L.nl = L.cnt + al(nb * 4);
    L.lb = L.nl + al(nb * 4); L.ls = L.lb + al(nb * 1000 * 8 * 2); L.lc = L.ls + al(nb * 1000 * 4);
    L.a_bytes = L.lc + al(nb * 4);
    L.refined = L.a_bytes;
    L.blocks = L.refined + al(px);
    L.rec_off = 64;
    L.lines_off = L.rec_off + al(size_t(CTD_MAX_BLOCKS) * sizeof(ctd_block));
    L.dist_off = L.lines_off + al(size_t(CTD_MAX_BLOCKS) * 32);
    L.blocks_stride = L.dist_off + al(size_t(CTD_MAX_BLOCK_DIST) * 8);
    L.total = L.blocks + nb * L.blocks_stride;
    h->results_bytes = L.total;
    CKC(cudaMalloc(&h->d_mask_u8, h->results_bytes));
    CKC(cudaMemset(h->d_mask_u8, 0, h->results_bytes));
    h->d_det = reinterpret_cast<float*>(h->d_mask_u8 + L.det);
    h->d_det_count = reinterpret_cast<int*>(h->d_mask_u8 + L.cnt);
    h->d_nlabels = reinterpret_cast<int32_t*>(h->d_mask_u8 + L.nl);
    h->d_line_boxes = reinterpret_cast<int16_t*>(h->d_mask_u8 + L.lb);
    h->d_line_scores = reinterpret_cast<float*>(h->d_mask_u8 + L.ls);
    h->d_line_count = reinterpret_cast<int32_t*>(h->d_mask_u8 + L.lc);
  }
  CKC(cudaMalloc(&h->d_ccl_scratch, px * 4 * 3));
  CKC(cudaMalloc(&h->d_segrep_scratch, segrep_scratch_bytes(int(nb), int(mh), int(mw), 1000)));
  const int cap = 4096;
  CKC(cudaMalloc(&h->d_nms_ws, nms_workspace_bytes(int(nb), cap)));
  nms_workspace_bind(h->nms, h->d_nms_ws, int(nb), cap);
  CKC(cudaDeviceSynchronize());
#undef CKC
  *out = h;
  return CTD_OK;
}

// -----------------------------------------------------------------------------------------
static int op_geom(ctd_handle* h, const ctd_op& op, int n, int ph, int pw, ConvGeom& g) {
  memset(&g, 0, sizeof(g));
  fill_conv_geom_taps(g, op.kind == CTD_OP_DETECT ? CTD_OP_CONV : op.kind, op.ksize, op.stride);
  g.n_img = n;
  g.n_src = op.n_src;
  const ctd_bufdesc& sb = h->bufs[op.src_buf[0]];
  g.src_h = ph / sb.down;
  g.src_w = pw / sb.down;
  g.cin_total = 0;
  for (int s = 0; s < op.n_src; ++s) {
    const ctd_bufdesc& b = h->bufs[op.src_buf[s]];
    if (b.down != sb.down) return ctd_fail(h, CTD_E_INVALID, "op sources differ in resolution");
    g.src_c[s] = op.src_c[s];
    g.src_cstride[s] = b.channels;
    g.cin_total += op.src_c[s];
  }
  g.k_total = g.taps * g.cin_total;
  g.gh = op.kind == CTD_OP_DECONV4 ? g.src_h : g.src_h / op.stride;
  g.gw = op.kind == CTD_OP_DECONV4 ? g.src_w : g.src_w / op.stride;
  g.dst_h = g.gh * g.out_mul;
  g.dst_w = g.gw * g.out_mul;
  g.cout = op.cout;
  g.cout_pad = op.cout_pad;
  if (op.dst_buf >= 0) {
    g.dst_cstride = h->bufs[op.dst_buf].channels;
    g.dst_coff = op.dst_coff;
  }
  g.act = op.act;
  g.residual = op.residual;
  return CTD_OK;
}

static int build_plans(ctd_handle* h, int n, int ph, int pw, ShapePlan& sp) {
  sp.tc.resize(h->ops.size());
  sp.has_tc.assign(h->ops.size(), 0);
  if (h->cfg.precision == CTD_PREC_SPLIT_TC) {
    // every GEMM-shaped op through conv_tc_kernel in split-fp16 form; stem / tails / thin ops stay on the fp32
    // CUDA-core kernels (run_one_op)
    for (size_t i = 0; i < h->ops.size(); ++i) {
      const ctd_op& op = h->ops[i];
      if (op.kind != CTD_OP_CONV && op.kind != CTD_OP_DECONV4 && op.kind != CTD_OP_DETECT) continue;
      ConvGeom g;
      if (int rc = op_geom(h, op, n, ph, pw, g)) return rc;
      const void* src[CTD_MAX_SRC];
      int coff[CTD_MAX_SRC];
      for (int s = 0; s < op.n_src; ++s) {
        src[s] = h->d_buf16[op.src_buf[s]];
        coff[s] = op.src_coff[s];
      }
      __half* dst = op.kind == CTD_OP_DETECT ? nullptr : static_cast<__half*>(h->d_buf[op.dst_buf]);
      const char* e = conv_tc_plan(sp.tc[i], h->enc, g, src, coff, h->d_wsplit + h->wsplit_off[i],
                                   reinterpret_cast<const float*>(h->d_blob + op.b_off), dst, 1);
      if (e) return ctd_fail(h, CTD_E_INVALID, "op %zu: %s", i, e);
      if (op.kind == CTD_OP_DETECT) {
        ConvTcParams& p = sp.tc[i].p;
        p.blks = h->d_blks;
        p.blks_rows_per_img = rows_per_image(ph, pw);
        int row0 = 0;
        for (int l = 0; l < op.aux; ++l) row0 += 3 * (ph / (8 << l)) * (pw / (8 << l));
        p.level_row0 = row0;
        p.nc = h->cfg.nc;
        float hp[7];
        cudaMemcpy(hp, h->d_blob + op.p_off, sizeof(hp), cudaMemcpyDeviceToHost);
        p.det_stride = hp[0];
        for (int k = 0; k < 6; ++k) p.anchor_wh[k] = hp[1 + k];
      }
      sp.has_tc[i] = 1;
    }
    return CTD_OK;
  }
  if (h->cfg.precision != CTD_PREC_FP16_TC) return CTD_OK;
  sp.bn.resize(h->ops.size());
  for (size_t i = 0; i < h->ops.size(); ++i) {
    const ctd_op& op = h->ops[i];
    if (op.kind == CTD_OP_BNECK) {
      // fused Bottleneck: 1x1 + 3x3 (+ residual) in one kernel, intermediate in shared memory (conv_fuse.cu)
      const ctd_bufdesc& sb = h->bufs[op.src_buf[0]];
      const ctd_bufdesc& db = h->bufs[op.dst_buf];
      if (op.n_src != 1 || op.src_c[0] != op.cout || op.cout != op.cout_pad || sb.down != db.down || op.src_buf[0] == op.dst_buf)
        return ctd_fail(h, CTD_E_INVALID, "op %zu: malformed fused Bottleneck", i);
      int nsm = 148;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->cfg.device);
      const char* e = ctd::conv_bneck_plan(sp.bn[i], h->enc, n, ph / sb.down, pw / sb.down, op.cout, h->d_buf[op.src_buf[0]],
                                           sb.channels, op.src_coff[0], h->d_blob + op.w16_off,
                                           reinterpret_cast<const float*>(h->d_blob + op.b_off),
                                           static_cast<__half*>(h->d_buf[op.dst_buf]), db.channels, op.dst_coff, op.act,
                                           op.residual, nsm);
      if (e) return ctd_fail(h, CTD_E_INVALID, "op %zu: %s", i, e);
      sp.has_tc[i] = 1;
      continue;
    }
    if (op.kind == CTD_OP_STEM) {
      const char* e = ((h->halo_mode & 1) ? conv_halo_plan_stem : conv_tc_plan_stem)(
          sp.tc[i], h->enc, h->d_buf[op.src_buf[0]], n, ph, pw, h->d_blob + op.w16_off,
          reinterpret_cast<const float*>(h->d_blob + op.b_off), static_cast<__half*>(h->d_buf[op.dst_buf]),
          h->bufs[op.dst_buf].channels, op.dst_coff, op.cout, op.act);
      if (e) return ctd_fail(h, CTD_E_INVALID, "stem: %s", e);
      sp.has_tc[i] = 1;
      continue;
    }
    if (op.kind == CTD_OP_SEG_TAIL && (h->halo_mode & 8) && op.b_off > 0 && op.src_c[0] == 64) {
      // final ConvT 4x4 s2 (64 -> 1) + sigmoid + u8 mask as one 1x1 GEMM over the 16 kernel positions + col2im epilogue
      const ctd_bufdesc& sb = h->bufs[op.src_buf[0]];
      int nsm = 148;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->cfg.device);
      const char* e = ctd::conv_segtail_plan(sp.seg, h->enc, n, ph / sb.down, pw / sb.down, h->d_buf[op.src_buf[0]], sb.channels,
                                             op.src_coff[0], h->d_blob + op.b_off, h->d_mask, h->d_mask_u8, nsm);
      if (e) return ctd_fail(h, CTD_E_INVALID, "seg tail: %s", e);
      sp.seg_op = int(i);
      sp.has_tc[i] = 1;
      continue;
    }
    if (op.kind == CTD_OP_SEG_TAIL && (h->halo_mode & 1) && op.w16_off > 0 && op.cout_pad == 16) {
      // final ConvT 4x4 s2 (C -> 1) + sigmoid + u8 mask as a 3x3 / 4-output halo convolution
      ctd_op c3 = op;
      c3.kind = CTD_OP_CONV; c3.ksize = 3; c3.stride = 1;
      ConvGeom g;
      if (int rc = op_geom(h, c3, n, ph, pw, g)) return rc;
      const void* src[CTD_MAX_SRC] = {h->d_buf[op.src_buf[0]]};
      int coff[CTD_MAX_SRC] = {op.src_coff[0]};
      const char* e = conv_halo_plan(sp.tc[i], h->enc, g, src, coff, h->d_blob + op.w16_off, nullptr, nullptr, h->d_mask,
                                     h->d_mask_u8);
      if (e) return ctd_fail(h, CTD_E_INVALID, "seg tail: %s", e);
      sp.has_tc[i] = sp.tc[i].halo ? 1 : 0;
      continue;
    }
    if (op.kind != CTD_OP_CONV && op.kind != CTD_OP_DECONV4 && op.kind != CTD_OP_DETECT) continue;
    ConvGeom g;
    if (int rc = op_geom(h, op, n, ph, pw, g)) return rc;
    const void* src[CTD_MAX_SRC];
    int coff[CTD_MAX_SRC];
    for (int s = 0; s < op.n_src; ++s) {
      src[s] = h->d_buf[op.src_buf[s]];
      coff[s] = op.src_coff[s];
    }
    __half* dst = op.kind == CTD_OP_DETECT ? nullptr : static_cast<__half*>(h->d_buf[op.dst_buf]);
    const char* e = nullptr;
    sp.tc[i].halo = 0;
    if ((h->halo_mode & 1) && op.kind != CTD_OP_DETECT)
      e = conv_halo_plan(sp.tc[i], h->enc, g, src, coff, h->d_blob + op.w16_off,
                         reinterpret_cast<const float*>(h->d_blob + op.b_off), dst);
    if (!e && !sp.tc[i].halo && (h->halo_mode & 4) && op.kind != CTD_OP_DETECT)
      e = conv_sw_plan(sp.tc[i], h->enc, g, src, coff, h->d_blob + op.w16_off,
                       reinterpret_cast<const float*>(h->d_blob + op.b_off), dst);
    if (!e && !sp.tc[i].halo && (h->halo_mode & 2) && op.kind != CTD_OP_DETECT)
      e = conv_hs_plan(sp.tc[i], h->enc, g, src, coff, h->d_blob + op.w16_off,
                       reinterpret_cast<const float*>(h->d_blob + op.b_off), dst);
    if (!e && !sp.tc[i].halo)
      e = conv_tc_plan(sp.tc[i], h->enc, g, src, coff, h->d_blob + op.w16_off,
                       reinterpret_cast<const float*>(h->d_blob + op.b_off), dst);
    if (e) return ctd_fail(h, CTD_E_INVALID, "op %zu: %s", i, e);
    if (op.kind == CTD_OP_DETECT) {
      ConvTcParams& p = sp.tc[i].p;
      p.blks = h->d_blks;
      p.blks_rows_per_img = rows_per_image(ph, pw);
      int row0 = 0;
      for (int l = 0; l < op.aux; ++l) row0 += 3 * (ph / (8 << l)) * (pw / (8 << l));
      p.level_row0 = row0;
      p.nc = h->cfg.nc;
      float hp[7];
      cudaMemcpy(hp, h->d_blob + op.p_off, sizeof(hp), cudaMemcpyDeviceToHost);
      p.det_stride = hp[0];
      for (int k = 0; k < 6; ++k) p.anchor_wh[k] = hp[1 + k];
    }
    sp.has_tc[i] = 1;
  }
  return CTD_OK;
}

template <typename T>
static int run_op_simt(ctd_handle* h, const ctd_op& op, int n, int ph, int pw) {
  cudaStream_t s = h->stream;
  const bool f32 = sizeof(T) == 4;
  switch (op.kind) {
    case CTD_OP_CONV:
    case CTD_OP_DECONV4:
    case CTD_OP_DETECT: {
      ConvSimtParams p;
      memset(&p, 0, sizeof(p));
      if (int rc = op_geom(h, op, n, ph, pw, p.g)) return rc;
      for (int k = 0; k < op.n_src; ++k)
        p.src[k] = static_cast<char*>(h->d_buf[op.src_buf[k]]) + size_t(op.src_coff[k]) * sizeof(T);
      p.w = h->d_blob + (f32 ? op.w32_off : op.w16_off);
      p.bias = reinterpret_cast<const float*>(h->d_blob + op.b_off);
      if (op.kind == CTD_OP_DETECT) {
        p.dst = nullptr;
        p.blks = h->d_blks;
        p.blks_rows_per_img = rows_per_image(ph, pw);
        int row0 = 0;
        for (int l = 0; l < op.aux; ++l) row0 += 3 * (ph / (8 << l)) * (pw / (8 << l));
        p.level_row0 = row0;
        p.nc = h->cfg.nc;
        float hp[7];
        cudaMemcpy(hp, h->d_blob + op.p_off, sizeof(hp), cudaMemcpyDeviceToHost);
        p.det_stride = hp[0];
        for (int k = 0; k < 6; ++k) p.anchor_wh[k] = hp[1 + k];
      } else {
        p.dst = h->d_buf[op.dst_buf];
      }
      CK(conv_simt_launch<T>(p, s));
      return CTD_OK;
    }
    default: return ctd_fail(h, CTD_E_INVALID, "run_op_simt: bad kind %d", op.kind);
  }
}

template <typename T>
static int run_op_thin(ctd_handle* h, const ctd_op& op, int n, int ph, int pw) {
  cudaStream_t s = h->stream;
  const ctd_bufdesc* sb = (op.kind == CTD_OP_STEM || op.kind == CTD_OP_S2D) ? nullptr : &h->bufs[op.src_buf[0]];
  (void)sb;
  const int sh = sb ? ph / sb->down : ph, sw = sb ? pw / sb->down : pw;
  const T* src = sb ? static_cast<const T*>(h->d_buf[op.src_buf[0]]) + op.src_coff[0] : nullptr;
  switch (op.kind) {
    case CTD_OP_STEM:
      CK(stem_launch<T>(h->d_pages, n, ph, pw, reinterpret_cast<const float*>(h->d_blob + op.w32_off),
                        reinterpret_cast<const float*>(h->d_blob + op.b_off), static_cast<T*>(h->d_buf[op.dst_buf]),
                        h->bufs[op.dst_buf].channels, op.dst_coff, op.cout, op.act, s));
      return CTD_OK;
    case CTD_OP_S2D:
      CK(s2d_launch<T>(h->d_pages, n, ph, pw, static_cast<T*>(h->d_buf[op.dst_buf]), h->bufs[op.dst_buf].channels,
                       op.dst_coff, pw / 2, 0, s));
      return CTD_OK;
    case CTD_OP_AVGPOOL2:
      CK(avgpool2_launch<T>(src, n, sh, sw, op.src_c[0], sb->channels,
                            static_cast<T*>(h->d_buf[op.dst_buf]) + op.dst_coff, h->bufs[op.dst_buf].channels, s));
      return CTD_OK;
    case CTD_OP_SPPF_POOL:
      CK(sppf_pool_launch<T>(static_cast<T*>(h->d_buf[op.src_buf[0]]) + op.src_coff[0], n, sh, sw, op.src_c[0],
                             sb->channels, s));
      return CTD_OK;
    case CTD_OP_UPSAMPLE2:
      CK(upsample2_launch<T>(src, n, sh, sw, op.src_c[0], sb->channels,
                             static_cast<T*>(h->d_buf[op.dst_buf]) + op.dst_coff, h->bufs[op.dst_buf].channels, s));
      return CTD_OK;
    case CTD_OP_SEG_TAIL:
      CK(seg_tail_launch<T>(src, n, sh, sw, op.src_c[0], sb->channels,
                            reinterpret_cast<const float*>(h->d_blob + op.p_off), h->d_mask, h->d_mask_u8, s));
      return CTD_OK;
    case CTD_OP_DB_TAIL:
      CK(db_tail_launch<T>(src, n, sh, sw, sb->channels, reinterpret_cast<const float*>(h->d_blob + op.p_off),
                           h->d_lines, h->d_bitmap, h->cfg.db_thresh, s));
      return CTD_OK;
    default: return ctd_fail(h, CTD_E_INVALID, "run_op_thin: bad kind %d", op.kind);
  }
}

// split-fp16 mode: refresh the fp16 hi | lo planes of the channel slice op `i` has just written
static int split_written_slice(ctd_handle* h, const ctd_op& op, int n, int ph, int pw, int* cnt) {
  int buf = op.dst_buf, coff = op.dst_coff, c = op.cout;
  if (op.kind == CTD_OP_SPPF_POOL) { buf = op.src_buf[0]; coff = op.src_coff[0] + op.src_c[0]; c = 3 * op.src_c[0]; }
  else if (op.kind == CTD_OP_AVGPOOL2 || op.kind == CTD_OP_UPSAMPLE2) c = op.src_c[0];
  else if (op.kind == CTD_OP_S2D) c = 16;
  if (buf < 0 || c <= 0) return CTD_OK;
  const ctd_bufdesc& b = h->bufs[buf];
  const size_t npix = size_t(n) * (ph / b.down) * (pw / b.down);
  __half* hi = static_cast<__half*>(h->d_buf16[buf]) + coff;
  CK(split_planes_launch(static_cast<const float*>(h->d_buf[buf]) + coff, hi, hi + npix * b.channels, npix, c, b.channels,
                         h->stream));
  ++*cnt;
  return CTD_OK;
}

static int run_one_op(ctd_handle* h, size_t i, int n, int ph, int pw, ShapePlan& sp, int* cnt) {
  const ctd_op& op = h->ops[i];
  const bool gemm = op.kind == CTD_OP_CONV || op.kind == CTD_OP_DECONV4 || op.kind == CTD_OP_DETECT;
  int rc = CTD_OK;
  if (h->cfg.precision == CTD_PREC_SPLIT_TC) {
    if (gemm) {
      cudaError_t e = conv_tc_launch(sp.tc[i], h->stream);
      rc = e == cudaSuccess ? CTD_OK : ctd_fail(h, CTD_E_CUDA, "conv_tc (split) op %zu: %s", i, cudaGetErrorString(e));
    } else {
      rc = run_op_thin<float>(h, op, n, ph, pw);
    }
    ++*cnt;
    if (rc) return rc;
    return split_written_slice(h, op, n, ph, pw, cnt);
  }
  if (op.kind == CTD_OP_BNECK) {
    if (h->cfg.precision != CTD_PREC_FP16_TC || !sp.has_tc[i])
      return ctd_fail(h, CTD_E_INVALID, "op %zu: fused Bottleneck ops run on the fp16 tensor-core engine only", i);
    cudaError_t e = ctd::conv_bneck_launch(sp.bn[i], h->stream);
    rc = e == cudaSuccess ? CTD_OK : ctd_fail(h, CTD_E_CUDA, "conv_bneck op %zu: %s", i, cudaGetErrorString(e));
  } else if (op.kind == CTD_OP_STEM && h->cfg.precision == CTD_PREC_FP16_TC) {
    // tensor-core stem: space-to-depth pre-pass into the padded window buffer, then the implicit GEMM
    cudaError_t e = s2d_launch<__half>(h->d_pages, n, ph, pw, static_cast<__half*>(h->d_buf[op.src_buf[0]]), 16, 0,
                                       pw / 2 + 4, 1, h->stream);
    if (e == cudaSuccess) e = conv_tc_launch(sp.tc[i], h->stream);
    rc = e == cudaSuccess ? CTD_OK : ctd_fail(h, CTD_E_CUDA, "stem op %zu: %s", i, cudaGetErrorString(e));
    ++*cnt;
  } else if (op.kind == CTD_OP_SEG_TAIL && h->cfg.precision == CTD_PREC_FP16_TC && sp.has_tc[i]) {
    cudaError_t e = sp.seg_op == int(i) ? ctd::conv_segtail_launch(sp.seg, h->stream) : conv_tc_launch(sp.tc[i], h->stream);
    rc = e == cudaSuccess ? CTD_OK : ctd_fail(h, CTD_E_CUDA, "seg tail op %zu: %s", i, cudaGetErrorString(e));
  } else if (gemm) {
    if (h->cfg.precision == CTD_PREC_FP16_TC) {
      cudaError_t e = conv_tc_launch(sp.tc[i], h->stream);
      rc = e == cudaSuccess ? CTD_OK : ctd_fail(h, CTD_E_CUDA, "conv_tc op %zu: %s", i, cudaGetErrorString(e));
    } else if (h->cfg.precision == CTD_PREC_FP32_SIMT) {
      rc = run_op_simt<float>(h, op, n, ph, pw);
    } else {
      rc = run_op_simt<__half>(h, op, n, ph, pw);
    }
  } else {
    rc = h->elem == 4 ? run_op_thin<float>(h, op, n, ph, pw) : run_op_thin<__half>(h, op, n, ph, pw);
  }
  ++*cnt;
  return rc;
}

static int run_ops(ctd_handle* h, int n, int ph, int pw, ShapePlan& sp, int* launches, bool record = false) {
  int cnt = 0;
  size_t evi = 0;
  const int rows = rows_per_image(ph, pw);
  if (h->overlap && !record && !h->cfg.debug_skip_postproc) {
    // Two-phase order.  Phase 1: every op the DB maps depend on (program order).  Then the DB post-processing
    // (CCL + contour boxes: latency-bound kernels that leave most SMs idle) forks to a side stream and runs
    // UNDER phase 2 = the rest of the network (PAN, Detect heads, the seg-head tail); NMS forks the same way once
    // the Detect rows exist.  No buffer is shared between the branches (the compiler never reuses buffers).
    for (size_t i = 0; i < h->ops.size(); ++i)
      if (h->db_ancestor[i])
        if (int rc = run_one_op(h, i, n, ph, pw, sp, &cnt)) return rc;
    CK(cudaEventRecord(h->ev_fork, h->stream));
    CK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    CK(ccl_launch(h->d_bitmap, n, ph, pw, h->d_labels, h->d_ccl_scratch, h->d_nlabels, h->side));
    CK(segrep_launch(h->d_bitmap, h->d_lines, size_t(2) * ph * pw, h->d_ccl_scratch, n, ph, pw, 1000, 1.5f,
                     h->d_segrep_scratch, h->d_line_boxes, h->d_line_scores, h->d_line_count, h->side));
    CK(cudaEventRecord(h->ev_join, h->side));
    cnt += 23;
    size_t last_detect = h->ops.size();
    for (size_t i = 0; i < h->ops.size(); ++i)
      if (!h->db_ancestor[i] && h->ops[i].kind == CTD_OP_DETECT) last_detect = i;
    bool nms_forked = false;
    for (size_t i = 0; i < h->ops.size(); ++i) {   // program order: PAN -> Detect heads -> seg-head tail
      if (h->db_ancestor[i]) continue;
      if (int rc = run_one_op(h, i, n, ph, pw, sp, &cnt)) return rc;
      if (i == last_detect) {
        CK(cudaEventRecord(h->ev_fork2, h->stream));
        CK(cudaStreamWaitEvent(h->side2, h->ev_fork2, 0));
        CK(nms_launch(h->d_blks, n, rows, h->cfg.nc, h->cfg.conf_thresh, h->cfg.nms_thresh, h->nms, h->d_det,
                      h->d_det_count, h->side2));
        CK(cudaEventRecord(h->ev_join2, h->side2));
        nms_forked = true;
        cnt += 5;
      }
    }
    if (!nms_forked) {
      CK(nms_launch(h->d_blks, n, rows, h->cfg.nc, h->cfg.conf_thresh, h->cfg.nms_thresh, h->nms, h->d_det,
                    h->d_det_count, h->stream));
      cnt += 5;
    } else {
      CK(cudaStreamWaitEvent(h->stream, h->ev_join2, 0));
    }
    CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    *launches = cnt;
    return CTD_OK;
  }
  if (record) CK(cudaEventRecord(h->op_events[evi++], h->stream));
  for (size_t i = 0; i < h->ops.size(); ++i) {
    if (int rc = run_one_op(h, i, n, ph, pw, sp, &cnt)) return rc;
    if (record) CK(cudaEventRecord(h->op_events[evi++], h->stream));
  }
  *launches = cnt;
  if (h->cfg.debug_skip_postproc) return CTD_OK;
  // post-processing on the same stream
  CK(nms_launch(h->d_blks, n, rows, h->cfg.nc, h->cfg.conf_thresh, h->cfg.nms_thresh, h->nms, h->d_det,
                h->d_det_count, h->stream));
  cnt += 5;
  if (record) CK(cudaEventRecord(h->op_events[evi++], h->stream));
  CK(ccl_launch(h->d_bitmap, n, ph, pw, h->d_labels, h->d_ccl_scratch, h->d_nlabels, h->stream));
  cnt += 8;
  CK(segrep_launch(h->d_bitmap, h->d_lines, size_t(2) * ph * pw, h->d_ccl_scratch, n, ph, pw, 1000, 1.5f,
                   h->d_segrep_scratch, h->d_line_boxes, h->d_line_scores, h->d_line_count, h->stream));
  cnt += 15;
  if (record) CK(cudaEventRecord(h->op_events[evi++], h->stream));
  *launches = cnt;
  return CTD_OK;
}

// shape checks + plan lookup + (first time) graph capture; the forward itself is enqueue_forward()
int prepare_forward(ctd_handle* h, int32_t n, int32_t ph, int32_t pw, ShapePlan** out) {
  if (n < 1 || n > h->cfg.max_batch) return ctd_fail(h, CTD_E_CAPACITY, "batch %d exceeds max_batch %d", n, h->cfg.max_batch);
  if (ph % 64 || pw % 64 || ph > h->cfg.max_h || pw > h->cfg.max_w || ph < 64 || pw < 64)
    return ctd_fail(h, CTD_E_SHAPE, "page %dx%d must be a multiple of 64 and <= %dx%d", ph, pw, h->cfg.max_h, h->cfg.max_w);
  CK(cudaSetDevice(h->cfg.device));
  auto key = std::make_tuple(int(n), int(ph), int(pw));
  auto it = h->plans.find(key);
  if (it == h->plans.end()) {
    ShapePlan sp;
    if (int rc = build_plans(h, n, ph, pw, sp)) return rc;
    it = h->plans.emplace(key, std::move(sp)).first;
  }
  ShapePlan& sp = it->second;
  if (h->cfg.use_graph && !sp.graph && (h->cfg.precision == CTD_PREC_FP16_TC || h->cfg.precision == CTD_PREC_SPLIT_TC)) {
    // DETECT params are fetched with a blocking memcpy in the SIMT path: plans are already built, so
    // capture only sees kernel launches (TC path).  SIMT paths run un-captured.
    cudaGraph_t graph;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = run_ops(h, n, ph, pw, sp, &sp.launches);
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc) return rc;
    CK(e);
    CK(cudaGraphInstantiate(&sp.graph, graph, 0));
    cudaGraphDestroy(graph);
  }
  *out = &sp;
  return CTD_OK;
}

int enqueue_forward(ctd_handle* h, int32_t n, int32_t ph, int32_t pw, ShapePlan& sp) {
  if (sp.graph) {
    CK(cudaGraphLaunch(sp.graph, h->stream));
  } else {
    if (int rc = run_ops(h, n, ph, pw, sp, &sp.launches)) return rc;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  h->last_launches = sp.launches;
  h->n = n; h->ph = ph; h->pw = pw;
  h->have_forward = true;
  return CTD_OK;
}

extern "C" int ctd_forward(ctd_handle* h, const uint8_t* pages, int32_t n, int32_t ph, int32_t pw,
                           int32_t pages_on_device) {
  if (!h || !pages) return CTD_E_INVALID;
  ShapePlan* sp = nullptr;
  if (int rc = prepare_forward(h, n, ph, pw, &sp)) return rc;
  const size_t bytes = size_t(n) * ph * pw * 3;
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(cudaMemcpyAsync(h->d_pages, pages, bytes, pages_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                     h->stream));
  return enqueue_forward(h, n, ph, pw, *sp);
}

// ---- pipelined host path ---------------------------------------------------------------------------
// submit(slot): copy_in: [wait slot's staging free] H2D pages -> stage_in[slot]
//               compute: [wait H2D] stage_in -> d_pages (D2D), forward, arena -> stage_out[slot] (D2D)
//               copy_out: [wait arena copy] D2H stage_out[slot] -> results_host
// so the H2D of batch i+1 and the D2H of batch i-1 run under the forward of batch i.
int ensure_pipeline(ctd_handle* h) {
  if (h->copy_in) return CTD_OK;
  CK(cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking));
  const size_t in_bytes = size_t(h->cfg.max_batch) * h->cfg.max_h * h->cfg.max_w * 3;
  for (int i = 0; i < 2; ++i) {
    CK(cudaMalloc(&h->d_stage_in[i], in_bytes));
    CK(cudaMalloc(&h->d_stage_out[i], h->results_bytes));
    CK(cudaEventCreateWithFlags(&h->ev_in_done[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_in_free[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_out_ready[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_out_done[i], cudaEventDisableTiming));
  }
  return CTD_OK;
}

extern "C" int ctd_submit(ctd_handle* h, int32_t slot, const uint8_t* pages_host, int32_t n, int32_t ph, int32_t pw,
                          void* results_host) {
  if (!h || !pages_host || !results_host || slot < 0 || slot > 1) return CTD_E_INVALID;
  if (h->cfg.debug_skip_postproc) return ctd_fail(h, CTD_E_INVALID, "ctd_submit needs the full pipeline");
  if (h->slot_busy[slot]) return ctd_fail(h, CTD_E_INVALID, "slot %d has an uncollected submission", slot);
  ShapePlan* sp = nullptr;
  if (int rc = prepare_forward(h, n, ph, pw, &sp)) return rc;
  if (int rc = ensure_pipeline(h)) return rc;
  const size_t bytes = size_t(n) * ph * pw * 3;
  CK(cudaStreamWaitEvent(h->copy_in, h->ev_in_free[slot], 0));   // no-op before the slot's first use
  CK(cudaMemcpyAsync(h->d_stage_in[slot], pages_host, bytes, cudaMemcpyHostToDevice, h->copy_in));
  CK(cudaEventRecord(h->ev_in_done[slot], h->copy_in));
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(cudaStreamWaitEvent(h->stream, h->ev_in_done[slot], 0));
  CK(cudaMemcpyAsync(h->d_pages, h->d_stage_in[slot], bytes, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaEventRecord(h->ev_in_free[slot], h->stream));
  if (int rc = enqueue_forward(h, n, ph, pw, *sp)) return rc;
  CK(cudaStreamWaitEvent(h->stream, h->ev_out_done[slot], 0));  // previous D2H of this slot has drained
  CK(cudaMemcpyAsync(h->d_stage_out[slot], h->d_mask_u8, h->layout.a_bytes, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaEventRecord(h->ev_out_ready[slot], h->stream));
  CK(cudaStreamWaitEvent(h->copy_out, h->ev_out_ready[slot], 0));
  CK(cudaMemcpyAsync(results_host, h->d_stage_out[slot], h->layout.a_bytes, cudaMemcpyDeviceToHost, h->copy_out));
  CK(cudaEventRecord(h->ev_out_done[slot], h->copy_out));
  h->slot_busy[slot] = true;
  return CTD_OK;
}

extern "C" int ctd_collect(ctd_handle* h, int32_t slot) {
  if (!h || slot < 0 || slot > 1) return CTD_E_INVALID;
  if (!h->slot_busy[slot]) return ctd_fail(h, CTD_E_INVALID, "slot %d has nothing in flight", slot);
  CK(cudaSetDevice(h->cfg.device));
  if (h->slot_full[slot]) {
    const int rc = ctd_collect_full(h, slot);
    h->slot_busy[slot] = false;
    return rc;
  }
  CK(cudaEventSynchronize(h->ev_out_done[slot]));
  h->slot_busy[slot] = false;
  return CTD_OK;
}

int ensure_io_scratch(ctd_handle* h, size_t bytes) {
  if (bytes <= h->io_scratch_cap) return CTD_OK;
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(h->d_io_scratch);
  h->d_io_scratch = nullptr;
  h->io_scratch_cap = 0;
  CK(cudaMalloc(&h->d_io_scratch, bytes + bytes / 4));
  h->io_scratch_cap = bytes + bytes / 4;
  return CTD_OK;
}

extern "C" int ctd_forward_resized(ctd_handle* h, const uint8_t* page, int32_t ih, int32_t iw, int32_t unpad_h,
                                   int32_t unpad_w, int32_t net_h, int32_t net_w) {
  if (!h || !page) return CTD_E_INVALID;
  if (ih < 1 || iw < 1 || unpad_h < 1 || unpad_w < 1 || unpad_h > net_h || unpad_w > net_w)
    return ctd_fail(h, CTD_E_SHAPE, "letterbox %dx%d -> %dx%d does not fit the %dx%d net input", ih, iw, unpad_h, unpad_w, net_h, net_w);
  ShapePlan* sp = nullptr;
  if (int rc = prepare_forward(h, 1, net_h, net_w, &sp)) return rc;
  const size_t bytes = size_t(ih) * iw * 3;
  if (int rc = ensure_io_scratch(h, bytes)) return rc;
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(cudaMemcpyAsync(h->d_io_scratch, page, bytes, cudaMemcpyHostToDevice, h->stream));
  CK(resize_linear_u8_launch(h->d_io_scratch, ih, iw, size_t(iw) * 3, 3, h->d_pages, unpad_h, unpad_w, net_h, net_w, h->stream));
  return enqueue_forward(h, 1, net_h, net_w, *sp);
}

extern "C" int ctd_get_mask_u8_resized(ctd_handle* h, int32_t crop_h, int32_t crop_w, int32_t out_h, int32_t out_w,
                                       uint8_t* mask_out) {
  if (!h || !mask_out) return CTD_E_INVALID;
  if (!h->have_forward) return ctd_fail(h, CTD_E_INVALID, "no forward pass has been run on this handle");
  if (crop_h < 1 || crop_w < 1 || crop_h > h->ph || crop_w > h->pw || out_h < 1 || out_w < 1)
    return ctd_fail(h, CTD_E_SHAPE, "bad crop %dx%d of the %dx%d mask", crop_h, crop_w, h->ph, h->pw);
  CK(cudaSetDevice(h->cfg.device));
  const size_t bytes = size_t(out_h) * out_w;
  if (int rc = ensure_io_scratch(h, bytes)) return rc;
  CK(resize_linear_u8_launch(h->d_mask_u8, crop_h, crop_w, size_t(h->pw), 1, h->d_io_scratch, out_h, out_w, out_h, out_w, h->stream));
  CK(cudaMemcpyAsync(mask_out, h->d_io_scratch, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

extern "C" int ctd_resize_linear_u8(ctd_handle* h, const uint8_t* src, int32_t sh, int32_t sw, int32_t channels, uint8_t* dst,
                                    int32_t dh, int32_t dw) {
  if (!h || !src || !dst) return CTD_E_INVALID;
  if ((channels != 1 && channels != 3) || sh < 1 || sw < 1 || dh < 1 || dw < 1) return ctd_fail(h, CTD_E_SHAPE, "bad resize shape");
  CK(cudaSetDevice(h->cfg.device));
  const size_t sb = size_t(sh) * sw * channels, db = size_t(dh) * dw * channels;
  const size_t so = (sb + 255) / 256 * 256;
  if (int rc = ensure_io_scratch(h, so + db)) return rc;
  CK(cudaMemcpyAsync(h->d_io_scratch, src, sb, cudaMemcpyHostToDevice, h->stream));
  CK(resize_linear_u8_launch(h->d_io_scratch, sh, sw, size_t(sw) * channels, channels, h->d_io_scratch + so, dh, dw, dh, dw, h->stream));
  CK(cudaMemcpyAsync(dst, h->d_io_scratch + so, db, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

extern "C" int ctd_join(ctd_handle* h, ctd_handle* other) {
  if (!h || !other) return CTD_E_INVALID;
  if (h == other) return CTD_OK;
  if (h->cfg.device != other->cfg.device) return ctd_fail(h, CTD_E_INVALID, "ctd_join: handles live on different devices");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventRecord(other->ev_xjoin, other->stream));
  CK(cudaStreamWaitEvent(h->stream, other->ev_xjoin, 0));
  return CTD_OK;
}

extern "C" int ctd_results_bytes(ctd_handle* h, size_t* bytes) {
  if (!h || !bytes) return CTD_E_INVALID;
  *bytes = h->results_bytes;
  return CTD_OK;
}

#define NEED_FWD()                                                                     \
  if (!h) return CTD_E_INVALID;                                                        \
  if (!h->have_forward) return ctd_fail(h, CTD_E_INVALID, "no forward pass has been run"); \
  CK(cudaSetDevice(h->cfg.device));

extern "C" int ctd_get_net_outputs(ctd_handle* h, float* blks, float* mask, float* lines) {
  NEED_FWD();
  const size_t px = size_t(h->n) * h->ph * h->pw;
  if (blks)
    CK(cudaMemcpyAsync(blks, h->d_blks, size_t(h->n) * rows_per_image(h->ph, h->pw) * (5 + h->cfg.nc) * 4,
                       cudaMemcpyDeviceToHost, h->stream));
  if (mask) CK(cudaMemcpyAsync(mask, h->d_mask, px * 4, cudaMemcpyDeviceToHost, h->stream));
  if (lines) CK(cudaMemcpyAsync(lines, h->d_lines, px * 8, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

extern "C" int ctd_get_mask_u8(ctd_handle* h, uint8_t* mask_u8) {
  NEED_FWD();
  if (!mask_u8) return CTD_E_INVALID;
  CK(cudaMemcpyAsync(mask_u8, h->d_mask_u8, size_t(h->n) * h->ph * h->pw, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

extern "C" int ctd_get_detections(ctd_handle* h, float* det, int32_t* det_count) {
  NEED_FWD();
  if (!det || !det_count) return CTD_E_INVALID;
  CK(cudaMemcpyAsync(det, h->d_det, size_t(h->n) * 300 * 6 * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(det_count, h->d_det_count, size_t(h->n) * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

extern "C" int ctd_get_nms_status(ctd_handle* h, int32_t* cand_total, int32_t* cap) {
  if (!h) return CTD_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  if (cap) *cap = h->nms.cap;
  if (cand_total) {
    const int n = h->have_forward ? h->n : 1;
    CK(cudaMemcpyAsync(cand_total, h->nms.cand_total, size_t(n) * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return CTD_OK;
}

extern "C" int ctd_get_db_components(ctd_handle* h, uint8_t* bitmap, int32_t* labels, int32_t* n_labels) {
  NEED_FWD();
  const size_t px = size_t(h->n) * h->ph * h->pw;
  if (bitmap) CK(cudaMemcpyAsync(bitmap, h->d_bitmap, px, cudaMemcpyDeviceToHost, h->stream));
  if (labels) CK(cudaMemcpyAsync(labels, h->d_labels, px * 4, cudaMemcpyDeviceToHost, h->stream));
  if (n_labels) CK(cudaMemcpyAsync(n_labels, h->d_nlabels, size_t(h->n) * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

extern "C" int ctd_get_text_lines(ctd_handle* h, int16_t* boxes, float* scores, int32_t* counts) {
  NEED_FWD();
  if (!boxes || !scores || !counts) return CTD_E_INVALID;
  CK(cudaMemcpyAsync(boxes, h->d_line_boxes, size_t(h->n) * 1000 * 8 * 2, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(scores, h->d_line_scores, size_t(h->n) * 1000 * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(counts, h->d_line_count, size_t(h->n) * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

extern "C" int ctd_seg_represent(ctd_handle* h, const float* pred, int32_t ih, int32_t iw, float thresh, int16_t* boxes,
                                 float* scores, int32_t* count) {
  if (!h || !pred || !boxes || !scores || !count) return CTD_E_INVALID;
  if (ih < 1 || iw < 1 || size_t(ih) * iw > size_t(h->cfg.max_h) * h->cfg.max_w || ih > 2048 || iw > 2048)
    return ctd_fail(h, CTD_E_CAPACITY, "map larger than the workspace");
  CK(cudaSetDevice(h->cfg.device));
  const size_t px = size_t(ih) * iw;
  CK(cudaMemcpyAsync(h->d_lines, pred, px * 4, cudaMemcpyHostToDevice, h->stream));
  CK(binarize_launch(h->d_lines, px, thresh, h->d_bitmap, h->stream));
  CK(ccl_launch(h->d_bitmap, 1, ih, iw, h->d_labels, h->d_ccl_scratch, h->d_nlabels, h->stream));
  CK(segrep_launch(h->d_bitmap, h->d_lines, px, h->d_ccl_scratch, 1, ih, iw, 1000, 1.5f, h->d_segrep_scratch,
                   h->d_line_boxes, h->d_line_scores, h->d_line_count, h->stream));
  CK(cudaMemcpyAsync(boxes, h->d_line_boxes, 1000 * 8 * 2, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(scores, h->d_line_scores, 1000 * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(count, h->d_line_count, 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_forward = false;
  return CTD_OK;
}

extern "C" int ctd_last_forward_ms(ctd_handle* h, float* ms) {
  NEED_FWD();
  if (!ms) return CTD_E_INVALID;
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return CTD_OK;
}

extern "C" int ctd_last_launch_count(ctd_handle* h, int32_t* launches) {
  NEED_FWD();
  if (!launches) return CTD_E_INVALID;
  *launches = h->last_launches;
  return CTD_OK;
}

extern "C" int ctd_timer_start(ctd_handle* h) {
  if (!h) return CTD_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventRecord(h->tev0, h->stream));
  return CTD_OK;
}
extern "C" int ctd_timer_stop(ctd_handle* h, float* ms) {
  if (!h || !ms) return CTD_E_INVALID;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventRecord(h->tev1, h->stream));
  CK(cudaEventSynchronize(h->tev1));
  CK(cudaEventElapsedTime(ms, h->tev0, h->tev1));
  return CTD_OK;
}

extern "C" int ctd_profile_forward(ctd_handle* h, const uint8_t* pages, int32_t n, int32_t ph, int32_t pw,
                                   int32_t pages_on_device, float* op_ms, int32_t cap) {
  if (!h || !pages || !op_ms) return CTD_E_INVALID;
  const int need = int(h->ops.size()) + 2;
  if (cap < need) return ctd_fail(h, CTD_E_INVALID, "op_ms needs %d entries", need);
  if (n < 1 || n > h->cfg.max_batch || ph % 64 || pw % 64 || ph > h->cfg.max_h || pw > h->cfg.max_w)
    return ctd_fail(h, CTD_E_SHAPE, "bad shape");
  CK(cudaSetDevice(h->cfg.device));
  while (int(h->op_events.size()) < need + 1) {
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    h->op_events.push_back(e);
  }
  auto key = std::make_tuple(int(n), int(ph), int(pw));
  auto it = h->plans.find(key);
  if (it == h->plans.end()) {
    ShapePlan sp;
    if (int rc = build_plans(h, n, ph, pw, sp)) return rc;
    it = h->plans.emplace(key, std::move(sp)).first;
  }
  CK(cudaMemcpyAsync(h->d_pages, pages, size_t(n) * ph * pw * 3,
                     pages_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  int launches = 0;
  if (int rc = run_ops(h, n, ph, pw, it->second, &launches, true)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  const int nev = h->cfg.debug_skip_postproc ? int(h->ops.size()) : need;
  for (int i = 0; i < need; ++i) op_ms[i] = 0.f;
  for (int i = 0; i < nev; ++i) CK(cudaEventElapsedTime(&op_ms[i], h->op_events[i], h->op_events[i + 1]));
  h->n = n; h->ph = ph; h->pw = pw;
  h->have_forward = true;
  h->last_launches = launches;
  return CTD_OK;
}

extern "C" int ctd_get_device_outputs(ctd_handle* h, ctd_device_outputs* out) {
  NEED_FWD();
  if (!out) return CTD_E_INVALID;
  out->stream = h->stream;
  out->mask_u8 = h->d_mask_u8;
  out->det = h->d_det;
  out->det_count = h->d_det_count;
  out->bitmap = h->d_bitmap;
  out->labels = h->d_labels;
  out->n_labels = h->d_nlabels;
  out->line_boxes = h->d_line_boxes;
  out->line_scores = h->d_line_scores;
  out->line_count = h->d_line_count;
  out->results_base = h->d_mask_u8;
  out->results_bytes = h->results_bytes;
  return CTD_OK;
}

template <typename T>
__global__ void to_f32_kernel(const T* src, float* dst, size_t n) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = float(src[i]);
}

extern "C" int ctd_debug_read_buffer(ctd_handle* h, int32_t buf, float* out, size_t out_elems) {
  NEED_FWD();
  if (buf < 0 || buf >= int(h->bufs.size()) || !out) return ctd_fail(h, CTD_E_INVALID, "bad buffer id");
  const ctd_bufdesc& b = h->bufs[buf];
  const size_t elems = size_t(h->n) * (h->ph / b.down) * (h->pw / b.down) * b.channels;
  if (out_elems < elems) return ctd_fail(h, CTD_E_INVALID, "buffer %d holds %zu elements", buf, elems);
  float* tmp = nullptr;
  CK(cudaMalloc(&tmp, elems * 4));
  if (h->elem == 4) to_f32_kernel<float><<<unsigned((elems + 255) / 256), 256, 0, h->stream>>>(static_cast<float*>(h->d_buf[buf]), tmp, elems);
  else to_f32_kernel<__half><<<unsigned((elems + 255) / 256), 256, 0, h->stream>>>(static_cast<__half*>(h->d_buf[buf]), tmp, elems);
  cudaError_t e = cudaMemcpyAsync(out, tmp, elems * 4, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  CK(e);
  return CTD_OK;
}

template <typename T>
__global__ void from_f32_kernel(const float* src, T* dst, size_t n) {
  size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = T(src[i]);
}

extern "C" int ctd_debug_write_buffer(ctd_handle* h, int32_t buf, const float* in, int32_t n, int32_t ph, int32_t pw) {
  if (!h || !in) return CTD_E_INVALID;
  if (buf < 0 || buf >= int(h->bufs.size())) return ctd_fail(h, CTD_E_INVALID, "bad buffer id");
  if (n < 1 || n > h->cfg.max_batch || ph > h->cfg.max_h || pw > h->cfg.max_w) return ctd_fail(h, CTD_E_CAPACITY, "shape");
  CK(cudaSetDevice(h->cfg.device));
  const ctd_bufdesc& b = h->bufs[buf];
  const size_t elems = size_t(n) * (ph / b.down) * (pw / b.down) * b.channels;
  float* tmp = nullptr;
  CK(cudaMalloc(&tmp, elems * 4));
  cudaError_t e = cudaMemcpyAsync(tmp, in, elems * 4, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    if (h->elem == 4) from_f32_kernel<float><<<unsigned((elems + 255) / 256), 256, 0, h->stream>>>(tmp, static_cast<float*>(h->d_buf[buf]), elems);
    else from_f32_kernel<__half><<<unsigned((elems + 255) / 256), 256, 0, h->stream>>>(tmp, static_cast<__half*>(h->d_buf[buf]), elems);
    if (h->cfg.precision == CTD_PREC_SPLIT_TC) {
      const size_t npix = elems / b.channels;
      __half* hi = static_cast<__half*>(h->d_buf16[buf]);
      e = split_planes_launch(static_cast<const float*>(h->d_buf[buf]), hi, hi + elems, npix, b.channels, b.channels, h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(tmp);
  CK(e);
  return CTD_OK;
}

extern "C" int ctd_debug_run_ops(ctd_handle* h, const uint8_t* pages, int32_t n, int32_t ph, int32_t pw, int32_t first_op,
                                 int32_t last_op) {
  if (!h) return CTD_E_INVALID;
  if (first_op < 0 || last_op >= int(h->ops.size()) || first_op > last_op) return ctd_fail(h, CTD_E_INVALID, "bad op range");
  if (n < 1 || n > h->cfg.max_batch || ph % 64 || pw % 64 || ph > h->cfg.max_h || pw > h->cfg.max_w || ph < 64 || pw < 64)
    return ctd_fail(h, CTD_E_SHAPE, "bad shape");
  CK(cudaSetDevice(h->cfg.device));
  auto key = std::make_tuple(int(n), int(ph), int(pw));
  auto it = h->plans.find(key);
  if (it == h->plans.end()) {
    ShapePlan sp;
    if (int rc = build_plans(h, n, ph, pw, sp)) return rc;
    it = h->plans.emplace(key, std::move(sp)).first;
  }
  if (pages) CK(cudaMemcpyAsync(h->d_pages, pages, size_t(n) * ph * pw * 3, cudaMemcpyHostToDevice, h->stream));
  int cnt = 0;
  for (int i = first_op; i <= last_op; ++i)
    if (int rc = run_one_op(h, size_t(i), n, ph, pw, it->second, &cnt)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  h->n = n; h->ph = ph; h->pw = pw;
  h->have_forward = true;
  h->last_launches = cnt;
  return CTD_OK;
}

int cc_device(ctd_handle* h, const uint8_t* d_img, int ih, int iw, int stats_cap, int32_t** d_stats, int32_t* n_labels) {
  const size_t px = size_t(ih) * iw;
  // own grow-on-demand scratch (any page size, independent of the net-input workspace; the results of the last
  // forward stay intact): labels | 3 ints/px of CCL scratch (reused for the stats table) | n_labels
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t o_scr = al(px * 4);
  const size_t scr_bytes = al(std::max(px * 12, size_t(stats_cap > 0 ? stats_cap : 0) * 5 * 4));
  const size_t o_nl = o_scr + scr_bytes, need = o_nl + 256;
  if (need > h->cc_scratch_cap) {
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_cc_scratch);
    h->d_cc_scratch = nullptr;
    h->cc_scratch_cap = 0;
    CK(cudaMalloc(&h->d_cc_scratch, need + need / 4));
    h->cc_scratch_cap = need + need / 4;
  }
  uint8_t* base = static_cast<uint8_t*>(h->d_cc_scratch);
  int32_t* d_labels = reinterpret_cast<int32_t*>(base);
  int32_t* d_scr = reinterpret_cast<int32_t*>(base + o_scr);
  int32_t* d_nl = reinterpret_cast<int32_t*>(base + o_nl);
  CK(ccl_launch(d_img, 1, ih, iw, d_labels, d_scr, d_nl, h->stream));
  if (stats_cap > 0) CK(ccl_stats_launch(d_labels, ih, iw, d_scr, stats_cap, h->stream));   // scratch is free after ccl_launch
  CK(cudaMemcpyAsync(n_labels, d_nl, 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (d_stats) *d_stats = d_scr;
  return CTD_OK;
}

extern "C" int ctd_connected_components(ctd_handle* h, const uint8_t* img, int32_t ih, int32_t iw, int32_t* labels,
                                        int32_t* stats, int32_t stats_cap, int32_t* n_labels) {
  if (!h || !img || !labels || !n_labels) return CTD_E_INVALID;
  if (ih < 1 || iw < 1 || size_t(ih) * iw > (size_t(1) << 28)) return ctd_fail(h, CTD_E_SHAPE, "bad image size %dx%d", ih, iw);
  CK(cudaSetDevice(h->cfg.device));
  const size_t px = size_t(ih) * iw;
  if (int rc = ensure_io_scratch(h, px + 256)) return rc;
  CK(cudaMemcpyAsync(h->d_io_scratch, img, px, cudaMemcpyHostToDevice, h->stream));
  int32_t* d_stats = nullptr;
  if (int rc = cc_device(h, h->d_io_scratch, ih, iw, (stats && stats_cap > 0) ? stats_cap : 0, &d_stats, n_labels)) return rc;
  CK(cudaMemcpyAsync(labels, h->d_cc_scratch, px * 4, cudaMemcpyDeviceToHost, h->stream));
  if (stats && stats_cap > 0) CK(cudaMemcpyAsync(stats, d_stats, size_t(stats_cap) * 5 * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return CTD_OK;
}

extern "C" int ctd_nms(ctd_handle* h, const float* pred, int32_t rows, float conf_thresh, float iou_thresh, float* det,
                       int32_t* det_count) {
  if (!h || !pred || !det || !det_count) return CTD_E_INVALID;
  const int no = 5 + h->cfg.nc;
  if (rows > rows_per_image(h->cfg.max_h, h->cfg.max_w) * h->cfg.max_batch)
    return ctd_fail(h, CTD_E_CAPACITY, "too many prediction rows");
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(h->d_blks, pred, size_t(rows) * no * 4, cudaMemcpyHostToDevice, h->stream));
  CK(nms_launch(h->d_blks, 1, rows, h->cfg.nc, conf_thresh, iou_thresh, h->nms, h->d_det, h->d_det_count, h->stream));
  CK(cudaMemcpyAsync(det, h->d_det, 300 * 6 * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(det_count, h->d_det_count, 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_forward = false;
  return CTD_OK;
}
