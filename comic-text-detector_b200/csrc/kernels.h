// Launch-parameter structs and host entry points of every kernel in the engine.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctd_b200.h"

namespace ctd {

constexpr int kMaxTaps = 9;
constexpr int kMaxPhases = 4;

// Geometry shared by the tensor-core and CUDA-core convolution kernels.  A "grid" pixel is an
// output pixel for CONV and an input pixel (= one sub-pixel phase output) for DECONV4.
struct ConvGeom {
  int n_img;             // images in the batch
  int gh, gw;            // grid height/width (see above)
  int dst_h, dst_w;      // destination buffer spatial size
  int out_mul;           // 1 (conv) or 2 (deconv): dst pixel = grid pixel * out_mul + phase
  int n_phase;           // 1 or 4
  int taps;              // taps per phase: 1, 9 or 4
  int cin_total;         // sum of src_c
  int k_total;           // taps * cin_total
  int n_src;
  int src_c[CTD_MAX_SRC];
  int src_cstride[CTD_MAX_SRC];  // channels of the source buffer (element stride between pixels)
  int src_h, src_w;              // source spatial size (all sources agree)
  int in_stride;                 // 1 or 2 (conv stride)
  // per phase, per tap: source pixel = grid pixel * in_stride + (dy, dx)
  int8_t tap_dy[kMaxPhases][kMaxTaps];
  int8_t tap_dx[kMaxPhases][kMaxTaps];
  int cout, cout_pad;
  int dst_cstride, dst_coff;
  int act, residual;
};

// Fill taps/offset tables for an op (conv k=1/3 stride 1/2, deconv 4x4 s2 p1).
void fill_conv_geom_taps(ConvGeom& g, int kind, int ksize, int stride);

// ---------------------------------------------------------------------------------------
// tcgen05 implicit-GEMM convolution (conv_tc.cu)
struct alignas(64) ConvTcParams {
  CUtensorMap a_map[CTD_MAX_SRC][4];  // [source][parity]: parity maps only for stride-2 convs
  CUtensorMap b_map;                  // packed weights [n_phase*cout_pad][k_total], K-major
  CUtensorMap o_map[4];               // destination slice, one map per deconv phase (TMA-store epilogue)
  int use_tma_store;                  // 1: epilogue stages 64-channel chunks in smem and stores them by TMA
  ConvGeom g;
  int kb_elems;                       // channels per K block: 64 / 32 / 16 (swizzle 128/64/32 B)
  int src_kblocks[CTD_MAX_SRC];
  int tiles_x, tiles_y;               // 16x8-pixel tiles per image
  int8_t tap_map[kMaxPhases][kMaxTaps];  // parity map index per tap (stride 2), else 0
  __half* dst;
  const float* bias;
  // DETECT epilogue (dst == nullptr): decoded rows go to blks
  float* blks;
  int blks_rows_per_img;  // A
  int level_row0;         // first row of this pyramid level
  float det_stride;
  float anchor_wh[6];     // pixels
  int nc;
  // halo variant (conv_halo_kernel): weights resident in smem, ONE activation box per K block holds the tile
  // plus its halo and every filter tap is an MMA operand view into it
  int halo_lox, halo_loy;             // halo pixels before the tile in x / y
  int halo_w, halo_h;                 // halo block size in pixels (tile 8 x 16 + halo)
  int halo_stages, halo_stage_bytes;  // activation ring
  int halo_w_bytes;                   // resident weights of one phase: taps * kblocks * BN * kb * 2
  int hs_b_stages;                    // conv_hs_kernel: depth of the streamed-weight ring (halo_stages = A ring)
  // seg-tail epilogue (halo kernel, BN = 16): accumulator columns 0..3 are the sub-pixel phases of the final
  // ConvT 4x4 s2 (C -> 1); sigmoid -> f32 mask + truncated u8 mask at (2y+py, 2x+px)
  float* seg_f32;
  uint8_t* seg_u8;
  // split-fp16 mode (CTD_PREC_SPLIT_TC, conv_tc_kernel only): every fp32 operand x is carried as two fp16 planes
  // hi = fp16(x), lo = fp16(x - hi); a K block issues (hi,hi) + (lo,hi) + (hi,lo) into the same fp32 accumulator
  // (~22 significant bits per operand), and the epilogue writes FP32.  Activation planes: image n_img + i of the
  // same tensor map holds the lo plane of image i; weight lo rows follow the hi rows (row + split_row_off).
  int split;
  int split_img_off;
  int split_row_off;
};

struct ConvTcPlan {
  ConvTcParams p;
  int halo = 0;                       // 1: conv_halo_kernel, 2: conv_hs_kernel (halo A, streamed B), 3: conv_sw_kernel
  int block_n;
  dim3 grid;
  size_t smem_bytes;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Builds tensor maps + launch shape.  Returns nullptr on success, else an error string.
// split != 0: split-fp16 mode (see ConvTcParams::split): src_ptr are the [2*n_img][h][w][C] fp16 hi|lo plane buffers,
// w16 holds hi rows then lo rows, dst is an FP32 NHWC buffer.
const char* conv_tc_plan(ConvTcPlan& plan, PFN_encodeTiled enc, const ConvGeom& g, const void* const src_ptr[],
                         const int src_coff[], const void* w16, const float* bias, __half* dst, int split = 0);
// Stem in tensor-core form: 3 filter rows x (4-pixel window x 16 channels) over the padded space-to-depth page
// (`s2d`: [n][ph/2][pw/2 + 4][16] fp16), output [n][ph/2][pw/2][cstride] at channel offset `dst_coff`.
const char* conv_tc_plan_stem(ConvTcPlan& plan, PFN_encodeTiled enc, const void* s2d, int n, int ph, int pw,
                              const void* w16, const float* bias, __half* dst, int dst_cstride, int dst_coff, int cout,
                              int act);
// Halo variant for stride-1 3x3 convolutions and the 2x2-tap deconvolution phases whose weights fit in shared
// memory.  Sets plan.halo = 1 when the op is eligible, leaves it 0 (and returns nullptr) when not.
const char* conv_halo_plan(ConvTcPlan& plan, PFN_encodeTiled enc, const ConvGeom& g, const void* const src_ptr[],
                           const int src_coff[], const void* w16, const float* bias, __half* dst,
                           float* seg_f32 = nullptr, uint8_t* seg_u8 = nullptr);
// Halo activations + STREAMED weights for the wide stride-1 3x3 convolutions / deconvolution phases whose weights
// do not fit in shared memory (BN = 128 / 256).  Sets plan.halo = 2 when eligible.
const char* conv_hs_plan(ConvTcPlan& plan, PFN_encodeTiled enc, const ConvGeom& g, const void* const src_ptr[],
                         const int src_coff[], const void* w16, const float* bias, __half* dst);
// Swapped operands for the 128-wide stride-1 3x3 convolutions / deconvolution phases: the 128 output channels are the
// MMA's M (weights = A operand), 256 PIXELS (8 x 32 tile, halo views) are its N, so one instruction does twice the
// work of the pixel-major form at N = 128; the epilogue transposes through shared memory.  Sets plan.halo = 3.
const char* conv_sw_plan(ConvTcPlan& plan, PFN_encodeTiled enc, const ConvGeom& g, const void* const src_ptr[],
                         const int src_coff[], const void* w16, const float* bias, __half* dst);
// Stem through the halo kernel (window map of conv_tc_plan_stem, halo in y only).
const char* conv_halo_plan_stem(ConvTcPlan& plan, PFN_encodeTiled enc, const void* s2d, int n, int ph, int pw,
                                const void* w16, const float* bias, __half* dst, int dst_cstride, int dst_coff, int cout,
                                int act);
cudaError_t conv_tc_launch(const ConvTcPlan& plan, cudaStream_t s);
cudaError_t conv_tc_init();  // sets max dynamic smem attributes once

// ---------------------------------------------------------------------------------------
// Fused Bottleneck (conv_fuse.cu): y = [x +] act(conv3x3(act(conv1x1(x)))), c -> c -> c channels, c in {32, 64}; the
// intermediate stays in shared memory.  w16: W1 [c][c] followed by W2 [c][9*c] (K-major fp16, K = (tap, ci));
// bias: bias1[c] | bias2[c].  Source and destination are different buffers.
struct alignas(64) BneckParams {
  CUtensorMap x_map, w1_map, w2_map, o_map;
  int n_img, gh, gw;
  int tiles_x, tiles_y;     // 8 x 16-pixel tiles
  int act, residual;
  int dst_cstride, dst_coff;
  __half* dst;
  const __half* src;        // residual: x re-read per output pixel
  int src_cstride, src_coff;
  const float* bias;
};
struct BneckPlan {
  BneckParams p;
  int c;
  dim3 grid;
  size_t smem_bytes;
};
bool conv_bneck_supported(int c);
const char* conv_bneck_plan(BneckPlan& plan, PFN_encodeTiled enc, int n_img, int gh, int gw, int c, const void* src,
                            int src_cstride, int src_coff, const void* w16, const float* bias, __half* dst,
                            int dst_cstride, int dst_coff, int act, int residual, int num_sms);
cudaError_t conv_bneck_init();   // also sets the attributes of conv_segtail_kernel

// Seg tail (conv_fuse.cu): ConvTranspose2d(64 -> 1, 4x4, s2, p1) + sigmoid + u8 mask as ONE 1x1 GEMM over the 16 kernel
// positions + a col2im epilogue.  w16: [16 = ky*4+kx][64 channels] fp16; source must have exactly 64 channels.
struct alignas(64) SegTailParams {
  CUtensorMap x_map, w_map;
  int n_img, gh, gw;
  int tiles_x, tiles_y;     // 16 x 12 input pixels per tile
  float* seg_f32;
  uint8_t* seg_u8;
};
struct SegTailPlan {
  SegTailParams p;
  dim3 grid;
  size_t smem_bytes;
};
const char* conv_segtail_plan(SegTailPlan& plan, PFN_encodeTiled enc, int n_img, int gh, int gw, const void* src, int src_cstride,
                              int src_coff, const void* w16, float* seg_f32, uint8_t* seg_u8, int num_sms);
cudaError_t conv_segtail_launch(const SegTailPlan& plan, cudaStream_t s);
cudaError_t conv_bneck_launch(const BneckPlan& plan, cudaStream_t s);

// ---------------------------------------------------------------------------------------
// CUDA-core kernels (simt.cu): accurate/bisecting path and the thin layers.  T = float | __half.
struct ConvSimtParams {
  ConvGeom g;
  const void* src[CTD_MAX_SRC];  // already offset to the first channel read
  const void* w;                 // [n_phase*cout_pad][k_total] float (T=float) or __half (T=__half)
  const float* bias;
  void* dst;
  float* blks;                   // DETECT
  int blks_rows_per_img, level_row0, nc;
  float det_stride;
  float anchor_wh[6];
};
template <typename T>
cudaError_t conv_simt_launch(const ConvSimtParams& p, cudaStream_t s);

template <typename T>
cudaError_t stem_launch(const uint8_t* pages, int n, int h, int w, const float* wgt /*[32][108] (ky,kx,c)*/,
                        const float* bias, T* dst, int dst_cstride, int dst_coff, int cout, int act, cudaStream_t s);
// u8 BGR HWC page -> /255 -> space-to-depth(2): dst[n][h/2][w/2][16], channel = (dy*2+dx)*3 + c, 12..15 = 0
template <typename T>
cudaError_t s2d_launch(const uint8_t* pages, int n, int h, int w, T* dst, int dst_cstride, int dst_coff, int pitch_px,
                       int xoff, cudaStream_t s);
// fp32 NHWC channel slice -> fp16 hi / lo planes (split-fp16 mode): hi = fp16(x), lo = fp16(x - hi).
// src / hi / lo point at the first channel of the slice; `cstride` elements between pixels (same in all three).
cudaError_t split_planes_launch(const float* src, __half* hi, __half* lo, size_t npix, int c, int cstride,
                                cudaStream_t s);
template <typename T>
cudaError_t avgpool2_launch(const T* src, int n, int h, int w, int c, int src_cstride, T* dst, int dst_cstride,
                            cudaStream_t s);
template <typename T>
cudaError_t sppf_pool_launch(T* buf, int n, int h, int w, int c, int cstride, cudaStream_t s);  // in-place slots
template <typename T>
cudaError_t upsample2_launch(const T* src, int n, int h, int w, int c, int src_cstride, T* dst, int dst_cstride,
                             cudaStream_t s);
// seg tail: ConvT4x4s2p1 C->1 + sigmoid; writes f32 mask [n][2h][2w] and u8 mask (p*255 truncated)
template <typename T>
cudaError_t seg_tail_launch(const T* src, int n, int h, int w, int c, int cstride, const float* wgt /*[c][4][4]*/,
                            float* mask_f32, uint8_t* mask_u8, cudaStream_t s);
// DB tail on the 32-channel (binarize|thresh) map at 1/4 resolution -> lines f32 [n][2][4h][4w]
// and the thresholded bitmap u8 [n][4h][4w] (shrink > db_thresh).
template <typename T>
cudaError_t db_tail_launch(const T* src, int n, int h, int w, int cstride, const float* params, float* lines,
                           uint8_t* bitmap, float db_thresh, cudaStream_t s);

// ---------------------------------------------------------------------------------------
// post-processing (postproc.cu)
struct NmsWorkspace {
  float* cand;     // [n][cap][6]
  int* cand_count; // [n] candidates in `cand` (<= cap after nms_overflow_kernel)
  int* cand_total; // [n] candidates the page really had (> cap: the best `cap` by score were kept)
  float* sorted;   // [n][cap][6]
  unsigned long long* mask;  // [n][cap][cap/64]
  int cap;
};
cudaError_t nms_launch(const float* blks, int n, int rows, int nc, float conf, float iou, NmsWorkspace& ws,
                       float* det /*[n][300][6]*/, int* det_count, cudaStream_t s);
size_t nms_workspace_bytes(int n, int cap);
void nms_workspace_bind(NmsWorkspace& ws, void* base, int n, int cap);

// 8-connectivity labelling with OpenCV's label numbering; labels i32, n_labels incl. background.
cudaError_t ccl_launch(const uint8_t* img, int n, int h, int w, int32_t* labels, int32_t* scratch /*3*n*h*w ints*/,
                       int32_t* n_labels, cudaStream_t s);
cudaError_t ccl_stats_launch(const int32_t* labels, int h, int w, int32_t* stats, int cap, cudaStream_t s);

// SegDetectorRepresenter.boxes_from_bitmap (segrep.cu).  Lf = foreground union-find roots left in the CCL
// scratch by ccl_launch (first n*h*w ints).  boxes i16 [n][max_cand][4][2], scores f32 [n][max_cand].
size_t segrep_scratch_bytes(int n, int h, int w, int max_cand);
cudaError_t segrep_launch(const uint8_t* bitmap, const float* pred, size_t pred_page_stride, const int* Lf, int n, int h,
                          int w, int max_cand, float unclip_ratio, void* scratch, int16_t* boxes, float* scores,
                          int* n_contours, cudaStream_t s);
cudaError_t binarize_launch(const float* pred, size_t count, float thresh, uint8_t* bitmap, cudaStream_t s);

// refine_mask (refine.cu): one CTA per block window.  d_wins: n_wins x {x1,y1,x2,y2,(int64)pixel offset}
// cv2.resize INTER_LINEAR, uint8, 1 or 3 channels, bit-exact (resize.cu).  src rows are `src_pitch` bytes apart; the
// dh x dw result is written at the top-left of a canvas_h x canvas_w canvas whose remaining pixels are zeroed.
cudaError_t resize_linear_u8_launch(const uint8_t* src, int sh, int sw, size_t src_pitch, int channels, uint8_t* dst,
                                    int dh, int dw, int canvas_h, int canvas_w, cudaStream_t s);
size_t refine_scratch_bytes(size_t total_px);
// d_wins: RefineWin records {x1,y1,x2,y2,(int64) plane offset,(int) page,(int) pad} = 32 bytes; idx_small / idx_large:
// device index lists into d_wins (windows above refine_large_px() pixels are processed by a CTA cluster each).
cudaError_t refine_launch(const uint8_t* d_img, const uint8_t* d_mask, int H, int W, const void* d_wins, const int* d_idx_small,
                          int n_small, const int* d_idx_large, int n_large, size_t total_px, void* scratch, int refine_mode,
                          uint8_t* d_out, cudaStream_t s);
int refine_large_px();
size_t refine_win_bytes();
// phase-synchronous form (refine_mk.cu): one kernel per phase over all window pixels of the batch, windows cut into
// chunks of whole rows {win, y0, rows, pad} by the host; d_state: refine_mk_state_bytes(n_wins) bytes of per-window state
size_t refine_mk_state_bytes(int n_wins);
size_t refine_mk_chunk_bytes();
int refine_mk_chunk_px();
// the first n_multi_chunks records of d_chunks are the chunks of windows that span more than one chunk
cudaError_t refine_mk_launch(const uint8_t* d_img, const uint8_t* d_mask, int H, int W, const void* d_wins, int n_wins,
                             const void* d_chunks, int n_chunks, int n_multi_chunks, void* d_state, size_t total_px,
                             void* scratch, int refine_mode, uint8_t* d_out, cudaStream_t s);

}  // namespace ctd
