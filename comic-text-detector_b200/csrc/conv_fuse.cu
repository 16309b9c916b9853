// Fused Bottleneck for sm_100a: y = x + act(conv3x3(act(conv1x1(x)))) in ONE kernel (reference: Bottleneck.forward,
// models/yolov5/common.py:94-104, as used inside C3, common.py:126-138 and basemodel.py:21-45).
//
// The unfused engine runs the 1x1 (c_ -> c_) and the 3x3 (c_ -> c_, + residual) as two tcgen05 kernels and the
// c_-channel intermediate t makes a round trip through HBM (write, read with halo), plus one more launch with its
// prologue / tail per Bottleneck.  Here the intermediate never leaves the SM:
//
//   TMA      : ONE box per tile brings the (8+2) x (16+2) halo block of x (180 pixels x c_ channels, 128B/64B swizzle)
//   GEMM 1   : t = W1 * x on ALL 180 halo pixels: two M=128 tcgen05.mma row blocks over the plain rows of the box
//              (rows 180..255 of the second block read whatever follows in shared memory; their results are never used)
//   epilogue1: tcgen05.ld -> bias + activation -> fp16 -> written with st.shared into a second halo block T in exactly
//              the swizzled K-major layout a TMA load would have produced; pixels outside the image are written as ZERO
//              (they are the 3x3 convolution's padding, not act(bias)); fence.proxy.async + mbarrier hand-over
//   GEMM 2   : the nine taps of the 3x3 are matrix-descriptor views into T (start row (dy+1)*10 + (dx+1), SBO = 10 rows),
//              the same trick as conv_halo_kernel; W1 and the nine W2 tap matrices stay resident in shared memory
//   epilogue2: bias + activation + residual (x is still in shared memory: centre rows of the halo block) -> fp16 -> TMA
//              store (c_ = 64) or 64-byte rows (c_ = 32)
//
// Two tiles are in flight per CTA (one per epilogue warpgroup; TMEM: 2 x (2 c_ + c_) columns), the MMA thread
// interleaves GEMM 1 of tile i+2 behind GEMM 2 of tile i.  Storage points are identical to the unfused path (t is
// rounded to fp16 exactly where the unfused engine stores it, accumulation order per output is the same), so the
// results are BIT-IDENTICAL to the two-kernel sequence (tests/test_gpu_fuse.py).  The destination is a different
// buffer than the source (neighbouring CTAs read the halo of x while this one writes y).
#include <cstring>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "kernels.h"
#include "ptx.cuh"

namespace ctd {

namespace {

constexpr int kThreads = 384;   // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 / 8-11 epilogue warpgroups
constexpr int kEpiWarp0 = 4;
constexpr int kTW = 8, kTH = 16;          // output tile (pixels)
constexpr int kHW = kTW + 2, kHH = kTH + 2;
constexpr int kHaloRows = kHW * kHH;      // 180

template <int C>
struct BnCfg {
  static constexpr int kRowBytes = C * 2;                          // 128 (c_ = 64) or 64 (c_ = 32)
  static constexpr int kW1Bytes = C * kRowBytes;
  static constexpr int kW2Bytes = 9 * C * kRowBytes;
  static constexpr int kWBytes = (kW1Bytes + kW2Bytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = (kHaloRows * kRowBytes + 1023) / 1024 * 1024;
  static constexpr int kXStages = C == 64 ? 3 : 6;
  static constexpr int kTmemCols = C == 64 ? 512 : 256;            // 2 slots x (2C + C), power of two
  static constexpr int kBarBytes = 512;
  // the second GEMM-1 row block of the LAST x stage reads 76 rows past the stage: the two T blocks follow it
  static constexpr size_t kSmem = 1024 + size_t(kWBytes) + size_t(kXStages + 2) * kStageBytes + kBarBytes + 2 * C * 4 + 64;
};

template <int ACT>
__device__ __forceinline__ float act_fn(float v) {
  if constexpr (ACT == CTD_ACT_SILU) return __fdividef(v, 1.0f + exp_neg_fast(v));
  else if constexpr (ACT == CTD_ACT_LEAKY) return fmaxf(v, 0.1f * v);
  else if constexpr (ACT == CTD_ACT_RELU) return fmaxf(v, 0.f);
  else if constexpr (ACT == CTD_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + exp_neg_fast(v));
  else return v;
}

// byte offset of 16-byte chunk j of row r inside a swizzled K-major block whose base is 1024-byte aligned
template <int C>
__device__ __forceinline__ uint32_t swz(int r, int j) {
  if constexpr (C == 64) return uint32_t(r) * 128u + uint32_t((j ^ (r & 7)) << 4);        // SWIZZLE_128B
  else return uint32_t(r) * 64u + uint32_t((j ^ ((r >> 1) & 3)) << 4);                   // SWIZZLE_64B
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// 32 accumulator columns -> bias + activation (+ residual) -> 4 x 16-byte fp16 chunks
template <int ACT, bool RES>
__device__ __forceinline__ void finish32(const uint32_t (&v)[32], const float* __restrict__ bias_s, const uint4 (&res)[4],
                                         uint4 (&o)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float f[8];
    const float4 b0 = *reinterpret_cast<const float4*>(bias_s + q * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(bias_s + q * 8 + 4);
    f[0] = act_fn<ACT>(__uint_as_float(v[q * 8 + 0]) + b0.x);
    f[1] = act_fn<ACT>(__uint_as_float(v[q * 8 + 1]) + b0.y);
    f[2] = act_fn<ACT>(__uint_as_float(v[q * 8 + 2]) + b0.z);
    f[3] = act_fn<ACT>(__uint_as_float(v[q * 8 + 3]) + b0.w);
    f[4] = act_fn<ACT>(__uint_as_float(v[q * 8 + 4]) + b1.x);
    f[5] = act_fn<ACT>(__uint_as_float(v[q * 8 + 5]) + b1.y);
    f[6] = act_fn<ACT>(__uint_as_float(v[q * 8 + 6]) + b1.z);
    f[7] = act_fn<ACT>(__uint_as_float(v[q * 8 + 7]) + b1.w);
    if constexpr (RES) {
      const __half2* rh = reinterpret_cast<const __half2*>(&res[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 rf = __half22float2(rh[e]);
        f[2 * e] += rf.x;
        f[2 * e + 1] += rf.y;
      }
    }
    o[q].x = pack2(f[0], f[1]); o[q].y = pack2(f[2], f[3]); o[q].z = pack2(f[4], f[5]); o[q].w = pack2(f[6], f[7]);
  }
}

__device__ __forceinline__ void issue_k(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc0, int ksteps) {
  umma_f16(tmem_d, ad, bd, idesc, acc0);
  umma_f16(tmem_d, ad + 2, bd + 2, idesc, 1u);
  if (ksteps >= 4) {
    umma_f16(tmem_d, ad + 4, bd + 4, idesc, 1u);
    umma_f16(tmem_d, ad + 6, bd + 6, idesc, 1u);
  }
}

template <int C>
__global__ void __launch_bounds__(kThreads, 1) conv_bneck_kernel(const __grid_constant__ BneckParams p) {
  using Cfg = BnCfg<C>;
  constexpr uint32_t RB = Cfg::kRowBytes;
  constexpr int S = Cfg::kXStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base;
  const uint32_t x_base = w_base + Cfg::kWBytes;
  const uint32_t t_base = x_base + S * Cfg::kStageBytes;
  const uint32_t bar_base = t_base + 2 * Cfg::kStageBytes;
  // barriers: x_full[8] | x_empty[8] | acc1_full[2] | t_full[2] | acc2_full[2] | acc_empty[2] | w | tmem ptr
  const uint32_t x_full = bar_base, x_empty = bar_base + 64, acc1_full = bar_base + 128, t_full = bar_base + 144;
  const uint32_t acc2_full = bar_base + 160, acc_empty = bar_base + 176, w_bar = bar_base + 192, tmem_ptr_addr = bar_base + 200;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const size_t bar_off = size_t(Cfg::kWBytes) + size_t(S + 2) * Cfg::kStageBytes;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 200);
  float* bias_s = reinterpret_cast<float*>(smem_gen + bar_off + Cfg::kBarBytes);   // bias1[C] | bias2[C]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.n_img * tiles_per_img;
  const int my_tiles = (total_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.x_map);
    prefetch_tensormap(&p.w1_map);
    prefetch_tensormap(&p.w2_map);
    if (C == 64) prefetch_tensormap(&p.o_map);
    for (int s = 0; s < S; ++s) {
      mbar_init(x_full + 8 * s, 1);
      mbar_init(x_empty + 8 * s, 1 + 128);   // GEMM 1 commit + the 128 epilogue-2 threads (residual reads)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc1_full + 8 * s, 1);
      mbar_init(t_full + 8 * s, 128);
      mbar_init(acc2_full + 8 * s, 1);
      mbar_init(acc_empty + 8 * s, 128);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_addr, Cfg::kTmemCols);
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) bias_s[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  auto decode = [&](int t, int& img, int& y0, int& x0) {
    img = t / tiles_per_img;
    const int trem = t - img * tiles_per_img;
    const int ty = trem / p.tiles_x;
    y0 = ty * kTH;
    x0 = (trem - ty * p.tiles_x) * kTW;
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (elect_one()) {
      mbar_arrive_expect_tx(w_bar, uint32_t(Cfg::kW1Bytes + Cfg::kW2Bytes));
      tma_load_2d(w_base, &p.w1_map, w_bar, 0, 0);
      for (int tap = 0; tap < 9; ++tap)
        tma_load_2d(w_base + uint32_t(Cfg::kW1Bytes) + uint32_t(tap) * uint32_t(C) * RB, &p.w2_map, w_bar, tap * C, 0);
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        int img, y0, x0;
        decode(t, img, y0, x0);
        const int stage = it % S;
        mbar_wait_relaxed(x_empty + 8 * stage, ((it / S) & 1) ^ 1);
        mbar_arrive_expect_tx(x_full + 8 * stage, uint32_t(kHaloRows) * RB);
        tma_load_4d(x_base + uint32_t(stage) * Cfg::kStageBytes, &p.x_map, x_full + 8 * stage, 0, x0 - 1, y0 - 1, img);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (one lane) =======================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(C);
      constexpr int ksteps = C / 16;
      const uint64_t a1_desc0 = make_kmajor_desc(x_base, RB);
      const uint64_t b1_desc = make_kmajor_desc(w_base, RB);
      const uint64_t a2_desc0 = make_kmajor_desc_ex(t_base, RB, uint32_t(kHW) * RB, 0u);
      const uint64_t b2_desc0 = make_kmajor_desc(w_base + Cfg::kW1Bytes, RB);
      uint32_t tap_a[9], tap_b[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        tap_a[tap] = (uint32_t((dy + 1) * kHW + (dx + 1)) * RB) >> 4;
        tap_b[tap] = (uint32_t(tap) * uint32_t(C) * RB) >> 4;
      }
      mbar_wait(w_bar, 0);
      auto gemm1 = [&](int ti) {
        const int slot = ti & 1, stage = ti % S;
        mbar_wait(x_full + 8 * stage, (ti / S) & 1);
        tc_fence_after();
        // acc1[slot] was drained by epilogue 1 of tile ti-2: this thread has already waited on t_full for it
        const uint64_t ad = a1_desc0 + uint64_t((uint32_t(stage) * Cfg::kStageBytes) >> 4);
        issue_k(tmem_base + uint32_t((slot * 3 + 0) * C), ad, b1_desc, idesc, 0u, ksteps);
        issue_k(tmem_base + uint32_t((slot * 3 + 1) * C), ad + uint64_t((128u * RB) >> 4), b1_desc, idesc, 0u, ksteps);
        umma_commit(acc1_full + 8 * slot);
        umma_commit(x_empty + 8 * stage);
      };
      auto gemm2 = [&](int ti) {
        const int slot = ti & 1, use = ti >> 1;
        mbar_wait(t_full + 8 * slot, use & 1);            // T[slot] written (and acc1[slot] drained)
        mbar_wait(acc_empty + 8 * slot, (use & 1) ^ 1);   // acc2[slot] drained by epilogue 2 of tile ti-2
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t((slot * 3 + 2) * C);
        const uint64_t ad = a2_desc0 + uint64_t((uint32_t(slot) * Cfg::kStageBytes) >> 4);
        uint32_t acc = 0u;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          issue_k(tmem_d, ad + tap_a[tap], b2_desc0 + tap_b[tap], idesc, acc, ksteps);
          acc = 1u;
        }
        umma_commit(acc2_full + 8 * slot);
      };
      if (my_tiles > 0) gemm1(0);
      if (my_tiles > 1) gemm1(1);
      for (int ti = 0; ti < my_tiles; ++ti) {
        gemm2(ti);
        if (ti + 2 < my_tiles) gemm1(ti + 2);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // =============================== epilogue warpgroups =========================
    const int quad = warp & 3;
    const int group = (warp - kEpiWarp0) >> 2;   // tile parity = TMEM slot = T block
    const int row = quad * 32 + lane;
    const uint32_t lane_off = uint32_t(quad * 32) << 16;
    const bool lead = (threadIdx.x & 127) == 0;
    const uint32_t tblk = t_base + uint32_t(group) * Cfg::kStageBytes;
    const int py = row / kTW, px = row - py * kTW;
    const int hr_c = (py + 1) * kHW + (px + 1);   // this thread's output pixel inside the halo block
    int ti = group;
    for (int t = int(blockIdx.x) + group * int(gridDim.x); t < total_tiles; t += 2 * int(gridDim.x), ti += 2) {
      int img, y0, x0;
      decode(t, img, y0, x0);
      const int use = ti >> 1, stage = ti % S;
      // ---------------- epilogue 1: t = act(W1 x + b1) on the halo pixels -> T block (swizzled, zero outside) -------
      mbar_wait_relaxed(acc1_full + 8 * group, use & 1);
      tc_fence_after();
      if (C == 64) {
        // the T block doubles as the TMA-store staging tile of epilogue 2: the previous store must have read it
        if (lead) tma_store_wait_read();
        named_barrier_sync(1 + group, 128);
      }
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        if (m == 1 && quad * 32 >= kHaloRows - 128) break;   // warp-uniform: rows 128 + 32*quad .. are all >= 180
        const int hr = m * 128 + row;
        const int hy = hr / kHW, hx = hr - hy * kHW;
        const int gy = y0 - 1 + hy, gx = x0 - 1 + hx;
        const bool live = hr < kHaloRows;
        const bool inside = live && gy >= 0 && gy < p.gh && gx >= 0 && gx < p.gw;
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + uint32_t((group * 3 + m) * C + c0) + lane_off, v);
          tmem_ld_wait();
          if (!live) continue;
          uint4 o[4];
          const uint4 nores[4] = {};
          if (inside) {
            switch (p.act) {
              case CTD_ACT_SILU: finish32<CTD_ACT_SILU, false>(v, bias_s + c0, nores, o); break;
              case CTD_ACT_LEAKY: finish32<CTD_ACT_LEAKY, false>(v, bias_s + c0, nores, o); break;
              case CTD_ACT_RELU: finish32<CTD_ACT_RELU, false>(v, bias_s + c0, nores, o); break;
              default: finish32<CTD_ACT_NONE, false>(v, bias_s + c0, nores, o); break;
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) st_shared_v4(tblk + swz<C>(hr, (c0 >> 3) + q), o[q].x, o[q].y, o[q].z, o[q].w);
        }
      }
      fence_proxy_async();   // generic-proxy writes of T -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(t_full + 8 * group);
      // ---------------- epilogue 2: y = act(W2 * T + b2) (+ x) ---------------------------------------------------------
      const int gy = y0 + py, gx = x0 + px;
      const bool valid = gy < p.gh && gx < p.gw;
      const uint32_t xblk = x_base + uint32_t(stage) * Cfg::kStageBytes;
      mbar_wait_relaxed(acc2_full + 8 * group, use & 1);
      tc_fence_after();
      __half* out = p.dst + (size_t(img) * p.gh * p.gw + size_t(valid ? gy : 0) * p.gw + (valid ? gx : 0)) * p.dst_cstride + p.dst_coff;
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint4 res[4];
        if (p.residual) {
#pragma unroll
          for (int q = 0; q < 4; ++q) res[q] = ld_shared_v4(xblk + swz<C>(hr_c, (c0 >> 3) + q));
        }
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + uint32_t((group * 3 + 2) * C + c0) + lane_off, v);
        tmem_ld_wait();
        uint4 o[4];
#define CTD_FIN(ACT)                                                           \
  if (p.residual) finish32<ACT, true>(v, bias_s + C + c0, res, o);             \
  else finish32<ACT, false>(v, bias_s + C + c0, res, o);
        switch (p.act) {
          case CTD_ACT_SILU: CTD_FIN(CTD_ACT_SILU) break;
          case CTD_ACT_LEAKY: CTD_FIN(CTD_ACT_LEAKY) break;
          case CTD_ACT_RELU: CTD_FIN(CTD_ACT_RELU) break;
          default: CTD_FIN(CTD_ACT_NONE) break;
        }
#undef CTD_FIN
        if constexpr (C == 64) {
          // staging tile = the first 128 rows of this group's T block (GEMM 2 has completed: acc2_full), 128B swizzle
#pragma unroll
          for (int q = 0; q < 4; ++q) st_shared_v4(tblk + swz<64>(row, (c0 >> 3) + q), o[q].x, o[q].y, o[q].z, o[q].w);
        } else {
          if (valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(out + c0 + q * 8) = o[q];
          }
        }
      }
      // this thread is done with TMEM (tcgen05.wait::ld above) and with the x block
      tc_fence_before();
      mbar_arrive(acc_empty + 8 * group);
      mbar_arrive(x_empty + 8 * stage);
      if constexpr (C == 64) {
        fence_proxy_async();
        named_barrier_sync(1 + group, 128);
        if (lead) {
          tma_store_4d(&p.o_map, tblk, 0, x0, y0, img);
          tma_store_commit();
        }
      }
    }
    if (C == 64 && lead) tma_store_wait_all();   // shared memory must outlive the bulk stores
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =========================================================================================
// Seg tail as ONE GEMM + col2im: the final ConvTranspose2d(64 -> 1, 4x4, s2, p1) + Sigmoid of UnetHead
// (basemodel.py:57-60) + postprocess_mask (inference.py:96-99).
//
// conv_halo_kernel<16> runs it as a 3x3 convolution with the four sub-pixel phases as N: 9 taps x 4 K steps = 36
// M128 x N16 MMAs per 128 pixels, and an N <= 64 MMA costs the same ~70 issue cycles as N = 64, so the layer is
// issue-bound at 12 % of the tensor pipe (242 us per 16 pages).  Here the 16 kernel positions are the N of a 1x1 GEMM,
//     P[pixel][ky*4+kx] = sum_c x[pixel][c] * w[c][ky][kx]            (4 MMAs per 128 pixels),
// computed on a (16+2) x (12+2) halo block of input pixels (one TMA box, two M = 128 row blocks), and the epilogue does
// the col2im: P goes through shared memory and every input pixel of the tile interior gathers the 2 x 2 taps of its
// four output pixels from itself and its 8 neighbours,
//     out[2y+py][2x+px] = sum over (dy,ky) in T[py], (dx,kx) in T[px] of P[y+dy][x+dx][ky][kx],  T[0] = {(0,1),(-1,3)},
//     T[1] = {(0,2),(+1,0)}, then sigmoid -> f32 mask and trunc(255 * s) -> u8 mask.  The kernel is HBM-bound (reads the
// 64-channel input once, plus the halo from L2).
constexpr int kSTW = 16, kSTH = 12;                 // interior input pixels per tile
constexpr int kSHW = kSTW + 2, kSHH = kSTH + 2;     // 18 x 14 = 252 halo rows (two M = 128 row blocks)
constexpr int kSRows = kSHW * kSHH;
constexpr int kSStageBytes = 32 * 1024;
constexpr int kSStages = 4;
constexpr int kSAcc = 8;                            // TMEM ring: 8 tiles x (2 x 16 columns)
constexpr int kSPBytes = 256 * 64;                  // P of one tile: 256 rows x 16 fp32
constexpr size_t kSegSmem = 1024 + 2048 + size_t(kSStages) * kSStageBytes + 2 * kSPBytes + 512;

__device__ __forceinline__ uint32_t p_addr(uint32_t base, int row, int q) {   // 16-byte chunk q of P row `row`
  return base + uint32_t(row) * 64u + uint32_t((q ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ float p_ld(uint32_t base, int row, int k) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(p_addr(base, row, k >> 2) + uint32_t((k & 3) << 2)) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) conv_segtail_kernel(const __grid_constant__ SegTailParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = smem_base;                       // [16 kernel positions][64 channels] fp16, 128B swizzle
  const uint32_t x_base = w_base + 2048;
  const uint32_t p_base = x_base + kSStages * kSStageBytes;
  const uint32_t bar_base = p_base + 2 * kSPBytes;
  // barriers: x_full[4] | x_empty[4] | acc_full[8] | acc_empty[8] | w | tmem ptr
  const uint32_t x_full = bar_base, x_empty = bar_base + 32, acc_full = bar_base + 64, acc_empty = bar_base + 128;
  const uint32_t w_bar = bar_base + 192, tmem_ptr_addr = bar_base + 200;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + 2048 + size_t(kSStages) * kSStageBytes + 2 * kSPBytes + 200);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int total_tiles = p.n_img * tiles_per_img;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.x_map);
    prefetch_tensormap(&p.w_map);
    for (int s = 0; s < kSStages; ++s) { mbar_init(x_full + 8 * s, 1); mbar_init(x_empty + 8 * s, 1); }
    for (int s = 0; s < kSAcc; ++s) { mbar_init(acc_full + 8 * s, 1); mbar_init(acc_empty + 8 * s, 128); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr_addr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  auto decode = [&](int t, int& img, int& y0, int& x0) {
    img = t / tiles_per_img;
    const int trem = t - img * tiles_per_img;
    const int ty = trem / p.tiles_x;
    y0 = ty * kSTH;
    x0 = (trem - ty * p.tiles_x) * kSTW;
  };

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(w_bar, 2048u);
      tma_load_2d(w_base, &p.w_map, w_bar, 0, 0);
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        int img, y0, x0;
        decode(t, img, y0, x0);
        const int stage = it % kSStages;
        mbar_wait_relaxed(x_empty + 8 * stage, ((it / kSStages) & 1) ^ 1);
        mbar_arrive_expect_tx(x_full + 8 * stage, uint32_t(kSRows) * 128u);
        tma_load_4d(x_base + uint32_t(stage) * kSStageBytes, &p.x_map, x_full + 8 * stage, 0, x0 - 1, y0 - 1, img);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(16);
      const uint64_t a_desc0 = make_kmajor_desc(x_base, 128);
      const uint64_t b_desc = make_kmajor_desc(w_base, 128);
      mbar_wait(w_bar, 0);
      int ti = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
        const int stage = ti % kSStages, as = ti % kSAcc;
        mbar_wait(acc_empty + 8 * as, ((ti / kSAcc) & 1) ^ 1);
        mbar_wait(x_full + 8 * stage, (ti / kSStages) & 1);
        tc_fence_after();
        const uint64_t ad = a_desc0 + uint64_t((uint32_t(stage) * kSStageBytes) >> 4);
        issue_k(tmem_base + uint32_t(as * 32), ad, b_desc, idesc, 0u, 4);
        issue_k(tmem_base + uint32_t(as * 32 + 16), ad + uint64_t((128u * 128u) >> 4), b_desc, idesc, 0u, 4);
        umma_commit(acc_full + 8 * as);
        umma_commit(x_empty + 8 * stage);
      }
    }
  } else if (warp >= kEpiWarp0) {
    const int quad = warp & 3;
    const int group = (warp - kEpiWarp0) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = uint32_t(quad * 32) << 16;
    const uint32_t pb = p_base + uint32_t(group) * kSPBytes;
    const size_t ow = size_t(p.gw) * 2, oh = size_t(p.gh) * 2;
    int ti = group;
    for (int t = int(blockIdx.x) + group * int(gridDim.x); t < total_tiles; t += 2 * int(gridDim.x), ti += 2) {
      int img, y0, x0;
      decode(t, img, y0, x0);
      const int as = ti % kSAcc;
      mbar_wait_relaxed(acc_full + 8 * as, (ti / kSAcc) & 1);
      tc_fence_after();
      named_barrier_sync(1 + group, 128);   // every thread of the group has finished reading the previous tile's P
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + uint32_t(as * 32 + m * 16) + lane_off, v);
        tmem_ld_wait();
        const int r = m * 128 + row;
#pragma unroll
        for (int q = 0; q < 4; ++q) st_shared_v4(p_addr(pb, r, q), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      tc_fence_before();
      mbar_arrive(acc_empty + 8 * as);
      named_barrier_sync(1 + group, 128);   // P of this tile is complete
#pragma unroll 1
      for (int e = row; e < kSTW * kSTH; e += 128) {
        const int iy = e / kSTW, ix = e - iy * kSTW;
        const int gy = y0 + iy, gx = x0 + ix;
        if (gy >= p.gh || gx >= p.gw) continue;
        const int r = (iy + 1) * kSHW + (ix + 1);
        float o[2][2];
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
          for (int px = 0; px < 2; ++px) {
            // taps (d, k): parity 0 -> (0,1), (-1,3); parity 1 -> (0,2), (+1,0)
            const int dy1 = py ? 1 : -1, ky0 = py ? 2 : 1, ky1 = py ? 0 : 3;
            const int dx1 = px ? 1 : -1, kx0 = px ? 2 : 1, kx1 = px ? 0 : 3;
            float a = p_ld(pb, r, ky0 * 4 + kx0);
            a += p_ld(pb, r + dx1, ky0 * 4 + kx1);
            a += p_ld(pb, r + dy1 * kSHW, ky1 * 4 + kx0);
            a += p_ld(pb, r + dy1 * kSHW + dx1, ky1 * 4 + kx1);
            o[py][px] = 1.0f / (1.0f + expf(-a));
          }
        const size_t o0 = (size_t(img) * oh + size_t(gy) * 2) * ow + size_t(gx) * 2;
#pragma unroll
        for (int py = 0; py < 2; ++py) {
          *reinterpret_cast<float2*>(p.seg_f32 + o0 + py * ow) = make_float2(o[py][0], o[py][1]);
          *reinterpret_cast<uchar2*>(p.seg_u8 + o0 + py * ow) = make_uchar2((uint8_t)(o[py][0] * 255.0f), (uint8_t)(o[py][1] * 255.0f));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

const char* encode(PFN_encodeTiled enc, CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims,
                   const cuuint64_t* strides_bytes, const cuuint32_t* box, int c) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = c == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled failed (bottleneck)";
}

}  // namespace

bool conv_bneck_supported(int c) { return c == 32 || c == 64; }

const char* conv_bneck_plan(BneckPlan& plan, PFN_encodeTiled enc, int n_img, int gh, int gw, int c, const void* src,
                            int src_cstride, int src_coff, const void* w16, const float* bias, __half* dst,
                            int dst_cstride, int dst_coff, int act, int residual, int num_sms) {
  if (!conv_bneck_supported(c)) return "bottleneck: unsupported channel count";
  if (src_coff % 8 || dst_coff % 8 || src_cstride % 8 || dst_cstride % 8) return "bottleneck: slices must be 16-byte aligned";
  if (act == CTD_ACT_SIGMOID) return "bottleneck: sigmoid activation not supported";
  BneckParams& p = plan.p;
  memset(&p, 0, sizeof(p));
  p.n_img = n_img; p.gh = gh; p.gw = gw;
  p.tiles_x = (gw + kTW - 1) / kTW;
  p.tiles_y = (gh + kTH - 1) / kTH;
  p.act = act; p.residual = residual;
  p.dst = dst; p.dst_cstride = dst_cstride; p.dst_coff = dst_coff;
  p.bias = bias;
  {
    const size_t cs = size_t(src_cstride);
    const char* base = static_cast<const char*>(src) + size_t(src_coff) * 2;
    cuuint64_t dims[4] = {cuuint64_t(c), cuuint64_t(gw), cuuint64_t(gh), cuuint64_t(n_img)};
    cuuint64_t str[3] = {cs * 2, cs * 2 * gw, cs * 2 * size_t(gw) * gh};
    cuuint32_t box[4] = {cuuint32_t(c), kHW, kHH, 1};
    if (const char* e = encode(enc, &p.x_map, base, 4, dims, str, box, c)) return e;
  }
  {
    cuuint64_t dims[2] = {cuuint64_t(c), cuuint64_t(c)};
    cuuint64_t str[1] = {cuuint64_t(c) * 2};
    cuuint32_t box[2] = {cuuint32_t(c), cuuint32_t(c)};
    if (const char* e = encode(enc, &p.w1_map, w16, 2, dims, str, box, c)) return e;
  }
  {
    const char* w2 = static_cast<const char*>(w16) + size_t(c) * c * 2;
    cuuint64_t dims[2] = {cuuint64_t(9 * c), cuuint64_t(c)};
    cuuint64_t str[1] = {cuuint64_t(9 * c) * 2};
    cuuint32_t box[2] = {cuuint32_t(c), cuuint32_t(c)};
    if (const char* e = encode(enc, &p.w2_map, w2, 2, dims, str, box, c)) return e;
  }
  if (c == 64) {
    const size_t cs = size_t(dst_cstride);
    const char* base = reinterpret_cast<const char*>(dst) + size_t(dst_coff) * 2;
    cuuint64_t dims[4] = {64, cuuint64_t(gw), cuuint64_t(gh), cuuint64_t(n_img)};
    cuuint64_t str[3] = {cs * 2, cs * 2 * gw, cs * 2 * size_t(gw) * gh};
    cuuint32_t box[4] = {64, kTW, kTH, 1};
    if (const char* e = encode(enc, &p.o_map, base, 4, dims, str, box, 64)) return e;
  }
  const int total = n_img * p.tiles_x * p.tiles_y;
  plan.c = c;
  plan.grid = dim3(unsigned(total < num_sms ? total : num_sms), 1, 1);
  plan.smem_bytes = c == 64 ? BnCfg<64>::kSmem : BnCfg<32>::kSmem;
  return nullptr;
}

const char* conv_segtail_plan(SegTailPlan& plan, PFN_encodeTiled enc, int n_img, int gh, int gw, const void* src, int src_cstride,
                              int src_coff, const void* w16, float* seg_f32, uint8_t* seg_u8, int num_sms) {
  if (src_coff % 8 || src_cstride % 8) return "seg tail: source slice must be 16-byte aligned";
  SegTailParams& p = plan.p;
  memset(&p, 0, sizeof(p));
  p.n_img = n_img; p.gh = gh; p.gw = gw;
  p.tiles_x = (gw + kSTW - 1) / kSTW;
  p.tiles_y = (gh + kSTH - 1) / kSTH;
  p.seg_f32 = seg_f32; p.seg_u8 = seg_u8;
  {
    const size_t cs = size_t(src_cstride);
    const char* base = static_cast<const char*>(src) + size_t(src_coff) * 2;
    cuuint64_t dims[4] = {64, cuuint64_t(gw), cuuint64_t(gh), cuuint64_t(n_img)};
    cuuint64_t str[3] = {cs * 2, cs * 2 * gw, cs * 2 * size_t(gw) * gh};
    cuuint32_t box[4] = {64, kSHW, kSHH, 1};
    if (const char* e = encode(enc, &p.x_map, base, 4, dims, str, box, 64)) return e;
  }
  {
    cuuint64_t dims[2] = {64, 16};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, 16};
    if (const char* e = encode(enc, &p.w_map, w16, 2, dims, str, box, 64)) return e;
  }
  const int total = n_img * p.tiles_x * p.tiles_y;
  plan.grid = dim3(unsigned(total < num_sms ? total : num_sms), 1, 1);
  plan.smem_bytes = kSegSmem;
  return nullptr;
}

cudaError_t conv_segtail_launch(const SegTailPlan& plan, cudaStream_t s) {
  conv_segtail_kernel<<<plan.grid, kThreads, plan.smem_bytes, s>>>(plan.p);
  return cudaGetLastError();
}

cudaError_t conv_bneck_init() {
  if (cudaError_t e0 = cudaFuncSetAttribute(conv_segtail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSegSmem));
      e0 != cudaSuccess)
    return e0;
  cudaError_t e = cudaFuncSetAttribute(conv_bneck_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BnCfg<64>::kSmem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv_bneck_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BnCfg<32>::kSmem));
}

cudaError_t conv_bneck_launch(const BneckPlan& plan, cudaStream_t s) {
  if (plan.c == 64) conv_bneck_kernel<64><<<plan.grid, kThreads, plan.smem_bytes, s>>>(plan.p);
  else conv_bneck_kernel<32><<<plan.grid, kThreads, plan.smem_bytes, s>>>(plan.p);
  return cudaGetLastError();
}

}  // namespace ctd
