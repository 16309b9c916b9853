// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, operands in shared memory) as a
// function of N and of how many DIFFERENT TMEM accumulators consecutive instructions rotate over.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cuda_runtime.h>

#include "../../comic-text-detector_b200/csrc/ptx.cuh"

using namespace ctd;

template <int N>
__global__ void __launch_bounds__(128, 1) probe(int iters, int n_acc, int distinct_ab, int halo_view, int writers, int commit_every, const uint8_t* __restrict__ gsrc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 4 * 16384;          // 4 A tiles (128 x 64 fp16), 4 B tiles (N x 64)
  const uint32_t bar = b_base + 4 * N * 128, tptr = bar + 16, cbar = bar + 32, dbar = bar + 64;   // dbar: 8 barriers
  const uint32_t scratch = bar + 1024;   // 32 KB landing zone of the bulk copies
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(cbar, 1); for (int q = 0; q < 8; ++q) mbar_init(dbar + 8 * q, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(tptr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tptr - base));
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(N);
    // halo_view: A seen through a 10-pixel-wide halo block (8-row groups 1280 B apart, start 11 rows in)
    const uint64_t ad0 = halo_view ? make_kmajor_desc_ex(a_base + 11 * 128, 128, 1280, 0) : make_kmajor_desc(a_base, 128);
    const uint64_t bd0 = make_kmajor_desc(b_base, 128);
    long long t0 = 0, t1 = 0;
    if (leader) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t acc = tmem + uint32_t((i % n_acc) * N);
        const int tile = (distinct_ab && !halo_view) ? (i >> 2) & 3 : 0;
        const uint64_t ad = ad0 + uint64_t(tile * (16384 >> 4)) + uint64_t((i & 3) * 2);
        const uint64_t bd = bd0 + uint64_t(tile * ((N * 128) >> 4)) + uint64_t((i & 3) * 2);
        umma_f16(acc, ad, bd, idesc, i >= n_acc ? 1u : 0u);
        // every `commit_every` MMAs (power of two): one commit + `writers` cycles of issue-thread stall, the per-stage
        // overhead of a real main loop (barrier wait, descriptor bookkeeping)
        if ((i & (commit_every - 1)) == commit_every - 1) {
          umma_commit(dbar + 8 * ((i >> 3) & 7));
          const long long ts = clock64();
          while (clock64() - ts < writers) {}
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (leader) {
      t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  else if (warp == 2 && false) {
    // emulate the TMA operand fill: async-proxy bulk copies global -> shared, `writers` KB per MMA, 16 KB at a time
    if (elect_one()) {
      const int n_copies = iters * writers / 16;
      uint32_t par = 0;
      for (int i = 0; i < n_copies; ++i) {
        mbar_arrive_expect_tx(cbar, 16384);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         scratch + uint32_t(i & 1) * 16384),
                     "l"(gsrc + (size_t(blockIdx.x) * 64 + (i & 63)) * 16384), "r"(16384), "r"(cbar)
                     : "memory");
        mbar_wait(cbar, par);
        par ^= 1;
      }
    }
  } else if (warp == 3 && false) {
    // emulate the TMA fill traffic: `writers` x 16 bytes per thread per iteration into the operand area's tail
    const uint32_t dstp = a_base + 3 * 16384 + (threadIdx.x - 64) * 16;
    for (int i = 0; i < iters * writers; ++i)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dstp + uint32_t((i & 7) * 1024)), "r"(0x3c003c00u) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N>
void run(long long* d_out) {
  const size_t smem = 1024 + 4 * 16384 + 4 * N * 128 + 1024 + 32768 + 64;
  static uint8_t* gsrc = nullptr;
  if (!gsrc) { cudaMalloc(&gsrc, size_t(148) * 64 * 16384); cudaMemset(gsrc, 0, size_t(148) * 64 * 16384); }
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  const int iters = 4096;
  for (int commit_every = 4; commit_every <= 16; commit_every *= 2)
    for (int writers = 0; writers <= 600; writers += 300) {
      const int halo = 0;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        probe<N><<<148, 128, smem>>>(iters, 1, 1, halo, writers, commit_every, gsrc, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
      }
      printf("N=%3d  commit every %2d MMAs  + %3d cycles of issue-thread stall per commit : %.1f cycles per MMA (floor %d)\n", N,
             commit_every, writers, double(h) / iters, N / 2);
    }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run<64>(d_out);
  run<128>(d_out);
  run<256>(d_out);
  return 0;
}
