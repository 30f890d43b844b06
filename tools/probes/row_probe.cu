// Micro-benchmark of the row-marching MMA pattern (conv_row.cu): per input row one N=48 overwrite + one N=96 accumulate +
// 8 x N=144 accumulate MMAs whose D blocks move through a ring of 10 x 48 TMEM columns, against plain N=144 MMAs on a fixed D.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../larvanet_b200/csrc/lv_common.cuh"
namespace lv { void set_error(const char*, ...) {} void count_launch(int) {} int sm_count() { return 148; } }
using namespace lv;

__device__ __forceinline__ uint32_t idesc_n(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

// VAR 0: 10 x N=144 on D column 0                      (baseline)
// VAR 1: 10 x N=144, D column = ring position          (moving D)
// VAR 2: N=48 (overwrite) + N=96 + 8 x N=144, ring     (the kernel's pattern, no commits)
// VAR 3: VAR 2 + two tcgen05.commit per row
// VAR 4: VAR 3 + fence.proxy.async + tcgen05 fence per row
// VAR 5: 30 x N=48 on D column 0 (tap-major cost for reference)
__device__ int g_random;
template <int VAR>
__global__ void __launch_bounds__(128, 1) probe(int rows, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[16];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = g_random ? (0x3c003c00u ^ ((i * 2654435761u) >> 9 & 0x03ff83ffu)) : 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bar[i]), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
      constexpr int A_PLANE = 2080, A_STAGE = 6 * 2080, W_PLANE = 144 * 16, W_KX = 6 * W_PLANE;
      long long t0 = clock64();
      for (int r = 0; r < rows; ++r) {
        const uint32_t kk0 = r + 1;
        const uint32_t c0 = (9 - (kk0 % 10)) * 48;            // block of output row r+1 (lowest column of the run)
        const bool wrap = ((kk0 % 10) == 0) || (((kk0 - 1) % 10) == 0);
        const uint32_t col = (VAR == 0 || VAR == 5) ? 0u : (wrap ? 0u : c0);
        const uint64_t ad0 = umma_smem_desc(a0 + (r % 8) * A_STAGE, A_PLANE, 128);
        const uint64_t bd0 = umma_smem_desc(b0, W_PLANE, 128);
        if (VAR == 4) { fence_proxy_async_smem(); tc_fence_after_sync(); }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            const uint64_t ad = ad0 + ((kx * 16 + 2 * ks * A_PLANE) >> 4);
            const uint64_t bd = bd0 + ((kx * W_KX + 2 * ks * W_PLANE) >> 4);
            if (VAR == 5) {
              umma_bf16(tm, ad, bd, idesc_n(48), 1u);
              umma_bf16(tm, ad, bd + 48, idesc_n(48), 1u);
              umma_bf16(tm, ad, bd + 96, idesc_n(48), 1u);
            } else if (VAR >= 2 && kx == 0 && ks == 0) {
              umma_bf16(tm + col, ad, bd, idesc_n(48), 0u);
              umma_bf16(tm + col + 48, ad, bd + 48, idesc_n(96), 1u);
            } else {
              umma_bf16(tm + col, ad, bd, idesc_n(144), 1u);
            }
          }
        }
        if (VAR >= 3) {
          umma_commit(smem_u32(&bar[r & 7]));
          umma_commit(smem_u32(&bar[8 + (r & 7)]));
        }
      }
      umma_commit(smem_u32(&bar[15]));
      long long t1 = clock64();
      uint32_t par = (VAR >= 3) ? (((rows + 7 - 7) / 8) & 1) : 0;   // bar[15] = 8 + 7: count its completions
      if (VAR >= 3) {
        int uses = 0;
        for (int r = 0; r < rows; ++r) uses += ((r & 7) == 7);
        par = uses & 1;
      }
      mbar_wait(smem_u32(&bar[15]), par);
      long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

static int g_grid = 1;
template <int VAR>
void run(long long* d, const char* what) {
  cudaFuncSetAttribute(probe<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rows = 512;
  probe<VAR><<<g_grid, 128, 200 * 1024>>>(rows, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("var %d: %s\n", VAR, cudaGetErrorString(e)); exit(1); }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-3d %-60s issue %8.1f  total %8.1f clk per row\n", VAR, what, double(h[0]) / rows, double(h[1]) / rows);
}

int main(int argc, char** argv) {
  long long* d;
  cudaMalloc(&d, 64);
  g_grid = argc > 1 ? atoi(argv[1]) : 1;
  const int rnd = argc > 2 ? atoi(argv[2]) : 0;
  cudaMemcpyToSymbol(g_random, &rnd, sizeof(int));
  printf("grid %d, %s operand data\n", g_grid, rnd ? "pseudo-random" : "constant");
  run<0>(d, "9 x N=144, fixed D");
  run<1>(d, "9 x N=144, D moves through the 10-block ring");
  run<2>(d, "N=48 overwrite + N=96 + 8 x N=144, ring");
  run<3>(d, "  + 2 commits per row");
  run<4>(d, "  + fence.proxy.async + tcgen05 fence per row");
  run<5>(d, "27 x N=48, fixed D (tap-major cost)");
  return 0;
}
