// Which concurrent activity slows the row-marching MMA stream (conv_row.cu) down?  One CTA: warp 1 issues the kernel's MMA
// pattern (N=48 overwrite + N=96 + 8 x N=144 per row, D ring, 2 commits per row); the other warps optionally run one of the
// kernel's side activities in a loop until the MMA warp is done.  Prints clk per row for every combination asked for.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../larvanet_b200/csrc/lv_common.cuh"
namespace lv { void set_error(const char*, ...) {} void count_launch(int) {} int sm_count() { return 148; } }
using namespace lv;

__device__ __forceinline__ uint32_t idesc_n(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}
constexpr int F_CPASYNC = 1, F_LDTM = 2, F_GMEM = 4, F_SPIN = 8, F_STS = 16, F_LDS = 32;

__global__ void __launch_bounds__(384, 1) probe(int rows, int flags, const uint4* __restrict__ gsrc, uint4* __restrict__ gdst,
                                                long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[17];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 17; ++i) mbar_init(smem_u32(&bar[i]), 1); mbar_fence_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  constexpr int A_PLANE = 2080, A_STAGE = 6 * 2080, W_PLANE = 144 * 16, W_KX = 6 * W_PLANE;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
      long long t0 = clock64();
      for (int r = 0; r < rows; ++r) {
        const uint32_t kk0 = r + 1;
        const bool wrap = ((kk0 % 10) == 0) || (((kk0 - 1) % 10) == 0);
        const uint32_t col = wrap ? 0u : (9 - (kk0 % 10)) * 48;
        const uint64_t ad0 = umma_smem_desc(a0 + (r % 8) * A_STAGE, A_PLANE, 128);
        const uint64_t bd0 = umma_smem_desc(b0, W_PLANE, 128);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            const uint64_t ad = ad0 + ((kx * 16 + 2 * ks * A_PLANE) >> 4);
            const uint64_t bd = bd0 + ((kx * W_KX + 2 * ks * W_PLANE) >> 4);
            if (kx == 0 && ks == 0) {
              umma_bf16(tm + col, ad, bd, idesc_n(48), 0u);
              umma_bf16(tm + col + 48, ad, bd + 48, idesc_n(96), 1u);
            } else {
              umma_bf16(tm + col, ad, bd, idesc_n(144), 1u);
            }
          }
        }
        umma_commit(smem_u32(&bar[r & 7]));
        umma_commit(smem_u32(&bar[8 + (r & 7)]));
      }
      umma_commit(smem_u32(&bar[16]));
      mbar_wait(smem_u32(&bar[16]), 0);
      out[0] = clock64() - t0;
      stop = 1;
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 4 && (flags & F_CPASYNC)) {
    // two producer warps: 13 x 16 B cp.async per thread per "row" into the A ring, like conv_row.cu's producers
    const int ptid = threadIdx.x - 64;
    int r = 0;
    while (!stop) {
      const uint32_t dst0 = smem_u32(smem + (r % 8) * A_STAGE) + ptid * 16;
#pragma unroll
      for (int i = 0; i < 13; ++i)
        if (ptid + i * 64 < 780) cp_async16(dst0 + i * 1024, gsrc + ((r * 780 + ptid + i * 64) & 0xfffff), 16u);
      cp_async_commit();
      cp_async_wait<4>();
      ++r;
    }
    cp_async_wait<0>();
  } else if (warp >= 4 && (flags & F_LDTM)) {
    // eight epilogue-like warps reading TMEM: 3 x (32 lanes x 16 columns) per iteration
    const uint32_t ta = tm + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    float acc = 0.f;
    while (!stop) {
      float v[48];
      tmem_ld16(ta, v); tmem_ld16(ta + 16, v + 16); tmem_ld16(ta + 32, v + 32);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 48; ++i) acc += v[i];
      __nanosleep(200);
    }
    if (acc == 123.f) out[7] = 1;
  } else if (warp >= 4 && (flags & F_GMEM)) {
    // eight epilogue-like warps: 6 x 16 B global loads (L2 only) + 6 x 16 B global stores per thread per iteration
    const int et = threadIdx.x - 128;
    int r = 0;
    while (!stop) {
      uint4 q[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) q[j] = __ldcg(gsrc + (((r * 6 + j) * 256 + et) & 0xfffff));
#pragma unroll
      for (int j = 0; j < 6; ++j) { q[j].x ^= r; gdst[(((r * 6 + j) * 256 + et) & 0xfffff)] = q[j]; }
      ++r;
    }
  } else if (warp >= 4 && (flags & F_STS)) {
    // eight warps storing to / loading from an unused shared-memory area (what register spills would do to the L1 arrays)
    uint32_t sa = smem_u32(smem + 180 * 1024) + (threadIdx.x - 128) * 16;
    int r = 0;
    while (!stop) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sa), "r"(r) : "memory");
      ++r;
      __nanosleep(20);
    }
  } else if (warp >= 4 && (flags & F_SPIN)) {
    // eight warps polling an mbarrier without back-off (what a mbar_wait() spin does)
    uint32_t spins = 0;
    while (!stop) { (void)mbar_test_wait(smem_u32(&bar[15]), 0); ++spins; }
    if (spins == 0xffffffffu) out[6] = 1;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

int main() {
  long long* d;
  uint4 *gs, *gd;
  cudaMalloc(&d, 64);
  cudaMalloc(&gs, 16 << 20);
  cudaMalloc(&gd, 16 << 20);
  cudaMemset(gs, 1, 16 << 20);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rows = 256;
  const int combos[] = {0, F_CPASYNC, F_LDTM, F_GMEM, F_STS, F_SPIN, F_CPASYNC | F_LDTM, F_CPASYNC | F_GMEM};
  const char* names[] = {"MMA stream alone", "+ 2 warps cp.async 12.5 KB/row", "+ 8 warps tcgen05.ld", "+ 8 warps global ld.cg/st 16 B",
                         "+ 8 warps st.shared", "+ 8 warps mbarrier polling", "+ cp.async + tcgen05.ld", "+ cp.async + global ld/st"};
  for (int c = 0; c < 8; ++c) {
    probe<<<1, 384, 200 * 1024>>>(rows, combos[c], gs, gd, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", names[c], cudaGetErrorString(e)); return 1; }
    long long h;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %8.1f clk per row\n", names[c], double(h) / rows);
  }
  return 0;
}
