// Micro-benchmark: true tensor-pipe cost of tcgen05.mma (M=128, K=16, bf16, SS) vs N and vs the shared-memory layout /
// alignment of the A operand.  The issue loop mirrors conv_tc.cu (27 compile-time (tap,kstep) descriptors, elect.sync
// leader) so the issuing thread is not the bottleneck.  Timing only -- operands are whatever is in shared memory.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../larvanet_b200/csrc/lv_common.cuh"
namespace lv { void set_error(const char*, ...) {} void count_launch(int) {} int sm_count() { return 148; } }
using namespace lv;

__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off = 0) {
  uint64_t d = umma_smem_desc(saddr, lbo, sbo);
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

// MODE 0: conv_tc today  : no swizzle, halo pitch 10 px (SBO 160), tap shift (ky*10+kx)*16, plane 2880
// MODE 1: dense aligned  : no swizzle, SBO 128, no tap shift (all core matrices 128 B aligned)
// MODE 2: pitch 16       : no swizzle, halo pitch 16 px (SBO 256), tap shift (ky*16+kx)*16, plane 4608
// MODE 3: SW128 shifted  : 128B swizzle, pixel pitch 128 B, halo pitch 16 px (SBO 2048), tap shift (ky*16+kx)*128
// MODE 4: pitch 16, kx=0 : like 2 but only aligned taps (kx forced 0)
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) probe(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
      constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      long long t0 = clock64();
      for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            uint64_t ad;
            const int ky = tap / 3, kx = tap % 3;
            if (MODE == 0) ad = desc_sw(a0 + (ky * 10 + kx) * 16 + 2 * ks * 2880, 2880, 160, 0);
            else if (MODE == 1) ad = desc_sw(a0 + 2 * ks * 2048 + tap * 8192, 2048, 128, 0);
            else if (MODE == 2) ad = desc_sw(a0 + (ky * 16 + kx) * 16 + 2 * ks * 4608, 4608, 256, 0);
            else if (MODE == 3) ad = desc_sw(a0 + (ky * 16 + kx) * 128 + ks * 32, 16, 2048, 2, kx);
            else ad = desc_sw(a0 + (ky * 16) * 16 + 2 * ks * 4608, 4608, 256, 0);
            const uint64_t bd = desc_sw(b0 + (tap % 3) * 6 * N * 16 + 2 * ks * N * 16, N * 16, 128, 0);
            umma_bf16(tm + (rep & 1) * 256, ad, bd, idesc, (tap | ks) ? 1u : 0u);
          }
        }
      }
      umma_commit(smem_u32(&bar));
      long long t1 = clock64();
      mbar_wait(smem_u32(&bar), 0);
      long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

template <int MODE, int N>
void run(long long* d) {
  cudaFuncSetAttribute(probe<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 16;
  probe<MODE, N><<<1, 128, 200 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("mode %d N %d: %s\n", MODE, N, cudaGetErrorString(e)); exit(1); }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-6d %-4d %10.1f %10.1f\n", MODE, N, double(h[0]) / (27 * reps), double(h[1]) / (27 * reps));
}

template <int MODE>
void run_mode(long long* d) {
  run<MODE, 16>(d); run<MODE, 48>(d); run<MODE, 64>(d); run<MODE, 96>(d); run<MODE, 128>(d); run<MODE, 144>(d);
  run<MODE, 192>(d); run<MODE, 256>(d);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  printf("cycles per MMA (M=128,K=16,bf16,SS), 432 MMAs; issue = until the issuing thread finished issuing\n");
  printf("%-6s %-4s %10s %10s\n", "mode", "N", "issue/mma", "total/mma");
  run_mode<0>(d); run_mode<1>(d); run_mode<2>(d); run_mode<3>(d); run_mode<4>(d);
  return 0;
}
