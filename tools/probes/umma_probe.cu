// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, of how many independent TMEM
// accumulators the issue loop rotates over, and of the shared-memory operand layout.  Timing only -- operands are
// whatever is in shared memory.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../larvanet_b200/csrc/lv_common.cuh"

namespace lv { void set_error(const char*, ...) {} void count_launch(int) {} int sm_count() { return 148; } }
using namespace lv;

__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = umma_smem_desc(saddr, lbo, sbo);
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

// mode 0: conv-like no-swizzle (A: LBO 2880, SBO 160; B: LBO N*16, SBO 128)
// mode 1: dense no-swizzle     (A: LBO 2048, SBO 128; B: LBO N*16, SBO 128)
// mode 2: SWIZZLE_128B K-major (A,B: SBO 1024, layout 2), K advance = +32 B
// mode 3: SWIZZLE_32B  K-major (A,B: SBO 256, layout 6)
__global__ void __launch_bounds__(128, 1) probe(int N, int R, int mode, int T, int m64, int stride, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ uint64_t sad[256], sbd[256];
  __shared__ uint32_t sacc[256];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = umma_idesc_bf16(m64 ? 64 : 128, N, 0, 0);
    for (int i = 0; i < T; ++i) {
        const int ks = i % 3, tap = (i / 3) % 9;
        uint64_t ad, bd;
        if (mode == 0) {
          ad = desc_sw(a0 + ((tap / 3) * 10 + tap % 3) * 16 + 2 * ks * 2880, 2880, 160, 0);
          bd = desc_sw(b0 + (tap % 3) * 6 * N * 16 + 2 * ks * N * 16, N * 16, 128, 0);
        } else if (mode == 1) {
          ad = desc_sw(a0 + 2 * ks * 2048 + (tap % 3) * 16384, 2048, 128, 0);
          bd = desc_sw(b0 + (tap % 3) * 6 * N * 16 + 2 * ks * N * 16, N * 16, 128, 0);
        } else if (mode == 2) {
          ad = desc_sw(a0 + ks * 32 + (tap % 3) * 16384, 16, 1024, 2);
          bd = desc_sw(b0 + ks * 32 + (tap % 3) * 32768, 16, 1024, 2);
        } else {
          ad = desc_sw(a0 + (tap % 3) * 4096 + ks * 4096 * 3, 16, 256, 6);
          bd = desc_sw(b0 + (tap % 3) * 8192 + ks * 8192 * 3, 16, 256, 6);
        }
        sad[i] = ad; sbd[i] = bd; sacc[i] = tm + (i % R) * stride;
    }
    for (int rep = 0; rep < 2; ++rep) {
      long long t0 = clock64();
#pragma unroll 8
      for (int i = 0; i < T; ++i) umma_bf16(sacc[i], sad[i], sbd[i], idesc, i >= R ? 1u : 0u);
      umma_commit(smem_u32(&bar));
      long long t1 = clock64();
      mbar_wait(smem_u32(&bar), rep & 1);
      long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int T = 216;
  printf("cycles per MMA (M=128,K=16 bf16), T=%d MMAs; issue = cycles until the issuing thread is done\n", T);
  printf("%-6s %-4s %-3s %-5s %10s %10s\n", "mode", "N", "R", "M", "issue/mma", "total/mma");
  for (int m64 = 0; m64 < 2; ++m64)
    for (int mode = 0; mode < 4; ++mode)
      for (int N : {16, 48, 96, 128, 256})
        for (int R : {1, 2, 4}) {
          const int stride = (N + 63) / 64 * 64;
          if (R * stride > 512) continue;
          if (m64 && (mode == 2 || mode == 3)) continue;
          probe<<<1, 128, 200 * 1024>>>(N, R, mode, T, m64, stride, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("mode %d N %d R %d: %s\n", mode, N, R, cudaGetErrorString(e)); return 1; }
          long long h[4];
          cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
          printf("%-6d %-4d %-3d %-5d %10.1f %10.1f\n", mode, N, R, m64 ? 64 : 128, double(h[2]) / T, double(h[3]) / T);
        }
  return 0;
}
