// Probe 2: (a) tcgen05.ld throughput with 4 warps, (b) cta_group::2 M=256 MMA cost vs N, (c) two issuing threads.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "../../larvanet_b200/csrc/lv_common.cuh"
namespace lv { void set_error(const char*, ...) {} void count_launch(int) {} int sm_count() { return 148; } }
using namespace lv;

// ---------------- (a) LDTM throughput ----------------
__global__ void __launch_bounds__(128, 1) ldtm_probe(int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot + (static_cast<uint32_t>(warp * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      tmem_ld16(tm + ((i * 8 + j) % 24) * 16, v);
      tmem_ld_wait();
      acc += v[0] + v[15];
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc;
  // batched: issue 8 loads then one wait
  __syncthreads();
  t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float v[8][16];
#pragma unroll
    for (int j = 0; j < 8; ++j) tmem_ld16(tm + ((i * 8 + j) % 24) * 16, v[j]);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j][0] + v[j][15];
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = t1 - t0;
  sink[threadIdx.x] = acc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc<512>(slot); }
}

// ---------------- (b) cta_group::2 ----------------
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_probe(int N, int T, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster.block_rank();
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  cluster.sync();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
    const uint32_t idesc = umma_idesc_bf16(256, N, 0, 0);
    for (int rep = 0; rep < 2; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < T; ++i) {
        const uint64_t ad = umma_smem_desc(a0 + (i % 3) * 4096, 2048, 128);
        const uint64_t bd = umma_smem_desc(b0 + (i % 3) * (N / 2) * 32, (N / 2) * 16, 128);
        umma_bf16_2cta(tm, ad, bd, idesc, i > 0 ? 1u : 0u);
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
      mbar_wait(smem_u32(&bar), rep & 1);
      long long t2 = clock64();
      out[rep] = t2 - t0;
    }
  }
  tc_fence_before_sync();
  cluster.sync();
  if (warp == 0) {
    tc_fence_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
  }
}

// ---------------- (c) two issuing threads in one CTA (different warps, different accumulators) ----------------
__global__ void __launch_bounds__(128, 1) mma_two_issuers(int N, int T, int nissue, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = slot;
  if ((threadIdx.x & 31) == 0 && warp < nissue) {
    const uint32_t a0 = smem_u32(smem) + warp * 16384, b0 = smem_u32(smem + 48 * 1024) + warp * 16384;
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    long long t0 = clock64();
    for (int i = 0; i < T; ++i) {
      const uint64_t ad = umma_smem_desc(a0 + (i % 3) * 4096, 2048, 128);
      const uint64_t bd = umma_smem_desc(b0 + (i % 3) * N * 32, N * 16, 128);
      umma_bf16(tm + warp * 256, ad, bd, idesc, i > 0 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar[warp]));
    mbar_wait(smem_u32(&bar[warp]), 0);
    out[warp] = clock64() - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc<512>(tm); }
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 64); cudaMalloc(&sink, 4096);
  long long h[4];
  ldtm_probe<<<1, 128>>>(200, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ldtm: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("LDTM 32x32b.x16 (4 warps, 2 KB per warp-instr): serial %.1f cyc/instr, batched(8) %.1f cyc/instr  => %.1f B/clk/SM batched\n",
         double(h[0]) / 1600, double(h[1]) / 1600, 4 * 2048.0 / (double(h[1]) / 1600));
  cudaFuncSetAttribute(mma2_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(mma_two_issuers, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int T = 216;
  for (int N : {32, 48, 96, 144, 192, 256}) {
    mma2_probe<<<2, 128, 96 * 1024>>>(N, T, d);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mma2 N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("cta_group::2 M=256 N=%d: %.1f cyc/mma\n", N, double(h[1]) / T);
  }
  for (int N : {48, 144})
    for (int ni : {1, 2}) {
      mma_two_issuers<<<1, 128, 96 * 1024>>>(N, T, ni, d);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("two issuers: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("one CTA, %d issuing thread(s), N=%d: %.1f / %.1f cyc per own mma\n", ni, N, double(h[0]) / T, ni > 1 ? double(h[1]) / T : 0.0);
    }
  return 0;
}
