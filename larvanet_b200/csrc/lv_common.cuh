// Shared helpers: error plumbing, launch counting and the sm_100a PTX wrappers (mbarrier, cp.async,
// cp.async.bulk, tcgen05.{alloc,mma,commit,ld,fence}) used by the tensor-core kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/larvanet_b200.h"

namespace lv {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int sm_count();

// Per-DEVICE one-time configuration (cudaFuncSetAttribute applies to the current device only): returns the slot of the
// current device in a kernel-specific `state[64]` table.
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  return dev;
}

#define LV_CHECK_ARG(cond, ...)                   \
  do {                                            \
    if (!(cond)) {                                \
      lv::set_error(__VA_ARGS__);                 \
      return LV_ERR_INVALID;                      \
    }                                             \
  } while (0)

#define LV_CUDA_OK(expr)                                                            \
  do {                                                                              \
    cudaError_t e_ = (expr);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      lv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return LV_ERR_CUDA;                                                           \
    }                                                                               \
  } while (0)

#define LV_LAUNCH_OK()                                                              \
  do {                                                                              \
    cudaError_t e_ = cudaGetLastError();                                            \
    if (e_ != cudaSuccess) {                                                        \
      lv::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return LV_ERR_CUDA;                                                           \
    }                                                                               \
    lv::count_launch();                                                             \
  } while (0)

// ------------------------------------------------------------------------------------------------
// dtype helpers
// ------------------------------------------------------------------------------------------------
template <typename T> struct DType;
template <> struct DType<float> { static constexpr int id = LV_F32; };
template <> struct DType<__nv_bfloat16> { static constexpr int id = LV_BF16; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// load / store 16 consecutive channels of one pixel as fp32 registers
__device__ __forceinline__ void load16(const float* p, float* v) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4 t = q[i];
    v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void load16(const __nv_bfloat16* p, float* v) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 t = q[i];
    v[8 * i + 0] = bf16_lo(t.x); v[8 * i + 1] = bf16_hi(t.x);
    v[8 * i + 2] = bf16_lo(t.y); v[8 * i + 3] = bf16_hi(t.y);
    v[8 * i + 4] = bf16_lo(t.z); v[8 * i + 5] = bf16_hi(t.z);
    v[8 * i + 6] = bf16_lo(t.w); v[8 * i + 7] = bf16_hi(t.w);
  }
}
__device__ __forceinline__ void store16(float* p, const float* v) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void store16(__nv_bfloat16* p, const float* v) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint4 t;
    t.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    t.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    t.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    t.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    q[i] = t;
  }
}

// 8 consecutive bf16 channels (16 B)
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* v) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
  v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* v) {
  uint4 t;
  t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
  t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = t;
}

// ------------------------------------------------------------------------------------------------
// Activation layout ("planar-8"): [N][H][C/8][W][8] -- for every image row, each 8-channel chunk is a contiguous run
// of W pixels x 16 B (bf16).  Why not NHWC: (1) a halo-tile row of one chunk is one contiguous 160 B run, so a tile is
// 108 TMA box rows instead of 1080 16-byte gathers and lands in shared memory already in UMMA operand order;
// (2) a warp of the epilogue (4 tile rows x 8 pixels, one pixel per thread) reads/writes four full 128 B lines per
// 8-channel chunk straight from registers -- no shared-memory staging, which matters because shared-memory bandwidth
// is what bounds the conv kernel.  Element offset of channel chunk `ch8` of pixel (n,y,x):
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t act_off(int n, int y, int x, int ch8, int H, int W, int CH) {
  return (((static_cast<size_t>(n) * H + y) * CH + ch8) * W + x) * 8;
}
// 16 consecutive channels (two chunks) of one pixel
template <typename T>
__device__ __forceinline__ void load16_act(const T* base, int n, int y, int x, int co0, int H, int W, int C, float* v) {
  const size_t o = act_off(n, y, x, co0 >> 3, H, W, C >> 3);
  if constexpr (sizeof(T) == 2) {
    load8(base + o, v);
    load8(base + o + static_cast<size_t>(W) * 8, v + 8);
  } else {
    const float4* p0 = reinterpret_cast<const float4*>(base + o);
    const float4* p1 = reinterpret_cast<const float4*>(base + o + static_cast<size_t>(W) * 8);
    const float4 a = p0[0], b = p0[1], c = p1[0], d = p1[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z; v[15] = d.w;
  }
}
template <typename T>
__device__ __forceinline__ void store16_act(T* base, int n, int y, int x, int co0, int H, int W, int C, const float* v) {
  const size_t o = act_off(n, y, x, co0 >> 3, H, W, C >> 3);
  if constexpr (sizeof(T) == 2) {
    store8(base + o, v);
    store8(base + o + static_cast<size_t>(W) * 8, v + 8);
  } else {
    float4* p0 = reinterpret_cast<float4*>(base + o);
    float4* p1 = reinterpret_cast<float4*>(base + o + static_cast<size_t>(W) * 8);
    p0[0] = make_float4(v[0], v[1], v[2], v[3]); p0[1] = make_float4(v[4], v[5], v[6], v[7]);
    p1[0] = make_float4(v[8], v[9], v[10], v[11]); p1[1] = make_float4(v[12], v[13], v[14], v[15]);
  }
}

// clip(round_half_even(v), 0, 255) of four fp32 samples packed into one 32-bit word (little endian: x first), the same
// arithmetic as lv_image_to_uint8 / validate._image_to_uint8 (np.round rounds half to even)
__device__ __forceinline__ uint32_t pack_u8x4(float4 v) {
  const int a = min(max(__float2int_rn(v.x), 0), 255), b = min(max(__float2int_rn(v.y), 0), 255);
  const int c = min(max(__float2int_rn(v.z), 0), 255), d = min(max(__float2int_rn(v.w), 0), 255);
  return static_cast<uint32_t>(a) | (static_cast<uint32_t>(b) << 8) | (static_cast<uint32_t>(c) << 16) |
         (static_cast<uint32_t>(d) << 24);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must surface as a trapped kernel (CUDA error at the next sync), never as a hung GPU.
#ifndef LV_SPIN_LIMIT
#define LV_SPIN_LIMIT (1u << 22)
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
  }
}
// Relaxed wait for roles that are normally far ahead of the barrier (producers on `empty`, epilogue on `tmem_full`):
// the suspend-time hint lets the hardware park the thread instead of re-issuing try_wait, so spinning warps do not eat
// the issue slots of the warps that are doing the work (measured: ~35 % of all issued instructions were spin loops).
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
#ifndef LV_RELAXED_SLEEP_NS
#define LV_RELAXED_SLEEP_NS 128
#endif
#ifndef LV_RELAXED_HINT_NS
#define LV_RELAXED_HINT_NS 400
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, LV_RELAXED_HINT_NS)) {
    if (LV_RELAXED_SLEEP_NS > 0) __nanosleep(LV_RELAXED_SLEEP_NS);   // back off: eight spinning epilogue warps otherwise issue a third of the SM's instructions
    if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
  }
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- proxy / tcgen05 fences ----
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- cp.async (LDGSTS), 16 B, zero-fill when src_bytes == 0 ----
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- TMA bulk copy global -> shared (1-D, no tensor map), completes on an mbarrier ----
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- TMA bulk copy shared -> global (1-D), bulk-group completion ----
__device__ __forceinline__ void tma_bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the sources of all but the N most recent bulk groups have been read (smem reusable)
template <int N> __device__ __forceinline__ void tma_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// named barrier among a subset of the CTA's warps (id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- programmatic dependent launch (PDL) ----
// wait: blocks until every grid this one programmatically depends on has completed and flushed (no-op otherwise);
// launch_dependents: lets the next PDL-launched grid start its prologue while this one is still running.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these) ----
template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// ---- UMMA descriptors -------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved" 8x16B core matrices), version 1 (sm_100).
//   K-major operand : rows of a core matrix are 16 B apart; SBO = byte distance between 8-row groups (M/N),
//                     LBO = byte distance between the two 16 B K-halves of one K=16 step.
//   MN-major operand: 8 K-rows of a core matrix are 16 B apart; SBO = distance between 8-element MN chunks,
//                     LBO = distance between 8-row K groups.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  return d;                             // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4)                                  // c_format = F32
         | (1u << 7)                                // a_format = BF16
         | (1u << 10)                               // b_format = BF16
         | (static_cast<uint32_t>(a_mn_major) << 15)
         | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(n >> 3) << 17)
         | (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// make the mbarrier track completion of all tcgen05.mma issued so far by this thread
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
        "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
      : "r"(taddr)
      : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace lv
