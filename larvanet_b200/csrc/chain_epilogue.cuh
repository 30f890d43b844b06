// Tile epilogues shared by the persistent chain kernels (conv_chain.cu: data flow through global flags; conv_strip.cu:
// cluster-resident strips with activations in shared memory).
#pragma once
#include "conv_epilogue.cuh"
#include "lv_common.cuh"

namespace lv {
namespace chain {

__device__ __forceinline__ uint4 ldcg16(const __nv_bfloat16* p) {
  return __ldcg(reinterpret_cast<const uint4*>(p));   // L2 only: other CTAs rewrite these buffers during the kernel
}

// Fast planar epilogue of one tile for one thread (one pixel x NT channels), straight-line for a compile-time flag
// set EPI (bit0 ReLU, bit1 ReLU-mask, bit2 res1, bit3 res2; EPI < 0: flags read at run time).  Operand loads are
// issued before the accumulator wait so that their latency hides behind the MMAs.
struct FastEpi {
  const __nv_bfloat16* mask;
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
  __nv_bfloat16* out;
  float res_scale;
  int relu;
};

// Optional second destination of a tile's bf16 results: the next layer's A operand in shared memory (own CTA) and the
// halo column of the left / right neighbour CTA of the cluster (mapa'd shared::cluster addresses; 0 = none).
struct SmemOut {
  uint32_t own, left, right;   // address of this pixel's chunk 0
  uint32_t plane;              // bytes between 8-channel chunk planes
};
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& t) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, const uint4& t) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
}

template <int EPI, int NT, bool BIAS_REGS, bool SMEM_OUT = false>
__device__ __forceinline__ void fast_tile(const FastEpi& e, const float* breg, const float* bias_g, bool valid, size_t o0,
                                          size_t chunk_stride,
                                          uint32_t taddr, uint32_t tfull, uint32_t tempty, uint32_t parity,
                                          const SmemOut so = SmemOut{0, 0, 0, 0}, uint4* deferred = nullptr) {
  constexpr int NCH = NT / 8;
  const bool unit_scale = (EPI >= 0) || (e.res_scale == 1.0f);
  const bool do_relu = (EPI >= 0) ? ((EPI & 1) != 0) : (e.relu != 0);
  const bool do_mask = (EPI >= 0) ? ((EPI & 2) != 0) : (e.mask != nullptr);
  const bool do_res1 = (EPI >= 0) ? ((EPI & 4) != 0) : (e.res1 != nullptr);
  const bool do_res2 = (EPI >= 0) ? ((EPI & 8) != 0) : (e.res2 != nullptr);
  uint4 qm[NCH], q1[NCH], q2[NCH];
  if (valid) {
    if (do_mask) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) qm[j] = ldcg16(e.mask + o0 + j * chunk_stride);
    }
    if (do_res1) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) q1[j] = ldcg16(e.res1 + o0 + j * chunk_stride);
    }
    if (do_res2) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) q2[j] = ldcg16(e.res2 + o0 + j * chunk_stride);
    }
  }
  mbar_wait_relaxed(tfull, parity);
  tc_fence_after_sync();
  float v[NT];
#pragma unroll
  for (int j = 0; j < NT / 16; ++j) tmem_ld16(taddr + j * 16, v + j * 16);
  tmem_ld_wait();
  tc_fence_before_sync();
  mbar_arrive(tempty);   // accumulator stage free: this stage's next tile may be accumulated
  if (valid) {
    __nv_bfloat16* po = e.out + o0;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      float* vj = v + 8 * j;
      if constexpr (BIAS_REGS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vj[i] += breg[8 * j + i];
      } else if (bias_g != nullptr) {   // L1-resident after the first tile
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_g) + 2 * j);
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_g) + 2 * j + 1);
        vj[0] += b0.x; vj[1] += b0.y; vj[2] += b0.z; vj[3] += b0.w;
        vj[4] += b1.x; vj[5] += b1.y; vj[6] += b1.z; vj[7] += b1.w;
      }
      if (!unit_scale) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vj[i] *= e.res_scale;
      }
      if (do_relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vj[i] = fmaxf(vj[i], 0.f);
      }
      if (do_mask) {
        const uint32_t w4[4] = {qm[j].x, qm[j].y, qm[j].z, qm[j].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          vj[2 * t] = (bf16_lo(w4[t]) > 0.f) ? vj[2 * t] : 0.f;
          vj[2 * t + 1] = (bf16_hi(w4[t]) > 0.f) ? vj[2 * t + 1] : 0.f;
        }
      }
      if (do_res1) {
        const uint32_t w4[4] = {q1[j].x, q1[j].y, q1[j].z, q1[j].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { vj[2 * t] += bf16_lo(w4[t]); vj[2 * t + 1] += bf16_hi(w4[t]); }
      }
      if (do_res2) {
        const uint32_t w4[4] = {q2[j].x, q2[j].y, q2[j].z, q2[j].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { vj[2 * t] += bf16_lo(w4[t]); vj[2 * t + 1] += bf16_hi(w4[t]); }
      }
      if constexpr (SMEM_OUT) {
        uint4 t;
        t.x = pack_bf16x2(vj[0], vj[1]); t.y = pack_bf16x2(vj[2], vj[3]);
        t.z = pack_bf16x2(vj[4], vj[5]); t.w = pack_bf16x2(vj[6], vj[7]);
        // `deferred`: the caller stores to global memory itself, AFTER it has signalled the neighbours (a
        // release at cluster scope waits for every earlier global store of the thread to be acknowledged)
        if (deferred != nullptr) deferred[j] = t;
        else if (e.out != nullptr) *reinterpret_cast<uint4*>(po + j * chunk_stride) = t;
        st_shared_v4(so.own + j * so.plane, t);
        if (so.left != 0) st_cluster_v4(so.left + j * so.plane, t);
        if (so.right != 0) st_cluster_v4(so.right + j * so.plane, t);
      } else {
        store8(po + j * chunk_stride, vj);
      }
    }
  }
}

// PixelShuffle(4) + bicubic base (+ L1 loss / sign gradient) epilogue of one tile for one thread, same arithmetic as
// conv_epilogue16's LV_EPI_PS4_NCHW branch but software-pipelined: the base / truth lines of colour plane c+1 are in
// flight while plane c is computed, and the first plane's are issued before the accumulator wait.
template <int NT>
__device__ __forceinline__ float ps4_tile(const lv_conv_args& a, const float* bias_g, bool valid, int n, int y, int x, int H,
                                          int W, size_t o0, size_t chunk_stride, uint32_t taddr, uint32_t tfull,
                                          uint32_t tempty, uint32_t parity) {
  static_assert(NT == 48, "three colour planes of 16 sub-pixels");
  const size_t W4 = static_cast<size_t>(W) * 4;
  const size_t plane = static_cast<size_t>(H) * 4 * W4;
  const size_t hr0 = (static_cast<size_t>(n) * 3 * (static_cast<size_t>(H) * 4) + 4 * y) * W4 + 4 * x;   // colour 0, dy 0
  const bool has_base = a.base_hr != nullptr, has_truth = a.truth_hr != nullptr, has_out = a.out_hr != nullptr;
  const bool has_sign = has_truth && a.grad_sign != nullptr;
  float4 qb[4], qt[4];
  auto load_plane = [&](int c, float4* b4, float4* t4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t off = hr0 + c * plane + i * W4;
      b4[i] = (valid && has_base) ? __ldg(reinterpret_cast<const float4*>(a.base_hr + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
      t4[i] = (valid && has_truth) ? __ldg(reinterpret_cast<const float4*>(a.truth_hr + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_plane(0, qb, qt);
  mbar_wait_relaxed(tfull, parity);
  tc_fence_after_sync();
  float loss = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[16];
    tmem_ld16(taddr + c * 16, v);
    float4 nb[4], nt[4];
    if (c < 2) load_plane(c + 1, nb, nt);
    tmem_ld_wait();
    if (c == 2) {
      tc_fence_before_sync();
      mbar_arrive(tempty);
    }
    if (valid) {
      float gs[16];
      float closs = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 bb = (bias_g != nullptr) ? __ldg(reinterpret_cast<const float4*>(bias_g) + c * 4 + i)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o = make_float4(v[4 * i] + bb.x, v[4 * i + 1] + bb.y, v[4 * i + 2] + bb.z, v[4 * i + 3] + bb.w);
        if (has_base) { o.x += qb[i].x; o.y += qb[i].y; o.z += qb[i].z; o.w += qb[i].w; }
        if (has_out) *reinterpret_cast<float4*>(a.out_hr + hr0 + c * plane + i * W4) = o;
        if (a.out_u8 != nullptr) *reinterpret_cast<uint32_t*>(a.out_u8 + hr0 + c * plane + i * W4) = pack_u8x4(o);
        if (has_truth) {
          const float d0 = o.x - qt[i].x, d1 = o.y - qt[i].y, d2 = o.z - qt[i].z, d3 = o.w - qt[i].w;
          closs += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
          gs[4 * i + 0] = (d0 > 0.f) ? 1.f : ((d0 < 0.f) ? -1.f : 0.f);
          gs[4 * i + 1] = (d1 > 0.f) ? 1.f : ((d1 < 0.f) ? -1.f : 0.f);
          gs[4 * i + 2] = (d2 > 0.f) ? 1.f : ((d2 < 0.f) ? -1.f : 0.f);
          gs[4 * i + 3] = (d3 > 0.f) ? 1.f : ((d3 < 0.f) ? -1.f : 0.f);
        }
      }
      loss += closs;
      if (has_sign) {
        __nv_bfloat16* gp = reinterpret_cast<__nv_bfloat16*>(a.grad_sign) + o0 + (2 * c) * chunk_stride;
        store8(gp, gs);
        store8(gp + chunk_stride, gs + 8);
      }
    }
    if (c < 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { qb[i] = nb[i]; qt[i] = nt[i]; }
    }
  }
  return loss;
}


}  // namespace chain
}  // namespace lv
