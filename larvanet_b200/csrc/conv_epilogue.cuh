// Fused conv epilogue shared by the tcgen05 kernel (conv_tc.cu) and the CUDA-core kernel (conv_simt.cu).
// Unit of work: one output pixel x 16 consecutive output channels held as fp32 accumulators.
//
// Restates, per element, what the reference spreads over separate torch ops:
//   bias add (nn.Conv2d), nn.ReLU, torch.add(x,res) / x+fea (models/LarvaNet.py:217-220,246-248),
//   nn.PixelShuffle(4) + `out += base` (:263-267), nn.L1Loss (:85,108) and its sign gradient,
//   nn.PixelShuffle(2) (models/edsr.py:164), final_conv + mean_inverse_shift 1x1 (models/edsr.py:204-205).
#pragma once

#include "lv_common.cuh"

namespace lv {

// geometry the kernels precompute once
struct ConvGeom {
  int tiles_x, tiles_y, ntiles_n;  // pixel tiles per image in x / y, N tiles
  int nt;                          // output channels per N tile
  int cout_pad;                    // cout rounded up to 16
  int total_tiles;
  long long* timeline;             // debug: per-role clock64 stamps of CTA 0 (nullptr in production)
};

// debug timeline record: [role][slot] = clock64; role 0 producer, 1 mma, 2 epilogue; 4 stamps per tile
__device__ __forceinline__ void tl_stamp(const ConvGeom& g, int role, uint32_t tile_k, int ev) {
  if (g.timeline != nullptr && blockIdx.x == 0 && tile_k < 64) g.timeline[(role * 64 + tile_k) * 4 + ev] = clock64();
}

template <typename T, bool ADD_BIAS = true>
__device__ __forceinline__ float conv_epilogue16(const lv_conv_args& a, int n, int y, int x, int co0, float* v) {
  const int H = a.h, W = a.w, C = a.cout;
  // 1. bias + scale
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int co = co0 + i;
    float b = 0.f;
    if (ADD_BIAS && a.bias != nullptr && co < C) b = __ldg(a.bias + co);
    v[i] = (co < C) ? a.res_scale * (v[i] + b) : 0.f;
  }
  // 2. ReLU
  if (a.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  // 3. ReLU mask of the forward activation (backward-data through nn.ReLU)
  if (a.mask != nullptr) {
    float t[16];
    load16_act(reinterpret_cast<const T*>(a.mask), n, y, x, co0, H, W, C, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (t[i] > 0.f) ? v[i] : 0.f;
  }
  // 4. residual / skip adds
  if (a.res1 != nullptr) {
    float t[16];
    load16_act(reinterpret_cast<const T*>(a.res1), n, y, x, co0, H, W, C, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += t[i];
  }
  if (a.res2 != nullptr) {
    float t[16];
    load16_act(reinterpret_cast<const T*>(a.res2), n, y, x, co0, H, W, C, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += t[i];
  }
  float loss = 0.f;
  // 5. store
  if (a.epilogue == LV_EPI_NHWC) {
    store16_act(reinterpret_cast<T*>(a.out), n, y, x, co0, H, W, C, v);
  } else if (a.epilogue == LV_EPI_PS4_NCHW) {
    const int c = co0 >> 4;
    const int CH = C >> 4;
    const size_t W4 = static_cast<size_t>(W) * 4;
    float g[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t off = ((static_cast<size_t>(n) * CH + c) * (static_cast<size_t>(H) * 4) + (4 * y + i)) * W4 + 4 * x;
      float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if (a.base_hr != nullptr) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.base_hr + off));
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      if (a.out_hr != nullptr) *reinterpret_cast<float4*>(a.out_hr + off) = o;
      if (a.out_u8 != nullptr) *reinterpret_cast<uint32_t*>(a.out_u8 + off) = pack_u8x4(o);
      if (a.truth_hr != nullptr) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a.truth_hr + off));
        const float d0 = o.x - t.x, d1 = o.y - t.y, d2 = o.z - t.z, d3 = o.w - t.w;
        loss += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
        g[4 * i + 0] = (d0 > 0.f) ? 1.f : ((d0 < 0.f) ? -1.f : 0.f);
        g[4 * i + 1] = (d1 > 0.f) ? 1.f : ((d1 < 0.f) ? -1.f : 0.f);
        g[4 * i + 2] = (d2 > 0.f) ? 1.f : ((d2 < 0.f) ? -1.f : 0.f);
        g[4 * i + 3] = (d3 > 0.f) ? 1.f : ((d3 < 0.f) ? -1.f : 0.f);
      }
    }
    if (a.truth_hr != nullptr && a.grad_sign != nullptr) {
      store16_act(reinterpret_cast<T*>(a.grad_sign), n, y, x, co0, H, W, C, g);
    }
  } else if (a.epilogue == LV_EPI_PS2_NHWC) {
    const int CO = C >> 2;        // output channels after the shuffle
    const int c0 = co0 >> 2;      // 4 consecutive output channels per 16-channel input chunk (half of an 8-chunk)
    T* out = reinterpret_cast<T*>(a.out);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        T* p = out + act_off(n, 2 * y + i, 2 * x + j, c0 >> 3, 2 * H, 2 * W, CO >> 3) + (c0 & 7);
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = from_f32<T>(v[4 * k + 2 * i + j]);
      }
    }
  } else {  // LV_EPI_RGB_NCHW
    if (co0 == 0) {
      float o[3] = {v[0], v[1], v[2]};
      if (a.post_w != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          o[c] = __ldg(a.post_w + 3 * c) * v[0] + __ldg(a.post_w + 3 * c + 1) * v[1] + __ldg(a.post_w + 3 * c + 2) * v[2];
          if (a.post_b != nullptr) o[c] += __ldg(a.post_b + c);
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) a.out_hr[((static_cast<size_t>(n) * 3 + c) * H + y) * W + x] = o[c];
    }
  }
  return loss;
}

// PixelShuffle(2) epilogue for 32 consecutive conv channels of one pixel (bf16 output): they are the 8 consecutive OUTPUT
// channels co0/4 .. co0/4+7 of the four sub-pixels, i.e. one full 16-byte chunk per sub-pixel -- four vector stores
// instead of the 64 two-byte stores of the generic 16-channel routine (models/edsr.py:156-173, nn.PixelShuffle(2)).
// Bias, res_scale and ReLU as in conv_epilogue16; no mask / residual operands (EDSR's UpsampleBlock has none).
__device__ __forceinline__ void conv_epilogue_ps2_32(const lv_conv_args& a, int n, int y, int x, int co0, float* v) {
  const int H = a.h, W = a.w, CO = a.cout >> 2;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float b = (a.bias != nullptr) ? __ldg(a.bias + co0 + i) : 0.f;
    v[i] = a.res_scale * (v[i] + b);
    if (a.relu) v[i] = fmaxf(v[i], 0.f);
  }
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = v[4 * k + 2 * i + j];
      store8(out + act_off(n, 2 * y + i, 2 * x + j, co0 >> 5, 2 * H, 2 * W, CO >> 3), o);
    }
  }
}

}  // namespace lv
