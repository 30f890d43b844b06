// CUDA-core direct 3x3 convolution with the same fused epilogues as the tensor-core kernel.
//   T = float          : the fp32 validation mode (north star: <= 1e-4 relative to the fp32 reference).
//   T = __nv_bfloat16  : on-GPU cross-check of conv_tc.cu at sizes the CPU oracle cannot reach (reads the very
//                        same packed bf16 weights and bf16 activations, accumulates in fp32).  Tests only.
// Tile: 8 x 16 output pixels per 256-thread block; thread = (pixel, parity of the 16-channel output chunk).
#include "conv_epilogue.cuh"
#include "lv_common.cuh"

namespace lv {

int pick_ntile(int cout_pad);

constexpr int kSTileH = 8, kSTileW = 16;
constexpr int kSHaloW = kSTileW + 2, kSHaloH = kSTileH + 2, kSHaloPix = kSHaloW * kSHaloH;

template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(const __grid_constant__ lv_conv_args a, int tiles_x, int tiles_y, int nt) {
  extern __shared__ float sx[];  // [kSHaloPix][cin+1]
  const int cin = a.cin, cinp = cin + 1;
  const int cout_pad = (a.cout + 15) / 16 * 16;
  const int tiles_per_img = tiles_x * tiles_y;
  const int n = blockIdx.x / tiles_per_img;
  const int rem = blockIdx.x % tiles_per_img;
  const int y0 = (rem / tiles_x) * kSTileH, x0 = (rem % tiles_x) * kSTileW;
  const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int py = p / kSTileW, px = p % kSTileW;
  const int y = y0 + py, x = x0 + px;
  const bool valid = (y < a.h) && (x < a.w);
  const int nchunks = cout_pad / 16;
  float loss = 0.f;

  for (int jb = 0; jb < nchunks; jb += 2) {
    const int j = jb + half;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int s = 0; s < a.num_src; ++s) {
      if (a.num_src > 1 || jb == 0) {  // block-uniform: a single source is staged once for all chunks
        __syncthreads();
        const T* src = reinterpret_cast<const T*>(a.src[s]);
        for (int idx = threadIdx.x; idx < kSHaloPix * cin; idx += 256) {
          const int hp = idx / cin, ci = idx % cin;
          const int gy = y0 - 1 + hp / kSHaloW, gx = x0 - 1 + hp % kSHaloW;
          float v = 0.f;
          if (gy >= 0 && gy < a.h && gx >= 0 && gx < a.w)
            v = to_f32(src[act_off(n, gy, gx, ci >> 3, a.h, a.w, cin >> 3) + (ci & 7)]);
          sx[hp * cinp + ci] = v;
        }
        __syncthreads();
      }
      if (j < nchunks) {
        const int co0 = j * 16;
        for (int tap = 0; tap < 9; ++tap) {
          const float* xp = sx + ((py + tap / 3) * kSHaloW + px + tap % 3) * cinp;
          if constexpr (sizeof(T) == 4) {
            // fp32 operand layout: [src][tap][ci][cout_pad]
            const float* wp = reinterpret_cast<const float*>(a.weights) +
                              (static_cast<size_t>(s * 9 + tap) * cin) * cout_pad + co0;
            for (int ci = 0; ci < cin; ++ci) {
              const float xv = xp[ci];
              const float4* w4 = reinterpret_cast<const float4*>(wp + static_cast<size_t>(ci) * cout_pad);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 w = __ldg(w4 + q);
                acc[4 * q + 0] = fmaf(xv, w.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(xv, w.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(xv, w.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(xv, w.w, acc[4 * q + 3]);
              }
            }
          } else {
            // tensor-core operand layout: [ntile][src][tap][chunk][co_in_tile][8] bf16
            const int ntile = co0 / nt, co_in = co0 % nt, ch = cin / 8;
            const uint4* wp;
            size_t wstride;   // uint4 units between consecutive 8-channel chunks
            if (a.wlayout == LV_W_KY_STACKED) {  // [src][kx][chunk][ky*cout_pad + co][8]
              wp = reinterpret_cast<const uint4*>(a.weights) +
                   (static_cast<size_t>(s * 3 + tap % 3) * ch) * (3 * cout_pad) + (tap / 3) * cout_pad + co0;
              wstride = 3 * cout_pad;
            } else {
              wp = reinterpret_cast<const uint4*>(a.weights) +
                   ((static_cast<size_t>(ntile) * a.num_src + s) * 9 + tap) * ch * nt + co_in;
              wstride = nt;
            }
            for (int c8 = 0; c8 < ch; ++c8) {
              float xv[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) xv[e] = xp[c8 * 8 + e];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const uint4 w = __ldg(wp + static_cast<size_t>(c8) * wstride + i);
                acc[i] = fmaf(xv[0], bf16_lo(w.x), acc[i]);
                acc[i] = fmaf(xv[1], bf16_hi(w.x), acc[i]);
                acc[i] = fmaf(xv[2], bf16_lo(w.y), acc[i]);
                acc[i] = fmaf(xv[3], bf16_hi(w.y), acc[i]);
                acc[i] = fmaf(xv[4], bf16_lo(w.z), acc[i]);
                acc[i] = fmaf(xv[5], bf16_hi(w.z), acc[i]);
                acc[i] = fmaf(xv[6], bf16_lo(w.w), acc[i]);
                acc[i] = fmaf(xv[7], bf16_hi(w.w), acc[i]);
              }
            }
          }
        }
      }
    }
    if (j < nchunks && valid) loss += conv_epilogue16<T>(a, n, y, x, j * 16, acc);
  }
  if (a.truth_hr != nullptr && a.loss_sum != nullptr) {
    loss = warp_sum(loss);
    if ((threadIdx.x & 31) == 0) atomicAdd(a.loss_sum, static_cast<double>(loss));
  }
}

int conv3x3_simt(const lv_conv_args& a, cudaStream_t stream) {
  const int tiles_x = (a.w + kSTileW - 1) / kSTileW, tiles_y = (a.h + kSTileH - 1) / kSTileH;
  const long long blocks = static_cast<long long>(a.n) * tiles_x * tiles_y;
  if (blocks == 0) return LV_OK;
  LV_CHECK_ARG(blocks < (1ll << 31), "conv3x3: too many tiles");
  LV_CHECK_ARG(a.cin <= 64 && a.cin % 8 == 0, "conv3x3 CUDA-core path supports cin <= 64, multiple of 8 (got %d)", a.cin);
  const size_t smem = static_cast<size_t>(kSHaloPix) * (a.cin + 1) * sizeof(float);
  const int cout_pad = (a.cout + 15) / 16 * 16;
  if (a.dtype == LV_F32) {
    conv3x3_simt_kernel<float><<<static_cast<unsigned>(blocks), 256, smem, stream>>>(a, tiles_x, tiles_y, cout_pad);
  } else {
    LV_CHECK_ARG(a.cin % 8 == 0, "bf16 conv needs cin %% 8 == 0");
    conv3x3_simt_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, smem, stream>>>(a, tiles_x, tiles_y,
                                                                                              pick_ntile(cout_pad));
  }
  LV_LAUNCH_OK();
  return LV_OK;
}

}  // namespace lv
