// LarvaHead conv (3 -> cout, K = 27, fp32 math on CUDA cores: 0.18 % of the MACs, HBM-bound) fused with the
// bicubic x4 base image -- both read the same LR tile -- plus the head weight gradient.
//
// Bicubic restates ATen upsample_bicubic2d for scale_factor=4, align_corners=False (reference
// models/LarvaNet.py:283-285): src = (d+0.5)/4-0.5, taps floor-1..floor+2, Keys cubic A=-0.75, indices clamped.
// Output row 4y+i uses t = {0.625, 0.875, 0.125, 0.375}[i] and first tap row y + {-2,-2,-1,-1}[i].
#include "lv_common.cuh"

namespace lv {

constexpr int kHT = 16;            // LR tile edge
constexpr int kHH = kHT + 4;       // +-2 halo for bicubic (conv uses +-1 of it)

__device__ __forceinline__ void cubic_coeffs(float t, float* c) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

template <typename T>
__global__ void __launch_bounds__(kHT* kHT)
head_bicubic_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                    const float* __restrict__ pre_w, const float* __restrict__ pre_b, T* __restrict__ fea,
                    float* __restrict__ base_hr, int N, int H, int W, int cout) {
  __shared__ float sx[3][kHH][kHH + 1];   // index-clamped LR tile (what bicubic reads)
  __shared__ float sc[3][kHT + 2][kHT + 3];  // conv input: optional 1x1 pre-conv, ZERO outside the image
  extern __shared__ float sw[];           // [27][cout] transposed weights, then [cout] bias
  const int tx = threadIdx.x % kHT, ty = threadIdx.x / kHT;
  const int x0 = blockIdx.x * kHT, y0 = blockIdx.y * kHT, n = blockIdx.z;

  for (int i = threadIdx.x; i < 27 * cout; i += blockDim.x) {
    const int k = i / cout, co = i % cout;   // k = c*9 + ky*3 + kx  (OIHW inner order)
    sw[i] = w[co * 27 + k];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sw[27 * cout + i] = (b != nullptr) ? b[i] : 0.f;
  for (int i = threadIdx.x; i < 3 * kHH * kHH; i += blockDim.x) {
    const int c = i / (kHH * kHH), r = (i / kHH) % kHH, q = i % kHH;
    const int gy = min(max(y0 - 2 + r, 0), H - 1), gx = min(max(x0 - 2 + q, 0), W - 1);
    sx[c][r][q] = x[((static_cast<size_t>(n) * 3 + c) * H + gy) * W + gx];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (kHT + 2) * (kHT + 2); i += blockDim.x) {
    const int r = i / (kHT + 2), q = i % (kHT + 2);
    const int gy = y0 - 1 + r, gx = x0 - 1 + q;
    const bool inb = gy >= 0 && gy < H && gx >= 0 && gx < W;
    float v0 = sx[0][r + 1][q + 1], v1 = sx[1][r + 1][q + 1], v2 = sx[2][r + 1][q + 1];
    if (pre_w != nullptr) {  // EDSR mean_shift: general 1x1 conv (models/edsr.py:129-136,197)
      const float u0 = pre_w[0] * v0 + pre_w[1] * v1 + pre_w[2] * v2 + pre_b[0];
      const float u1 = pre_w[3] * v0 + pre_w[4] * v1 + pre_w[5] * v2 + pre_b[1];
      const float u2 = pre_w[6] * v0 + pre_w[7] * v1 + pre_w[8] * v2 + pre_b[2];
      v0 = u0; v1 = u1; v2 = u2;
    }
    sc[0][r][q] = inb ? v0 : 0.f;
    sc[1][r][q] = inb ? v1 : 0.f;
    sc[2][r][q] = inb ? v2 : 0.f;
  }
  __syncthreads();

  const int gy = y0 + ty, gx = x0 + tx;
  if (gy >= H || gx >= W) return;

  // ---- conv 3 -> cout ----
  float in[27];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) in[c * 9 + ky * 3 + kx] = sc[c][ty + ky][tx + kx];
  for (int co0 = 0; co0 < cout; co0 += 16) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = sw[27 * cout + co0 + i];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(in[k], sw[k * cout + co0 + i], acc[i]);
    }
    store16_act(fea, n, gy, gx, co0, H, W, cout, acc);
  }

  // ---- bicubic x4 base ----
  if (base_hr == nullptr) return;
  const float tph[4] = {0.625f, 0.875f, 0.125f, 0.375f};
  const int first[4] = {-2, -2, -1, -1};
  float cw[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) cubic_coeffs(tph[i], cw[i]);
  const size_t W4 = static_cast<size_t>(W) * 4, H4 = static_cast<size_t>(H) * 4;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // horizontal pass on the 5 LR rows y-2..y+2 -> hz[row][j]
    float hz[5][4];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* p = &sx[c][ty + r][tx + 2 + first[j]];
        hz[r][j] = p[0] * cw[j][0] + p[1] * cw[j][1] + p[2] * cw[j][2] + p[3] * cw[j][3];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r0 = first[i] + 2;
      float4 o;
      o.x = hz[r0][0] * cw[i][0] + hz[r0 + 1][0] * cw[i][1] + hz[r0 + 2][0] * cw[i][2] + hz[r0 + 3][0] * cw[i][3];
      o.y = hz[r0][1] * cw[i][0] + hz[r0 + 1][1] * cw[i][1] + hz[r0 + 2][1] * cw[i][2] + hz[r0 + 3][1] * cw[i][3];
      o.z = hz[r0][2] * cw[i][0] + hz[r0 + 1][2] * cw[i][1] + hz[r0 + 2][2] * cw[i][2] + hz[r0 + 3][2] * cw[i][3];
      o.w = hz[r0][3] * cw[i][0] + hz[r0 + 1][3] * cw[i][1] + hz[r0 + 2][3] * cw[i][2] + hz[r0 + 3][3] * cw[i][3];
      *reinterpret_cast<float4*>(base_hr + ((static_cast<size_t>(n) * 3 + c) * H4 + (4 * gy + i)) * W4 + 4 * gx) = o;
    }
  }
}

// bicubic only, any channel count (LarvaNetModule.base on its own)
__global__ void __launch_bounds__(kHT* kHT)
bicubic_kernel(const float* __restrict__ x, float* __restrict__ base_hr, int NC, int H, int W) {
  __shared__ float sx[kHH][kHH + 1];
  const int tx = threadIdx.x % kHT, ty = threadIdx.x / kHT;
  const int x0 = blockIdx.x * kHT, y0 = blockIdx.y * kHT, nc = blockIdx.z;
  for (int i = threadIdx.x; i < kHH * kHH; i += blockDim.x) {
    const int r = i / kHH, q = i % kHH;
    const int gy = min(max(y0 - 2 + r, 0), H - 1), gx = min(max(x0 - 2 + q, 0), W - 1);
    sx[r][q] = x[(static_cast<size_t>(nc) * H + gy) * W + gx];
  }
  __syncthreads();
  const int gy = y0 + ty, gx = x0 + tx;
  if (gy >= H || gx >= W) return;
  const float tph[4] = {0.625f, 0.875f, 0.125f, 0.375f};
  const int first[4] = {-2, -2, -1, -1};
  float cw[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) cubic_coeffs(tph[i], cw[i]);
  float hz[5][4];
#pragma unroll
  for (int r = 0; r < 5; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* p = &sx[ty + r][tx + 2 + first[j]];
      hz[r][j] = p[0] * cw[j][0] + p[1] * cw[j][1] + p[2] * cw[j][2] + p[3] * cw[j][3];
    }
  const size_t W4 = static_cast<size_t>(W) * 4, H4 = static_cast<size_t>(H) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r0 = first[i] + 2;
    float4 o;
    o.x = hz[r0][0] * cw[i][0] + hz[r0 + 1][0] * cw[i][1] + hz[r0 + 2][0] * cw[i][2] + hz[r0 + 3][0] * cw[i][3];
    o.y = hz[r0][1] * cw[i][0] + hz[r0 + 1][1] * cw[i][1] + hz[r0 + 2][1] * cw[i][2] + hz[r0 + 3][1] * cw[i][3];
    o.z = hz[r0][2] * cw[i][0] + hz[r0 + 1][2] * cw[i][1] + hz[r0 + 2][2] * cw[i][2] + hz[r0 + 3][2] * cw[i][3];
    o.w = hz[r0][3] * cw[i][0] + hz[r0 + 1][3] * cw[i][1] + hz[r0 + 2][3] * cw[i][2] + hz[r0 + 3][3] * cw[i][3];
    *reinterpret_cast<float4*>(base_hr + (static_cast<size_t>(nc) * H4 + (4 * gy + i)) * W4 + 4 * gx) = o;
  }
}

// head weight gradient: dw[co][c][ky][kx] (+)= scale * sum_px dy[px][co] * x[c][y+ky-1][x+kx-1];  db[co] (+)= scale * sum dy
// Two passes, no atomics (deterministic): blocks loop over 8x16 pixel tiles and keep their partial sums in registers --
// thread = (pair of output channels, one (c,ky) row of three kx taps): 6 FMAs per 5 shared-memory loads -- and store ONE
// partial vector [cout*27 + cout] per block; head_wgrad_reduce_kernel sums the blocks' vectors in a fixed order, scales
// and stores (or adds to) dw / db.  Up to four blocks per SM overlap each other's load and compute phases.
constexpr int kGW = 16, kGH = 8;  // pixel tile of the head wgrad
template <typename T>
__global__ void __launch_bounds__(288)
head_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* __restrict__ partial,
                  int N, int H, int W, int cout) {
  __shared__ float sx[3][kGH + 2][kGW + 2];
  extern __shared__ float sdy[];  // [kGH*kGW][cout+2]
  const int cp = cout + 2;
  const int half = cout / 2;
  const int nthr = half * 9;      // active threads: 216 (cout 48) or 288 (cout 64)
  const int t = threadIdx.x;
  const bool active = t < nthr;
  const int cog = t % half, kg = t / half;          // channel pair, (c,ky) index 0..8
  const int kc = kg / 3, ky = kg % 3;
  const int tiles_x = (W + kGW - 1) / kGW, tiles_y = (H + kGH - 1) / kGH;
  const int tiles_per_img = tiles_x * tiles_y, total = N * tiles_per_img;
  float acc[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  float bacc[2] = {0.f, 0.f};
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
    const int y0 = (rem / tiles_x) * kGH, x0 = (rem % tiles_x) * kGW;
    __syncthreads();
    for (int i = t; i < 3 * (kGH + 2) * (kGW + 2); i += blockDim.x) {
      const int c = i / ((kGH + 2) * (kGW + 2)), r = (i / (kGW + 2)) % (kGH + 2), q = i % (kGW + 2);
      const int gy = y0 - 1 + r, gx = x0 - 1 + q;
      sx[c][r][q] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? x[((static_cast<size_t>(n) * 3 + c) * H + gy) * W + gx] : 0.f;
    }
    // one 8-channel chunk (16 B in bf16) per thread and step: consecutive threads read consecutive pixels of a chunk row
    const int chunks = cout >> 3;
    for (int i = t; i < kGH * kGW * chunks; i += blockDim.x) {
      const int q = i % kGW, rc = i / kGW;
      const int ch = rc % chunks, r = rc / chunks;
      const int gy = y0 + r, gx = x0 + q;
      float v[8];
      if (gy < H && gx < W) {
        const T* src = dy + act_off(n, gy, gx, ch, H, W, chunks);
        if constexpr (sizeof(T) == 2) {
          load8(reinterpret_cast<const __nv_bfloat16*>(src), v);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = to_f32(src[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
      float* d = &sdy[(r * kGW + q) * cp + ch * 8];
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = v[e];
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int p = 0; p < kGH * kGW; ++p) {
        const float2 d = *reinterpret_cast<const float2*>(&sdy[p * cp + 2 * cog]);
        const float* xr = &sx[kc][(p >> 4) + ky][p & 15];
        const float x0v = xr[0], x1v = xr[1], x2v = xr[2];
        acc[0][0] = fmaf(d.x, x0v, acc[0][0]); acc[0][1] = fmaf(d.x, x1v, acc[0][1]); acc[0][2] = fmaf(d.x, x2v, acc[0][2]);
        acc[1][0] = fmaf(d.y, x0v, acc[1][0]); acc[1][1] = fmaf(d.y, x1v, acc[1][1]); acc[1][2] = fmaf(d.y, x2v, acc[1][2]);
        if (kg == 0) { bacc[0] += d.x; bacc[1] += d.y; }
      }
    }
  }
  if (active) {
    float* pv = partial + static_cast<size_t>(blockIdx.x) * (cout * 28);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int co = 2 * cog + j;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) pv[co * 27 + kc * 9 + ky * 3 + kx] = acc[j][kx];
      if (kg == 0) pv[cout * 27 + co] = bacc[j];
    }
  }
}

// second pass: out[i] (+)= scale * sum over blocks of partial[block][i], i in [0, cout*28): 16 threads per output split the
// blocks, fixed summation order
__global__ void __launch_bounds__(1024)
head_wgrad_reduce_kernel(const float* __restrict__ partial, int nblocks, int nout, int cout, float* __restrict__ dw,
                         float* __restrict__ db, float scale, int overwrite) {
  __shared__ float red[16][64];
  const int o = blockIdx.x * 64 + (threadIdx.x & 63), part = threadIdx.x >> 6;
  float s = 0.f;
  if (o < nout) {
    for (int b = part; b < nblocks; b += 16) s += partial[static_cast<size_t>(b) * nout + o];
  }
  red[part][threadIdx.x & 63] = s;
  __syncthreads();
  if (part == 0 && o < nout) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += red[i][threadIdx.x & 63];
    t *= scale;
    float* dst = (o < cout * 27) ? (dw + o) : (db != nullptr ? db + (o - cout * 27) : nullptr);
    if (dst != nullptr) *dst = overwrite ? t : (*dst + t);
  }
}

int head_bicubic_fwd(const float* x, const float* w, const float* b, const float* pre_w, const float* pre_b, void* fea,
                     float* base_hr, int n, int h, int w_, int cout, int dtype, cudaStream_t stream) {
  if (n == 0 || h == 0 || w_ == 0) return LV_OK;
  LV_CHECK_ARG(cout % 16 == 0 && cout <= 256, "head conv: cout must be a multiple of 16 (got %d)", cout);
  LV_CHECK_ARG(n <= 65535, "head conv: batch too large for one launch");
  dim3 grid((w_ + kHT - 1) / kHT, (h + kHT - 1) / kHT, n);
  const size_t smem = static_cast<size_t>(28) * cout * sizeof(float);
  if (dtype == LV_F32)
    head_bicubic_kernel<float><<<grid, kHT * kHT, smem, stream>>>(x, w, b, pre_w, pre_b, static_cast<float*>(fea), base_hr,
                                                                 n, h, w_, cout);
  else
    head_bicubic_kernel<__nv_bfloat16><<<grid, kHT * kHT, smem, stream>>>(
        x, w, b, pre_w, pre_b, static_cast<__nv_bfloat16*>(fea), base_hr, n, h, w_, cout);
  LV_LAUNCH_OK();
  return LV_OK;
}

int bicubic_x4(const float* x, float* base_hr, int n, int c, int h, int w_, cudaStream_t stream) {
  if (n * c == 0 || h == 0 || w_ == 0) return LV_OK;
  LV_CHECK_ARG(static_cast<long long>(n) * c <= 65535, "bicubic: n*c too large for one launch");
  dim3 grid((w_ + kHT - 1) / kHT, (h + kHT - 1) / kHT, n * c);
  bicubic_kernel<<<grid, kHT * kHT, 0, stream>>>(x, base_hr, n * c, h, w_);
  LV_LAUNCH_OK();
  return LV_OK;
}

constexpr int kHeadWgradMaxBlocks = 4 * 148;

long long head_wgrad_workspace_bytes(int cout) {
  return static_cast<long long>(kHeadWgradMaxBlocks) * cout * 28 * sizeof(float);
}

int head_wgrad(const float* x, const void* dy, float* dw, float* db, int n, int h, int w_, int cout, int dtype,
               float scale, void* workspace, long long workspace_bytes, int overwrite, cudaStream_t stream) {
  if (n == 0 || h == 0 || w_ == 0) return LV_OK;
  LV_CHECK_ARG(cout <= 64 && cout % 2 == 0, "head wgrad: cout must be even and <= 64 (got %d)", cout);
  const long long tiles = static_cast<long long>(n) * ((w_ + kGW - 1) / kGW) * ((h + kGH - 1) / kGH);
  LV_CHECK_ARG(tiles < (1ll << 31), "head wgrad: too many tiles");
  // up to four co-resident blocks per SM (shared memory 27 KB, 288 threads): their load / compute phases overlap
  long long cap = 4ll * sm_count();
  if (cap > kHeadWgradMaxBlocks) cap = kHeadWgradMaxBlocks;
  const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
  const int nout = cout * 28;
  LV_CHECK_ARG(workspace != nullptr && workspace_bytes >= static_cast<long long>(grid) * nout * 4,
               "head wgrad: workspace too small (%lld bytes; lv_head_wgrad_workspace_bytes(cout) gives the size)", workspace_bytes);
  float* partial = static_cast<float*>(workspace);
  const size_t smem = static_cast<size_t>(kGH * kGW) * (cout + 2) * sizeof(float);
  if (dtype == LV_F32)
    head_wgrad_kernel<float><<<grid, 288, smem, stream>>>(x, static_cast<const float*>(dy), partial, n, h, w_, cout);
  else
    head_wgrad_kernel<__nv_bfloat16><<<grid, 288, smem, stream>>>(x, static_cast<const __nv_bfloat16*>(dy), partial, n, h,
                                                                  w_, cout);
  LV_LAUNCH_OK();
  head_wgrad_reduce_kernel<<<(nout + 63) / 64, 1024, 0, stream>>>(partial, static_cast<int>(grid), nout, cout, dw, db, scale,
                                                                  overwrite);
  LV_LAUNCH_OK();
  return LV_OK;
}

}  // namespace lv
