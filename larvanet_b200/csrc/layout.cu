// Operand packing and layout conversion helpers (HBM-bound, tiny).
//   * weight packing: fp32 OIHW master weights -> the operand layout each conv kernel consumes
//       LV_BF16: [ntile][src][tap][cin/8][cout_in_tile][8] bf16  (K-major SWIZZLE_NONE UMMA B operand, see conv_tc.cu)
//       LV_F32 : [src][tap][cin][cout_pad] fp32                  (conv_simt.cu)
//     with an optional "transpose" that yields the backward-data operand (180-degree rotated taps, in/out swapped).
//   * NCHW fp32 <-> NHWC {bf16,fp32} at the Python module boundary (the reference's tensors are NCHW fp32,
//     models/LarvaNet.py:163-171).
//   * stand-alone L1 loss + sign gradient, fused AdamW.
#include "lv_common.cuh"

namespace lv {

int pick_ntile(int cout_pad);

constexpr int kMaxPackItems = 128;   // LV_PACK_MAX_ITEMS: 128 x 80 B kernel parameter
struct PackItemDev {
  const float* w;
  void* packed;
  int O, I, transpose, i_off, i_cnt, cin, dtype, wlayout;
  int p_cout, p_cin_total, cout_pad, nt, nsrc;
  long long total;  // packed elements
};
struct PackBatch {
  PackItemDev it[kMaxPackItems];
};

__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackBatch batch) {
  const PackItemDev& p = batch.it[blockIdx.y];
  for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < p.total; idx += static_cast<long long>(gridDim.x) * 256) {
    int co, t, tap;
    if (p.dtype == LV_BF16 && p.wlayout == LV_W_KY_STACKED) {
      // [src][kx][chunk][ky*cout_pad + co][8]
      long long r = idx;
      const int e = static_cast<int>(r % 8); r /= 8;
      const int nrow = static_cast<int>(r % (3 * p.cout_pad)); r /= 3 * p.cout_pad;
      const int ch = p.cin / 8;
      const int chunk = static_cast<int>(r % ch); r /= ch;
      const int kx = static_cast<int>(r % 3); r /= 3;
      const int s = static_cast<int>(r);
      co = nrow % p.cout_pad;
      tap = (nrow / p.cout_pad) * 3 + kx;
      t = s * p.cin + chunk * 8 + e;
    } else if (p.dtype == LV_BF16) {
      long long r = idx;
      const int e = static_cast<int>(r % 8); r /= 8;
      const int co_in = static_cast<int>(r % p.nt); r /= p.nt;
      const int ch = p.cin / 8;
      const int chunk = static_cast<int>(r % ch); r /= ch;
      tap = static_cast<int>(r % 9); r /= 9;
      const int s = static_cast<int>(r % p.nsrc); r /= p.nsrc;
      const int ntile = static_cast<int>(r);
      co = ntile * p.nt + co_in;
      t = s * p.cin + chunk * 8 + e;
    } else {
      long long r = idx;
      co = static_cast<int>(r % p.cout_pad); r /= p.cout_pad;
      const int ci = static_cast<int>(r % p.cin); r /= p.cin;
      tap = static_cast<int>(r % 9); r /= 9;
      const int s = static_cast<int>(r);
      t = s * p.cin + ci;
    }
    float v = 0.f;
    if (co < p.p_cout) {
      int o, i, ky = tap / 3, kx = tap % 3;
      if (p.transpose) {
        o = t; i = p.i_off + co; ky = 2 - ky; kx = 2 - kx;
      } else {
        o = co; i = p.i_off + t;
      }
      v = p.w[((static_cast<size_t>(o) * p.I + i) * 3 + ky) * 3 + kx];
    }
    if (p.dtype == LV_BF16)
      reinterpret_cast<__nv_bfloat16*>(p.packed)[idx] = __float2bfloat16_rn(v);
    else
      reinterpret_cast<float*>(p.packed)[idx] = v;
  }
}

int pack_weights(const lv_pack_item* items, int count, cudaStream_t stream) {
  if (count == 0) return LV_OK;
  LV_CHECK_ARG(count > 0 && count <= kMaxPackItems, "pack: count must be in 1..%d", kMaxPackItems);
  PackBatch batch;
  long long max_total = 0;
  for (int k = 0; k < count; ++k) {
    const lv_pack_item& s = items[k];
    PackItemDev& d = batch.it[k];
    LV_CHECK_ARG(s.w != nullptr && s.packed != nullptr, "pack: null pointer in item %d", k);
    LV_CHECK_ARG(s.cin > 0 && s.i_cnt > 0 && s.i_off >= 0 && s.i_off + s.i_cnt <= s.I, "pack: bad slice in item %d", k);
    d.w = s.w; d.packed = s.packed; d.O = s.O; d.I = s.I; d.transpose = s.transpose; d.i_off = s.i_off;
    d.i_cnt = s.i_cnt; d.cin = s.cin; d.dtype = s.dtype; d.wlayout = s.wlayout;
    d.p_cout = s.transpose ? s.i_cnt : s.O;
    d.p_cin_total = s.transpose ? s.O : s.i_cnt;
    LV_CHECK_ARG(d.p_cin_total % s.cin == 0, "pack: cin_total %d not a multiple of per-source cin %d", d.p_cin_total, s.cin);
    if (s.dtype == LV_BF16) LV_CHECK_ARG(s.cin % 16 == 0, "pack: bf16 operands need cin %% 16 == 0 (got %d)", s.cin);
    d.nsrc = d.p_cin_total / s.cin;
    d.cout_pad = (d.p_cout + 15) / 16 * 16;
    d.nt = pick_ntile(d.cout_pad);
    if (s.wlayout == LV_W_KY_STACKED)
      LV_CHECK_ARG(s.dtype == LV_BF16 && 3 * d.cout_pad <= 256, "pack: ky-stacked layout needs bf16 and cout <= 80 (item %d)", k);
    d.total = static_cast<long long>(d.cout_pad) * d.p_cin_total * 9;
    if (d.total > max_total) max_total = d.total;
  }
  long long bx = (max_total + 255) / 256;
  if (bx > 1024) bx = 1024;
  pack_weights_kernel<<<dim3(static_cast<unsigned>(bx), count), 256, 0, stream>>>(batch);
  LV_LAUNCH_OK();
  return LV_OK;
}

// ---------------------------------------------------------------------------------------------
// NCHW fp32 <-> planar-8 activation layout [N][H][C/8][W][8]; one thread per destination element
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int H, int W, long long total) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    long long r = i;
    const int e = static_cast<int>(r % 8); r /= 8;
    const int x = static_cast<int>(r % W); r /= W;
    const int ch = static_cast<int>(r % (C / 8)); r /= (C / 8);
    const int y = static_cast<int>(r % H);
    const long long n = r / H;
    dst[i] = from_f32<T>(src[((n * C + ch * 8 + e) * H + y) * W + x]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int H, int W, long long total) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    long long r = i;
    const int x = static_cast<int>(r % W); r /= W;
    const int y = static_cast<int>(r % H); r /= H;
    const int c = static_cast<int>(r % C);
    const long long n = r / C;
    dst[i] = to_f32(src[act_off(static_cast<int>(n), y, x, c >> 3, H, W, C >> 3) + (c & 7)]);
  }
}

static unsigned grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return static_cast<unsigned>(b < 1 ? 1 : b);
}

int nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w, int dtype, cudaStream_t stream) {
  const long long total = static_cast<long long>(n) * c * h * w;
  if (total == 0) return LV_OK;
  LV_CHECK_ARG(c % 8 == 0, "activation layout needs channels %% 8 == 0 (got %d)", c);
  if (dtype == LV_F32)
    nchw_to_nhwc_kernel<float><<<grid_for(total), 256, 0, stream>>>(src, static_cast<float*>(dst), c, h, w, total);
  else
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), c, h, w,
                                                                            total);
  LV_LAUNCH_OK();
  return LV_OK;
}
int nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w, int dtype, cudaStream_t stream) {
  const long long total = static_cast<long long>(n) * c * h * w;
  if (total == 0) return LV_OK;
  LV_CHECK_ARG(c % 8 == 0, "activation layout needs channels %% 8 == 0 (got %d)", c);
  if (dtype == LV_F32)
    nhwc_to_nchw_kernel<float><<<grid_for(total), 256, 0, stream>>>(static_cast<const float*>(src), dst, c, h, w, total);
  else
    nhwc_to_nchw_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), dst, c,
                                                                            h, w, total);
  LV_LAUNCH_OK();
  return LV_OK;
}

// ---------------------------------------------------------------------------------------------
// L1 loss on HR fp32 NCHW images; thread = (n, c, y, x) LR position covering its 4x4 HR block.
template <typename T>
__global__ void __launch_bounds__(256)
l1_loss_grad_kernel(const float* __restrict__ out, const float* __restrict__ truth, double* __restrict__ loss_sum,
                    T* __restrict__ grad_sign, int N, int C, int H, int W) {
  const long long total = static_cast<long long>(N) * C * H * W;
  float loss = 0.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const int c = static_cast<int>((i / (static_cast<long long>(W) * H)) % C);
    const long long n = i / (static_cast<long long>(W) * H * C);
    float g[16];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const size_t off = ((n * C + c) * (4ll * H) + (4 * y + r)) * (4ll * W) + 4 * x;
      const float4 o = *reinterpret_cast<const float4*>(out + off);
      const float4 t = *reinterpret_cast<const float4*>(truth + off);
      const float d[4] = {o.x - t.x, o.y - t.y, o.z - t.z, o.w - t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        loss += fabsf(d[j]);
        g[4 * r + j] = (d[j] > 0.f) ? 1.f : ((d[j] < 0.f) ? -1.f : 0.f);
      }
    }
    if (grad_sign != nullptr) store16_act(grad_sign, static_cast<int>(n), y, x, 16 * c, H, W, 16 * C, g);
  }
  loss = warp_sum(loss);
  if ((threadIdx.x & 31) == 0 && loss_sum != nullptr) atomicAdd(loss_sum, static_cast<double>(loss));
}

int l1_loss_grad(const float* out_hr, const float* truth_hr, double* loss_sum, void* grad_sign, int n, int c, int h, int w,
                 int dtype, cudaStream_t stream) {
  const long long total = static_cast<long long>(n) * c * h * w;
  if (total == 0) return LV_OK;
  if (dtype == LV_F32)
    l1_loss_grad_kernel<float><<<grid_for(total), 256, 0, stream>>>(out_hr, truth_hr, loss_sum, static_cast<float*>(grad_sign),
                                                                   n, c, h, w);
  else
    l1_loss_grad_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, stream>>>(
        out_hr, truth_hr, loss_sum, static_cast<__nv_bfloat16*>(grad_sign), n, c, h, w);
  LV_LAUNCH_OK();
  return LV_OK;
}

// ---------------------------------------------------------------------------------------------
// uint8 image helpers (validate._image_to_uint8 / _image_psnr, reference validate.py:17-27): np.round is
// round-half-to-even == __float2int_rn; then clip to 0..255.
__device__ __forceinline__ int to_u8(float v) {
  const int r = __float2int_rn(v);
  return r < 0 ? 0 : (r > 255 ? 255 : r);
}

__global__ void __launch_bounds__(256)
image_to_uint8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, long long total) {
  // four pixels per thread: one 16 B load, one 4 B store
  const long long quads = total / 4;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < quads; i += static_cast<long long>(gridDim.x) * 256) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uchar4 o;
    o.x = static_cast<unsigned char>(to_u8(v.x)); o.y = static_cast<unsigned char>(to_u8(v.y));
    o.z = static_cast<unsigned char>(to_u8(v.z)); o.w = static_cast<unsigned char>(to_u8(v.w));
    reinterpret_cast<uchar4*>(dst)[i] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x < static_cast<int>(total - quads * 4)) {
    const long long i = quads * 4 + threadIdx.x;
    dst[i] = static_cast<unsigned char>(to_u8(src[i]));
  }
}

int image_to_uint8(const float* src, uint8_t* dst, long long total, cudaStream_t stream) {
  if (total == 0) return LV_OK;
  LV_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0,
               "image_to_uint8: src must be 16-byte and dst 4-byte aligned");
  image_to_uint8_kernel<<<grid_for((total + 3) / 4), 256, 0, stream>>>(src, dst, total);
  LV_LAUNCH_OK();
  return LV_OK;
}

// sum over (c, y < h, x < w) of (u8(truth[c][y][x]) - u8(out[c][y][x]))^2, truth cropped to the output's size
// (validate._fit_truth_image_size); exact in integers, accumulated as double
__global__ void __launch_bounds__(256)
psnr_sqsum_kernel(const float* __restrict__ out, const float* __restrict__ truth, double* __restrict__ sq_sum, int c, int h,
                  int w, int th, int tw) {
  const long long total = static_cast<long long>(c) * h * w;
  unsigned long long acc = 0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int x = static_cast<int>(i % w);
    const int y = static_cast<int>((i / w) % h);
    const int ch = static_cast<int>(i / (static_cast<long long>(w) * h));
    const int d = to_u8(truth[(static_cast<long long>(ch) * th + y) * tw + x]) - to_u8(out[i]);
    acc += static_cast<unsigned long long>(d * d);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(sq_sum, static_cast<double>(acc));
}

int psnr_sqsum(const float* out, const float* truth, double* sq_sum, int c, int h, int w, int th, int tw, cudaStream_t stream) {
  LV_CHECK_ARG(th >= h && tw >= w, "psnr: truth (%d x %d) smaller than output (%d x %d)", th, tw, h, w);
  const long long total = static_cast<long long>(c) * h * w;
  if (total == 0) return LV_OK;
  psnr_sqsum_kernel<<<grid_for(total), 256, 0, stream>>>(out, truth, sq_sum, c, h, w, th, tw);
  LV_LAUNCH_OK();
  return LV_OK;
}

// ---------------------------------------------------------------------------------------------
// Device-side training-patch pipeline: random crop + rot90 + horizontal flip of an LR / HR image pair that is resident
// in HBM, written straight into the step's batch tensors (reference dataloaders/div2k_train_loader.py:72-98 on the host,
// dataloaders/div2k_train_loader_tensor.py:57-97 with torch ops).  One item per patch; out(i, j) of a P x P patch:
//   undo the flip (j' = P-1-j), undo torch.rot90(k, dims=(1,2)) (k counter-clockwise quarter turns), read the crop.
struct PatchItem {           // mirrors lv_patch_item
  const float* lr;           // fp32 [3, h, w]
  const float* hr;           // fp32 [3, scale*h, scale*w]
  int h, w, y, x, rot, flip;
};

__device__ __forceinline__ void unrotate(int i, int j, int P, int rot, int flip, int& si, int& sj) {
  if (flip) j = P - 1 - j;
  switch (rot & 3) {
    case 0: si = i; sj = j; break;
    case 1: si = j; sj = P - 1 - i; break;            // out = rot90(in, 1): out[i][j] = in[j][P-1-i]
    case 2: si = P - 1 - i; sj = P - 1 - j; break;
    default: si = P - 1 - j; sj = i; break;           // rot90(in, 3): out[i][j] = in[P-1-j][i]
  }
}

__global__ void __launch_bounds__(256)
crop_augment_kernel(const PatchItem* __restrict__ items, float* __restrict__ out_lr, float* __restrict__ out_hr, int P, int scale) {
  const PatchItem it = items[blockIdx.y];
  const int PH = P * scale;
  const long long n_lr = 3ll * P * P, n_hr = 3ll * PH * PH;
  float* dl = out_lr + static_cast<long long>(blockIdx.y) * n_lr;
  float* dh = out_hr + static_cast<long long>(blockIdx.y) * n_hr;
  for (long long e = blockIdx.x * 256ll + threadIdx.x; e < n_lr + n_hr; e += static_cast<long long>(gridDim.x) * 256) {
    const bool hi = e >= n_lr;
    const long long q = hi ? e - n_lr : e;
    const int S = hi ? PH : P;
    const int c = static_cast<int>(q / (static_cast<long long>(S) * S));
    const int r = static_cast<int>(q - static_cast<long long>(c) * S * S);
    int si, sj;
    unrotate(r / S, r % S, S, it.rot, it.flip, si, sj);
    if (hi) {
      const int W = it.w * scale, H = it.h * scale;
      dh[q] = it.hr[(static_cast<long long>(c) * H + it.y * scale + si) * W + it.x * scale + sj];
    } else {
      dl[q] = it.lr[(static_cast<long long>(c) * it.h + it.y + si) * it.w + it.x + sj];
    }
  }
}

int crop_augment(const lv_patch_item* items_dev, int count, float* out_lr, float* out_hr, int patch, int scale,
                 cudaStream_t stream) {
  if (count == 0) return LV_OK;
  LV_CHECK_ARG(count > 0 && count <= 65535 && patch > 0 && scale > 0, "crop_augment: bad count / patch / scale");
  static_assert(sizeof(PatchItem) == sizeof(lv_patch_item), "lv_patch_item layout");
  const long long per = 3ll * patch * patch * (1 + scale * scale);
  long long bx = (per + 255) / 256;
  if (bx > 64) bx = 64;
  crop_augment_kernel<<<dim3(static_cast<unsigned>(bx), count), 256, 0, stream>>>(
      reinterpret_cast<const PatchItem*>(items_dev), out_lr, out_hr, patch, scale);
  LV_LAUNCH_OK();
  return LV_OK;
}

// ---------------------------------------------------------------------------------------------
// AdamW over a flat arena, torch.optim.AdamW semantics (decoupled weight decay, bias-corrected).
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long numel, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < numel; i += static_cast<long long>(gridDim.x) * 256) {
    const float grad = g[i] * gscale;
    float pv = p[i] * (1.f - lr * wd);
    const float mv = m[i] + (grad - m[i]) * (1.f - b1);            // lerp_, as torch does
    const float vv = v[i] * b2 + (1.f - b2) * grad * grad;
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= (lr / bc1) * (mv / denom);
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
}

int adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long numel, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, float grad_scale, cudaStream_t stream) {
  if (numel == 0) return LV_OK;
  LV_CHECK_ARG(step >= 1, "adamw: step must be >= 1");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adamw_kernel<<<grid_for(numel), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps,
                                                    weight_decay, static_cast<float>(bc1), static_cast<float>(sqrt(bc2)),
                                                    grad_scale);
  LV_LAUNCH_OK();
  return LV_OK;
}


// ---------------------------------------------------------------------------------------------
// AdamW + bf16 operand re-pack in ONE kernel.  A block owns "bricks" of a conv weight: 8 output x 8
// input channels x 9 taps = 576 fp32 values that are 8 contiguous 288 B runs of the OIHW master weight.  It updates
// them (same arithmetic as adamw_kernel), writes the parameter and both moments back, and emits the brick's part of the
// forward operand ([src][tap][chunk][co][8]: 9 x 128 B) and of the 180-degree-rotated, in/out-swapped backward-data
// operand ([tap'][chunk = o/8][co' = i][8 = o%8]: 9 x 128 B) from shared memory -- both sides coalesced.  Everything
// that is not a packed conv weight (biases, the head conv) is updated by a few tail blocks from a list of 48-element rows.
constexpr int kFusedMaxConvs = 64, kFusedMaxRows = 512, kFusedRow = 48, kFusedRowsPerBlock = 4;
struct FusedConv {
  long long w_off;            // element offset of the OIHW weight [48, I, 3, 3] in the parameter arena
  int I, brick0;              // input channels (48 * sources); index of its first brick
  __nv_bfloat16* fwd;         // forward operand
  __nv_bfloat16* bwd[LV_MAX_SRC];   // backward-data operand per 48-channel source slice
};
struct FusedBatch {
  FusedConv conv[kFusedMaxConvs];
  // everything that is not a packed conv weight, cut into rows of <= 48 elements (4 rows per tail block: a serial
  // loop over ~45 bias vectors in one block would cost ~1.5 us of memory latency each)
  long long row_off[kFusedMaxRows];
  short row_cnt[kFusedMaxRows];
  int nconv, nrows, nbricks, brick_blocks;
};

// ---- data-parallel variant: the gradient of element i is the SUM over all ranks' gradient arenas, read straight out of
// the peers' HBM over NVLink (symmetric / peer-mapped memory), in rank order on every rank -- so every rank computes the
// bit-identical sum and the replicas' weights never drift apart.  One kernel = device-side barrier ("every rank's
// gradients are complete") + all-reduce + AdamW + operand re-pack + device-side barrier ("every rank is done reading").
// Two ways to read the sum (template TWO):
//   one-shot (world 2): every rank reads all ranks' arenas -- (world-1) x 3.3 MB over NVLink, one barrier pair;
//   two-shot (world >= 4; at 8 ranks the one-shot read is 23 MB = 30 us): rank r first reduces ITS 1/world slice from all
//     peers into its `reduced` buffer, a second device-side barrier follows, and the update then reads every element
//     from its owner's reduced buffer -- 2 x (world-1)/world x 3.3 MB per rank.
constexpr int kDpMaxRanks = 16;
struct DpCtx {
  const float* grad[kDpMaxRanks];     // gradient arena of rank r (peer-mapped address on this device)
  float* reduced[kDpMaxRanks];        // two-shot: rank r's buffer of reduced slices (only slice r is ever written by r)
  uint32_t* flags[kDpMaxRanks];       // flag block of rank r: three barriers x one slot per writer rank (start, end, middle)
  const double* loss[kDpMaxRanks];    // local loss accumulator of rank r
  double* loss_out;                   // this rank: sum over ranks
  uint32_t* ctl;                      // this rank, device-local control words: [0] epoch, [1] go, [2] blocks arrived (end),
                                      // [3] blocks arrived (middle), [4] go (middle)
  long long slice;                    // two-shot: elements per rank's slice (multiple of 4; world * slice >= numel)
  int world, rank;
};
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer memory is not cached in the local L2; bypass L1 too so that a line read in an earlier step is never reused
__device__ __forceinline__ float ld_peer_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until_sys(const uint32_t* p, uint32_t need) {
  uint32_t spins = 0;
  while (ld_acquire_sys(p) < need) {
    __nanosleep(64);
    if (++spins > (1u << 24)) asm volatile("trap;");   // ~1-2 s: a peer died; fail loudly instead of hanging the GPU
  }
}

__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

template <int NP, bool TWO>
__device__ __forceinline__ float grad_at(const float* g, const DpCtx& dp, long long i) {
  if constexpr (NP == 0) {
    return g[i];
  } else if constexpr (TWO) {
    // the element's owner reduced it in phase A
    const unsigned owner = static_cast<unsigned>(i) / static_cast<unsigned>(dp.slice);
    return ld_peer_f32(dp.reduced[owner] + i);
  } else {
    float part[NP];
#pragma unroll
    for (int r = 0; r < NP; ++r) part[r] = ld_peer_f32(dp.grad[r] + i);   // all loads in flight before the first add
    float s = part[0];
#pragma unroll
    for (int r = 1; r < NP; ++r) s += part[r];
    return s;
  }
}

template <int NP, bool TWO>
__device__ __forceinline__ float adamw_one(float* p, const float* g, const DpCtx& dp, float* m, float* v, long long i, float lr,
                                           float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  const float grad = grad_at<NP, TWO>(g, dp, i) * gscale;
  float pv = p[i] * (1.f - lr * wd);
  const float mv = m[i] + (grad - m[i]) * (1.f - b1);            // lerp_, as torch does
  const float vv = v[i] * b2 + (1.f - b2) * grad * grad;
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  pv -= (lr / bc1) * (mv / denom);
  p[i] = pv; m[i] = mv; v[i] = vv;
  return pv;
}

// two-shot phase A + middle barrier: this rank's slice of the sum, then "every rank's slice is reduced and visible"
template <int NP>
__device__ __forceinline__ void dp_reduce_slice(const DpCtx& dp, uint32_t e) {
  __shared__ uint32_t s_last_mid;
  const int t = threadIdx.x;
  const long long base = static_cast<long long>(dp.rank) * dp.slice;
  const long long n4 = dp.slice / 4;
  float* dst = dp.reduced[dp.rank] + base;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + t; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 part[NP];
#pragma unroll
    for (int r = 0; r < NP; ++r) part[r] = ld_peer_v4(dp.grad[r] + base + 4 * i);
    float4 s = part[0];
#pragma unroll
    for (int r = 1; r < NP; ++r) { s.x += part[r].x; s.y += part[r].y; s.z += part[r].z; s.w += part[r].w; }   // rank order: identical sums
    *reinterpret_cast<float4*>(dst + 4 * i) = s;
  }
  // grid barrier (every block of this rank is resident: grid <= co-resident blocks, checked on the host), then the
  // cross-rank barrier by the last block to arrive, then release the grid
  __syncthreads();
  if (t == 0) {
    __threadfence_system();
    const uint32_t prev = atomicAdd(dp.ctl + 3, 1u);
    s_last_mid = (prev == gridDim.x - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last_mid != 0u) {
    if (t < NP) {
      st_release_sys(dp.flags[t] + 2 * kDpMaxRanks + dp.rank, e);
      spin_until_sys(dp.flags[dp.rank] + 2 * kDpMaxRanks + t, e);
    }
    __syncthreads();
    if (t == 0) {
      dp.ctl[3] = 0u;
      st_release_gpu(dp.ctl + 4, e);
    }
  } else if (t == 0) {
    uint32_t spins = 0;
    while (ld_acquire_gpu_u32(dp.ctl + 4) < e) {
      __nanosleep(64);
      if (++spins > (1u << 24)) asm volatile("trap;");
    }
  }
  __syncthreads();
}

// entry barrier of the data-parallel kernel; returns this launch's epoch
template <int NP>
__device__ __forceinline__ uint32_t dp_enter(const DpCtx& dp) {
  __shared__ uint32_t s_epoch;
  const int t = threadIdx.x;
  if (t == 0) s_epoch = ld_acquire_gpu_u32(dp.ctl) + 1u;   // committed by the last block to leave
  __syncthreads();
  const uint32_t e = s_epoch;
  if (blockIdx.x == 0) {
    // tell every rank "my gradients (written by earlier kernels of my stream) are complete", wait for everybody's
    if (t < NP) {
      st_release_sys(dp.flags[t] + dp.rank, e);
      spin_until_sys(dp.flags[dp.rank] + t, e);
    }
    __syncthreads();
    if (t == 0) {
      double s = 0.0;
      for (int r = 0; r < NP; ++r) s += *reinterpret_cast<const volatile double*>(dp.loss[r]);
      *dp.loss_out = s;
      st_release_gpu(dp.ctl + 1, e);   // lets the other blocks of this grid go
    }
  } else {
    if (t == 0) {
      uint32_t spins = 0;
      while (ld_acquire_gpu_u32(dp.ctl + 1) < e) {
        __nanosleep(64);
        if (++spins > (1u << 24)) asm volatile("trap;");
      }
    }
  }
  __syncthreads();
  return e;
}

// exit barrier: the last block of the grid tells every rank "I am done reading your arena" and waits until every rank
// is done reading this one -- only then may the kernels that follow in this stream overwrite the gradients / loss
template <int NP>
__device__ __forceinline__ void dp_leave(const DpCtx& dp, uint32_t e) {
  __shared__ uint32_t s_last;
  const int t = threadIdx.x;
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const uint32_t prev = atomicAdd(dp.ctl + 2, 1u);
    s_last = (prev == gridDim.x - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last == 0u) return;
  if (t < NP) {
    st_release_sys(dp.flags[t] + kDpMaxRanks + dp.rank, e);
    spin_until_sys(dp.flags[dp.rank] + kDpMaxRanks + t, e);
  }
  __syncthreads();
  if (t == 0) {
    dp.ctl[2] = 0u;
    st_release_gpu(dp.ctl, e);
  }
}

template <int NP, bool TWO>
__global__ void __launch_bounds__(192)
adamw_pack_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  const __grid_constant__ FusedBatch fb, const __grid_constant__ DpCtx dp, float lr, float b1, float b2,
                  float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
  __shared__ float sm[8][72];   // [oo][ii*9 + tap]
  const int t = threadIdx.x;
  uint32_t epoch = 0;
  if constexpr (NP > 0) epoch = dp_enter<NP>(dp);
  if constexpr (TWO) dp_reduce_slice<NP>(dp, epoch);
  if (static_cast<int>(blockIdx.x) >= fb.brick_blocks) {   // biases, head conv, anything without a packed operand
    const int row = (static_cast<int>(blockIdx.x) - fb.brick_blocks) * kFusedRowsPerBlock + t / kFusedRow;
    const int i = t % kFusedRow;
    if (row < fb.nrows && i < fb.row_cnt[row])
      adamw_one<NP, TWO>(p, g, dp, m, v, fb.row_off[row] + i, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    if constexpr (NP > 0) dp_leave<NP>(dp, epoch);
    return;
  }
  for (int brick = blockIdx.x; brick < fb.nbricks; brick += fb.brick_blocks) {
    int c = 0;
    while (c + 1 < fb.nconv && fb.conv[c + 1].brick0 <= brick) ++c;
    const FusedConv& cv = fb.conv[c];
    const int ibricks = cv.I / 8;
    const int local = brick - cv.brick0;
    const int ob = local / ibricks, ib = local - ob * ibricks;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int e = t + 192 * r;                 // 0..575
      const int oo = e / 72, rem = e - oo * 72;  // rem = ii*9 + tap: 72 contiguous floats of output channel 8*ob+oo
      const long long idx = cv.w_off + (static_cast<long long>(8 * ob + oo) * cv.I + 8 * ib) * 9 + rem;
      sm[oo][rem] = adamw_one<NP, TWO>(p, g, dp, m, v, idx, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    }
    __syncthreads();
    const int s = ib / 6, chunk = ib - 6 * s;    // source slice and 8-channel chunk of the brick's input channels
    if (t < 72) {
      // forward operand: for tap, output channel 8*ob+oo: the 8 input channels of the brick
      const int tap = t / 8, oo = t % 8;
      uint4 q;
      q.x = pack_bf16x2(sm[oo][0 * 9 + tap], sm[oo][1 * 9 + tap]);
      q.y = pack_bf16x2(sm[oo][2 * 9 + tap], sm[oo][3 * 9 + tap]);
      q.z = pack_bf16x2(sm[oo][4 * 9 + tap], sm[oo][5 * 9 + tap]);
      q.w = pack_bf16x2(sm[oo][6 * 9 + tap], sm[oo][7 * 9 + tap]);
      const size_t off = ((static_cast<size_t>(s * 9 + tap) * 6 + chunk) * 48 + 8 * ob + oo) * 8;
      *reinterpret_cast<uint4*>(cv.fwd + off) = q;
    } else if (t < 144) {
      // backward-data operand of slice s: tap rotated, "output" channel = input channel 8*chunk+ii, the 8 = o%8
      const int u = t - 72, tap = u / 8, ii = u % 8;
      uint4 q;
      q.x = pack_bf16x2(sm[0][ii * 9 + tap], sm[1][ii * 9 + tap]);
      q.y = pack_bf16x2(sm[2][ii * 9 + tap], sm[3][ii * 9 + tap]);
      q.z = pack_bf16x2(sm[4][ii * 9 + tap], sm[5][ii * 9 + tap]);
      q.w = pack_bf16x2(sm[6][ii * 9 + tap], sm[7][ii * 9 + tap]);
      const size_t off = ((static_cast<size_t>(8 - tap) * 6 + ob) * 48 + 8 * chunk + ii) * 8;
      *reinterpret_cast<uint4*>(cv.bwd[s] + off) = q;
    }
  }
  if constexpr (NP > 0) dp_leave<NP>(dp, epoch);
}

static int build_fused_batch(FusedBatch& fb, long long numel, const lv_fused_conv* convs, int nconv) {
  LV_CHECK_ARG(convs != nullptr && nconv >= 1 && nconv <= kFusedMaxConvs, "adamw+pack: 1..%d convs per call", kFusedMaxConvs);
  // convs must be sorted by offset and disjoint; everything between them becomes a plain range
  long long pos = 0;
  int nr = 0, bricks = 0;
  for (int k = 0; k < nconv; ++k) {
    const lv_fused_conv& c = convs[k];
    LV_CHECK_ARG(c.cout == 48 && c.cin_total % 48 == 0 && c.cin_total >= 48 && c.cin_total <= 48 * LV_MAX_SRC,
                 "adamw+pack: conv %d must be 48 x (48*k) x 3 x 3", k);
    LV_CHECK_ARG(c.w_off >= pos && c.w_off + 48ll * c.cin_total * 9 <= numel, "adamw+pack: conv %d out of order / range", k);
    LV_CHECK_ARG(c.fwd != nullptr, "adamw+pack: conv %d has no forward operand", k);
    for (long long q = pos; q < c.w_off; q += kFusedRow) {
      LV_CHECK_ARG(nr < kFusedMaxRows, "adamw+pack: too many parameters outside the packed convs");
      fb.row_off[nr] = q;
      fb.row_cnt[nr] = static_cast<short>(c.w_off - q < kFusedRow ? c.w_off - q : kFusedRow);
      ++nr;
    }
    fb.conv[k].w_off = c.w_off;
    fb.conv[k].I = c.cin_total;
    fb.conv[k].brick0 = bricks;
    fb.conv[k].fwd = static_cast<__nv_bfloat16*>(c.fwd);
    for (int sidx = 0; sidx < LV_MAX_SRC; ++sidx) {
      fb.conv[k].bwd[sidx] = static_cast<__nv_bfloat16*>(c.bwd[sidx]);
      LV_CHECK_ARG(sidx >= c.cin_total / 48 || c.bwd[sidx] != nullptr, "adamw+pack: conv %d lacks a backward operand", k);
    }
    bricks += 6 * (c.cin_total / 8);
    pos = c.w_off + 48ll * c.cin_total * 9;
  }
  for (long long q = pos; q < numel; q += kFusedRow) {
    LV_CHECK_ARG(nr < kFusedMaxRows, "adamw+pack: too many parameters outside the packed convs");
    fb.row_off[nr] = q;
    fb.row_cnt[nr] = static_cast<short>(numel - q < kFusedRow ? numel - q : kFusedRow);
    ++nr;
  }
  fb.nconv = nconv; fb.nrows = nr; fb.nbricks = bricks;
  fb.brick_blocks = bricks < 148 * 8 ? bricks : 148 * 8;
  return LV_OK;
}

int adamw_pack_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long numel, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int step, float grad_scale, const lv_fused_conv* convs,
                    int nconv, cudaStream_t stream) {
  if (numel == 0) return LV_OK;
  LV_CHECK_ARG(step >= 1, "adamw: step must be >= 1");
  static thread_local FusedBatch fb;
  static thread_local DpCtx none;   // unused by the single-process instantiation
  int rc = build_fused_batch(fb, numel, convs, nconv);
  if (rc != LV_OK) return rc;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  const int tail_blocks = (fb.nrows + kFusedRowsPerBlock - 1) / kFusedRowsPerBlock;
  adamw_pack_kernel<0, false><<<fb.brick_blocks + tail_blocks, 192, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, fb, none, lr, beta1, beta2,
                                                                       eps, weight_decay, static_cast<float>(bc1),
                                                                       static_cast<float>(sqrt(bc2)), grad_scale);
  LV_LAUNCH_OK();
  return LV_OK;
}

// Data-parallel optimizer step: gradient all-reduce over peer memory + AdamW + operand re-pack in ONE kernel (see DpCtx).
int dp_adamw_pack_step(float* param, float* exp_avg, float* exp_avg_sq, long long numel, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, float grad_scale, const lv_fused_conv* convs, int nconv,
                       const void* const* peer_grads, void* const* peer_reduced, void* const* peer_flags,
                       const void* const* peer_loss, double* loss_out, uint32_t* ctl, long long slice, int world, int rank,
                       cudaStream_t stream) {
  if (numel == 0) return LV_OK;
  LV_CHECK_ARG(step >= 1, "adamw: step must be >= 1");
  LV_CHECK_ARG(world == 2 || world == 4 || world == 8, "dp adamw: world size must be 2, 4 or 8 (got %d)", world);
  LV_CHECK_ARG(rank >= 0 && rank < world, "dp adamw: rank %d outside 0..%d", rank, world - 1);
  LV_CHECK_ARG(peer_grads && peer_flags && peer_loss && loss_out && ctl, "dp adamw: null pointer");
  const bool two = peer_reduced != nullptr && slice > 0;
  if (two) LV_CHECK_ARG(slice % 4 == 0 && slice * world >= numel && numel < (1ll << 31), "dp adamw: bad slice %lld", slice);
  static thread_local FusedBatch fb;
  static thread_local DpCtx dp;
  int rc = build_fused_batch(fb, numel, convs, nconv);
  if (rc != LV_OK) return rc;
  for (int r = 0; r < world; ++r) {
    LV_CHECK_ARG(peer_grads[r] && peer_flags[r] && peer_loss[r] && (!two || peer_reduced[r]), "dp adamw: null pointer for rank %d", r);
    dp.grad[r] = static_cast<const float*>(peer_grads[r]);
    dp.reduced[r] = two ? static_cast<float*>(peer_reduced[r]) : nullptr;
    dp.flags[r] = static_cast<uint32_t*>(peer_flags[r]);
    dp.loss[r] = static_cast<const double*>(peer_loss[r]);
  }
  dp.loss_out = loss_out;
  dp.ctl = ctl;
  dp.slice = slice;
  dp.world = world;
  dp.rank = rank;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  if (two) {
    // the middle barrier is a grid barrier: every block must be resident at once (192 threads, ~56 registers: at least 4
    // blocks per SM)
    const int cap = 4 * sm_count();
    if (fb.brick_blocks > cap - 64) fb.brick_blocks = cap - 64;
  }
  const int tail_blocks = (fb.nrows + kFusedRowsPerBlock - 1) / kFusedRowsPerBlock;
  const int grid = fb.brick_blocks + tail_blocks;
  if (two) LV_CHECK_ARG(grid <= 4 * sm_count(), "dp adamw: %d blocks cannot be co-resident", grid);
  const float* g = dp.grad[rank];
#define LV_DP_LAUNCH(NPV, TWOV)                                                                                           \
  adamw_pack_kernel<NPV, TWOV><<<grid, 192, 0, stream>>>(param, g, exp_avg, exp_avg_sq, fb, dp, lr, beta1, beta2, eps,     \
                                                         weight_decay, static_cast<float>(bc1),                           \
                                                         static_cast<float>(sqrt(bc2)), grad_scale)
  if (world == 2 && !two) LV_DP_LAUNCH(2, false);
  else if (world == 2) LV_DP_LAUNCH(2, true);
  else if (world == 4 && !two) LV_DP_LAUNCH(4, false);
  else if (world == 4) LV_DP_LAUNCH(4, true);
  else if (world == 8 && !two) LV_DP_LAUNCH(8, false);
  else LV_DP_LAUNCH(8, true);
#undef LV_DP_LAUNCH
  LV_LAUNCH_OK();
  return LV_OK;
}

}  // namespace lv
