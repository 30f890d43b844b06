// Weight (and bias) gradients of the 3x3 convs:  dW[co][ci][ky][kx] = sum_px dY[px][co] * X[px + (ky-1,kx-1)][ci]
//
// Tensor-core path (bf16, cin == 48, cout <= 64): per 16x8 pixel tile the dY tile and the X halo tile are staged
// channel-chunk-planar ([chunk][pixel][16 B]) exactly like conv_tc.cu; read as MN-major SWIZZLE_NONE UMMA operands
// that same layout gives  A = dY^T (M = cout padded to 64, K = 16 pixels = two tile rows) and, for every tap, the
// shifted view  B = X[. + tap] (N = 48, same K) -- so wgrad needs no transposes and no im2col either.  The nine tap
// products accumulate in nine 48-column TMEM accumulators (432 of 512 columns) across ALL tiles of the CTA; a tenth
// 8-column accumulator against an all-ones B tile yields the bias gradient.  Split-K over pixel tiles: each
// (item, split) CTA writes its partial to a workspace, a second kernel reduces deterministically and accumulates into
// the fp32 OIHW gradient.  Batched: one launch covers `count` layers (grid.y).
//
// CUDA-core path (fp32 validation mode, and bf16 cross-check in tests): register-tiled direct accumulation.
#include "lv_common.cuh"

namespace lv {

// ============================================================================================
// tensor-core path
// ============================================================================================
constexpr int kWT_H = 16, kWT_W = 8, kWHaloW = 10, kWHaloPix = 180;
constexpr int kWCin = 48;
constexpr int kWCols = 9 * kWCin + 8;       // 440 accumulator columns (9 taps x 48 + bias)
constexpr int kWStages = 3;
constexpr int kWThreads = 288;
constexpr int kDyPlane = 128 * 16;          // 2048 B
constexpr int kXPlane = kWHaloPix * 16;     // 2880 B
constexpr int kDyBytes = 8 * kDyPlane;      // reserve 8 chunk planes so M=64 never reads outside the stage
constexpr int kXBytes = (kWCin / 8) * kXPlane;
constexpr int kWStageBytes = kDyBytes + kXBytes;  // 16384 + 17280 = 33664
constexpr int kWSmem = kWStages * kWStageBytes + 256 /*ones tile*/ + 256 /*barriers*/;

__global__ void __launch_bounds__(kWThreads, 1)
wgrad_tc_kernel(const lv_wgrad_item* __restrict__ items, float* __restrict__ workspace, int splits) {
  extern __shared__ __align__(128) uint8_t smem[];
  const lv_wgrad_item it = items[blockIdx.y];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int tiles_x = (it.w + kWT_W - 1) / kWT_W, tiles_y = (it.h + kWT_H - 1) / kWT_H;
  const int tiles_per_img = tiles_x * tiles_y;
  const int total_tiles = it.n * tiles_per_img;
  const int cho = it.cout / 8;

  uint8_t* sOnes = smem + kWStages * kWStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + 256);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kWStages + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * kWStages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWStages + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(full_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 128) {
    reinterpret_cast<uint16_t*>(sOnes)[threadIdx.x] = 0x3F80u;  // bf16 1.0: [16 pixels][8] MN-major ones tile
  }
  if (warp == 4) tmem_alloc<512>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 5) {
    // ------------------------------- producers -------------------------------
    const int ptid = threadIdx.x - 160;
    const __nv_bfloat16* X = reinterpret_cast<const __nv_bfloat16*>(it.x);
    const __nv_bfloat16* DY = reinterpret_cast<const __nv_bfloat16*>(it.dy);
    // tile-invariant piece tables: dY tile pieces first (128 px x cout/8 chunks), then X halo pieces (180 px x 6)
    constexpr int kMaxPieces = (128 * 8 + kWHaloPix * (kWCin / 8) + 127) / 128;  // 17 for cout = 64
    uint32_t pc_dst[kMaxPieces];
    int pc_rel[kMaxPieces], pc_rc[kMaxPieces];
    const int n_dy = 128 * cho, n_all = n_dy + kWHaloPix * (kWCin / 8);
#pragma unroll
    for (int i = 0; i < kMaxPieces; ++i) {
      const int idx = ptid + i * 128;
      if (idx < n_dy) {
        const int p = idx / cho, c = idx - p * cho;
        pc_dst[i] = c * kDyPlane + p * 16;
        pc_rel[i] = (((p >> 3) * cho + c) * it.w + (p & 7)) * 8;
        pc_rc[i] = ((p >> 3) << 8) | (p & 7);                       // bit 30 clear: dY piece (tile coordinates)
      } else if (idx < n_all) {
        const int j = idx - n_dy;
        const int p = j / (kWCin / 8), c = j - p * (kWCin / 8);
        const int r = p / kWHaloW, col = p - r * kWHaloW;
        pc_dst[i] = kDyBytes + c * kXPlane + p * 16;
        pc_rel[i] = (((r - 1) * (kWCin / 8) + c) * it.w + (col - 1)) * 8;
        pc_rc[i] = (1 << 30) | (r << 8) | col;                      // bit 30 set: X halo piece (halo coordinates)
      } else {
        pc_dst[i] = 0; pc_rel[i] = 0; pc_rc[i] = -1;
      }
    }
    uint32_t fill = 0;
    for (int tile = split; tile < total_tiles; tile += splits, ++fill) {
      const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
      const int ty = rem / tiles_x;
      const int y0 = ty * kWT_H, x0 = (rem - ty * tiles_x) * kWT_W;
      const int stage = fill % kWStages;
      mbar_wait_relaxed(empty_bar(stage), ((fill / kWStages) & 1) ^ 1);
      const uint32_t st0 = smem_u32(smem + stage * kWStageBytes);
      // planar-8 layout: element offset of (n, y0, chunk 0, x0)
      const __nv_bfloat16* dy_org = DY + ((static_cast<long long>(n) * it.h + y0) * cho * it.w + x0) * 8;
      const __nv_bfloat16* x_org = X + ((static_cast<long long>(n) * it.h + y0) * (kWCin / 8) * it.w + x0) * 8;
#pragma unroll
      for (int i = 0; i < kMaxPieces; ++i) {
        if (pc_rc[i] >= 0) {
          const bool isx = (pc_rc[i] >> 30) != 0;
          const int rr = (pc_rc[i] >> 8) & 0xff, cc = pc_rc[i] & 0xff;
          const int gy = y0 + rr - (isx ? 1 : 0), gx = x0 + cc - (isx ? 1 : 0);
          const bool inb = (static_cast<unsigned>(gy) < static_cast<unsigned>(it.h)) &&
                           (static_cast<unsigned>(gx) < static_cast<unsigned>(it.w));
          const __nv_bfloat16* src = isx ? x_org : dy_org;
          cp_async16(st0 + pc_dst[i], inb ? (src + pc_rel[i]) : X, inb ? 16u : 0u);
        }
      }
      cp_async_mbar_arrive_noinc(full_bar(stage));
    }
    cp_async_wait<0>();
  } else if (warp == 4) {
    // ------------------------------- MMA issuer -------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_w = umma_idesc_bf16(64, kWCin, 1, 1);
      constexpr uint32_t idesc_b = umma_idesc_bf16(64, 8, 1, 1);
      const uint32_t ones_addr = smem_u32(sOnes);
      uint32_t fill = 0;
      for (int tile = split; tile < total_tiles; tile += splits, ++fill) {
        const int stage = fill % kWStages;
        mbar_wait(full_bar(stage), (fill / kWStages) & 1);
        fence_proxy_async_smem();   // consumer-side: cp.async (generic proxy) writes -> UMMA (async proxy) reads
        tc_fence_after_sync();
        const uint32_t dy0 = smem_u32(smem + stage * kWStageBytes);
        const uint32_t xs0 = dy0 + kDyBytes;
#pragma unroll 1
        for (int k8 = 0; k8 < 8; ++k8) {       // 16 pixels = tile rows 2*k8, 2*k8+1
          const uint32_t acc = (fill > 0 || k8 > 0) ? 1u : 0u;
          const uint64_t adesc = umma_smem_desc(dy0 + (2 * k8) * 128, /*LBO: next K group*/ 128, /*SBO: next M chunk*/ kDyPlane);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t b_addr = xs0 + ((2 * k8 + tap / 3) * kWHaloW + tap % 3) * 16;
            const uint64_t bdesc = umma_smem_desc(b_addr, kWHaloW * 16, kXPlane);
            umma_bf16(tmem_base + tap * kWCin, adesc, bdesc, idesc_w, acc);
          }
          const uint64_t odesc = umma_smem_desc(ones_addr, 128, 256);
          umma_bf16(tmem_base + 9 * kWCin, adesc, odesc, idesc_b, acc);
        }
        umma_commit(empty_bar(stage));
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    // ------------------------------- epilogue: TMEM -> workspace -------------------------------
    mbar_wait_relaxed(done_bar, 0);
    tc_fence_after_sync();
    // M = 64 accumulator layout: row r lives in TMEM lane 32*(r/16) + r%16
    const int co = warp * 16 + lane;
    const bool has = (lane < 16) && (co < it.cout);
    float* ws = workspace + (static_cast<size_t>(blockIdx.y) * splits + split) * (static_cast<size_t>(kWCols) * 64);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const bool any_tiles = split < total_tiles;
#pragma unroll 1
    for (int j = 0; j < kWCols / 8; j += 2) {   // 55 groups of 8 columns; read 16 at a time (last read covers 8)
      float v[16];
      if (j + 1 < kWCols / 8) {
        tmem_ld16(taddr + j * 8, v);
      } else {
        tmem_ld16(taddr + (j - 1) * 8, v);      // overlap the previous 8 columns; keep the upper half
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i + 8];
      }
      tmem_ld_wait();
      const int ncol = (j + 1 < kWCols / 8) ? 16 : 8;
      if (has) {
        for (int i = 0; i < ncol; ++i) ws[static_cast<size_t>(j * 8 + i) * 64 + co] = any_tiles ? v[i] : 0.f;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// workspace [item][split][col 440][co 64] -> dw / db (+=).  One block per (item, group of 8 output channels): the
// split sums are read as 32-byte runs of 8 channels per column, permuted through shared memory from the accumulator's
// (tap, ci) column order to OIHW's (ci, tap), and added to dw as 432 contiguous floats per output channel.
constexpr int kRedCo = 8;
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const lv_wgrad_item* __restrict__ items, const float* __restrict__ workspace, int splits) {
  __shared__ float sm[kRedCo][9 * kWCin + 1];
  const lv_wgrad_item it = items[blockIdx.y];
  const int co0 = blockIdx.x * kRedCo;
  if (co0 >= it.cout) return;
  const float* ws0 = workspace + static_cast<size_t>(blockIdx.y) * splits * (static_cast<size_t>(kWCols) * 64);
  constexpr int kCols = 9 * kWCin + 1;   // 432 weight columns + the first bias column
  for (int idx = threadIdx.x; idx < kCols * kRedCo; idx += 256) {
    const int col = idx / kRedCo, c = idx % kRedCo;
    const float* ws = ws0 + static_cast<size_t>(col) * 64 + co0 + c;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += ws[static_cast<size_t>(k) * kWCols * 64];   // fixed order: deterministic
    s *= it.scale;
    // column (tap, ci) -> OIHW offset ci*9 + tap inside the channel's row; the bias column keeps index 432
    const int dst = (col < 9 * kWCin) ? (col % kWCin) * 9 + col / kWCin : col;
    sm[c][dst] = s;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 9 * kWCin * kRedCo; idx += 256) {
    const int c = idx / (9 * kWCin), j = idx % (9 * kWCin);
    const int co = co0 + c;
    if (co < it.cout) {
      float* p = it.dw + (static_cast<size_t>(co) * it.cin_total + it.cin_off) * 9 + j;
      *p += sm[c][j];
    }
  }
  if (threadIdx.x < kRedCo && it.db != nullptr && co0 + threadIdx.x < it.cout) it.db[co0 + threadIdx.x] += sm[threadIdx.x][9 * kWCin];
}

// ============================================================================================
// CUDA-core path
// ============================================================================================
constexpr int kSW_H = 8, kSW_W = 16, kSWHaloW = 18, kSWHaloPix = 180;
constexpr int kPP = 9;  // (co,ci) pairs per thread per pass

template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const lv_wgrad_item* __restrict__ items, int splits) {
  extern __shared__ float sm[];
  const lv_wgrad_item it = items[blockIdx.y];
  const int cin = it.cin, cout = it.cout, cinp = cin + 1, cop = cout + 1;
  float* sx = sm;                       // [180][cin+1]
  float* sdy = sm + kSWHaloPix * cinp;  // [128][cout+1]
  const int tiles_x = (it.w + kSW_W - 1) / kSW_W, tiles_y = (it.h + kSW_H - 1) / kSW_H;
  const int tiles_per_img = tiles_x * tiles_y, total_tiles = it.n * tiles_per_img;
  const int npairs = cin * cout;
  const T* X = reinterpret_cast<const T*>(it.x);
  const T* DY = reinterpret_cast<const T*>(it.dy);

  for (int pbase = 0; pbase < npairs; pbase += 256 * kPP) {
    float acc[kPP][9];
    float bacc[kPP];
#pragma unroll
    for (int q = 0; q < kPP; ++q) {
      bacc[q] = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[q][t] = 0.f;
    }
    for (int tile = blockIdx.x; tile < total_tiles; tile += splits) {
      const int n = tile / tiles_per_img, rem = tile % tiles_per_img;
      const int y0 = (rem / tiles_x) * kSW_H, x0 = (rem % tiles_x) * kSW_W;
      __syncthreads();
      for (int idx = threadIdx.x; idx < kSWHaloPix * cin; idx += 256) {
        const int hp = idx / cin, ci = idx % cin;
        const int gy = y0 - 1 + hp / kSWHaloW, gx = x0 - 1 + hp % kSWHaloW;
        float v = 0.f;
        if (gy >= 0 && gy < it.h && gx >= 0 && gx < it.w) v = to_f32(X[act_off(n, gy, gx, ci >> 3, it.h, it.w, cin >> 3) + (ci & 7)]);
        sx[hp * cinp + ci] = v;
      }
      for (int idx = threadIdx.x; idx < 128 * cout; idx += 256) {
        const int p = idx / cout, co = idx % cout;
        const int gy = y0 + p / kSW_W, gx = x0 + p % kSW_W;
        float v = 0.f;
        if (gy < it.h && gx < it.w) v = to_f32(DY[act_off(n, gy, gx, co >> 3, it.h, it.w, cout >> 3) + (co & 7)]);
        sdy[p * cop + co] = v;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < kPP; ++q) {
        const int pair = pbase + q * 256 + threadIdx.x;
        if (pair < npairs) {
          const int co = pair % cout, ci = pair / cout;
          for (int p = 0; p < 128; ++p) {
            const float d = sdy[p * cop + co];
            const float* xp = sx + ((p / kSW_W) * kSWHaloW + (p % kSW_W)) * cinp + ci;
            if (ci == 0) bacc[q] += d;
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[q][t] = fmaf(d, xp[((t / 3) * kSWHaloW + (t % 3)) * cinp], acc[q][t]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < kPP; ++q) {
      const int pair = pbase + q * 256 + threadIdx.x;
      if (pair < npairs) {
        const int co = pair % cout, ci = pair / cout;
        float* p = it.dw + (static_cast<size_t>(co) * it.cin_total + it.cin_off + ci) * 9;
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(p + t, it.scale * acc[q][t]);
        if (ci == 0 && it.db != nullptr) atomicAdd(it.db + co, it.scale * bacc[q]);
      }
    }
  }
}

// ============================================================================================
// host side
// ============================================================================================
static int check_items(const lv_wgrad_item* items, int count) {
  LV_CHECK_ARG(items != nullptr && count > 0, "wgrad: empty item list");
  LV_CHECK_ARG(count <= 65535, "wgrad: too many items");
  for (int k = 0; k < count; ++k) {
    const lv_wgrad_item& a = items[k];
    LV_CHECK_ARG(a.x && a.dy && a.dw, "wgrad: null pointer in item %d", k);
    LV_CHECK_ARG(a.dtype == items[0].dtype, "wgrad: mixed dtypes in one batch");
    LV_CHECK_ARG(a.cin > 0 && a.cout > 0 && a.cin <= 64 && a.cout <= 64 && a.cin % 8 == 0 && a.cout % 8 == 0,
                 "wgrad: cin/cout must be multiples of 8 in 8..64 (item %d)", k);
    LV_CHECK_ARG(a.cin_off >= 0 && a.cin_off + a.cin <= a.cin_total, "wgrad: bad cin slice (item %d)", k);
  }
  return LV_OK;
}

static bool tc_eligible(const lv_wgrad_item* items, int count) {
  for (int k = 0; k < count; ++k)
    if (items[k].dtype != LV_BF16 || items[k].cin != kWCin || items[k].cout % 8 != 0) return false;
  return true;
}

long long wgrad_workspace_bytes(const lv_wgrad_item* items, int count, int splits) {
  if (count <= 0 || splits <= 0 || !tc_eligible(items, count)) return 0;
  return static_cast<long long>(count) * splits * kWCols * 64 * sizeof(float);
}

int wgrad_simt(const lv_wgrad_item* items_host, const lv_wgrad_item* items_dev, int count, int splits, cudaStream_t stream) {
  int rc = check_items(items_host, count);
  if (rc != LV_OK) return rc;
  LV_CHECK_ARG(splits > 0, "wgrad: splits must be > 0");
  int cin = 0, cout = 0;
  for (int k = 0; k < count; ++k) {
    cin = items_host[k].cin > cin ? items_host[k].cin : cin;
    cout = items_host[k].cout > cout ? items_host[k].cout : cout;
  }
  const size_t smem = (static_cast<size_t>(kSWHaloPix) * (cin + 1) + 128 * (cout + 1)) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    LV_CUDA_OK(cudaFuncSetAttribute(wgrad_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    LV_CUDA_OK(cudaFuncSetAttribute(wgrad_simt_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  if (items_host[0].dtype == LV_F32)
    wgrad_simt_kernel<float><<<dim3(splits, count), 256, smem, stream>>>(items_dev, splits);
  else
    wgrad_simt_kernel<__nv_bfloat16><<<dim3(splits, count), 256, smem, stream>>>(items_dev, splits);
  LV_LAUNCH_OK();
  return LV_OK;
}

int wgrad(const lv_wgrad_item* items_host, const lv_wgrad_item* items_dev, int count, int splits, void* workspace,
          cudaStream_t stream) {
  int rc = check_items(items_host, count);
  if (rc != LV_OK) return rc;
  LV_CHECK_ARG(splits > 0 && splits <= 65535, "wgrad: splits must be in 1..65535");
  if (!tc_eligible(items_host, count)) {
    LV_CHECK_ARG(items_host[0].dtype == LV_F32, "wgrad: the bf16 tensor-core path needs cin == 48 and cout %% 8 == 0");
    return wgrad_simt(items_host, items_dev, count, splits, stream);
  }
  LV_CHECK_ARG(workspace != nullptr, "wgrad: workspace required");
  static bool configured = false;
  if (!configured) {
    LV_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem));
    configured = true;
  }
  wgrad_tc_kernel<<<dim3(splits, count), kWThreads, kWSmem, stream>>>(items_dev, static_cast<float*>(workspace), splits);
  LV_LAUNCH_OK();
  wgrad_reduce_kernel<<<dim3(64 / kRedCo, count), 256, 0, stream>>>(items_dev, static_cast<const float*>(workspace),
                                                                                  splits);
  LV_LAUNCH_OK();
  return LV_OK;
}

}  // namespace lv
