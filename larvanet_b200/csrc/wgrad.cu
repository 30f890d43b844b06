// Weight (and bias) gradients of the 3x3 convs:  dW[co][ci][ky][kx] = sum_px dY[px][co] * X[px + (ky-1,kx-1)][ci]
//
// Tensor-core path (bf16, cin == 48, cout in {16,32,48,64}).  Per 16x8 pixel tile two TMA tensor-map boxes land in
// shared memory:  the X halo tile {10 px, 6 chunks, 18 rows} as [row][chunk][10 px x 16 B] and the dY tile
// {8 px, cout/8 chunks, 16 rows} as [row][chunk][8 px x 16 B]; zero padding is the TMA unit's out-of-bounds fill.
// Read as MN-major SWIZZLE_NONE UMMA operands (K = 16 pixels = two tile rows):
//   * B = dY:  N = cout, chunk stride 128 B, K-group (tile row) stride cout/8 * 128 B;
//   * A = X with the vertical taps STACKED ALONG M.  Chunk-plane stride is 160 B and a halo row holds exactly 6 planes
//     (960 B), so plane index m of the descriptor addresses (row + m/6, chunk m%6): M = 128 = 16 planes covers
//     (ky=0, 48 ch), (ky=1, 48 ch), (ky=2, ch 0..31) of one horizontal tap kx (the descriptor start address moves by
//     kx*16 B), and a second MMA of M = 64 starting at plane 16 the remaining (ky=2, ch 32..47).
// An M=64 MMA costs as much as an M=128 one (~40-45 clk at N=48), so this is 6 MMAs per 16 pixels instead of 9 for the
// nine taps (+1 against an all-ones A tile for the bias gradient).  Accumulators: D[(ky,ci)][(kx,co)] in 336 TMEM
// columns, accumulated across the tiles of one layer.  Work split (stream-K style): the tile jobs of all layers of a launch
// form one list that is cut into equal contiguous ranges, one per CTA (<= one CTA per SM), so every SM gets the same
// number of tiles however many layers there are; a CTA whose range crosses a layer boundary drains its accumulators to
// its next workspace slot and starts over.  A second kernel sums, per layer, the slots of the CTAs that touched it (fixed
// order: deterministic) into the fp32 OIHW gradient.  Up to 64 layers per launch; descriptors and the job schedule
// travel as kernel parameters.
//
// CUDA-core path (fp32 validation mode, and bf16 cross-check in tests): register-tiled direct accumulation.
#include <cuda.h>

#include <unordered_map>

#include "lv_common.cuh"

namespace lv {

// ============================================================================================
// tensor-core path
// ============================================================================================
constexpr int kWT_H = 16, kWT_W = 8, kWHaloW = 10, kWHaloH = 18;
constexpr int kWCin = 48, kWCh = kWCin / 8;
constexpr int kWStages = 3;
constexpr int kWThreads = 192;               // 4 TMEM-drain warps, 1 MMA warp, 1 TMA loader warp
constexpr int kWMaxItems = 64;               // tensor maps per launch (2 x 64 x 128 B kernel parameter)
constexpr int kXPlaneStride = kWHaloW * 16;  // 160 B: next 8-channel chunk of the same halo row
constexpr int kXRowPitch = kWCh * kXPlaneStride;   // 960 B: next halo row == 6 planes further
constexpr int kXBytes = kWHaloH * kXRowPitch;      // 17280
constexpr int kDyBytesMax = 16 * 8 * 128;          // 16384 (cout 64)
constexpr int kWStageBytes = kDyBytesMax + kXBytes + 128;   // 33792 (keeps every stage 128 B aligned)
constexpr int kOnesBytes = 2048;             // bf16 1.0 everywhere: the A tile of the bias MMA
// the M=64 MMA's unused planes reach up to 5 rows past the X tile of the last stage
constexpr int kWSmem = kWStages * kWStageBytes + kOnesBytes + 6 * kXRowPitch + 256 /*barriers*/;
// accumulator columns: [kx*cout + co] from the M=128 MMAs, then the same for the M=64 MMAs, then the bias block
constexpr int kWLanes = 128;
__host__ __device__ constexpr int wcols(int cout) { return 7 * cout; }
constexpr int kWColsMax = 7 * 64;            // 448 <= 512 TMEM columns

struct WMaps {
  CUtensorMap dy[kWMaxItems];   // planar-8 dY [n][h][cout/8][w][8], box {8 px, cout/8 chunks, 16 rows}
  CUtensorMap x[kWMaxItems];    // planar-8 X  [n][h][6][w][8],      box {10 px, 6 chunks, 18 rows}
};

struct WSched {
  long long jobs;                 // total tile jobs of the launch
  int count, smax;                // layers, workspace slots per CTA
  int prefix[kWMaxItems + 1];     // first job of every layer
  // which workspace slots hold a layer's partial sums: CTAs cta_lo .. cta_lo + ncta - 1; the layer is segment `seg0` of
  // the first of them and segment 0 of the others (their ranges start inside the layer)
  short cta_lo[kWMaxItems], ncta[kWMaxItems], seg0[kWMaxItems];
};
// first job of CTA b out of G
__host__ __device__ inline long long wstart(long long jobs, int b, int G) { return jobs * b / G; }
__host__ __device__ inline int witem(const WSched& sc, long long j) {
  int i = 0;
  while (i + 1 < sc.count && sc.prefix[i + 1] <= j) ++i;
  return i;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(kWThreads, 1)
wgrad_tc_kernel(const lv_wgrad_item* __restrict__ items, const __grid_constant__ WMaps maps, const __grid_constant__ WSched sc,
                float* __restrict__ workspace) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = static_cast<int>(gridDim.x);
  const long long j0 = wstart(sc.jobs, blockIdx.x, G), j1 = wstart(sc.jobs, blockIdx.x + 1, G);
  const int item0 = witem(sc, j0);

  uint8_t* sOnes = smem + kWStages * kWStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + kOnesBytes + 6 * kXRowPitch);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kWStages + s); };
  const uint32_t done_bar = bar0 + 8u * (2 * kWStages);        // MMAs of a segment retired -> drain warps
  const uint32_t drained_bar = bar0 + 8u * (2 * kWStages + 1); // accumulators read out -> MMA warp may overwrite them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWStages + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    mbar_init(drained_bar, 128);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < kOnesBytes / 4; i += kWThreads) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
  if (warp == 4) tmem_alloc<512>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // every role walks the same segments: (layer `item`, its tiles [t0, t1)) for consecutive layers from item0
  if (warp >= 5) {
    // ------------------------------- loader: two TMA boxes per tile -------------------------------
    if (elect_one()) {
      uint32_t fill = 0;
      int item = item0;
      for (long long j = j0; j < j1; ++item) {
        const long long seg_end = (sc.prefix[item + 1] < j1) ? sc.prefix[item + 1] : j1;
        const lv_wgrad_item it = items[item];
        const int tiles_x = (it.w + kWT_W - 1) / kWT_W, tiles_y = (it.h + kWT_H - 1) / kWT_H;
        const int tiles_per_img = tiles_x * tiles_y;
        const uint32_t bytes = static_cast<uint32_t>(it.cout / 8) * 16u * 128u + kXBytes;
        for (int tile = static_cast<int>(j - sc.prefix[item]), t1 = static_cast<int>(seg_end - sc.prefix[item]); tile < t1;
             ++tile, ++fill) {
          const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
          const int ty = rem / tiles_x;
          const int y0 = ty * kWT_H, x0 = (rem - ty * tiles_x) * kWT_W;
          const int stage = fill % kWStages;
          mbar_wait_relaxed(empty_bar(stage), ((fill / kWStages) & 1) ^ 1);
          const uint32_t st0 = smem_u32(smem + stage * kWStageBytes);
          mbar_arrive_expect_tx(full_bar(stage), bytes);
          tma_load_4d(st0, &maps.dy[item], x0 * 8, 0, y0, n, full_bar(stage));
          tma_load_4d(st0 + kDyBytesMax, &maps.x[item], (x0 - 1) * 8, 0, y0 - 1, n, full_bar(stage));
        }
        j = seg_end;
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ------------------------------- MMA issuer -------------------------------
    if (elect_one()) {
      const uint32_t ones_addr = smem_u32(sOnes);
      uint32_t fill = 0;
      int item = item0, seg = 0;
      for (long long j = j0; j < j1; ++item, ++seg) {
        const long long seg_end = (sc.prefix[item + 1] < j1) ? sc.prefix[item + 1] : j1;
        const int cout = items[item].cout;
        const uint32_t idesc_128 = umma_idesc_bf16(128, cout, 1, 1);
        const uint32_t idesc_64 = umma_idesc_bf16(64, cout, 1, 1);
        const uint32_t dy_row = static_cast<uint32_t>(cout / 8) * 128u;     // bytes of one dY tile row (all chunks)
        if (seg > 0) {   // the previous layer's accumulators must have been read out
          mbar_wait(drained_bar, (seg - 1) & 1);
          tc_fence_after_sync();
        }
        const int ntile = static_cast<int>(seg_end - j);
        for (int q = 0; q < ntile; ++q, ++fill) {
          const int stage = fill % kWStages;
          mbar_wait(full_bar(stage), (fill / kWStages) & 1);
          tc_fence_after_sync();
          const uint32_t dy0 = smem_u32(smem + stage * kWStageBytes);
          const uint32_t xs0 = dy0 + kDyBytesMax;
#pragma unroll 1
          for (int k8 = 0; k8 < 8; ++k8) {       // 16 pixels = tile rows 2*k8, 2*k8+1
            const uint32_t acc = (q > 0 || k8 > 0) ? 1u : 0u;
            // B = dY: N chunks 128 B apart, K groups (tile rows) one dY row apart
            const uint64_t bdesc = umma_smem_desc(dy0 + (2 * k8) * dy_row, /*LBO: next K group*/ dy_row, /*SBO: next N chunk*/ 128);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              // A = X: plane m -> (halo row 2*k8 + m/6, chunk m%6), pixels kx .. kx+7 of the row
              const uint32_t a0 = xs0 + (2 * k8) * kXRowPitch + kx * 16;
              umma_bf16(tmem_base + kx * cout, umma_smem_desc(a0, kXRowPitch, kXPlaneStride), bdesc, idesc_128, acc);
              umma_bf16(tmem_base + (3 + kx) * cout, umma_smem_desc(a0 + 16 * kXPlaneStride, kXRowPitch, kXPlaneStride),
                        bdesc, idesc_64, acc);
            }
            umma_bf16(tmem_base + 6 * cout, umma_smem_desc(ones_addr, 128, 256), bdesc, idesc_64, acc);
          }
          umma_commit(empty_bar(stage));
        }
        umma_commit(done_bar);
        j = seg_end;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------- epilogue: TMEM -> workspace slot [col][lane 128] -------------------------------
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int l128 = warp * 32 + lane;
    int item = item0, seg = 0;
    for (long long j = j0; j < j1; ++item, ++seg) {
      const long long seg_end = (sc.prefix[item + 1] < j1) ? sc.prefix[item + 1] : j1;
      const int cout = items[item].cout;
      const int ncols = wcols(cout);
      float* ws = workspace + (static_cast<size_t>(blockIdx.x) * sc.smax + seg) * (static_cast<size_t>(kWColsMax) * kWLanes);
      mbar_wait_relaxed(done_bar, seg & 1);
      tc_fence_after_sync();
      // columns [0, 3*cout): all 128 lanes ((ky, ci) rows of the M=128 MMAs); columns [3*cout, 7*cout): only TMEM lanes
      // 0..15 carry data (rows 0..15 of an M=64 accumulator), i.e. warp 0
      for (int c = 0; c < ncols; c += 16) {
        if (c >= 3 * cout && warp != 0) break;
        float v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        if (c < 3 * cout || lane < 16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) ws[static_cast<size_t>(c + i) * kWLanes + l128] = v[i];
        }
      }
      tc_fence_before_sync();
      mbar_arrive(drained_bar);
      j = seg_end;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

// workspace [cta][slot][col][lane 128] -> dw / db (+=).  One block per (layer, output channel): sums the slots of the
// CTAs whose job range touched the layer (ascending CTA order: deterministic); for each kx the 128 lanes of column
// kx*cout+co are contiguous, permuted through shared memory to OIHW's (ci, ky, kx) order and added to dw as 432
// contiguous floats.
__global__ void __launch_bounds__(128)
wgrad_reduce_kernel(const lv_wgrad_item* __restrict__ items, const __grid_constant__ WSched sc, const float* __restrict__ workspace,
                    int G) {
  (void)G;
  __shared__ float sm[9 * kWCin];
  const int item = blockIdx.y;
  const lv_wgrad_item it = items[item];
  const int co = blockIdx.x;
  const long long jb = sc.prefix[item], je = sc.prefix[item + 1];
  if (co >= it.cout || je <= jb) return;
  const int cout = it.cout;
  const size_t slot = static_cast<size_t>(kWColsMax) * kWLanes;
  const int t = threadIdx.x;
  const int nslot = sc.ncta[item];
  const float* wsb = workspace + (static_cast<size_t>(sc.cta_lo[item]) * sc.smax + sc.seg0[item]) * slot;
  const size_t step = static_cast<size_t>(sc.smax) * slot;
  const size_t c0 = static_cast<size_t>(0 * cout + co) * kWLanes + t, c1 = static_cast<size_t>(1 * cout + co) * kWLanes + t,
               c2 = static_cast<size_t>(2 * cout + co) * kWLanes + t;
  const size_t ce = static_cast<size_t>((3 + (t % 48) / 16) * cout + co) * kWLanes + (t % 16);
  const size_t cb = static_cast<size_t>(6 * cout + co) * kWLanes;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, e = 0.f, bsum = 0.f;
#pragma unroll 4
  for (int k = 0; k < nslot; ++k) {   // ascending CTA order: deterministic
    // slot of CTA cta_lo + k: the first one at segment seg0, the others at segment 0
    const float* ws = (k == 0) ? wsb : wsb + k * step - static_cast<size_t>(sc.seg0[item]) * slot;
    a0 += ws[c0];
    a1 += ws[c1];
    a2 += ws[c2];
    if (t < 48) e += ws[ce];
    if (t == 64) bsum += ws[cb];
  }
  {  // rows t = ky*48 + ci of the M=128 accumulators (ky 0, 1 and ky 2 / ci < 32)
    const int ky = t / kWCin, ci = t % kWCin;
    sm[ci * 9 + ky * 3 + 0] = a0 * it.scale;
    sm[ci * 9 + ky * 3 + 1] = a1 * it.scale;
    sm[ci * 9 + ky * 3 + 2] = a2 * it.scale;
  }
  if (t < 48) sm[(32 + t % 16) * 9 + 6 + t / 16] = e * it.scale;   // rows 0..15 of the M=64 accumulators: ky 2, ci 32..47
  if (t == 64 && it.db != nullptr) it.db[co] = bsum * it.scale + (it.overwrite ? 0.f : it.db[co]);
  __syncthreads();
  float* dst = it.dw + (static_cast<size_t>(co) * it.cin_total + it.cin_off) * 9;
  if (it.overwrite) {
    for (int j = t; j < 9 * kWCin; j += 128) dst[j] = sm[j];
  } else {
    for (int j = t; j < 9 * kWCin; j += 128) dst[j] += sm[j];
  }
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
    if (items[k].dtype != LV_BF16 || items[k].cin != kWCin || items[k].cout % 16 != 0 || items[k].cout > 64) return false;
  return true;
}

static long long item_tiles(const lv_wgrad_item& a) {
  return static_cast<long long>(a.n) * ((a.h + kWT_H - 1) / kWT_H) * ((a.w + kWT_W - 1) / kWT_W);
}

// job schedule of items[0..count) (count <= kWMaxItems) on at most `max_ctas` CTAs; returns the grid size (0: no work)
static int make_sched(const lv_wgrad_item* items, int count, long long max_ctas, WSched* sc) {
  sc->count = count;
  long long j = 0;
  for (int k = 0; k < count; ++k) {
    sc->prefix[k] = static_cast<int>(j);
    j += item_tiles(items[k]);
  }
  for (int k = count; k <= kWMaxItems; ++k) sc->prefix[k] = static_cast<int>(j);
  sc->jobs = j;
  if (j == 0) { sc->smax = 1; return 0; }
  long long G = max_ctas < sm_count() ? max_ctas : sm_count();
  if (G > 512) G = 512;   // wgrad_reduce_kernel's slot list
  if (G > j) G = j;
  if (G < 1) G = 1;
  int smax = 1;
  for (int b = 0; b < G; ++b) {
    const long long s0 = wstart(j, b, static_cast<int>(G)), s1 = wstart(j, b + 1, static_cast<int>(G));
    const int n = witem(*sc, s1 - 1) - witem(*sc, s0) + 1;
    if (n > smax) smax = n;
  }
  sc->smax = smax;
  for (int k = 0; k < count; ++k) {
    const long long jb = sc->prefix[k], je = sc->prefix[k + 1];
    sc->cta_lo[k] = sc->ncta[k] = sc->seg0[k] = 0;
    if (je <= jb) continue;
    // CTA owning job x: the largest b with jobs*b/G <= x
    auto owner = [&](long long x) { return static_cast<int>(((x + 1) * G + j - 1) / j - 1); };
    const int lo = owner(jb), hi = owner(je - 1);
    sc->cta_lo[k] = static_cast<short>(lo);
    sc->ncta[k] = static_cast<short>(hi - lo + 1);
    sc->seg0[k] = static_cast<short>(k - witem(*sc, wstart(j, lo, static_cast<int>(G))));
  }
  return static_cast<int>(G);
}

long long wgrad_workspace_bytes(const lv_wgrad_item* items, int count, int splits) {
  if (count <= 0 || splits <= 0 || !tc_eligible(items, count)) return 0;
  long long total = 0;
  WSched sc;
  for (int first = 0; first < count; first += kWMaxItems) {
    const int cnt = (count - first < kWMaxItems) ? count - first : kWMaxItems;
    for (int k = 0; k < cnt; ++k)
      if (item_tiles(items[first + k]) >= (1ll << 31) / kWMaxItems) return -1;
    const int G = make_sched(items + first, cnt, static_cast<long long>(splits) * cnt, &sc);
    total += static_cast<long long>(G) * sc.smax * kWColsMax * kWLanes * sizeof(float);
  }
  return total > 16 ? total : 16;
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
  static bool configured[64] = {false};   // per device
  const int dev = current_device_slot();
  if (!configured[dev]) {
    LV_CUDA_OK(cudaFuncSetAttribute(wgrad_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    LV_CUDA_OK(cudaFuncSetAttribute(wgrad_simt_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured[dev] = true;
  }
  if (items_host[0].dtype == LV_F32)
    wgrad_simt_kernel<float><<<dim3(splits, count), 256, smem, stream>>>(items_dev, splits);
  else
    wgrad_simt_kernel<__nv_bfloat16><<<dim3(splits, count), 256, smem, stream>>>(items_dev, splits);
  LV_LAUNCH_OK();
  return LV_OK;
}

// ---- host: TMA descriptors (cached per tensor) ----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// box {`px` pixels, all chunks, `rows` rows} of a planar-8 bf16 activation tensor [n][h][chunks][w][8]
int activation_tile_map(const void* src, int n, int h, int w, int chunks, int px, int rows, CUtensorMap* out) {
  struct Key {
    const void* p; int n, h, w, chunks, px, rows;
    bool operator==(const Key& o) const {
      return p == o.p && n == o.n && h == o.h && w == o.w && chunks == o.chunks && px == o.px && rows == o.rows;
    }
  };
  struct Hash {
    size_t operator()(const Key& k) const {
      return std::hash<const void*>()(k.p) ^ (static_cast<size_t>(k.n) * 1000003u) ^ (static_cast<size_t>(k.h) << 20) ^
             (static_cast<size_t>(k.w) << 8) ^ (static_cast<size_t>(k.chunks) << 4) ^ static_cast<size_t>(k.rows + 64 * k.px);
    }
  };
  static thread_local std::unordered_map<Key, CUtensorMap, Hash> cache;
  const Key key{src, n, h, w, chunks, px, rows};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return LV_OK; }
  EncodeTiledFn enc = encode_fn();
  LV_CHECK_ARG(enc != nullptr, "tensor map: cuTensorMapEncodeTiled is not available from this driver");
  LV_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15u) == 0, "tensor map: activation pointers must be 16-byte aligned");
  const cuuint64_t gdim[4] = {static_cast<cuuint64_t>(w) * 8, static_cast<cuuint64_t>(chunks), static_cast<cuuint64_t>(h),
                              static_cast<cuuint64_t>(n)};
  const cuuint64_t gstr[3] = {static_cast<cuuint64_t>(w) * 16, static_cast<cuuint64_t>(w) * 16 * chunks,
                              static_cast<cuuint64_t>(w) * 16 * chunks * h};
  const cuuint32_t box[4] = {static_cast<cuuint32_t>(px) * 8, static_cast<cuuint32_t>(chunks), static_cast<cuuint32_t>(rows), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tm;
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(src), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LV_CHECK_ARG(r == CUDA_SUCCESS, "tensor map: cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d x %d-chunk tensor",
               static_cast<int>(r), n, h, w, chunks);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, tm);
  *out = tm;
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
  static bool configured[64] = {false};   // per device
  const int dev = current_device_slot();
  if (!configured[dev]) {
    LV_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmem));
    configured[dev] = true;
  }
  static thread_local WMaps maps;   // staging only; the launches copy them by value
  static thread_local WSched sc;
  const size_t slot = static_cast<size_t>(kWColsMax) * kWLanes;
  float* ws = static_cast<float*>(workspace);
  for (int first = 0; first < count; first += kWMaxItems) {
    const int cnt = (count - first < kWMaxItems) ? count - first : kWMaxItems;
    for (int k = 0; k < cnt; ++k) {
      const lv_wgrad_item& a = items_host[first + k];
      LV_CHECK_ARG(item_tiles(a) < (1ll << 31) / kWMaxItems, "wgrad: item %d has too many tiles", first + k);
      if (item_tiles(a) == 0) continue;
      rc = activation_tile_map(a.dy, a.n, a.h, a.w, a.cout / 8, kWT_W, kWT_H, &maps.dy[k]);
      if (rc != LV_OK) return rc;
      rc = activation_tile_map(a.x, a.n, a.h, a.w, kWCh, kWHaloW, kWHaloH, &maps.x[k]);
      if (rc != LV_OK) return rc;
    }
    const int G = make_sched(items_host + first, cnt, static_cast<long long>(splits) * cnt, &sc);
    if (G == 0) continue;
    wgrad_tc_kernel<<<G, kWThreads, kWSmem, stream>>>(items_dev + first, maps, sc, ws);
    LV_LAUNCH_OK();
    wgrad_reduce_kernel<<<dim3(64, cnt), 128, 0, stream>>>(items_dev + first, sc, ws, G);
    LV_LAUNCH_OK();
    ws += static_cast<size_t>(G) * sc.smax * slot;
  }
  return LV_OK;
}

}  // namespace lv
