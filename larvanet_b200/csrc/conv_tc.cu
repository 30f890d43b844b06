// 3x3 convolution as an implicit GEMM on the sm_100a tensor cores (tcgen05.mma, fp32 accumulators in TMEM).
//
//   D[128 pixels x NT couts] = sum over (source s, tap t, 16-channel K step k)  A_{s,t,k}[128 x 16] * B_{s,t,k}[16 x NT]
//
// Tile: 16 rows x 8 columns of output pixels (M = 128).  The (16+2) x (8+2) input halo tile of one source is
// staged in shared memory ONCE, channel-chunk-planar: [chunk of 8 channels][halo pixel][16 B].  That is exactly the
// SWIZZLE_NONE K-major UMMA operand layout with "8-row group" = one tile row, so the A operand of tap (ky,kx) is
// the same buffer with the descriptor start address moved by (ky*10+kx)*16 bytes and SBO = 160 B (halo row pitch):
// no im2col, no per-tap copies, zero padding comes from zero-filled halo pixels.
//
// Weights (bf16, pre-packed as [ntile][src][tap][chunk][cout][8]) stay resident in shared memory for the whole
// persistent CTA; they arrive with one TMA bulk copy per (src,tap) block.
//
// Warp roles (288 threads): warps 0-3 epilogue (one TMEM lane quarter each), warp 4 TMEM alloc + single-thread
// MMA issue, warps 5-8 halo-tile producers (cp.async 16 B, zero-fill outside the image).  Two TMEM accumulator
// stages let the epilogue of tile i overlap the MMAs of tile i+1; NSTAGE halo stages decouple loads from MMAs.
#include "conv_epilogue.cuh"
#include "lv_common.cuh"

namespace lv {

constexpr int kTileH = 16, kTileW = 8;
constexpr int kHaloW = kTileW + 2, kHaloH = kTileH + 2, kHaloPix = kHaloW * kHaloH;  // 10 x 18 = 180
constexpr int kEpiThreads = 128, kProdThreads = 128;
constexpr int kTcThreads = kEpiThreads + 32 + kProdThreads;  // 288

template <int CIN, int NT, int NSTAGE>
struct TcCfg {
  static constexpr int CH = CIN / 8;                    // 16-byte channel chunks per pixel
  static constexpr int KSTEPS = CIN / 16;               // UMMA K steps per tap
  static constexpr int A_PLANE = kHaloPix * 16;         // bytes of one chunk plane
  static constexpr int A_STAGE = CH * A_PLANE;
  static constexpr int W_TAP = CH * NT * 16;            // bytes of one (src,tap) weight block
  static constexpr int ACC_STRIDE = (NT <= 64) ? 64 : 128;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int LAG = (NSTAGE >= 3) ? 2 : 1;     // cp.async groups kept in flight per producer thread
  static size_t smem_bytes(int num_src) {
    return static_cast<size_t>(num_src) * 9 * W_TAP + static_cast<size_t>(NSTAGE) * A_STAGE + 256;
  }
};

template <int CIN, int NT, int NSTAGE>
__global__ void __launch_bounds__(kTcThreads, (CIN == 48 && NT == 48) ? 2 : 1)
conv3x3_tc_kernel(const __grid_constant__ lv_conv_args a, const ConvGeom g) {
  using Cfg = TcCfg<CIN, NT, NSTAGE>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t w_bytes = static_cast<uint32_t>(a.num_src) * 9u * Cfg::W_TAP;
  uint8_t* sW = smem;
  uint8_t* sA = smem + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + NSTAGE * Cfg::A_STAGE);
  // bars: [0,NSTAGE) full, [NSTAGE,2NSTAGE) empty, then tmem_full[2], tmem_empty[2], wbar
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 + s); };
  const uint32_t wbar = bar0 + 8u * (2 * NSTAGE + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 5);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), kProdThreads);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiThreads);
    }
    mbar_init(wbar, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc<Cfg::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int ntile = static_cast<int>(blockIdx.x % g.ntiles_n);  // fixed per CTA (grid is a multiple of ntiles_n)
  const int tiles_per_img = g.tiles_x * g.tiles_y;

  if (warp >= 5) {
    // =============================== producers: halo tiles -> smem ===============================
    const int ptid = threadIdx.x - (kEpiThreads + 32);
    uint32_t fill = 0;      // running (tile, source) counter
    uint32_t arrived = 0;   // fills already signalled on their full barrier
    for (long long tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const long long pt = tile / g.ntiles_n;
      const int n = static_cast<int>(pt / tiles_per_img);
      const int rem = static_cast<int>(pt % tiles_per_img);
      const int y0 = (rem / g.tiles_x) * kTileH - 1;
      const int x0 = (rem % g.tiles_x) * kTileW - 1;
      for (int s = 0; s < a.num_src; ++s, ++fill) {
        const int stage = fill % NSTAGE;
        mbar_wait(empty_bar(stage), ((fill / NSTAGE) & 1) ^ 1);
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.src[s]);
        const uint32_t dst0 = smem_u32(sA + stage * Cfg::A_STAGE);
#pragma unroll 4
        for (int idx = ptid; idx < kHaloPix * Cfg::CH; idx += kProdThreads) {
          const int p = idx / Cfg::CH, c = idx % Cfg::CH;
          const int r = p / kHaloW, col = p % kHaloW;
          const int gy = y0 + r, gx = x0 + col;
          const bool inb = (gy >= 0) && (gy < a.h) && (gx >= 0) && (gx < a.w);
          const size_t off = inb ? ((static_cast<size_t>(n) * a.h + gy) * a.w + gx) * CIN + c * 8 : 0;
          cp_async16(dst0 + c * Cfg::A_PLANE + p * 16, src + off, inb ? 16u : 0u);
        }
        cp_async_commit();
        if (fill >= static_cast<uint32_t>(Cfg::LAG)) {
          cp_async_wait<Cfg::LAG>();
          fence_proxy_async_smem();
          mbar_arrive(full_bar(arrived % NSTAGE));
          ++arrived;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (; arrived < fill; ++arrived) mbar_arrive(full_bar(arrived % NSTAGE));
  } else if (warp == 4) {
    // =============================== MMA issuer (one elected thread) ==============================
    if (lane == 0) {
      // resident weights: one bulk copy per (src,tap) block
      mbar_arrive_expect_tx(wbar, w_bytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.weights) + static_cast<size_t>(ntile) * w_bytes;
      for (int b = 0; b < a.num_src * 9; ++b)
        tma_bulk_g2s(smem_u32(sW + b * Cfg::W_TAP), wsrc + static_cast<size_t>(b) * Cfg::W_TAP, Cfg::W_TAP, wbar);
      mbar_wait(wbar, 0);

      constexpr uint32_t idesc = umma_idesc_bf16(128, NT, 0, 0);
      const uint32_t sW_addr = smem_u32(sW);
      uint32_t fill = 0, k = 0;
      for (long long tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++k) {
        const uint32_t as = k & 1;
        mbar_wait(tempty_bar(as), ((k >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
        uint32_t accumulate = 0;
        for (int s = 0; s < a.num_src; ++s, ++fill) {
          const int stage = fill % NSTAGE;
          mbar_wait(full_bar(stage), (fill / NSTAGE) & 1);
          tc_fence_after_sync();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_STAGE);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_tap = a_addr + ((tap / 3) * kHaloW + (tap % 3)) * 16;
            const uint32_t b_tap = sW_addr + (s * 9 + tap) * Cfg::W_TAP;
#pragma unroll
            for (int ks = 0; ks < Cfg::KSTEPS; ++ks) {
              const uint64_t adesc = umma_smem_desc(a_tap + 2 * ks * Cfg::A_PLANE, Cfg::A_PLANE, kHaloW * 16);
              const uint64_t bdesc = umma_smem_desc(b_tap + 2 * ks * (NT * 16), NT * 16, 128);
              umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(empty_bar(stage));  // halo stage reusable once these MMAs retire
        }
        umma_commit(tfull_bar(as));       // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue: TMEM -> registers -> global ========================
    const int m = threadIdx.x;            // TMEM lane == tile pixel index
    const int r = m >> 3, c = m & 7;
    float loss = 0.f;
    uint32_t k = 0;
    for (long long tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++k) {
      const long long pt = tile / g.ntiles_n;
      const int n = static_cast<int>(pt / tiles_per_img);
      const int rem = static_cast<int>(pt % tiles_per_img);
      const int y = (rem / g.tiles_x) * kTileH + r;
      const int x = (rem % g.tiles_x) * kTileW + c;
      const bool valid = (y < a.h) && (x < a.w);
      const uint32_t as = k & 1;
      mbar_wait(tfull_bar(as), (k >> 1) & 1);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + as * Cfg::ACC_STRIDE;
      if constexpr (NT <= 64) {
        // read the whole accumulator first so the TMEM stage is released before the global traffic
        float v[NT];
#pragma unroll
        for (int j = 0; j < NT / 16; ++j) tmem_ld16(taddr + j * 16, v + j * 16);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(tempty_bar(as));
        if (valid) {
#pragma unroll
          for (int j = 0; j < NT / 16; ++j)
            loss += conv_epilogue16<__nv_bfloat16>(a, n, y, x, ntile * NT + j * 16, v + j * 16);
        }
      } else {
#pragma unroll 1
        for (int j = 0; j < NT / 16; ++j) {
          float v[16];
          tmem_ld16(taddr + j * 16, v);
          tmem_ld_wait();
          if (valid) loss += conv_epilogue16<__nv_bfloat16>(a, n, y, x, ntile * NT + j * 16, v);
        }
        tc_fence_before_sync();
        mbar_arrive(tempty_bar(as));
      }
    }
    if (a.truth_hr != nullptr && a.loss_sum != nullptr) {
      loss = warp_sum(loss);
      if (lane == 0) atomicAdd(a.loss_sum, static_cast<double>(loss));
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after_sync();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
int pick_ntile(int cout_pad) {
  int nt = cout_pad <= 128 ? cout_pad : 128;
  while (cout_pad % nt != 0 || nt % 16 != 0) nt -= 16;
  return nt;
}

template <int CIN, int NT, int NSTAGE>
static int launch_tc(const lv_conv_args& a, const ConvGeom& g, int max_ctas, cudaStream_t stream) {
  using Cfg = TcCfg<CIN, NT, NSTAGE>;
  const size_t smem = Cfg::smem_bytes(a.num_src);
  auto kern = conv3x3_tc_kernel<CIN, NT, NSTAGE>;
  static size_t configured = 0;
  if (smem > configured) {
    LV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  long long ctas = max_ctas > 0 ? max_ctas : sm_count();
  if (ctas > g.total_tiles) ctas = g.total_tiles;
  ctas = (ctas / g.ntiles_n) * g.ntiles_n;
  if (ctas < g.ntiles_n) ctas = g.ntiles_n;
  kern<<<static_cast<unsigned>(ctas), kTcThreads, smem, stream>>>(a, g);
  LV_LAUNCH_OK();
  return LV_OK;
}

int conv3x3_tc(const lv_conv_args& a, int max_ctas, cudaStream_t stream) {
  ConvGeom g;
  g.cout_pad = (a.cout + 15) / 16 * 16;
  g.nt = pick_ntile(g.cout_pad);
  g.ntiles_n = g.cout_pad / g.nt;
  g.tiles_x = (a.w + kTileW - 1) / kTileW;
  g.tiles_y = (a.h + kTileH - 1) / kTileH;
  g.total_tiles = static_cast<long long>(a.n) * g.tiles_x * g.tiles_y * g.ntiles_n;
  if (g.total_tiles == 0) return LV_OK;
#define LV_TC_CASE(CI, NTV, NS) \
  if (a.cin == CI && g.nt == NTV) return launch_tc<CI, NTV, NS>(a, g, max_ctas, stream);
  if (a.num_src <= 2) {
    LV_TC_CASE(48, 48, 4)
  } else {
    LV_TC_CASE(48, 48, 3)
  }
  LV_TC_CASE(48, 96, 3)
  LV_TC_CASE(64, 64, 3)
  LV_TC_CASE(64, 128, 3)
  LV_TC_CASE(64, 16, 3)
  LV_TC_CASE(48, 16, 3)
#undef LV_TC_CASE
  set_error("conv3x3 tensor-core path: unsupported shape cin=%d cout=%d (n-tile %d)", a.cin, a.cout, g.nt);
  return LV_ERR_INVALID;
}

}  // namespace lv
