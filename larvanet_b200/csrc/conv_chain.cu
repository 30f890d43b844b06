// A whole chain of 3x3 convolutions (the residual-block bodies and early-exit legs of LarvaNet, forward or input-gradient)
// in ONE persistent kernel.
//
// On the shapes LarvaNet is trained and served on (16 x 48x48 patches, one 320x180 frame) a conv layer is only 2-3 tiles
// per SM, so a per-layer launch is dominated by what surrounds the math: launch + prologue (barriers, TMEM, 41 KB of
// weights), pipeline fill and drain (measured 6-8 us per layer for ~2 us of tensor work).  Here the CTAs stay resident
// for all layers and the layers are chained by data flow instead of kernel boundaries:
//
//   * every layer has the same tile grid (16x8 output pixels, conv_tc.cu's implicit GEMM); done[tile] counts finished
//     (layer, tile) jobs, 4 per job.  The epilogue warps arrive on a CTA-scope mbarrier after their stores and move on; a
//     dedicated publisher warp turns "all 4 warps arrived" into one red.release.gpu, which is what waits for the SM's
//     store acknowledgements (1.1-1.5k clk);
//   * the job (layer l, tile X) may start when all 9 tiles around X have finished every earlier layer (done >= 4*l;
//     the halo producer warp polls with relaxed loads and finishes with one ld.acquire.gpu, then fetches the 18x10 halo
//     tile with ONE tensor-map TMA copy whose out-of-bounds fill is the conv's zero padding).  That one rule covers the
//     halo reads, the same-tile residual / mask reads of the epilogue, and the write-after-read hazards of ping-pong
//     activation buffers;
//   * the halo / MMA / epilogue pipeline never drains between layers: the producers run ahead into layer l+1 while
//     the epilogue of layer l is still in flight, and the weights of layer l+1 are prefetched by TMA into the next of
//     three shared-memory weight buffers at the start of layer l;
//   * tile -> CTA assignment is rotated by (tiles mod grid) every layer, so the partial last wave is spread over all
//     CTAs instead of making the same few CTAs the critical path of every layer.
//
// Jobs are processed layer-major by every CTA and the grid never exceeds one CTA per SM, so the CTA owning the oldest
// unfinished job can always proceed (no cyclic waits).  The last CTA to leave re-zeroes the flags, so the workspace is
// ready for the next launch (also under CUDA-graph replay, where arguments are frozen).
//
// The layer descriptors travel as a kernel parameter (up to 96 x lv_conv_args = 17.7 KB in the constant bank).
#include <cuda.h>

#include "chain_epilogue.cuh"
#include "conv_epilogue.cuh"
#include "lv_common.cuh"

// LV_CHAIN_TMA=0 builds the cp.async halo producers of the first version (two warps, 17 x 16-byte copies per thread and
// tile) for A/B runs
#ifndef LV_CHAIN_TMA
#define LV_CHAIN_TMA 1
#endif

namespace lv {

// wgrad.cu: cached tensor map with box {px pixels, all chunks, rows rows} over a planar-8 bf16 tensor [n][h][chunks][w][8]
int activation_tile_map(const void* src, int n, int h, int w, int chunks, int px, int rows, CUtensorMap* out);

extern long long* g_timeline;
extern int g_use_pdl;

namespace chain {

constexpr int kMaxLayers = 96;   // LV_CHAIN_MAX_LAYERS
constexpr int kTileH = 16, kTileW = 8;
constexpr int kHaloW = kTileW + 2, kHaloH = kTileH + 2, kHaloPix = kHaloW * kHaloH;  // 10 x 18 = 180
constexpr bool kTmaHalo = LV_CHAIN_TMA != 0;
constexpr int kEpiWarps = 8, kEpiThreads = kEpiWarps * 32, kProdThreads = kTmaHalo ? 32 : 64;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kPubWarp = kMmaWarp + 1 + kProdThreads / 32;    // warp 10 (11): publishes finished tiles
constexpr int kThreads = kEpiThreads + 32 + kProdThreads + 32;  // 352 (384): <= 3 warps per scheduler, 168 registers per thread
constexpr int kWBufs = 3;
constexpr uint32_t kWarpsPerTile = 4;                      // epilogue warps per (layer, tile) = done[] increments

struct Params {
  lv_conv_args layer[kMaxLayers];
};
// one tensor map per layer: the layer's input, box = one halo tile {10 px, all chunks, 18 rows}; out-of-image pixels
// (the conv's zero padding) are filled in by the TMA unit
struct Maps {
  CUtensorMap src[kMaxLayers];
};
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

template <int CIN, int NT, int NSTAGE>
struct Cfg {
  static constexpr int CH = CIN / 8;
  static constexpr int KSTEPS = CIN / 16;
  // halo tile in shared memory: cp.async build [chunk][row][px] (planes), TMA build [row][chunk][px] (the box order)
  static constexpr int A_PLANE = kHaloPix * 16;
  static constexpr int A_STAGE = CH * A_PLANE;
  static constexpr int A_ROW = kTmaHalo ? CH * kHaloW * 16 : kHaloW * 16;   // bytes between tile rows of one chunk
  static constexpr int A_CHUNK = kTmaHalo ? kHaloW * 16 : A_PLANE;          // bytes between 8-channel chunks of one row
  static constexpr int W_TAP = CH * NT * 16;
  static constexpr int W_LAYER = 9 * W_TAP;
  static constexpr int ACC_STRIDE = 64;
  static constexpr int TMEM_COLS = 128;
  static constexpr int PROD_PIECES = (kHaloPix * CH + kProdThreads - 1) / kProdThreads;
  static constexpr size_t smem_bytes() { return static_cast<size_t>(kWBufs) * W_LAYER + static_cast<size_t>(NSTAGE) * A_STAGE + 512; }
};

// Measured on B200: ld.acquire.gpu polls + red.release.gpu publishes are markedly cheaper than relaxed accesses
// bracketed by explicit fence.acq_rel.gpu (6.1 vs 6.6 us per layer on a 320x180 frame), and handing the tile over
// through one CTA-level publisher (one fence per tile) is slower still: every extra intra-CTA hop costs more than the
// fences it saves, because the per-layer time is the latency of the dependency chain, not its throughput.  A release
// STORE of the new count (st.release.gpu; single writer per tile) instead of the release reduction is far slower too
// (6.6 vs 4.2 us per layer).
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// ld.acquire.gpu compiles to LD + CCTL.IVALL (invalidate the SM's whole L1) and that is where a polling warp's samples pile
// up in the ncu source view: poll with relaxed loads, then ONE acquire load of the satisfied flag (every writer is a
// release RMW, so whatever value it reads continues the release sequence).
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t need) {
  if (ld_acquire_gpu(p) >= need) return;      // common case: already there
  uint32_t spins = 0;
  while (ld_relaxed_gpu(p) < need) {
    __nanosleep(32);
    if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
  }
  (void)ld_acquire_gpu(p);
}
// Everything an epilogue thread needs that does not change from layer to layer.
struct EpiCtx {
  ConvGeom g;
  const uint32_t* done;
  volatile uint32_t* pub_seen;
  uint32_t tfull, tempty, pub_bar0, taddr;   // this group's accumulator-stage barriers, publish-barrier ring, TMEM address
  int G, H, W, tiles_per_img, r, c, lane, eg;
  bool tl0;
  size_t chunk_stride;
};

constexpr int kKindPs4 = 100, kKindGeneric = -1;

// All tiles of one layer for one epilogue thread.  KIND: 0/1/2/4/12 = straight-line planar epilogue with that flag
// set, kKindPs4 = PixelShuffle(4)+base(+loss), kKindGeneric = shared 16-channel routine.  The register-hungry kinds
// (two residuals, PixelShuffle) read the bias through L1 instead of holding 48 more registers across the tile loop.
template <int KIND, int NT>
__device__ __forceinline__ uint32_t run_layer(const EpiCtx& cx, const lv_conv_args& a, int l, int first, uint32_t k) {
  constexpr int NCH = NT / 8;
  const bool has_ops = (a.mask != nullptr) || (a.res1 != nullptr) || (a.res2 != nullptr);
  FastEpi fe;
  fe.mask = reinterpret_cast<const __nv_bfloat16*>(a.mask);
  fe.res1 = reinterpret_cast<const __nv_bfloat16*>(a.res1);
  fe.res2 = reinterpret_cast<const __nv_bfloat16*>(a.res2);
  fe.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  fe.res_scale = a.res_scale;
  fe.relu = a.relu;
  float loss = 0.f;
  constexpr bool kBiasRegs = (KIND == 0 || KIND == 1 || KIND == 2 || KIND == 4);
  float breg[kBiasRegs ? NT : 1];
  const float* bias_g = a.bias;
  if constexpr (kBiasRegs) {
#pragma unroll
    for (int i = 0; i < NT / 4; ++i) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(a.bias) + i);
      breg[4 * i] = b4.x; breg[4 * i + 1] = b4.y; breg[4 * i + 2] = b4.z; breg[4 * i + 3] = b4.w;
    }
  }
  for (int tile = first; tile < cx.g.total_tiles; tile += cx.G, ++k) {
    if ((k & 1u) != static_cast<uint32_t>(cx.eg)) continue;
    const int n = tile / cx.tiles_per_img;
    const int rem = tile - n * cx.tiles_per_img;
    const int tyi = rem / cx.g.tiles_x;
    const int y = tyi * kTileH + cx.r, x = (rem - tyi * cx.g.tiles_x) * kTileW + cx.c;
    const bool valid = (y < cx.H) && (x < cx.W);
    const size_t o0 = valid ? act_off(n, y, x, 0, cx.H, cx.W, NCH) : 0;
    // pub_bar ring safety: the publisher must have seen this barrier's previous use (job k-4); read the counter now,
    // test it after the stores
    const uint32_t seen = *cx.pub_seen;
    if (l > 0 && (KIND == kKindGeneric || has_ops)) {
      // same-tile operands come from earlier layers of this chain, possibly written by another CTA
      if (cx.lane == 0) wait_flag(cx.done + tile, kWarpsPerTile * static_cast<uint32_t>(l));
      __syncwarp();
    }
    if (cx.tl0) tl_stamp(cx.g, 2, k, 0);
    const uint32_t par = (k >> 1) & 1;
    if constexpr (KIND == kKindPs4) {
      loss += ps4_tile<NT>(a, bias_g, valid, n, y, x, cx.H, cx.W, o0, cx.chunk_stride, cx.taddr, cx.tfull, cx.tempty, par);
    } else if constexpr (KIND == kKindGeneric) {
      mbar_wait_relaxed(cx.tfull, par);
      tc_fence_after_sync();
#pragma unroll 1
      for (int j = 0; j < NT / 16; ++j) {
        float v[16];
        tmem_ld16(cx.taddr + j * 16, v);
        tmem_ld_wait();
        if (valid) loss += conv_epilogue16<__nv_bfloat16>(a, n, y, x, j * 16, v);
      }
      tc_fence_before_sync();
      mbar_arrive(cx.tempty);
    } else {
      fast_tile<KIND, NT, kBiasRegs>(fe, breg, bias_g, valid, o0, cx.chunk_stride, cx.taddr, cx.tfull, cx.tempty, par);
    }
    // this warp's quarter of (layer l, tile) is on its way to global memory: hand it to the publisher warp
    __syncwarp();
    if (cx.lane == 0) {
      if (k >= 4 && seen + 3u < k) {
        uint32_t spins = 0;
        while (*cx.pub_seen + 3u < k) {
          if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
        }
      }
      mbar_arrive(cx.pub_bar0 + 8u * (k & 3u));
    }
    if (cx.tl0) tl_stamp(cx.g, 2, k, 3);
  }
  if (a.truth_hr != nullptr && a.loss_sum != nullptr) {
    loss = warp_sum(loss);
    if (cx.lane == 0) atomicAdd(a.loss_sum, static_cast<double>(loss));
  }
  return k;
}

template <int CIN, int NT, int NSTAGE>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_chain_kernel(const __grid_constant__ Params P, const __grid_constant__ Maps M, const int nlayers, const ConvGeom g,
                     uint32_t* __restrict__ done, const int rot) {
  using C_ = Cfg<CIN, NT, NSTAGE>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint8_t* sW = smem;
  uint8_t* sA = sW + kWBufs * C_::W_LAYER;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + NSTAGE * C_::A_STAGE);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 + s); };
  auto wfull_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 4 + b); };
  auto wfree_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 4 + kWBufs + b); };
  auto pub_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 4 + 2 * kWBufs + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4 + 2 * kWBufs + 4);
  uint32_t* s_last = tmem_slot + 1;
  volatile uint32_t* pub_seen = tmem_slot + 2;   // jobs whose completion the publisher warp has observed

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), kTmaHalo ? 1 : kProdThreads);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiThreads / 2);
    }
    for (int b = 0; b < kWBufs; ++b) {
      mbar_init(wfull_bar(b), 1);
      mbar_init(wfree_bar(b), 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(pub_bar(s), kEpiWarps / 2);
    tmem_slot[2] = 0u;
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc<C_::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  const int G = static_cast<int>(gridDim.x);
  const int tiles_per_img = g.tiles_x * g.tiles_y;
  const int H = P.layer[0].h, W = P.layer[0].w;   // every layer of a chain has the same geometry (checked on the host)
  // first tile of this CTA in layer l: tile X of layer l belongs to CTA (X + l*rot) mod G
  auto first_tile = [&](int l) {
    const int sh = static_cast<int>((static_cast<long long>(l) * rot) % G);
    return (static_cast<int>(blockIdx.x) + G - sh) % G;
  };

  if (warp == kPubWarp) {
    // =============================== publisher: GPU-scope release of finished tiles ===============================
    // The epilogue warps arrive on a CTA-scope mbarrier after their stores and go on to the next tile; this warp turns
    // "all 4 warps of the tile arrived" into ONE red.release.gpu (which is what waits for the store acknowledgements).
    if (lane == 0) {
      int l = 0, tile = first_tile(0);
      uint32_t k = 0;
      auto advance = [&]() {
        tile += G;
        ++k;
        while (l < nlayers && tile >= g.total_tiles) {
          ++l;
          if (l < nlayers) tile = first_tile(l);
        }
      };
      while (l < nlayers) {
        mbar_wait(pub_bar(k & 3u), (k >> 2) & 1u);
        const int t0 = tile;
        advance();
        *pub_seen = k;
        tl_stamp(g, 3, k - 1, 0);
        red_release_gpu_add(done + t0, kWarpsPerTile);
        tl_stamp(g, 3, k - 1, 1);
      }
    }
    __syncwarp();
  } else if (warp > kMmaWarp) {
    // =============================== producers: dependency wait + halo tiles -> smem ===============================
    const int ptid = threadIdx.x - (kEpiThreads + 32);
    const int ddy = lane / 3 - 1, ddx = lane % 3 - 1;   // lanes 0..8 watch the 3x3 tile neighbourhood
    uint32_t fill = 0;
    if constexpr (kTmaHalo) {
      // ONE tensor-map copy per tile (box {10 px, all chunks, 18 rows}, zero fill outside the image) issued by one lane:
      // against 17 cp.async per thread on two warps it takes ~500-1,000 clk of LSU-path instructions per tile off the
      // producer, which was the pacing role of small chains (it could not run ahead of the MMAs)
      pdl_wait();
      for (int l = 0; l < nlayers; ++l) {
        for (int tile = first_tile(l); tile < g.total_tiles; tile += G, ++fill) {
          const int n = tile / tiles_per_img;
          const int rem = tile - n * tiles_per_img;
          const int ty = rem / g.tiles_x, tx = rem - ty * g.tiles_x;
          if (l > 0) {
            if (lane < 9) {
              const int yy = ty + ddy, xx = tx + ddx;
              if (yy >= 0 && yy < g.tiles_y && xx >= 0 && xx < g.tiles_x) {
                wait_flag(done + tile + ddy * g.tiles_x + ddx, kWarpsPerTile * static_cast<uint32_t>(l));
              }
            }
            __syncwarp();
          }
          if (lane == 0) {
            const int stage = fill % NSTAGE;
            tl_stamp(g, 0, fill, 0);
            mbar_wait_relaxed(empty_bar(stage), ((fill / NSTAGE) & 1) ^ 1);
            tl_stamp(g, 0, fill, 1);
            // the flags were acquired through the generic proxy; the copy below reads global memory through the async one
            // (the .global form: 140-500 clk; the unrestricted fence.proxy.async costs ~1,000 clk here)
            asm volatile("fence.proxy.async.global;" ::: "memory");
            mbar_arrive_expect_tx(full_bar(stage), C_::A_STAGE);
            tma_load_4d(smem_u32(sA + stage * C_::A_STAGE), &M.src[l], (tx * kTileW - 1) * 8, 0, ty * kTileH - 1, n, full_bar(stage));
            tl_stamp(g, 0, fill, 2);
          }
          __syncwarp();
        }
      }
    } else {
    uint32_t pc_dst[C_::PROD_PIECES];
    int pc_rel[C_::PROD_PIECES], pc_rc[C_::PROD_PIECES];
#pragma unroll
    for (int i = 0; i < C_::PROD_PIECES; ++i) {
      const int idx = ptid + i * kProdThreads;
      const int col = idx % kHaloW, rc = idx / kHaloW;
      const int c = rc % C_::CH, r = rc / C_::CH;
      pc_dst[i] = c * C_::A_PLANE + (r * kHaloW + col) * 16;
      pc_rel[i] = ((r * C_::CH + c) * W + col) * 8;
      pc_rc[i] = (idx < kHaloPix * C_::CH) ? ((r << 8) | col) : -1;
    }
    pdl_wait();
    for (int l = 0; l < nlayers; ++l) {
      const __nv_bfloat16* src_base = reinterpret_cast<const __nv_bfloat16*>(P.layer[l].src[0]);
      for (int tile = first_tile(l); tile < g.total_tiles; tile += G, ++fill) {
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int ty = rem / g.tiles_x, tx = rem - ty * g.tiles_x;
        if (l > 0) {
          // every producer warp watches the 3x3 neighbourhood itself (no intra-CTA hand-over on the critical path)
          if (lane < 9) {
            const int yy = ty + ddy, xx = tx + ddx;
            if (yy >= 0 && yy < g.tiles_y && xx >= 0 && xx < g.tiles_x) {
              wait_flag(done + tile + ddy * g.tiles_x + ddx, kWarpsPerTile * static_cast<uint32_t>(l));
            }
          }
          __syncwarp();
        }
        const int y0 = ty * kTileH - 1, x0 = tx * kTileW - 1;
        const long long origin = ((static_cast<long long>(n) * H + y0) * C_::CH * W + x0) * 8;
        const int stage = fill % NSTAGE;
        if (ptid == 0) tl_stamp(g, 0, fill, 0);
        mbar_wait_relaxed(empty_bar(stage), ((fill / NSTAGE) & 1) ^ 1);
        if (ptid == 0) tl_stamp(g, 0, fill, 1);
        const __nv_bfloat16* src = src_base + origin;
        const uint32_t dst0 = smem_u32(sA + stage * C_::A_STAGE);
#pragma unroll
        for (int i = 0; i < C_::PROD_PIECES; ++i) {
          if (pc_rc[i] >= 0) {
            const int gy = y0 + (pc_rc[i] >> 8), gx = x0 + (pc_rc[i] & 0xff);
            const bool inb = (static_cast<unsigned>(gy) < static_cast<unsigned>(H)) &&
                             (static_cast<unsigned>(gx) < static_cast<unsigned>(W));
            cp_async16(dst0 + pc_dst[i], inb ? (src + pc_rel[i]) : src_base, inb ? 16u : 0u);
          }
        }
        cp_async_mbar_arrive_noinc(full_bar(stage));
        if (ptid == 0) tl_stamp(g, 0, fill, 2);
      }
    }
    cp_async_wait<0>();
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer (one elected lane) ================================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, NT, 0, 0);
      auto load_weights = [&](int l) {
        const int b = l % kWBufs;
        mbar_arrive_expect_tx(wfull_bar(b), C_::W_LAYER);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.layer[l].weights);
        for (int t = 0; t < 9; ++t)
          tma_bulk_g2s(smem_u32(sW + b * C_::W_LAYER + t * C_::W_TAP), wsrc + static_cast<size_t>(t) * C_::W_TAP, C_::W_TAP,
                       wfull_bar(b));
      };
      load_weights(0);   // packed weights are never written while a launch chain is in flight: no pdl_wait needed
      uint32_t fill = 0;
      bool next_tempty = false, next_full = false;   // the next job's barriers were seen complete by an early probe
      uint32_t free_pending = 0, free_phase = 0;   // bit b: MMAs reading weight buffer b outstanding / wfree parity
      for (int l = 0; l < nlayers; ++l) {
        const int wb = l % kWBufs;
        if (l + 1 < nlayers) {
          const int nb = (l + 1) % kWBufs;
          if (free_pending & (1u << nb)) {   // last used by layer l-2: its MMAs retired long ago
            mbar_wait(wfree_bar(nb), (free_phase >> nb) & 1u);
            free_phase ^= 1u << nb;
            free_pending &= ~(1u << nb);
          }
          load_weights(l + 1);
        }
        mbar_wait(wfull_bar(wb), (l / kWBufs) & 1);
        const uint32_t sW_addr = smem_u32(sW + wb * C_::W_LAYER);
        bool any = false;
        for (int tile = first_tile(l); tile < g.total_tiles; tile += G, ++fill) {
          const uint32_t k = fill;
          const uint32_t as = k & 1;
          tl_stamp(g, 1, k, 0);
          // both barriers of this job were probed while the previous job's last MMAs were being issued (below): when
          // the pipeline is ahead, neither wait costs a shared-memory round trip here (~190 clk each under MMA load)
          if (!next_tempty) mbar_wait(tempty_bar(as), ((k >> 1) & 1) ^ 1);
          tl_stamp(g, 1, k, 1);
          const uint32_t d_tmem = tmem_base + as * C_::ACC_STRIDE;
          const int stage = fill % NSTAGE;
          if (!next_full) mbar_wait(full_bar(stage), (fill / NSTAGE) & 1);
          if (!kTmaHalo) fence_proxy_async_smem();   // cp.async (generic proxy) writes -> UMMA (async proxy) reads
          tc_fence_after_sync();
          tl_stamp(g, 1, k, 2);
          const uint32_t a_addr = smem_u32(sA + stage * C_::A_STAGE);
          const uint32_t nk = k + 1u, nstage = (fill + 1u) % NSTAGE;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_tap = a_addr + (tap / 3) * C_::A_ROW + (tap % 3) * 16;
            const uint32_t b_tap = sW_addr + tap * C_::W_TAP;
            if (tap == 7) {
              // the tensor queue is full from here on: the probes' round trips hide behind the remaining issue stalls
              next_tempty = mbar_try_wait(tempty_bar(nk & 1u), ((nk >> 1) & 1u) ^ 1u);
              next_full = mbar_try_wait(full_bar(nstage), ((fill + 1u) / NSTAGE) & 1u);
            }
#pragma unroll
            for (int ks = 0; ks < C_::KSTEPS; ++ks) {
              const uint64_t adesc = umma_smem_desc(a_tap + 2 * ks * C_::A_CHUNK, C_::A_CHUNK, C_::A_ROW);
              const uint64_t bdesc = umma_smem_desc(b_tap + 2 * ks * (NT * 16), NT * 16, 128);
              umma_bf16(d_tmem, adesc, bdesc, idesc, (tap | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          umma_commit(tfull_bar(as));
          tl_stamp(g, 1, k, 3);
          any = true;
        }
        if (any) {
          umma_commit(wfree_bar(wb));
          free_pending |= 1u << wb;
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue: TMEM -> registers -> global ========================
    EpiCtx cx;
    cx.eg = warp >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    cx.r = m >> 3;
    cx.c = m & 7;
    cx.lane = lane;
    cx.tl0 = (q == 0 && lane == 0);
    cx.g = g;
    cx.done = done;
    cx.pub_seen = pub_seen;
    cx.tfull = tfull_bar(cx.eg);
    cx.tempty = tempty_bar(cx.eg);
    cx.pub_bar0 = pub_bar(0);
    cx.taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cx.eg * C_::ACC_STRIDE;
    cx.G = G;
    cx.H = H;
    cx.W = W;
    cx.tiles_per_img = tiles_per_img;
    cx.chunk_stride = static_cast<size_t>(W) * 8;

    pdl_wait();
    uint32_t k = 0;   // CTA-local job counter; this group handles the jobs with k % 2 == eg
    for (int l = 0; l < nlayers; ++l) {
      const lv_conv_args& a = P.layer[l];
      const bool has_ops = (a.mask != nullptr) || (a.res1 != nullptr) || (a.res2 != nullptr);
      int kind = kKindGeneric;
      if (a.cout == NT && a.res_scale == 1.0f) {
        if (a.epilogue == LV_EPI_NHWC) {
          const int code = (a.relu ? 1 : 0) | (a.mask ? 2 : 0) | (a.res1 ? 4 : 0) | (a.res2 ? 8 : 0);
          if (code == 0 || code == 1 || code == 2 || code == 4 || code == 12) kind = code;
        } else if (a.epilogue == LV_EPI_PS4_NCHW && !a.relu && !has_ops) {
          kind = kKindPs4;
        }
      }
      const int first = first_tile(l);
      switch (kind) {
        case 0: k = run_layer<0, NT>(cx, a, l, first, k); break;
        case 1: k = run_layer<1, NT>(cx, a, l, first, k); break;
        case 2: k = run_layer<2, NT>(cx, a, l, first, k); break;
        case 4: k = run_layer<4, NT>(cx, a, l, first, k); break;
        case 12: k = run_layer<12, NT>(cx, a, l, first, k); break;
        case kKindPs4: k = run_layer<kKindPs4, NT>(cx, a, l, first, k); break;
        default: k = run_layer<kKindGeneric, NT>(cx, a, l, first, k); break;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<C_::TMEM_COLS>(tmem_base);
  }
  // self-cleaning workspace: done[total_tiles] is the exit counter; the last CTA out resets everything
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t prev = atomicAdd(done + g.total_tiles, 1u);
    *s_last = (prev == static_cast<uint32_t>(G) - 1u) ? 1u : 0u;
    __threadfence();
  }
  __syncthreads();
  if (*s_last != 0u) {
    for (int i = threadIdx.x; i <= g.total_tiles; i += kThreads) done[i] = 0u;
  }
}

}  // namespace chain

#ifdef LV_EXPERIMENTAL
int conv3x3_strip(const lv_conv_args* layers, int count, cudaStream_t stream);
#endif

long long conv3x3_chain_workspace_bytes(int n, int h, int w) {
  const long long tiles = static_cast<long long>(n) * ((h + chain::kTileH - 1) / chain::kTileH) * ((w + chain::kTileW - 1) / chain::kTileW);
  return (tiles + 1) * 4;
}

// layers[0..count): 48 -> 48 bf16 convs with tap-major weights on one common (n, h, w); executed in order with
// data-flow synchronisation.  `sync_ws` must be zero before the first launch (the kernel re-zeroes it on exit).
int conv3x3_chain(const lv_conv_args* layers, int count, void* sync_ws, long long sync_ws_bytes, int max_ctas,
                  cudaStream_t stream) {
  using C_ = chain::Cfg<48, 48, 4>;
  LV_CHECK_ARG(count >= 1 && count <= chain::kMaxLayers, "conv chain: 1..%d layers per call (got %d)", chain::kMaxLayers, count);
  const lv_conv_args& a0 = layers[0];
  for (int i = 0; i < count; ++i) {
    const lv_conv_args& a = layers[i];
    LV_CHECK_ARG(a.dtype == LV_BF16 && a.cin == 48 && a.num_src == 1 && a.cout == 48 && a.wlayout == LV_W_TAP_MAJOR,
                 "conv chain: layer %d is not a single-source bf16 48->48 conv with tap-major weights", i);
    LV_CHECK_ARG(a.n == a0.n && a.h == a0.h && a.w == a0.w, "conv chain: layer %d has a different geometry", i);
  }
  ConvGeom g;
  g.timeline = g_timeline;
  g.cout_pad = 48;
  g.nt = 48;
  g.ntiles_n = 1;
  g.tiles_x = (a0.w + chain::kTileW - 1) / chain::kTileW;
  g.tiles_y = (a0.h + chain::kTileH - 1) / chain::kTileH;
  const long long tt = static_cast<long long>(a0.n) * g.tiles_x * g.tiles_y;
  if (tt == 0) return LV_OK;
  LV_CHECK_ARG(tt < (1ll << 30), "conv chain: too many tiles (%lld)", tt);
  LV_CHECK_ARG(sync_ws != nullptr && sync_ws_bytes >= (tt + 1) * 4, "conv chain: sync workspace too small (%lld < %lld bytes)",
               sync_ws_bytes, (tt + 1) * 4);
  g.total_tiles = static_cast<int>(tt);

#ifdef LV_EXPERIMENTAL
  // small images (patch training): one cluster per image, activations resident in shared memory
  // (tools/experiments/conv_strip.cu, opt-in at run time with LARVANET_B200_STRIP=1)
  if (static_cast<long long>(a0.n) * a0.h * a0.w > 0) {
    const int rc = conv3x3_strip(layers, count, stream);
    if (rc != 1) return rc;
  }
#endif
  static thread_local chain::Params params;   // staging only; the launch copies it by value
  static thread_local chain::Maps maps;
  for (int i = 0; i < count; ++i) {
    params.layer[i] = layers[i];
    if (chain::kTmaHalo) {
      const int rc = activation_tile_map(layers[i].src[0], a0.n, a0.h, a0.w, C_::CH, chain::kHaloW, chain::kHaloH, &maps.src[i]);
      if (rc != LV_OK) return rc;
    }
  }
  auto kern = chain::conv3x3_chain_kernel<48, 48, 4>;
  // once per DEVICE: opt into the dynamic shared memory and make sure one CTA per SM can really be resident -- the
  // data-flow waits need every CTA of the grid running at the same time
  static bool configured[64] = {false};
  int dev = 0;
  LV_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (!configured[dev]) {
    LV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(C_::smem_bytes())));
    int per_sm = 0;
    LV_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, chain::kThreads, C_::smem_bytes()));
    if (per_sm < 1) {
      set_error("conv chain: a CTA (%zu B shared memory) does not fit on an SM of device %d", C_::smem_bytes(), dev);
      return LV_ERR_UNSUPPORTED;
    }
    configured[dev] = true;
  }
  long long ctas = max_ctas > 0 ? max_ctas : sm_count();
  if (ctas > sm_count()) ctas = sm_count();   // one resident CTA per SM: the data-flow waits need every CTA running
  if (ctas > tt) ctas = tt;
  const int rot = static_cast<int>(tt % ctas);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(chain::kThreads);
  cfg.dynamicSmemBytes = C_::smem_bytes();
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, params, maps, count, g, static_cast<uint32_t*>(sync_ws), rot));
  count_launch();
  return LV_OK;
}

}  // namespace lv
