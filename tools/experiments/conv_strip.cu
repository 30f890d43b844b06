// Cluster-resident variant of the conv chain (conv_chain.cu) for small images (patch training): one thread-block
// cluster per image, one CTA per 8-pixel-wide tile column ("strip", <= 4 tiles high), resident for ALL layers.
//
// conv_chain.cu hands a tile from one layer to the next through global memory and per-tile release/acquire flags:
// ~3.8k clk of store-acknowledge + flag + poll + L2 load per layer for ~2.6k clk of MMA work.  Here the activations never
// leave the SM on the conv-input path:
//   * each CTA keeps two activation buffers in shared memory, laid out exactly like conv_tc.cu's halo tile but as tall
//     as the image ([8-channel chunk][rows + 2][10 px][16 B]); the A operand of tile ty / tap (ky,kx) is a descriptor
//     into the layer's input buffer;
//   * the epilogue writes its bf16 results to global memory as before (saved activations / outputs) AND into the
//     other buffer (the next layer's input), and the threads of the strip's first / last pixel column also store theirs
//     into the right / left halo column of the neighbour CTA through distributed shared memory;
//   * one mbarrier per layer parity counts the epilogue warps of the CTA itself and of its two neighbours (remote
//     arrive, release.cluster): when it completes, the layer's output strip and both halo columns are in place and the
//     previous layer's input buffer is free -- ~0.5k clk instead of the global hand-over;
//   * layers whose input is not the buffer content (the first layer, the exits' gradients in the backward pass) load
//     their strip from global memory with cp.async.
// Clusters are independent (zero padding at the image border), so no inter-cluster synchronisation exists and it does
// not matter how many clusters are resident at once (measured: 15 clusters of 8-9 CTAs, 33 of 4 fit on 148 SMs).
//
// STATUS: experimental, opt-in with LARVANET_B200_STRIP=1.  Bit-identical to per-layer launches on every shape of the
// chain tests, but slower than conv_chain.cu today (see conv3x3_strip below for the measured reasons).
#include <cstdlib>

#include "chain_epilogue.cuh"
#include "conv_epilogue.cuh"
#include "lv_common.cuh"

namespace lv {

extern long long* g_timeline;
extern int g_use_pdl;

namespace strip {

constexpr int kMaxLayers = 96;
constexpr int kTileH = 16, kTileW = 8, kHaloW = 10;
constexpr int kEpiWarps = 8, kEpiThreads = kEpiWarps * 32, kProdThreads = 96;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kThreads = kEpiThreads + 32 + kProdThreads;   // 384
constexpr int kMaxTilesY = 4, kMaxTilesX = 8;
constexpr int kCin = 48, kNT = 48, kCH = kCin / 8, kKSteps = kCin / 16;
constexpr int kWTap = kCH * kNT * 16, kWLayer = 9 * kWTap;       // 41472
constexpr int kRowBytes = kHaloW * 16;                           // 160
constexpr int kKindPs4 = 100, kKindGeneric = -1;

struct Params {
  lv_conv_args layer[kMaxLayers];
  signed char in_buf[kMaxLayers];     // activation buffer the layer reads
  signed char out_buf[kMaxLayers];    // buffer its results go to (-1: none, e.g. PixelShuffle exits)
  signed char from_global[kMaxLayers];// 1: the input strip is loaded from a.src[0] first
  signed char defer_store[kMaxLayers];// 1: global stores of the results may follow the hand-over (nobody reads them as an
                                      //    epilogue operand in the very next layer)
};

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, 400;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_strip_kernel(const __grid_constant__ Params P, const int nlayers, const ConvGeom g) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = P.layer[0].h, W = P.layer[0].w;
  const int tiles_y = g.tiles_y;
  const int R2 = tiles_y * kTileH + 2;                 // buffer rows: image rows -1 .. tiles_y*16
  const uint32_t plane = static_cast<uint32_t>(R2) * kRowBytes;
  const uint32_t buf_bytes = kCH * plane;

  uint8_t* sW = smem;                                  // two weight buffers
  uint8_t* sAct = sW + 2 * kWLayer;                    // two activation buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAct + 2 * buf_bytes);
  const uint32_t bar0 = smem_u32(bars);
  auto tfull_bar = [&](int s) { return bar0 + 8u * s; };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 + s); };
  auto wfull_bar = [&](int b) { return bar0 + 8u * (4 + b); };
  auto layer_bar = [&](int p) { return bar0 + 8u * (6 + p); };
  const uint32_t load_bar = bar0 + 8u * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const uint32_t rank = cluster_rank();                // tile column of this CTA
  const int tx = static_cast<int>(rank);
  const int n = static_cast<int>(blockIdx.x) / g.tiles_x;   // image of this cluster
  const bool has_left = tx > 0, has_right = tx + 1 < g.tiles_x;
  const int x0 = tx * kTileW;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiThreads / 2);
      mbar_init(wfull_bar(s), 1);
      // every epilogue warp of this CTA and of its neighbours arrives once per tile
      mbar_init(layer_bar(s), 4u * tiles_y * (1u + (has_left ? 1u : 0u) + (has_right ? 1u : 0u)));
    }
    mbar_init(load_bar, kProdThreads);
    mbar_fence_init();
  }
  // zero both activation buffers: halo rows / columns outside the image are never written again
  for (uint32_t i = threadIdx.x; i < 2 * buf_bytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(sAct)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == kMmaWarp) tmem_alloc<128>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  cluster_sync_all();          // neighbours' barriers and buffers are initialised before anyone stores into them
  pdl_launch_dependents();
  pdl_wait();

  if (warp > kMmaWarp) {
    // =============================== loaders: strips that come from global memory ===============================
    const int ptid = threadIdx.x - (kEpiThreads + 32);
    const int pieces = R2 * kHaloW * kCH;      // 16 B pieces of one strip incl. halo
    uint32_t nload = 0;
    for (int l = 0; l < nlayers; ++l) {
      // Follow EVERY layer's barrier, also when there is nothing to load: a parity wait is only meaningful for a waiter
      // that has seen all earlier phases.  Once the previous layer is complete everybody (own epilogues and both
      // neighbours') is done with it: the buffer is free and the tensor -- whoever in the cluster wrote it -- is visible.
      if (l > 0) {
        if (lane == 0) mbar_wait_cluster(layer_bar((l - 1) & 1), ((l - 1) >> 1) & 1);
        __syncwarp();
      }
      if (!P.from_global[l]) continue;
      const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(P.layer[l].src[0]);
      const uint32_t dst0 = smem_u32(sAct + P.in_buf[l] * buf_bytes);
      for (int idx = ptid; idx < pieces; idx += kProdThreads) {
        const int col = idx % kHaloW, rc = idx / kHaloW;
        const int c = rc % kCH, r = rc / kCH;
        const int gy = r - 1, gx = x0 - 1 + col;
        const bool inb = (static_cast<unsigned>(gy) < static_cast<unsigned>(H)) && (static_cast<unsigned>(gx) < static_cast<unsigned>(W));
        const __nv_bfloat16* p = inb ? src + act_off(n, gy, gx, c, H, W, kCH) : src;
        cp_async16(dst0 + c * plane + (r * kHaloW + col) * 16, p, inb ? 16u : 0u);
      }
      cp_async_mbar_arrive_noinc(load_bar);
      ++nload;
    }
    (void)nload;
    cp_async_wait<0>();
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer (one elected lane) ================================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kNT, 0, 0);
      auto load_weights = [&](int l) {
        const int b = l & 1;
        mbar_arrive_expect_tx(wfull_bar(b), kWLayer);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.layer[l].weights);
        for (int t = 0; t < 9; ++t)
          tma_bulk_g2s(smem_u32(sW + b * kWLayer + t * kWTap), wsrc + static_cast<size_t>(t) * kWTap, kWTap, wfull_bar(b));
      };
      load_weights(0);
      uint32_t k = 0, nload = 0;
      for (int l = 0; l < nlayers; ++l) {
        tl_stamp(g, 0, k, 0);
        if (l > 0) mbar_wait_cluster(layer_bar((l - 1) & 1), ((l - 1) >> 1) & 1);   // input strip + halo columns in place
        tl_stamp(g, 0, k, 1);
        if (l + 1 < nlayers) load_weights(l + 1);   // buffer (l+1)&1 was read by layer l-1, whose MMAs have retired
        mbar_wait(wfull_bar(l & 1), (l >> 1) & 1);
        if (P.from_global[l]) {
          mbar_wait(load_bar, nload & 1);
          ++nload;
        }
        tl_stamp(g, 0, k, 2);
        fence_proxy_async_smem();   // epilogue / DSMEM / cp.async writes (generic proxy) -> UMMA reads (async proxy)
        tc_fence_after_sync();
        tl_stamp(g, 0, k, 3);
        const uint32_t sW_addr = smem_u32(sW + (l & 1) * kWLayer);
        const uint32_t in_addr = smem_u32(sAct + P.in_buf[l] * buf_bytes);
        for (int ty = 0; ty < tiles_y; ++ty, ++k) {
          const uint32_t as = k & 1;
          tl_stamp(g, 1, k, 0);
          mbar_wait(tempty_bar(as), ((k >> 1) & 1) ^ 1);
          tc_fence_after_sync();
          tl_stamp(g, 1, k, 1);
          const uint32_t d_tmem = tmem_base + as * 64;
          const uint32_t a_addr = in_addr + ty * kTileH * kRowBytes;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_tap = a_addr + ((tap / 3) * kHaloW + (tap % 3)) * 16;
            const uint32_t b_tap = sW_addr + tap * kWTap;
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks) {
              const uint64_t adesc = umma_smem_desc(a_tap + 2 * ks * plane, plane, kRowBytes);
              const uint64_t bdesc = umma_smem_desc(b_tap + 2 * ks * (kNT * 16), kNT * 16, 128);
              umma_bf16(d_tmem, adesc, bdesc, idesc, (tap | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(tfull_bar(as));
          tl_stamp(g, 1, k, 3);
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue: TMEM -> registers -> global + shared + DSMEM ========================
    const int eg = warp >> 2, q = warp & 3;
    const int m = q * 32 + lane;
    const int r = m >> 3, c = m & 7;
    const uint32_t as = eg;
    constexpr int NCH = kNT / 8;
    const size_t chunk_stride = static_cast<size_t>(W) * 8;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 64;
    const uint32_t act0 = smem_u32(sAct);
    // neighbour addresses of this CTA's buffers and barriers
    const uint32_t act_left = has_left ? mapa(act0, rank - 1) : 0u, act_right = has_right ? mapa(act0, rank + 1) : 0u;
    const uint32_t lbar_left = has_left ? mapa(layer_bar(0), rank - 1) : 0u;
    const uint32_t lbar_right = has_right ? mapa(layer_bar(0), rank + 1) : 0u;

    uint32_t k = 0;
    for (int l = 0; l < nlayers; ++l) {
      const lv_conv_args& a = P.layer[l];
      const bool has_ops = (a.mask != nullptr) || (a.res1 != nullptr) || (a.res2 != nullptr);
      int kind = kKindGeneric;
      if (a.cout == kNT && a.res_scale == 1.0f) {
        if (a.epilogue == LV_EPI_NHWC) {
          const int code = (a.relu ? 1 : 0) | (a.mask ? 2 : 0) | (a.res1 ? 4 : 0) | (a.res2 ? 8 : 0);
          if (code == 0 || code == 1 || code == 2 || code == 4 || code == 12) kind = code;
        } else if (a.epilogue == LV_EPI_PS4_NCHW && !a.relu && !has_ops) {
          kind = kKindPs4;
        }
      }
      chain::FastEpi fe;
      fe.mask = reinterpret_cast<const __nv_bfloat16*>(a.mask);
      fe.res1 = reinterpret_cast<const __nv_bfloat16*>(a.res1);
      fe.res2 = reinterpret_cast<const __nv_bfloat16*>(a.res2);
      fe.out = reinterpret_cast<__nv_bfloat16*>(a.out);
      fe.res_scale = a.res_scale;
      fe.relu = a.relu;
      const int ob = P.out_buf[l];
      const float* bias_g = a.bias;
      float loss = 0.f;
      // same-pixel operands (residuals, masks) were written by this CTA's epilogue warps in earlier layers: the
      // previous layer's barrier orders them before the loads below
      if (l > 0) {
        if (lane == 0) mbar_wait_cluster(layer_bar((l - 1) & 1), ((l - 1) >> 1) & 1);
        __syncwarp();
      }
      for (int ty = 0; ty < tiles_y; ++ty, ++k) {
        if ((k & 1u) != static_cast<uint32_t>(eg)) continue;
        const int y = ty * kTileH + r, x = x0 + c;
        const bool valid = (y < H) && (x < W);
        const size_t o0 = valid ? act_off(n, y, x, 0, H, W, NCH) : 0;
        const uint32_t par = (k >> 1) & 1;
        const bool tl0 = (q == 0 && lane == 0);
        if (tl0) tl_stamp(g, 2, k, 0);
        chain::SmemOut so{0u, 0u, 0u, plane};
        uint4 held[NCH];
        uint4* deferred = (P.defer_store[l] && ob >= 0) ? held : nullptr;
        if (ob >= 0) {
          const uint32_t pix = static_cast<uint32_t>(ob) * buf_bytes + ((y + 1) * kHaloW + (c + 1)) * 16;
          so.own = act0 + pix;
          // first / last column of the strip = right / left halo column of the neighbour
          if (c == 0 && has_left) so.left = act_left + static_cast<uint32_t>(ob) * buf_bytes + ((y + 1) * kHaloW + 9) * 16;
          if (c == kTileW - 1 && has_right) so.right = act_right + static_cast<uint32_t>(ob) * buf_bytes + ((y + 1) * kHaloW + 0) * 16;
        }
        switch (kind) {
          case 0: chain::fast_tile<0, kNT, false, true>(fe, nullptr, bias_g, valid, o0, chunk_stride, taddr, tfull_bar(as), tempty_bar(as), par, so, deferred); break;
          case 1: chain::fast_tile<1, kNT, false, true>(fe, nullptr, bias_g, valid, o0, chunk_stride, taddr, tfull_bar(as), tempty_bar(as), par, so, deferred); break;
          case 2: chain::fast_tile<2, kNT, false, true>(fe, nullptr, bias_g, valid, o0, chunk_stride, taddr, tfull_bar(as), tempty_bar(as), par, so, deferred); break;
          case 4: chain::fast_tile<4, kNT, false, true>(fe, nullptr, bias_g, valid, o0, chunk_stride, taddr, tfull_bar(as), tempty_bar(as), par, so, deferred); break;
          case 12: chain::fast_tile<12, kNT, false, true>(fe, nullptr, bias_g, valid, o0, chunk_stride, taddr, tfull_bar(as), tempty_bar(as), par, so, deferred); break;
          case kKindPs4:
            loss += chain::ps4_tile<kNT>(a, bias_g, valid, n, y, x, H, W, o0, chunk_stride, taddr, tfull_bar(as), tempty_bar(as), par);
            break;
          default: {
            mbar_wait_relaxed(tfull_bar(as), par);
            tc_fence_after_sync();
#pragma unroll 1
            for (int j = 0; j < kNT / 16; ++j) {
              float v[16];
              tmem_ld16(taddr + j * 16, v);
              tmem_ld_wait();
              if (valid) loss += conv_epilogue16<__nv_bfloat16>(a, n, y, x, j * 16, v);
            }
            tc_fence_before_sync();
            mbar_arrive(tempty_bar(as));
          } break;
        }
        if (tl0) tl_stamp(g, 2, k, 2);
        // this warp's part of (layer l, tile ty) is stored: tell this CTA's MMA warp and both neighbours'
        __syncwarp();
        if (lane == 0) {
          const uint32_t off = 8u * (l & 1);
          mbar_arrive_cluster(mapa(layer_bar(0), rank) + off);
          if (has_left) mbar_arrive_cluster(lbar_left + off);
          if (has_right) mbar_arrive_cluster(lbar_right + off);
        }
        if (deferred != nullptr && kind >= 0 && valid && fe.out != nullptr) {
#pragma unroll
          for (int j = 0; j < NCH; ++j) *reinterpret_cast<uint4*>(fe.out + o0 + j * chunk_stride) = held[j];
        }
        if (tl0) tl_stamp(g, 2, k, 3);
      }
      if (a.truth_hr != nullptr && a.loss_sum != nullptr) {
        loss = warp_sum(loss);
        if (lane == 0) atomicAdd(a.loss_sum, static_cast<double>(loss));
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<128>(tmem_base);
  }
  cluster_sync_all();   // nobody leaves while a neighbour may still store into its shared memory or barriers
}

}  // namespace strip

// Returns LV_OK when the chain was launched on the cluster-resident kernel, 1 when it is not eligible (the caller falls
// back to the general chain kernel), or an error code.
int conv3x3_strip(const lv_conv_args* layers, int count, cudaStream_t stream) {
  // Opt-in (LARVANET_B200_STRIP=1): bit-exact, but measured SLOWER than the flag-based chain on B200 (8.4 vs 4.2 us per
  // layer at 16 x 48x48): the per-tile epilogue with shared-memory + DSMEM stores takes ~3k clk and every
  // mbarrier.arrive.release.cluster another ~2k clk (it waits for the remote stores), so the hand-over is no cheaper
  // than release/acquire flags through L2.  Kept for the next round's work on cheaper cluster hand-overs.
  static const bool enabled = [] { const char* e = getenv("LARVANET_B200_STRIP"); return e != nullptr && e[0] == '1'; }();
  if (!enabled || count > strip::kMaxLayers) return 1;
  const lv_conv_args& a0 = layers[0];
  const int tiles_x = (a0.w + strip::kTileW - 1) / strip::kTileW, tiles_y = (a0.h + strip::kTileH - 1) / strip::kTileH;
  if (tiles_x < 1 || tiles_x > strip::kMaxTilesX || tiles_y < 1 || tiles_y > strip::kMaxTilesY || a0.n < 1) return 1;
  if (static_cast<long long>(a0.n) * tiles_x > 65535) return 1;
  static thread_local strip::Params params;
  // which tensor each shared-memory buffer holds while the chain runs
  const void* holds[2] = {nullptr, nullptr};
  int last_out = -1;
  for (int i = 0; i < count; ++i) {
    const lv_conv_args& a = layers[i];
    if (a.epilogue != LV_EPI_NHWC && a.epilogue != LV_EPI_PS4_NCHW) return 1;
    params.layer[i] = a;
    int in = -1;
    if (holds[0] == a.src[0] && a.src[0] != nullptr) in = 0;
    else if (holds[1] == a.src[0] && a.src[0] != nullptr) in = 1;
    params.from_global[i] = static_cast<signed char>(in < 0);
    if (in < 0) {
      in = (last_out >= 0) ? 1 - last_out : 0;   // overwrite the older buffer
      holds[in] = a.src[0];
    }
    params.in_buf[i] = static_cast<signed char>(in);
    int out = -1;
    if (a.epilogue == LV_EPI_NHWC && a.out != nullptr) {
      out = 1 - in;
      if (holds[in] == a.out) holds[in] = nullptr;   // the tensor is being rewritten: the old copy is stale
      holds[out] = a.out;
      last_out = out;
    }
    // side outputs that later layers may read as their input (sign gradient of an exit): never resident
    if (a.grad_sign != nullptr) {
      for (int b = 0; b < 2; ++b)
        if (holds[b] == a.grad_sign) holds[b] = nullptr;
    }
    params.out_buf[i] = static_cast<signed char>(out);
  }
  for (int i = 0; i < count; ++i) {
    // the next layer's epilogue may read this layer's output as residual / mask: then the stores must precede the hand-over
    bool next_reads = false;
    if (i + 1 < count && layers[i].out != nullptr) {
      const lv_conv_args& nx = layers[i + 1];
      next_reads = (nx.res1 == layers[i].out) || (nx.res2 == layers[i].out) || (nx.mask == layers[i].out);
    }
    params.defer_store[i] = static_cast<signed char>(!next_reads);
  }
  ConvGeom g;
  g.timeline = g_timeline;
  g.cout_pad = 48;
  g.nt = 48;
  g.ntiles_n = 1;
  g.tiles_x = tiles_x;
  g.tiles_y = tiles_y;
  g.total_tiles = a0.n * tiles_x * tiles_y;
  const int R2 = tiles_y * strip::kTileH + 2;
  const size_t smem = 2 * static_cast<size_t>(strip::kWLayer) + 2 * static_cast<size_t>(strip::kCH) * R2 * strip::kRowBytes + 256;
  auto kern = strip::conv3x3_strip_kernel;
  static size_t configured = 0;
  if (smem > configured) {
    LV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(a0.n * tiles_x));
  cfg.blockDim = dim3(strip::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(tiles_x);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  LV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, params, count, g));
  count_launch();
  return LV_OK;
}

}  // namespace lv
