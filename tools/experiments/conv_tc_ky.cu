// ky-stacked variant of the tensor-core 3x3 convolution (conv_tc.cu): three times fewer, three times wider MMAs.
//
// conv_tc.cu issues 27 MMAs (9 taps x 3 K steps) of N = 48 per tile; each costs max(N/2, 32+N/4) = 45 clk, almost all
// of it the 4 KB re-read of the A tile from shared memory (the tensor pipe itself would need 24 clk).  Here the three
// vertical taps are stacked along N instead:  for every horizontal tap kx and K step,
//     P[128 px x (3*48)] += A_kx[128 x 16] * [W(ky=0,kx) | W(ky=1,kx) | W(ky=2,kx)][16 x 144]
// i.e. 9 MMAs of N = 144 (73 clk each = N/2: tensor-pipe bound, A is read 9 instead of 27 times).  The accumulator
// lanes are INPUT rows: lane (rho, c) holds P_ky(rho, c) = sum_kx,cin X[rho][c+kx-1] W[ky][kx], and the output row r is
//     out(r) = P_0(r) + P_1(r+1) + P_2(r+2)          (rho = r .. r+2, rows counted from the halo row R0-1)
// so the epilogue thread of lane-row rho produces output row rho-1 from its own P_1, P_0 of the lane 8 below (one
// warp shuffle) and P_2 of the lane 8 above; the two lane-rows at every warp boundary go through a 9 KB shared-memory
// exchange.  A tile therefore yields 14 output rows x 8 columns from 16 x 10 input pixels.
#include "conv_epilogue.cuh"
#include "lv_common.cuh"

namespace lv {

extern long long* g_timeline;
extern int g_use_pdl;

namespace ky {

constexpr int kRowsIn = 16, kRowsOut = 14, kTileW = 8, kHaloW = 10, kHaloPix = kRowsIn * kHaloW;  // 160
constexpr int kEpiThreads = 256, kProdThreads = 96, kMmaWarp = 8;
constexpr int kThreads = kEpiThreads + 32 + kProdThreads;  // 384
constexpr int kAccStride = 256;                            // TMEM columns between the two accumulator stages
constexpr size_t kMaxSmem = 227 * 1024;

template <int CIN, int NT, int NSTAGE>
struct Cfg {
  static constexpr int CH = CIN / 8;
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int N3 = 3 * NT;                      // stacked N
  static constexpr int A_PLANE = kHaloPix * 16;          // 2560 B
  static constexpr int A_STAGE = CH * A_PLANE;
  static constexpr int W_KX = CH * N3 * 16;              // bytes of one (src,kx) weight block
  static constexpr int XPITCH = NT + 4;                  // floats per exchanged lane (+4: conflict-free float4 rows)
  static constexpr int XCHG = 3 * 2 * 8 * XPITCH * 4;    // one buffer: 3 warp boundaries x {P0 up, P2 down} x 8 lanes
  static constexpr int PROD_PIECES = (kHaloPix * CH + kProdThreads - 1) / kProdThreads;
  // xbuf = exchange buffers per epilogue group: 2 (one named barrier per tile) or 1 (two barriers, when smem is tight)
  static size_t smem_bytes(int num_src, int cout_pad, int xbuf) {
    return static_cast<size_t>(num_src) * 3 * W_KX + static_cast<size_t>(NSTAGE) * A_STAGE + 2 * xbuf * XCHG +
           static_cast<size_t>(cout_pad) * 4 + 256;
  }
};

template <int CIN, int NT, int NSTAGE, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_ky_kernel(const __grid_constant__ lv_conv_args a, const ConvGeom g, const int xbuf) {
  using C_ = Cfg<CIN, NT, NSTAGE>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const uint32_t w_bytes = static_cast<uint32_t>(a.num_src) * 3u * C_::W_KX;
  uint8_t* sW = smem;
  uint8_t* sA = sW + w_bytes;
  float* sX = reinterpret_cast<float*>(sA + NSTAGE * C_::A_STAGE);       // exchange buffers, two per epilogue group
  float* sBias = sX + 2 * xbuf * (C_::XCHG / 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + g.cout_pad);
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
      mbar_init(tempty_bar(s), kEpiThreads / 2);
    }
    mbar_init(wbar, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < g.cout_pad; i += kThreads) sBias[i] = (a.bias != nullptr && i < a.cout) ? a.bias[i] : 0.f;
  if (warp == kMmaWarp) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  const int tiles_per_img = g.tiles_x * g.tiles_y;

  if (warp > kMmaWarp) {
    // =============================== producers ===============================
    const int ptid = threadIdx.x - (kEpiThreads + 32);
    uint32_t pc_dst[C_::PROD_PIECES];
    int pc_rel[C_::PROD_PIECES], pc_rc[C_::PROD_PIECES];
#pragma unroll
    for (int i = 0; i < C_::PROD_PIECES; ++i) {
      const int idx = ptid + i * kProdThreads;
      const int col = idx % kHaloW, rc = idx / kHaloW;
      const int c = rc % C_::CH, r = rc / C_::CH;
      pc_dst[i] = c * C_::A_PLANE + (r * kHaloW + col) * 16;
      pc_rel[i] = ((r * C_::CH + c) * a.w + col) * 8;
      pc_rc[i] = (idx < kHaloPix * C_::CH) ? ((r << 8) | col) : -1;
    }
    uint32_t fill = 0;
    pdl_wait();
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int ty = rem / g.tiles_x;
      const int y0 = ty * kRowsOut - 1;
      const int x0 = (rem - ty * g.tiles_x) * kTileW - 1;
      const long long origin = ((static_cast<long long>(n) * a.h + y0) * C_::CH * a.w + x0) * 8;
      for (int s = 0; s < a.num_src; ++s, ++fill) {
        const int stage = fill % NSTAGE;
        if (ptid == 0) tl_stamp(g, 0, fill, 0);
        mbar_wait_relaxed(empty_bar(stage), ((fill / NSTAGE) & 1) ^ 1);
        if (ptid == 0) tl_stamp(g, 0, fill, 1);
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.src[s]) + origin;
        const uint32_t dst0 = smem_u32(sA + stage * C_::A_STAGE);
#pragma unroll
        for (int i = 0; i < C_::PROD_PIECES; ++i) {
          if (pc_rc[i] >= 0) {
            const int gy = y0 + (pc_rc[i] >> 8), gx = x0 + (pc_rc[i] & 0xff);
            const bool inb = (static_cast<unsigned>(gy) < static_cast<unsigned>(a.h)) &&
                             (static_cast<unsigned>(gx) < static_cast<unsigned>(a.w));
            cp_async16(dst0 + pc_dst[i], inb ? (src + pc_rel[i]) : reinterpret_cast<const __nv_bfloat16*>(a.src[s]),
                       inb ? 16u : 0u);
          }
        }
        cp_async_mbar_arrive_noinc(full_bar(stage));
        if (ptid == 0) tl_stamp(g, 0, fill, 2);
      }
    }
    cp_async_wait<0>();
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer ===============================
    if (elect_one()) {
      mbar_arrive_expect_tx(wbar, w_bytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.weights);
      for (int b = 0; b < a.num_src * 3; ++b)
        tma_bulk_g2s(smem_u32(sW + b * C_::W_KX), wsrc + static_cast<size_t>(b) * C_::W_KX, C_::W_KX, wbar);
      mbar_wait(wbar, 0);
      constexpr uint32_t idesc = umma_idesc_bf16(128, C_::N3, 0, 0);
      const uint32_t sW_addr = smem_u32(sW);
      uint32_t fill = 0, k = 0;
      for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++k) {
        const uint32_t as = k & 1;
        tl_stamp(g, 1, k, 0);
        mbar_wait(tempty_bar(as), ((k >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        tl_stamp(g, 1, k, 1);
        const uint32_t d_tmem = tmem_base + as * kAccStride;
        uint32_t accumulate = 0;
        for (int s = 0; s < a.num_src; ++s, ++fill) {
          const int stage = fill % NSTAGE;
          mbar_wait(full_bar(stage), (fill / NSTAGE) & 1);
          fence_proxy_async_smem();
          tc_fence_after_sync();
          tl_stamp(g, 1, k, 2);
          const uint32_t a_addr = smem_u32(sA + stage * C_::A_STAGE);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const uint32_t b_kx = sW_addr + (s * 3 + kx) * C_::W_KX;
#pragma unroll
            for (int ks = 0; ks < C_::KSTEPS; ++ks) {
              const uint64_t adesc = umma_smem_desc(a_addr + kx * 16 + 2 * ks * C_::A_PLANE, C_::A_PLANE, kHaloW * 16);
              const uint64_t bdesc = umma_smem_desc(b_kx + 2 * ks * (C_::N3 * 16), C_::N3 * 16, 128);
              umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(empty_bar(stage));
        }
        umma_commit(tfull_bar(as));
        tl_stamp(g, 1, k, 3);
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    const int eg = warp >> 2, q = warp & 3;
    const int m = q * 32 + lane;
    const int rho = m >> 3, c = m & 7;          // input (halo) row index 0..15 and column of this lane
    const int lrow = lane >> 3;                 // lane-row inside the warp 0..3
    float loss = 0.f;
    const uint32_t as = eg;
    // exchange slots (double buffered per group, one named barrier per tile):
    //   up[b] = P0 of lane-row 3 of warp b   (consumed by lane-row 0 of warp b+1), b = 0..2
    //   dn[b] = P2 of lane-row 0 of warp b+1 (consumed by lane-row 3 of warp b)
    constexpr int XP = C_::XPITCH;
    const bool pub_up = (lrow == 3) && (q < 3), pub_dn = (lrow == 0) && (q > 0);
    const bool need_up = pub_dn, need_dn = pub_up;
    // EPI >= 0: compile-time flag set (bit0 relu, bit1 mask, bit2 res1, bit3 res2) of the planar bf16 epilogue
    const bool fast = (EPI >= 0) || ((a.epilogue == LV_EPI_NHWC) && (a.cout == g.cout_pad));
    const bool unit_scale = (EPI >= 0) || (a.res_scale == 1.0f);
    const bool do_relu = (EPI >= 0) ? ((EPI & 1) != 0) : (a.relu != 0);
    const bool do_mask = (EPI >= 0) ? ((EPI & 2) != 0) : (a.mask != nullptr);
    const bool do_res1 = (EPI >= 0) ? ((EPI & 4) != 0) : (a.res1 != nullptr);
    const bool do_res2 = (EPI >= 0) ? ((EPI & 8) != 0) : (a.res2 != nullptr);
    constexpr int NCH = NT / 8;
    const size_t chunk_stride = static_cast<size_t>(a.w) * 8;

    pdl_wait();
    const int tile_stride = 2 * gridDim.x;
    uint32_t k = eg;
    for (int tile = blockIdx.x + eg * gridDim.x; tile < g.total_tiles; tile += tile_stride, k += 2) {
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int tyi = rem / g.tiles_x;
      const int y = tyi * kRowsOut + rho - 1, x = (rem - tyi * g.tiles_x) * kTileW + c;
      const bool valid = (rho >= 1) && (rho <= kRowsOut) && (y < a.h) && (x < a.w);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kAccStride;
      const bool tl0 = (threadIdx.x == 0);
      float* xb = sX + (eg * xbuf + ((k >> 1) & (xbuf - 1))) * (C_::XCHG / 4);
      float* my_up = xb + ((q < 3 ? q : 2) * 2 + 0) * 8 * XP + c * XP;          // slot_up(q): written by lrow 3
      float* my_dn = xb + ((q > 0 ? q - 1 : 0) * 2 + 1) * 8 * XP + c * XP;      // slot_dn(q-1): written by lrow 0
      const float* in_up = xb + ((q > 0 ? q - 1 : 0) * 2 + 0) * 8 * XP + c * XP; // slot_up(q-1): read by lrow 0
      const float* in_dn = xb + ((q < 3 ? q : 2) * 2 + 1) * 8 * XP + c * XP;     // slot_dn(q): read by lrow 3

      // issue every global load of the tile BEFORE waiting for the accumulator: their latency hides behind the MMAs
      const size_t o0 = (valid && fast) ? act_off(n, y, x, 0, a.h, a.w, NCH) : 0;
      uint4 qm[NCH], q1[NCH], q2[NCH];
      if (valid && fast) {
        if (do_mask) {
          const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(a.mask) + o0;
#pragma unroll
          for (int j = 0; j < NCH; ++j) qm[j] = *reinterpret_cast<const uint4*>(p + j * chunk_stride);
        }
        if (do_res1) {
          const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(a.res1) + o0;
#pragma unroll
          for (int j = 0; j < NCH; ++j) q1[j] = *reinterpret_cast<const uint4*>(p + j * chunk_stride);
        }
        if (do_res2) {
          const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(a.res2) + o0;
#pragma unroll
          for (int j = 0; j < NCH; ++j) q2[j] = *reinterpret_cast<const uint4*>(p + j * chunk_stride);
        }
      }
      if (fast && (do_mask || do_res1 || do_res2)) {
        // pull the NEXT tile's residual/mask lines of this group into L2 while this tile is processed
        const int tile2 = tile + tile_stride;
        if (tile2 < g.total_tiles) {
          const int n2 = tile2 / tiles_per_img;
          const int rem2 = tile2 - n2 * tiles_per_img;
          const int ty2 = rem2 / g.tiles_x;
          const int y2 = ty2 * kRowsOut + rho - 1, x2 = (rem2 - ty2 * g.tiles_x) * kTileW + c;
          if (rho >= 1 && rho <= kRowsOut && y2 < a.h && x2 < a.w) {
            const size_t o2 = act_off(n2, y2, x2, 0, a.h, a.w, NCH);
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
              if (do_mask) prefetch_l2(reinterpret_cast<const __nv_bfloat16*>(a.mask) + o2 + j * chunk_stride);
              if (do_res1) prefetch_l2(reinterpret_cast<const __nv_bfloat16*>(a.res1) + o2 + j * chunk_stride);
              if (do_res2) prefetch_l2(reinterpret_cast<const __nv_bfloat16*>(a.res2) + o2 + j * chunk_stride);
            }
          }
        }
      }

      if (tl0) tl_stamp(g, 2, k, 0);
      mbar_wait_relaxed(tfull_bar(as), (k >> 1) & 1);
      tc_fence_after_sync();
      if (tl0) tl_stamp(g, 2, k, 1);
      // phase A: v = bias + P1(rho) + P0(rho-1) + P2(rho+1) from the own warp's lanes (shuffles); the lane-rows at a warp
      // boundary publish what their neighbour warp is missing
      float v[NT];
#pragma unroll
      for (int jj = 0; jj < NT / 16; ++jj) {
        float p0[16], p1[16], p2[16];
        tmem_ld16(taddr + jj * 16, p0);
        tmem_ld16(taddr + NT + jj * 16, p1);
        tmem_ld16(taddr + 2 * NT + jj * 16, p2);
        tmem_ld_wait();
        if (jj == NT / 16 - 1) {
          tc_fence_before_sync();
          mbar_arrive(tempty_bar(as));        // accumulator stage free: this group's next tile may be accumulated
        }
        if (pub_up) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(my_up + jj * 16)[i] = make_float4(p0[4 * i], p0[4 * i + 1], p0[4 * i + 2], p0[4 * i + 3]);
        }
        if (pub_dn) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(my_dn + jj * 16)[i] = make_float4(p2[4 * i], p2[4 * i + 1], p2[4 * i + 2], p2[4 * i + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float up = __shfl_up_sync(0xffffffffu, p0[i], 8);     // P0 of the lane-row below (rho-1)
          float dn = __shfl_down_sync(0xffffffffu, p2[i], 8);   // P2 of the lane-row above (rho+1)
          if (lrow == 0) up = 0.f;
          if (lrow == 3) dn = 0.f;
          v[jj * 16 + i] = p1[i] + up + dn;
        }
      }
      if (tl0) tl_stamp(g, 2, k, 2);
      // bias: 12 broadcast LDS.128 whose latency overlaps the barrier (registers are too scarce to keep it resident)
      float4 bq[NT / 4];
#pragma unroll
      for (int i = 0; i < NT / 4; ++i) bq[i] = reinterpret_cast<const float4*>(sBias)[i];
      named_bar_sync(1 + eg, 128);
      if (tl0) tl_stamp(g, 3, k, 0);
#pragma unroll
      for (int i = 0; i < NT / 4; ++i) {
        v[4 * i] += bq[i].x; v[4 * i + 1] += bq[i].y; v[4 * i + 2] += bq[i].z; v[4 * i + 3] += bq[i].w;
      }
      // phase B: boundary lane-rows add the neighbour warp's term
      if (need_up) {
#pragma unroll
        for (int i = 0; i < NT / 4; ++i) {
          const float4 u = reinterpret_cast<const float4*>(in_up)[i];
          v[4 * i] += u.x; v[4 * i + 1] += u.y; v[4 * i + 2] += u.z; v[4 * i + 3] += u.w;
        }
      }
      if (need_dn) {
#pragma unroll
        for (int i = 0; i < NT / 4; ++i) {
          const float4 u = reinterpret_cast<const float4*>(in_dn)[i];
          v[4 * i] += u.x; v[4 * i + 1] += u.y; v[4 * i + 2] += u.z; v[4 * i + 3] += u.w;
        }
      }
      if (valid) {
        if (fast) {
          __nv_bfloat16* po = reinterpret_cast<__nv_bfloat16*>(a.out) + o0;
#pragma unroll
          for (int j = 0; j < NCH; ++j) {
            float* vj = v + 8 * j;
            if (!unit_scale) {
#pragma unroll
              for (int i = 0; i < 8; ++i) vj[i] *= a.res_scale;
            }
            if (do_relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) vj[i] = fmaxf(vj[i], 0.f);
            }
            if (do_mask) {
              const uint32_t w4[4] = {qm[j].x, qm[j].y, qm[j].z, qm[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                vj[2 * e] = (bf16_lo(w4[e]) > 0.f) ? vj[2 * e] : 0.f;
                vj[2 * e + 1] = (bf16_hi(w4[e]) > 0.f) ? vj[2 * e + 1] : 0.f;
              }
            }
            if (do_res1) {
              const uint32_t w4[4] = {q1[j].x, q1[j].y, q1[j].z, q1[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) { vj[2 * e] += bf16_lo(w4[e]); vj[2 * e + 1] += bf16_hi(w4[e]); }
            }
            if (do_res2) {
              const uint32_t w4[4] = {q2[j].x, q2[j].y, q2[j].z, q2[j].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) { vj[2 * e] += bf16_lo(w4[e]); vj[2 * e + 1] += bf16_hi(w4[e]); }
            }
            store8(po + j * chunk_stride, vj);
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < NT / 16; ++jj) {
            float t[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = v[jj * 16 + i];
            loss += conv_epilogue16<__nv_bfloat16, false>(a, n, y, x, jj * 16, t);
          }
        }
      }
      if (tl0) tl_stamp(g, 2, k, 3);
      if (xbuf == 1) named_bar_sync(3 + eg, 128);   // single exchange buffer: everyone has read it before it is rewritten
    }
    if (a.truth_hr != nullptr && a.loss_sum != nullptr) {
      loss = warp_sum(loss);
      if (lane == 0) atomicAdd(a.loss_sum, static_cast<double>(loss));
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int CIN, int NT, int NSTAGE, int EPI>
static int launch_epi(const lv_conv_args& a, const ConvGeom& g, int max_ctas, cudaStream_t stream) {
  using C_ = Cfg<CIN, NT, NSTAGE>;
  const int xbuf = C_::smem_bytes(a.num_src, g.cout_pad, 2) <= kMaxSmem ? 2 : 1;
  const size_t smem = C_::smem_bytes(a.num_src, g.cout_pad, xbuf);
  auto kern = conv3x3_tc_ky_kernel<CIN, NT, NSTAGE, EPI>;
  static size_t configured = 0;
  if (smem > configured) {
    LV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  long long ctas = max_ctas > 0 ? max_ctas : sm_count();
  if (ctas > g.total_tiles) ctas = g.total_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, a, g, xbuf));
  count_launch();
  return LV_OK;
}

// straight-line epilogues for the flag combinations the engines use; everything else takes the runtime-flag variant
template <int CIN, int NT, int NSTAGE>
static int launch(const lv_conv_args& a, const ConvGeom& g, int max_ctas, cudaStream_t stream) {
  if (a.epilogue == LV_EPI_NHWC && a.res_scale == 1.0f && a.cout == g.cout_pad) {
    const int e = (a.relu ? 1 : 0) | (a.mask ? 2 : 0) | (a.res1 ? 4 : 0) | (a.res2 ? 8 : 0);
    switch (e) {
      case 0: return launch_epi<CIN, NT, NSTAGE, 0>(a, g, max_ctas, stream);
      case 1: return launch_epi<CIN, NT, NSTAGE, 1>(a, g, max_ctas, stream);
      case 2: return launch_epi<CIN, NT, NSTAGE, 2>(a, g, max_ctas, stream);
      case 4: return launch_epi<CIN, NT, NSTAGE, 4>(a, g, max_ctas, stream);
      case 12: return launch_epi<CIN, NT, NSTAGE, 12>(a, g, max_ctas, stream);
      default: break;
    }
  }
  return launch_epi<CIN, NT, NSTAGE, -1>(a, g, max_ctas, stream);
}

}  // namespace ky

// ky-stacked weights: cout must be one N tile (3*cout_pad <= 256)
int conv3x3_tc_ky(const lv_conv_args& a, int max_ctas, cudaStream_t stream) {
  ConvGeom g;
  g.timeline = g_timeline;
  g.cout_pad = (a.cout + 15) / 16 * 16;
  g.nt = g.cout_pad;
  g.ntiles_n = 1;
  g.tiles_x = (a.w + ky::kTileW - 1) / ky::kTileW;
  g.tiles_y = (a.h + ky::kRowsOut - 1) / ky::kRowsOut;
  const long long tt = static_cast<long long>(a.n) * g.tiles_x * g.tiles_y;
  if (tt == 0) return LV_OK;
  LV_CHECK_ARG(tt < (1ll << 31), "conv3x3: too many tiles (%lld)", tt);
  g.total_tiles = static_cast<int>(tt);
  constexpr size_t kMax = ky::kMaxSmem;
#define LV_KY_CASE(CI, NTV)                                                                               \
  if (a.cin == CI && g.cout_pad == NTV) {                                                                 \
    if (ky::Cfg<CI, NTV, 4>::smem_bytes(a.num_src, g.cout_pad, 2) <= kMax) return ky::launch<CI, NTV, 4>(a, g, max_ctas, stream); \
    if (ky::Cfg<CI, NTV, 3>::smem_bytes(a.num_src, g.cout_pad, 2) <= kMax) return ky::launch<CI, NTV, 3>(a, g, max_ctas, stream); \
    if (ky::Cfg<CI, NTV, 2>::smem_bytes(a.num_src, g.cout_pad, 1) <= kMax) return ky::launch<CI, NTV, 2>(a, g, max_ctas, stream); \
  }
  LV_KY_CASE(48, 48)
  LV_KY_CASE(64, 64)
#undef LV_KY_CASE
  set_error("conv3x3 ky-stacked tensor-core path: unsupported shape cin=%d x %d, cout=%d", a.cin, a.num_src, a.cout);
  return LV_ERR_INVALID;
}

}  // namespace lv
