// Row-marching 3x3 convolution on the sm_100a tensor cores: one layer or a whole chain of layers per persistent launch.
//
// Why a second conv kernel: conv_tc.cu / conv_chain.cu issue 27 MMAs of N = 48 per 128-pixel tile (one per tap and 16-channel
// K step).  An M=128, K=16 SS MMA costs max(N/2, 32 + N/4) clk (profiles/r01_umma_microbench.txt): at N = 48 that is 45 clk,
// of which 32 are the 4 KB A-tile read from shared memory and only 24 tensor-pipe work -- the tile is shared-memory bound
// at a 53 % tensor ceiling because every input pixel is read NINE times.  Here every input pixel is read THREE times:
//
//   * M = 128 consecutive pixels of ONE image row ("strip").  The batch's rows are laid on a line with one zero pad pixel
//     between images (pitch W+1; the pad is the left AND right zero padding of the two images it separates), so 48-px
//     training patches fill the 128 lanes as well as 480-px frames do.
//   * B = the three vertical taps stacked along N (LV_W_KY_STACKED: [kx][cin/8][ky*cout + co][8], N = 3*cout = 144): the
//     MMAs of INPUT row y produce, side by side in TMEM, its contributions to OUTPUT rows y+1 (ky=0), y (ky=1), y-1 (ky=2).
//   * The accumulators of the output rows live in a ring of TMEM column blocks (10 x 48 or 8 x 64 columns), consecutive
//     rows in ADJACENT blocks in descending column order.  The N=144 MMA of input row y therefore lands on the blocks of
//     rows y+1, y, y-1 at once and the sum over the vertical taps happens inside the tensor core: no shuffle epilogue
//     (conv_tc_ky.cu's problem), 9 MMAs of 73 clk per row (tensor bound) instead of 27 of 45.  An output row is complete --
//     and drained by an epilogue group while the MMA warp marches on -- once input row y+1 has been issued.
//     A block is (re)initialised by a tiny MMA of its own -- D = ones[128 x 16] x bias_tile[16 x cout], overwrite mode -- issued
//     before the first row that touches it: the conv bias (split into bf16 hi + lo parts in two K columns, ~fp32 exact)
//     is then already in the accumulator, every other MMA accumulates, and the epilogue has no bias to fetch (any load
//     or shuffle there queues behind the MMAs' operand reads: ~300 clk each, measured).
//
// Work unit ("job") = one strip x `rows_per_job` rows; a job of R rows reads R+2 input rows, each ONCE, through a ring of
// row buffers ([8-channel chunk][130 px][16 B] = the SWIZZLE_NONE K-major A operand; the horizontal tap is a 16-byte shift of
// the descriptor start).  Jobs of one layer are dealt round-robin to the persistent CTAs (one per SM).
//
// Chains (count > 1): same data-flow protocol as conv_chain.cu -- done[job] counts finished layers, a job of layer l
// starts when its 3x3 job neighbourhood has finished layer l-1, a publisher warp turns "all rows of the job stored" into
// one red.release.gpu; next-layer weights are prefetched into a shared-memory ring.
//
// Epilogues are the chain kernels' (chain_epilogue.cuh): one pixel x NT channels per thread straight from TMEM to
// global memory; a warp writes 32 consecutive pixels = 512 contiguous bytes per 8-channel chunk.
// (included by conv_row.cu -- TMA row producer, 12 warps -- and conv_row_cp.cu -- cp.async row producers, 13 warps)
#include "chain_epilogue.cuh"
#include "conv_epilogue.cuh"
#include "lv_common.cuh"

// 48-channel build: 8 row buffers + 2 weight buffers (190 KB).  The third weight buffer of the first version bought
// nothing; eight instead of five row buffers cut the rows the scheduler finds "not staged yet" from 37 % to 21 %.
#ifndef LV_ROW_STAGES48
#define LV_ROW_STAGES48 8
#define LV_ROW_WBUFS48 2
#endif
// 64-channel build: 2 weight buffers (147 KB) + 4 row buffers = 224 KB.  That gives up the L1 carve-out, but two row
// buffers starve the pipeline more than the missing L1 costs (EDSR 1080p frame 0.923 -> 0.857 ms, batch of 4: 3.28 -> 2.84)
#ifndef LV_ROW_STAGES64
#define LV_ROW_STAGES64 4
#endif
#ifndef LV_ROW_DBG_NOSPLIT
#define LV_ROW_DBG_NOSPLIT 0
#endif
#ifndef LV_ROW_DBG_NOEPI
#define LV_ROW_DBG_NOEPI 0
#endif
#ifndef LV_ROW_DBG_NOPROD
#define LV_ROW_DBG_NOPROD 0
#endif

namespace lv {

extern int g_use_pdl;
extern long long* g_timeline;

namespace LV_ROW_NS {

constexpr int kMaxLayers = 96;   // LV_CHAIN_MAX_LAYERS
constexpr int kLanes = 128, kRowPx = kLanes + 2;
// Warp roles: 0-7 epilogue (two groups of 4 warps), 8 MMA issuer, 9 scheduler, 10(-11) row producer(s), last: publisher.
// Why a scheduler warp: the tensor pipe's instruction queue holds only ~4 MMAs (~270 clk of work), and ONE thread that
// waits for the accumulator block and the input row, computes the row's descriptors, issues and commits spends ~1,100 clk
// per row outside the issue loop (measured: 2,330 clk per row for 680 clk of MMAs, tensor pipe 29 % busy; the 16x8-tile
// kernels pay the same ~1,000 clk per tile).  So everything but the issue itself moves to a second thread: the scheduler
// waits, computes, and hands the issuer a ready-made command (descriptors, accumulate flags, barriers to commit to)
// through a 16-deep shared-memory ring guarded by two counters (`published`, `consumed`): the issuer takes every row
// that is ready with ONE poll and fetches the next command while the current row's MMAs sit in the tensor queue.
// Row loads: 1-D TMA bulk copies (one per 8-channel chunk and per run of consecutive pixels of one image, pad pixels
// from a zero page) issued by ONE producer warp -- against 26 warp-wide cp.async instructions per row on two warps they
// take the row fill off the LSU path (which the MMAs' operand reads starve, see below) and give a warp back: 12 warps
// = 168 registers per thread instead of 13 warps = 128.  -DLV_ROW_CPASYNC builds the cp.async producers (conv_row_cp.cu).
// Two measured properties of the tensor pipe shape the issue order (tools/probes/row_probe.cu, tools/row_trace.py):
//   * consecutive MMAs into the SAME accumulator address pipeline at the nominal rate (N=144: 73 clk), and moving to a
//     new address once per row is free, but ALTERNATING between two addresses costs ~110 clk per switch -- rows whose
//     three targets straddle the ring's wrap issue their two runs one after the other, never interleaved (interleaved
//     they cost +2,000 clk per such row, i.e. 40 % of the kernel's time);
//   * an N=144 MMA reads 8.6 KB of operands in 73 clk = 92 % of the shared-memory port, so while MMAs execute every other
//     warp's shared-memory / LSU instruction waits 100-300 clk: all other roles are written to need few of them.
#ifdef LV_ROW_CPASYNC
constexpr bool kTmaRows = false;
#else
constexpr bool kTmaRows = true;
#endif
#ifndef LV_ROW_EPI_GROUPS
#define LV_ROW_EPI_GROUPS 2
#endif
constexpr int kEpiGroups = LV_ROW_EPI_GROUPS;             // groups of 4 warps (one per TMEM lane quarter); output row kk
                                                          // is drained by group kk % kEpiGroups
constexpr int kEpiWarps = 4 * kEpiGroups, kEpiThreads = kEpiWarps * 32, kProdThreads = kTmaRows ? 32 : 64;
constexpr int kMmaWarp = kEpiWarps;                       // 8
constexpr int kSchedWarp = kMmaWarp + 1;                  // 9
constexpr int kProdWarp0 = kSchedWarp + 1;                // 10 (, 11)
constexpr int kPubWarp = kProdWarp0 + kProdThreads / 32;  // 11 (12)
constexpr int kThreads = (kPubWarp + 1) * 32;             // 384 (416)
constexpr int kCmdSlots = 16;

__device__ uint4 g_zero_page[kLanes + 2];                 // source of the pad / out-of-range pixels' bulk copies (zeros)

// one row of MMA work, written by the scheduler thread, read by the issuer thread
struct __align__(16) RowCmd {
  uint32_t a_lo, b_lo;                 // low words of the A / B shared-memory descriptors (start address + LBO fields)
  uint32_t ra_d, ra_b;                 // run A: TMEM address, weight-row offset
  uint32_t ra_i;                       // run A: instruction descriptor (N = cout x targets)
  uint32_t rb_d, rb_b, rb_i;           // run B (only when the accumulator ring wraps inside the row's targets), rb_i == 0: none
  uint32_t f0_d, f1_d;                 // blocks this row touches first: initialised with the bias MMA (kNoBlock: none)
  uint32_t bars;                       // barrier INDICES (8 bits each; +1, 0 = none, for all but the first):
                                       //   [0:8) row buffer free, [8:16) / [16:24) accumulator complete (output row yi-1 /
                                       //   row yi on the image's bottom row), [24:32) weight buffer free (layer's last row)
  uint32_t bias_lo;                    // low word of the bias tile's descriptor
  uint32_t pad[4];
};
static_assert(sizeof(RowCmd) == 64, "RowCmd is read with 16-byte shared-memory loads");
constexpr uint32_t kNoBlock = 0xffffffffu;
constexpr uint32_t kCmdStop = 0xffffffffu;   // a_lo of the terminating command

struct Params {
  lv_conv_args layer[kMaxLayers];
};

struct Geom {
  int N, H, W, P;           // P = W + 1: pitch of one image on the line
  int nstrips, rows_per_job, nblocks, total_jobs;
  long long* stats;         // debug (lv_debug_set_timeline): per-role cycle counters of CTA 0, nullptr in production
};

// debug counters (CTA 0 only): [role * 8 + i]; role 0 = MMA thread, 1 = producer thread 0, 2 = epilogue warp 0 lane 0,
// 3 = epilogue warp 4 lane 0
__device__ __forceinline__ void stat_add(const Geom& g, int slot, long long v) {
  if (g.stats != nullptr && blockIdx.x == 0) g.stats[slot] += v;
}
__device__ __forceinline__ long long stat_clk(const Geom& g) { return g.stats != nullptr ? clock64() : 0; }
// Developer trace (-DLV_ROW_TRACE=1, tools/row_trace.py): absolute clock64() stamps of CTA 0's roles for rows
// [kTraceFirst, kTraceFirst + kTraceRows) of the CTA's row counters, behind the 64 counters of the stats buffer.
#ifndef LV_ROW_TRACE
#define LV_ROW_TRACE 0
#endif
constexpr uint32_t kTraceFirst = 100, kTraceRows = 64;
__device__ __forceinline__ void trace(const Geom& g, int role, uint32_t idx, int ev) {
#if LV_ROW_TRACE
  if (g.stats != nullptr && blockIdx.x == 0 && idx - kTraceFirst < kTraceRows)
    g.stats[64 + (role * kTraceRows + (idx - kTraceFirst)) * 4 + ev] = clock64();
#endif
}

// Shared-memory budget: stay below the 196 KB carve-out step so that the SM keeps ~32 KB of L1 (kernel parameters,
// bias vectors, whatever the epilogue spills): with the 228 KB step nothing is left and every such access goes to L2.
template <int CIN, int NT, int NSTAGE, int WBUFS>
struct Cfg {
  static constexpr int CH = CIN / 8;
  static constexpr int KSTEPS = CIN / 16;
  static constexpr int N3 = 3 * NT;
  static constexpr int A_PLANE = kRowPx * 16;        // 2080 B: one 8-channel chunk of a row buffer
  static constexpr int A_STAGE = CH * A_PLANE;
  static constexpr int W_PLANE = N3 * 16;            // one 8-channel chunk of one kx block of the weights
  static constexpr int W_KX = CH * W_PLANE;
  static constexpr int W_LAYER = 3 * W_KX;
  static constexpr int RING = (512 / NT) & ~1;       // accumulator blocks (even: block parity == epilogue group)
  static constexpr int TMEM_COLS = 512;
  static constexpr int PIECES = (kRowPx * CH + kProdThreads - 1) / kProdThreads;
  static constexpr int ONES_TILE = 2 * kLanes * 16;  // A operand of the bias MMA: [2 K-halves][128 rows][16 B], (1,1,0,..) / 0
  static constexpr int BIAS_TILE = 2 * NT * 16;      // B operand: [2 K-halves][cout rows][16 B], row n = (hi(b_n), lo(b_n), 0,..)
  static constexpr int NBARS = 2 * NSTAGE + 2 * RING + 3 * WBUFS + 4;
  static constexpr size_t smem_bytes() {
    return static_cast<size_t>(WBUFS) * (W_LAYER + BIAS_TILE) + ONES_TILE + static_cast<size_t>(NSTAGE) * A_STAGE +
           kCmdSlots * sizeof(RowCmd) + NBARS * 8 + 64;
  }
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// poll relaxed, finish with one acquire (an acquire load invalidates the SM's L1, see conv_chain.cu)
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t need) {
  if (ld_acquire_gpu(p) >= need) return;
  uint32_t spins = 0;
  while (ld_relaxed_gpu(p) < need) {
    __nanosleep(32);
    if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
  }
  (void)ld_acquire_gpu(p);
}

__device__ __forceinline__ uint32_t ld_acquire_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(uint32_t addr, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& t) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 t;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(addr) : "memory");
  return t;
}

// runtime-N instruction descriptor (M = 128, bf16 x bf16 -> fp32, both operands K-major)
__device__ __forceinline__ uint32_t idesc_n(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
}

struct Job {
  int u, y0, y1;
};

struct EpiCtx {
  Geom g;
  const uint32_t* done;
  volatile uint32_t* pub_seen;
  uint32_t tfull0, tempty0, pub_bar0, tmem_lane;   // barrier ring bases, TMEM address of this warp's lane quarter
  int G, lane, eg, m, nlayers;
  size_t chunk_stride, row_stride;
};

constexpr int kKindPs4 = 100, kKindGeneric = -1;

// Planar epilogue of one output row for one thread (one pixel x NT channels), straight-line for a compile-time flag set
// EPI (bit0 ReLU, bit1 ReLU-mask, bit2 res1, bit3 res2).  Same arithmetic as chain::fast_tile, but the accumulator is read
// and processed in two halves of NT/2 channels and the bias is already in the accumulator (bias MMA): ~100 live registers
// instead of ~170, so the 416-thread CTA (128 registers per thread) does not spill in its hot loop.  All operand loads of the row are issued
// before the accumulator wait (their latency hides behind the MMAs).
template <int EPI, int NT>
__device__ __forceinline__ void row_tile(const chain::FastEpi& e, bool valid, size_t o0,
                                         size_t chunk_stride, uint32_t taddr, uint32_t tfull, uint32_t tempty, uint32_t parity,
                                         long long* dbg = nullptr) {
  long long d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0;
  if (dbg != nullptr) d0 = clock64();
  constexpr int NCH = NT / 8, HC = NCH / 2, HALF = NT / 2;
  constexpr bool do_relu = (EPI & 1) != 0, do_mask = (EPI & 2) != 0, do_res1 = (EPI & 4) != 0, do_res2 = (EPI & 8) != 0;
  uint4 qm[do_mask ? NCH : 1], q1[do_res1 ? NCH : 1], q2[do_res2 ? NCH : 1];
  if (valid) {
    if constexpr (do_mask) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) qm[j] = chain::ldcg16(e.mask + o0 + j * chunk_stride);
    }
    if constexpr (do_res1) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) q1[j] = chain::ldcg16(e.res1 + o0 + j * chunk_stride);
    }
    if constexpr (do_res2) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) q2[j] = chain::ldcg16(e.res2 + o0 + j * chunk_stride);
    }
  }
  if (dbg != nullptr) d1 = clock64();
  mbar_wait_relaxed(tfull, parity);
  tc_fence_after_sync();
  if (dbg != nullptr) d2 = clock64();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[HALF];
    if constexpr (HALF == 24) {
      tmem_ld16(taddr + h * HALF, v);
      tmem_ld8(taddr + h * HALF + 16, v + 16);
    } else {
#pragma unroll
      for (int j = 0; j < HALF / 16; ++j) tmem_ld16(taddr + h * HALF + j * 16, v + j * 16);
    }
    tmem_ld_wait();
    if (dbg != nullptr) { if (h == 0) d3 = clock64(); else d4 += clock64(); }
    if (h == 1) {
      tc_fence_before_sync();
      mbar_arrive(tempty);   // accumulator block free: the MMAs of a later row may overwrite it
    }
    if (dbg != nullptr && h == 1) d4 -= 0;
    if (valid) {
      __nv_bfloat16* po = e.out + o0;
#pragma unroll
      for (int jj = 0; jj < HC; ++jj) {
        const int j = h * HC + jj;
        float* vj = v + 8 * jj;
        if constexpr (do_relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) vj[i] = fmaxf(vj[i], 0.f);
        }
        if constexpr (do_mask) {
          const uint32_t w4[4] = {qm[j].x, qm[j].y, qm[j].z, qm[j].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            vj[2 * t] = (bf16_lo(w4[t]) > 0.f) ? vj[2 * t] : 0.f;
            vj[2 * t + 1] = (bf16_hi(w4[t]) > 0.f) ? vj[2 * t + 1] : 0.f;
          }
        }
        if constexpr (do_res1) {
          const uint32_t w4[4] = {q1[j].x, q1[j].y, q1[j].z, q1[j].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) { vj[2 * t] += bf16_lo(w4[t]); vj[2 * t + 1] += bf16_hi(w4[t]); }
        }
        if constexpr (do_res2) {
          const uint32_t w4[4] = {q2[j].x, q2[j].y, q2[j].z, q2[j].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) { vj[2 * t] += bf16_lo(w4[t]); vj[2 * t + 1] += bf16_hi(w4[t]); }
        }
        store8(po + j * chunk_stride, vj);
      }
    }
    if (dbg != nullptr && h == 0) d4 = -clock64() + 0;   // start of the second half: d4 becomes (t_ld1_done - t_half0_done)
  }
  if (dbg != nullptr) {
    const long long d5 = clock64();
    dbg[0] += d1 - d0;        // operand loads issued
    dbg[1] += d2 - d1;        // (second) accumulator wait: ~0 in stats mode
    dbg[2] += d3 - d2;        // first TMEM load + wait
    dbg[3] += d4;             // second TMEM load + wait (from the end of the first half's stores)
    dbg[4] += d5 - d0;        // whole tile function
  }
}

// All jobs of one layer for one epilogue thread (one line position = one pixel column of the strip).
template <int KIND, int NT, int RING>
__device__ __forceinline__ uint32_t run_layer(const EpiCtx& cx, const lv_conv_args& a, int l, int first, uint32_t k) {
  constexpr int NCH = NT / 8;
  const bool has_ops = (a.mask != nullptr) || (a.res1 != nullptr) || (a.res2 != nullptr);
  chain::FastEpi fe;
  fe.mask = reinterpret_cast<const __nv_bfloat16*>(a.mask);
  fe.res1 = reinterpret_cast<const __nv_bfloat16*>(a.res1);
  fe.res2 = reinterpret_cast<const __nv_bfloat16*>(a.res2);
  fe.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  fe.res_scale = a.res_scale;
  fe.relu = a.relu;
  float loss = 0.f;
  long long st_wait = 0, st_drain = 0, st_rows = 0;   // debug counters (registers; flushed once per layer)
  long long dbgv[5] = {0, 0, 0, 0, 0};
  for (int job = first; job < cx.g.total_jobs; job += cx.G) {
    const int bk = job / cx.g.nstrips, u = job - bk * cx.g.nstrips;
    const int y0 = bk * cx.g.rows_per_job;
    const int y1 = min(cx.g.H, y0 + cx.g.rows_per_job);
    const int p = u * kLanes + cx.m;
    const int n = p / cx.g.P, x = p - n * cx.g.P;
    const bool valid = (n < cx.g.N) && (x < cx.g.W);
    const size_t o_img = valid ? act_off(n, 0, x, 0, cx.g.H, cx.g.W, NCH) : 0;
    if (l > 0 && (KIND == kKindGeneric || has_ops)) {
      // same-position operands come from earlier layers of this chain, possibly written by another CTA
      if (cx.lane == 0) wait_flag(cx.done + job, static_cast<uint32_t>(l));
      __syncwarp();
    }
    for (int yo = y0; yo < y1; ++yo) {
      const uint32_t kk = k + static_cast<uint32_t>(yo - y0);
      if (kk % static_cast<uint32_t>(kEpiGroups) != static_cast<uint32_t>(cx.eg)) continue;
      const uint32_t blk = kk % RING;
      const uint32_t par = (kk / RING) & 1u;
      const uint32_t taddr = cx.tmem_lane + (RING - 1 - blk) * NT;
      const uint32_t tfull = cx.tfull0 + 8u * blk, tempty = cx.tempty0 + 8u * blk;
      const size_t o0 = o_img + static_cast<size_t>(yo) * cx.row_stride;
      const uint32_t seen = (cx.nlayers > 1) ? *cx.pub_seen : 0u;
      const bool st = cx.g.stats != nullptr && cx.lane == 0 && (cx.m == 0) && cx.eg < 2;
      long long c0 = 0, c1 = 0;
      if (cx.g.stats != nullptr) {       // debug only: separate "waiting for the accumulator" from "draining it"
        c0 = clock64();
        if (st) trace(cx.g, 3 + cx.eg, kk, 0);
        mbar_wait_relaxed(tfull, par);
        c1 = clock64();
        if (st) trace(cx.g, 3 + cx.eg, kk, 1);
      }
      if constexpr (KIND == kKindPs4) {
        loss += chain::ps4_tile<NT>(a, nullptr, valid, n, yo, x, cx.g.H, cx.g.W, o0, cx.chunk_stride, taddr, tfull, tempty, par);
      } else if constexpr (KIND == kKindGeneric) {
        mbar_wait_relaxed(tfull, par);
        tc_fence_after_sync();
#pragma unroll 1
        for (int j = 0; j < NT / 16; ++j) {
          float v[16];
          tmem_ld16(taddr + j * 16, v);
          tmem_ld_wait();
          if (valid) loss += conv_epilogue16<__nv_bfloat16, false>(a, n, yo, x, j * 16, v);   // bias: in the accumulator
        }
        tc_fence_before_sync();
        mbar_arrive(tempty);
      } else {
#if LV_ROW_DBG_NOEPI   // timing experiment: no TMEM drain, no global traffic (results are garbage)
        mbar_wait_relaxed(tfull, par);
        tc_fence_after_sync();
        tc_fence_before_sync();
        mbar_arrive(tempty);
#else
        row_tile<KIND, NT>(fe, valid, o0, cx.chunk_stride, taddr, tfull, tempty, par, st ? dbgv : nullptr);
#endif
      }
      if (st) {
        trace(cx.g, 3 + cx.eg, kk, 2);
        st_wait += c1 - c0;
        st_drain += clock64() - c1;
        st_rows += 1;
      }
      if (cx.nlayers > 1) {
        // this warp's quarter of the row is on its way to global memory: hand it to the publisher warp (ring of 4 rows)
        __syncwarp();
        if (cx.lane == 0) {
          if (kk >= 4 && seen + 3u < kk) {
            uint32_t spins = 0;
            while (*cx.pub_seen + 3u < kk) {
              if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
            }
          }
          mbar_arrive(cx.pub_bar0 + 8u * (kk & 3u));
        }
      }
    }
    k += static_cast<uint32_t>(y1 - y0);
  }
  if (a.truth_hr != nullptr && a.loss_sum != nullptr) {
    loss = warp_sum(loss);
    if (cx.lane == 0) atomicAdd(a.loss_sum, static_cast<double>(loss));
  }
  if (cx.g.stats != nullptr && cx.lane == 0 && cx.m == 0 && cx.eg < 2) {
    const int base = 16 + 8 * cx.eg;
    stat_add(cx.g, base + 0, st_wait);
    stat_add(cx.g, base + 1, st_drain);
    stat_add(cx.g, base + 2, st_rows);
    if (cx.eg == 0) for (int i = 0; i < 5; ++i) stat_add(cx.g, 32 + i, dbgv[i]);
  }
  return k;
}

// 13 warps: the register file is split per scheduler (16 K registers each) and one scheduler holds 4 of the 13 warps, so
// the budget is 128 registers per thread; the hot epilogues (row_tile) are written to stay below it (cp.async build;
// the TMA build has 12 warps and 168 registers)
template <int CIN, int NT, int NSTAGE, int WBUFS>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_row_kernel(const __grid_constant__ Params P, const int nlayers, const Geom g, uint32_t* __restrict__ done, const int rot) {
  using C_ = Cfg<CIN, NT, NSTAGE, WBUFS>;
  constexpr int RING = C_::RING;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint8_t* sW = smem;
  uint8_t* sBias = sW + WBUFS * C_::W_LAYER;                 // [WBUFS] bias tiles, same ring as the weight buffers
  uint8_t* sOnes = sBias + WBUFS * C_::BIAS_TILE;
  uint8_t* sA = sOnes + C_::ONES_TILE;
  RowCmd* cmds = reinterpret_cast<RowCmd*>(sA + NSTAGE * C_::A_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(cmds + kCmdSlots);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + RING + b); };
  auto wfull_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 2 * RING + b); };
  auto wfree_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 2 * RING + WBUFS + b); };
  auto bfull_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 2 * RING + 2 * WBUFS + b); };
  auto pub_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * RING + 3 * WBUFS + s); };
  auto empty_idx = [&](int s) { return static_cast<uint32_t>(NSTAGE + s); };
  auto tfull_idx = [&](uint32_t b) { return static_cast<uint32_t>(2 * NSTAGE) + b; };
  auto wfree_idx = [&](int b) { return static_cast<uint32_t>(2 * NSTAGE + 2 * RING + WBUFS + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C_::NBARS);
  uint32_t* s_last = tmem_slot + 1;
  volatile uint32_t* pub_seen = tmem_slot + 2;   // rows whose completion the publisher warp has observed
  volatile uint32_t* cmd_published = tmem_slot + 4;   // command ring: rows written by the scheduler ...
  volatile uint32_t* cmd_consumed = tmem_slot + 8;    // ... and rows the issuer has taken

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), kTmaRows ? 1 : kProdThreads);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < RING; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kEpiThreads / kEpiGroups);
    }
    for (int b = 0; b < WBUFS; ++b) {
      mbar_init(wfull_bar(b), 1);
      mbar_init(wfree_bar(b), 1);
      mbar_init(bfull_bar(b), kProdThreads);
    }
    for (int s = 0; s < 4; ++s) mbar_init(pub_bar(s), kEpiWarps / kEpiGroups);
    tmem_slot[2] = 0u;
    tmem_slot[4] = 0u;
    tmem_slot[8] = 0u;
    mbar_fence_init();
  }
  // constant A operand of the bias MMA: K columns 0 and 1 are ones (they meet the bias' hi and lo parts), the rest zero;
  // the bias tiles start out all zero (only the first K-half of a row is ever rewritten)
  for (int i = threadIdx.x; i < C_::ONES_TILE / 16; i += kThreads)
    st_shared_v4(smem_u32(sOnes) + i * 16, make_uint4(i < kLanes ? 0x3f803f80u : 0u, 0u, 0u, 0u));
  for (int i = threadIdx.x; i < WBUFS * C_::BIAS_TILE / 16; i += kThreads)
    st_shared_v4(smem_u32(sBias) + i * 16, make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();   // generic-proxy writes above -> UMMA (async proxy) reads
  if (warp == kMmaWarp) tmem_alloc<C_::TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  const int G = static_cast<int>(gridDim.x);
  // first job of this CTA in layer l: job X of layer l belongs to CTA (X + l*rot) mod G (spreads the partial last wave)
  auto first_job = [&](int l) {
    const int sh = static_cast<int>((static_cast<long long>(l) * rot) % G);
    return (static_cast<int>(blockIdx.x) + G - sh) % G;
  };
  auto decode = [&](int job) {
    Job j;
    const int bk = job / g.nstrips;
    j.u = job - bk * g.nstrips;
    j.y0 = bk * g.rows_per_job;
    j.y1 = min(g.H, j.y0 + g.rows_per_job);
    return j;
  };

  if (warp == kPubWarp) {
    // =============================== publisher: GPU-scope release of finished jobs ===============================
    if (lane == 0 && nlayers > 1) {
      uint32_t kk = 0;
      for (int l = 0; l < nlayers; ++l) {
        for (int job = first_job(l); job < g.total_jobs; job += G) {
          const Job j = decode(job);
          for (int yo = j.y0; yo < j.y1; ++yo, ++kk) {
            mbar_wait(pub_bar(kk & 3u), (kk >> 2) & 1u);
            *pub_seen = kk + 1u;
          }
          red_release_gpu_add(done + job, 1u);
        }
      }
    }
    __syncwarp();
  } else if (warp >= kProdWarp0) {
    // =============================== producers: dependency wait + input rows -> smem ===============================
    const int ptid = threadIdx.x - kProdWarp0 * 32;
    const int ddy = lane / 3 - 1, ddx = lane % 3 - 1;   // lanes 0..8 watch the 3x3 job neighbourhood
    const size_t row_stride = static_cast<size_t>(C_::CH) * g.W * 8;   // elements between image rows
    uint32_t fill = 0;
    long long sp_flag = 0, sp_empty = 0, sp_issue = 0, sp_rows = 0;   // debug counters (registers)
    pdl_wait();
    for (int l = 0; l < nlayers; ++l) {
      const __nv_bfloat16* src_base = reinterpret_cast<const __nv_bfloat16*>(P.layer[l].src[0]);
      {
        // this layer's bias tile (B operand of the bias MMA): row n = (hi(b_n), lo(b_n), 0, ...), b_n = hi + lo in bf16.
        // It shares the weight ring's slot: free once the MMAs of layer l - WBUFS have retired (every CTA has at least
        // one job per layer, so that layer's last row did commit `wfree`).
        const int wb = l % WBUFS;
        if (l >= WBUFS) mbar_wait_relaxed(wfree_bar(wb), ((l / WBUFS) - 1) & 1);
        for (int nrow = ptid; nrow < NT; nrow += kProdThreads) {
          const float* bias = P.layer[l].bias;
          const float b = (bias != nullptr && nrow < P.layer[l].cout) ? __ldg(bias + nrow) : 0.f;
          const __nv_bfloat16 hi = __float2bfloat16_rn(b);
          const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
          const uint32_t w = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16);
          st_shared_v4(smem_u32(sBias + wb * C_::BIAS_TILE) + nrow * 16, make_uint4(w, 0u, 0u, 0u));
        }
        fence_proxy_async_smem();
        mbar_arrive(bfull_bar(wb));
      }
      for (int job = first_job(l); job < g.total_jobs; job += G) {
        const Job j = decode(job);
        const long long pc0 = stat_clk(g);
        if (l > 0) {
          if (lane < 9) {
            const int bk = job / g.nstrips;
            const int yy = bk + ddy, xx = j.u + ddx;
            if (yy >= 0 && yy < g.nblocks && xx >= 0 && xx < g.nstrips)
              wait_flag(done + yy * g.nstrips + xx, static_cast<uint32_t>(l));
          }
          __syncwarp();
          // the rows were written by other CTAs' st.global (generic proxy); the bulk copies read through the async proxy
          if constexpr (kTmaRows) asm volatile("fence.proxy.async.global;" ::: "memory");   // the unrestricted fence.proxy.async costs ~1,000 clk here
        }
        sp_flag += stat_clk(g) - pc0;      // dependency (flag) wait
        const int ya = max(j.y0 - 1, 0), yb = min(j.y1 + 1, g.H);
        if constexpr (kTmaRows) {
          // The 130-pixel window of the line = at most THREE runs when images are at least 129 px wide (the launcher sends
          // narrower ones to the cp.async build): [tail of an image | the pad pixel (zeros) | head of the next image], or a
          // zero tail behind the last image.  One bulk copy per run and 8-channel chunk.
          const int p0 = j.u * kLanes - 1;
          int seg_px[3] = {0, 0, 0}, seg_len[3] = {0, 0, 0};
          long long seg_src[3] = {-1, -1, -1};
          {
            int px = 0;
#pragma unroll
            for (int sidx = 0; sidx < 3; ++sidx) {
              if (px < kRowPx) {
                const int pos = p0 + px;
                int len, n = -1, x = 0;
                if (pos < 0) {
                  len = -pos;
                } else {
                  n = pos / g.P;
                  x = pos - n * g.P;
                  if (n >= g.N) { n = -1; len = kRowPx - px; }
                  else if (x >= g.W) { n = -1; len = 1; }
                  else len = g.W - x;
                }
                len = min(len, kRowPx - px);
                seg_px[sidx] = px;
                seg_len[sidx] = len;
                seg_src[sidx] = (n < 0) ? -1 : (static_cast<long long>(n) * g.H * C_::CH * g.W + x) * 8;
                px += len;
              }
            }
          }
          // lane = (run, chunk): at most 3 * CH <= 24 copies per row, one per lane (measured: 584 clk per row against 890
          // when a single lane issues all of them)
          const int my_s = lane / C_::CH, my_c = lane - my_s * C_::CH;
          const bool mine = my_s < 3 && (my_s == 0 ? seg_len[0] : my_s == 1 ? seg_len[1] : seg_len[2]) > 0;
          const int m_px = my_s == 0 ? seg_px[0] : my_s == 1 ? seg_px[1] : seg_px[2];
          const int m_len = my_s == 0 ? seg_len[0] : my_s == 1 ? seg_len[1] : seg_len[2];
          const long long m_src = my_s == 0 ? seg_src[0] : my_s == 1 ? seg_src[1] : seg_src[2];
          const uint32_t m_dst = static_cast<uint32_t>(my_c * C_::A_PLANE + m_px * 16), m_bytes = static_cast<uint32_t>(m_len * 16);
          const long long m_off = (m_src >= 0) ? m_src + static_cast<long long>(my_c) * g.W * 8 : -1;
          for (int yi = ya; yi < yb; ++yi, ++fill) {
            const int stage = fill % NSTAGE;
            const long long pc1 = stat_clk(g);
            if (lane == 0) {
              mbar_wait_relaxed(empty_bar(stage), ((fill / NSTAGE) & 1) ^ 1);
#if LV_ROW_DBG_NOPROD   // timing experiment: rows are never copied (results are garbage)
              mbar_arrive(full_bar(stage));
#else
              mbar_arrive_expect_tx(full_bar(stage), C_::A_STAGE);
#endif
            }
            __syncwarp();
            const long long pc2 = stat_clk(g);
            if (lane == 0) trace(g, 0, fill, 0);
            if (mine && !LV_ROW_DBG_NOPROD) {
              const __nv_bfloat16* src = src_base + static_cast<size_t>(yi) * row_stride;
              tma_bulk_g2s(smem_u32(sA + stage * C_::A_STAGE) + m_dst,
                           m_off >= 0 ? static_cast<const void*>(src + m_off) : static_cast<const void*>(g_zero_page), m_bytes,
                           full_bar(stage));
            }
            sp_empty += pc2 - pc1;                 // waiting for a free row buffer
            if (lane == 0) trace(g, 0, fill, 1);
            sp_issue += stat_clk(g) - pc2;         // issuing the copies
            sp_rows += 1;
          }
        } else {
          // this thread's 16-byte pieces of a row buffer: piece idx = chunk * 130 + px  <->  shared-memory offset idx * 16
          long long pc_off[C_::PIECES];
#pragma unroll
          for (int i = 0; i < C_::PIECES; ++i) {
            const int idx = ptid + i * kProdThreads;
            const int c = idx / kRowPx, px = idx - c * kRowPx;
            const int p = j.u * kLanes - 1 + px;
            long long off = -1;
            if (idx < kRowPx * C_::CH && p >= 0) {
              const int n = p / g.P, x = p - n * g.P;
              if (n < g.N && x < g.W) off = ((static_cast<long long>(n) * g.H * C_::CH + c) * g.W + x) * 8;
            }
            pc_off[i] = off;
          }
          for (int yi = ya; yi < yb; ++yi, ++fill) {
            const int stage = fill % NSTAGE;
            const long long pc1 = stat_clk(g);
            mbar_wait_relaxed(empty_bar(stage), ((fill / NSTAGE) & 1) ^ 1);
            const long long pc2 = stat_clk(g);
            const __nv_bfloat16* src = src_base + static_cast<size_t>(yi) * row_stride;
            const uint32_t dst0 = smem_u32(sA + stage * C_::A_STAGE) + ptid * 16;
#pragma unroll
            for (int i = 0; i < C_::PIECES; ++i) {
              if (ptid + i * kProdThreads < kRowPx * C_::CH) {
                const bool inb = pc_off[i] >= 0;
                cp_async16(dst0 + i * (kProdThreads * 16), inb ? (src + pc_off[i]) : src_base, inb ? 16u : 0u);
              }
            }
            cp_async_mbar_arrive_noinc(full_bar(stage));
            sp_empty += pc2 - pc1;                 // waiting for a free row buffer
            sp_issue += stat_clk(g) - pc2;         // issuing the copies
            sp_rows += 1;
          }
        }
      }
    }
    cp_async_wait<0>();
    if (ptid == 0) {
      stat_add(g, 8, sp_flag); stat_add(g, 9, sp_empty); stat_add(g, 10, sp_issue); stat_add(g, 11, sp_rows);
    }
  } else if (warp == kSchedWarp) {
    // =============================== scheduler: waits + descriptors -> command ring ================================
    // One thread.  Under MMA load every shared-memory round trip of another warp (an mbarrier probe, a load) costs
    // 150-300 clk, so the loop is arranged to pay ONE per row: the row's barrier probes are issued first and their
    // results looked at only after the descriptor arithmetic; commands are handed over through a pair of counters
    // (`published` here, `consumed` in the issuer) instead of a barrier per slot, so that the issuer can take every row
    // that is ready with one poll.
    if (lane == 0) {
      auto load_weights = [&](int l) {
        const int b = l % WBUFS;
        mbar_arrive_expect_tx(wfull_bar(b), C_::W_LAYER);
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.layer[l].weights);
        for (int t = 0; t < 3 * C_::CH; ++t)
          tma_bulk_g2s(smem_u32(sW + b * C_::W_LAYER + t * C_::W_PLANE), wsrc + static_cast<size_t>(t) * C_::W_PLANE,
                       C_::W_PLANE, wfull_bar(b));
      };
      load_weights(0);   // packed weights are never written while a launch chain is in flight: no pdl_wait needed
      uint32_t k = 0, ncmd = 0, cons = 0, fph = 0;   // fph: parity of the row buffer's current use
      int stage = 0;
      const uint32_t a_lo0 = static_cast<uint32_t>(umma_smem_desc(smem_u32(sA), C_::A_PLANE, 128));
      const uint32_t cmd0 = smem_u32(cmds);
      const uint32_t pub_addr = smem_u32(const_cast<uint32_t*>(cmd_published)), cons_addr = smem_u32(const_cast<uint32_t*>(cmd_consumed));
      long long ss_setup = 0, ss_wait = 0, ss_slot = 0, ss_write = 0, ss_rows = 0, ss_nf = 0, ss_nt = 0;   // debug counters
      for (int l = 0; l < nlayers; ++l) {
        const int wb = l % WBUFS;
        if (l + 1 < nlayers) {
          // the next layer's weights go to the slot of layer l + 1 - WBUFS: wait until its MMAs have retired (every CTA
          // has at least one job per layer: grid <= jobs, so each layer's last row commits `wfree` exactly once)
          const int nb = (l + 1) % WBUFS;
          if (l + 1 >= WBUFS) mbar_wait(wfree_bar(nb), (((l + 1) / WBUFS) - 1) & 1);
          load_weights(l + 1);
        }
        mbar_wait(wfull_bar(wb), (l / WBUFS) & 1);
        mbar_wait(bfull_bar(wb), (l / WBUFS) & 1);
        const uint32_t b_lo = static_cast<uint32_t>(umma_smem_desc(smem_u32(sW + wb * C_::W_LAYER), C_::W_PLANE, 128));
        const uint32_t bias_lo = static_cast<uint32_t>(umma_smem_desc(smem_u32(sBias + wb * C_::BIAS_TILE), NT * 16, 128));
        // the layer's last row of this CTA also commits the "weight buffer free" barrier: find it first
        int last_job = -1;
        for (int job = first_job(l); job < g.total_jobs; job += G) last_job = job;
        for (int job = first_job(l); job < g.total_jobs; job += G) {
          const Job j = decode(job);
          const int ya = max(j.y0 - 1, 0), yb = min(j.y1 + 1, g.H);
          // Accumulator-ring position of "target 0" (output row yi + 1) of the job's first input row: counter kk0, block
          // m0 = kk0 % RING at column (RING-1 - m0) * NT, use parity ph0 = (kk0 / RING) & 1; all of it advances by one per
          // row, so the loop below carries it along instead of dividing (this thread's ALU time is on the critical path:
          // 400 clk of setup per row held the whole CTA at ~1,100 clk per row)
          const uint32_t kk_first = k + static_cast<uint32_t>(ya + 1 - j.y0);
          uint32_t m0 = kk_first % RING, ph0 = (kk_first / RING) & 1u;
          for (int yi = ya; yi < yb; ++yi, ++ncmd) {
            const long long c0 = stat_clk(g);
            trace(g, 1, ncmd, 0);
            // the row is staged; blocks that get their FIRST contribution from this input row (target 0 always; target 1
            // too on the image's top row) have been drained by the epilogue (their previous output row): probe the three
            // barriers now, look at the answers after the arithmetic below
            const uint32_t m1 = m0 ? m0 - 1u : RING - 1u, ph1 = m0 ? ph0 : ph0 ^ 1u;
            const uint32_t m2 = m1 ? m1 - 1u : RING - 1u;
            // targets t = 0,1,2: output row yi+1-t through vertical tap ky = t (weight rows [t*NT, (t+1)*NT)); the valid
            // ones (row inside the job) form the interval [ta, tb]
            const int ta = max(0, yi + 2 - j.y1), tb = min(2, yi + 1 - j.y0);
            const bool fresh0 = ta == 0, fresh1 = yi == 0 && ta <= 1 && tb >= 1;
            const uint32_t bar_t0 = tempty_bar(m0), bar_t1 = tempty_bar(m1), bar_f = full_bar(stage);
            bool ok_t0 = !fresh0 || mbar_try_wait(bar_t0, ph0 ^ 1u);
            bool ok_t1 = !fresh1 || mbar_try_wait(bar_t1, ph1 ^ 1u);
            bool ok_f = mbar_try_wait(bar_f, fph);
            // Runs = maximal groups of valid targets issued as ONE MMA (blocks of targets t, t+1 are adjacent, ascending,
            // unless target t sits in block 0 = the ring wraps there; N = NT * targets): A = [ta..ea], B = [ea+1..tb]
            // (only when the ring wraps inside the interval); fresh blocks are initialised by the bias MMA
            const uint32_t m_ta = ta == 0 ? m0 : ta == 1 ? m1 : m2;
            int ea = ta;
            if (ea < tb && (ea == 0 ? m0 : m1) != 0u) ++ea;
            if (ea < tb && (ea == 0 ? m0 : m1) != 0u) ++ea;
            const bool has_b = ea < tb;
            const uint32_t m_eb = ea == 0 ? m1 : m2;                      // first block of run B
            const uint32_t tfa = tb == 2 ? 1u + tfull_idx(m2) : 0u;       // output row yi-1 gets its last contribution
            const uint32_t tfb = (yi == g.H - 1 && ta <= 1 && tb >= 1) ? 1u + tfull_idx(m1) : 0u;   // bottom row: row yi too
            const uint32_t wfr = (job == last_job && yi == yb - 1) ? 1u + wfree_idx(wb) : 0u;
            const uint4 q0 = make_uint4(a_lo0 + static_cast<uint32_t>(stage) * (C_::A_STAGE >> 4), b_lo,
                                        tmem_base + (RING - 1u - m_ta) * NT, static_cast<uint32_t>(ta * NT));
            const uint4 q1 = make_uint4(idesc_n((ea - ta + 1) * NT), tmem_base + (RING - 1u - m_eb) * NT,
                                        static_cast<uint32_t>((ea + 1) * NT), has_b ? idesc_n((tb - ea) * NT) : 0u);
            const uint4 q2 = make_uint4(fresh0 ? tmem_base + (RING - 1u - m0) * NT : kNoBlock,
                                        fresh1 ? tmem_base + (RING - 1u - m1) * NT : kNoBlock,
                                        empty_idx(stage) | (tfa << 8) | (tfb << 16) | (wfr << 24), bias_lo);
            const long long c1 = stat_clk(g);
            if (g.stats != nullptr) { ss_nf += ok_f ? 0 : 1; ss_nt += (ok_t0 && ok_t1) ? 0 : 1; }
            {
              uint32_t spins = 0;
              while (!(ok_t0 && ok_t1 && ok_f)) {
                if (!ok_t0) ok_t0 = mbar_try_wait(bar_t0, ph0 ^ 1u);
                if (!ok_t1) ok_t1 = mbar_try_wait(bar_t1, ph1 ^ 1u);
                if (!ok_f) ok_f = mbar_try_wait(bar_f, fph);
                if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
              }
            }
            trace(g, 1, ncmd, 1);
            const long long c2 = stat_clk(g);
            // a free slot of the command ring: the issuer's counter is re-read only when the cached value says "full"
            if (ncmd - cons >= static_cast<uint32_t>(kCmdSlots)) {
              uint32_t spins = 0;
              while (ncmd - (cons = ld_acquire_shared(cons_addr)) >= static_cast<uint32_t>(kCmdSlots))
                if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
            }
            trace(g, 1, ncmd, 2);
            const long long c3 = stat_clk(g);
            const uint32_t dst = cmd0 + (ncmd % kCmdSlots) * static_cast<uint32_t>(sizeof(RowCmd));
            st_shared_v4(dst, q0);
            st_shared_v4(dst + 16, q1);
            st_shared_v4(dst + 32, q2);
            st_release_shared(pub_addr, ncmd + 1u);   // release: the command and the barrier completions observed above
            trace(g, 1, ncmd, 3);
            if (g.stats != nullptr) {
              ss_setup += c1 - c0; ss_wait += c2 - c1; ss_slot += c3 - c2; ss_write += clock64() - c3; ss_rows += 1;
            }
            if (++m0 == RING) { m0 = 0u; ph0 ^= 1u; }
            if (++stage == NSTAGE) { stage = 0; fph ^= 1u; }
          }
          k += static_cast<uint32_t>(j.y1 - j.y0);
        }
      }
      // terminating command
      {
        uint32_t spins = 0;
        while (ncmd - cons >= static_cast<uint32_t>(kCmdSlots)) {
          cons = ld_acquire_shared(cons_addr);
          if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
        }
        st_shared_v4(smem_u32(&cmds[ncmd % kCmdSlots]), make_uint4(kCmdStop, 0u, 0u, 0u));
        st_release_shared(pub_addr, ncmd + 1u);
      }
      stat_add(g, 0, ss_setup); stat_add(g, 4, ss_wait); stat_add(g, 1, ss_slot); stat_add(g, 5, ss_write);
      stat_add(g, 3, ss_rows); stat_add(g, 13, ss_nf); stat_add(g, 14, ss_nt);
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer (one elected lane): executes the command ring =======================
    if (elect_one()) {
      long long si_wait = 0, si_issue = 0, si_commit = 0, si_polls = 0;   // debug counters (registers)
      // high word of both operand descriptors: SBO = 128 B (8 rows x 16 B core matrices), descriptor version 1; the low
      // word (start address + LBO) comes with the command
      constexpr uint64_t kDescHi = (static_cast<uint64_t>((128 >> 4) & 0x3fff) << 32) | (static_cast<uint64_t>(1) << 46);
      const uint64_t ones_desc = umma_smem_desc(smem_u32(sOnes), kLanes * 16, 128);
      const uint32_t pub_addr = smem_u32(const_cast<uint32_t*>(cmd_published)), cons_addr = smem_u32(const_cast<uint32_t*>(cmd_consumed));
      // `pub` rows are known to be published; row n's command is in w0..w2.  The next command is fetched BEFORE this
      // row's MMAs are issued whenever it is already known to be there, so that its loads land while the issue loop is
      // blocked on the tensor queue; `published` is polled again only when the known rows run out.
      uint32_t pub = 0;
      uint4 w0, w1, w2;
      bool have = false;
      for (uint32_t n = 0;; ++n) {
        const long long c0 = stat_clk(g);
        if (!have) {
          uint32_t spins = 0;
          while (pub <= n) {
            pub = ld_acquire_shared(pub_addr);
            si_polls += 1;
            if (++spins > LV_SPIN_LIMIT) asm volatile("trap;");
          }
          const uint32_t ca = smem_u32(&cmds[n % kCmdSlots]);
          w0 = ld_shared_v4(ca); w1 = ld_shared_v4(ca + 16); w2 = ld_shared_v4(ca + 32);
        }
        if (w0.x == kCmdStop) break;
        trace(g, 2, n, 0);
        uint4 n0, n1, n2;
        have = pub > n + 1u;
        if (have) {
          const uint32_t nca = smem_u32(&cmds[(n + 1u) % kCmdSlots]);
          n0 = ld_shared_v4(nca); n1 = ld_shared_v4(nca + 16); n2 = ld_shared_v4(nca + 32);
        }
        if (!kTmaRows) fence_proxy_async_smem();   // cp.async (generic proxy) writes of the row -> UMMA (async proxy) reads;
                                                    // rows written by bulk copies are already in the async proxy
        tc_fence_after_sync();
        const long long c1 = stat_clk(g);
        // RowCmd fields: w0 = {a_lo, b_lo, ra_d, ra_b}, w1 = {ra_i, rb_d, rb_b, rb_i}, w2 = {f0_d, f1_d, bars, bias_lo}
        const uint64_t adesc0 = kDescHi | w0.x;
        const uint64_t bdesc0 = kDescHi | w0.y;
        // LV_ROW_DBG_NOSPLIT (bit 0: no split MMAs at ring wraps, bit 1: no bias MMAs): timing experiments, results garbage
        const bool has_b = (LV_ROW_DBG_NOSPLIT & 1) ? false : (w1.w != 0u);
        // blocks that get their first contribution from this row: D = ones x bias tile (overwrite)
        if (!(LV_ROW_DBG_NOSPLIT & 2)) {
          if (w2.x != kNoBlock) umma_bf16(w2.x, ones_desc, kDescHi | w2.w, idesc_n(NT), 0u);
          if (w2.y != kNoBlock) umma_bf16(w2.y, ones_desc, kDescHi | w2.w, idesc_n(NT), 0u);
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int ks = 0; ks < C_::KSTEPS; ++ks) {
            // descriptor start addresses move in 16-byte units: horizontal tap = one pixel, k-step = two chunk planes
            const uint64_t adesc = adesc0 + static_cast<uint64_t>((kx * 16 + 2 * ks * C_::A_PLANE) >> 4);
            const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((kx * C_::W_KX + 2 * ks * C_::W_PLANE) >> 4);
            umma_bf16(w0.z, adesc, bdesc + w0.w, w1.x, 1u);
            trace(g, 5 + ((kx * C_::KSTEPS + ks) >> 2), n, (kx * C_::KSTEPS + ks) & 3);
          }
        }
        // the second run of a row whose targets straddle the ring's wrap, AFTER the first one: alternating between two
        // accumulator addresses costs the tensor pipe ~110 clk per switch (measured: +2,000 clk per such row)
        if (has_b) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int ks = 0; ks < C_::KSTEPS; ++ks) {
              const uint64_t adesc = adesc0 + static_cast<uint64_t>((kx * 16 + 2 * ks * C_::A_PLANE) >> 4);
              const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((kx * C_::W_KX + 2 * ks * C_::W_PLANE) >> 4);
              umma_bf16(w1.y, adesc, bdesc + w1.z, w1.w, 1u);
            }
          }
        }
        const long long c2 = stat_clk(g);
        trace(g, 2, n, 1);
        const uint32_t bi = w2.z;
        umma_commit(bar0 + 8u * (bi & 0xffu));                                         // row buffer reusable once these MMAs retire
        if ((bi >> 8) & 0xffu) umma_commit(bar0 + 8u * (((bi >> 8) & 0xffu) - 1u));    // output row yi-1 complete
        if ((bi >> 16) & 0xffu) umma_commit(bar0 + 8u * (((bi >> 16) & 0xffu) - 1u));  // image's bottom row: row yi complete
        if (bi >> 24) umma_commit(bar0 + 8u * ((bi >> 24) - 1u));                      // layer's last row: weight buffer free
        st_release_shared(cons_addr, n + 1u);   // the command's words are in registers: its slot may be rewritten
        trace(g, 2, n, 2);
        if (have) { w0 = n0; w1 = n1; w2 = n2; }
        if (g.stats != nullptr) { si_wait += c1 - c0; si_issue += c2 - c1; si_commit += clock64() - c2; }
      }
      stat_add(g, 2, si_issue); stat_add(g, 6, si_commit); stat_add(g, 7, si_wait); stat_add(g, 12, si_polls);
    }
    __syncwarp();
  } else {
    // =============================== epilogue: TMEM -> registers -> global ========================
    EpiCtx cx;
    cx.eg = warp >> 2;
    const int q = warp & 3;
    cx.m = q * 32 + lane;
    cx.lane = lane;
    cx.g = g;
    cx.done = done;
    cx.pub_seen = pub_seen;
    cx.tfull0 = tfull_bar(0);
    cx.tempty0 = tempty_bar(0);
    cx.pub_bar0 = pub_bar(0);
    cx.tmem_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    cx.G = G;
    cx.nlayers = nlayers;
    cx.chunk_stride = static_cast<size_t>(g.W) * 8;
    cx.row_stride = static_cast<size_t>(NT / 8) * g.W * 8;

    pdl_wait();
    uint32_t k = 0;   // CTA-local output-row counter; this group handles the rows with k % 2 == eg
    for (int l = 0; l < nlayers; ++l) {
      const lv_conv_args& a = P.layer[l];
      const bool has_ops = (a.mask != nullptr) || (a.res1 != nullptr) || (a.res2 != nullptr);
      int kind = kKindGeneric;
      if (a.cout == NT && a.res_scale == 1.0f) {
        if (a.epilogue == LV_EPI_NHWC) {
          const int code = (a.relu ? 1 : 0) | (a.mask ? 2 : 0) | (a.res1 ? 4 : 0) | (a.res2 ? 8 : 0);
          if (code == 0 || code == 1 || code == 2 || code == 4 || code == 12) kind = code;
        } else if (NT == 48 && a.epilogue == LV_EPI_PS4_NCHW && !a.relu && !has_ops) {
          kind = kKindPs4;
        }
      }
      const int first = first_job(l);
      if constexpr (NT >= 48) {
        switch (kind) {
          case 0: k = run_layer<0, NT, RING>(cx, a, l, first, k); break;
          case 1: k = run_layer<1, NT, RING>(cx, a, l, first, k); break;
          case 2: k = run_layer<2, NT, RING>(cx, a, l, first, k); break;
          case 4: k = run_layer<4, NT, RING>(cx, a, l, first, k); break;
          case 12: k = run_layer<12, NT, RING>(cx, a, l, first, k); break;
          case kKindPs4:
            if constexpr (NT == 48) { k = run_layer<kKindPs4, NT, RING>(cx, a, l, first, k); break; }
          default: k = run_layer<kKindGeneric, NT, RING>(cx, a, l, first, k); break;
        }
      } else {
        // narrow outputs (EDSR's 64 -> 3 RGB conv, cout padded to 16): the shared 16-channel epilogue routine only
        (void)kind;
        k = run_layer<kKindGeneric, NT, RING>(cx, a, l, first, k);
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    tmem_dealloc<C_::TMEM_COLS>(tmem_base);
  }
  if (nlayers > 1) {
    // self-cleaning workspace: done[total_jobs] is the exit counter; the last CTA out resets everything
    if (threadIdx.x == 0) {
      __threadfence();
      const uint32_t prev = atomicAdd(done + g.total_jobs, 1u);
      *s_last = (prev == static_cast<uint32_t>(G) - 1u) ? 1u : 0u;
      __threadfence();
    }
    __syncthreads();
    if (*s_last != 0u) {
      for (int i = threadIdx.x; i <= g.total_jobs; i += kThreads) done[i] = 0u;
    }
  }
}

// rows per job: minimise (rounds of jobs per CTA) x (rows + ~1.7 halo-row equivalents); short jobs keep every SM busy on
// small problems, long jobs amortise the two halo rows on large ones
static Geom make_geom(int n, int h, int w, int ctas) {
  Geom g;
  g.stats = nullptr;
  g.N = n; g.H = h; g.W = w; g.P = w + 1;
  const long long line = static_cast<long long>(n) * g.P - 1;
  g.nstrips = static_cast<int>((line + kLanes - 1) / kLanes);
  if (g.nstrips < 1) g.nstrips = 1;
  double best = 1e300;
  int best_r = 1;
  for (int r = 1; r <= h; ++r) {
    const long long jobs = static_cast<long long>(g.nstrips) * ((h + r - 1) / r);
    const long long rounds = (jobs + ctas - 1) / ctas;
    const double cost = static_cast<double>(rounds) * (r + 1.7);
    if (cost < best - 1e-9) { best = cost; best_r = r; }
  }
  g.rows_per_job = best_r;
  g.nblocks = (h + best_r - 1) / best_r;
  g.total_jobs = g.nstrips * g.nblocks;
  return g;
}

}  // namespace LV_ROW_NS

// checks the persistent grid can be co-resident (the data-flow waits need every CTA running) and opts into the
// dynamic shared memory, once per device
template <typename K>
static int prepare_kernel(K kern, size_t smem, int threads, int grid, bool needs_coresidency, size_t (&configured)[64]) {
  int dev = 0;
  LV_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (configured[dev] < smem) {
    LV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    LV_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
    if (per_sm < 1) {
      set_error("conv row kernel: a CTA (%zu B shared memory, %d threads) does not fit on an SM of device %d", smem, threads, dev);
      return LV_ERR_UNSUPPORTED;
    }
    configured[dev] = smem;
  }
  if (needs_coresidency && grid > sm_count()) {
    set_error("conv row chain: grid %d exceeds the %d SMs that can hold one resident CTA each", grid, sm_count());
    return LV_ERR_UNSUPPORTED;
  }
  return LV_OK;
}

template <int CIN, int NT, int NSTAGE, int WBUFS>
static int launch_row(const lv_conv_args* layers, int count, void* sync_ws, long long sync_ws_bytes, int max_ctas,
                      cudaStream_t stream) {
  using C_ = LV_ROW_NS::Cfg<CIN, NT, NSTAGE, WBUFS>;
  const lv_conv_args& a0 = layers[0];
  int ctas = max_ctas > 0 ? max_ctas : sm_count();
  if (ctas > sm_count()) ctas = sm_count();
  LV_ROW_NS::Geom g = LV_ROW_NS::make_geom(a0.n, a0.h, a0.w, ctas);
  g.stats = g_timeline;
  if (g.total_jobs < ctas) ctas = g.total_jobs;
  if (count > 1) {
    LV_CHECK_ARG(sync_ws != nullptr && sync_ws_bytes >= (static_cast<long long>(g.total_jobs) + 1) * 4,
                 "conv row chain: sync workspace too small (%lld < %lld bytes)", sync_ws_bytes,
                 (static_cast<long long>(g.total_jobs) + 1) * 4);
  }
  auto kern = LV_ROW_NS::conv3x3_row_kernel<CIN, NT, NSTAGE, WBUFS>;
  static size_t configured[64] = {0};   // per kernel instantiation (this function is one) and per device
  int rc = prepare_kernel(kern, C_::smem_bytes(), LV_ROW_NS::kThreads, ctas, count > 1, configured);
  if (rc != LV_OK) return rc;
  static thread_local LV_ROW_NS::Params params;   // staging only; the launch copies it by value
  for (int i = 0; i < count; ++i) params.layer[i] = layers[i];
  const int rot = g.total_jobs % ctas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas));
  cfg.blockDim = dim3(LV_ROW_NS::kThreads);
  cfg.dynamicSmemBytes = C_::smem_bytes();
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, params, count, g, static_cast<uint32_t*>(sync_ws), rot));
  count_launch();
  return LV_OK;
}

// entry of this build of the kernel (LV_ROW_ENTRY: conv3x3_row_chain_tma / conv3x3_row_chain_cp); arguments are validated
// by the dispatcher conv3x3_row_chain (conv_row.cu)
int LV_ROW_ENTRY(const lv_conv_args* layers, int count, void* sync_ws, long long sync_ws_bytes, int max_ctas,
                 cudaStream_t stream) {
  if (layers[0].cin == 48) return launch_row<48, 48, LV_ROW_STAGES48, LV_ROW_WBUFS48>(layers, count, sync_ws, sync_ws_bytes, max_ctas, stream);
  if (layers[0].cout <= 16) return launch_row<64, 16, 4, 2>(layers, count, sync_ws, sync_ws_bytes, max_ctas, stream);
  return launch_row<64, 64, LV_ROW_STAGES64, 2>(layers, count, sync_ws, sync_ws_bytes, max_ctas, stream);
}

}  // namespace lv
