// Row-marching 3x3 convolution (see conv_row_impl.cuh): this translation unit holds the build with the TMA row producer
// (12 warps, 168 registers; images at least 129 pixels wide) and the dispatcher; conv_row_cp.cu holds the cp.async build
// (13 warps, 128 registers) that serves narrower images, where a 130-pixel window of the line spans several images and
// would need more bulk copies per row than one thread can issue in a row's time.
#define LV_ROW_NS row
#define LV_ROW_ENTRY conv3x3_row_chain_tma
#include "conv_row_impl.cuh"

namespace lv {

int conv3x3_row_chain_cp(const lv_conv_args* layers, int count, void* sync_ws, long long sync_ws_bytes, int max_ctas,
                         cudaStream_t stream);

long long conv3x3_row_workspace_bytes(int n, int h, int w) {
  const long long line = static_cast<long long>(n) * (w + 1) - 1;
  long long strips = (line + row::kLanes - 1) / row::kLanes;
  if (strips < 1) strips = 1;
  return (strips * (h > 0 ? h : 1) + 1) * 4;   // one flag per job at worst (1 row per job) + the exit counter
}

bool conv3x3_row_supported(const lv_conv_args& a) {
  if (a.dtype != LV_BF16 || a.wlayout != LV_W_KY_STACKED || a.num_src != 1) return false;
  // 64 -> (<= 16) with the RGB epilogue: EDSR's last conv (N = 3 * 16 per MMA, 12 MMAs per 128 pixels instead of 36)
  if (a.cin == 64 && a.cout <= 16) return a.epilogue == LV_EPI_RGB_NCHW;
  return a.cin == a.cout && (a.cin == 48 || a.cin == 64) &&
         (a.epilogue == LV_EPI_NHWC || (a.epilogue == LV_EPI_PS4_NCHW && a.cout == 48));
}

// layers[0..count): single-source bf16 C -> C convs (C = 48 or 64, the same for all; or 64 -> <= 16 with the RGB epilogue)
// with ky-stacked weights on one common (n, h, w); count == 1 needs no workspace.
int conv3x3_row_chain(const lv_conv_args* layers, int count, void* sync_ws, long long sync_ws_bytes, int max_ctas,
                      cudaStream_t stream) {
  LV_CHECK_ARG(count >= 1 && count <= row::kMaxLayers, "conv row chain: 1..%d layers per call (got %d)", row::kMaxLayers, count);
  const lv_conv_args& a0 = layers[0];
  for (int i = 0; i < count; ++i) {
    const lv_conv_args& a = layers[i];
    LV_CHECK_ARG(conv3x3_row_supported(a) && a.cin == a0.cin && (a.cout <= 16) == (a0.cout <= 16),
                 "conv row kernel: layer %d is not a single-source bf16 %d->%d conv with ky-stacked weights", i, a0.cin, a0.cout);
    LV_CHECK_ARG(a.n == a0.n && a.h == a0.h && a.w == a0.w, "conv row chain: layer %d has a different geometry", i);
  }
  if (static_cast<long long>(a0.n) * a0.h * a0.w == 0) return LV_OK;
  LV_CHECK_ARG(static_cast<long long>(a0.n) * (a0.w + 1) < (1ll << 30), "conv row kernel: batch x width too large");
  // a 130-pixel window holds at most three runs (image tail | pad | image head) only when the pitch W+1 is >= 130
  if (a0.w >= 129) return conv3x3_row_chain_tma(layers, count, sync_ws, sync_ws_bytes, max_ctas, stream);
  return conv3x3_row_chain_cp(layers, count, sync_ws, sync_ws_bytes, max_ctas, stream);
}

}  // namespace lv
