// extern "C" boundary of liblarvanet_b200.so (declared in include/larvanet_b200.h).
// Plain pointers and sizes only; no torch types.  Every entry validates its arguments, enqueues on the caller's stream
// and returns an LV_* status; failures leave a thread-local message for lv_last_error().
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "lv_common.cuh"

namespace lv {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

// implemented in the other translation units
int conv3x3_tc(const lv_conv_args& a, int max_ctas, cudaStream_t stream);
#ifdef LV_EXPERIMENTAL
int conv3x3_tc_ky(const lv_conv_args& a, int max_ctas, cudaStream_t stream);
#endif
bool conv3x3_row_supported(const lv_conv_args& a);
long long conv3x3_row_workspace_bytes(int n, int h, int w);
int conv3x3_row_chain(const lv_conv_args* layers, int count, void* sync_ws, long long sync_ws_bytes, int max_ctas,
                      cudaStream_t stream);
int conv3x3_simt(const lv_conv_args& a, cudaStream_t stream);
long long conv3x3_chain_workspace_bytes(int n, int h, int w);
int conv3x3_chain(const lv_conv_args* layers, int count, void* sync_ws, long long sync_ws_bytes, int max_ctas,
                  cudaStream_t stream);
int pick_ntile(int cout_pad);
extern long long* g_timeline;
extern int g_use_pdl;
int head_bicubic_fwd(const float*, const float*, const float*, const float*, const float*, void*, float*, int, int, int, int,
                     int, cudaStream_t);
int bicubic_x4(const float*, float*, int, int, int, int, cudaStream_t);
long long head_wgrad_workspace_bytes(int);
int head_wgrad(const float*, const void*, float*, float*, int, int, int, int, int, float, void*, long long, int, cudaStream_t);
int pack_weights(const lv_pack_item*, int, cudaStream_t);
int nchw_to_nhwc(const float*, void*, int, int, int, int, int, cudaStream_t);
int nhwc_to_nchw(const void*, float*, int, int, int, int, int, cudaStream_t);
int l1_loss_grad(const float*, const float*, double*, void*, int, int, int, int, int, cudaStream_t);
int adamw_step(float*, const float*, float*, float*, long long, float, float, float, float, float, int, float, cudaStream_t);
int image_to_uint8(const float*, uint8_t*, long long, cudaStream_t);
int adamw_pack_step(float*, const float*, float*, float*, long long, float, float, float, float, float, int, float,
                    const lv_fused_conv*, int, cudaStream_t);
int psnr_sqsum(const float*, const float*, double*, int, int, int, int, int, cudaStream_t);
int crop_augment(const lv_patch_item*, int, float*, float*, int, int, cudaStream_t);
int dp_adamw_pack_step(float*, float*, float*, long long, float, float, float, float, float, int, float, const lv_fused_conv*, int,
                       const void* const*, void* const*, void* const*, const void* const*, double*, uint32_t*, long long, int, int,
                       cudaStream_t);
long long wgrad_workspace_bytes(const lv_wgrad_item*, int, int);
int wgrad(const lv_wgrad_item*, const lv_wgrad_item*, int, int, void*, cudaStream_t);
int wgrad_simt(const lv_wgrad_item*, const lv_wgrad_item*, int, int, cudaStream_t);

static int check_conv(const lv_conv_args* a) {
  LV_CHECK_ARG(a != nullptr, "conv3x3: null args");
  LV_CHECK_ARG(a->n >= 0 && a->h >= 0 && a->w >= 0, "conv3x3: negative geometry");
  LV_CHECK_ARG(a->dtype == LV_F32 || a->dtype == LV_BF16, "conv3x3: bad dtype %d", a->dtype);
  LV_CHECK_ARG(a->num_src >= 1 && a->num_src <= LV_MAX_SRC, "conv3x3: num_src must be 1..%d", LV_MAX_SRC);
  LV_CHECK_ARG(a->cin > 0 && a->cout > 0, "conv3x3: bad channel counts");
  LV_CHECK_ARG(a->weights != nullptr, "conv3x3: null weights");
  LV_CHECK_ARG(a->wlayout == LV_W_TAP_MAJOR || (a->wlayout == LV_W_KY_STACKED && a->dtype == LV_BF16), "conv3x3: bad wlayout %d", a->wlayout);
  for (int s = 0; s < a->num_src; ++s) LV_CHECK_ARG(a->src[s] != nullptr, "conv3x3: null source %d", s);
  switch (a->epilogue) {
    case LV_EPI_NHWC:
      LV_CHECK_ARG(a->out != nullptr, "conv3x3: EPI_NHWC needs out");
      LV_CHECK_ARG(a->cout % 16 == 0, "conv3x3: EPI_NHWC needs cout %% 16 == 0 (got %d)", a->cout);
      break;
    case LV_EPI_PS4_NCHW:
      LV_CHECK_ARG(a->cout % 16 == 0, "conv3x3: EPI_PS4 needs cout %% 16 == 0");
      LV_CHECK_ARG(a->out_hr != nullptr || a->truth_hr != nullptr || a->out_u8 != nullptr,
                   "conv3x3: EPI_PS4 needs out_hr, out_u8 and/or truth_hr");
      LV_CHECK_ARG(a->truth_hr == nullptr || a->loss_sum != nullptr, "conv3x3: truth_hr given without loss_sum");
      break;
    case LV_EPI_PS2_NHWC:
      LV_CHECK_ARG(a->out != nullptr && a->cout % 16 == 0, "conv3x3: EPI_PS2 needs out and cout %% 16 == 0");
      break;
    case LV_EPI_RGB_NCHW:
      LV_CHECK_ARG(a->out_hr != nullptr && a->cout == 3, "conv3x3: EPI_RGB needs out_hr and cout == 3");
      LV_CHECK_ARG(a->mask == nullptr && a->res1 == nullptr && a->res2 == nullptr, "conv3x3: EPI_RGB takes no mask/residual");
      break;
    default:
      LV_CHECK_ARG(false, "conv3x3: unknown epilogue %d", a->epilogue);
  }
  return LV_OK;
}

}  // namespace lv

using namespace lv;

extern "C" {

const char* lv_last_error(void) { return g_err; }
int lv_abi_version(void) { return LV_ABI_VERSION; }
int64_t lv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
// developer hook (not in the public header): device buffer of 3*64*4 int64 that CTA 0 of the next tensor-core conv
// launches fills with clock64 stamps per role/tile; pass NULL to switch off
void lv_debug_set_timeline(long long* buf) { lv::g_timeline = buf; }

int lv_device_check(int dev, int* sm) {
  cudaDeviceProp prop;
  LV_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (sm != nullptr) *sm = prop.multiProcessorCount;
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; larvanet_b200 is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    return LV_ERR_UNSUPPORTED;
  }
  return LV_OK;
}

int64_t lv_packed_weight_bytes(int cout, int cin_total, int dtype) {
  const int64_t cout_pad = (cout + 15) / 16 * 16;
  return cout_pad * cin_total * 9 * (dtype == LV_BF16 ? 2 : 4);
}

int lv_pack_conv3x3_weights(const lv_pack_item* items, int count, void* stream) {
  LV_CHECK_ARG(items != nullptr || count == 0, "pack: null items");
  return pack_weights(items, count, static_cast<cudaStream_t>(stream));
}

int lv_conv3x3(const lv_conv_args* a, int max_ctas, void* stream) {
  static const bool pdl_env = [] {
    const char* e = getenv("LARVANET_B200_PDL");
    if (e != nullptr) lv::g_use_pdl = (e[0] != '0');
    return true;
  }();
  (void)pdl_env;
  int rc = check_conv(a);
  if (rc != LV_OK) return rc;
  if (a->dtype == LV_BF16) {
    LV_CHECK_ARG(a->cin % 16 == 0, "conv3x3: the tensor-core path needs cin %% 16 == 0 (got %d)", a->cin);
    if (a->wlayout == LV_W_KY_STACKED) {
      // row-marching kernel (conv_row.cu): 9 MMAs of N = 3*cout per 128 pixels, vertical taps summed in TMEM
      if (conv3x3_row_supported(*a)) return conv3x3_row_chain(a, 1, nullptr, 0, max_ctas, static_cast<cudaStream_t>(stream));
#ifdef LV_EXPERIMENTAL
      return conv3x3_tc_ky(*a, max_ctas, static_cast<cudaStream_t>(stream));
#else
      set_error("conv3x3: ky-stacked weights are only supported for single-source bf16 48->48 / 64->64 convs "
                "(EPI_NHWC, or EPI_PS4 at 48 channels); got cin=%d x %d sources, cout=%d, epilogue %d",
                a->cin, a->num_src, a->cout, a->epilogue);
      return LV_ERR_INVALID;
#endif
    }
    return conv3x3_tc(*a, max_ctas, static_cast<cudaStream_t>(stream));
  }
  return conv3x3_simt(*a, static_cast<cudaStream_t>(stream));
}

int64_t lv_conv_chain_workspace_bytes(int n, int h, int w) {
  if (n < 0 || h < 0 || w < 0) return -1;
  const long long a = conv3x3_chain_workspace_bytes(n, h, w), b = conv3x3_row_workspace_bytes(n, h, w);
  return a > b ? a : b;   // one workspace serves both chain kernels (tile flags / row-job flags)
}

int lv_conv3x3_chain(const lv_conv_args* layers, int count, void* sync_ws, int64_t sync_ws_bytes, int max_ctas,
                     void* stream) {
  LV_CHECK_ARG(layers != nullptr && count >= 1, "conv chain: no layers");
  for (int i = 0; i < count; ++i) {
    int rc = check_conv(&layers[i]);
    if (rc != LV_OK) return rc;
    LV_CHECK_ARG(layers[i].bias == nullptr || (reinterpret_cast<uintptr_t>(layers[i].bias) & 15u) == 0,
                 "conv chain: layer %d bias is not 16-byte aligned", i);
  }
  if (layers[0].wlayout == LV_W_KY_STACKED)
    return conv3x3_row_chain(layers, count, sync_ws, sync_ws_bytes, max_ctas, static_cast<cudaStream_t>(stream));
  return conv3x3_chain(layers, count, sync_ws, sync_ws_bytes, max_ctas, static_cast<cudaStream_t>(stream));
}

int lv_conv3x3_simt(const lv_conv_args* a, void* stream) {
  int rc = check_conv(a);
  if (rc != LV_OK) return rc;
  return conv3x3_simt(*a, static_cast<cudaStream_t>(stream));
}

int lv_head_bicubic_fwd(const float* x, const float* w, const float* b, const float* pre_w, const float* pre_b, void* fea,
                        float* base_hr, int n, int h, int w_, int cout, int dtype, void* stream) {
  LV_CHECK_ARG(x && w && fea, "head: null pointer");
  LV_CHECK_ARG((pre_w == nullptr) == (pre_b == nullptr), "head: pre_w and pre_b go together");
  return head_bicubic_fwd(x, w, b, pre_w, pre_b, fea, base_hr, n, h, w_, cout, dtype, static_cast<cudaStream_t>(stream));
}

int lv_bicubic_x4(const float* x, float* base_hr, int n, int c, int h, int w_, void* stream) {
  LV_CHECK_ARG(x && base_hr, "bicubic: null pointer");
  return bicubic_x4(x, base_hr, n, c, h, w_, static_cast<cudaStream_t>(stream));
}

int64_t lv_head_wgrad_workspace_bytes(int cout) { return cout > 0 ? head_wgrad_workspace_bytes(cout) : -1; }

int lv_head_wgrad(const float* x, const void* dy, float* dw, float* db, int n, int h, int w_, int cout, int dtype,
                  float scale, void* workspace, int64_t workspace_bytes, int overwrite, void* stream) {
  LV_CHECK_ARG(x && dy && dw, "head wgrad: null pointer");
  return head_wgrad(x, dy, dw, db, n, h, w_, cout, dtype, scale, workspace, workspace_bytes, overwrite,
                    static_cast<cudaStream_t>(stream));
}

int64_t lv_wgrad_workspace_bytes(const lv_wgrad_item* items_host, int count, int splits) {
  return wgrad_workspace_bytes(items_host, count, splits);
}

int lv_conv3x3_wgrad(const lv_wgrad_item* items_host, const lv_wgrad_item* items_dev, int count, int splits, void* workspace,
                     void* stream) {
  LV_CHECK_ARG(items_host && items_dev, "wgrad: null item arrays");
  return wgrad(items_host, items_dev, count, splits, workspace, static_cast<cudaStream_t>(stream));
}

int lv_conv3x3_wgrad_simt(const lv_wgrad_item* items_host, const lv_wgrad_item* items_dev, int count, int splits,
                          void* stream) {
  LV_CHECK_ARG(items_host && items_dev, "wgrad: null item arrays");
  return wgrad_simt(items_host, items_dev, count, splits, static_cast<cudaStream_t>(stream));
}

int lv_nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w_, int dtype, void* stream) {
  LV_CHECK_ARG(src && dst, "nchw_to_nhwc: null pointer");
  return nchw_to_nhwc(src, dst, n, c, h, w_, dtype, static_cast<cudaStream_t>(stream));
}
int lv_nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w_, int dtype, void* stream) {
  LV_CHECK_ARG(src && dst, "nhwc_to_nchw: null pointer");
  return nhwc_to_nchw(src, dst, n, c, h, w_, dtype, static_cast<cudaStream_t>(stream));
}

int lv_l1_loss_grad(const float* out_hr, const float* truth_hr, double* loss_sum, void* grad_sign, int n, int c, int h,
                    int w_, int dtype, void* stream) {
  LV_CHECK_ARG(out_hr && truth_hr, "l1: null pointer");
  return l1_loss_grad(out_hr, truth_hr, loss_sum, grad_sign, n, c, h, w_, dtype, static_cast<cudaStream_t>(stream));
}

int lv_adamw_pack_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                       const lv_fused_conv* convs, int nconv, void* stream) {
  LV_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adamw+pack: null pointer");
  return adamw_pack_step(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                         convs, nconv, static_cast<cudaStream_t>(stream));
}

int lv_dp_adamw_pack_step(float* param, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int step, float grad_scale, const lv_fused_conv* convs, int nconv,
                          const void* const* peer_grads, void* const* peer_reduced, void* const* peer_flags,
                          const void* const* peer_loss, double* loss_out, uint32_t* ctl, int64_t slice, int world, int rank,
                          void* stream) {
  LV_CHECK_ARG(param && exp_avg && exp_avg_sq, "dp adamw+pack: null pointer");
  return dp_adamw_pack_step(param, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, weight_decay, step, grad_scale, convs,
                            nconv, peer_grads, peer_reduced, peer_flags, peer_loss, loss_out, ctl, slice, world, rank,
                            static_cast<cudaStream_t>(stream));
}

int lv_crop_augment(const lv_patch_item* items_dev, int count, float* out_lr, float* out_hr, int patch, int scale,
                    void* stream) {
  LV_CHECK_ARG(count == 0 || (items_dev && out_lr && out_hr), "crop_augment: null pointer");
  return crop_augment(items_dev, count, out_lr, out_hr, patch, scale, static_cast<cudaStream_t>(stream));
}

int lv_image_to_uint8(const float* src, uint8_t* dst, int64_t numel, void* stream) {
  LV_CHECK_ARG(numel >= 0 && (numel == 0 || (src && dst)), "image_to_uint8: null pointer");
  return image_to_uint8(src, dst, numel, static_cast<cudaStream_t>(stream));
}

int lv_psnr_sqsum(const float* out, const float* truth, double* sq_sum, int c, int h, int w_, int truth_h, int truth_w,
                  void* stream) {
  LV_CHECK_ARG(out && truth && sq_sum, "psnr: null pointer");
  LV_CHECK_ARG(c >= 0 && h >= 0 && w_ >= 0, "psnr: negative size");
  return psnr_sqsum(out, truth, sq_sum, c, h, w_, truth_h, truth_w, static_cast<cudaStream_t>(stream));
}

int lv_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  LV_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adamw: null pointer");
  return adamw_step(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                    static_cast<cudaStream_t>(stream));
}

}  // extern "C"
