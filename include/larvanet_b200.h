/*
 * larvanet_b200 -- C-ABI of the B200-native (sm_100a) LarvaNet / LarvaNetV2 / EDSR-baseline x4 SR hot path.
 *
 * The reference (Geunwoo-Jeon/LarvaNet) is pure Python/PyTorch and has NO FFI of its own:
 * every entry point below replaces one or more `torch.nn` call sites inside the reference's modules; the
 * reference interface each one replaces is cited as file:line relative to the reference root.  Python binds
 * these with ctypes (`larvanet_b200/_lib.py`; stub shown in INTEGRATION.md).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless a parameter says "host"; the library never allocates, frees or
 *     synchronises; work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 = ok, otherwise an LV_ERR_* code; lv_last_error() returns a thread-local message.
 *   - activations are channels-last in 8-channel planes ("planar-8"): element (n, y, x, c) of a C-channel tensor lives
 *     at ((((n*H + y) * (C/8) + c/8) * W + x) * 8 + c%8), i.e. [N][H][C/8][W][8]; C must be a multiple of 8.  Comments
 *     and names below that say "NHWC [n,h,w,c]" mean this layout (logical NHWC indexing, planar-8 storage).
 *     dtype LV_BF16 is the product path (tcgen05 tensor cores, fp32 accumulate);
 *     dtype LV_F32 is the fp32 validation mode (CUDA-core direct convolution, same epilogues).
 *   - images at the Python boundary stay NCHW fp32 on the 0..255 scale like the reference
 *     (models/LarvaNet.py:163-171).
 */
#ifndef LARVANET_B200_H_
#define LARVANET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LV_ABI_VERSION 2

enum { LV_F32 = 0, LV_BF16 = 1 };

enum {
  LV_OK = 0,
  LV_ERR_INVALID = 1,     /* bad argument / unsupported shape */
  LV_ERR_CUDA = 2,        /* a CUDA runtime call failed (message has the cudaError string) */
  LV_ERR_UNSUPPORTED = 3  /* wrong device (needs sm_100) */
};

/* conv epilogue kinds */
enum {
  LV_EPI_NHWC = 0,      /* out[n,h,w,co]                       (ResidualBlock / merge / EDSR body convs)   */
  LV_EPI_PS4_NCHW = 1,  /* out_hr[n,c,4h+i,4w+j] = v[16c+4i+j] + base   (LarvaLeg / LarvaTail, fp32 NCHW)  */
  LV_EPI_PS2_NHWC = 2,  /* out[n,2h+i,2w+j,c]    = v[4c+2i+j]           (EDSR UpsampleBlock, NHWC)          */
  LV_EPI_RGB_NCHW = 3   /* out_hr[n,c,h,w] = post_w[c,:].v[0:3] + post_b[c]   (EDSR final_conv + 1x1)      */
};

/* packed weight layouts (LV_BF16): tap-major [ntile][src][tap][cin/8][cout][8] (27 MMAs of N=cout per 16x8-pixel tile) or
 * ky-stacked [src][kx][cin/8][ky*cout_pad+cout][8] (row-marching kernel, csrc/conv_row.cu: 9 MMAs of N=3*cout per 128 pixels
 * of an image row, the three vertical taps summed in TMEM; single-source 48->48 / 64->64 convs only) */
enum { LV_W_TAP_MAJOR = 0, LV_W_KY_STACKED = 1 };

#define LV_MAX_SRC 4

/*
 * One 3x3 / stride 1 / pad 1 convolution with a fused epilogue:
 *     v = res_scale * (conv(src..) + bias);  if (relu) v = max(v,0);  if (mask) v = mask>0 ? v : 0;
 *     v += res1;  v += res2;   then the epilogue kind decides how v is stored.
 * Replaces nn.Conv2d + nn.ReLU + torch.add + nn.PixelShuffle + `out += base` + nn.L1Loss at
 *   models/LarvaNet.py:209-220 (ResidualBlock), :246-248 (LarvaBody skip), :255-267 (LarvaLeg),
 *   models/LarvaNetV2.py:318-334 (LarvaTail; `src[0..num_src)` replaces torch.cat + merge_conv),
 *   models/edsr.py:139-153,156-173,182-207, and -- with flipped/transposed weights, `mask` and `res*` --
 *   the autograd conv backward-data of the same layers (loss.backward(), models/LarvaNet.py:113).
 */
typedef struct lv_conv_args {
  int32_t n, h, w;           /* conv grid (input == output resolution)                                   */
  int32_t cin;               /* channels per source; tensor-core path needs cin % 16 == 0                 */
  int32_t num_src;           /* 1..LV_MAX_SRC                                                            */
  int32_t cout;              /* real output channels                                                      */
  int32_t dtype;             /* LV_BF16 | LV_F32 : type of src/out/res/mask/grad_sign                     */
  int32_t relu;
  int32_t epilogue;          /* LV_EPI_*                                                                  */
  int32_t wlayout;           /* LV_W_TAP_MAJOR | LV_W_KY_STACKED: which packed-weight layout `weights` uses  */
  float res_scale;           /* 1.0 unless EDSR --edsr_res_weight                                        */
  float reserved1;
  const void* src[LV_MAX_SRC]; /* NHWC [n,h,w,cin]                                                        */
  const void* weights;       /* packed by lv_pack_conv3x3_weights for this dtype                          */
  const float* bias;         /* [cout] or NULL                                                            */
  const void* mask;          /* NHWC [n,h,w,cout] or NULL                                                 */
  const void* res1;          /* NHWC [n,h,w,cout] or NULL                                                 */
  const void* res2;          /* NHWC [n,h,w,cout] or NULL                                                 */
  void* out;                 /* EPI_NHWC: [n,h,w,cout]; EPI_PS2_NHWC: [n,2h,2w,cout/4]; else unused       */
  float* out_hr;             /* EPI_PS4_NCHW: fp32 [n,cout/16,4h,4w] (NULL = skip store); EPI_RGB: [n,3,h,w] */
  const float* base_hr;      /* EPI_PS4_NCHW: fp32 [n,cout/16,4h,4w] added to the shuffled output, or NULL */
  const float* truth_hr;     /* EPI_PS4_NCHW: if non-NULL, fuse nn.L1Loss: loss_sum += sum|out-truth| and   */
  double* loss_sum;          /*   write grad_sign = sign(out-truth) in the pre-shuffle NHWC layout          */
  void* grad_sign;           /* NHWC [n,h,w,cout] of dtype                                                 */
  const float* post_w;       /* EPI_RGB_NCHW: fp32 [3,3] 1x1 conv weight (row-major [out,in]) or NULL      */
  const float* post_b;       /* EPI_RGB_NCHW: fp32 [3] or NULL                                             */
  uint8_t* out_u8;           /* EPI_PS4_NCHW: if non-NULL, also store clip(round_half_even(out), 0, 255) as uint8 */
                             /*   [n,cout/16,4h,4w] -- the PNG-ready frame of get_sr.py:86-89 / validate.py:17-18 --  */
                             /*   straight from the epilogue (out_hr may then be NULL: no fp32 frame is written)  */
} lv_conv_args;

/* Weight-gradient work item (one conv layer, or one source slice of the V2 merge conv). */
typedef struct lv_wgrad_item {
  int32_t n, h, w;
  int32_t cin, cout;         /* cin of THIS slice (tensor-core path: 48 or 64), cout <= 64 on the tc path  */
  int32_t cin_total, cin_off;/* dw is [cout, cin_total, 3, 3]; this item fills ci in [cin_off, cin_off+cin) */
  int32_t dtype;
  const void* x;             /* NHWC [n,h,w,cin]  layer input saved by the forward pass                    */
  const void* dy;            /* NHWC [n,h,w,cout] gradient wrt the conv output (before bias)               */
  float* dw;                 /* fp32 OIHW, ACCUMULATED INTO (dw += scale * sum) unless `overwrite`         */
  float* db;                 /* fp32 [cout], same; NULL to skip                                            */
  float scale;
  int32_t overwrite;         /* bf16 tensor-core path only: 1 = store instead of accumulate (this item is the */
                             /* only writer of its dw slice / db, so the caller need not zero them first)  */
} lv_wgrad_item;

/* Weight re-pack work item: fp32 OIHW master weight [O, I, 3, 3] -> packed operand for lv_conv3x3.
 *   transpose = 0 (forward operand):        packed conv has cout = O and cin_total = i_cnt, reading the input-channel
 *                                            slice [i_off, i_off+i_cnt) of the master weight.
 *   transpose = 1 (backward-data operand):  taps rotated by 180 degrees, in/out swapped: packed conv has cout = i_cnt
 *                                            (the slice of ORIGINAL input channels it produces gradients for) and
 *                                            cin_total = O.
 *   `cin` = channels per source of the packed conv (cin_total must be a multiple of it). */
typedef struct lv_pack_item {
  const float* w;
  void* packed;              /* output, lv_packed_weight_bytes(packed cout, packed cin_total, dtype) bytes          */
  int32_t O, I;
  int32_t transpose;
  int32_t i_off, i_cnt;
  int32_t cin;
  int32_t dtype;
  int32_t wlayout;           /* LV_W_TAP_MAJOR | LV_W_KY_STACKED (bf16 only)                                          */
} lv_pack_item;

const char* lv_last_error(void);
int lv_abi_version(void);
/* 0 if device `dev` can run the library (compute capability 10.x); fills sm_count if non-NULL */
int lv_device_check(int dev, int* sm_count);

/* bytes of the packed weight for a conv with `cout` outputs and `cin_total` inputs */
int64_t lv_packed_weight_bytes(int cout, int cin_total, int dtype);
/* batched pack: `items` is a HOST array (copied by value into the launch); count <= LV_PACK_MAX_ITEMS per call */
#define LV_PACK_MAX_ITEMS 128
int lv_pack_conv3x3_weights(const lv_pack_item* items, int count, void* stream);

/* the conv (see lv_conv_args).  `max_ctas` <= 0 lets the library choose (persistent grid).
 * LV_BF16 -> tcgen05 tensor-core kernel; LV_F32 -> fp32 validation kernel.  There is no CPU fallback. */
int lv_conv3x3(const lv_conv_args* args, int max_ctas, void* stream);
/* TEST-ONLY cross-check: the same conv on CUDA cores for either dtype (bf16: identical operands, fp32 accumulate) */
int lv_conv3x3_simt(const lv_conv_args* args, void* stream);

/*
 * A chain of convolutions in ONE persistent launch: layers[0..count) are executed in order as if by `count` calls of
 * lv_conv3x3, but the CTAs stay resident and the layers are linked by per-tile data-flow flags instead of kernel
 * boundaries (no launch, prologue or pipeline drain per layer).  Replaces the Conv2d sequences of
 * models/LarvaNet.py:205-220 (ResidualBlock), :236-248 (LarvaBody) and :251-267 (LarvaLeg), forward and input-gradient.
 * Layers with LV_W_TAP_MAJOR weights run on the 16x8-tile data-flow kernel (csrc/conv_chain.cu, 48 channels); layers with
 * LV_W_KY_STACKED weights on the row-marching kernel (csrc/conv_row.cu, 48 or 64 channels); one layout per call.
 *   layers:  HOST array (copied by value into the launch), 1 <= count <= LV_CHAIN_MAX_LAYERS; every layer must be a single-source
 *            LV_BF16 48 -> 48 conv with LV_W_TAP_MAJOR weights and 16-byte aligned bias, all on the same (n, h, w).
 *            A layer may read (src / res1 / res2 / mask) anything written by an EARLIER layer of the chain or before
 *            the call, and may overwrite any buffer whose readers are earlier layers.
 *   sync_ws: device workspace of lv_conv_chain_workspace_bytes(n, h, w) bytes; zero it once after allocation, the
 *            kernel leaves it zeroed.  One workspace per stream.
 */
#define LV_CHAIN_MAX_LAYERS 96
int64_t lv_conv_chain_workspace_bytes(int n, int h, int w);
int lv_conv3x3_chain(const lv_conv_args* layers, int count, void* sync_ws, int64_t sync_ws_bytes, int max_ctas,
                     void* stream);

/*
 * LarvaHead conv (3->cout, fp32 math) fused with the bicubic x4 base.
 * Replaces models/LarvaNet.py:227,231-233 (LarvaHead) and :283-285 (F.interpolate bicubic, align_corners=False);
 * for EDSR (`pre_w`/`pre_b` non-NULL) also the 1x1 MeanShift before first_conv (models/edsr.py:197-198).
 *   x: fp32 NCHW [n,3,h,w];  fea: NHWC [n,h,w,cout] of dtype;  base_hr: fp32 NCHW [n,3,4h,4w] or NULL
 */
int lv_head_bicubic_fwd(const float* x, const float* w, const float* b, const float* pre_w, const float* pre_b,
                        void* fea, float* base_hr, int n, int h, int w_, int cout, int dtype, void* stream);
/* bicubic x4 only (LarvaNetModule.base, models/LarvaNet.py:283-285) */
int lv_bicubic_x4(const float* x, float* base_hr, int n, int c, int h, int w_, void* stream);
/* head weight/bias gradient (autograd of :227): dw[cout,3,3,3] = scale*sum_px dy*x, db = scale*sum dy when `overwrite`,
 * else accumulated into dw / db.  Deterministic two-pass reduction (no atomics); `workspace` holds the per-block partial
 * sums: lv_head_wgrad_workspace_bytes(cout) bytes. */
int64_t lv_head_wgrad_workspace_bytes(int cout);
int lv_head_wgrad(const float* x, const void* dy, float* dw, float* db, int n, int h, int w_, int cout,
                  int dtype, float scale, void* workspace, int64_t workspace_bytes, int overwrite, void* stream);

/*
 * Batched weight gradients.  `items_dev` is a DEVICE array of `count` items (same dtype; bf16 needs cin == 48);
 * `items_host` is the same array on the host (used for validation and launch geometry only).  Each item is
 * split over `splits` CTAs along its pixel tiles.
 * `workspace` holds split-K partials: lv_wgrad_workspace_bytes(...) bytes.  Replaces autograd's conv
 * backward-filter + bias reduction for every Conv2d(48|64 -> <=64) of the path.
 */
int64_t lv_wgrad_workspace_bytes(const lv_wgrad_item* items_host, int count, int splits);
int lv_conv3x3_wgrad(const lv_wgrad_item* items_host, const lv_wgrad_item* items_dev, int count, int splits,
                     void* workspace, void* stream);
/* TEST-ONLY cross-check of the tensor-core wgrad on CUDA cores (either dtype; fp32 atomics) */
int lv_conv3x3_wgrad_simt(const lv_wgrad_item* items_host, const lv_wgrad_item* items_dev, int count, int splits,
                          void* stream);

/* layout / dtype helpers at the module boundary (NCHW fp32 <-> NHWC dtype) */
int lv_nchw_to_nhwc(const float* src, void* dst, int n, int c, int h, int w_, int dtype, void* stream);
int lv_nhwc_to_nchw(const void* src, float* dst, int n, int c, int h, int w_, int dtype, void* stream);

/*
 * uint8 image helpers on fp32 0..255 images (any layout, element-wise):
 *   lv_image_to_uint8: dst = clip(round_half_even(src), 0, 255)       (validate._image_to_uint8, validate.py:17-18;
 *                      get_sr.py:86-89 before the PNG is written) -- 4x less device->host traffic than the fp32 image.
 *   lv_psnr_sqsum:     *sq_sum += sum over [c, h, w] of (u8(truth) - u8(out))^2, truth [c, truth_h, truth_w] cropped to
 *                      the output's [c, h, w] (validate._fit_truth_image_size + _image_psnr, validate.py:20-27);
 *                      PSNR = 10*log10(255^2 * c*h*w / sq_sum).  Lets LarvaNet.validate_for_train (models/LarvaNet.py:
 *                      141-161) score on the device and read back 8 bytes per image.
 */
int lv_image_to_uint8(const float* src, uint8_t* dst, int64_t numel, void* stream);
int lv_psnr_sqsum(const float* out, const float* truth, double* sq_sum, int c, int h, int w_, int truth_h, int truth_w,
                  void* stream);

/*
 * Stand-alone L1 loss + gradient on fp32 NCHW HR images (nn.L1Loss, models/LarvaNet.py:85,108):
 *   loss_sum += sum|out-truth|;  if grad_sign: sign(out-truth) un-shuffled (PixelShuffle(4) backward) to
 *   NHWC [n,h,w,16*c] of dtype.
 */
int lv_l1_loss_grad(const float* out_hr, const float* truth_hr, double* loss_sum, void* grad_sign,
                    int n, int c, int h, int w_, int dtype, void* stream);

/*
 * Fused multi-tensor AdamW over a flat fp32 parameter/gradient arena (torch.optim.AdamW defaults,
 * models/LarvaNet.py:86-88,114).  `grad_scale` multiplies the gradient first (DP mean / loss scaling).
 */
/* One packed conv weight for lv_adamw_pack_step: OIHW fp32 [48, cin_total, 3, 3] at element offset `w_off` of the parameter
 * arena, its LV_BF16 / LV_W_TAP_MAJOR forward operand and the backward-data operand of every 48-channel source slice. */
typedef struct lv_fused_conv {
  int64_t w_off;
  int32_t cout, cin_total;
  void* fwd;
  void* bwd[LV_MAX_SRC];
} lv_fused_conv;

/* optim.AdamW.step() (models/LarvaNet.py:86-88,114; models/LarvaNetV2.py:83-85,123) for the whole parameter arena AND the
 * re-pack of the listed convs' operands (= lv_adamw_step + lv_pack_conv3x3_weights, forward and backward-data forms) in
 * ONE kernel; `convs` is a HOST array (<= 64, sorted by w_off, disjoint); parameters outside the listed weights get the
 * plain update.  Bit-identical to the two separate calls. */
int lv_adamw_pack_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                       const lv_fused_conv* convs, int nconv, void* stream);

int lv_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                  void* stream);

/*
 * Data-parallel optimizer step (new capability; the reference is single-process): the SUM all-reduce of the
 * gradient arena, optim.AdamW.step() (models/LarvaNet.py:86-88,114) and the operand re-pack of lv_adamw_pack_step in ONE
 * kernel.  Every rank's gradient arena, flag block and loss accumulator live in peer-mapped (symmetric) memory:
 *   peer_grads[r]   fp32 [numel]     gradient arena of rank r as mapped on THIS device (index = rank, own entry included)
 *   peer_reduced[r] fp32 [world*slice] rank r's buffer of reduced slices, or NULL array: one-shot mode (every rank reads all
 *                                    arenas; fine for 2 ranks).  Two-shot mode (`slice` = elements per rank, multiple of 4,
 *                                    world*slice >= numel): rank r reduces slice r into its buffer, a device-side barrier
 *                                    follows, the update reads each element from its owner -- 2*(world-1)/world arena
 *                                    reads per rank instead of world-1
 *   peer_flags[r]   uint32 [48]      zero-initialised flag block of rank r (device-side barriers, slot per writer rank)
 *   peer_loss[r]    double [1]       rank r's local loss accumulator; *loss_out (local) receives the sum over ranks
 *   ctl             uint32 [8]       zero-initialised device-LOCAL control words of this rank
 * The kernel waits (device side, over NVLink) until every rank has reached it -- i.e. all gradients are complete --, sums
 * the gradients of all ranks in rank order (bit-identical sums on every rank), and returns only when every rank is done
 * reading, so the caller may overwrite its arena right after.  Every rank must call it the same number of times.
 * world must be 2, 4 or 8; `host` arrays are read during the call only.
 */
int lv_dp_adamw_pack_step(float* param, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr, float beta1, float beta2,
                          float eps, float weight_decay, int step, float grad_scale, const lv_fused_conv* convs, int nconv,
                          const void* const* peer_grads, void* const* peer_reduced, void* const* peer_flags,
                          const void* const* peer_loss, double* loss_out, uint32_t* ctl, int64_t slice, int world, int rank,
                          void* stream);

/*
 * Device-side training-patch pipeline: random crop + rot90 + horizontal flip of HBM-resident LR / HR image pairs, written
 * straight into the step's batch tensors -- replaces the host-side numpy pipeline of dataloaders/div2k_train_loader.py:72-98
 * (and the torch-op variant dataloaders/div2k_train_loader_tensor.py:57-97) plus the per-step host->device copy of
 * train_larva.py:123-124.  `items_dev`: DEVICE array of `count` items (one per patch; the random draws are the caller's):
 *   out_lr[b] = flip?( rot90( lr[:, y:y+patch, x:x+patch], rot ) ),  out_hr[b] likewise at `scale` times the coordinates
 * (rot = number of counter-clockwise quarter turns as in torch.rot90(k, dims=(1,2)); flip reverses the last axis).
 */
typedef struct lv_patch_item {
  const float* lr;           /* fp32 [3, h, w], 0..255                       */
  const float* hr;           /* fp32 [3, scale*h, scale*w]                   */
  int32_t h, w;              /* LR image size                                */
  int32_t y, x;              /* LR crop origin (0 <= y <= h-patch, same for x) */
  int32_t rot, flip;
} lv_patch_item;
int lv_crop_augment(const lv_patch_item* items_dev, int count, float* out_lr, float* out_hr, int patch, int scale,
                    void* stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t lv_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LARVANET_B200_H_ */
