/* libb200cls.so - C ABI of the B200-native classification training step.
 *
 * The reference (KKKSQJ/DeepLearning) has no operator registry on this path: every op is a stock PyTorch call
 * (nn.Conv2d / nn.BatchNorm2d / nn.Linear / ... -> ATen -> cuDNN / cuBLAS, or oneDNN on CPU), and the only FFI it owns is
 * the pybind module `swin_window_process` (classification/swin_transformer/kernels/window_process/swin_window_process.cpp:70-131).
 * Each entry point below names the reference call site whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C, no C++ types, no torch types; all pointers are DEVICE pointers unless stated otherwise.
 *   - activations are NHWC bf16 (channel count a multiple of 8); parameters/gradients/statistics are fp32.
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates device memory.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no device synchronisation inside, so every call
 *     is CUDA-Graph capturable.
 *   - return 0 on success, negative on failure (B200_EINVAL / B200_EUNSUPPORTED / B200_ECUDA); b200_last_error() returns
 *     a thread-local message. Launch errors are detected with cudaPeekAtLastError only.
 */
#ifndef B200CLS_H_
#define B200CLS_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_EINVAL (-1)
#define B200_EUNSUPPORTED (-2)
#define B200_ECUDA (-3)

#define B200_ACT_NONE 0
#define B200_ACT_RELU 1
#define B200_ACT_GELU 2
#define B200_ACT_GELU_GRAD 3 /* multiply the accumulator by aux_in, which holds GELU'(pre-activation) as written through aux_out
                                by the forward GEMM with B200_ACT_GELU - backward of B200_ACT_GELU */

const char* b200_last_error(void);
int b200_abi_version(void);
/* Number of SMs of the current device (used by callers to size workspaces). */
int b200_sm_count(void);
/* Kernels launched by this library so far in this process (bench.py reports the per-step delta as gpu_launches). */
unsigned long long b200_launch_count(void);

/* ---- convolution / linear as implicit GEMM on tcgen05 (ksize in {1,3} pad ksize/2, stride in {1,2}; or 2x2/s2 unpadded) ----
 * forward:  y[B,Ho,Wo,Cout] = conv(x[B,H,W,Cin], w) (+bias) (act) (+residual)
 *   w        bf16 [Cout][ksize*ksize*Cin]   (b200_pack_weight mode 0)
 *   stats    optional fp32 [b200_conv2d_fwd_stats_rows()][2][Cout]: partial sum / sum of squares of y (as stored), one row
 *            per (persistent CTA group, 32-lane TMEM quadrant) - a few hundred rows; feed them to b200_bn_finalize
 *   residual optional bf16, same shape as y, added after bias/act
 *   out_f32  optional fp32 [B*Ho*Wo][ld_out] - when given the result is written there instead of y (ksize 1 only)
 * replaces nn.Conv2d.forward / nn.Linear.forward: classification/resnet/models/networks.py:107,111,115,119,218;
 * classification/vision_transformer/vit_model.py:66,95,109,129,132. A linear layer is the case H=W=1, B=rows. */
int b200_conv2d_fwd(const void* x, const void* w, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int stride,
                    float* stats, const float* bias, int act, const void* residual, float* out_f32, long long ld_out,
                    void* stream);
int b200_conv2d_fwd_stats_rows(int B, int H, int W, int Cout, int ksize, int stride);
/* one-shot (this thread, next b200_conv2d_fwd call): fold BatchNorm with FIXED statistics into the epilogue,
 * y = act(conv(x) * scale[c] + shift[c] (+ residual)) with act (0 / B200_ACT_RELU) applied after the residual add - the
 * eval-mode forward of conv -> bn -> (+identity) -> relu (classification/resnet/models/networks.py:104-124, utils.py:61-83
 * `evaluate`) without any BatchNorm pass.  scale / shift from b200_bn_eval_coeffs; Cout % 64 == 0. */
int b200_conv2d_fwd_set_bn(const float* scale, const float* shift);
/* one-shot (this thread, next b200_conv2d_dgrad with stride 1 or b200_gemm_dual): the output is the gradient of
 * relu(bn(x_raw)) - the kernel zeroes it where x_raw * scale + shift <= 0 (dz) and writes the per-CTA partial rows
 * stats[rows][2][C] = sum(dz), sum(dz * x_raw), rows = b200_conv2d_fwd_stats_rows(B, H, W, C, ksize, 1) of the dx geometry:
 * the reduce half of F.batch_norm's backward (classification/resnet/models/networks.py:108,112 bn1 / bn2 under loss.backward(),
 * utils.py:33) without a pass over the gradient.  Feed the rows to b200_bn_bwd_finalize, then b200_bn_bwd_apply(src_is_dz=1). */
int b200_dgrad_set_bn_mask(const void* x_raw, const float* scale, const float* shift, float* stats);
/* same convolution writing an fp32 NHWC output (+bias) through TMA - ConvNeXt downsample conv feeding the fp32 stream */
int b200_conv2d_fwd_f32(const void* x, const void* w, float* y, int B, int H, int W, int Cin, int Cout, int ksize,
                        int stride, const float* bias, void* stream);

/* data gradient: dx[B,H,W,Cin] = conv_transpose(dy[B,Ho,Wo,Cout], w) (+residual, same shape as dx; may alias dx)
 *   wd  bf16 [Cin][ksize*ksize*Cout]  (b200_pack_weight mode 1)
 *   ksize 1 & stride 2 writes only the even (h,w) pixels of dx; the others keep their previous contents.
 * replaces the cuDNN backward-data / cuBLAS dgrad autograd runs inside loss.backward() (classification/resnet/utils.py:43). */
int b200_conv2d_dgrad(const void* dy, const void* wd, void* dx, int B, int H, int W, int Cin, int Cout, int ksize,
                      int stride, const void* residual, void* stream);

/* weight gradient: dw[Cout][Cin][ksize][ksize] (fp32, OIHW) (+)= sum_pixels dy (x) x
 *   workspace: b200_conv2d_wgrad_workspace_bytes() bytes of scratch for the split-K partial tiles.
 * replaces the cuDNN backward-filter / cuBLAS wgrad inside loss.backward(). */
int b200_conv2d_wgrad(const void* dy, const void* x, float* dw, void* workspace, size_t workspace_bytes, int B, int H,
                      int W, int Cin, int Cout, int ksize, int stride, int accumulate, void* stream);
size_t b200_conv2d_wgrad_workspace_bytes(int B, int H, int W, int Cin, int Cout, int ksize, int stride);
/* one-shot (this thread, next b200_conv2d_wgrad call): multiply gradient row `cout` by rowscale[cout] (layer scale) */
int b200_conv2d_wgrad_set_rowscale(const float* rowscale);
/* one-shot (this thread, next b200_conv2d_wgrad call): also produce the BIAS gradient = column sums of dy, as per-split
 * partial sums bias_partial fp32 [b200_conv2d_wgrad_splits()][2][Cout] (plane 0; fold with b200_bn_bwd_finalize).  The dy tiles
 * are summed from shared memory by four extra warps of the wgrad kernel: no separate pass over dy
 * (replaces the bias part of loss.backward() for nn.Linear / nn.Conv2d(bias=True): vit_model.py:95,109,127-133). */
int b200_conv2d_wgrad_set_bias_partial(float* bias_partial);
/* one-shot, together with set_bias_partial: the split reduction that follows the wgrad GEMM also folds bias_partial into the
 * finished bias gradient bias_out[Cout] (nn.Linear / nn.Conv2d bias under loss.backward()) - no launch of its own */
int b200_conv2d_wgrad_set_bias_out(float* bias_out);
int b200_conv2d_wgrad_splits(int B, int H, int W, int Cin, int Cout, int ksize, int stride);

/* ---- general GEMM with strided pixel views (transformer layers, patch embedding) ----------------------------------------
 * out[pixel, n] = epilogue( sum_k a[pixel, k] * w[n, k] ), pixels = dim[0] x dim[1] x dim[2] (w fastest), channel stride 1.
 * `a` and `out` must have identical pixel extents; strides are in elements and let `out`/`residual` be token-offset or
 * batch-broadcast views (e.g. ViT: rows 1.. of [B,197,D], residual = pos_embed with batch stride 0).
 * replaces nn.Linear / PatchEmbed conv + the surrounding bias / GELU / residual-add elementwise ops of
 * classification/vision_transformer/vit_model.py:66,95,109,127-133,159-160. */
typedef struct {
  const void* base;     /* first element of the view (bf16 unless stated otherwise) */
  long long dim[3];     /* pixel extents (w, h, n); use 1 for unused dims */
  long long stride[3];  /* element strides of the pixel dims */
} b200_view_t;

typedef struct {
  const void* w;               /* bf16 [N][K] (b200_pack_weight mode 0) */
  int N, K;
  const float* bias;           /* [N] or NULL */
  const float* colscale;       /* [N] multiplier applied after bias/act, before the residual (ConvNeXt layer scale), or NULL */
  int act;                     /* B200_ACT_* */
  int out_f32;                 /* 1: `out` is an fp32 tensor (residual stream) */
  const b200_view_t* residual; /* added after bias/act, or NULL */
  int residual_f32;
  const b200_view_t* aux_out;  /* second bf16 output kept for the backward pass, or NULL: with act == GELU it receives the derivative
                                  GELU'(pre-activation) (evaluated together with the value), otherwise the pre-activation itself */
  const b200_view_t* aux_in;   /* act == B200_ACT_GELU_GRAD: the GELU'(pre-activation) tensor a forward call wrote through aux_out */
  float* stats;                /* optional per-32-row-slab column sum / sum of squares, or NULL */
  const float* rowscale;       /* stochastic depth: per-SAMPLE multiplier [n_samples] applied after bias/act/colscale and
                                  before the residual (drop_path: convNext/models/networks.py:11-26, vit_model.py:12-40,
                                  swin_transformer.py:282,285), or NULL */
  int rows_per_sample;         /* output pixels per sample (sample = flat pixel index / rows_per_sample) */
} b200_gemm_args_t;

int b200_gemm_ex(const b200_view_t* a, const b200_view_t* out, const b200_gemm_args_t* args, void* stream);

/* ---- LayerNorm over the last dim, one warp per row (vit_model.py:194 eps 1e-6; swin_transformer.py:509 eps 1e-5) ------------
 * x is fp32 (x_f32) or bf16, y bf16; mean/rstd [rows] are kept for the backward pass.
 * backward: dx = rstd*(dy*gamma - mean(dy*gamma) - xhat*mean(dy*gamma*xhat)) (+ add), dx/add fp32 or bf16;
 * partial[b200_layernorm_bwd_blocks()][2][C] = per-block (sum dy, sum dy*xhat), folded by b200_bn_bwd_finalize. */
int b200_layernorm_fwd(const void* x, int x_f32, const float* gamma, const float* beta, void* y, int y_f32, float* mean,
                       float* rstd, long long rows, int C, float eps, void* stream);
int b200_layernorm_bwd_blocks(long long rows, int C);
int b200_layernorm_bwd(const void* dy, const void* x, int x_f32, const float* mean, const float* rstd, const float* gamma,
                       const void* add, void* dx, int dx_f32, float* partial, long long rows, int C, void* stream);

/* ---- ViT patch embedding helpers (vit_model.py:56-66,244-250) ---------------------------------------------------------------
 * patchify: NCHW fp32 -> bf16 [B*(H/ps)*(W/ps)][Cin*ps*ps] with k = c*ps*ps + kh*ps + kw (= conv weight.view(D,-1) order) */
int b200_patchify_nchw(const float* x, void* a, int B, int Cin, int H, int W, int ps, void* stream);
int b200_cls_row(const float* cls, const float* pos, float* tokens, int B, int T, int D, void* stream);
int b200_batch_rowsum(const void* g, int g_f32, long long stride_b, int B, int D, float* out, int accumulate,
                      void* stream);
/* strided 2-D copy (16-byte granularity), e.g. gathering the class-token rows of a [B,T,D] tensor */
int b200_copy_rows(const void* src, long long src_pitch_bytes, void* dst, long long dst_pitch_bytes, long long rows,
                   long long row_bytes, void* stream);
/* bias gradients of tall matrices: partial[b200_colsum_partial_slices(rows)][2][cols], folded by b200_bn_bwd_finalize */
int b200_colsum_partial_slices(long long rows);
int b200_colsum_partial(const void* m, long long rows, long long ld, int cols, float* partial, void* stream);

/* ---- multi-head self-attention, head_dim 64, T <= 256 tokens, on tcgen05 (vit_model.py:95-108) ---------------------------------
 * qkv bf16 [B][T][3][H][64] (the qkv Linear output as is), out bf16 [B][T][H*64], lse fp32 [B][H][T].
 * backward: dqkv bf16 [B][T][3][H][64]; delta fp32 [B][H][T] is scratch. Scores / probabilities never touch HBM. */
int b200_attention_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, float scale, void* stream);
int b200_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                       int B, int T, int H, float scale, void* stream);

/* ---- Swin (classification/swin_transformer/models/swin_transformer.py) -----------------------------------------------------
 * Shifted-window attention, 7x7 windows, head_dim 32, on tcgen05. qkv bf16 [B][H][W][3*nH*32] in natural (un-rolled) pixel
 * order; torch.roll / window_partition / window_reverse (:251-280) are folded into the gather / scatter addressing.
 * bias_tab = b200_window_bias_gather(): fp32 [nH][masked ? nW : 1][49 (query i)][64 (key j, 49 used)] holding
 *   log2(e) x ( relative_position_bias_table[relative_position_index[i][j]][h] (:131-134) plus, for shifted blocks, the
 *   attn_mask buffer value mask[w][i][j] (0 / -100, :215-238, :142-147) ) - one small launch per block and step; a soft-max thread
 *   (one query row) fetches its 256-byte row with 13 vector loads. masked = 1 when a mask was folded in.
 * Two windows are processed per tensor-core step (block-diagonal 128x128 score tile). lse fp32 [B][nW][nH][49] is the
 * base-2 log-sum-exp of the (scaled, biased) score rows - an opaque hand-over from forward to backward.
 * backward: dqkv same layout as qkv; dbias dense [nH][49 i][49 j] must be zeroed by the caller (atomics), then
 * b200_window_bias_scatter adds it into the table gradient. */
int b200_window_attention_fwd(const void* qkv, void* out, const float* bias_tab, int masked, float* lse, int B, int H,
                              int W, int nH, int shift, float scale, void* stream);
int b200_window_attention_bwd(const void* qkv, const void* out, const void* dout, const float* bias_tab, int masked,
                              const float* lse, void* dqkv, float* dbias, int B, int H, int W, int nH, int shift,
                              float scale, void* stream);
int b200_window_bias_gather(const float* table, const long long* index, const float* mask, int nW, float* bias_tab, int nH,
                            void* stream);
int b200_window_bias_scatter(const float* dbias, const long long* index, float* dtable, int nH, void* stream);
/* The reference's own operator FFI (kernels/window_process/swin_window_process.cpp:70-131), any 2/4-byte element type:
 *   partition: out[B*nW][ws][ws][C] = window_partition(roll(in[B][H][W][C], shifts=(shift, shift)))   (forward)
 *   merge:     out[B][H][W][C] = roll(window_reverse(in[B*nW][ws][ws][C]), shifts=(shift, shift))
 * roll_and_window_partition_backward(g, s) == merge(g, -s), window_merge_and_roll_backward(g, s) == partition(g, -s). */
int b200_window_partition(const void* in, void* out, int B, int H, int W, int C, int shift, int ws, int elem_bytes,
                          void* stream);
int b200_window_merge(const void* in, void* out, int B, int H, int W, int C, int shift, int ws, int elem_bytes,
                      void* stream);
/* PatchMerging front half (:333-343): 2x2 gather-concat of the fp32 stream + LayerNorm(4C) -> bf16 [B*H/2*W/2][4C] */
int b200_patch_merge_ln_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                            int B, int H, int W, int C, float eps, void* stream);
int b200_patch_merge_ln_bwd_blocks(long long rows);
int b200_patch_merge_ln_bwd(const void* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                            void* dx, float* partial, int B, int H, int W, int C, void* stream);

/* ---- ConvNeXt (classification/convNext/models/networks.py:92-105,160-165) ---------------------------------------------------
 * 7x7 depthwise conv, pad 3, NHWC: out = bias + sum_taps wt[tap][c]*in[...]; wt = tap-major [49][C] copy (b200_dwconv7_pack).
 * flip != 0 correlates with the flipped kernel (data gradient); `add` (same type as out) is summed into the result. */
int b200_dwconv7_pack(const float* w, float* wt, int C, void* stream);
int b200_dwconv7(const void* in, int in_f32, const float* wt, const float* bias, const void* add, void* out, int out_f32,
                 int flip, int B, int H, int W, int C, void* stream);
/* weight gradient dw [C][49] (+)= sum du * x_shifted; workspace = b200_dwconv7_wgrad_workspace_bytes() */
size_t b200_dwconv7_wgrad_workspace_bytes(int B, int H, int W, int C);
int b200_dwconv7_wgrad(const void* du, const float* x, float* dw, void* workspace, size_t workspace_bytes, int B, int H,
                       int W, int C, int accumulate, void* stream);
/* global average pool of [B][HW][C] (fp32 or bf16) -> fp32 [B][C] */
int b200_avgpool_any(const void* x, int x_f32, float* y, int B, int HW, int C, void* stream);
/* partial[b200_colsum_partial_slices(rows)][2][cols] of column sums of a*b (b optional): layer-scale / bias gradients */
int b200_colsum_prod_partial(const void* a, const void* b, long long rows, long long ld, int cols, float* partial,
                             void* stream);
/* layer-scale gradients from the unscaled pwconv2 weight gradient G [C][K]: dgamma, dW2 = gamma*G, db2 = gamma*gsum */
int b200_layerscale_grads(const float* G, const float* W2, const float* b2, const float* gsum, const float* gamma,
                          float* dW2, float* db2, float* dgamma, int C, int K, void* stream);
/* fused AdamW over flat fp32 arenas; wd = per-element weight decay (0 for the no-decay group, convNext/utils.py:144-166);
 * hyper = device {lr, 1-beta1^t, 1-beta2^t, beta1^t, beta2^t}; b200_adamw_tick advances t by one on the device
 * (initialise hyper to {lr, 0, 0, 1, 1}), so the whole update is CUDA-graph replayable. */
int b200_adamw_tick(float* hyper, float beta1, float beta2, void* stream);
int b200_adamw(float* p, const float* g, float* m, float* v, const float* wd, long long n, const float* hyper,
               float beta1, float beta2, float eps, float gscale, const float* clip_coef, void* stream);
/* Global-norm gradient clipping of the Swin recipe (torch.nn.utils.clip_grad_norm_: swin_transformer/utils/torch_utils.py:303-317,
 * main.py:197) without rewriting the gradients: clip[0] = min(1, max_norm / (gscale*||g||_2 + 1e-6)), clip[1] = the norm.
 * The optimizer entries multiply their gradient scale by clip_coef[0] when the pointer is non-null.
 * partial: fp32 scratch of b200_grad_clip_blocks() floats; g must be 16-byte aligned. */
int b200_grad_clip_blocks(void);
int b200_grad_clip_coef(const float* g, long long n, float gscale, float max_norm, float* partial, float* clip, void* stream);

/* ---- BatchNorm2d (train: batch statistics, eval: running statistics) ----------------------------------------------------
 * replaces nn.BatchNorm2d + nn.ReLU (+ residual add) of Bottleneck.forward, classification/resnet/models/networks.py:108-124 */
/* partial[T][2][C] column reductions run on a 2-D grid; `scratch` (b200_reduce_scratch_bytes(T, C) bytes, its first 1024
 * bytes zero on first use - the kernel leaves them zero) carries the slice sums and a ticket counter. One per stream. */
size_t b200_reduce_scratch_bytes(int T, int C);
int b200_bn_finalize(const float* partial, int T, int C, double count, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                     float* mean, float* invstd, float* scale, float* shift, void* scratch, size_t scratch_bytes,
                     void* stream);
int b200_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* scale, float* shift, void* stream);
/* y = act(x*scale[c]+shift[c] (+residual)); x,y,residual bf16 [rows][C] */
int b200_bn_apply(const void* x, const void* residual, void* y, const float* scale, const float* shift, long long rows,
                  int C, int relu, void* stream);
/* backward pass 1: partial[b200_bn_bwd_blocks()][2][C] = per-block sums of dz and dz*x (raw x), dz = g*relu_mask;
 * b200_bn_bwd_finalize(mean, invstd) turns the second into sum(dz*xhat).
 *   y_out (optional) = saved post-activation output used for the mask; otherwise the mask is recomputed from x.
 *   dz_out (optional) receives dz as bf16. */
int b200_bn_bwd_reduce(const void* g, const void* x, const void* y_out, void* dz_out, const float* scale,
                       const float* shift, int relu, long long rows, int C, float* partial, void* stream);
int b200_bn_bwd_blocks(long long rows, int C);
int b200_bn_bwd_finalize(const float* partial, int T, int C, double count, float* dgamma, float* dbeta, int accumulate,
                         float* m1, float* m2, const float* mean, const float* invstd, void* scratch,
                         size_t scratch_bytes, void* stream);
/* backward pass 2: dx = scale*(dz - m1 - xhat*m2); g_is_dz != 0 means `g` already holds dz (mask applied). */
int b200_bn_bwd_apply(const void* g, const void* x, const void* y_out, int g_is_dz, void* dx, const float* scale,
                      const float* shift, const float* mean, const float* invstd, const float* m1, const float* m2,
                      int relu, long long rows, int C, void* stream);
/* ---- pooling ----------------------------------------------------------------------------------------------------------
 * stem: y[B,Ho,Wo,C] = maxpool3x3/s2/p1(relu(x*scale+shift)); idx = one byte arg-max tap per element (uint64 per 8 ch)
 * replaces bn1 -> relu -> maxpool, classification/resnet/models/networks.py:207-209 */
int b200_bn_relu_maxpool_fwd(const void* x, void* y, void* idx, const float* scale, const float* shift, int B, int H,
                             int W, int C, void* stream);
int b200_maxpool_bwd(const void* g_out, const void* idx, void* g_in, int B, int H, int W, int C, void* stream);
/* global average pool, classification/resnet/models/networks.py:216 */
int b200_avgpool_fwd(const void* x, void* y, int B, int HW, int C, void* stream);
int b200_avgpool_bwd(const void* gy, void* gx, int B, int HW, int C, void* stream);

/* ---- loss / misc ------------------------------------------------------------------------------------------------------
 * CrossEntropyLoss(mean) forward+backward; classification/resnet/train.py:104, utils.py:39-42.
 *   loss_rows[B] per-sample loss; dlogits (optional) bf16 [B][ld_d] = (softmax - onehot)*gscale; correct (optional) int[B] */
int b200_softmax_xent(const float* logits, long long ld, const long long* labels, int B, int N, float gscale,
                      float* loss_rows, void* dlogits, long long ld_d, int* correct, void* stream);
/* the same with a target DISTRIBUTION per sample: soft_targets fp32 [B][ld_soft] (timm SoftTargetCrossEntropy behind
 * Mixup / CutMix, classification/swin_transformer/main.py:111-113) or, with soft_targets == NULL, hard labels smoothed by
 * `smoothing` (LabelSmoothingCrossEntropy, main.py:114-115):  loss_b = sum_c t_c (lse - x_c),
 * dlogits = (softmax * sum_c t_c - t) * gscale; correct[b] compares the arg-max with the label (or the arg-max of t). */
int b200_softmax_xent_soft(const float* logits, long long ld, const long long* labels, const float* soft_targets,
                           long long ld_soft, float smoothing, int B, int N, float gscale, float* loss_rows, void* dlogits,
                           long long ld_d, int* correct, void* stream);
int b200_mean(const float* v, int n, float* out, void* stream);
int b200_colsum_bf16(const void* m, long long rows, long long ld, int cols, float* out, int accumulate, void* stream);

/* weight packing fp32 OIHW -> bf16 GEMM operand; mode 0: [O][taps*I] (pitch ld_dst), mode 1: [I][taps*O] */
int b200_pack_weight(const float* src, void* dst, int O, int I, int taps, int mode, long long ld_dst, void* stream);
/* all weights of a model in one launch: table[n][10] int64 {src, dst, O, I, taps, mode, ld_dst, first_block, rows_out,
 * oscale (optional fp32 [O] multiplier per output channel, 0 = none)}; mode 0 = forward / wgrad operand [O][tap*I+i],
 * 1 = dgrad operand [I][tap*O+o], 2 = space-to-depth stem operand [O][256] of a [O][3][7][7] kernel */
int b200_pack_weights_multi(const void* table, int n_entries, int total_blocks, void* stream);
int b200_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
int b200_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream);
/* stem im2col from the user's NCHW fp32 batch: a bf16 [B*Ho*Wo][ldk], k=(kh*KW+kw)*Cin+c (networks.py:206 conv1 7x7/2) */
int b200_im2col_nchw(const float* x, void* a, int B, int Cin, int H, int W, int KH, int KW, int stride, int pad,
                     int ldk, void* stream);

/* Space-to-depth stem - the conv1 7x7 / stride 2 / pad 3 of ResNet (classification/resnet/models/networks.py:150,206) without a
 * patch matrix: b200_stem_s2d writes z bf16 [B][H/2+3][W/2+3][16] (zero-padded input, 2x2 pixel phase folded into 12 + 4
 * zero channels); the conv is then a 4x4 / stride-1 conv whose four x-taps are 64 contiguous elements of z, presented to the
 * implicit-GEMM kernels through a tensor map with overlapping rows. w = b200_pack_weights_multi mode 2 ([64][256]).
 * y bf16 [B][Ho][Wo][64] (Ho = H/2), stats as for b200_conv2d_fwd (rows: b200_conv2d_fwd_stats_rows(B, Ho, Wo, 64, 3, 1)).
 * wgrad: g fp32 [64][64][4] scratch gradient in the operand layout -> b200_stem_s2d_wgrad_relayout -> dW [64][3][7][7]. */
int b200_stem_s2d(const float* x, void* z, int B, int H, int W, void* stream);
/* GPU input pipeline (replaces the CPU ToTensor + Normalize of classification/resnet/train.py:46-71):
 * decoded uint8 NHWC [B][H][W][3] -> the same space-to-depth operand (ResNet), or -> normalised fp32 NCHW (other families).
 * mean3 / std3 are HOST pointers to 3 floats each (the reference's [0.485, 0.456, 0.406] / [0.229, 0.224, 0.225]). */
int b200_stem_s2d_u8(const void* x_u8_nhwc, void* z, int B, int H, int W, const float* mean3, const float* std3, void* stream);
int b200_normalize_u8_nhwc(const void* x_u8_nhwc, float* y_nchw, int B, int H, int W, const float* mean3, const float* std3,
                           void* stream);
int b200_stem_s2d_conv_fwd(const void* z, const void* w, void* y, float* stats, int B, int Ho, int Wo, void* stream);
size_t b200_stem_s2d_conv_wgrad_workspace_bytes(int B, int Ho, int Wo);
int b200_stem_s2d_conv_wgrad(const void* dy, const void* z, float* g, void* workspace, size_t workspace_bytes, int B, int Ho,
                             int Wo, void* stream);
int b200_stem_s2d_wgrad_relayout(const float* g, float* dw, int accumulate, void* stream);

/* stem weight gradient [Cout][ldk] (k = tap*Cin + c, as produced by b200_conv2d_wgrad on the patch matrix) -> OIHW */
int b200_stem_wgrad_relayout(const float* src, float* dst, int Cout, int Cin, int taps, int ldk, int accumulate,
                              void* stream);

/* fused SGD(momentum) over a flat fp32 arena; torch.optim.SGD semantics (classification/resnet/train.py:96).
 * lr_dev (optional device float*) overrides lr, so a captured CUDA graph can follow the reference's LambdaLR schedule. */
int b200_sgd_momentum(float* p, const float* g, float* buf, long long n, float lr, const float* lr_dev, float momentum,
                      float weight_decay, float gscale, int first_step, const float* clip_coef, void* stream);

/* ---- train-mode BatchNorm folded through a 1x1 convolution (ResNet bottleneck conv3 -> bn3 -> +identity -> ReLU,
 * classification/resnet/models/networks.py:116-124): the wide conv output is never written; see csrc/bn_algebra.cuh.
 * forward: G = y2^T y2 (b200_conv2d_wgrad(y2, y2)), s = colsum(y2) -> b200_bn_gram_stats -> scale / shift ->
 *          b200_conv1x1_bn_act_fwd: y = relu(conv1x1(y2, w) * scale + shift + residual)       (w bf16 [Cout][Cin])
 * backward: dz = relu-mask * gradient (b200_conv1x1_dgrad_masked of the NEXT block, with its per-CTA column sums),
 *          D = dz^T y2 (b200_conv2d_wgrad) -> b200_bn_conv1x1_bwd -> dgamma, dbeta, dW, wcat [Cin][Cout + Cin] bf16, bias [Cin]
 *          -> b200_gemm_dual: g2 = [dz | y2] wcat^T + bias                                     */
int b200_bn_gram_stats(const float* G, const float* s, const void* w_bf16, int N, int K, double count, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                       long long* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift, void* stream);
int b200_conv1x1_bn_act_fwd(const void* x, const void* w, const float* scale, const float* shift, const void* residual,
                            void* y, long long pixels, int Cin, int Cout, int relu, void* stream);
/* y[pixels][Cout] = conv1x1(x, w) * scale + shift: the downsample branch conv -> BatchNorm (networks.py:118-119, 196-199) with the
 * batch statistics from b200_bn_gram_stats of its (compact, see b200_subsample2) input; no residual, no ReLU */
int b200_conv1x1_bn_fwd(const void* x, const void* w, const float* scale, const float* shift, void* y, long long pixels,
                        int Cin, int Cout, void* stream);
/* xs[b][i][j][:] = x[b][2i][2j][:] (the pixels a 1x1 / stride-2 convolution reads);  gx[b][2i][2j][:] += gs[b][i][j][:] */
int b200_subsample2(const void* x, void* xs, int B, int H, int W, int C, void* stream);
int b200_add_even_pixels(void* gx, const void* gs, int B, int H, int W, int C, void* stream);
/* dx[pixels][Cin] = (mask_src > 0) ? (dy[pixels][Cout] * wd^T + residual) : 0  (wd bf16 [Cin][Cout], b200_pack_weight mode 1);
 * stats fp32 [b200_conv1x1_dgrad_masked_stats_rows()][2][Cin]: per-CTA column sums (plane 0) of dx as stored */
int b200_conv1x1_dgrad_masked_stats_rows(long long pixels, int Cin);
int b200_conv1x1_dgrad_masked(const void* dy, const void* wd, void* dx, long long pixels, int Cin, int Cout,
                              const void* residual, const void* mask_src, float* stats, void* stream);
/* dz_partial fp32 [T][2][N] (plane 0 = partial column sums of dz); scratch: b200_bn_conv1x1_bwd_scratch_bytes(N, K) bytes;
 * tickets: 64 uint32 counters, zero before the first call (the kernel leaves them zero) */
size_t b200_bn_conv1x1_bwd_scratch_bytes(int N, int K);
int b200_bn_conv1x1_bwd(const float* dz_partial, int T, const float* D, const float* G, const float* s, const void* w_bf16,
                        const float* w_f32, int N, int K, double count, const float* gamma, const float* mean,
                        const float* invstd, float* dgamma, float* dbeta, float* dW, int accumulate, void* wcat, float* bias,
                        void* scratch, size_t scratch_bytes, void* tickets, void* stream);
/* out[pixels][N] (bf16) = [a0[pixels][K0] | a1[pixels][K1]] * wcat[N][K0 + K1]^T + bias[N] */
int b200_gemm_dual(const void* a0, int K0, const void* a1, int K1, const void* wcat, const float* bias, void* out,
                   long long pixels, int N, void* stream);

/* ---- stochastic depth / pre_logits helpers -------------------------------------------------------------------------------
 * y[b, :] = x[b, :] * scale[b] over bf16 samples of elems_per_sample elements: the gradient entering a residual branch whose
 * forward was scaled per sample by drop_path (convNext/models/networks.py:11-26, vit_model.py:12-40, swin_transformer.py:282,285;
 * the forward scaling itself is b200_gemm_args_t::rowscale).  Samples with scale 0 are not read. */
int b200_rowscale_bf16(const void* x, const float* scale, void* y, long long n_samples, long long elems_per_sample,
                       void* stream);
/* ViT pre_logits = Linear + Tanh on the class-token row (vit_model.py:218-221): t = tanh(u) fp32 (kept for the backward) and
 * its bf16 copy (operand of the classifier GEMM); backward du = dt * (1 - t^2), bf16 in / out. */
int b200_tanh_fwd(const float* u, float* t, void* t_bf16, long long n, void* stream);
int b200_tanh_bwd(const void* dt_bf16, const float* t, void* du_bf16, long long n, void* stream);

/* bring-up only: override the UMMA shared-memory descriptor strides (which: 0 = forward K-major, 1 = wgrad MN-major) */
int b200_debug_set_desc(int which, unsigned lbo, unsigned sbo, unsigned kstep);

#ifdef __cplusplus
}
#endif
#endif /* B200CLS_H_ */
