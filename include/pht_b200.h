/*
 * pht_b200.h -- C ABI of libpht_b200.so: the B200 (sm_100a) kernels behind the
 * AFGSA denoiser hot path of goodbadwolf/pixel_heal_thyself (PHT).
 *
 * The reference has no native code: every entry point below replaces a stock
 * torch call site of the reference (cited as file:line relative to the
 * reference checkout).  The host side (pixel_heal_thyself_b200/, Python) binds
 * these with ctypes; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless
 *     the name ends in _host.  No allocation, no ownership transfer: every
 *     buffer (workspaces included) is owned by the caller.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 on success, negative pht_status on failure
 *     (pht_last_error() returns a thread-local message).  Python raises.
 *   - activations are channels-last ("NHWC") with an explicit strided view so
 *     that padded buffers, interior views and channel slices need no copies.
 *   - dtype: PHT_F32 (parity mode, CUDA-core fp32 math) or PHT_BF16 (production:
 *     bf16 storage, fp32 accumulation, tcgen05 tensor cores where shapes allow).
 */
#ifndef PHT_B200_H
#define PHT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHT_ABI_VERSION 3

enum pht_status {
  PHT_OK = 0,
  PHT_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  PHT_ERR_CUDA = -2,      /* CUDA runtime / driver error */
  PHT_ERR_UNSUPPORTED = -3
};

enum pht_dtype { PHT_F32 = 0, PHT_BF16 = 1 };
enum pht_pad_mode { PHT_PAD_REPLICATE = 0, PHT_PAD_REFLECT = 1 };

/* Strided channels-last view.  Element (b, y, x, c) lives at
 * ptr + b*sb + y*sy + x*sx + c  (strides in ELEMENTS, channel stride 1).
 * Reads outside 0<=y<H, 0<=x<W return 0 (this is how zero padding, dgrad's
 * implicit zero border and ragged tiles are expressed).  (oy, ox) is added to
 * the pixel coordinate before the bounds test: an output pixel (y, x) with tap
 * (dy, dx) reads view pixel (y + dy + oy, x + dx + ox). */
typedef struct pht_view {
  void* ptr;
  int32_t H, W, C;
  int32_t oy, ox;
  int32_t dtype;            /* pht_dtype of the elements */
  int64_t sb, sy, sx;
} pht_view;

/* Epilogue flags of pht_conv_gemm */
#define PHT_EPI_RESID_PRE  1u  /* v += resid before the activation            */
#define PHT_EPI_RESID_POST 2u  /* out2 = v + resid                            */
#define PHT_EPI_MASK       4u  /* out2 *= (mask > 0 ? 1 : mslope[n])          */
/* Backward of REPLICATE padding fused into a data-gradient: the output domain (Ho, Wo) = (H+2, W+2) is the padded
 * domain; before the epilogue runs, the value of every border pixel is added to the interior pixel it was replicated
 * from (== pht_pad_fold with PHT_PAD_REPLICATE).  resid / mask are views of the INTERIOR (H x W; padded pixel (y, x)
 * reads view pixel (y-1, x-1)); out1 / out2 are views of PADDED buffers (H+2 x W+2): their interior holds the result,
 * their 1-pixel frame receives don't-care values.  bf16 tensor-core path only (ksize 3, N <= 256, H % 8 == 0,
 * W % 8 == 0); otherwise PHT_ERR_UNSUPPORTED. */
#define PHT_EPI_PADFOLD    8u
/* out1 (RING1) / out2 (RING2) is the INTERIOR view (H x W) of a padded buffer [B][H+2][W+2][C]: besides the interior,
 * the kernel also writes the buffer's 1-pixel frame -- the replicate (default) or reflect (RING_REFLECT) padding that
 * pht_border_fill would write afterwards (every frame pixel is a copy of one output pixel, stored by the thread that owns
 * it).  bf16 tensor-core path only, H, W >= 4, not together with PHT_EPI_PADFOLD; otherwise PHT_ERR_UNSUPPORTED. */
#define PHT_EPI_RING1        16u
#define PHT_EPI_RING2        32u
#define PHT_EPI_RING_REFLECT 64u

/* Implicit-GEMM convolution over pixels with virtual concat:
 *   acc[p, n] = sum_t sum_s sum_k src[s](p + tap_t)[k] * w[t][n][koff_s + k]
 *   v    = act(acc + bias[n] (+ resid[p,n] if RESID_PRE)),  act(v) = v>0 ? v : v*slope[n]
 *   out1 = v                                   (if out1.ptr)
 *   out2 = (v (+ resid if RESID_POST)) (* dact(mask) if MASK)   (if out2.ptr)
 * Taps are the ksize x ksize centred stencil (ksize in {1,3,5}), t = ky*ksize+kx.
 * Weights are packed [ksize*ksize][N][Ktot] in `dtype`, Ktot = sum of src[s].C.
 * bias / slope / mslope are fp32 [N] or NULL (NULL slope = no activation).
 * Output domain is B x Ho x Wo pixels, N channels.
 *
 * Replaces: nn.Conv2d(+ReLU/LeakyReLU) forward and data-gradient for every
 * 1x1 / 3x3 conv of AFGSANet (model.py:606-658, 449-452, 552-569, 690-706),
 * torch.cat (model.py:462,722,727), the residual adds (model.py:576-581) and
 * the ReLU backward masks (autograd). */
typedef struct pht_conv_gemm_args {
  int32_t dtype;            /* compute/storage dtype of src, w, out            */
  int32_t B, Ho, Wo, N;
  int32_t ksize;
  int32_t n_src;
  uint32_t flags;
  pht_view src[3];
  const void* w;
  const float* bias;
  const float* slope;
  const float* mslope;
  pht_view resid;
  pht_view mask;
  pht_view out1;
  pht_view out2;
} pht_conv_gemm_args;

int pht_conv_gemm(const pht_conv_gemm_args* args, void* stream);

/* Weight gradient of the op above:
 *   dw[t][n][koff_s + k] (+)= sum_p dy[p, n] * src[s](p + tap_t)[k],  dbias[n] = sum_p dy[p, n]
 * dw is fp32 [ksize*ksize][N][Ktot] and is OVERWRITTEN; dbias fp32 [N] or NULL.
 * workspace: pht_wgrad_workspace_bytes(args) bytes (split-K partials).
 * Replaces cuDNN convolution_backward(weight, bias) (autograd of model.py convs). */
typedef struct pht_wgrad_args {
  int32_t dtype;
  int32_t B, Ho, Wo, N;
  int32_t ksize;
  int32_t n_src;
  pht_view dy;
  pht_view src[3];
  float* dw;
  float* dbias;
  void* workspace;
  size_t workspace_bytes;
} pht_wgrad_args;

size_t pht_wgrad_workspace_bytes(const pht_wgrad_args* args);
int pht_wgrad(const pht_wgrad_args* args, void* stream);

/* Deferred variant for batching: pht_wgrad_partial() runs only the split GEMM and leaves the fp32 partials in
 * args->workspace (which must stay untouched until the reduce); *job receives the descriptor of the pending
 * fixed-order reduction (into args->dw / args->dbias).  Returns PHT_ERR_UNSUPPORTED when the shape is not taken by the
 * split tensor-core kernel (call pht_wgrad instead).  pht_wgrad_reduce_batched() then finishes ANY number of pending
 * jobs in ONE launch (`jobs` is a host array; `table_dev` caller-owned device scratch of n * sizeof(job) bytes,
 * upload != 0 (re)writes it). */
typedef struct pht_wgrad_reduce_job {
  const float* partials;    /* [splits][elems] */
  float* dw;                /* [elems] */
  const float* bias_partials; /* [bias_rows][N] or NULL */
  float* dbias;             /* [N] */
  int64_t elems;            /* multiple of 4 */
  int32_t splits, bias_rows, N, pad_;
} pht_wgrad_reduce_job;
int pht_wgrad_partial(const pht_wgrad_args* args, pht_wgrad_reduce_job* job, void* stream);
int pht_wgrad_reduce_batched(const pht_wgrad_reduce_job* jobs, int32_t n, void* table_dev, size_t table_bytes, int32_t upload,
                             void* stream);

/* Fill the 1-pixel border of a padded NHWC buffer [B][H+2][W+2][C] from its
 * interior (replicate = clamp to edge, reflect = mirror without the edge).
 * Replaces F.pad inside nn.Conv2d(padding_mode=...) (model.py:607-622, 552-569,
 * 690-706; mode chosen at base_trainer.py:334). */
int pht_border_fill(void* buf, int32_t dtype, int32_t B, int32_t H, int32_t W, int32_t C, int32_t mode, void* stream);

/* FiLM modulation of the AFGSA layer's FiLM variant (use_film=True): replaces FiLM.forward's
 * `gamma, beta = torch.chunk(gamma_beta, 2, dim=1); return gamma * x + beta` (pht/models/afgsa/film.py:36-45, spatial
 * gamma/beta as built by AFGSA, model.py:443-449) and its autograd.  gb holds [gamma | beta] (2C channels), x / out C
 * channels, all views of one dtype.
 *   fwd: out = gb[..., :C] * x + gb[..., C:]
 *   bwd: dgb[..., :C] = dout * x, dgb[..., C:] = dout, and (dx non-null) dx = (dx_in ? dx_in : 0) + gb[..., :C] * dout
 *        (dx may alias dx_in). */
int pht_film_fwd(const pht_view* gb, const pht_view* x, const pht_view* out, int32_t B, int32_t C, void* stream);
int pht_film_bwd(const pht_view* gb, const pht_view* x, const pht_view* dout, const pht_view* dgb, const pht_view* dx_in,
                 const pht_view* dx, int32_t B, int32_t C, void* stream);

/* Backward of the padding: fold the border of a padded-domain gradient
 * gpad[B][H+2][W+2][C] into the interior, then
 *   out1 = fold(gpad) (+ resid)          (if out1.ptr)
 *   out2 = out1 * dact(mask)             (if out2.ptr; mslope as in conv_gemm)
 * Replaces replication_pad2d_backward / reflection_pad2d_backward + ReLU bwd. */
int pht_pad_fold(const void* gpad, int32_t dtype, int32_t B, int32_t H, int32_t W, int32_t C, int32_t mode,
                 const pht_view* resid, const pht_view* mask, const float* mslope,
                 const pht_view* out1, const pht_view* out2, void* stream);

/* 5x5 im2col of a tiny-channel NCHW fp32 image (Cin = 3 or 7) into a row-major
 * [B*H*W][Kpad] matrix of `dtype`, k = (ky*5+kx)*Cin + ci, zero-filled to Kpad;
 * out-of-image taps follow `mode` (clamp / mirror).  The 1x1, 3x3 and 5x5
 * encoder branches (model.py:606-646, 719-726) then run as ONE dense GEMM with
 * their kernels embedded in the 5x5 tap grid. */
int pht_im2col5(const float* x_nchw, void* col, int32_t dtype, int32_t B, int32_t Cin, int32_t H, int32_t W,
                int32_t Kpad, int32_t mode, void* stream);

/* Block-local auxiliary-feature-guided self attention, model.py:474-516.
 * q (pre-scaled by d^-1/2), k, v: views of C = heads*64 channels; window =
 * block + 2*halo; keys outside the image are zero but NOT masked; rel_h /
 * rel_w fp32 [win][d/2] are added to the key halves after padding.
 *   out[p, h*d+j] = resid[p, ...] + sum_key softmax(q.k')[key] v[key]
 * lse: fp32 [B*H*W][heads] (log-sum-exp per query/head, saved for backward). */
typedef struct pht_attn_args {
  int32_t dtype;
  int32_t B, H, W;
  int32_t heads, head_dim, block, halo;
  pht_view q, k, v;
  const float* rel_h;
  const float* rel_w;
  pht_view resid;           /* optional (ptr may be NULL)                      */
  pht_view out;
  float* lse;
  int32_t ring;             /* forward only: 1 (replicate) / 2 (reflect): `out` is the INTERIOR view of a padded buffer
                             * [B][H+2][W+2][C] and the kernel also writes that buffer's 1-pixel frame (what pht_border_fill
                             * would write afterwards); bf16 tensor-core path only, else PHT_ERR_UNSUPPORTED.  0 = no frame */
  int32_t pad_;
} pht_attn_args;

int pht_attn_fwd(const pht_attn_args* args, void* stream);

/* Recompute-based backward of the op above (autograd of model.py:474-516):
 * given d_out, recomputes P from q, k, lse and produces dq, dk, dv (views in
 * the activations' dtype, OVERWRITTEN; the up-to-4 overlapping window
 * contributions of a key pixel are summed inside the op) and
 * d_rel_h / d_rel_w fp32 [win][d/2] (OVERWRITTEN).
 * workspace: pht_attn_bwd_workspace_bytes() bytes (relative-position partial sums; window-major dK/dV scratch when the
 * "attn_bwd_direct" option is 0; fp32 accumulators for the CUDA-core path).
 * With "attn_bwd_direct" = 1 (default) the tcgen05 path accumulates the window contributions straight into dk / dv
 * with vector reductions, so they must be zero when the kernel starts: pht_attn_bwd does that itself unless `prezeroed`
 * is set, in which case the caller has already run pht_attn_bwd_zero with the same arguments (e.g. on a side stream,
 * overlapped with earlier kernels). */
typedef struct pht_attn_bwd_args {
  pht_attn_args fwd;        /* q, k, v, rel_*, lse as in the forward; out/resid unused */
  pht_view d_out;
  pht_view dq;
  pht_view dk;
  pht_view dv;
  float* d_rel_h;
  float* d_rel_w;
  void* workspace;
  size_t workspace_bytes;
  int32_t prezeroed;        /* 1: pht_attn_bwd_zero(args) has already been run for this launch */
  int32_t pad_;
} pht_attn_bwd_args;

size_t pht_attn_bwd_workspace_bytes(const pht_attn_bwd_args* args);
int pht_attn_bwd_zero(const pht_attn_bwd_args* args, void* stream);
int pht_attn_bwd(const pht_attn_bwd_args* args, void* stream);

/* Decoder tail: 3x3 conv 256->3 with ZERO padding, no activation, plus the
 * residual with the network input (model.py:707-714, 732).
 *   out_nchw[b,co,y,x] = x_nchw[b,co,y,x] + bias[co] + sum_{t,c} h(y+dy,x+dx)[c] * w[co][t][c]
 * w fp32 [3][9][C]. */
int pht_dec_tail_fwd(const pht_view* h, const float* w, const float* bias, const float* x_nchw, float* out_nchw,
                     int32_t B, int32_t H, int32_t W, void* stream);
/* dh_pre[p,c] = (sum_{t,co} dout(p - tap_t)[co] * w[co][t][c]) * (h[p,c] > 0)  (ReLU of decoder.1 fused) */
int pht_dec_tail_bwd_data(const float* dout_nchw, const float* w, const pht_view* h, const pht_view* dh_pre,
                          int32_t B, int32_t H, int32_t W, void* stream);
/* dw fp32 [3][9][C] and dbias fp32 [3], both OVERWRITTEN. workspace >= pht_dec_tail_ws_bytes(). */
size_t pht_dec_tail_ws_bytes(int32_t B, int32_t H, int32_t W, int32_t C);
int pht_dec_tail_bwd_weight(const float* dout_nchw, const pht_view* h, float* dw, float* dbias, void* workspace,
                            size_t workspace_bytes, int32_t B, int32_t H, int32_t W, void* stream);

/* L1 reconstruction loss, fused forward+backward (losses.py:175-184,
 * base_trainer.py:423):  loss[0] = mean|a-b| (OVERWRITTEN),
 * grad[i] = grad_scale * sign(a[i]-b[i]) / n  (grad may be NULL).  The block partials of the deterministic
 * two-level reduction live in library scratch private to (device, stream): launches on different streams may overlap. */
int pht_l1_loss(const float* a, const float* b, int64_t n, float grad_scale, float* loss, float* grad, void* stream);

/* BatchNorm2d (training mode, affine) + LeakyReLU of the WGAN-GP critic's conv blocks (DiscriminatorVGG,
 * model.py:264-344; norm :52-61, act :64-83), with the first- and the second-order backward that the gradient penalty
 * (losses.py:12-57) needs.  All tensors fp32, channels-last: x / z / gz / gx / h [m = B*H*W][C]; gamma / beta /
 * run_mean / run_var [C]; stat [2][C] = batch mean and 1/sqrt(var + eps), written by the forward and read by both backward
 * passes.  C = 4 x a power of two, <= 1024.  workspace >= pht_bn_act_ws_bytes(C), 16-byte aligned, ZEROED by the
 * caller before its first use (it holds the "last block done" ticket of the deterministic reduction, which every call leaves
 * at zero again; calls sharing a workspace must be stream-ordered).
 *   fwd      z = leaky(gamma (x - mean) rstd + beta); run_mean / run_var (may be NULL) updated like nn.BatchNorm2d.
 *            pre_bias (may be NULL): the bias [C] of the convolution in front, NOT added to x by the caller -- batch
 *            normalisation cancels a per-channel constant exactly (z, stat and every gradient are those of x), only the
 *            running mean sees it: run_mean tracks mean(x) + pre_bias, as nn.Conv2d(bias=True) + nn.BatchNorm2d would
 *   bwd      gx, g_gamma, g_beta (the latter two may be NULL) from gz
 *   bwd_bwd  cotangent h of gx -> h_gz (w.r.t. gz), h_x (w.r.t. x, the dependence of mean / rstd on x included),
 *            h_gamma (may be NULL).  Per-channel sums accumulate in fp64 in a fixed order (deterministic). */
size_t pht_bn_act_ws_bytes(int32_t C);
/* out[c] = sum over m rows of x[m][C] (fp32): the conv bias gradients of the critic; same workspace / channel rules */
int pht_colsum_f32(const float* x, float* out, int64_t m, int32_t C, void* workspace, size_t workspace_bytes, void* stream);
int pht_bn_act_fwd(const float* x, const float* gamma, const float* beta, const float* pre_bias, float* run_mean, float* run_var,
                   float* stat, float* z, int64_t m, int32_t C, float eps, float momentum, float slope, void* workspace,
                   size_t workspace_bytes, void* stream);
int pht_bn_act_bwd(const float* x, const float* gz, const float* gamma, const float* beta, const float* stat, float* gx,
                   float* g_gamma, float* g_beta, int64_t m, int32_t C, float slope, void* workspace, size_t workspace_bytes,
                   void* stream);
int pht_bn_act_bwd_bwd(const float* x, const float* gz, const float* h, const float* gamma, const float* beta, const float* stat,
                       float* h_gz, float* h_x, float* h_gamma, int64_t m, int32_t C, float slope, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Optional MS-SSIM + L1 image loss, fused forward + backward (SSIMLoss, losses.py:248-263, used at
 * base_trainer.py:450-452 with weight 0.1): per-pixel scale = max(channel-max of gt, 1); kornia 0.8.0
 * MS_SSIMLoss(reduction="mean") arithmetic on out / scale, gt / scale (kornia is a third-party dependency that is not
 * vendored with the reference: PARITY UNPINNED, the checker is oracle/msssim_oracle.py).  out / gt: fp32 NCHW
 * [B][3][H][W]; loss[0] OVERWRITTEN; grad (may be NULL) = grad_scale * d loss / d out, fp32 NCHW, OVERWRITTEN.
 * workspace >= pht_msssim_ws_bytes(B, H, W), 16-byte aligned. */
size_t pht_msssim_ws_bytes(int32_t B, int32_t H, int32_t W);
int pht_msssim_loss(const float* out_nchw, const float* gt_nchw, int32_t B, int32_t H, int32_t W, float grad_scale, float* loss,
                    float* grad, void* workspace, size_t workspace_bytes, void* stream);

/* Batch preprocessing (base_trainer.py:373-383; preprocessing.py:19-22,34-38):
 * NHWC fp32 patches -> NCHW fp32: noisy/gt = log(v+1); aux[0:3] =
 * clamp((nan_to_num(n)+1)/2, 0, 1); aux[3:7] unchanged.  gt may be NULL. */
int pht_preprocess(const float* noisy_nhwc, const float* gt_nhwc, const float* aux_nhwc, float* noisy_nchw,
                   float* gt_nchw, float* aux_nchw, int32_t B, int32_t H, int32_t W, void* stream);

/* Patch crop + the preprocessing above in one pass (preprocessing.py:325-344,
 * gen_hdf5.py:135-139 + base_trainer.py:373-383): frames NHWC fp32
 * [n_img][Hf][Wf][3|7] resident in HBM; patch i is the PxP window centred at
 * centres[i] = (x, y) (int32 [n][2]) of frame img_idx[i] (int32 [n], NULL = frame 0);
 * output NCHW fp32 [n][C][P][P]. */
int pht_crop_preprocess(const float* noisy_f, const float* gt_f, const float* aux_f, int32_t Hf, int32_t Wf,
                        const int32_t* centres, const int32_t* img_idx, int32_t n, int32_t P, float* noisy_nchw,
                        float* gt_nchw, float* aux_nchw, void* stream);

/* Adam over a flat fp32 parameter arena (base_trainer.py:182-187; torch.optim
 * Adam, betas (0.9,0.999), eps 1e-8, no weight decay).  step is 1-based. */
int pht_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
             int32_t step, float grad_scale, void* stream);
/* The same update with the step-dependent quantities resident on the device, so that the launch can be replayed from a
 * CUDA graph: hyper = float[4] {lr, step count so far (int32 bit pattern), scratch, scratch}.  A one-thread prelude
 * kernel increments the step count and derives lr / (1 - beta1^step) and 1 / sqrt(1 - beta2^step) in fp64 (as
 * pht_adam does on the host); the caller changes the learning rate by writing hyper[0]. */
int pht_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, float beta1, float beta2, float eps, float grad_scale,
                 float* hyper, void* stream);

/* Weight (re)packing between the reference's OIHW fp32 parameters and the
 * kernels' [tap][N][K] layout.
 *   pack:    dst[t][n_off+o][k_off+i] = scale * w[o][i][ky][kx]   (transpose=0)
 *            dst[t'][n_off+i][k_off+o] = scale * w[o][i][ky][kx], t' = flipped tap (transpose=1, dgrad)
 *            dst[n_off+i][k_off + t*O + o] = scale * w[o][i][ky][kx]  (transpose=2: taps x outputs folded into K,
 *                                             no flip; the decoder tail's data-/weight-gradient GEMMs)
 *   embed:   the ksize x ksize kernel is centred inside a `grid` x `grid` tap
 *            grid whose taps are folded into K (used for the im2col5 encoders):
 *            dst[n_off+o][((ky+e)*grid + kx+e)*I + i], e = (grid-ksize)/2.
 *   unpack:  the inverse, accumulating nothing: w_grad[o][i][ky][kx] = scale * src[...]. */
typedef struct pht_pack_args {
  const float* w;           /* OIHW fp32 (pack: source; unpack: destination, cast away const) */
  void* packed;             /* packed tensor (dtype below for pack; fp32 for unpack)          */
  int32_t dtype;
  int32_t O, I, ksize;
  int32_t Ntot, Ktot;       /* extents of the packed tensor                                   */
  int32_t n_off, k_off;
  int32_t transpose;
  int32_t grid;             /* 0 = taps stay a separate leading dim; 5 = embed into K         */
  int32_t i_begin, i_count; /* input-channel slice of w that is packed (i_count 0 = all of I);  */
                            /* packed input index = i - i_begin                                 */
  float scale;
} pht_pack_args;

int pht_pack_weight(const pht_pack_args* a, void* stream);
int pht_unpack_wgrad(const pht_pack_args* a, void* stream);

/* All weight (re)packs of a step in ONE launch.  `jobs` is a HOST array of n descriptors (same meaning as
 * pht_pack_weight; jobs[i].dtype is the dtype of jobs[i].packed).  `table_dev` is caller-owned device scratch of
 * pht_pack_table_bytes(n) bytes; upload != 0 (re)writes the device copy of the descriptor table (needed the first
 * time and whenever a pointer or shape in `jobs` changed), upload == 0 reuses it. */
size_t pht_pack_table_bytes(int32_t n);
int pht_pack_weights_batched(const pht_pack_args* jobs, int32_t n, void* table_dev, size_t table_bytes, int32_t upload,
                             void* stream);
/* the same for pht_unpack_wgrad (jobs[i].packed is the fp32 packed gradient, jobs[i].w the OIHW destination) */
int pht_unpack_wgrads_batched(const pht_pack_args* jobs, int32_t n, void* table_dev, size_t table_bytes, int32_t upload,
                              void* stream);

/* Decoder tail on the tensor-core path (model.py:707-714, 732): the 256->3 zero-padded 3x3 conv runs as a 64-wide
 * pht_conv_gemm with fp32 output y [B*H*W][ldy]; this adds bias and the residual and transposes to NCHW:
 *   out_nchw[b,co,y,x] = y[p][co] + bias[co] + x_nchw[b,co,y,x]. */
int pht_tail_finish(const float* y, int32_t ldy, const float* bias, const float* x_nchw, float* out_nchw, int32_t B,
                    int32_t H, int32_t W, void* stream);
/* The same decoder tail from a 1x1 GEMM: y fp32 [B][H][W][ldy] with y[p][t * 3 + co] = sum_c in(p)[c] w[co][c][t]
 * (t = ky * 3 + kx) -> out_nchw(p)[co] = bias[co] + x_nchw(p)[co] + sum_t y[p + (ky - 1, kx - 1)][t * 3 + co], zero
 * padding (model.py:707-714, 731-732).  The 256-channel decoder activation is read once instead of nine times. */
int pht_tail_gather(const float* y, int32_t ldy, const float* bias, const float* x_nchw, float* out_nchw, int32_t B, int32_t H,
                    int32_t W, void* stream);
/* Backward side: a[p][t*3+co] = dout[p - tap_t][co] (bf16 [B*H*W][64], columns 27..63 zero) so that
 * d(decoder.1 output) = a @ Wt and d(decoder.2 weight) = h^T a are plain GEMMs; dbias[co] = sum_p dout[p][co]. */
int pht_tail_im2col_bwd(const float* dout_nchw, void* a_bf16, float* dbias, int32_t B, int32_t H, int32_t W, void* stream);

/* dtype conversion of a contiguous array (fp32 <-> bf16), n elements. */
int pht_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream);
/* same for a rows x cols matrix with leading dimensions (elements) src_ld / dst_ld */
int pht_cast2d(const void* src, int32_t src_dtype, int64_t src_ld, void* dst, int32_t dst_dtype, int64_t dst_ld,
               int64_t rows, int64_t cols, void* stream);

/* NHWC(dtype, C = 3) -> NCHW fp32 and back are folded into the ops above. */

/* Poisson-disk dart throwing, bit-exact with
 * sample_patches_dart_throwing(shape, P, n, random.Random(seed)) of
 * preprocessing.py:179-213 (CPython MT19937 randint semantics).  One CTA per
 * image: seeds int64 [n_img]; out int32 [n_img][n][2] = (x, y). */
int pht_sample_patches(const int64_t* seeds, int32_t n_img, int32_t Hf, int32_t Wf, int32_t P, int32_t n,
                       int32_t max_iter, int32_t* out, void* stream);

/* Importance map of the patch sampler (get_importance_map as called by importance_sampling,
 * preprocessing.py:119-168, 293-300): relative variance of the noisy radiance + variance of the normals over a
 * P x P box (scipy.ndimage.uniform_filter semantics: 'reflect' borders, window [i-P/2, i-P/2+P), double accumulation,
 * float32 after each separable pass), channel max, gamma 1/2.2, each map and the sum normalised by its maximum.
 * noisy_f [n_img][Hf][Wf][3], aux_f [n_img][Hf][Wf][7] (normal = channels 0..2) are the raw frames (the cleaning of
 * preprocess_data, :97-103, is applied on the fly); imp fp32 [n_img][Hf][Wf] is OVERWRITTEN.
 * workspace: pht_importance_map_ws_bytes() bytes. */
size_t pht_importance_map_ws_bytes(int32_t n_img, int32_t Hf, int32_t Wf);
int pht_importance_map(const float* noisy_f, const float* aux_f, int32_t n_img, int32_t Hf, int32_t Wf, int32_t P,
                       float* imp, void* workspace, size_t workspace_bytes, void* stream);

/* importance_sampling(data, P, n, random.Random(seed)) of preprocessing.py:284-322, bit-exact given the map:
 * pht_sample_patches' dart throwing followed, on the same RNG stream, by prune_patches (:259-281: serpentine
 * 4P-regions with inclusive bounds, float32 error diffusion against random.random()).
 * out_centres int32 [n_img][n][2] = kept patch CENTRES (x, y) in the reference's order, -1 padded;
 * out_counts int32 [n_img] = number kept (-1 if dart throwing failed). */
int pht_importance_sample(const int64_t* seeds, int32_t n_img, int32_t Hf, int32_t Wf, int32_t P, int32_t n,
                          int32_t max_iter, const float* imp, int32_t* out_centres, int32_t* out_counts, void* stream);

/* ---- validation metrics (pht/models/base_trainer.py:549-571) ----
 * pht_tonemap_u8: tensor2img (pht/models/afgsa/util.py:77-119): NCHW fp32 -> uint8 NHWC;
 *   v = post_spec ? exp(x) - 1 (postprocess_specular, preprocessing.py:46-48) : x;
 *   out = uint8(clip(clip(v ** (1/2.2), 0, 1) * 255, 0, 255)), NaN -> 0.
 * pht_image_metrics_u8: per image of the batch, on uint8 NHWC pairs:
 *   sqdiff[b] = sum (a - b)^2 (exact integer; calculate_psnr = 20 log10(255 / sqrt(sqdiff / (H W C))), metric.py:9-24);
 *   ssim_sum[b] = sum of the SSIM map over the valid (H-10) x (W-10) x C region, 11x11 gaussian (sigma 1.5) window,
 *   fp64 (metric.py:27-48; calculate_ssim = ssim_sum / ((H-10)(W-10)C), :51-73).
 * pht_mrse: out_sum[b] = sum (a - b)^2 / (b^2 + 0.01) with a = a_is_log ? exp(a) - 1 : a (calculate_rmse = 0.5 * mean,
 *   metric.py:76-94).  workspace: pht_image_metrics_ws_bytes(B) bytes, 8-byte aligned.  Deterministic reductions. */
int pht_tonemap_u8(const float* x_nchw, uint8_t* out_nhwc, int32_t B, int32_t C, int32_t H, int32_t W, int32_t post_spec,
                   void* stream);
size_t pht_image_metrics_ws_bytes(int32_t B);
int pht_image_metrics_u8(const uint8_t* a, const uint8_t* b, int32_t B, int32_t H, int32_t W, int32_t C, uint64_t* sqdiff,
                         double* ssim_sum, void* workspace, size_t workspace_bytes, void* stream);
int pht_mrse(const float* a, const float* b, int32_t B, int64_t n_per_image, int32_t a_is_log, double* out_sum, void* workspace,
             size_t workspace_bytes, void* stream);

/* Introspection */
int pht_abi_version(void);
const char* pht_last_error(void);
/* counters[0] = tcgen05 conv_gemm launches, [1] = CUDA-core conv_gemm launches,
 * [2] = tcgen05 wgrad, [3] = CUDA-core wgrad, [4] = tcgen05 attention, [5] = CUDA-core attention,
 * [6] = all other kernel launches.  Reset with pht_reset_counters(). */
void pht_get_counters(uint64_t* counters8);
void pht_reset_counters(void);
/* The counters are bumped by the host-side launch calls.  A caller that captured launches into a CUDA graph and replays
 * them adds the captured counts (the pht_get_counters delta over the capture) once per replay. */
void pht_add_counters(const uint64_t* counters8);
/* 1 = never use tcgen05 paths (debug / A-B testing) */
void pht_set_force_simple(int on);
/* tuning / A-B knobs: "tc_cfg" = 0 auto, 1 prefer the deep-ring conv_gemm config, 2 force the wide-epilogue one;
 * "wgrad_split_div" = d: 1x1 weight-gradients use 1/d of the pixel splits (fewer fp32 partials);
 * "pdl" = 0 / 1 (default 1): launch the tcgen05 kernels with programmatic dependent launch (their preamble overlaps
 * the previous kernel's tail);
 * "serpentine" = 0 / 1 (default 1): pht_conv_gemm walks its tiles opposite to the direction its first source was last
 * written in (the most recently written part of the input is still in L2);
 * "strips" = 0 / 1 (default 1): PHT_EPI_PADFOLD launches cover the last two rows of the padded domain with 2 x 64-pixel
 * tiles (fewer, fuller tiles); none of these three changes a result bit;
 * "cta_pairs" = 0 / 1 (default 0): 1 = pht_conv_gemm launches with 256-wide output tiles run as CTA pairs (2-CTA clusters,
 * tcgen05 cta_group::2, M = 256): the two CTAs of a pair take adjacent pixel tiles, each loads only half of every weight
 * tile, and the operand rings are 6 / 4 stages deep instead of 4 / 3.  Bit-identical results; measured no faster on the
 * AFGSA shapes (see profiles/README.md), hence opt-in;
 * "half_ring" = 0 / 1 (default 0): 1 = the 3x3 launches of pht_conv_gemm use an operand ring of 8 half stages (32 channels
 * deep, 64-byte swizzle rows) instead of 4 full ones.  Bit-identical; measured slower (the tensor core reads 64-byte swizzle
 * rows at about half the rate), kept for the record; "conv_grid_cap" = n: at most n CTAs per pht_conv_gemm launch (diagnostics);
 * "conv_trace" = 1: pht_conv_gemm records clock64 stamps per tile of CTA 0 and (start, end, SM, entry) times of every
 * CTA (diagnostics, read with pht_conv_gemm_trace);
 * "attn_trace" = 1 / 2: CTA 0 of pht_attn_bwd / pht_attn_fwd records clock64 stamps of its pipeline events (diagnostics);
 * "attn_bwd_direct" = 0 / 1 (default 1): 1 = the tcgen05 pht_attn_bwd adds the (up to four) overlapping window
 * contributions of a key pixel straight into dk / dv with vector reductions, in arrival order (fastest; every add rounds
 * to bf16, so the last bits depend on the order: not bit-reproducible from run to run); 0 = window-major bf16 scratch + a
 * fold kernel that sums them in a fixed order in fp32 (bit-reproducible; needs the larger workspace that
 * pht_attn_bwd_workspace_bytes reports while the option is 0);
 * "bf16_fallback" = 0 / 1 (default 0): a PHT_BF16 launch of pht_conv_gemm / pht_wgrad / pht_attn_fwd / pht_attn_bwd whose
 * shape or views the tcgen05 kernels do not take returns PHT_ERR_UNSUPPORTED instead of silently running the ~20x slower
 * CUDA-core kernel; 1 re-enables that fallback (pht_set_force_simple(1) always selects the CUDA-core kernels) */
int pht_set_option(const char* name, int value);
/* Copies the stamps recorded under "attn_trace" ([iteration][12 events], first 48 iterations of CTA 0) to host memory
 * after a device synchronise; returns the number of values copied or a negative status. */
int pht_attn_bwd_trace(int64_t* host, int32_t n);
/* Same for pht_conv_gemm under "conv_trace" = 1: [tile][8 events] of CTA 0 (first 32 tiles = 256 values), followed by
 * [cta][4] = (globaltimer ns after the dependency wait, ns at the end, SM id, ns at kernel entry) for up to 160 CTAs. */
int pht_conv_gemm_trace(int64_t* host, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* PHT_B200_H */
