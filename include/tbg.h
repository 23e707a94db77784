/*
 * tbg.h — C ABI of the B200-native TextBoxGAN training-step hot path.
 *
 * Every entry point takes raw device pointers + explicit shapes + a cudaStream_t (as void*),
 * allocates nothing, returns 0 on success and a negative code on failure (message through
 * tbg_last_error(), thread-local), and never throws across the boundary.  These are the
 * conventions of the reference's one native plugin, the TF op `UpFirDn2D`
 * (models/custom_stylegan2/layers/upfirdn/upfirdn_2d.cu:232-324: shape checks -> Status,
 * output owned by the framework, launch on the framework's stream), extended to the library
 * calls TensorFlow made on the reference's behalf (cuDNN convs, cuBLAS GEMMs, Eigen
 * element-wise kernels, ResourceApplyAdam) — see SURVEY.md §2.1 and INTEGRATION.md.
 *
 * Layout conventions: activations are NHWC bf16 ([B, H, W, C], C contiguous); weights handed to
 * the GEMM kernels are bf16 "K-major" matrices [N, K] with K = (tap_h, tap_w, Cin) flattened;
 * parameters, gradients of parameters, modulation vectors and optimiser state are fp32.
 */
#ifndef TBG_H_
#define TBG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TBG_OK 0
#define TBG_ERR_INVALID_ARG (-1)
#define TBG_ERR_CUDA (-2)
#define TBG_ERR_UNSUPPORTED (-3)

/* Last error message of the calling thread ("" if none). */
const char* tbg_last_error(void);
/* Library/ABI version (bumped when a signature changes). */
int tbg_version(void);
/* Number of kernel launches issued through this library by the calling process (bench.py's
 * gpu_launches counter). tbg_reset_launch_count() zeroes it. */
long long tbg_launch_count(void);
void tbg_reset_launch_count(void);
/* CRC-32C (Castagnoli) of a HOST buffer, continuing from `crc` (0 to start): the tensor checksum of TensorFlow
 * checkpoints (tf.train.Checkpoint files of train.py:94-108, read by textboxgan_b200/tf_checkpoint.py).  No device work. */
unsigned int tbg_crc32c(const void* data, unsigned long long n, unsigned int crc);

/* ------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *
 * Replaces: tf.nn.conv2d / conv2d_transpose as called from
 *   modulated_conv2d.py:99-112 (ModulatedConv2D.call, non-fused algebra :95-96,119-121),
 *   upfirdn_2d_v2.py:65-103 (upsample_conv_2d), :106-113 (conv_downsample_2d),
 *   conv.py:51-73 (Conv2D.call), and their cuDNN backward passes (dgrad is the same kernel
 *   on re-laid-out weights).
 *
 * GEMM view: rows = output pixels (128-pixel boxes of one or several images), columns = n_total
 * output channels, K = taps_h*taps_w*Cin.  up_h/up_w treat columns as (phase_y, phase_x, cout) and
 * scatter phase (py,px) of GEMM row (b,i,j) to output pixel (2i+py, 2j+px): the fused
 * transposed-conv + FIR of upsample_conv_2d, and the input-gradient of every stride-2
 * convolution (DESIGN.md §"geometry algebra").
 *
 * Epilogue, in reference order (modulated_conv2d.py:119-121, noise.py:21, bias_act.py:25-34,
 * discriminator.py:82): v = acc*col_scale[b,c]; v += noise[b,y,x]*noise_strength[0];
 * v += bias[c]; v = act(v)*act_gain; v = (v + residual)*res_scale (if residual; res_first
 * moves the residual merge in front of the activation, the ResNet-unit order of the OCR head).
 * ------------------------------------------------------------------------------------------ */
typedef struct tbg_conv_args {
  const void* x;   /* bf16 [B, H, W, Cin] */
  const void* w;   /* bf16 [n_total, taps_h*taps_w*Cin] */
  void* out;       /* bf16 (or fp32 if out_fp32) [B, out_H, out_W, cout] */
  int B, H, W, Cin;
  int Ho, Wo;      /* GEMM row grid (output pixels; along an up axis: the input grid) */
  int n_total;     /* GEMM columns: cout * (1+up_h) * (1+up_w) */
  int cout;        /* channels of the output tensor */
  int taps_h, taps_w;
  int pad_h, pad_w;       /* input coord = o*stride - pad + tap */
  int stride_h, stride_w; /* 1 or 2 */
  int up_h, up_w;         /* 0 | 1 per axis: 2-phase transposed conv along that axis */
  const float* col_scale;      /* [B, cout] or NULL */
  const float* bias;           /* [cout] or NULL */
  const float* noise;          /* [B, out_H, out_W] or NULL */
  const float* noise_strength; /* device scalar, required if noise */
  const void* residual;        /* bf16, same shape as out, or NULL */
  float res_scale;
  int res_first;   /* 0: v = (act(v)*gain + residual)*res_scale; 1: v = act((v + residual)*res_scale)*gain */
  int act;         /* 0 linear, 1 leaky-relu(0.2), 2 relu */
  float act_gain;  /* multiplies after act (sqrt(2) for lrelu) */
  int out_fp32;    /* 0: bf16 output, 1: fp32 output */
  const void* relu_mask;  /* bf16, same shape as out, or NULL: the final value is zeroed where relu_mask <= 0 — the
                             backward of a ReLU whose output is relu_mask, fused into the input-gradient conv */
  unsigned long long tap_mask[4]; /* per output phase (py*2+px along up axes, else [0]): bit (th*taps_w+tw) set = the tap's
                           weight block is non-zero and is computed; 0 = all taps.  Lets a transposed
                           stride-2 convolution skip the taps a phase does not have. */
} tbg_conv_args;

int tbg_conv2d_igemm(const tbg_conv_args* args, void* stream);

/* Weight gradient of the same convolution (split over pixel blocks, fp32 atomics):
 *   gw[n, (th,tw), c] += sum_{b,ho,wo} gy[b, ho, wo, n] * x[b, ho*s-pad+th, wo*s-pad+tw, c]
 * Replaces cuDNN's backward-filter pass behind tape.gradient (training_step.py:224-235).
 * gw is fp32 [n_total, taps_h*taps_w*Cin] and must be zeroed (or hold a running sum) by the
 * caller.  Along an up axis gy is the 2x-resolution tensor and n indexes (py,px,cout). */
typedef struct tbg_wgrad_args {
  const void* x;   /* bf16 [B, H, W, Cin] */
  const void* gy;  /* bf16 [B, gy_H, gy_W, cout] */
  float* gw;       /* fp32 [n_total, taps_h*taps_w*Cin] */
  int B, H, W, Cin;
  int Ho, Wo;
  int n_total, cout;
  int taps_h, taps_w, pad_h, pad_w, stride_h, stride_w;
  int up_h, up_w;
} tbg_wgrad_args;

int tbg_conv2d_wgrad(const tbg_wgrad_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * upfirdn2d — drop-in for the reference's TF custom op
 *   UpFirDn2D(x: T[major,inH,inW,minor], k: T[kH,kW]; upx,upy,downx,downy,padx0,padx1,pady0,pady1)
 *   -> y: T[major,outH,outW,minor]          (upfirdn_2d.cu:310-324, Compute :232-307)
 * pad (crop if negative) -> zero-insert upsample -> correlate with the flipped FIR -> decimate,
 * fp32 accumulation.  dtype_bf16 = 0: T = float, 1: T = bf16; k is always fp32.
 * outW = (inW*upx + padx0 + padx1 - kW + downx) / downx (same for H).  Errors mirror the op's
 * OP_REQUIRES checks and are returned as TBG_ERR_INVALID_ARG.
 * ------------------------------------------------------------------------------------------ */
int tbg_upfirdn2d(const void* x, const float* k, void* y, int dtype_bf16, int major, int inH, int inW, int minor,
                  int kH, int kW, int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1,
                  void* stream);

/* tf.keras.optimizers.Adam (optimizer_v2) update of a flat fp32 buffer — replaces the per-variable
 * ResourceApplyAdam ops behind optimizer.apply_gradients (training_step.py:235):
 *   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g^2;  p -= lr_t*m/(sqrt(v)+eps),
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t) supplied by the host, either by value or — when lr_t_dev is
 * non-NULL — through a device scalar (so that a captured CUDA graph can be replayed with the
 * step-dependent bias correction updated outside the graph). */
int tbg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr_t, const float* lr_t_dev,
                  float beta1, float beta2, float eps, void* stream);

/* dst = src + (dst - src)*beta over a flat fp32 buffer — Generator.set_as_moving_average_of
 * (generator.py:48-59). */
int tbg_ema_step(float* dst, const float* src, long long n, float beta, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-sequence LSTM recurrence of the frozen OCR head (BiLSTM encoder of ASTER, reached through
 * AsterInferer.call, aster_inferer.py:28-37).  xp = x @ W_ih + b for all steps is computed by the
 * caller; gate order i, f, g, o; H must be 256.
 *   fwd:  xp f32 [D,B,T,4H], w_packed bf16 [D,H(k),H(j),4] -> h f32 [D,B,T,H],
 *         gates f32 [D,B,T,4H] (post-activation), c f32 [D,B,T,H]
 *   bwd:  g_h f32 [D,B,T,H] (+ saved gates, c), wT_packed bf16 [D,H(j),H(k),4] -> g_xp f32 [D,B,T,4H]
 * (input gradient only: the head is frozen, training_step.py:201-206).
 * ------------------------------------------------------------------------------------------ */
int tbg_lstm_seq_fwd(const float* xp, const void* w_packed, float* h_out, float* gates_out, float* c_out, int D, int B,
                     int T, int H, void* stream);
int tbg_lstm_seq_bwd(const float* g_h, const float* gates, const float* c_saved, const void* wT_packed, float* g_xp,
                     int D, int B, int T, int H, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused HBM-bound passes of the plain (first-order) step.  x/xs/g* activations are bf16
 * [B, HW, C] (NHWC flattened), C % 8 == 0; per-sample vectors are fp32 [B, C].
 *   tbg_modulate       xs = x * s[b,c]                                   (modulated_conv2d.py:96)
 *   tbg_modulate_bwd   gx = gxs * s ;  gs[b,c] += sum_hw gxs*x           (gs zeroed, or initialised by
 *                      tbg_demod_bwd with the demodulation term of dL/ds)
 *   tbg_bias_act_bwd   backward of out = act(y0*d + noise*ns + bias)*gain (+ residual)
 *                      (modulated_conv2d.py:121, noise.py:21, bias_act.py:25-34, discriminator.py:82):
 *                      gy0 = g_pre*d (bf16) and, when S1 != NULL, the zero-initialised sums
 *                      S1 = sum_hw g_pre, Spre = sum_hw g_pre*pre, Snz = sum_hw g_pre*noise, from
 *                      which d(bias), d(noise strength) and d(d) follow on [B, C] tensors
 *                      (Spre / Snz may be NULL).  s1_over_batch != 0: S1 is [C], summed over the
 *                      batch as well (= the bias gradient of Conv2D + BiasAct; Spre, noise NULL).
 *                      act: 0 linear, 1 leaky-relu(0.2), 2 relu.
 *   tbg_torgb_fwd/bwd  y[p,j] = sum_c x[p,c]*ws[b,c,j] (+ bias[j]), j < 3   (to_rgb.py:28-33);
 *                      bwd: gx = gy . ws^T (bf16), gws[b,c,j] += sum_p x*gy (gws must be zeroed)
 * ------------------------------------------------------------------------------------------ */
int tbg_modulate(const void* x, const float* s, void* xs, int B, int HW, int C, void* stream);
int tbg_modulate_bwd(const void* gxs, const void* x, const float* s, void* gx, float* gs, int B, int HW, int C,
                     void* stream);
int tbg_bias_act_bwd(const void* g_out, const void* out, const void* residual, const float* noise, const float* d,
                     void* gy0, float* S1, float* Spre, float* Snz, int B, int HW, int C, int act, float gain,
                     int s1_over_batch, void* stream);
/* tbg_bias_act_bwd (all three sums) for a modulated layer whose output also feeds a ToRGB (synthesis_block.py:143-152):
 * the gradient reaching `out` is g_out (NULL on the last block) + g_rgb (x) ws, formed here from the image gradient
 * g_rgb fp32 [B,HW,3] and ws fp32 [B,C,3] instead of being materialised; the same pass yields the ToRGB weight gradient
 *   gws[b,c,j] += sum_p out[b,p,c] * g_rgb[b,p,j]        (gws, S1, Spre, Snz zeroed by the caller; d is required). */
int tbg_bias_act_rgb_bwd(const void* g_out, const void* out, const float* noise, const float* d, const float* g_rgb,
                         const float* ws, void* gy0, float* S1, float* Spre, float* Snz, float* gws, int B, int HW, int C,
                         int act, float gain, void* stream);
/* 4x4 separable FIR k = [1,3,3,1] (x) [1,3,3,1] on NHWC bf16 — the resample kernel upfirdn_2d applies after
 * the transposed convolution of upsample_conv_2d and before the strided convolution of
 * conv_downsample_2d (upfirdn_2d_v2.py:65-113):
 *   out[b,y,x,c] = scale * sum_{m,n<4} k[m] k[n] in[b, y+m+offy, x+n+offx, c]      (in = 0 out of bounds)
 * optionally followed by v = act(v*d[b,c] + noise[b,y,x]*ns + bias[c]) * gain (d, noise, bias may be NULL;
 * act: 0 linear, 1 leaky-relu(0.2)).  k is symmetric: the adjoint is the same call with
 * off' = -3 - off and the roles of in/out exchanged. */
int tbg_fir4(const void* in, void* out, int B, int IH, int IW, int OH, int OW, int C, int offy, int offx, float scale,
             const float* d, const float* noise, const float* noise_strength, const float* bias, int act, float gain,
             void* stream);
/* The skip branch of a residual block (discriminator.py:126-131: conv_downsample_2d with a 1x1 kernel,
 * upfirdn_2d_v2.py:106-113) needs the filtered tensor only at the pixels its stride-(sy, 2) convolution reads:
 *   out[b,p,q,c] = scale * sum_{m,n<4} k[m] k[n] in[b, sy*p+m+offy, 2*q+n+offx, c]      (in = 0 out of bounds; sy = 1 | 2)
 * in [B,IH,IW,C] bf16 -> out [B,OH,OW,C] bf16.  tbg_fir4_down_adjoint is its transpose, optionally added to the
 * gradient that arrives through the main branch (add may be NULL):
 *   out[b,y,x,c] = add[b,y,x,c] + scale * sum k[m] k[n] g[b,p,q,c]  over sy*p+m+offy = y, 2*q+n+offx = x
 * g [B,OH,OW,C], add / out [B,IH,IW,C]. */
int tbg_fir4_down(const void* in, void* out, int B, int IH, int IW, int OH, int OW, int C, int sy, int offy, int offx,
                  float scale, void* stream);
int tbg_fir4_down_adjoint(const void* g, const void* add, void* out, int B, int IH, int IW, int OH, int OW, int C, int sy,
                          int offy, int offx, float scale, void* stream);
/* Tuning switches of the library (explicit calls; the library never reads environment variables).  Keys:
 *   "conv_halo" (default 1)   tbg_conv2d_igemm runs 3x3 stride-1 pad-1 convolutions whose grid is a multiple of 16 x 16
 *                             pixels (Cin % 64 == 0, cout % 32 == 0, no residual / relu_mask / fp32 output) on the
 *                             halo-reuse kernel of csrc/conv_halo.cu: same arguments, same results;
 *   "wgrad_staged" (0), "wgrad_items_per_sm" (0 = heuristic), "lstm_cluster" (1);
 *   "wgrad_halo" (1)          tbg_conv2d_wgrad runs 3x3 stride-1 pad-1 weight gradients on large grids on the halo-reuse
 *                             kernel of csrc/conv_wgrad_halo.cu;
 *   "halo_a_stages" (2..3), "halo_b_stages" (2..8), "halo_staged" (0|1): pipeline depth / store path of conv_halo.
 *   "halo_cta2" (0)           conv3x3_halo on CTA pairs (thread-block clusters of 2, tcgen05.mma.cta_group::2, M = 256,
 *                             each CTA stages half of every weight box): same results bit for bit; measured slower at
 *                             N = 128 (profiles/r02ag_halo_cta2.txt), hence off.
 * tbg_get_tuning returns the current value or -1 for an unknown key. */
int tbg_set_tuning(const char* key, int value);
int tbg_get_tuning(const char* key);

/* AsterInferer.convert_inputs (aster_inferer.py:153-190): NCHW fp32 image [B,3,H,W] -> per-sample crop at
 * floor(first_blank * cw_num / cw_den) columns (clamped to [1, W]; W when `labels` [B, mcn] has no `blank`) ->
 * bilinear resize (half-pixel centres, no antialias) -> NHWC fp32 [B, oh, ow, 3].  bwd scatters into gimg (zeroed by
 * the caller) with fp32 atomics. */
int tbg_crop_resize_fwd(const float* img, const int* labels, float* out, int B, int H, int W, int oh, int ow, int mcn,
                        int blank, int cw_num, int cw_den, void* stream);
int tbg_crop_resize_bwd(const float* g, const int* labels, float* gimg, int B, int H, int W, int oh, int ow, int mcn,
                        int blank, int cw_num, int cw_den, void* stream);

/* Loader transform of a whole batch on the device (dataset_utils/training_data_loader.py:64-86): sample b = BGR uint8 image
 * src + offsets[b], src_h[b] x src_w[b] x 3 (HWC) -> cv2.resize(INTER_LINEAR) to dst_w[b] x H, rounded to the uint8 grid,
 * / 127.5 - 1, zero-padded on the right to W, written CHW into out fp32 [B,3,H,W].  All arrays are DEVICE pointers. */
int tbg_batch_resize_normalize(const unsigned char* src, const long long* offsets, const int* src_h, const int* src_w,
                               const int* dst_w, float* out, int B, int H, int W, void* stream);

/* FromRGB (from_rgb.py:26-29): 1x1 conv 3 -> C of the NCHW fp32 image + bias + leaky-ReLU(0.2)*gain -> NHWC bf16:
 *   out[b,p,c] = lrelu(coef * sum_j img[b,j,p] w[j,c] + bias[c]) * gain,   w fp32 [3, C]
 * bwd: gpre = g_out*gain*slope(out); gimg[b,j,p] = coef sum_c gpre w[j,c] (NCHW fp32, may be NULL);
 *      gw[j,c] += coef sum_{b,p} img gpre, gb[c] += sum_{b,p} gpre (zeroed by the caller; both NULL to skip). */
int tbg_fromrgb_fwd(const float* img, const float* w, const float* bias, void* out, int B, int HW, int C, float coef,
                    float gain, void* stream);
int tbg_fromrgb_bwd(const float* img, const float* w, const void* g_out, const void* out, float* gimg, float* gw, float* gb,
                    int B, int HW, int C, float coef, float gain, void* stream);
/* Stand-alone nodes of the twice-differentiable path (path-length / R1 regularisers, training_step.py:300-373):
 *   tbg_bias_act_fwd: out = act(t + noise[b,p]*noise_strength + bias[c]) * gain      (noise.py:21, bias_act.py:25-34)
 *   tbg_rowdot:       out[b,c] += sum_p a[b,p,c] * b[b,p,c]   (fp32 [B,C], zeroed by the caller; the adjoint of the
 *                     style modulation x * s[b,c], modulated_conv2d.py:96) */
int tbg_bias_act_fwd(const void* t, const float* noise, const float* noise_strength, const float* bias, void* out, int B,
                     int HW, int C, int act, float gain, void* stream);
int tbg_rowdot(const void* a, const void* b, float* out, int B, int HW, int C, void* stream);
int tbg_torgb_fwd(const void* x, const float* ws, const float* bias, float* y, int B, int HW, int C, void* stream);
int tbg_torgb_bwd(const void* x, const float* ws, const float* gy, void* gx, float* gws, int B, int HW, int C,
                  void* stream);

/* ------------------------------------------------------------------------------------------
 * Weight preparation / gradient fold.  From the fp32 HWIO master weight w[KH,KW,I,O] (KH,KW <= 3)
 * build, in one launch, the bf16 GEMM matrix of a convolution geometry, the matrix of its adjoint
 * geometry and (optionally) q[i,o] = coef^2 * sum_taps w^2 (demodulation, modulated_conv2d.py:80-82):
 *   fwd[(p,q,o),(t,u,i)] = coef * sum_{kh,kw} Ty[p,t,kh] Tx[q,u,kw] w[kh,kw,i,o]     (rows Opad, cols Ipad)
 *   adj[(p,q,i),(t,u,o)] = coef * sum_{kh,kw} Ty'[p,t,kh] Tx'[q,u,kw] w[kh,kw,i,o]
 * `tables` is a HOST array of 4 blocks {P, T, K, 36 floats [P][T][K]} for Ty, Tx, Ty', Tx' (identity
 * for plain convs; FIR-folded tables for upsample_conv_2d / conv_downsample_2d,
 * upfirdn_2d_v2.py:65-113).  coef is the equalised-LR runtime coefficient (commons.py:4-12).
 * tbg_wfold is the transpose: gw += coef * fold(gfwd) (+ 2 coef^2 w gq), gfwd fp32 in fwd layout;
 * gq[i,o] = dL/dq is either given or formed in the kernel from (s [nb,I], t [nb,O]) of tbg_demod_bwd as
 * sum_b s[b,i]^2 t[b,o] (pass gq NULL).  accumulate != 0: gw += ...; 0: gw = ... (no zero-initialisation needed).
 * ------------------------------------------------------------------------------------------ */
int tbg_wprep(const float* w, const float* tables, float coef, int KH, int KW, int I, int O, int Ipad, int Opad,
              void* fwd, void* adj, float* q, void* stream);
/* Every weight of an iteration in one launch (the per-weight launches are latency-bound): tbg_wprep_make_job writes the
 * description of ONE tbg_wprep call (same arguments, device pointers) into a HOST buffer of tbg_wprep_job_bytes() bytes and
 * returns the number of CTAs it needs (> 0) or a negative status; block_begin is the sum of the counts of the jobs before
 * it.  The caller copies the packed jobs to device memory (16-byte aligned) once and replays
 * tbg_wprep_group(jobs_dev, n_jobs, total_blocks) every iteration; results equal n_jobs tbg_wprep calls. */
int tbg_wprep_job_bytes(void);
int tbg_wprep_make_job(void* job_host, int block_begin, const float* w, const float* tables, float coef, int KH, int KW,
                       int I, int O, int Ipad, int Opad, void* fwd, void* adj, float* q);
int tbg_wprep_group(const void* jobs_dev, int n_jobs, int total_blocks, void* stream);
int tbg_wfold(const float* gfwd, const float* gq, const float* w, const float* tables, float coef, int KH, int KW,
              int I, int O, int Ipad, int Opad, float* gw, const float* s, const float* t, int nb, int accumulate,
              void* stream);

/* Fold of a gradient held in the adjoint-matrix layout [Ipad, (tap, Opad)] of an identity-table geometry
 * (role-swapped weight gradient of a transposed convolution): gw[tap,i,o] += coef*gadj[i, tap'*Opad+o], tap' =
 * tap, or the spatially mirrored tap when flip != 0 (upsample_conv_2d flips w, upfirdn_2d_v2.py:80)
 * (+ 2 coef^2 w dL/dq with dL/dq formed from (s, t) as in tbg_wfold). */
int tbg_wfold_adj(const float* gadj, const float* w, float coef, int KH, int KW, int I, int O, int Opad, float* gw,
                  const float* s, const float* t, int nb, int flip, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------
 * Demodulation coefficient of ModulatedConv2D and its gradient (modulated_conv2d.py:75-82), fp32:
 *   tbg_demod_coef  d[b,o] = rsqrt(sum_i s[b,i]^2 q[i,o] + eps)           s [B,I], q [I,O] (tbg_wprep)
 *   tbg_demod_bwd   from S1/Spre/Snz [B,O] of tbg_bias_act_bwd, d, the noise strength ns (device
 *                   scalar) and bias [O]:
 *                     t[b,o]   = dL/d(s^2 @ q) = -0.5 (Spre - ns Snz - bias S1) d^2
 *                     gs[b,i]  = 2 s[b,i] sum_o t[b,o] q[i,o]   (written; tbg_modulate_bwd adds to it)
 *                     gbias[o] = sum_b S1[b,o],  gns[0] = sum_{b,o} Snz[b,o]
 * ------------------------------------------------------------------------------------------ */
int tbg_demod_coef(const float* s, const float* q, float* d, int B, int I, int O, float eps, void* stream);
int tbg_demod_bwd(const float* S1, const float* Spre, const float* Snz, const float* d, const float* ns,
                  const float* bias, const float* s, const float* q, float* t, float* gbias, float* gns, float* gs,
                  int B, int I, int O, void* stream);

/* ------------------------------------------------------------------------------------------
 * Grouped style projection of the synthesis network: for every modulated convolution l,
 *   s_l[b,i] = coef * sum_k style[b, idx_l, k] * w_l[k,i] + b_l[i] + 1        (modulated_conv2d.py:75-76,
 * dense.py:23-29; coef = 1/sqrt(S)), all layers in ONE launch; tbg_style_dense_bwd returns, in two
 * launches, gw_l = coef * style[:,idx_l]^T gs_l, gb_l = sum_b gs_l and
 * gstyle[b,j,:] = coef * sum_{l: idx_l = j} gs_l[b,:] w_l^T (rows no layer uses are zeroed).
 * style / gstyle: fp32 [B, n_style, S]; at most 32 layers; the table is a HOST array.
 * ------------------------------------------------------------------------------------------ */
typedef struct tbg_style_layer {
  const float* w;   /* [S, I] */
  const float* b;   /* [I] */
  float* s;         /* fwd out [B, I] */
  const float* gs;  /* bwd in  [B, I] */
  float* gw;        /* bwd out [S, I] */
  float* gb;        /* bwd out [I] */
  int I, idx;
} tbg_style_layer;
int tbg_style_dense_fwd(const tbg_style_layer* layers, int n_layers, const float* style, int B, int n_style, int S,
                        float coef, void* stream);
int tbg_style_dense_bwd(const tbg_style_layer* layers, int n_layers, const float* style, float* gstyle, int B,
                        int n_style, int S, float coef, void* stream);

/* ------------------------------------------------------------------------------------------
 * Greedy Bahdanau-attention LSTM decoder of the frozen OCR head, all decode steps in one launch,
 * one CTA per sample (reached through AsterInferer.call, aster_inferer.py:28-37; the head is frozen,
 * so backward yields only d/d mem and d/d keys).  Dimensions are fixed: hidden = attention units =
 * embedding = 256, memory features 512, 96 classes, T <= 64.  Weights bf16 (v, b, bd fp32).
 *   fwd: mem f32 [B,T,512], keys f32 [B,T,256] (= mem Wm, computed by the caller)
 *        -> logits f32 [B,steps,96] + saved a [B,steps,T], ctx [B,steps,512], gates [B,steps,1024],
 *           c, h [B,steps,256], prev int32 [B,steps]
 *   bwd: g_logits f32 [B,steps,96] (+ saved) -> g_mem f32 [B,T,512], g_keys f32 [B,T,256]
 * ------------------------------------------------------------------------------------------ */
typedef struct tbg_dec_weights {
  const void *wq, *wqT;   /* bf16 [256,256] query layer and its transpose */
  const void* v;          /* f32  [256] attention_v */
  const void* emb;        /* bf16 [96,256] previous-symbol embedding */
  const void *wg, *wgT;   /* bf16 [1024,1024]: rows [emb | ctx | h] = LSTM W_ih stacked on W_hh (-> i,f,g,o), and its transpose */
  const void* b;          /* f32  [1024] */
  const void *wd, *wdT;   /* bf16 [768,96], [96,768]  output dense on [h, ctx] */
  const void* bd;         /* f32  [96] */
} tbg_dec_weights;

int tbg_attn_decoder_fwd(const float* mem, const float* keys, const tbg_dec_weights* w, float* logits, float* sv_a,
                         float* sv_ctx, float* sv_gates, float* sv_c, float* sv_h, int* sv_prev, int B, int T,
                         int steps, void* stream);
int tbg_attn_decoder_bwd(const float* mem, const float* keys, const tbg_dec_weights* w, const float* g_logits,
                         const float* sv_a, const float* sv_ctx, const float* sv_gates, const float* sv_c,
                         const float* sv_h, float* g_mem, float* g_keys, int B, int T, int steps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small fp32 dense layers (exact fp32 FMA; the reference computes them in fp32).
 *
 * Replaces the cuBLAS GEMMs TensorFlow ran for Dense.call (dense.py:23-29), the mapping network
 * (mapping_block.py:15-45), the word encoder's Keras Dense (word_encoder.py:21,52) and the discriminator head
 * (discriminator.py:132-142, 213).  Row-major fp32, x [M,K], w [K,N]:
 *   fwd: y = act((x @ w) * coef + bias * bias_coef) * gain          act: 0 linear, 1 leaky-relu(0.2), 2 relu
 *   bwd: gpre = gy * gain * act'(y) (workspace [M,N], written);  gx = coef * gpre @ w^T (or NULL);
 *        gw (+)= coef * x^T @ gpre (or NULL; accumulate_gw adds into gw);  gb = bias_coef * sum_m gpre (or NULL)
 * tbg_pixel_norm_*: y = x * rsqrt(mean_k x^2 + 1e-8) per row (mapping_block.py:15-18) and its gradient.
 * ------------------------------------------------------------------------------------------ */
int tbg_dense_fwd(const float* x, const float* w, const float* bias, float* y, int M, int K, int N, float coef,
                  float bias_coef, int act, float gain, void* stream);
int tbg_dense_bwd(const float* x, const float* w, const float* y, const float* gy, float* gpre, float* gx, float* gw,
                  float* gb, int M, int K, int N, float coef, float bias_coef, int act, float gain, int accumulate_gw,
                  void* stream);
int tbg_pixel_norm_fwd(const float* x, float* y, int M, int K, void* stream);
int tbg_pixel_norm_bwd(const float* x, const float* gy, float* gx, int M, int K, void* stream);

/* WordEncoder.call (word_encoder.py:39-63) in one launch: embedding lookup in concat(w0 [1,E] frozen, table [V-1,E])
 * (words int32 [B,mcn] in [0, V)), dropout (mask fp32 [B,mcn,E] of 0/1 or NULL, kept values scaled by 1/keep), Keras
 * Dense(E -> D) + ReLU, then reshape [B, out_w, out_c, out_h] / transpose to the base feature map, written NHWC bf16
 * [B, out_h, out_w, out_c] (mcn*D == out_h*out_w*out_c).  emb [B*mcn,E] and act [B*mcn,D] (fp32) are saved for bwd.
 * bwd: g_out NHWC bf16 -> gpre workspace [B*mcn,D]; g_table [V-1,E] accumulated with atomics (caller zeroes it; the
 * frozen row w0 gets none); g_fc_w [E,D] and g_fc_b [D] written. */
int tbg_word_encoder_fwd(const int* words, const float* w0, const float* table, const float* mask, float keep,
                         const float* fc_w, const float* fc_b, float* emb, float* act, void* out, int B, int mcn, int E,
                         int D, int out_h, int out_w, int out_c, void* stream);
int tbg_word_encoder_bwd(const int* words, const float* mask, float keep, const float* fc_w, const float* emb,
                         const float* act, const void* g_out, float* gpre, float* g_table, float* g_fc_w, float* g_fc_b,
                         int B, int mcn, int E, int D, int out_h, int out_w, int out_c, void* stream);

/* MinibatchStd.call (mini_batch_std.py:10-35) on NHWC bf16 x [n_calls*B, HW, C]: per call of B samples, groups of
 * G = min(group_size, B) samples {m, m + B/G, ...}; statistic = mean over all HW*C features of sqrt(var_group + 1e-8).
 * fwd writes xcat [n_calls*B, HW, Cpad] = [x | statistic | zeros] (the extra channel padded up to the GEMM K block) and
 * stat [n_calls*B]; bwd: gx = gxcat[..., :C] + d statistic / d x * sum_{group, pixels} gxcat[..., C]. */
int tbg_minibatch_std_fwd(const void* x, void* xcat, float* stat, int B, int n_calls, int group_size, int HW, int C,
                          int Cpad, void* stream);
int tbg_minibatch_std_bwd(const void* x, const void* gxcat, void* gx, int B, int n_calls, int group_size, int HW, int C,
                          int Cpad, void* stream);

/* R1 penalty reduction (training_step.py:363-372): out[b] = sum_{c,h,w} g[b]^2 on the fp32 image gradient
 * [B, per_sample]; bwd: gg = 2 * g * gout[b]. */
int tbg_r1_sqnorm(const float* g, float* out, int B, long long per_sample, void* stream);
int tbg_r1_sqnorm_bwd(const float* g, const float* gout, float* gg, int B, long long per_sample, void* stream);

/* ToRGB.call (to_rgb.py:28-33) fused with the skip sum of SynthesisBlock.call (synthesis_block.py:152) and, for the last
 * block, mask_text_box (utils/utils.py:11-45) + the NHWC -> NCHW layout change:
 *   out[b,p,:] = x[b,p,:] @ ws[b] + bias + upsample_2d(y_prev)[b,p,:]         (y_prev fp32 NHWC [B,H/2,W/2,3] or NULL)
 *   masked (words int32 [B,mcn] or NULL): column w of sample b is zero unless words[b, floor(w*mcn/W)] != 0
 *   out fp32 NHWC [B,H,W,3], or NCHW [B,3,H,W] when nchw != 0.  C in {64,128,192,256,512}.
 * tbg_image_grad_nhwc is the adjoint of the mask + layout change: g fp32 NCHW -> masked NHWC. */
int tbg_torgb_skip_fwd(const void* x, const float* ws, const float* bias, const float* y_prev, const int* words,
                       float* out, int B, int H, int W, int C, int mcn, int nchw, void* stream);
int tbg_image_grad_nhwc(const float* g, const int* words, float* out, int B, int H, int W, int mcn, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TBG_H_ */
