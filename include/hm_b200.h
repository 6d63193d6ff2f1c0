/* hm_b200.h -- C ABI of libhm_b200.so, the B200 (sm_100a) compute library behind the mask2image
 * (pix2pixHD-style layout->image GAN) training hot path of xcyan/neurips18_hierchical_image_manipulation.
 *
 * The reference has no FFI of its own: every op below replaces a stock torch.nn call made by the
 * reference's Python modules (file:line cited per entry, relative to the reference tree).  A maintainer
 * binds these with ctypes; the host side that mirrors the reference's Python
 * interface lives in neurips18_hierchical_image_manipulation_b200/.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *  - the library never allocates, never synchronises and never throws: workspaces are passed in,
 *    work is enqueued on `stream` (a cudaStream_t passed as void*), the return value is 0 or a
 *    negative hm_status;
 *  - activations are NHWC.  fp32 tensors are dense [N,H,W,C].  bf16 operand tensors are
 *    [N,H,W,Cs] with Cs (channel stride) a multiple of 8 and C <= Cs valid channels; a bf16 operand
 *    is a (hi, lo) pair of planes: lo == NULL selects plain bf16 products, lo != NULL selects the
 *    3-product split (hi*hi + lo*hi + hi*lo, fp32 accumulate) that reproduces fp32 convolution to
 *    ~1e-5 relative -- the parity mode;
 *  - weights/grads cross the boundary in the reference's own layouts (OIHW fp32 for Conv2d,
 *    IOHW for ConvTranspose2d); hm_pack_weight produces the bf16 [tap][rows_pad][k_pad] slabs the
 *    engines consume.
 */
#ifndef HM_B200_H_
#define HM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum hm_status {
  HM_OK = 0,
  HM_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  HM_ERR_TENSORMAP = -2, /* cuTensorMapEncodeTiled failed */
  HM_ERR_LAUNCH = -3,    /* kernel launch failed */
  HM_ERR_DRIVER = -4     /* CUDA driver entry point unavailable */
} hm_status;

typedef enum hm_act { HM_ACT_NONE = 0, HM_ACT_RELU = 1, HM_ACT_LRELU = 2, HM_ACT_TANH = 3 } hm_act;

/* bf16 NHWC operand (read through TMA). dims are those of the STORED tensor, i.e. they include any
 * border the producer materialised (reflection padding). */
typedef struct hm_operand {
  const void* hi;
  const void* lo; /* NULL => single bf16 product */
  int n, h, w, c; /* c = valid channels */
  int cs;         /* channel stride in elements, multiple of 8 */
  int lo_c0;      /* channels [0, lo_c0) are exactly representable in bf16 (their lo plane is all zero, e.g. the one-hot
                   * label map of pix2pixHD_condImg_model.py:151-152): the engines skip the lo product there.  0 = no
                   * such guarantee.  Zero padding channels (>= c) are skipped at 16-channel granularity likewise. */
} hm_operand;

/* destination of an engine epilogue: element (n, y, x, ch) of the logical result is written to
 * ptr[((n*H + y + h_off)*W + x + w_off)*C + c_off + ch].  ptr == NULL disables that output. */
typedef struct hm_out_f32 { float* ptr; int H, W, C, h_off, w_off, c_off; } hm_out_f32;
typedef struct hm_out_bf16 { void* hi; void* lo; int H, W, C, h_off, w_off, c_off; } hm_out_bf16;

const char* hm_version(void);
int hm_last_cuda_error(void); /* cudaError_t of the last failed runtime call on this thread */

/* Optional device scratch for the engines (the library never allocates): hm_scratch_bytes() bytes, 256 B aligned,
 * owned by the caller, ZERO-FILLED before registration (its first 4 KB hold self-resetting arrival counters) and kept
 * alive until it is replaced (ptr == NULL unregisters).  With it (and hm_set_streamk(1)) the CTA-pair K-engine
 * balances the last, partly filled wave of output tiles by splitting their contraction range over all SMs
 * (deterministic two-phase stream-K: partial accumulators go to the scratch and are summed in a fixed order);
 * without it the engines fall back to whole tiles.  One scratch per process: launches that use it must be ordered on
 * one stream. */
size_t hm_scratch_bytes(void);
int hm_set_scratch(void* ptr, size_t bytes);
/* The stream-K tail is OPT-IN (default off, or HM_STREAMK=1 in the environment): measured on B200 it does not beat the
 * whole-tile schedule on this path's shapes.  hm_set_streamk(1/0) switches it at run time. */
int hm_set_streamk(int on);
/* Persistent engines launch one CTA (or CTA pair) per SM.  While a collective runs concurrently (the bucketed gradient
 * allreduce of the data-parallel step, whose NCCL CTAs need SMs of their own), a full-width grid would leave some of its
 * CTAs waiting for the SMs NCCL holds and double the kernel's time; hm_set_sm_limit(n) makes the engines size their
 * grids for n SMs (rounded down to even) until it is reset with 0.  Host-side setting, read at launch (or capture) time. */
int hm_set_sm_limit(int n);

/* ---- weight packing --------------------------------------------------------------------------
 * N-tile width the K-engine uses for `rows` output rows, and the padded slab dims. */
int hm_pick_bn(int rows);
int hm_rows_pad(int rows);   /* round_up(rows, hm_pick_bn(rows)) */
int hm_k_pad(int k);         /* round_up(k, 64) */
/* src is fp32 with element (row r, contraction index k, tap t) at src[r*s_row + k*s_k + t*s_tap];
 * dst_{hi,lo} are bf16 [taps][rows_pad][k_pad], zero padded.  dst_lo may be NULL.
 *   Conv2d  W[co][ci][kh][kw] (models/Pix2Pix_NET.py:74-91) fprop : rows=co, k=ci   (s_row=ci*kh*kw, s_k=kh*kw, s_tap=1)
 *                                                            dgrad : rows=ci, k=co
 *   ConvTranspose2d W[ci][co][kh][kw] (Pix2Pix_NET.py:89)    fwd   : rows=co, k=ci ; dgrad: rows=ci, k=co */
int hm_pack_weight(const float* src, int rows, int k, int taps, long s_row, long s_k, long s_tap, void* dst_hi,
                   void* dst_lo, void* stream);
/* Both roles of one weight tensor in one pass over the fp32 source W[A][B][taps] (Conv2d: A = co, B = ci;
 * ConvTranspose2d: A = ci, B = co): p1 = slabs with rows = A, k = B; p2 = slabs with rows = B, k = A (each
 * [taps][hm_rows_pad(rows)][hm_k_pad(k)], lo planes nullable).  Same values as two hm_pack_weight calls. */
int hm_pack_weight_pair(const float* src, int A, int B, int taps, void* p1_hi, void* p1_lo, void* p2_hi, void* p2_lo,
                        void* stream);

/* ---- tcgen05 implicit-GEMM convolution engines ----------------------------------------------------
 * hm_conv_fprop: y[n,ho,wo,co] = act(bias[co] + sum_{kh,kw,ci} x[n, ho*stride+kh-pad, wo*stride+kw-pad, ci] * Wp[kh,kw][co][ci])
 *   x is the stored operand; coordinates outside it read as zero (TMA fill), so `pad` is the ZERO padding
 *   and reflection padding is a border materialised by the producer (then pad = 0).
 *   Replaces nn.Conv2d forward (Pix2Pix_NET.py:74,78,91; layer_util.py:350,367; Discriminator_NET.py:72-93;
 *   torchvision VGG19 convs via layer_util.py:384-399) and ConvTranspose2d backward-data.
 * hm_conv_dgrad: the adjoint of hm_conv_fprop w.r.t. x (dy plays "x", Wp packed with rows=ci,k=co):
 *   dx[n,h,w,ci] = act(bias + sum_{kh,kw,co : (h+pad-kh) % stride == 0 ...} dy[n,(h+pad-kh)/stride,(w+pad-kw)/stride,co] * Wp[kh,kw][ci][co])
 *   for h < Hout, w < Wout.  Replaces Conv2d backward-data and ConvTranspose2d FORWARD (Pix2Pix_NET.py:89,
 *   stride 2, pad 1, output_padding 1: Hout = 2*H).
 * stride is 1 or 2.  out32 / out16 may each be NULL.  err_flag: optional device int set non-zero
 * if a pipeline wait timed out (kernel bug guard); 0 otherwise. */
int hm_conv_fprop(const hm_operand* x, const void* w_hi, const void* w_lo, int k_pad, int rows_pad,
                  const float* bias, int KH, int KW, int stride, int pad, int Hout, int Wout, int Cout, int act,
                  float slope, const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, void* stream);
int hm_conv_dgrad(const hm_operand* dy, const void* w_hi, const void* w_lo, int k_pad, int rows_pad,
                  const float* bias, int KH, int KW, int stride, int pad, int Hout, int Wout, int Cout, int act,
                  float slope, const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, void* stream);

/* hm_conv_wgrad: G[(kh,kw, cp)][cq] = sum_{n,y,x} P[n, y*stride+kh-pad, x*stride+kw-pad, cp] * Q[n,y,x,cq]
 *   Conv2d: P = stored input x, Q = dy  -> dW[co=cq][ci=cp][kh][kw];  ConvTranspose2d: P = dy, Q = x -> dW[ci=cq][co=cp][kh][kw].
 *   G (fp32 workspace, hm_wgrad_ws_bytes) is then scattered by hm_wgrad_unpack into the reference layout
 *   dst[cq][cp][kh][kw] (accumulate != 0 adds to dst: D is run on real and fake). Replaces Conv2d /
 *   ConvTranspose2d backward-weight. */
size_t hm_wgrad_ws_bytes(int KH, int KW, int cp, int cq);
int hm_conv_wgrad(const hm_operand* P, const hm_operand* Q, int KH, int KW, int stride, int pad, float* G_ws,
                  int* err_flag, void* stream);
/* Dilated variants (nn.Conv2d(..., dilation=d, padding=d): DilatedResnetBlock, models/layer_util.py:255-293, the
 * --add_dilated_layers latent blocks of MaskTwoStreamConvSwitch_NET.py:101-104): tap (kh, kw) reads x[h + kh*d - pad,
 * w + kw*d - pad]; dilation == 1 is hm_conv_fprop / hm_conv_dgrad / hm_conv_wgrad exactly.  dilation > 1 needs stride 1 for
 * the gradients and runs on the generic / CTA-pair engines (the row-streaming engines need adjacent horizontal taps). */
int hm_conv_fprop_dil(const hm_operand* x, const void* w_hi, const void* w_lo, int k_pad, int rows_pad,
                      const float* bias, int KH, int KW, int stride, int pad, int dilation, int Hout, int Wout, int Cout,
                      int act, float slope, const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, void* stream);
int hm_conv_dgrad_dil(const hm_operand* dy, const void* w_hi, const void* w_lo, int k_pad, int rows_pad,
                      const float* bias, int KH, int KW, int stride, int pad, int dilation, int Hout, int Wout, int Cout,
                      int act, float slope, const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, void* stream);
int hm_conv_wgrad_dil(const hm_operand* P, const hm_operand* Q, int KH, int KW, int stride, int pad, int dilation, float* G_ws,
                      int* err_flag, void* stream);
int hm_wgrad_unpack(const float* G_ws, int KH, int KW, int cp, int cq, float* dst, int accumulate, void* stream);

/* ---- tap-unrolled lowering of thin (<= 4 channel) convolutions (hm_thin.cu) ---------------------------------
 * The engines contract over 64-channel chunks per filter tap and emit >= 16-column N tiles, so a 3-channel side
 * wastes most of every MMA.  Folding the horizontal taps of the thin side into its channel index gives the same
 * arithmetic with KW x fewer MMAs.  Used for the generator head Conv2d(ngf,3,7) (models/Pix2Pix_NET.py:91), the
 * PatchGAN output Conv2d(512,1,4) (models/Discriminator_NET.py:93) and VGG19 conv1_1 (layer_util.py:384-390).
 *
 * hm_pack_weight_ex: hm_pack_weight with two-level row / contraction indices and an optional DEVICE scalar
 *   multiplier (*scale, e.g. 1/sigma of spectral normalisation, models/sn_utils.py:49-72):
 *   element (r, k, t) = *scale * src[(r / r_div)*s_r_hi + (r % r_div)*s_r_lo + (k / k_div)*s_k_hi + (k % k_div)*s_k_lo + t*s_tap]. */
int hm_pack_weight_ex(const float* src, int rows, int r_div, long s_r_hi, long s_r_lo, int k, int k_div, long s_k_hi,
                      long s_k_lo, int taps, long s_tap, const float* scale, void* dst_hi, void* dst_lo, void* stream);
/* G[(kh*round_up(cp,64) + p)][kw*cq + q] (ld = round_up(KW*cq,64), the hm_conv_wgrad workspace of P = x, Q = unrolled dy
 * with KW' = 1) -> dst[q][p][kh][kw] (OIHW), optionally accumulating. */
int hm_wgrad_unpack_cols(const float* G_ws, int KH, int KW, int cp, int cq, float* dst, int accumulate, void* stream);
/* dst[n,h,w,(j*KW+i)*C + c] = src[n, h+oh+sh*j, w+ow+sw*i, c] (zero outside src; channels >= KH*KW*C of dst zero).
 * src: operand [N,Hs,Ws,s_cs], dst: operand [N,Hd,Wd,d_cs]; s_lo / d_lo may be NULL. */
int hm_tap_unroll(const void* s_hi, const void* s_lo, int N, int Hs, int Ws, int C, int s_cs, int KH, int KW, int oh,
                  int ow, int sh, int sw, void* d_hi, void* d_lo, int Hd, int Wd, int d_cs, void* stream);
/* out[n,h,w,c] = act(bias[c] + sum_{j<KH,i<KW} T[n, h+oh+sh*j, w+ow+sw*i, (j*KW+i)*C + c]), terms outside T skipped;
 * T fp32 [N,Ht,Wt,ldT], out fp32 [N,Ho,Wo,ldo], C <= 4, bias may be NULL. */
int hm_tap_combine(const float* T, int N, int Ht, int Wt, int ldT, int KH, int KW, int C, int oh, int ow, int sh, int sw,
                   const float* bias, int act, float slope, float* out, int Ho, int Wo, int ldo, void* stream);

/* ---- HBM-bound kernels (hm_elementwise.cu) --------------------------------------------------------------
 * fp32 tensors are dense NHWC; "operand" outputs are bf16 (hi, lo) planes with channel stride *_cs (multiple of 8),
 * lo may be NULL. */

/* K11. Pix2PixHDModel_condImg.encode_input + get_edges (models/pix2pixHD_condImg_model.py:144-174, 285-291).
 * label / inst / mask_in are fp32 [B,1,H,W], image fp32 [B,3,H,W] (NCHW, as the reference's data loader emits them,
 * already on the device); inst == NULL means --no_instance.  Writes
 *   g  : generator input operand [B, H+2*g_border, W+2*g_border, g_cs], channels one-hot | edge | (1-mask)*image,
 *        ReflectionPad2d(g_border) materialised (Pix2Pix_NET.py:74);
 *   d  : (optional) discriminator input operand [2B,H,W,d_cs]: conditioning channels in both halves, the real
 *        image in channels [cin, cin+3) of the second half (the fake half is filled by hm_finish_fake);
 *   v  : (optional) VGG input operand [2B,H,W,v_cs]: real image in the second half.
 *   d_no_imgcond bit 0 (--no_imgCond, :213-214): the D operand is [label | edge | image] without the masked image;
 *   bit 1 (netG global_twostream with which_encoder 'ctx', :71-72,178-179,227-228): the D operand is the bare [image];
 *   d_mask != NULL (--mask_gan_input, :180-181; mask_in, or mask_out with --use_soft_mask, :217): fp32 [B,1,H,W]
 *   multiplied into every channel of the D operand (here, in hm_finish_fake and, for the gradient, in hm_fake_bwd). */
int hm_encode_input(const float* label, const float* inst, const float* image, const float* mask_in, int B, int H,
                    int W, int label_nc, void* g_hi, void* g_lo, int g_cs, int g_border, void* d_hi, void* d_lo,
                    int d_cs, void* v_hi, void* v_lo, int v_cs, int d_no_imgcond, const float* d_mask, void* stream);

/* K7. nn.InstanceNorm2d(C, affine=False) (models/layer_util.py:19-26), split in statistics / apply / backward.
 * ws: hm_in_ws_bytes(N, H*W, C) bytes of scratch. */
size_t hm_in_ws_bytes(int N, int HW, int C);
int hm_in_stats(const float* y, int N, int HW, int C, float eps, float* ws, float* mean, float* rstd, void* stream);
/* out = act((y-mean)*rstd) [+ skip]; mean == NULL skips the normalisation.  Emits the dense fp32 result (optional)
 * and/or the bf16 operand with a materialised border (reflect != 0: ReflectionPad2d(border), else zeros). */
int hm_in_apply(const float* y, const float* mean, const float* rstd, const float* skip, int N, int H, int W, int C,
                int act, float slope, float* out32, void* o_hi, void* o_lo, int o_cs, int border, int reflect,
                void* stream);
/* Backward of [InstanceNorm] -> act:  dz = fold_reflect(g1) + g2 + l1coef*sign(z - tref);
 * dy = IN'(dz * act'(.)).  g1 is fp32 [N,H+2b,W+2b,g1_ld] (read at channel offset g1_coff), g2 dense [N,H,W,C];
 * act' is taken from y (with stats) if given, else z, else the sign of the bf16 plane mask_hi.  Result: operand
 * [N,H,W,o_cs] and/or dense fp32. */
int hm_in_bwd(const float* y, const float* mean, const float* rstd, const float* z, const void* mask_hi, int mask_cs,
              const float* g1, int g1_border, int g1_ld, int g1_coff, const float* g2, const float* tref, float l1coef,
              int N, int H, int W, int C, int act, float slope, float* ws, void* o_hi, void* o_lo, int o_cs,
              float* out32, void* stream);
/* out = base + fold_reflect(g_padded)  (residual skip gradient, layer_util.py:376-378); base may be NULL */
int hm_fold_add(const float* g_padded, int border, int N, int H, int W, int C, const float* base, float* out,
                void* stream);

/* K8. nn.AvgPool2d(3, stride=2, padding=1, count_include_pad=False) (Discriminator_NET.py:31-32, Pix2Pix_NET.py:45)
 * on operands (H, W = interior extent; the input may carry a materialised border in_border, the output is written
 * with a ReflectionPad2d(out_border) border), and its adjoint accumulated into channels [c0,c1) of a finer fp32
 * gradient. */
int hm_avgpool3s2(const void* i_hi, const void* i_lo, int N, int H, int W, int cs, int in_border, void* o_hi, void* o_lo,
                  int out_border, void* stream);
int hm_avgpool3s2_bwd(const float* g_coarse, int N, int Ho, int Wo, int ld_coarse, float* g_fine, int H, int W,
                      int ld_fine, int c0, int c1, void* stream);
/* VGG19 MaxPool2d(2,2) (torchvision features, layer_util.py:384-399) and its adjoint (first maximum wins). */
int hm_maxpool2(const void* i_hi, const void* i_lo, int N, int H, int W, int cs, void* o_hi, void* o_lo, void* stream);
int hm_maxpool2_bwd(const float* g, int N, int H, int W, int C, const void* a_hi, const void* a_lo, int cs, float* dz,
                    void* stream);

/* K9. loss reductions into fp64 accumulators: *acc += coef*sum|a-b| (nn.L1Loss, losses.py:75-82 and
 * pix2pixHD_condImg_model.py:235-251) and *acc += coef*sum (a-target)^2 (LSGAN nn.MSELoss, losses.py:40-50);
 * hm_mse_grad writes scale*(y-target) as an operand (the caller folds 2*coef/numel into scale). */
int hm_l1_sum(const float* a, const float* b, long n, double coef, double* acc, void* stream);
int hm_mse_sum(const float* a, long n, float target, double coef, double* acc, void* stream);
int hm_mse_grad(const float* y, long P, int C, float target, float scale, void* o_hi, void* o_lo, int o_cs, void* stream);
/* Vanilla GAN (--no_lsgan: GANLoss with nn.BCELoss, models/losses.py:17-20) on the discriminator's raw last-layer output:
 * hm_bce_sum: *acc += coef * sum BCE(sigmoid(x), target) (logs clamped at -100 like nn.BCELoss);
 * hm_bce_grad: operand = scale * d/dx BCE(sigmoid(x), target), same conventions as hm_mse_grad. */
int hm_bce_sum(const float* x, long n, float target, double coef, double* acc, void* stream);
int hm_bce_grad(const float* x, long P, int C, float target, float scale, void* o_hi, void* o_lo, int o_cs, void* stream);

/* K12. generator head epilogue: output gate (Pix2Pix_NET.py:96-99), NCHW fp32 copy for the caller, fake image into
 * channels [d_coff, d_coff+3) of the first half of the D operand and into the first half of the VGG operand;
 * hm_fake_bwd is its adjoint: d/d(pre-tanh) of everything that reads the fake image. */
int hm_finish_fake(const float* t, const float* image, const float* mask, int use_gate, int B, int H, int W,
                   float* fake_nchw, void* d_hi, void* d_lo, int d_cs, int d_coff, void* v_hi, void* v_lo, int v_cs,
                   const float* d_mask, void* stream);
int hm_fake_bwd(const float* t, const float* mask, int use_gate, const float* gD, int gD_ld, int gD_coff,
                const float* gV, int gV_ld, const float* real_nchw, float rec_coef, int B, int H, int W, void* o_hi,
                void* o_lo, int o_cs, const float* d_mask, void* stream);

/* misc: dense fp32 [P][ld] (channels [coff, coff+C)) * scale -> operand; per-channel sums (bias gradients). */
int hm_f32_to_operand(const float* x, long P, int C, int ld, int coff, float scale, void* o_hi, void* o_lo, int o_cs,
                      void* stream);
int hm_colsum(const float* x, long P, int C, float* out, int accumulate, void* stream);
int hm_colsum_operand(const void* hi, const void* lo, long P, int C, int cs, float* out, int accumulate, void* stream);

/* ---- GlobalTwoStreamGenerator glue (hm_twostream.cu; models/Pix2Pix_NET.py:103-247) ---------------------------
 * hm_mask_maxpool: nn.MaxPool2d(f, f) of the object mask (:133): mask fp32 [B,1,H,W] -> out fp32 [B,H/f,W/f].
 * hm_mask_blend: (1-m)*a + m*b ('early_add' fusion of the context and label streams, :207-209 and
 *   FeatureFusionBlock 'add', layer_util.py:322-323); a, b fp32 [N,H,W,C], m fp32 [N,H,W]; either of a / b may be
 *   NULL (single-stream variants: plain copy).  Emits the dense fp32 result and/or the operand with
 *   ReflectionPad2d(border) materialised (input of the first ResnetBlock).
 * hm_mask_blend_bwd: da = (1-m)*g, db = m*g.
 * hm_concat_operands: torch.cat((enc, dec), 1) of two operands over P pixels (:221, skip connections).
 * hm_cond_image_operand: (1 - mask)*image (pix2pixHD_condImg_model.py:165-166) as the context-stream input operand
 *   [B,H+2b,W+2b,o_cs] with ReflectionPad2d(b); image fp32 NCHW [B,3,H,W], mask fp32 [B,1,H,W]. */
int hm_mask_maxpool(const float* mask, int B, int H, int W, int f, float* out, void* stream);
int hm_mask_blend(const float* a, const float* b, const float* m, int N, int H, int W, int C, float* out32, void* o_hi,
                  void* o_lo, int o_cs, int border, void* stream);
int hm_mask_blend_bwd(const float* g, const float* m, long P, int C, float* da, float* db, void* stream);
/* hm_pool_exchange: ImagePool.query (util/image_pool.py:11-31, --pool_size > 0; discriminate(..., use_pool=True),
 * pix2pixHD_condImg_model.py:182-184) for image b of a batch of bf16 operands with elems_per_image elements each (% 8 == 0).
 * The decisions are drawn on the host and read from DEVICE memory (int32 dec[2b] = action, dec[2b+1] = pool slot):
 * 0 out = cur; 1 pool[slot] = cur, out = cur (pool filling); 2 out = pool[slot], pool[slot] = cur (exchange).  Images of
 * one batch must be processed in order (one launch per image). */
int hm_pool_exchange(const void* cur_hi, const void* cur_lo, void* pool_hi, void* pool_lo, void* out_hi, void* out_lo,
                     const int* dec, int b, long elems_per_image, void* stream);
/* FeatureFusionBlock 'concat' (layer_util.py:305-327; feat_fusion '*_concat' of GlobalTwoStreamGenerator):
 * hm_mask_concat: operand [P pixels, o_cs >= 2C] = relu(cat((1 - m) * a, m * b)) from dense fp32 a, b [P, C] and m [P];
 * hm_mask_concat_bwd: g [P, ld >= 2C] = gradient w.r.t. that operand -> da, db [P, C] (through the ReLU and the mask). */
int hm_mask_concat(const float* a, const float* b, const float* m, long P, int C, void* o_hi, void* o_lo, int o_cs,
                   void* stream);
int hm_mask_concat_bwd(const float* g, int ld, const float* m, const float* a, const float* b, long P, int C, float* da,
                       float* db, void* stream);
int hm_concat_operands(const void* a_hi, const void* a_lo, int a_cs, int Ca, const void* b_hi, const void* b_lo, int b_cs,
                       int Cb, void* o_hi, void* o_lo, int o_cs, long P, void* stream);
int hm_cond_image_operand(const float* image, const float* mask, int B, int H, int W, void* o_hi, void* o_lo, int o_cs,
                          int border, void* stream);

/* K13 (opt-in). Spectral normalisation of the PatchGAN convolutions: models/sn_utils.py:11-25 max_singular_value
 * (Ip = 1) and :49-72 SNConv2d.W_bar = W / sigma, with W viewed as [n = Cout][m = Cin*KH*KW] as stored (OIHW).
 * `layers` is a DEVICE array with one entry per convolution; every layer of the discriminator is handled by one
 * launch.  stash (hm_sn_stash_floats(n, m) floats per layer) receives u0 | b = W v | v | sigma, 1/sigma, |a|, |b|:
 * stash + 2n + m + 1 is the device scalar 1/sigma that hm_pack_weight_ex takes as `scale` (the normalisation is fused
 * into the weight load; W_bar is never materialised).  update_u != 0 stores u' back into u (SNConv2d in training mode).
 * hm_sn_weight_grad turns grad = dL/dW_bar (as accumulated by hm_conv_wgrad + hm_wgrad_unpack) into dL/dW in place,
 * differentiating through the power iteration exactly as autograd does in the reference. */
typedef struct hm_sn_layer { const float* W; float* grad; float* u; float* stash; int n, m; } hm_sn_layer;
size_t hm_sn_stash_floats(int n, int m);
int hm_sn_power_iteration(const hm_sn_layer* layers, int n_layers, int max_n, int max_m, int update_u, void* stream);
int hm_sn_weight_grad(const hm_sn_layer* layers, int n_layers, int max_n, int max_m, void* stream);

/* K10. torch.optim.Adam(lr, betas=(beta1, 0.999)) step (pix2pixHD_condImg_model.py:135,139) over a flat fp32
 * segment; grad_scale multiplies the gradient first (1/world_size after a sum-allreduce). step is 1-based. */
int hm_adam_step(float* param, const float* grad, float* m, float* v, long n, float lr, float beta1, float beta2,
                 float eps, int step, float grad_scale, void* stream);
/* Same step with the (1-based) step count read from DEVICE memory, so that a training step captured in a CUDA graph
 * replays with the right bias corrections. */
int hm_adam_step_dev(float* param, const float* grad, float* m, float* v, long n, float lr, float beta1, float beta2,
                     float eps, const int* step_dev, float grad_scale, void* stream);

/* ---- box2mask generator glue (hm_box2mask.cu; BASELINE config #5) ------------------------------------
 * hm_box2mask_encode: the generator input of TwoStreamAE_mask (models/TwoStreamAE_mask.py:127-151,331-338, cond_in ==
 *   'ctx_obj'): operand [B,H,W,o_cs] = cat(object box mask in the object's class channel, one-hot(context labels)).
 *   mask_ctx_in / mask_in are [B,1,H,W] fp32, cls is [B] fp32 class ids.  All values are 0/1 (exact in bf16).
 * hm_bn_fold: nn.BatchNorm2d(affine) in training mode (layer_util.py:19-21): batch statistics mean/rstd [C] (from
 *   hm_in_stats on the tensor viewed as ONE sample of N*H*W pixels) + gamma/beta -> per-(n,c) rows for hm_in_apply:
 *   rstd' = rstd*gamma, mean' = mean - beta/rstd'.  With running_mean / running_var (both or neither) it also performs the
 *   module's buffer update: running = (1-momentum)*running + momentum*statistic (variance unbiased by count/(count-1),
 *   recovered from rstd and eps), `repeat` times, and *num_batches_tracked += repeat (nullable).
 * hm_upsample2_add: out = deep + nn.Upsample(scale_factor=2, mode='bilinear')(small), the DeconvResnetBlock tail
 *   (layer_util.py:178-179,236-242); fp32 NHWC, small is [N,h,w,C], deep/out are [N,2h,2w,C], C % 4 == 0.
 * hm_box2mask_head: MaskTwoStreamConv_NET.forward :190-217 -- obj_prob = sigmoid(obj_logit), comb = (1-p)*ctx + p*obj,
 *   log-softmax over the C classes (outputs in NCHW, any may be NULL) -- and the reconstruction losses of
 *   TwoStreamAE_mask.forward :188-203: acc[0] += sum of NLL over pixels with mask_out >= 0.5 (mask_losses.py:12-27),
 *   acc[1] += their count, acc[2] += sum of BCE(obj_prob [* mask_out when gated], inst) terms (logs clamped at -100).
 *   `use_gate` is a bit set: bit 0 = --use_output_gate; bit 1 = --no_comb (MaskTwoStreamConvSwitch_NET.forward :208: the
 *   context logits are returned as they are, comb = ctx); bit 2 = --objReconLoss l1 (acc[2] accumulates |q - inst|
 *   instead of the BCE terms, TwoStreamAE_mask.py:50-51); in hm_box2mask_head_bwd too. */
int hm_box2mask_encode(const float* mask_ctx_in, const float* mask_in, const float* cls, int B, int H, int W, int label_nc,
                       void* o_hi, void* o_lo, int o_cs, void* stream);
int hm_bn_fold(const float* mean, const float* rstd, const float* gamma, const float* beta, int N, int C, float* mean_out,
               float* rstd_out, float* running_mean, float* running_var, long long* num_batches_tracked, float count,
               float momentum, float eps, int repeat, void* stream);
/* hm_bn_stats = hm_in_stats over the batch-folded tensor (one sample of N*HW pixels) + hm_bn_fold in two launches instead of
 * three: y fp32 [N, HW, C]; mean / rstd [C] (the batch statistics hm_bn_bwd needs), mean_rows / rstd_rows [N][C] for
 * hm_in_apply, optional running-buffer update (both or neither; unbiased variance, momentum, `repeat` evaluations).
 * ws: hm_in_ws_bytes(1, N*HW, C). */
int hm_bn_stats(const float* y, int N, int HW, int C, float eps, float* ws, float* mean, float* rstd, const float* gamma,
                const float* beta, float* mean_rows, float* rstd_rows, float* running_mean, float* running_var,
                long long* num_batches_tracked, float momentum, int repeat, void* stream);
int hm_upsample2_add(const float* small, const float* deep, int N, int h, int w, int C, float* out, void* stream);
/* Backward halves (training step of TwoStreamAE_mask.forward :233-248 without the GAN terms):
 * hm_bn_bwd: BatchNorm2d(affine) + activation backward, batch statistics (mean / rstd [C] over N*H*W): the gradient w.r.t.
 *   the BN OUTPUT is g1 [+ g2] (dense fp32 [N,H,W,C]); the activation mask is the sign of gamma*yhat+beta (or of z /
 *   mask_hi when given); result = gradient w.r.t. the conv output y as a bf16 operand and / or dense fp32; dgamma / dbeta
 *   are ACCUMULATED.  ws: hm_in_ws_bytes(1, N*H*W, C).
 * hm_upsample2_bwd: adjoint of the bilinear x2 upsample of hm_upsample2_add (d_deep is g itself).
 * hm_box2mask_head_bwd: gradient of  w_obj*loss_obj + w_comb*loss_comb  w.r.t. the context logits (operand [N,H,W,c_cs])
 *   and the object logit (operand [N,H,W,o_cs], channel 0); acc is the accumulator hm_box2mask_head filled (acc[1] = box
 *   pixel count). */
int hm_bn_bwd(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* z,
              const void* mask_hi, int mask_cs, const float* g1, const float* g2, int N, int H, int W, int C, int act,
              float slope, float* ws, void* o_hi, void* o_lo, int o_cs, float* out32, float* dgamma, float* dbeta,
              void* stream);
int hm_upsample2_bwd(const float* g, int N, int h, int w, int C, float* dsmall, void* stream);
/* g_prob (nullable, fp32 with leading dimension g_ld): gradient w.r.t. channel 0 of the discriminator input of the
 * --use_gan branch, added to d/d obj_prob (times mask^2 when use_gate). */
int hm_box2mask_head_bwd(const float* ctx_logit, const float* obj_logit, int obj_ld, const float* label_map,
                         const float* mask_out, const float* inst, int N, int H, int W, int C, int use_gate,
                         const double* acc, float w_comb, float w_obj, const float* g_prob, int g_ld, void* c_hi,
                         void* c_lo, int c_cs, void* o_hi, void* o_lo, int o_cs, void* stream);
/* hm_box2mask_d_input: discriminator input of --use_gan (TwoStreamAE_mask.py:153-157,205-213): operand [B,H,W,o_cs] =
 *   cat(x * mask^x_mask_power, cond * mask) with cond as in hm_box2mask_encode; mask_out == NULL: no gating. */
int hm_box2mask_d_input(const float* x, const float* mask_ctx_in, const float* mask_in, const float* cls,
                        const float* mask_out, int x_mask_power, int B, int H, int W, int label_nc, void* o_hi, void* o_lo,
                        int o_cs, void* stream);
int hm_box2mask_head(const float* ctx_logit, const float* obj_logit, int obj_ld, const float* label_map,
                     const float* mask_out, const float* inst, int N, int H, int W, int C, int use_gate, float* comb_logit,
                     float* comb_logprob, float* obj_prob, double* acc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HM_B200_H_ */
