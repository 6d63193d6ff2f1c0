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
} hm_operand;

/* destination of an engine epilogue: element (n, y, x, ch) of the logical result is written to
 * ptr[((n*H + y + h_off)*W + x + w_off)*C + c_off + ch].  ptr == NULL disables that output. */
typedef struct hm_out_f32 { float* ptr; int H, W, C, h_off, w_off, c_off; } hm_out_f32;
typedef struct hm_out_bf16 { void* hi; void* lo; int H, W, C, h_off, w_off, c_off; } hm_out_bf16;

const char* hm_version(void);
int hm_last_cuda_error(void); /* cudaError_t of the last failed runtime call on this thread */

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
int hm_wgrad_unpack(const float* G_ws, int KH, int KW, int cp, int cq, float* dst, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HM_B200_H_ */
