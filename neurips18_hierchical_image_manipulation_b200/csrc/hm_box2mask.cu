// hm_box2mask.cu -- HBM-bound glue of the box2mask generator (BASELINE config #5; models/MaskTwoStreamConv_NET.py,
// models/TwoStreamAE_mask.py, models/layer_util.py:119-242): input encoding, BatchNorm folding, the bilinear x2 upsample
// of the DeconvResnetBlock shortcut fused with the residual add, and the two-stream output head with its losses.
// Convolutions run on the tcgen05 engines (hm_conv.cu); normalise + activation + operand emission reuse hm_in_apply
// with the folded statistics.
#include "../../include/hm_b200.h"
#include "hm_ptx.cuh"

#include <algorithm>

namespace {

using bf16 = __nv_bfloat16;
constexpr int kBlock = 256;
inline int grid_for(long items, int block = kBlock, int max_blocks = 148 * 32) {
  long g = (items + block - 1) / block;
  return int(std::max<long>(1, std::min<long>(g, max_blocks)));
}
#define HM_LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? HM_OK : HM_ERR_LAUNCH)

// cond = cat(object box mask placed in the object's class channel, one-hot(context label map))
// (TwoStreamAE_mask.encode_input :127-151 + construct_input_cond :331-338, cond_in == 'ctx_obj').  Every value is 0 or 1,
// hence exact in bf16: only the hi plane is written (the lo plane, if any, is zeroed).
__global__ void b2m_encode_kernel(const float* __restrict__ mask_ctx_in, const float* __restrict__ mask_in,
                                  const float* __restrict__ cls, int B, int H, int W, int label_nc, bf16* o_hi, bf16* o_lo,
                                  int cs) {
  const int G = cs >> 3;
  const long total = long(B) * H * W * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % G);
    const long pix = i / G;
    const int n = int(pix / (long(H) * W));
    const int c0 = g * 8;
    const int ctx_cls = int(__ldg(mask_ctx_in + pix));
    const int obj_cls = int(__ldg(cls + n));
    const float box = __ldg(mask_in + pix);
    alignas(16) bf16 h[8];
    alignas(16) bf16 z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float v = 0.f;
      if (c < label_nc) v = (c == obj_cls) ? box : 0.f;
      else if (c < 2 * label_nc) v = (c - label_nc == ctx_cls) ? 1.f : 0.f;
      h[j] = __float2bfloat16_rn(v);
      z[j] = __float2bfloat16_rn(v - __bfloat162float(h[j]));
    }
    *reinterpret_cast<uint4*>(o_hi + pix * cs + c0) = *reinterpret_cast<const uint4*>(h);
    if (o_lo) *reinterpret_cast<uint4*>(o_lo + pix * cs + c0) = *reinterpret_cast<const uint4*>(z);
  }
}

// Discriminator input of the --use_gan branch (TwoStreamAE_mask.py:153-157, 205-213): cat(x, cond) with x the object mask
// (instance mask or generated probability) and cond the generator's own conditioning tensor, everything multiplied by
// the box mask when use_output_gate is set (the generated mask twice: it is gated once for the loss, once more here).
__global__ void b2m_d_input_kernel(const float* __restrict__ x, const float* __restrict__ mask_ctx_in,
                                   const float* __restrict__ mask_in, const float* __restrict__ cls,
                                   const float* __restrict__ mask_out, int x_mask_power, int B, int H, int W, int label_nc,
                                   bf16* o_hi, bf16* o_lo, int cs) {
  const int G = cs >> 3;
  const long total = long(B) * H * W * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % G);
    const long pix = i / G;
    const int n = int(pix / (long(H) * W));
    const int c0 = g * 8;
    const int ctx_cls = int(__ldg(mask_ctx_in + pix));
    const int obj_cls = int(__ldg(cls + n));
    const float box = __ldg(mask_in + pix);
    const float m = mask_out ? __ldg(mask_out + pix) : 1.f;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float t = 0.f;
      if (c == 0) { t = __ldg(x + pix) * m; if (x_mask_power > 1) t *= m; }
      else if (c <= label_nc) t = ((c - 1 == obj_cls) ? box : 0.f) * m;
      else if (c <= 2 * label_nc) t = ((c - 1 - label_nc == ctx_cls) ? 1.f : 0.f) * m;
      v[j] = t;
    }
    alignas(16) bf16 h[8];
    alignas(16) bf16 l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hm::split_bf16(v[j], h[j], l[j]);
    *reinterpret_cast<uint4*>(o_hi + pix * cs + c0) = *reinterpret_cast<const uint4*>(h);
    if (o_lo) *reinterpret_cast<uint4*>(o_lo + pix * cs + c0) = *reinterpret_cast<const uint4*>(l);
  }
}

// BatchNorm2d(affine) in training mode = normalise with the BATCH statistics (mean / rstd over N*H*W, computed by
// hm_in_stats on the tensor viewed as one sample) then scale and shift:  (x - mean) * rstd * gamma + beta
//   == (x - mean') * rstd'   with  rstd' = rstd * gamma,  mean' = mean - beta / rstd'   -> per-(n, c) rows for hm_in_apply.
// With running_mean / running_var the kernel also performs BatchNorm's buffer update (momentum 0.1 in the reference's
// nn.BatchNorm2d defaults): running = (1 - m) * running + m * batch statistic, the variance UNBIASED (x count/(count-1)),
// `repeat` times (the discriminator evaluates the generated batch twice per iteration), and num_batches_tracked += repeat.
__global__ void bn_fold_kernel(const float* __restrict__ mean, const float* __restrict__ rstd,
                               const float* __restrict__ gamma, const float* __restrict__ beta, int N, int C,
                               float* __restrict__ mean_out, float* __restrict__ rstd_out, float* running_mean,
                               float* running_var, long long* num_batches_tracked, float count, float momentum, float eps,
                               int repeat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int c = i % C;
  const float rs = rstd[c] * (gamma ? gamma[c] : 1.f);
  const float b = beta ? beta[c] : 0.f;
  rstd_out[i] = rs;
  mean_out[i] = (rs != 0.f) ? mean[c] - b / rs : mean[c];
  if (i < C && running_mean && running_var) {
    const float r = rstd[c];
    const float var_b = fmaxf(1.f / (r * r) - eps, 0.f);
    const float var_u = count > 1.f ? var_b * (count / (count - 1.f)) : var_b;
    float rm = running_mean[c], rv = running_var[c];
    for (int k = 0; k < repeat; ++k) {
      rm = (1.f - momentum) * rm + momentum * mean[c];
      rv = (1.f - momentum) * rv + momentum * var_u;
    }
    running_mean[c] = rm;
    running_var[c] = rv;
    if (i == 0 && num_batches_tracked) *num_batches_tracked += repeat;
  }
}

// out = deep + Upsample(scale 2, bilinear, align_corners = False)(small): the tail of DeconvResnetBlock.forward
// (layer_util.py:236-242, shortcut = [...] + nn.Upsample, :178-179).  One thread per output pixel x 4 channels.
__global__ void upsample2_add_kernel(const float* __restrict__ small, const float* __restrict__ deep, int N, int h, int w,
                                     int C, float* __restrict__ out) {
  const int H = 2 * h, W = 2 * w, G = C >> 2;
  const long total = long(N) * H * W * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % G);
    long r = i / G;
    const int x = int(r % W); r /= W;
    const int y = int(r % H);
    const int n = int(r / H);
    // source coordinate 0.5 * (dst + 0.5) - 0.5, clamped at 0 (ATen area_pixel_compute_source_index)
    const float sy = fmaxf(0.5f * (y + 0.5f) - 0.5f, 0.f), sx = fmaxf(0.5f * (x + 0.5f) - 0.5f, 0.f);
    const int y0 = int(sy), x0 = int(sx);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float fy = sy - y0, fx = sx - x0;
    const float* base = small + size_t(n) * h * w * C + g * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(base + (size_t(y0) * w + x0) * C));
    const float4 b = __ldg(reinterpret_cast<const float4*>(base + (size_t(y0) * w + x1) * C));
    const float4 c = __ldg(reinterpret_cast<const float4*>(base + (size_t(y1) * w + x0) * C));
    const float4 d = __ldg(reinterpret_cast<const float4*>(base + (size_t(y1) * w + x1) * C));
    const size_t o = ((size_t(n) * H + y) * W + x) * C + g * 4;
    const float4 e = __ldg(reinterpret_cast<const float4*>(deep + o));
    const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
    float4 v;
    v.x = e.x + w00 * a.x + w01 * b.x + w10 * c.x + w11 * d.x;
    v.y = e.y + w00 * a.y + w01 * b.y + w10 * c.y + w11 * d.y;
    v.z = e.z + w00 * a.z + w01 * b.z + w10 * c.z + w11 * d.z;
    v.w = e.w + w00 * a.w + w01 * b.w + w10 * c.w + w11 * d.w;
    *reinterpret_cast<float4*>(out + o) = v;
  }
}

// Two-stream output head (MaskTwoStreamConv_NET.forward :190-217) + the reconstruction losses of
// TwoStreamAE_mask.forward :188-203 (MaskReconLoss = NLL over the pixels inside the box, mask_losses.py:12-27;
// nn.BCELoss on the gated object mask).  One thread per pixel; ctx logits are read twice (second pass from L1/L2).
//   acc[0] += sum of -log p(label) over box pixels, acc[1] += number of box pixels, acc[2] += sum of BCE terms
__global__ void b2m_head_kernel(const float* __restrict__ ctx_logit /*[N,H,W,C]*/, const float* __restrict__ obj_logit /*[N,H,W,ldo]*/,
                                int ldo, const float* __restrict__ label_map, const float* __restrict__ mask_out,
                                const float* __restrict__ inst, int N, int H, int W, int C, int flags,
                                float* __restrict__ comb_logit /*NCHW*/, float* __restrict__ comb_logprob /*NCHW*/,
                                float* __restrict__ obj_prob /*[N,1,H,W]*/, double* __restrict__ acc) {
  const bool use_gate = flags & 1, no_comb = flags & 2;   // bit 0: --use_output_gate, bit 1: --no_comb
  const long HW = long(H) * W, total = long(N) * HW;
  double nll = 0.0, cnt = 0.0, bce = 0.0;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int n = int(i / HW);
    const long p = i - long(n) * HW;
    const float o = __ldg(obj_logit + i * ldo);
    const float pr = 1.f / (1.f + __expf(-o));
    const float* cl = ctx_logit + i * C;
    float mx = -3.402823466e38f;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, no_comb ? __ldg(cl + c) : (1.f - pr) * __ldg(cl + c) + pr * o);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf((no_comb ? __ldg(cl + c) : (1.f - pr) * __ldg(cl + c) + pr * o) - mx);
    const float lse = mx + logf(se);
    const float mo = mask_out ? __ldg(mask_out + i) : 1.f;
    const int lab = label_map ? int(__ldg(label_map + i)) : -1;
    for (int c = 0; c < C; ++c) {
      const float v = no_comb ? __ldg(cl + c) : (1.f - pr) * __ldg(cl + c) + pr * o;
      const size_t oi = (size_t(n) * C + c) * HW + p;
      if (comb_logit) comb_logit[oi] = v;
      if (comb_logprob) comb_logprob[oi] = v - lse;
      if (c == lab && mo >= 0.5f) { nll -= double(v - lse); cnt += 1.0; }
    }
    if (obj_prob) obj_prob[i] = pr;
    if (inst) {
      const float q = use_gate ? pr * mo : pr;
      const float t = __ldg(inst + i);
      if (flags & 4) {                       // --objReconLoss l1 (nn.L1Loss, TwoStreamAE_mask.py:50-51)
        bce += double(fabsf(q - t));
      } else {
        const float lq = fmaxf(logf(q), -100.f), l1q = fmaxf(logf(1.f - q), -100.f);   // nn.BCELoss clamps its logs at -100
        bce -= double(t * lq + (1.f - t) * l1q);
      }
    }
  }
  __shared__ double sh[3][kBlock];
  sh[0][threadIdx.x] = nll; sh[1][threadIdx.x] = cnt; sh[2][threadIdx.x] = bce;
  __syncthreads();
  for (int s = kBlock / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int k = 0; k < 3; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && acc)
    for (int k = 0; k < 3; ++k) atomicAdd(acc + k, sh[k][0]);
}

// Adjoint of the bilinear x2 upsample (align_corners = False): d_small[n, i, j, :] = sum over the output pixels whose
// interpolation footprint contains (i, j) of weight * g.  Per axis, output 2i-1 ... 2i+2 can touch source i; the weights
// are recomputed with the forward's own formula so that the clamped borders come out right.
__device__ __forceinline__ float up2_weight(int dst, int src, int n) {   // weight of source index `src` in output `dst`
  const float s = fmaxf(0.5f * (dst + 0.5f) - 0.5f, 0.f);
  const int i0 = int(s), i1 = min(i0 + 1, n - 1);
  const float f = s - i0;
  return (src == i0 ? 1.f - f : 0.f) + (src == i1 ? f : 0.f);
}
__global__ void upsample2_bwd_kernel(const float* __restrict__ g, int N, int h, int w, int C, float* __restrict__ dsmall) {
  const int H = 2 * h, W = 2 * w, G = C >> 2;
  const long total = long(N) * h * w * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int gq = int(i % G);
    long r = i / G;
    const int x = int(r % w); r /= w;
    const int y = int(r % h);
    const int n = int(r / h);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int Y = max(2 * y - 1, 0); Y <= min(2 * y + 2, H - 1); ++Y) {
      const float wy = up2_weight(Y, y, h);
      if (wy == 0.f) continue;
      for (int X = max(2 * x - 1, 0); X <= min(2 * x + 2, W - 1); ++X) {
        const float wx = up2_weight(X, x, w);
        if (wx == 0.f) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + ((size_t(n) * H + Y) * W + X) * C + gq * 4));
        const float ww = wy * wx;
        acc.x += ww * v.x; acc.y += ww * v.y; acc.z += ww * v.z; acc.w += ww * v.w;
      }
    }
    *reinterpret_cast<float4*>(dsmall + ((size_t(n) * h + y) * w + x) * C + gq * 4) = acc;
  }
}

// Backward of b2m_head_kernel for  L = w_obj * loss_obj + w_comb * loss_comb  (TwoStreamAE_mask.py:233-235 without the GAN
// term):  d/d ctx_logit [N,H,W,C]  and  d/d obj_logit [N,H,W,1]  as bf16 gradient operands.
//   comb_c = (1-p) ctx_c + p o,  lp = log_softmax(comb);  loss_comb = -(1/cnt) sum_{box} lp[label]
//   q = p * mask_out (gate);     loss_obj = -(1/NHW) sum ( t log q + (1-t) log(1-q) )   (logs clamped at -100: zero slope there)
__global__ void b2m_head_bwd_kernel(const float* __restrict__ ctx_logit, const float* __restrict__ obj_logit, int ldo,
                                    const float* __restrict__ label_map, const float* __restrict__ mask_out,
                                    const float* __restrict__ inst, int N, int H, int W, int C, int flags,
                                    const double* __restrict__ acc /* acc[1] = number of box pixels */, float w_comb,
                                    float w_obj, const float* __restrict__ g_prob, int g_ld, bf16* c_hi, bf16* c_lo,
                                    int c_cs, bf16* o_hi, bf16* o_lo, int o_cs) {
  const bool use_gate = flags & 1, no_comb = flags & 2;   // bit 0: --use_output_gate, bit 1: --no_comb
  const long total = long(N) * H * W;
  const float inv_cnt = acc[1] > 0.5 ? float(1.0 / acc[1]) : 0.f;
  const float inv_n = 1.f / float(total);
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const float o = __ldg(obj_logit + i * ldo);
    const float pr = 1.f / (1.f + __expf(-o));
    const float* cl = ctx_logit + i * C;
    const float mo = mask_out ? __ldg(mask_out + i) : 1.f;
    const int lab = int(__ldg(label_map + i));
    const bool in_box = mo >= 0.5f;
    float mx = -3.402823466e38f;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, no_comb ? __ldg(cl + c) : (1.f - pr) * __ldg(cl + c) + pr * o);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf((no_comb ? __ldg(cl + c) : (1.f - pr) * __ldg(cl + c) + pr * o) - mx);
    const float inv_se = 1.f / se;
    float dp = 0.f, dsum = 0.f;
    bf16* ch = c_hi + i * c_cs;
    bf16* clo = c_lo ? c_lo + i * c_cs : nullptr;
    for (int c = 0; c < c_cs; ++c) {
      float dctx = 0.f;
      if (c < C && in_box) {
        const float x = __ldg(cl + c);
        const float sm = expf((no_comb ? x : (1.f - pr) * x + pr * o) - mx) * inv_se;
        const float dcomb = w_comb * inv_cnt * (sm - (c == lab ? 1.f : 0.f));
        dctx = no_comb ? dcomb : (1.f - pr) * dcomb;
        if (!no_comb) { dp += dcomb * (o - x); dsum += dcomb; }
      }
      bf16 hh, ll;
      hm::split_bf16(dctx, hh, ll);
      ch[c] = hh;
      if (clo) clo[c] = ll;
    }
    float d_o = pr * dsum;                       // direct path of the object logit into comb
    {
      const float q = use_gate ? pr * mo : pr;
      const float t = __ldg(inst + i);
      float dq = 0.f;
      if (flags & 4) {                       // --objReconLoss l1: sign(q - t), 0 at equality like torch
        dq = q > t ? 1.f : (q < t ? -1.f : 0.f);
      } else {
        if (logf(q) > -100.f) dq -= t / q;
        if (logf(1.f - q) > -100.f) dq += (1.f - t) / (1.f - q);
      }
      dp += w_obj * inv_n * dq * (use_gate ? mo : 1.f);
    }
    // GAN term (--use_gan): g_prob is the gradient w.r.t. channel 0 of the discriminator input = p * mask^2 (gated)
    if (g_prob) dp += __ldg(g_prob + i * g_ld) * (use_gate ? mo * mo : 1.f);
    d_o += dp * pr * (1.f - pr);
    for (int c = 0; c < o_cs; ++c) {
      bf16 hh, ll;
      hm::split_bf16(c == 0 ? d_o : 0.f, hh, ll);
      o_hi[i * o_cs + c] = hh;
      if (o_lo) o_lo[i * o_cs + c] = ll;
    }
  }
}

}  // namespace

extern "C" {

int hm_upsample2_bwd(const float* g, int N, int h, int w, int C, float* dsmall, void* stream) {
  if (!g || !dsmall || (C & 3)) return HM_ERR_INVALID;
  upsample2_bwd_kernel<<<grid_for(long(N) * h * w * (C >> 2)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(g, N, h, w, C,
                                                                                                               dsmall);
  return HM_LAUNCH_OK();
}

int hm_box2mask_head_bwd(const float* ctx_logit, const float* obj_logit, int obj_ld, const float* label_map,
                         const float* mask_out, const float* inst, int N, int H, int W, int C, int use_gate,
                         const double* acc, float w_comb, float w_obj, const float* g_prob, int g_ld, void* c_hi,
                         void* c_lo, int c_cs, void* o_hi, void* o_lo, int o_cs, void* stream) {
  if (!ctx_logit || !obj_logit || !label_map || !inst || !acc || !c_hi || !o_hi || c_cs < C || o_cs < 1) return HM_ERR_INVALID;
  b2m_head_bwd_kernel<<<grid_for(long(N) * H * W, kBlock, 148 * 8), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      ctx_logit, obj_logit, obj_ld, label_map, mask_out, inst, N, H, W, C, use_gate, acc, w_comb, w_obj, g_prob, g_ld,
      static_cast<bf16*>(c_hi), static_cast<bf16*>(c_lo), c_cs, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs);
  return HM_LAUNCH_OK();
}

int hm_box2mask_d_input(const float* x, const float* mask_ctx_in, const float* mask_in, const float* cls,
                        const float* mask_out, int x_mask_power, int B, int H, int W, int label_nc, void* o_hi, void* o_lo,
                        int o_cs, void* stream) {
  if (!x || !mask_ctx_in || !mask_in || !cls || !o_hi || (o_cs & 7) || o_cs < 1 + 2 * label_nc) return HM_ERR_INVALID;
  b2m_d_input_kernel<<<grid_for(long(B) * H * W * (o_cs >> 3)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      x, mask_ctx_in, mask_in, cls, mask_out, x_mask_power, B, H, W, label_nc, static_cast<bf16*>(o_hi),
      static_cast<bf16*>(o_lo), o_cs);
  return HM_LAUNCH_OK();
}

int hm_box2mask_encode(const float* mask_ctx_in, const float* mask_in, const float* cls, int B, int H, int W, int label_nc,
                       void* o_hi, void* o_lo, int o_cs, void* stream) {
  if (!mask_ctx_in || !mask_in || !cls || !o_hi || (o_cs & 7) || o_cs < 2 * label_nc) return HM_ERR_INVALID;
  b2m_encode_kernel<<<grid_for(long(B) * H * W * (o_cs >> 3)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      mask_ctx_in, mask_in, cls, B, H, W, label_nc, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs);
  return HM_LAUNCH_OK();
}

int hm_bn_fold(const float* mean, const float* rstd, const float* gamma, const float* beta, int N, int C, float* mean_out,
               float* rstd_out, float* running_mean, float* running_var, long long* num_batches_tracked, float count,
               float momentum, float eps, int repeat, void* stream) {
  if (!mean || !rstd || !mean_out || !rstd_out || N <= 0 || C <= 0 || (!running_mean != !running_var)) return HM_ERR_INVALID;
  bn_fold_kernel<<<(N * C + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      mean, rstd, gamma, beta, N, C, mean_out, rstd_out, running_mean, running_var, num_batches_tracked, count, momentum, eps,
      repeat);
  return HM_LAUNCH_OK();
}

int hm_upsample2_add(const float* small, const float* deep, int N, int h, int w, int C, float* out, void* stream) {
  if (!small || !deep || !out || (C & 3)) return HM_ERR_INVALID;
  upsample2_add_kernel<<<grid_for(long(N) * 4 * h * w * (C >> 2)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      small, deep, N, h, w, C, out);
  return HM_LAUNCH_OK();
}

int hm_box2mask_head(const float* ctx_logit, const float* obj_logit, int obj_ld, const float* label_map,
                     const float* mask_out, const float* inst, int N, int H, int W, int C, int use_gate, float* comb_logit,
                     float* comb_logprob, float* obj_prob, double* acc, void* stream) {
  if (!ctx_logit || !obj_logit || C <= 0 || obj_ld <= 0) return HM_ERR_INVALID;
  b2m_head_kernel<<<grid_for(long(N) * H * W, kBlock, 148 * 8), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      ctx_logit, obj_logit, obj_ld, label_map, mask_out, inst, N, H, W, C, use_gate, comb_logit, comb_logprob, obj_prob,
      acc);
  return HM_LAUNCH_OK();
}

}  // extern "C"
