// hm_thin.cu -- "tap-unrolled" lowering of convolutions with a THIN side (<= 4 channels): the generator head
// Conv2d(ngf, 3, 7) (models/Pix2Pix_NET.py:91), the PatchGAN output Conv2d(512, 1, 4) (models/Discriminator_NET.py:93)
// and VGG19 conv1_1 Conv2d(3, 64, 3) (torchvision features[0] via models/layer_util.py:384-390).
//
// Why: the tcgen05 engines contract over 64-channel chunks per filter tap and emit N tiles of >= 16 columns, so a
// 3-channel side wastes 95 % of every MMA (head: fprop 3.2 ms, dgrad 3.7 ms, wgrad 4.4 ms per step at 512x1024 x4 in
// bf16x3 -- 11 % of the step for 0.6 % of its FLOPs).  Folding the KW horizontal taps of the thin side into its
// channel index turns the same arithmetic into GEMMs with KW x fewer MMAs:
//
//   thin OUTPUT (Cout = C <= 4):   T[n,h,w',(kw,co)] = sum_{kh,ci} x[n,h+kh-p,w'-p,ci] W[co,ci,kh,kw]   (KH x 1 conv, N = KW*C)
//                                  y[n,h,w,co]       = act(b[co] + sum_kw T[n,h,w+kw,(kw,co)])            (hm_tap_combine)
//        gradients use  U[n,h,w',(kw,co)] = dy[n,h,w'-kw,co]  (hm_tap_unroll):
//                                  dx = KH x 1 dgrad of U,   dW[co,ci,kh,kw] = sum_q x[q+(kh-p,-p)][ci] U[q][(kw,co)]
//   thin INPUT  (Cin = C <= 4):    U[n,h,w,(kh,kw,ci)] = x[n,h+kh-p,w+kw-p,ci]  (hm_tap_unroll), y = 1x1 conv of U (K = KH*KW*C),
//                                  dU = 1x1 dgrad,  dx[n,h,w,ci] = sum_{kh,kw} dU[n,h-kh+p,w-kw+p,(kh,kw,ci)]  (hm_tap_combine)
//
// The kernels here are the HBM-bound glue (a few bytes per pixel); the contractions stay on the tcgen05 engines.
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

__global__ void pack_weight_ex_kernel(const float* __restrict__ src, int rows, int r_div, long s_r_hi, long s_r_lo, int kk,
                                      int k_div, long s_k_hi, long s_k_lo, int taps, long s_tap, int rows_pad, int k_pad,
                                      const float* __restrict__ scale, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  const long total = long(taps) * rows_pad * k_pad;
  const float sc = scale ? __ldg(scale) : 1.f;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int k = int(i % k_pad);
    const long rt = i / k_pad;
    const int r = int(rt % rows_pad);
    const int t = int(rt / rows_pad);
    float v = 0.f;
    if (r < rows && k < kk)
      v = __ldg(src + (r / r_div) * s_r_hi + (r % r_div) * s_r_lo + (k / k_div) * s_k_hi + (k % k_div) * s_k_lo + t * s_tap) * sc;
    bf16 h, l;
    hm::split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// G[(kh*cp_pad + p)][kw*cq + q] (ld = ldg) -> dst[q][p][kh][kw]
__global__ void wgrad_unpack_cols_kernel(const float* __restrict__ G, int KH, int KW, int cp, int cq, int cp_pad, int ldg,
                                         float* __restrict__ dst, int accumulate) {
  const long total = long(cq) * cp * KH * KW;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int kw = int(i % KW);
    long r = i / KW;
    const int kh = int(r % KH); r /= KH;
    const int pch = int(r % cp);
    const int q = int(r / cp);
    const float v = __ldg(G + (long(kh) * cp_pad + pch) * ldg + kw * cq + q);
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

// one thread per (destination pixel, group of 8 destination channels)
__global__ void tap_unroll_kernel(const bf16* __restrict__ s_hi, const bf16* __restrict__ s_lo, int N, int Hs, int Ws, int C,
                                  int s_cs, int KH, int KW, int oh, int ow, int sh, int sw, bf16* __restrict__ d_hi,
                                  bf16* __restrict__ d_lo, int Hd, int Wd, int d_cs) {
  const int groups = d_cs >> 3;
  const long total = long(N) * Hd * Wd * groups;
  const int nch = KH * KW * C;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % groups);
    long pix = i / groups;
    const int w = int(pix % Wd);
    long r = pix / Wd;
    const int h = int(r % Hd);
    const int n = int(r / Hd);
    alignas(16) bf16 vh[8];
    alignas(16) bf16 vl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int d = g * 8 + e;
      bf16 a = __float2bfloat16_rn(0.f), b = a;
      if (d < nch) {
        const int tap = d / C, c = d - tap * C;
        const int j = tap / KW, t = tap - j * KW;
        const int hs = h + oh + sh * j, ws = w + ow + sw * t;
        if (hs >= 0 && hs < Hs && ws >= 0 && ws < Ws) {
          const size_t off = ((size_t(n) * Hs + hs) * Ws + ws) * s_cs + c;
          a = s_hi[off];
          if (s_lo) b = s_lo[off];
        }
      }
      vh[e] = a; vl[e] = b;
    }
    const size_t doff = size_t(pix) * d_cs + g * 8;
    *reinterpret_cast<uint4*>(d_hi + doff) = *reinterpret_cast<const uint4*>(vh);
    if (d_lo) *reinterpret_cast<uint4*>(d_lo + doff) = *reinterpret_cast<const uint4*>(vl);
  }
}

// Fast path for an 8-channel-stride source (the 3-channel image / gradient operands): one thread per destination pixel
// reads each source pixel of its window with ONE 16 B load per plane and writes its d_cs (<= 32) channels with 16 B
// stores (the generic kernel above spends a div/mod and a 2-byte load per destination element).
template <int DG>   // destination channel groups of 8 (d_cs = 8 * DG)
__global__ void tap_unroll8_kernel(const bf16* __restrict__ s_hi, const bf16* __restrict__ s_lo, int N, int Hs, int Ws, int C,
                                   int KH, int KW, int oh, int ow, int sh, int sw, bf16* __restrict__ d_hi,
                                   bf16* __restrict__ d_lo, int Hd, int Wd) {
  const long total = long(N) * Hd * Wd;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int w = int(i % Wd);
    long r = i / Wd;
    const int h = int(r % Hd);
    const int n = int(r / Hd);
    alignas(16) bf16 vh[8 * DG];
    alignas(16) bf16 vl[8 * DG];
#pragma unroll
    for (int e = 0; e < 8 * DG; ++e) { vh[e] = __float2bfloat16_rn(0.f); vl[e] = vh[e]; }
    int d = 0;
    for (int j = 0; j < KH; ++j) {
      const int hs = h + oh + sh * j;
      for (int t = 0; t < KW; ++t, d += C) {
        const int ws = w + ow + sw * t;
        if (hs < 0 || hs >= Hs || ws < 0 || ws >= Ws) continue;
        const size_t off = ((size_t(n) * Hs + hs) * Ws + ws) * 8;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(s_hi + off));
        const bf16* ah = reinterpret_cast<const bf16*>(&a);
        uint4 b = make_uint4(0, 0, 0, 0);
        if (s_lo && d_lo) b = __ldg(reinterpret_cast<const uint4*>(s_lo + off));
        const bf16* al = reinterpret_cast<const bf16*>(&b);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < C) {
            // d + c < 8 * DG is guaranteed by the launcher (KH*KW*C <= d_cs); constant-index the register arrays
#pragma unroll
            for (int e = 0; e < 8 * DG; ++e)
              if (e == d + c) { vh[e] = ah[c]; vl[e] = al[c]; }
          }
      }
    }
    const size_t doff = size_t(i) * (8 * DG);
#pragma unroll
    for (int g = 0; g < DG; ++g) {
      *reinterpret_cast<uint4*>(d_hi + doff + g * 8) = *reinterpret_cast<const uint4*>(vh + g * 8);
      if (d_lo) *reinterpret_cast<uint4*>(d_lo + doff + g * 8) = *reinterpret_cast<const uint4*>(vl + g * 8);
    }
  }
}

// one thread per output pixel, C <= 4 channels each
__global__ void tap_combine_kernel(const float* __restrict__ T, int N, int Ht, int Wt, int ldT, int KH, int KW, int C, int oh,
                                   int ow, int sh, int sw, const float* __restrict__ bias, int act, float slope,
                                   float* __restrict__ out, int Ho, int Wo, int ldo) {
  const long total = long(N) * Ho * Wo;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int w = int(i % Wo);
    long r = i / Wo;
    const int h = int(r % Ho);
    const int n = int(r / Ho);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < KH; ++j) {
      const int ht = h + oh + sh * j;
      if (ht < 0 || ht >= Ht) continue;
      for (int t = 0; t < KW; ++t) {
        const int wt = w + ow + sw * t;
        if (wt < 0 || wt >= Wt) continue;
        const float* p = T + ((size_t(n) * Ht + ht) * Wt + wt) * ldT + (j * KW + t) * C;
#pragma unroll
        for (int c = 0; c < 4; ++c) if (c < C) acc[c] += __ldg(p + c);
      }
    }
    float* o = out + size_t(i) * ldo;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c < C) {
        float v = acc[c] + (bias ? __ldg(bias + c) : 0.f);
        if (act == HM_ACT_RELU) v = fmaxf(v, 0.f);
        else if (act == HM_ACT_LRELU) v = v > 0.f ? v : v * slope;
        else if (act == HM_ACT_TANH) v = tanhf(v);
        o[c] = v;
      }
  }
}

}  // namespace

extern "C" {

int hm_pack_weight_ex(const float* src, int rows, int r_div, long s_r_hi, long s_r_lo, int k, int k_div, long s_k_hi,
                      long s_k_lo, int taps, long s_tap, const float* scale, void* dst_hi, void* dst_lo, void* stream) {
  if (!src || !dst_hi || rows <= 0 || k <= 0 || taps <= 0 || r_div <= 0 || k_div <= 0) return HM_ERR_INVALID;
  const int rows_pad = hm_rows_pad(rows), k_pad = hm_k_pad(k);
  const long total = long(taps) * rows_pad * k_pad;
  pack_weight_ex_kernel<<<grid_for(total, kBlock, 148 * 16), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      src, rows, r_div, s_r_hi, s_r_lo, k, k_div, s_k_hi, s_k_lo, taps, s_tap, rows_pad, k_pad, scale,
      static_cast<bf16*>(dst_hi), static_cast<bf16*>(dst_lo));
  return HM_LAUNCH_OK();
}

int hm_wgrad_unpack_cols(const float* G_ws, int KH, int KW, int cp, int cq, float* dst, int accumulate, void* stream) {
  if (!G_ws || !dst || KH <= 0 || KW <= 0 || cp <= 0 || cq <= 0) return HM_ERR_INVALID;
  const int cp_pad = (cp + 63) / 64 * 64, ldg = (KW * cq + 63) / 64 * 64;
  const long total = long(cq) * cp * KH * KW;
  wgrad_unpack_cols_kernel<<<grid_for(total, kBlock, 148 * 16), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      G_ws, KH, KW, cp, cq, cp_pad, ldg, dst, accumulate);
  return HM_LAUNCH_OK();
}

int hm_tap_unroll(const void* s_hi, const void* s_lo, int N, int Hs, int Ws, int C, int s_cs, int KH, int KW, int oh,
                  int ow, int sh, int sw, void* d_hi, void* d_lo, int Hd, int Wd, int d_cs, void* stream) {
  if (!s_hi || !d_hi || C <= 0 || KH <= 0 || KW <= 0 || (d_cs & 7) || KH * KW * C > d_cs || N <= 0 || Hd <= 0 || Wd <= 0)
    return HM_ERR_INVALID;
  if (s_cs == 8 && C <= 4 && (d_cs == 8 || d_cs == 16 || d_cs == 24 || d_cs == 32)) {
    const long px = long(N) * Hd * Wd;
    const bf16* sh_ = static_cast<const bf16*>(s_hi);
    const bf16* sl_ = static_cast<const bf16*>(s_lo);
    bf16* dh_ = static_cast<bf16*>(d_hi);
    bf16* dl_ = static_cast<bf16*>(d_lo);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (d_cs >> 3) {
      case 1: tap_unroll8_kernel<1><<<grid_for(px), kBlock, 0, st>>>(sh_, sl_, N, Hs, Ws, C, KH, KW, oh, ow, sh, sw, dh_, dl_, Hd, Wd); break;
      case 2: tap_unroll8_kernel<2><<<grid_for(px), kBlock, 0, st>>>(sh_, sl_, N, Hs, Ws, C, KH, KW, oh, ow, sh, sw, dh_, dl_, Hd, Wd); break;
      case 3: tap_unroll8_kernel<3><<<grid_for(px), kBlock, 0, st>>>(sh_, sl_, N, Hs, Ws, C, KH, KW, oh, ow, sh, sw, dh_, dl_, Hd, Wd); break;
      default: tap_unroll8_kernel<4><<<grid_for(px), kBlock, 0, st>>>(sh_, sl_, N, Hs, Ws, C, KH, KW, oh, ow, sh, sw, dh_, dl_, Hd, Wd); break;
    }
    return HM_LAUNCH_OK();
  }
  const long total = long(N) * Hd * Wd * (d_cs >> 3);
  tap_unroll_kernel<<<grid_for(total), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(s_hi), static_cast<const bf16*>(s_lo), N, Hs, Ws, C, s_cs, KH, KW, oh, ow, sh, sw,
      static_cast<bf16*>(d_hi), static_cast<bf16*>(d_lo), Hd, Wd, d_cs);
  return HM_LAUNCH_OK();
}

int hm_tap_combine(const float* T, int N, int Ht, int Wt, int ldT, int KH, int KW, int C, int oh, int ow, int sh, int sw,
                   const float* bias, int act, float slope, float* out, int Ho, int Wo, int ldo, void* stream) {
  if (!T || !out || C <= 0 || C > 4 || KH <= 0 || KW <= 0 || KH * KW * C > ldT || C > ldo || N <= 0) return HM_ERR_INVALID;
  const long total = long(N) * Ho * Wo;
  tap_combine_kernel<<<grid_for(total), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      T, N, Ht, Wt, ldT, KH, KW, C, oh, ow, sh, sw, bias, act, slope, out, Ho, Wo, ldo);
  return HM_LAUNCH_OK();
}

}  // extern "C"
