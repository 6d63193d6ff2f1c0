// hm_twostream.cu -- HBM-bound glue of GlobalTwoStreamGenerator (models/Pix2Pix_NET.py:103-247), the generator the
// reference's shipped scripts train (scripts/train_mask2image_city.sh): the object-mask max-pool, the masked fusion of
// the context and label streams ('early_add': (1-m)*ctx + m*obj, :207-209 + FeatureFusionBlock 'add'), its adjoint,
// the skip concatenation of encoder features into the decoder (:221) and the context-stream input operand.
// Convolutions / InstanceNorm / activations reuse the engines and K7 kernels.
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

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void mask_maxpool_kernel(const float* __restrict__ mask, int B, int H, int W, int f, float* __restrict__ out) {
  const int Ho = H / f, Wo = W / f;
  const long total = long(B) * Ho * Wo;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int w = int(i % Wo);
    long r = i / Wo;
    const int h = int(r % Ho);
    const int n = int(r / Ho);
    float m = -3.402823466e38f;
    for (int y = 0; y < f; ++y)
      for (int x = 0; x < f; ++x) m = fmaxf(m, __ldg(mask + (long(n) * H + h * f + y) * W + w * f + x));
    out[i] = m;
  }
}

// one thread per (padded pixel, group of 8 channels)
__global__ void mask_blend_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ m,
                                  int N, int H, int W, int C, float* __restrict__ out32, bf16* __restrict__ o_hi,
                                  bf16* __restrict__ o_lo, int o_cs, int border) {
  const int Hp = H + 2 * border, Wp = W + 2 * border;
  const int groups = (o_hi ? o_cs : ((C + 7) & ~7)) >> 3;
  const long total = long(N) * Hp * Wp * groups;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % groups);
    long r = i / groups;
    const int wp = int(r % Wp); r /= Wp;
    const int hp = int(r % Hp);
    const int n = int(r / Hp);
    const int h = reflect_idx(hp - border, H), w = reflect_idx(wp - border, W);
    const bool interior = (hp - border == h) && (wp - border == w);
    const size_t pix = (size_t(n) * H + h) * W + w;
    const float mm = a && b ? __ldg(m + pix) : 0.f;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      float x = 0.f;
      if (c < C) {
        const float av = a ? __ldg(a + pix * C + c) : 0.f;
        const float bv = b ? __ldg(b + pix * C + c) : 0.f;
        x = (a && b) ? (1.f - mm) * av + mm * bv : (a ? av : bv);
      }
      v[j] = x;
    }
    if (out32 && interior) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (g * 8 + j < C) out32[pix * C + g * 8 + j] = v[j];
    }
    if (o_hi) {
      alignas(16) bf16 hh[8];
      alignas(16) bf16 ll[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) hm::split_bf16(v[j], hh[j], ll[j]);
      const size_t off = ((size_t(n) * Hp + hp) * Wp + wp) * o_cs + g * 8;
      *reinterpret_cast<uint4*>(o_hi + off) = *reinterpret_cast<const uint4*>(hh);
      if (o_lo) *reinterpret_cast<uint4*>(o_lo + off) = *reinterpret_cast<const uint4*>(ll);
    }
  }
}

__global__ void mask_blend_bwd_kernel(const float* __restrict__ g, const float* __restrict__ m, long P, int C,
                                      float* __restrict__ da, float* __restrict__ db) {
  const long total = P * C;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const float mm = __ldg(m + i / C), gv = __ldg(g + i);
    if (da) da[i] = (1.f - mm) * gv;
    if (db) db[i] = mm * gv;
  }
}

// FeatureFusionBlock 'concat' (layer_util.py:305-327): operand [N,H,W,2C] = relu(cat((1 - m) * a, m * b)), no border (it
// feeds the 1x1 fuse convolution).  One thread per (pixel, group of 8 output channels); C % 8 == 0.
__global__ void mask_concat_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ m,
                                   long P, int C, bf16* __restrict__ o_hi, bf16* __restrict__ o_lo, int o_cs) {
  const int groups = o_cs >> 3;
  const long total = P * groups;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % groups);
    const long pix = i / groups;
    const float mm = __ldg(m + pix);
    const int c0 = g * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float x = 0.f;
      if (c < C) x = (1.f - mm) * __ldg(a + pix * C + c);
      else if (c < 2 * C) x = mm * __ldg(b + pix * C + c - C);
      v[j] = fmaxf(x, 0.f);
    }
    alignas(16) bf16 hh[8];
    alignas(16) bf16 ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hm::split_bf16(v[j], hh[j], ll[j]);
    const size_t off = size_t(pix) * o_cs + c0;
    *reinterpret_cast<uint4*>(o_hi + off) = *reinterpret_cast<const uint4*>(hh);
    if (o_lo) *reinterpret_cast<uint4*>(o_lo + off) = *reinterpret_cast<const uint4*>(ll);
  }
}

// its adjoint: g [P, ld >= 2C] is the gradient w.r.t. the concatenated operand;
// da = (1 - m) * g[:, :C] where (1 - m) * a > 0, db = m * g[:, C:2C] where m * b > 0
__global__ void mask_concat_bwd_kernel(const float* __restrict__ g, int ld, const float* __restrict__ m,
                                       const float* __restrict__ a, const float* __restrict__ b, long P, int C,
                                       float* __restrict__ da, float* __restrict__ db) {
  const long total = P * C;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const long pix = i / C;
    const int c = int(i - pix * C);
    const float mm = __ldg(m + pix);
    const float wa = 1.f - mm;
    da[i] = (wa * __ldg(a + i) > 0.f) ? wa * __ldg(g + pix * ld + c) : 0.f;
    db[i] = (mm * __ldg(b + i) > 0.f) ? mm * __ldg(g + pix * ld + C + c) : 0.f;
  }
}

// ImagePool.query for ONE image (util/image_pool.py:11-31): `dec` holds the host-drawn decision of every image of the batch
// in device memory (so that the launch can sit in a CUDA graph): dec[2b] = 0 pass through, 1 store in slot dec[2b+1] and
// pass through (pool still filling), 2 exchange with slot dec[2b+1] (return the stored image, keep the new one).
__global__ void pool_exchange_kernel(const uint4* __restrict__ cur_hi, const uint4* __restrict__ cur_lo, uint4* __restrict__ pool_hi,
                                     uint4* __restrict__ pool_lo, uint4* __restrict__ out_hi, uint4* __restrict__ out_lo,
                                     const int* __restrict__ dec, int b, long vecs) {
  const int action = dec[2 * b];
  const long slot = dec[2 * b + 1];
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < vecs; i += long(gridDim.x) * blockDim.x) {
    const long ci = long(b) * vecs + i, pi = slot * vecs + i;
    const uint4 ch = cur_hi[ci];
    uint4 cl = make_uint4(0, 0, 0, 0);
    if (cur_lo) cl = cur_lo[ci];
    uint4 oh = ch, ol = cl;
    if (action == 2) {
      oh = pool_hi[pi];
      if (pool_lo) ol = pool_lo[pi];
    }
    if (action != 0) {
      pool_hi[pi] = ch;
      if (pool_lo) pool_lo[pi] = cl;
    }
    out_hi[ci] = oh;
    if (out_lo) out_lo[ci] = ol;
  }
}

// one thread per (pixel, group of 8 output channels); groups never straddle the a / b boundary when Ca % 8 == 0,
// otherwise elements are gathered one by one
__global__ void concat_operands_kernel(const bf16* __restrict__ a_hi, const bf16* __restrict__ a_lo, int a_cs, int Ca,
                                       const bf16* __restrict__ b_hi, const bf16* __restrict__ b_lo, int b_cs, int Cb,
                                       bf16* __restrict__ o_hi, bf16* __restrict__ o_lo, int o_cs, long P) {
  const int groups = o_cs >> 3;
  const long total = P * groups;
  const bf16 zero = __float2bfloat16_rn(0.f);
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % groups);
    const long pix = i / groups;
    alignas(16) bf16 hh[8];
    alignas(16) bf16 ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = g * 8 + j;
      bf16 h = zero, l = zero;
      if (c < Ca) { h = a_hi[pix * a_cs + c]; if (a_lo) l = a_lo[pix * a_cs + c]; }
      else if (c < Ca + Cb) { h = b_hi[pix * b_cs + c - Ca]; if (b_lo) l = b_lo[pix * b_cs + c - Ca]; }
      hh[j] = h; ll[j] = l;
    }
    const size_t off = size_t(pix) * o_cs + g * 8;
    *reinterpret_cast<uint4*>(o_hi + off) = *reinterpret_cast<const uint4*>(hh);
    if (o_lo) *reinterpret_cast<uint4*>(o_lo + off) = *reinterpret_cast<const uint4*>(ll);
  }
}

// cond = (1 - mask) * image (NULLVAL = 0, pix2pixHD_condImg_model.py:165-166) as an operand with ReflectionPad2d(border)
__global__ void cond_image_kernel(const float* __restrict__ image, const float* __restrict__ mask, int B, int H, int W,
                                  bf16* __restrict__ o_hi, bf16* __restrict__ o_lo, int o_cs, int border) {
  const int Hp = H + 2 * border, Wp = W + 2 * border;
  const long total = long(B) * Hp * Wp;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int wp = int(i % Wp);
    long r = i / Wp;
    const int hp = int(r % Hp);
    const int n = int(r / Hp);
    const int h = reflect_idx(hp - border, H), w = reflect_idx(wp - border, W);
    const float mm = __ldg(mask + (long(n) * H + h) * W + w);
    alignas(16) bf16 hh[8];
    alignas(16) bf16 ll[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = 0.f;
      if (j < 3) x = (1.f - mm) * __ldg(image + (long(n) * 3 + j) * H * W + long(h) * W + w);
      hm::split_bf16(x, hh[j], ll[j]);
    }
    for (int g = 0; g < (o_cs >> 3); ++g) {
      const size_t off = size_t(i) * o_cs + g * 8;
      if (g == 0) {
        *reinterpret_cast<uint4*>(o_hi + off) = *reinterpret_cast<const uint4*>(hh);
        if (o_lo) *reinterpret_cast<uint4*>(o_lo + off) = *reinterpret_cast<const uint4*>(ll);
      } else {
        *reinterpret_cast<uint4*>(o_hi + off) = make_uint4(0, 0, 0, 0);
        if (o_lo) *reinterpret_cast<uint4*>(o_lo + off) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

}  // namespace

extern "C" {

int hm_mask_maxpool(const float* mask, int B, int H, int W, int f, float* out, void* stream) {
  if (!mask || !out || f <= 0 || H % f || W % f) return HM_ERR_INVALID;
  mask_maxpool_kernel<<<grid_for(long(B) * (H / f) * (W / f)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(mask, B, H, W, f, out);
  return HM_LAUNCH_OK();
}

int hm_mask_blend(const float* a, const float* b, const float* m, int N, int H, int W, int C, float* out32, void* o_hi,
                  void* o_lo, int o_cs, int border, void* stream) {
  if ((!a && !b) || (a && b && !m) || (!out32 && !o_hi) || (o_hi && ((o_cs & 7) || o_cs < C))) return HM_ERR_INVALID;
  if (border > 0 && (border >= H || border >= W)) return HM_ERR_INVALID;
  const int groups = (o_hi ? o_cs : ((C + 7) & ~7)) >> 3;
  const long total = long(N) * (H + 2 * border) * (W + 2 * border) * groups;
  mask_blend_kernel<<<grid_for(total), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      a, b, m, N, H, W, C, out32, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs, border);
  return HM_LAUNCH_OK();
}

int hm_mask_blend_bwd(const float* g, const float* m, long P, int C, float* da, float* db, void* stream) {
  if (!g || !m || (!da && !db)) return HM_ERR_INVALID;
  mask_blend_bwd_kernel<<<grid_for(P * C), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(g, m, P, C, da, db);
  return HM_LAUNCH_OK();
}

int hm_pool_exchange(const void* cur_hi, const void* cur_lo, void* pool_hi, void* pool_lo, void* out_hi, void* out_lo,
                     const int* dec, int b, long elems_per_image, void* stream) {
  if (!cur_hi || !pool_hi || !out_hi || !dec || b < 0 || elems_per_image <= 0 || (elems_per_image & 7)) return HM_ERR_INVALID;
  const long vecs = elems_per_image >> 3;
  pool_exchange_kernel<<<grid_for(vecs), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(cur_hi), static_cast<const uint4*>(cur_lo), static_cast<uint4*>(pool_hi),
      static_cast<uint4*>(pool_lo), static_cast<uint4*>(out_hi), static_cast<uint4*>(out_lo), dec, b, vecs);
  return HM_LAUNCH_OK();
}

int hm_mask_concat(const float* a, const float* b, const float* m, long P, int C, void* o_hi, void* o_lo, int o_cs,
                   void* stream) {
  if (!a || !b || !m || !o_hi || (o_cs & 7) || o_cs < 2 * C || C <= 0) return HM_ERR_INVALID;
  mask_concat_kernel<<<grid_for(P * (o_cs >> 3)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      a, b, m, P, C, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs);
  return HM_LAUNCH_OK();
}

int hm_mask_concat_bwd(const float* g, int ld, const float* m, const float* a, const float* b, long P, int C, float* da,
                       float* db, void* stream) {
  if (!g || !m || !a || !b || !da || !db || ld < 2 * C) return HM_ERR_INVALID;
  mask_concat_bwd_kernel<<<grid_for(P * C), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(g, ld, m, a, b, P, C, da, db);
  return HM_LAUNCH_OK();
}

int hm_concat_operands(const void* a_hi, const void* a_lo, int a_cs, int Ca, const void* b_hi, const void* b_lo, int b_cs,
                       int Cb, void* o_hi, void* o_lo, int o_cs, long P, void* stream) {
  if (!a_hi || !b_hi || !o_hi || (o_cs & 7) || Ca + Cb > o_cs || Ca > a_cs || Cb > b_cs) return HM_ERR_INVALID;
  concat_operands_kernel<<<grid_for(P * (o_cs >> 3)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(a_hi), static_cast<const bf16*>(a_lo), a_cs, Ca, static_cast<const bf16*>(b_hi),
      static_cast<const bf16*>(b_lo), b_cs, Cb, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs, P);
  return HM_LAUNCH_OK();
}

int hm_cond_image_operand(const float* image, const float* mask, int B, int H, int W, void* o_hi, void* o_lo, int o_cs,
                          int border, void* stream) {
  if (!image || !mask || !o_hi || (o_cs & 7) || o_cs < 8) return HM_ERR_INVALID;
  cond_image_kernel<<<grid_for(long(B) * (H + 2 * border) * (W + 2 * border)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      image, mask, B, H, W, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs, border);
  return HM_LAUNCH_OK();
}

}  // extern "C"
