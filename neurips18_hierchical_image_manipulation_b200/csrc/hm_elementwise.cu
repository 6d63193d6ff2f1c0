// hm_elementwise.cu -- the HBM-bound kernels of the mask2image hot path (everything that is not a
// tensor-core contraction): input encoding, InstanceNorm statistics / apply / backward, pooling,
// loss reductions and gradients, output gate, fused Adam.  All tensors are NHWC; one thread owns a
// group of 8 consecutive channels of one pixel so that fp32 traffic is 2 x 16 B and bf16 traffic 16 B
// per access, fully coalesced across a warp.  Reference lines are cited per kernel.
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

__device__ __forceinline__ int reflect_idx(int i, int n) {  // index into [0,n) of reflect-padded coordinate i
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// load 8 consecutive fp32 channels [c, c+8) of a pixel whose channel count is C (row pointer p); zeros past C
__device__ __forceinline__ void load8(const float* __restrict__ p, int c, int C, float (&v)[8]) {
  if (((C & 3) == 0) && c + 8 <= C) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + c + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (c + i < C) ? __ldg(p + c + i) : 0.f;
  }
}
__device__ __forceinline__ void store8(float* __restrict__ p, int c, int C, const float (&v)[8]) {
  if (((C & 3) == 0) && c + 8 <= C) {
    *reinterpret_cast<float4*>(p + c) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (c + i < C) p[c + i] = v[i];
  }
}
// bf16 operand planes: channel stride cs is a multiple of 8, so a group of 8 is one aligned 16 B word
__device__ __forceinline__ void store_op8(bf16* __restrict__ hi, bf16* __restrict__ lo, size_t off, const float (&v)[8]) {
  alignas(16) bf16 h[8];
  alignas(16) bf16 l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hm::split_bf16(v[i], h[i], l[i]);
  *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(h);
  if (lo) *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<const uint4*>(l);
}
__device__ __forceinline__ void load_op8(const bf16* __restrict__ hi, const bf16* __restrict__ lo, size_t off, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(hi + off));
  const bf16* h = reinterpret_cast<const bf16*>(&a);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __bfloat162float(h[i]);
  if (lo) {
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(lo + off));
    const bf16* l = reinterpret_cast<const bf16*>(&b);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += __bfloat162float(l[i]);
  }
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == HM_ACT_RELU) return fmaxf(v, 0.f);
  if (act == HM_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == HM_ACT_TANH) return tanhf(v);
  return v;
}
__device__ __forceinline__ float act_grad(float pre_sign_src, int act, float slope) {
  // derivative of relu / lrelu from anything with the sign of the pre-activation
  if (act == HM_ACT_RELU) return pre_sign_src > 0.f ? 1.f : 0.f;
  if (act == HM_ACT_LRELU) return pre_sign_src > 0.f ? 1.f : slope;
  return 1.f;
}

// ================================================================================================
// K11  encode_input  (models/pix2pixHD_condImg_model.py:144-174, get_edges :285-291)
// one thread per (padded) generator-input pixel: one-hot(label) | edge(inst) | (1-mask)*image
// ================================================================================================
// One thread per (pixel, group of 8 output channels) of one of the three operands, groups fastest, so that the 16 B
// stores of a warp are contiguous (a thread-per-pixel version wrote 80-96 B apart and ran at 1.4 TB/s); the few scalar
// inputs of a pixel are re-read by its 5-6 channel-group threads through L1.  The grid is (row segment, row, image) per
// operand kind, so the only index arithmetic left is one division by the group count, done as a float multiply (the
// flat-index version spent ~4 integer divisions by run-time values per 16 B store and ran at 29 % of the HBM peak).
struct EncodeArgs {
  const float* label; const float* inst; const float* image; const float* mask; const float* d_mask;
  int B, H, W, label_nc, d_no_imgcond;
  bf16* g_hi; bf16* g_lo; int g_cs, gb;
  bf16* d_hi; bf16* d_lo; int d_cs;
  bf16* v_hi; bf16* v_lo; int v_cs;
};

template <int KIND>   // 0: generator operand (with reflect border), 1: discriminator operand (2B images), 2: VGG operand
__global__ void __launch_bounds__(kBlock) encode_kernel(EncodeArgs a, int groups, float inv_groups) {
  const int H = a.H, W = a.W;
  const int Wrow = KIND == 0 ? W + 2 * a.gb : W;
  const int t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= Wrow * groups) return;
  int wp = __float2int_rd((t + 0.5f) * inv_groups);
  int grp = t - wp * groups;
  if (grp < 0) { --wp; grp += groups; } else if (grp >= groups) { ++wp; grp -= groups; }
  const int hp = blockIdx.y;
  int n = blockIdx.z, half = 0;
  int h = hp, w = wp;
  size_t off;
  if (KIND == 0) {
    h = reflect_idx(hp - a.gb, H); w = reflect_idx(wp - a.gb, W);
    off = ((size_t(n) * (H + 2 * a.gb) + hp) * Wrow + wp) * a.g_cs + grp * 8;
  } else if (KIND == 1) {
    half = n >= a.B ? 1 : 0;           // 0 = fake (image channels filled by hm_finish_fake), 1 = real
    off = ((size_t(n) * H + h) * W + w) * a.d_cs + grp * 8;
    n -= half * a.B;
  } else {
    off = ((size_t(n) + a.B) * H * W + size_t(h) * W + w) * a.v_cs + grp * 8;
  }
  const int label_nc = a.label_nc;
  const int n_cond = label_nc + (a.inst ? 1 : 0);
  const long pix = (long(n) * H + h) * W + w;
  const int c0 = grp * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  const bool d_img_only = KIND == 1 && (a.d_no_imgcond & 2);          // which_encoder == 'ctx': D sees the bare image
  const bool need_img = (KIND == 2 || d_img_only) ? (c0 < 3) : (c0 + 8 > n_cond);   // (also covers the shifted no_imgCond layout)
  float img[3] = {0.f, 0.f, 0.f}, cond[3] = {0.f, 0.f, 0.f};
  if (need_img) {
    const float m = __ldg(a.mask + pix);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      img[c] = __ldg(a.image + (long(n) * 3 + c) * H * W + long(h) * W + w);
      cond[c] = (1.f - m) * img[c];  // NULLVAL == 0 (:18, :165-166)
    }
  }
  if (KIND == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c0 + j < 3) v[j] = img[c0 + j];
  } else if (d_img_only) {             // [image] only: real half = the image, fake half filled by hm_finish_fake
#pragma unroll
    for (int j = 0; j < 8; ++j) if (half == 1 && c0 + j < 3) v[j] = img[c0 + j];
    if (a.d_mask) {
      const float dm = __ldg(a.d_mask + pix);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= dm;
    }
  } else {
    if (c0 < label_nc) {
      const int cls = int(__ldg(a.label + pix));
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c0 + j < label_nc && c0 + j == cls) v[j] = 1.f;
    }
    if (a.inst && c0 <= label_nc && label_nc < c0 + 8) {  // :285-291: 4-neighbour instance boundary
      const float tt = __ldg(a.inst + pix);
      bool e = false;
      if (w > 0) e |= (tt != __ldg(a.inst + pix - 1));
      if (w < W - 1) e |= (tt != __ldg(a.inst + pix + 1));
      if (h > 0) e |= (tt != __ldg(a.inst + pix - W));
      if (h < H - 1) e |= (tt != __ldg(a.inst + pix + W));
      v[label_nc - c0] = e ? 1.f : 0.f;
    }
    // D operand with no_imgCond (pix2pixHD_condImg_model.py:213-214): [label | edge | image], else [.. | cond | image]
    const bool no_ic = KIND == 1 && (a.d_no_imgcond & 1);
    const int cimg = no_ic ? n_cond : n_cond + 3;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      if (!no_ic && c >= n_cond && c < n_cond + 3) v[j] = cond[c - n_cond];
      else if (KIND == 1 && half == 1 && c >= cimg && c < cimg + 3) v[j] = img[c - cimg];
    }
    if (KIND == 1 && a.d_mask) {   // mask_gan_input (:180-181): the whole D input is multiplied by the mask
      const float dm = __ldg(a.d_mask + pix);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= dm;
    }
  }
  if (KIND == 0) store_op8(a.g_hi, a.g_lo, off, v);
  else if (KIND == 1) store_op8(a.d_hi, a.d_lo, off, v);
  else store_op8(a.v_hi, a.v_lo, off, v);
}

// ================================================================================================
// K7  InstanceNorm2d(affine=False) statistics (models/layer_util.py:19-26): biased variance per (n,c)
// pass 1: per-block partial (sum, sumsq) ; pass 2: combine in fp64 -> mean, rstd
// ================================================================================================
__global__ void in_stats_kernel(const float* __restrict__ y, int HW, int C, int gx_log2, float* __restrict__ partial) {
  // grid (nblk, N, cgroups); thread (tx = channel quad, ty = pixel lane)
  const int gx = 1 << gx_log2, rows = kBlock >> gx_log2;
  const int tx = threadIdx.x & (gx - 1), ty = threadIdx.x >> gx_log2;
  const int n = blockIdx.y, nblk = gridDim.x;
  const int c4 = blockIdx.z * gx + tx;
  const bool cvalid = c4 * 4 < C;
  const int per = (HW + nblk - 1) / nblk;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (cvalid) {
    const float* base = y + size_t(n) * HW * C + c4 * 4;
    // shifted sums: accumulate (x - pivot) with the plane's first pixel as pivot, so that
    // var = E[d^2] - E[d]^2 does not cancel catastrophically when |mean| >> std
    const float4 pv = __ldg(reinterpret_cast<const float4*>(base));
#pragma unroll 4
    for (int p = p0 + ty; p < p1; p += rows) {
      float4 v = __ldg(reinterpret_cast<const float4*>(base + size_t(p) * C));
      v.x -= pv.x; v.y -= pv.y; v.z -= pv.z; v.w -= pv.w;
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] += v.x * v.x; q[1] += v.y * v.y; q[2] += v.z * v.z; q[3] += v.w * v.w;
    }
  }
  __shared__ float sm[kBlock * 8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { sm[threadIdx.x * 8 + i] = s[i]; sm[threadIdx.x * 8 + 4 + i] = q[i]; }
  __syncthreads();
  // tree over ty
  for (int step = rows >> 1; step >= 1; step >>= 1) {
    if (ty < step) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm[threadIdx.x * 8 + i] += sm[(threadIdx.x + step * gx) * 8 + i];
    }
    __syncthreads();
  }
  if (ty == 0 && cvalid) {
    float* dst = partial + ((size_t(n) * nblk + blockIdx.x) * 2) * C + c4 * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) { dst[i] = sm[tx * 8 + i]; dst[C + i] = sm[tx * 8 + 4 + i]; }
  }
}
// 256 threads = 32 (n, c) pairs x 8 lanes over the partial blocks: the nblk (<= 256) dependent L2 loads of the old
// thread-per-channel loop made this trivial kernel take ~18 us on the full-resolution layers.
__global__ void in_stats_finalize_kernel(const float* __restrict__ y, const float* __restrict__ partial, int N, int nblk,
                                         int C, int HW, float eps, float* __restrict__ mean, float* __restrict__ rstd) {
  __shared__ double ss[8][32], sq[8][32];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  const bool ok = i < N * C;
  const int n = ok ? i / C : 0, c = ok ? i % C : 0;
  double s = 0, q = 0;
  if (ok) {
    for (int b = bl; b < nblk; b += 8) {
      const float* p = partial + ((size_t(n) * nblk + b) * 2) * C + c;
      s += p[0]; q += p[C];
    }
  }
  ss[bl][cl] = s; sq[bl][cl] = q;
  __syncthreads();
  if (bl == 0 && ok) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { s += ss[k][cl]; q += sq[k][cl]; }
    const double m = s / HW;   // mean of (x - pivot)
    double var = q / HW - m * m;
    if (var < 0) var = 0;
    mean[i] = float(m + double(__ldg(y + size_t(n) * HW * C + c)));
    rstd[i] = float(1.0 / sqrt(var + double(eps)));
  }
}

// BatchNorm2d in training mode (box2mask): finalize of the batch-folded statistics (one "sample" of P = N * HW pixels) fused
// with what hm_bn_fold does -- per-(n, c) rows for hm_in_apply (rstd' = rstd * gamma, mean' = mean - beta / rstd') and the
// module's running-buffer update (momentum, unbiased variance, `repeat` evaluations) -- one launch instead of two.
__global__ void bn_stats_finalize_kernel(const float* __restrict__ y, const float* __restrict__ partial, int nblk, int C, int P,
                                         float eps, float* __restrict__ mean, float* __restrict__ rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, int N,
                                         float* __restrict__ mean_rows, float* __restrict__ rstd_rows, float* running_mean,
                                         float* running_var, long long* num_batches_tracked, float momentum, int repeat) {
  __shared__ double ss[8][32], sq[8][32];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const bool ok = c < C;
  double s = 0, q = 0;
  if (ok) {
    for (int b = bl; b < nblk; b += 8) {
      const float* p = partial + (size_t(b) * 2) * C + c;
      s += p[0]; q += p[C];
    }
  }
  ss[bl][cl] = s; sq[bl][cl] = q;
  __syncthreads();
  if (bl == 0 && ok) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { s += ss[k][cl]; q += sq[k][cl]; }
    const double m = s / P;                                // mean of (x - pivot)
    double var = q / P - m * m;
    if (var < 0) var = 0;
    const float mu = float(m + double(__ldg(y + c)));
    const float rs = float(1.0 / sqrt(var + double(eps)));
    mean[c] = mu;
    rstd[c] = rs;
    const float rsg = rs * (gamma ? gamma[c] : 1.f);
    const float b = beta ? beta[c] : 0.f;
    const float mg = (rsg != 0.f) ? mu - b / rsg : mu;
    for (int n = 0; n < N; ++n) { rstd_rows[size_t(n) * C + c] = rsg; mean_rows[size_t(n) * C + c] = mg; }
    if (running_mean && running_var) {
      const float var_u = P > 1 ? float(var) * (float(P) / float(P - 1)) : float(var);
      float rm = running_mean[c], rv = running_var[c];
      for (int k = 0; k < repeat; ++k) {
        rm = (1.f - momentum) * rm + momentum * mu;
        rv = (1.f - momentum) * rv + momentum * var_u;
      }
      running_mean[c] = rm;
      running_var[c] = rv;
      if (c == 0 && num_batches_tracked) *num_batches_tracked += repeat;
    }
  }
}

// ================================================================================================
// K7  normalise + activation (+ residual) + operand emission with materialised border
//   out = act((y - mean) * rstd) [+ skip]     (Pix2Pix_NET.py:74-90, layer_util.py:341-378, Discriminator_NET.py:80-90)
// ================================================================================================
// grid (pixel blocks, N, cgroups); thread (tx = 8-channel group, ty = padded-pixel lane); (hp, wp) advance incrementally
__global__ void __launch_bounds__(kBlock) in_apply_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ skip, int N, int H, int W,
                                int C, int act, float slope, float* __restrict__ out32, bf16* o_hi, bf16* o_lo, int o_cs,
                                int border, int reflect, int gx_log2) {
  const int gx = 1 << gx_log2, rows = kBlock >> gx_log2;
  const int tx = threadIdx.x & (gx - 1), ty = threadIdx.x >> gx_log2;
  const int n = blockIdx.y;
  const int c = (blockIdx.z * gx + tx) * 8;
  const int Cout = o_hi ? o_cs : ((C + 7) & ~7);
  if (c >= Cout) return;
  const int Hp = H + 2 * border, Wp = W + 2 * border, HWp = Hp * Wp;
  float mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool ok = mean && (c + j < C);
    mu[j] = ok ? __ldg(mean + n * C + c + j) : 0.f;
    rs[j] = ok ? __ldg(rstd + n * C + c + j) : 1.f;
  }
  int pp = blockIdx.x * rows + ty;
  int hp = pp / Wp, wp = pp - hp * Wp;
  const int step = gridDim.x * rows;
  const int sh = step / Wp, sw = step - sh * Wp;
  for (; pp < HWp; pp += step, hp += sh, wp += sw) {
    if (wp >= Wp) { wp -= Wp; ++hp; }
    int h = hp - border, w = wp - border;
    bool interior = true, zero = false;
    if (h < 0 || h >= H || w < 0 || w >= W) {
      interior = false;
      if (reflect) { h = reflect_idx(h, H); w = reflect_idx(w, W); } else zero = true;
    }
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!zero && c < C) {
      const size_t pix = (size_t(n) * H + h) * W + w;
      load8(y + pix * C, c, C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = act_fwd((v[j] - mu[j]) * rs[j], act, slope);
      if (skip) {
        float sk[8];
        load8(skip + pix * C, c, C, sk);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += sk[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c + j >= C) v[j] = 0.f;
      if (interior && out32) store8(out32 + pix * C, c, C, v);
    }
    if (o_hi) store_op8(o_hi, o_lo, (size_t(n) * HWp + size_t(hp) * Wp + wp) * o_cs + c, v);
  }
}

// ================================================================================================
// K7 backward.  dz = fold_reflect(g1) + g2 + l1coef*sign(z - tref);  dyhat = dz * act'(.)
//   with IN:  dy = rstd * (dyhat - mean(dyhat) - yhat * mean(dyhat*yhat))   else  dy = dyhat
// ================================================================================================
struct BwdArgs {
  const float* y; const float* mean; const float* rstd;   // pre-norm conv output + stats (nullable)
  const float* z;                                          // post-activation fp32 value (taps), nullable
  const bf16* mask_hi; int mask_cs;                        // relu mask source when neither y nor z exist
  const float* g1; int g1_border; int g1_ld; int g1_coff;  // gradient in (reflect-)padded space
  const float* g2;                                         // dense gradient, same dims as y
  const float* tref; float l1coef;                         // L1 feature-matching term
  int N, H, W, C, act; float slope;
  // BatchNorm (box2mask, BWD_GENERIC variants only): out = act(gamma * yhat + beta); the activation mask is taken from
  // gamma * yhat + beta and the result is scaled by gamma:  dy = rstd * gamma * (dact - mean(dact) - yhat * mean(dact * yhat))
  const float* gamma; const float* beta;
};

// ---- V-wide (V = 4 or 8 channels per thread) helpers: V = 4 halves the register footprint of the backward kernels
// (64 regs -> 4 resident blocks/SM), which is what lets them cover HBM latency.
// V == 4 is only launched when every tensor is float4-addressable and C % 4 == 0 (hm_in_bwd checks), so a valid
// thread (c < C) always owns 4 full channels: no per-access bounds / alignment tests in the hot loops.
template <int V>
__device__ __forceinline__ void loadv(const float* __restrict__ p, int c, int C, float (&v)[V]) {
  if constexpr (V == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + c));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
    if (((C & 3) == 0) && c + V <= C) {
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p + c + i));
        v[i] = a.x; v[i + 1] = a.y; v[i + 2] = a.z; v[i + 3] = a.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = (c + i < C) ? __ldg(p + c + i) : 0.f;
    }
  }
}
template <int V>
__device__ __forceinline__ void storev(float* __restrict__ p, int c, int C, const float (&v)[V]) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p + c) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    if (((C & 3) == 0) && c + V <= C) {
#pragma unroll
      for (int i = 0; i < V; i += 4) *reinterpret_cast<float4*>(p + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) if (c + i < C) p[c + i] = v[i];
    }
  }
}
template <int V>
__device__ __forceinline__ void store_opv(bf16* __restrict__ hi, bf16* __restrict__ lo, size_t off, const float (&v)[V]) {
  alignas(16) bf16 h[V];
  alignas(16) bf16 l[V];
#pragma unroll
  for (int i = 0; i < V; ++i) hm::split_bf16(v[i], h[i], l[i]);
  if constexpr (V == 8) {
    *reinterpret_cast<uint4*>(hi + off) = *reinterpret_cast<const uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + off) = *reinterpret_cast<const uint4*>(l);
  } else {
    *reinterpret_cast<uint2*>(hi + off) = *reinterpret_cast<const uint2*>(h);
    if (lo) *reinterpret_cast<uint2*>(lo + off) = *reinterpret_cast<const uint2*>(l);
  }
}

// per-thread constants of one (n, V-channel group): hoisted out of the pixel loops
template <int V>
struct ChanStats { float mu[V], rs[V]; float gneg; /* act'(x <= 0): 0 relu, slope lrelu, 1 none */ };
template <int V>
__device__ __forceinline__ void load_chan_stats(const BwdArgs& a, int n, int c, ChanStats<V>& cs) {
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const bool ok = a.mean && (c + j < a.C);
    cs.mu[j] = ok ? __ldg(a.mean + n * a.C + c + j) : 0.f;
    cs.rs[j] = ok ? __ldg(a.rstd + n * a.C + c + j) : 1.f;
  }
  cs.gneg = (a.act == HM_ACT_RELU) ? 0.f : ((a.act == HM_ACT_LRELU) ? a.slope : 1.f);
}

// Which inputs exist is known on the host, so the hot variants are compiled with the presence flags as template
// constants (F_* mask): with run-time flags the loop body re-tested ~10 pointers and rebuilt every 64-bit address
// per pixel (~160 instructions per 4-channel pixel, issue bound at 25-35 % of HBM bandwidth; ncu r01).  BWD_GENERIC
// keeps the run-time tests for the combinations that are not instantiated.
enum : int { F_IN = 1, F_G1D = 2, F_G1B = 4, F_G2 = 8, F_Z = 16, F_TREF = 32, F_MASK = 64, BWD_GENERIC = 128, F_BN = 256 };
// F_BN: the affine (BatchNorm) form as a compile-time variant: gamma is known to be present
template <int F> __device__ __forceinline__ bool has_in(const BwdArgs& a) { return (F & BWD_GENERIC) ? a.y != nullptr : (F & F_IN) != 0; }
template <int F> __device__ __forceinline__ bool has_g1d(const BwdArgs& a) { return (F & BWD_GENERIC) ? (a.g1 && a.g1_border == 0) : (F & F_G1D) != 0; }
template <int F> __device__ __forceinline__ bool has_g1b(const BwdArgs& a) { return (F & BWD_GENERIC) ? (a.g1 && a.g1_border != 0) : (F & F_G1B) != 0; }
template <int F> __device__ __forceinline__ bool has_g2(const BwdArgs& a) { return (F & BWD_GENERIC) ? a.g2 != nullptr : (F & F_G2) != 0; }
template <int F> __device__ __forceinline__ bool has_z(const BwdArgs& a) { return (F & BWD_GENERIC) ? a.z != nullptr : (F & F_Z) != 0; }
template <int F> __device__ __forceinline__ bool has_tref(const BwdArgs& a) { return (F & BWD_GENERIC) ? a.tref != nullptr : (F & F_TREF) != 0; }
template <int F> __device__ __forceinline__ bool has_mask(const BwdArgs& a) { return (F & BWD_GENERIC) ? a.mask_hi != nullptr : (F & F_MASK) != 0; }

template <int V, int F>
__device__ __forceinline__ void bwd_dyhat(const BwdArgs& a, const ChanStats<V>& cs, size_t pix, int n, int h, int w, int c,
                                          float (&dyh)[V], float (&yh)[V]) {
  float dz[V];
#pragma unroll
  for (int j = 0; j < V; ++j) dz[j] = 0.f;
  if (has_g1d<F>(a)) {
    const float* p = a.g1 + pix * a.g1_ld + a.g1_coff;
    if (V == 4 || (((a.g1_ld & 3) == 0) && ((a.g1_coff & 3) == 0))) loadv<V>(p, c, a.C, dz);
    else {
#pragma unroll
      for (int j = 0; j < V; ++j) dz[j] = (c + j < a.C) ? __ldg(p + c + j) : 0.f;
    }
  } else if (has_g1b<F>(a)) {
    const int b = a.g1_border, Hp = a.H + 2 * b, Wp = a.W + 2 * b;
    const bool vec = V == 4 || (((a.g1_ld & 3) == 0) && ((a.g1_coff & 3) == 0));
    // adjoint of ReflectionPad2d(b): the interior cell plus up to one mirrored border cell per side and axis.
    // Fast path first: only pixels in the 2b rows / columns next to the frame receive mirrored contributions
    // (the generic 3x3 candidate loop cost ~250 instructions per pixel and made these kernels issue bound).
    {
      const float* p = a.g1 + ((size_t(n) * Hp + h + b) * Wp + w + b) * a.g1_ld + a.g1_coff;
      if (vec) loadv<V>(p, c, a.C, dz);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j) dz[j] = (c + j < a.C) ? __ldg(p + c + j) : 0.f;
      }
    }
    const bool near_h = (h >= 1 && h <= b) || (h >= a.H - 1 - b && h <= a.H - 2);
    const bool near_w = (w >= 1 && w <= b) || (w >= a.W - 1 - b && w <= a.W - 2);
    if (near_h || near_w) {
#pragma unroll 1
      for (int ih = 0; ih < 3; ++ih) {
        int hs;
        if (ih == 0) hs = h + b;
        else if (ih == 1) { if (!(h >= 1 && h <= b)) continue; hs = b - h; }
        else { if (!(h >= a.H - 1 - b && h <= a.H - 2)) continue; hs = b + 2 * (a.H - 1) - h; }
#pragma unroll 1
        for (int iw = 0; iw < 3; ++iw) {
          int ws;
          if (iw == 0) { if (ih == 0) continue; ws = w + b; }   // the interior cell is already in dz
          else if (iw == 1) { if (!(w >= 1 && w <= b)) continue; ws = b - w; }
          else { if (!(w >= a.W - 1 - b && w <= a.W - 2)) continue; ws = b + 2 * (a.W - 1) - w; }
          float t[V];
          const float* p = a.g1 + ((size_t(n) * Hp + hs) * Wp + ws) * a.g1_ld + a.g1_coff;
          if (vec) loadv<V>(p, c, a.C, t);
          else {
#pragma unroll
            for (int j = 0; j < V; ++j) t[j] = (c + j < a.C) ? __ldg(p + c + j) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < V; ++j) dz[j] += t[j];
        }
      }
    }
  }
  if (has_g2<F>(a)) {
    float t[V];
    loadv<V>(a.g2 + pix * a.C, c, a.C, t);
#pragma unroll
    for (int j = 0; j < V; ++j) dz[j] += t[j];
  }
  const bool have_z = has_z<F>(a), have_y = has_in<F>(a);
  float src[V];   // anything with the sign of the pre-activation
  if (have_y) {
    loadv<V>(a.y + pix * a.C, c, a.C, yh);
#pragma unroll
    for (int j = 0; j < V; ++j) { yh[j] = (yh[j] - cs.mu[j]) * cs.rs[j]; src[j] = yh[j]; }
    if ((F & F_BN) || ((F & BWD_GENERIC) && a.gamma)) {
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (c + j < a.C) src[j] = yh[j] * __ldg(a.gamma + c + j) + (a.beta ? __ldg(a.beta + c + j) : 0.f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) yh[j] = 0.f;
  }
  if (have_z) {
    float zv[V];
    loadv<V>(a.z + pix * a.C, c, a.C, zv);
    if (has_tref<F>(a)) {
      float t[V];
      loadv<V>(a.tref + pix * a.C, c, a.C, t);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float d = zv[j] - t[j];
        dz[j] += a.l1coef * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      }
    }
    if (!have_y) {
#pragma unroll
      for (int j = 0; j < V; ++j) src[j] = zv[j];
    }
  } else if (has_tref<F>(a)) {
    float t[V];
    loadv<V>(a.tref + pix * a.C, c, a.C, t);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float d = act_fwd(yh[j], a.act, a.slope) - t[j];
      dz[j] += a.l1coef * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
  }
  if (!have_y && !have_z) {
    if (has_mask<F>(a)) {
      const bf16* mh = a.mask_hi + pix * a.mask_cs + c;
      if constexpr (V == 4) {
        const uint2 mv = __ldg(reinterpret_cast<const uint2*>(mh));
        const bf16* m4 = reinterpret_cast<const bf16*>(&mv);
#pragma unroll
        for (int j = 0; j < 4; ++j) src[j] = __bfloat162float(m4[j]);
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) src[j] = __bfloat162float(mh[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) src[j] = 1.f;
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) dyh[j] = (V == 4 || c + j < a.C) ? dz[j] * (src[j] > 0.f ? 1.f : cs.gneg) : 0.f;
}

// grid (nblk, N, cgroups); thread (tx = V-channel group, ty = pixel lane): per-thread channel constants are loaded once
template <int V, int F>
__global__ void __launch_bounds__(kBlock) in_bwd_reduce_kernel(BwdArgs a, int gx_log2, float* __restrict__ partial) {
  const int gx = 1 << gx_log2, rows = kBlock >> gx_log2;
  const int tx = threadIdx.x & (gx - 1), ty = threadIdx.x >> gx_log2;
  const int n = blockIdx.y, nblk = gridDim.x;
  const int c = (blockIdx.z * gx + tx) * V;
  const bool cvalid = c < a.C;
  const int HW = a.H * a.W;
  const int per = (HW + nblk - 1) / nblk;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  float s1[V], s2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (cvalid) {
    ChanStats<V> cs;
    load_chan_stats<V>(a, n, c, cs);
    // (h, w) advance incrementally: an integer division per pixel made these kernels instruction bound
    int p = p0 + ty;
    int h = p / a.W, w = p - h * a.W;
    const int sh = rows / a.W, sw = rows - sh * a.W;
    const size_t base = size_t(n) * HW;
    for (; p < p1; p += rows) {
      float dyh[V], yh[V];
      bwd_dyhat<V, F>(a, cs, base + p, n, h, w, c, dyh, yh);
#pragma unroll
      for (int j = 0; j < V; ++j) { s1[j] += dyh[j]; s2[j] += dyh[j] * yh[j]; }
      h += sh; w += sw;
      if (w >= a.W) { w -= a.W; ++h; }
    }
  }
  __shared__ float sm[kBlock * 2 * V];
#pragma unroll
  for (int j = 0; j < V; ++j) { sm[threadIdx.x * 2 * V + j] = s1[j]; sm[threadIdx.x * 2 * V + V + j] = s2[j]; }
  __syncthreads();
  for (int step = rows >> 1; step >= 1; step >>= 1) {
    if (ty < step) {
#pragma unroll
      for (int j = 0; j < 2 * V; ++j) sm[threadIdx.x * 2 * V + j] += sm[(threadIdx.x + step * gx) * 2 * V + j];
    }
    __syncthreads();
  }
  if (ty == 0 && cvalid) {
    const int C8 = (a.C + 7) & ~7;
    float* dst = partial + ((size_t(n) * nblk + blockIdx.x) * 2) * C8 + c;
#pragma unroll
    for (int j = 0; j < V; ++j) { dst[j] = sm[tx * 2 * V + j]; dst[C8 + j] = sm[tx * 2 * V + V + j]; }
  }
}
// BatchNorm (N == 1): the same sums are d(beta) = sum dact and d(gamma) = sum dact * xhat -- accumulated here when asked
// for (C valid channels), which saves the separate bn_param_grad launch per layer.
__global__ void in_bwd_finalize_kernel(const float* __restrict__ partial, int N, int nblk, int C8, int HW,
                                       float* __restrict__ sums /*[N][2][C8] means*/, float* __restrict__ dgamma = nullptr,
                                       float* __restrict__ dbeta = nullptr, int C = 0) {
  __shared__ double ss[8][32], sq[8][32];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cl;
  const bool ok = i < N * C8;
  const int n = ok ? i / C8 : 0, c = ok ? i % C8 : 0;
  double s = 0, q = 0;
  if (ok) {
    for (int b = bl; b < nblk; b += 8) {
      const float* p = partial + ((size_t(n) * nblk + b) * 2) * C8 + c;
      s += p[0]; q += p[C8];
    }
  }
  ss[bl][cl] = s; sq[bl][cl] = q;
  __syncthreads();
  if (bl == 0 && ok) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { s += ss[k][cl]; q += sq[k][cl]; }
    sums[(size_t(n) * 2) * C8 + c] = float(s / HW);
    sums[(size_t(n) * 2 + 1) * C8 + c] = float(q / HW);
    if (n == 0 && c < C) {
      if (dbeta) dbeta[c] += float(s);
      if (dgamma) dgamma[c] += float(q);
    }
  }
}
// same thread decomposition; grid (pixel blocks, N, cgroups) with a grid-stride loop over the pixels of image n
template <int V, int F>
__global__ void __launch_bounds__(kBlock) in_bwd_apply_kernel(BwdArgs a, int gx_log2, const float* __restrict__ sums,
                                                              bf16* o_hi, bf16* o_lo, int o_cs, float* __restrict__ out32) {
  const int gx = 1 << gx_log2, rows = kBlock >> gx_log2;
  const int tx = threadIdx.x & (gx - 1), ty = threadIdx.x >> gx_log2;
  const int n = blockIdx.y;
  const int c = (blockIdx.z * gx + tx) * V;
  const int Cout = o_hi ? o_cs : ((a.C + 7) & ~7);
  if (c >= Cout) return;
  const int C8 = (a.C + 7) & ~7;
  const int HW = a.H * a.W;
  const bool cvalid = c < a.C;
  ChanStats<V> cs;
  float m1[V], m2[V];
  if (cvalid) {
    load_chan_stats<V>(a, n, c, cs);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const bool ok = sums && (c + j < a.C);
      m1[j] = ok ? __ldg(sums + (size_t(n) * 2) * C8 + c + j) : 0.f;
      m2[j] = ok ? __ldg(sums + (size_t(n) * 2 + 1) * C8 + c + j) : 0.f;
    }
  }
  int p = blockIdx.x * rows + ty;
  int h = p / a.W, w = p - h * a.W;
  const int step = gridDim.x * rows;
  const int sh = step / a.W, sw = step - sh * a.W;
  const size_t base = size_t(n) * HW;
  for (; p < HW; p += step, h += sh, w += sw) {
    if (w >= a.W) { w -= a.W; ++h; }
    float dy[V];
#pragma unroll
    for (int j = 0; j < V; ++j) dy[j] = 0.f;
    const size_t pix = base + p;
    if (cvalid) {
      float dyh[V], yh[V];
      bwd_dyhat<V, F>(a, cs, pix, n, h, w, c, dyh, yh);
      if (sums) {
#pragma unroll
        for (int j = 0; j < V; ++j) dy[j] = (V == 4 || c + j < a.C) ? cs.rs[j] * (dyh[j] - m1[j] - yh[j] * m2[j]) : 0.f;
        if ((F & F_BN) || ((F & BWD_GENERIC) && a.gamma)) {
#pragma unroll
          for (int j = 0; j < V; ++j) if (c + j < a.C) dy[j] *= __ldg(a.gamma + c + j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) dy[j] = dyh[j];
      }
    }
    if (o_hi) store_opv<V>(o_hi, o_lo, pix * o_cs + c, dy);
    if (out32 && cvalid) storev<V>(out32 + pix * a.C, c, a.C, dy);
  }
}

// ================================================================================================
// K8  AvgPool2d(3, stride 2, pad 1, count_include_pad=False)  (Discriminator_NET.py:31-32, Pix2Pix_NET.py:45)
// ================================================================================================
__global__ void avgpool_kernel(const bf16* __restrict__ i_hi, const bf16* __restrict__ i_lo, int N, int H, int W,
                               int cs, int ib, bf16* o_hi, bf16* o_lo, int Ho, int Wo, int ob) {
  // H, W / Ho, Wo are interior extents; the input may carry a materialised border ib (skipped), the output gets a
  // ReflectionPad2d(ob) border (LocalEnhancer feeds the pooled input to a 7x7 reflect-padded stem)
  const int G = cs >> 3;
  const int Hs = H + 2 * ib, Ws = W + 2 * ib, Hop = Ho + 2 * ob, Wop = Wo + 2 * ob;
  const long total = long(N) * Hop * Wop * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % G);
    long r = i / G;
    const int wop = int(r % Wop); r /= Wop;
    const int hop = int(r % Hop);
    const int n = int(r / Hop);
    const int ho = reflect_idx(hop - ob, Ho), wo = reflect_idx(wop - ob, Wo);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (int dh = -1; dh <= 1; ++dh) {
      const int h = 2 * ho + dh;
      if (h < 0 || h >= H) continue;
      for (int dw = -1; dw <= 1; ++dw) {
        const int w = 2 * wo + dw;
        if (w < 0 || w >= W) continue;
        float v[8];
        load_op8(i_hi, i_lo, ((size_t(n) * Hs + h + ib) * Ws + w + ib) * cs + g * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
        ++cnt;
      }
    }
    const float inv = 1.f / float(cnt);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    store_op8(o_hi, o_lo, ((size_t(n) * Hop + hop) * Wop + wop) * cs + g * 8, acc);
  }
}
// adjoint, accumulated into the finer gradient for channels [c0, c1)
__global__ void avgpool_bwd_kernel(const float* __restrict__ gc, int N, int Ho, int Wo, int ldc, float* __restrict__ gf,
                                   int H, int W, int ldf, int c0, int c1) {
  const int nc = c1 - c0;
  const long total = long(N) * H * W * nc;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int c = c0 + int(i % nc);
    long r = i / nc;
    const int w = int(r % W); r /= W;
    const int h = int(r % H);
    const int n = int(r / H);
    float acc = 0.f;
    for (int ho = (h) / 2; ho <= (h + 1) / 2; ++ho) {
      if (ho >= Ho) continue;
      const int ch = min(2 * ho + 1, H - 1) - max(2 * ho - 1, 0) + 1;
      for (int wo = (w) / 2; wo <= (w + 1) / 2; ++wo) {
        if (wo >= Wo) continue;
        const int cw = min(2 * wo + 1, W - 1) - max(2 * wo - 1, 0) + 1;
        acc += __ldg(gc + ((size_t(n) * Ho + ho) * Wo + wo) * ldc + c) / float(ch * cw);
      }
    }
    gf[((size_t(n) * H + h) * W + w) * ldf + c] += acc;
  }
}

// ================================================================================================
// MaxPool2d(2,2) of the VGG19 tower (torchvision features via layer_util.py:384-399) and its adjoint
// ================================================================================================
__global__ void maxpool_kernel(const bf16* __restrict__ i_hi, const bf16* __restrict__ i_lo, int N, int H, int W, int cs,
                               bf16* o_hi, bf16* o_lo) {
  const int Ho = H / 2, Wo = W / 2, G = cs >> 3;
  const long total = long(N) * Ho * Wo * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int g = int(i % G);
    long r = i / G;
    const int wo = int(r % Wo); r /= Wo;
    const int ho = int(r % Ho);
    const int n = int(r / Ho);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int dh = 0; dh < 2; ++dh)
      for (int dw = 0; dw < 2; ++dw) {
        float v[8];
        load_op8(i_hi, i_lo, ((size_t(n) * H + 2 * ho + dh) * W + 2 * wo + dw) * cs + g * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    store_op8(o_hi, o_lo, ((size_t(n) * Ho + ho) * Wo + wo) * cs + g * 8, m);
  }
}
__global__ void maxpool_bwd_kernel(const float* __restrict__ g, int N, int H, int W, int C, const bf16* __restrict__ a_hi,
                                   const bf16* __restrict__ a_lo, int cs, float* __restrict__ dz) {
  // one thread per pooled window x 8 channels; writes the 2x2 window of dz (first maximum wins, as ATen)
  const int Ho = H / 2, Wo = W / 2, G = (C + 7) >> 3;
  const long total = long(N) * Ho * Wo * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int gq = int(i % G);
    long r = i / G;
    const int wo = int(r % Wo); r /= Wo;
    const int ho = int(r % Ho);
    const int n = int(r / Ho);
    const int c = gq * 8;
    float v[4][8];
    for (int k = 0; k < 4; ++k)
      load_op8(a_hi, a_lo, ((size_t(n) * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1)) * cs + c, v[k]);
    float gg[8];
    load8(g + ((size_t(n) * Ho + ho) * Wo + wo) * C, c, C, gg);
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int a = 0; float m = v[0][j];
      for (int k = 1; k < 4; ++k) if (v[k][j] > m) { m = v[k][j]; a = k; }
      arg[j] = a;
    }
    for (int k = 0; k < 4; ++k) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (arg[j] == k) ? gg[j] : 0.f;
      store8(dz + ((size_t(n) * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1)) * C, c, C, o);
    }
  }
}

// ================================================================================================
// K9  loss reductions: acc[slot] += coef * sum|a-b|   /   coef * sum (a-t)^2     (models/losses.py:40-50,75-82;
//     pix2pixHD_condImg_model.py:235-251).  fp32 per-thread, fp64 block + global accumulation.
// ================================================================================================
__global__ void l1_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, double coef, double* acc) {
  float s = 0.f;
  double sd = 0.0;
  int k = 0;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) {
    s += fabsf(__ldg(a + i) - __ldg(b + i));
    if (++k == 64) { sd += s; s = 0.f; k = 0; }
  }
  sd += s;
  __shared__ double sm[kBlock];
  sm[threadIdx.x] = sd;
  __syncthreads();
  for (int st = kBlock / 2; st >= 1; st >>= 1) {
    if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(acc, sm[0] * coef);
}
__global__ void mse_sum_kernel(const float* __restrict__ a, long n, float target, double coef, double* acc) {
  float s = 0.f;
  double sd = 0.0;
  int k = 0;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) {
    const float d = __ldg(a + i) - target;
    s += d * d;
    if (++k == 64) { sd += s; s = 0.f; k = 0; }
  }
  sd += s;
  __shared__ double sm[kBlock];
  sm[threadIdx.x] = sd;
  __syncthreads();
  for (int st = kBlock / 2; st >= 1; st >>= 1) {
    if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(acc, sm[0] * coef);
}
// d/dy of coef*mean((y-t)^2) written as a bf16 operand (dense [P][C] fp32 -> [P][cs])
__global__ void mse_grad_kernel(const float* __restrict__ y, long P, int C, float target, float scale, bf16* o_hi,
                                bf16* o_lo, int cs) {
  const int G = cs >> 3;
  const long total = P * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int c = int(i % G) * 8;
    const long p = i / G;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < C) {
      load8(y + p * C, c, C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (c + j < C) ? scale * (v[j] - target) : 0.f;
    }
    store_op8(o_hi, o_lo, size_t(p) * cs + c, v);
  }
}

// Vanilla GAN (--no_lsgan; GANLoss with nn.BCELoss, models/losses.py:17-20) on the discriminator's raw last-layer output x:
// p = sigmoid(x) is the nn.Sigmoid the reference appends (Discriminator_NET.py:95-96); BCELoss clamps its logs at -100.
__global__ void bce_sum_kernel(const float* __restrict__ x, long n, float target, double coef, double* acc) {
  double sd = 0.0;
  float s = 0.f;
  int k = 0;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) {
    const float p = 1.f / (1.f + expf(-__ldg(x + i)));
    s -= target * fmaxf(logf(p), -100.f) + (1.f - target) * fmaxf(logf(1.f - p), -100.f);
    if (++k == 64) { sd += s; s = 0.f; k = 0; }
  }
  sd += s;
  __shared__ double sm[kBlock];
  sm[threadIdx.x] = sd;
  __syncthreads();
  for (int st = kBlock / 2; st >= 1; st >>= 1) {
    if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(acc, sm[0] * coef);
}
// d/dx of scale * sum BCE(sigmoid(x), t) as a bf16 operand: torch's BCELoss backward (p - t) / max(p (1 - p), 1e-12) times
// the sigmoid's p (1 - p)
__global__ void bce_grad_kernel(const float* __restrict__ x, long P, int C, float target, float scale, bf16* o_hi,
                                bf16* o_lo, int cs) {
  const int G = cs >> 3;
  const long total = P * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int c = int(i % G) * 8;
    const long p = i / G;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < C) {
      load8(x + p * C, c, C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float pr = 1.f / (1.f + expf(-v[j]));
        const float q = pr * (1.f - pr);
        v[j] = (c + j < C) ? scale * (pr - target) * (q / fmaxf(q, 1e-12f)) : 0.f;
      }
    }
    store_op8(o_hi, o_lo, size_t(p) * cs + c, v);
  }
}

// ================================================================================================
// K12  generator output: gate (Pix2Pix_NET.py:96-99), NCHW copy for the caller, D / VGG operand slots
// ================================================================================================
__global__ void finish_fake_kernel(const float* __restrict__ t /*[B,H,W,3] tanh output*/, const float* __restrict__ image,
                                   const float* __restrict__ mask, int use_gate, int B, int H, int W,
                                   float* __restrict__ fake_nchw, bf16* d_hi, bf16* d_lo, int d_cs, int d_coff, bf16* v_hi,
                                   bf16* v_lo, int v_cs, const float* __restrict__ d_mask) {
  const long total = long(B) * H * W;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int w = int(i % W);
    const int h = int((i / W) % H);
    const int n = int(i / (long(W) * H));
    float o[3];
    const float m = use_gate ? __ldg(mask + i) : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __ldg(t + i * 3 + c);
      if (use_gate) {
        const float img = (1.f - m) * __ldg(image + (long(n) * 3 + c) * H * W + long(h) * W + w);  // input[:, -3:] = cond image
        v = (1.f - m) * img + m * v;
      }
      o[c] = v;
      if (fake_nchw) fake_nchw[(long(n) * 3 + c) * H * W + long(h) * W + w] = v;
    }
    if (d_hi) {
      const float dm = d_mask ? __ldg(d_mask + i) : 1.f;   // mask_gan_input (:229-230)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        bf16 hh, ll;
        hm::split_bf16(o[c] * dm, hh, ll);
        d_hi[size_t(i) * d_cs + d_coff + c] = hh;
        if (d_lo) d_lo[size_t(i) * d_cs + d_coff + c] = ll;
      }
    }
    if (v_hi) {
      for (int c0 = 0; c0 < v_cs; c0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c0 + j < 3) ? o[c0 + j] : 0.f;
        store_op8(v_hi, v_lo, size_t(i) * v_cs + c0, v);
      }
    }
  }
}
// d(loss)/d(head pre-activation) = (gD[:, coff:coff+3] + gV[:, :3] + rec term) * gate * (1 - t^2)
__global__ void fake_bwd_kernel(const float* __restrict__ t, const float* __restrict__ mask, int use_gate,
                                const float* __restrict__ gD, int gD_ld, int gD_coff, const float* __restrict__ gV, int gV_ld,
                                const float* __restrict__ real_nchw, float rec_coef, int B, int H, int W, bf16* o_hi,
                                bf16* o_lo, int o_cs, const float* __restrict__ d_mask) {
  const long total = long(B) * H * W;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int w = int(i % W);
    const int h = int((i / W) % H);
    const int n = int(i / (long(W) * H));
    const float m = use_gate ? __ldg(mask + i) : 1.f;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float tv = __ldg(t + i * 3 + c);
      float g = 0.f;
      if (gD) g += __ldg(gD + size_t(i) * gD_ld + gD_coff + c) * (d_mask ? __ldg(d_mask + i) : 1.f);
      if (gV) g += __ldg(gV + size_t(i) * gV_ld + c);
      if (rec_coef != 0.f) {
        const float img = __ldg(real_nchw + (long(n) * 3 + c) * H * W + long(h) * W + w);
        const float fake = use_gate ? ((1.f - m) * (1.f - m) * img + m * tv) : tv;
        const float d = fake - img;
        g += rec_coef * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      }
      v[c] = g * m * (1.f - tv * tv);
    }
    for (int c0 = 0; c0 < o_cs; c0 += 8) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (c0 == 0) ? v[j] : 0.f;
      store_op8(o_hi, o_lo, size_t(i) * o_cs + c0, o);
    }
  }
}

// dense fp32 [P][C] -> bf16 operand [P][cs]  (and the reverse direction is never needed)
__global__ void f32_to_op_kernel(const float* __restrict__ x, long P, int C, int ld, int coff, float scale, bf16* o_hi,
                                 bf16* o_lo, int cs) {
  const int G = cs >> 3;
  const long total = P * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int c = int(i % G) * 8;
    const long p = i / G;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c + j < C) ? scale * __ldg(x + p * ld + coff + c + j) : 0.f;
    store_op8(o_hi, o_lo, size_t(p) * cs + c, v);
  }
}

// per-channel sum over pixels of a dense fp32 [P][C] tensor -> bias gradient (accumulate)
__global__ void colsum_kernel(const float* __restrict__ x, long P, int C, float* __restrict__ out, int accumulate) {
  // grid.x = channel, block reduces over pixels
  const int c = blockIdx.x;
  double s = 0.0;
  float f = 0.f;
  int k = 0;
  for (long p = threadIdx.x; p < P; p += blockDim.x) {
    f += __ldg(x + p * C + c);
    if (++k == 64) { s += f; f = 0.f; k = 0; }
  }
  s += f;
  __shared__ double sm[kBlock];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int st = kBlock / 2; st >= 1; st >>= 1) {
    if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = (accumulate ? out[c] : 0.f) + float(sm[0]);
}
// same, reading a bf16 operand (hi+lo): grid (pixel blocks, cgroups), thread (tx = 8-channel group, ty = pixel lane);
// coalesced 16 B loads, block tree reduction, one fp32 atomicAdd per (block, channel)
__global__ void __launch_bounds__(kBlock) colsum_op_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, long P,
                                                           int C, int cs, int gx_log2, float* __restrict__ out) {
  const int gx = 1 << gx_log2, rows = kBlock >> gx_log2;
  const int tx = threadIdx.x & (gx - 1), ty = threadIdx.x >> gx_log2;
  const int c = (blockIdx.y * gx + tx) * 8;
  const bool cvalid = c < cs;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cvalid) {
    for (long p = blockIdx.x * long(rows) + ty; p < P; p += long(gridDim.x) * rows) {
      float v[8];
      load_op8(hi, lo, size_t(p) * cs + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
  }
  __shared__ float sm[kBlock * 8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[threadIdx.x * 8 + j] = s[j];
  __syncthreads();
  for (int step = rows >> 1; step >= 1; step >>= 1) {
    if (ty < step) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sm[threadIdx.x * 8 + j] += sm[(threadIdx.x + step * gx) * 8 + j];
    }
    __syncthreads();
  }
  if (ty == 0 && cvalid) {
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c + j < C) atomicAdd(out + c + j, sm[tx * 8 + j]);
  }
}

// ================================================================================================
// K10  fused Adam over a flat fp32 parameter segment (torch.optim.Adam semantics, lr/betas of
//      pix2pixHD_condImg_model.py:135,139; no weight decay, no amsgrad)
// ================================================================================================
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, float gscale) {
  const long n4 = n >> 2;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n4; i += long(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = G[j] * gscale;
      M[j] = b1 * M[j] + (1.f - b1) * gr;
      V[j] = b2 * V[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(V[j]) / bc2_sqrt + eps;
      P[j] -= (lr / bc1) * (M[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long i = (n4 << 2) + blockIdx.x * long(blockDim.x) + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) {
    const float gr = g[i] * gscale;
    const float mj = b1 * m[i] + (1.f - b1) * gr;
    const float vj = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mj; v[i] = vj;
    p[i] -= (lr / bc1) * (mj / (sqrtf(vj) / bc2_sqrt + eps));
  }
}

// CUDA-graph friendly variant: the (1-based) step count lives in device memory, so a captured training step replays
// with the right bias corrections (a host-side `step` argument would be frozen into the graph).
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long n, float lr, float b1, float b2, float eps,
                                const int* __restrict__ step_dev, float gscale) {
  const float step = float(__ldg(step_dev));
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float lrc = lr / bc1;
  const long n4 = n >> 2;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n4; i += long(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; const float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = G[j] * gscale;
      M[j] = b1 * M[j] + (1.f - b1) * gr;
      V[j] = b2 * V[j] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(V[j]) / bc2_sqrt + eps;
      P[j] -= lrc * (M[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (long i = (n4 << 2) + blockIdx.x * long(blockDim.x) + threadIdx.x; i < n; i += long(gridDim.x) * blockDim.x) {
    const float gr = g[i] * gscale;
    const float mj = b1 * m[i] + (1.f - b1) * gr;
    const float vj = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mj; v[i] = vj;
    p[i] -= lrc * (mj / (sqrtf(vj) / bc2_sqrt + eps));
  }
}

// out += fold_reflect(g)  (dense [N,H,W,C] += padded [N,H+2b,W+2b,C]); used for the residual skip gradient
__global__ void fold_add_kernel(const float* __restrict__ g, int b, int N, int H, int W, int C, const float* __restrict__ base,
                                float* __restrict__ out) {
  BwdArgs a{};
  a.g1 = g; a.g1_border = b; a.g1_ld = C; a.g1_coff = 0; a.g2 = base;
  a.N = N; a.H = H; a.W = W; a.C = C; a.act = HM_ACT_NONE; a.slope = 0.f;
  const int G = (C + 7) >> 3;
  const long total = long(N) * H * W * G;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int gq = int(i % G);
    long r = i / G;
    const int w = int(r % W); r /= W;
    const int h = int(r % H);
    const int n = int(r / H);
    float dyh[8], yh[8];
    ChanStats<8> cs;
    load_chan_stats<8>(a, n, gq * 8, cs);
    bwd_dyhat<8, BWD_GENERIC>(a, cs, (size_t(n) * H + h) * W + w, n, h, w, gq * 8, dyh, yh);
    store8(out + ((size_t(n) * H + h) * W + w) * C, gq * 8, C, dyh);
  }
}

int stats_geometry(int C, int quad, int* gx_log2, int* cgroups) {
  // quad = channels per thread (4 for stats, 8 for backward)
  const int groups = (C + quad - 1) / quad;
  int gx = 1, l = 0;
  while (gx < groups && gx < 32) { gx <<= 1; ++l; }
  *gx_log2 = l;
  *cgroups = (groups + gx - 1) / gx;
  return 0;
}
inline int stats_nblk(int N, int HW, int cgroups = 1) {
  // enough blocks to fill the chip (~8 per SM: full occupancy at <= 32 registers) while keeping >= 64 pixels per block
  const int want = (148 * 8 + N * cgroups - 1) / std::max(1, N * cgroups);
  int nblk = std::max(1, std::min(want, HW / 64));
  return std::min(nblk, 256);
}
inline int stats_nblk_max(int N, int HW) { return std::max(1, std::min(std::min((148 * 8 + N - 1) / N, HW / 64), 256)); }

}  // namespace

extern "C" {

int hm_encode_input(const float* label, const float* inst, const float* image, const float* mask_in, int B, int H,
                    int W, int label_nc, void* g_hi, void* g_lo, int g_cs, int g_border, void* d_hi, void* d_lo,
                    int d_cs, void* v_hi, void* v_lo, int v_cs, int d_no_imgcond, const float* d_mask, void* stream) {
  if (!label || !image || !mask_in || !g_hi || (g_cs & 7) || (d_hi && (d_cs & 7)) || (v_hi && (v_cs & 7)))
    return HM_ERR_INVALID;
  const int cin = label_nc + (inst ? 1 : 0) + 3;
  if (cin > g_cs || (d_hi && ((d_no_imgcond & 2) ? 3 : cin + ((d_no_imgcond & 1) ? 0 : 3)) > d_cs)) return HM_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EncodeArgs a;
  a.label = label; a.inst = inst; a.image = image; a.mask = mask_in; a.d_mask = d_mask;
  a.B = B; a.H = H; a.W = W; a.label_nc = label_nc; a.d_no_imgcond = d_no_imgcond;
  a.g_hi = static_cast<bf16*>(g_hi); a.g_lo = static_cast<bf16*>(g_lo); a.g_cs = g_cs; a.gb = g_border;
  a.d_hi = static_cast<bf16*>(d_hi); a.d_lo = static_cast<bf16*>(d_lo); a.d_cs = d_cs;
  a.v_hi = static_cast<bf16*>(v_hi); a.v_lo = static_cast<bf16*>(v_lo); a.v_cs = v_cs;
  if (H + 2 * g_border > 65535 || 2 * B > 65535) return HM_ERR_INVALID;
  {
    const int groups = g_cs >> 3, wrow = W + 2 * g_border;
    encode_kernel<0><<<dim3((wrow * groups + kBlock - 1) / kBlock, H + 2 * g_border, B), kBlock, 0, st>>>(a, groups, 1.f / groups);
  }
  if (d_hi) {
    const int groups = d_cs >> 3;
    encode_kernel<1><<<dim3((W * groups + kBlock - 1) / kBlock, H, 2 * B), kBlock, 0, st>>>(a, groups, 1.f / groups);
  }
  if (v_hi) {
    const int groups = v_cs >> 3;
    encode_kernel<2><<<dim3((W * groups + kBlock - 1) / kBlock, H, B), kBlock, 0, st>>>(a, groups, 1.f / groups);
  }
  return HM_LAUNCH_OK();
}

size_t hm_in_ws_bytes(int N, int HW, int C) {
  const int C8 = (C + 7) & ~7;
  return size_t(N) * stats_nblk_max(N, HW) * 2 * C8 * sizeof(float) + size_t(N) * 2 * C8 * sizeof(float);
}

int hm_bn_stats(const float* y, int N, int HW, int C, float eps, float* ws, float* mean, float* rstd, const float* gamma,
                const float* beta, float* mean_rows, float* rstd_rows, float* running_mean, float* running_var,
                long long* num_batches_tracked, float momentum, int repeat, void* stream) {
  if (!y || !ws || !mean || !rstd || !mean_rows || !rstd_rows || (C & 3) || N <= 0 || (!running_mean != !running_var) ||
      long(N) * HW > 0x7fffffffL)
    return HM_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int gx_log2, cgroups;
  stats_geometry(C, 4, &gx_log2, &cgroups);
  const int P = N * HW;                                   // the batch folded into one "sample" of N * HW pixels
  const int nblk = stats_nblk(1, P, cgroups);
  in_stats_kernel<<<dim3(nblk, 1, cgroups), kBlock, 0, st>>>(y, P, C, gx_log2, ws);
  bn_stats_finalize_kernel<<<(C + 31) / 32, 256, 0, st>>>(y, ws, nblk, C, P, eps, mean, rstd, gamma, beta, N, mean_rows,
                                                          rstd_rows, running_mean, running_var, num_batches_tracked,
                                                          momentum, repeat);
  return HM_LAUNCH_OK();
}

int hm_in_stats(const float* y, int N, int HW, int C, float eps, float* ws, float* mean, float* rstd, void* stream) {
  if (!y || !ws || !mean || !rstd || (C & 3)) return HM_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int gx_log2, cgroups;
  stats_geometry(C, 4, &gx_log2, &cgroups);
  const int nblk = stats_nblk(N, HW, cgroups);
  in_stats_kernel<<<dim3(nblk, N, cgroups), kBlock, 0, st>>>(y, HW, C, gx_log2, ws);
  in_stats_finalize_kernel<<<(N * C + 31) / 32, 256, 0, st>>>(y, ws, N, nblk, C, HW, eps, mean, rstd);
  return HM_LAUNCH_OK();
}

int hm_in_apply(const float* y, const float* mean, const float* rstd, const float* skip, int N, int H, int W, int C,
                int act, float slope, float* out32, void* o_hi, void* o_lo, int o_cs, int border, int reflect,
                void* stream) {
  if (!y || (o_hi && ((o_cs & 7) || o_cs < C)) || (!o_hi && !out32)) return HM_ERR_INVALID;
  if (reflect && (border >= H || border >= W)) return HM_ERR_INVALID;
  const int Cout = o_hi ? o_cs : ((C + 7) & ~7);
  int gx_log2, cgroups;
  stats_geometry(Cout, 8, &gx_log2, &cgroups);
  const int rows = kBlock >> gx_log2;
  const int HWp = (H + 2 * border) * (W + 2 * border);
  const int pblocks = std::min((HWp + rows - 1) / rows, std::max(1, (148 * 16) / std::max(1, N * cgroups)));
  in_apply_kernel<<<dim3(pblocks, N, cgroups), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      y, mean, rstd, skip, N, H, W, C, act, slope, out32, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs,
      border, reflect, gx_log2);
  return HM_LAUNCH_OK();
}

}  // extern "C"

namespace {

int in_bwd_impl(const float* y, const float* mean, const float* rstd, const float* z, const void* mask_hi, int mask_cs,
                const float* g1, int g1_border, int g1_ld, int g1_coff, const float* g2, const float* tref, float l1coef,
                int N, int H, int W, int C, int act, float slope, float* ws, void* o_hi, void* o_lo, int o_cs,
                float* out32, const float* gamma, const float* beta, float** sums_out, void* stream,
                float* dgamma = nullptr, float* dbeta = nullptr) {
  if ((!o_hi && !out32) || (o_hi && ((o_cs & 7) || o_cs < C))) return HM_ERR_INVALID;
  if (mean && (!y || !ws)) return HM_ERR_INVALID;
  if (g1 && g1_border > 0 && (g1_border >= H || g1_border >= W)) return HM_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BwdArgs a;
  a.y = y; a.mean = mean; a.rstd = rstd; a.z = z; a.mask_hi = static_cast<const bf16*>(mask_hi); a.mask_cs = mask_cs;
  a.g1 = g1; a.g1_border = g1_border; a.g1_ld = g1_ld; a.g1_coff = g1_coff; a.g2 = g2; a.tref = tref; a.l1coef = l1coef;
  a.N = N; a.H = H; a.W = W; a.C = C; a.act = act; a.slope = slope;
  a.gamma = gamma; a.beta = beta;
  const int C8 = (C + 7) & ~7;
  float* sums = nullptr;
  int gx_log2, cgroups;
  // 4 channels per thread whenever every tensor is float4-addressable (all InstanceNorm layers); 8 otherwise
  const bool v4 = ((C & 3) == 0) && (!o_hi || (o_cs & 3) == 0) && (!mask_hi || (mask_cs & 3) == 0) &&
                  (!g1 || (((g1_ld & 3) == 0) && ((g1_coff & 3) == 0)));
  const int V = v4 ? 4 : 8;
  // presence mask of this call; instantiated combinations run with compile-time flags (the affine / BatchNorm form
  // always takes the generic kernels: -1 matches no instantiated case)
  const int F = (gamma ? F_BN : 0) |
                (y ? F_IN : 0) | ((g1 && g1_border == 0) ? F_G1D : 0) | ((g1 && g1_border != 0) ? F_G1B : 0) |
                (g2 ? F_G2 : 0) | (z ? F_Z : 0) | (tref ? F_TREF : 0) | ((mask_hi && !y && !z) ? F_MASK : 0);
#define HM_BWD_CASES(X)                                                                                      \
  X(F_IN | F_G2) X(F_IN | F_G1D) X(F_IN | F_G1B) X(F_IN | F_G1D | F_Z) X(F_IN | F_G1D | F_Z | F_TREF)          \
  X(F_G1D | F_Z) X(F_G1D | F_Z | F_TREF) X(F_Z | F_TREF) X(F_Z | F_TREF | F_G2) X(F_MASK | F_G2) X(F_MASK)       \
  X(F_BN | F_IN | F_G1D) X(F_BN | F_IN | F_G1D | F_G2) X(F_BN | F_IN | F_G2)
  if (mean) {
    stats_geometry(C, V, &gx_log2, &cgroups);
    const int nblk = stats_nblk(N, H * W, cgroups);
    sums = ws + size_t(N) * nblk * 2 * C8;
    const dim3 grid(nblk, N, cgroups);
    bool done = false;
    if (v4) {
      switch (F) {
#define X(f) case (f): in_bwd_reduce_kernel<4, (f)><<<grid, kBlock, 0, st>>>(a, gx_log2, ws); done = true; break;
        HM_BWD_CASES(X)
#undef X
        default: break;
      }
      if (!done) in_bwd_reduce_kernel<4, BWD_GENERIC><<<grid, kBlock, 0, st>>>(a, gx_log2, ws);
    } else {
      in_bwd_reduce_kernel<8, BWD_GENERIC><<<grid, kBlock, 0, st>>>(a, gx_log2, ws);
    }
    in_bwd_finalize_kernel<<<(N * C8 + 31) / 32, 256, 0, st>>>(ws, N, nblk, C8, H * W, sums, dgamma, dbeta, C);
  }
  const int Cout = o_hi ? o_cs : C8;
  stats_geometry(Cout, V, &gx_log2, &cgroups);
  const int rows = kBlock >> gx_log2;
  int pblocks = std::min((H * W + rows - 1) / rows, std::max(1, (148 * 8) / std::max(1, N * cgroups)));
  {
    const dim3 grid(pblocks, N, cgroups);
    bf16* ohi = static_cast<bf16*>(o_hi);
    bf16* olo = static_cast<bf16*>(o_lo);
    bool done = false;
    if (v4) {
      switch (F) {
#define X(f) case (f): in_bwd_apply_kernel<4, (f)><<<grid, kBlock, 0, st>>>(a, gx_log2, sums, ohi, olo, o_cs, out32); done = true; break;
        HM_BWD_CASES(X)
#undef X
        default: break;
      }
      if (!done) in_bwd_apply_kernel<4, BWD_GENERIC><<<grid, kBlock, 0, st>>>(a, gx_log2, sums, ohi, olo, o_cs, out32);
    } else {
      in_bwd_apply_kernel<8, BWD_GENERIC><<<grid, kBlock, 0, st>>>(a, gx_log2, sums, ohi, olo, o_cs, out32);
    }
  }
#undef HM_BWD_CASES
  if (sums_out) *sums_out = sums;
  return HM_LAUNCH_OK();
}
}  // namespace

extern "C" {

int hm_in_bwd(const float* y, const float* mean, const float* rstd, const float* z, const void* mask_hi, int mask_cs,
              const float* g1, int g1_border, int g1_ld, int g1_coff, const float* g2, const float* tref, float l1coef,
              int N, int H, int W, int C, int act, float slope, float* ws, void* o_hi, void* o_lo, int o_cs,
              float* out32, void* stream) {
  return in_bwd_impl(y, mean, rstd, z, mask_hi, mask_cs, g1, g1_border, g1_ld, g1_coff, g2, tref, l1coef, N, H, W, C, act,
                     slope, ws, o_hi, o_lo, o_cs, out32, nullptr, nullptr, nullptr, stream);
}

/* BatchNorm2d(affine) + activation backward in training mode (box2mask): the tensor is treated as ONE sample of N*H*W
 * pixels (mean / rstd are the batch statistics [C]); dgamma / dbeta are ACCUMULATED. */
int hm_bn_bwd(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* z,
              const void* mask_hi, int mask_cs, const float* g1, const float* g2, int N, int H, int W, int C, int act,
              float slope, float* ws, void* o_hi, void* o_lo, int o_cs, float* out32, float* dgamma, float* dbeta,
              void* stream) {
  if (!y || !mean || !rstd || !gamma || !ws || long(N) * H > 0x7fffffffL) return HM_ERR_INVALID;
  float* sums = nullptr;
  // d(gamma) / d(beta) are accumulated by the finalize launch of the reduction (N == 1: one "sample" of N*H*W pixels)
  int rc = in_bwd_impl(y, mean, rstd, z, mask_hi, mask_cs, g1, 0, C, 0, g2, nullptr, 0.f, 1, N * H, W, C, act, slope, ws,
                       o_hi, o_lo, o_cs, out32, gamma, beta, &sums, stream, dgamma, dbeta);
  if (rc != HM_OK) return rc;
  return HM_LAUNCH_OK();
}

int hm_fold_add(const float* g_padded, int border, int N, int H, int W, int C, const float* base, float* out,
                void* stream) {
  if (!g_padded || !out || (border > 0 && (border >= H || border >= W))) return HM_ERR_INVALID;
  const long total = long(N) * H * W * ((C + 7) >> 3);
  fold_add_kernel<<<grid_for(total), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(g_padded, border, N, H, W, C, base, out);
  return HM_LAUNCH_OK();
}

int hm_avgpool3s2(const void* i_hi, const void* i_lo, int N, int H, int W, int cs, int in_border, void* o_hi, void* o_lo,
                  int out_border, void* stream) {
  if (!i_hi || !o_hi || (cs & 7)) return HM_ERR_INVALID;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  if (out_border >= Ho || out_border >= Wo) return HM_ERR_INVALID;
  const long total = long(N) * (Ho + 2 * out_border) * (Wo + 2 * out_border) * (cs >> 3);
  avgpool_kernel<<<grid_for(total), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(i_hi), static_cast<const bf16*>(i_lo), N, H, W, cs, in_border, static_cast<bf16*>(o_hi),
      static_cast<bf16*>(o_lo), Ho, Wo, out_border);
  return HM_LAUNCH_OK();
}

int hm_avgpool3s2_bwd(const float* g_coarse, int N, int Ho, int Wo, int ld_coarse, float* g_fine, int H, int W,
                      int ld_fine, int c0, int c1, void* stream) {
  if (!g_coarse || !g_fine || c1 <= c0) return HM_ERR_INVALID;
  avgpool_bwd_kernel<<<grid_for(long(N) * H * W * (c1 - c0)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      g_coarse, N, Ho, Wo, ld_coarse, g_fine, H, W, ld_fine, c0, c1);
  return HM_LAUNCH_OK();
}

int hm_maxpool2(const void* i_hi, const void* i_lo, int N, int H, int W, int cs, void* o_hi, void* o_lo, void* stream) {
  if (!i_hi || !o_hi || (cs & 7)) return HM_ERR_INVALID;
  maxpool_kernel<<<grid_for(long(N) * (H / 2) * (W / 2) * (cs >> 3)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(i_hi), static_cast<const bf16*>(i_lo), N, H, W, cs, static_cast<bf16*>(o_hi),
      static_cast<bf16*>(o_lo));
  return HM_LAUNCH_OK();
}

int hm_maxpool2_bwd(const float* g, int N, int H, int W, int C, const void* a_hi, const void* a_lo, int cs, float* dz,
                    void* stream) {
  // odd extents: the last row / column lies outside every 2x2 window (MaxPool2d floors), its gradient is zero -- the
  // kernel only writes the windows, so the caller passes a zero-filled dz then
  if (!g || !a_hi || !dz) return HM_ERR_INVALID;
  maxpool_bwd_kernel<<<grid_for(long(N) * (H / 2) * (W / 2) * ((C + 7) >> 3)), kBlock, 0,
                       static_cast<cudaStream_t>(stream)>>>(g, N, H, W, C, static_cast<const bf16*>(a_hi),
                                                           static_cast<const bf16*>(a_lo), cs, dz);
  return HM_LAUNCH_OK();
}

int hm_l1_sum(const float* a, const float* b, long n, double coef, double* acc, void* stream) {
  if (!a || !b || !acc) return HM_ERR_INVALID;
  l1_sum_kernel<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, coef, acc);
  return HM_LAUNCH_OK();
}
int hm_mse_sum(const float* a, long n, float target, double coef, double* acc, void* stream) {
  if (!a || !acc) return HM_ERR_INVALID;
  mse_sum_kernel<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(a, n, target, coef, acc);
  return HM_LAUNCH_OK();
}
int hm_mse_grad(const float* y, long P, int C, float target, float scale, void* o_hi, void* o_lo, int o_cs, void* stream) {
  if (!y || !o_hi || (o_cs & 7) || o_cs < C) return HM_ERR_INVALID;
  mse_grad_kernel<<<grid_for(P * (o_cs >> 3)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      y, P, C, target, scale, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs);
  return HM_LAUNCH_OK();
}

int hm_bce_sum(const float* x, long n, float target, double coef, double* acc, void* stream) {
  if (!x || !acc) return HM_ERR_INVALID;
  bce_sum_kernel<<<grid_for(n, kBlock, 148 * 8), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(x, n, target, coef, acc);
  return HM_LAUNCH_OK();
}
int hm_bce_grad(const float* x, long P, int C, float target, float scale, void* o_hi, void* o_lo, int o_cs, void* stream) {
  if (!x || !o_hi || (o_cs & 7) || o_cs < C) return HM_ERR_INVALID;
  bce_grad_kernel<<<grid_for(P * (o_cs >> 3)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      x, P, C, target, scale, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs);
  return HM_LAUNCH_OK();
}

int hm_finish_fake(const float* t, const float* image, const float* mask, int use_gate, int B, int H, int W,
                   float* fake_nchw, void* d_hi, void* d_lo, int d_cs, int d_coff, void* v_hi, void* v_lo, int v_cs,
                   const float* d_mask, void* stream) {
  if (!t || (use_gate && (!image || !mask))) return HM_ERR_INVALID;
  finish_fake_kernel<<<grid_for(long(B) * H * W), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      t, image, mask, use_gate, B, H, W, fake_nchw, static_cast<bf16*>(d_hi), static_cast<bf16*>(d_lo), d_cs, d_coff,
      static_cast<bf16*>(v_hi), static_cast<bf16*>(v_lo), v_cs, d_mask);
  return HM_LAUNCH_OK();
}

int hm_fake_bwd(const float* t, const float* mask, int use_gate, const float* gD, int gD_ld, int gD_coff,
                const float* gV, int gV_ld, const float* real_nchw, float rec_coef, int B, int H, int W, void* o_hi,
                void* o_lo, int o_cs, const float* d_mask, void* stream) {
  if (!t || !o_hi || (o_cs & 7) || (use_gate && !mask) || (rec_coef != 0.f && !real_nchw)) return HM_ERR_INVALID;
  fake_bwd_kernel<<<grid_for(long(B) * H * W), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      t, mask, use_gate, gD, gD_ld, gD_coff, gV, gV_ld, real_nchw, rec_coef, B, H, W, static_cast<bf16*>(o_hi),
      static_cast<bf16*>(o_lo), o_cs, d_mask);
  return HM_LAUNCH_OK();
}

int hm_f32_to_operand(const float* x, long P, int C, int ld, int coff, float scale, void* o_hi, void* o_lo, int o_cs,
                      void* stream) {
  if (!x || !o_hi || (o_cs & 7) || o_cs < C) return HM_ERR_INVALID;
  f32_to_op_kernel<<<grid_for(P * (o_cs >> 3)), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      x, P, C, ld, coff, scale, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo), o_cs);
  return HM_LAUNCH_OK();
}

int hm_colsum(const float* x, long P, int C, float* out, int accumulate, void* stream) {
  if (!x || !out) return HM_ERR_INVALID;
  colsum_kernel<<<C, kBlock, 0, static_cast<cudaStream_t>(stream)>>>(x, P, C, out, accumulate);
  return HM_LAUNCH_OK();
}
int hm_colsum_operand(const void* hi, const void* lo, long P, int C, int cs, float* out, int accumulate, void* stream) {
  if (!hi || !out || (cs & 7)) return HM_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!accumulate && cudaMemsetAsync(out, 0, size_t(C) * sizeof(float), st) != cudaSuccess) return HM_ERR_LAUNCH;
  int gx_log2, cgroups;
  stats_geometry(cs, 8, &gx_log2, &cgroups);
  const int rows = kBlock >> gx_log2;
  const long pb = std::max<long>(1, std::min<long>((P + rows - 1) / rows, (148 * 4) / cgroups + 1));
  colsum_op_kernel<<<dim3(unsigned(pb), cgroups), kBlock, 0, st>>>(static_cast<const bf16*>(hi), static_cast<const bf16*>(lo),
                                                                    P, C, cs, gx_log2, out);
  return HM_LAUNCH_OK();
}

int hm_adam_step(float* param, const float* grad, float* m, float* v, long n, float lr, float beta1, float beta2,
                 float eps, int step, float grad_scale, void* stream) {
  if (!param || !grad || !m || !v || step < 1) return HM_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return HM_ERR_INVALID;
  const float bc1 = 1.f - powf(beta1, float(step));
  const float bc2s = sqrtf(1.f - powf(beta2, float(step)));
  adam_kernel<<<grid_for(n / 4 + 1, kBlock, 148 * 16), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, m, v, n, lr, beta1, beta2, eps, bc1, bc2s, grad_scale);
  return HM_LAUNCH_OK();
}

int hm_adam_step_dev(float* param, const float* grad, float* m, float* v, long n, float lr, float beta1, float beta2,
                     float eps, const int* step_dev, float grad_scale, void* stream) {
  if (!param || !grad || !m || !v || !step_dev) return HM_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return HM_ERR_INVALID;
  adam_dev_kernel<<<grid_for(n / 4 + 1, kBlock, 148 * 16), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, m, v, n, lr, beta1, beta2, eps, step_dev, grad_scale);
  return HM_LAUNCH_OK();
}

}  // extern "C"
