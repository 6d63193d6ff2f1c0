// hm_sn.cu -- K13: spectral normalisation of the PatchGAN convolutions (opt-in; reference models/sn_utils.py:11-25
// max_singular_value and :49-72 SNConv2d.W_bar).  One power iteration per discriminator evaluation:
//     a = W^T u0,  v = a / (|a| + eps),  b = W v,  u' = b / (|b| + eps),  sigma = u'^T W v,   W_bar = W / sigma
// with W viewed as [n = Cout][m = Cin*KH*KW] exactly as it is stored (OIHW).  sigma never leaves the device: the
// weight-pack kernel (hm_pack_weight_ex) multiplies by *(1/sigma) while it converts W to the bf16 tap slabs, i.e.
// the normalisation is fused into the weight load and W_bar is never materialised in fp32.
// The reference differentiates THROUGH the power iteration (u0 is a constant, v and u' are functions of W):
//     dL/dW = ( G - <G, W_bar> * (g_b v^T + u0 g_a^T) ) / sigma,        G = dL/dW_bar,
//     g_b = b (1/s_b + eps/s_b^2),  g_v = W^T g_b,  g_a = g_v/s_a - a (a.g_v)/(|a| s_a^2),  s_x = |x| + eps,
// which hm_sn_weight_grad applies in place to the gradient the weight-gradient engine accumulated for W_bar.
// All layers of the discriminator are processed by ONE launch (one CTA per layer; the matrices are <= 8 MB).
#include "../../include/hm_b200.h"
#include "hm_ptx.cuh"

namespace {

constexpr int kThreads = 1024;
constexpr float kEps = 1e-12f;  // sn_utils.py:8

__device__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (kThreads >> 5)) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// out[j] = sum_i W[i][j] * x[i]   (thread per column, coalesced across the warp)
__device__ void matvec_t(const float* __restrict__ W, int n, int m, const float* x, float* out) {
  for (int j = threadIdx.x; j < m; j += kThreads) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc = fmaf(__ldg(W + size_t(i) * m + j), x[i], acc);
    out[j] = acc;
  }
}
// out[i] = sum_j W[i][j] * x[j]   (warp per row)
__device__ void matvec(const float* __restrict__ W, int n, int m, const float* x, float* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < n; i += (kThreads >> 5)) {
    float acc = 0.f;
    for (int j = lane; j < m; j += 32) acc = fmaf(__ldg(W + size_t(i) * m + j), x[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[i] = acc;
  }
}

// stash layout (floats): [0,n) u0 | [n,2n) b | [2n,2n+m) v | 2n+m: sigma, 1/sigma, |a|, |b|
__global__ void __launch_bounds__(kThreads) sn_power_kernel(const hm_sn_layer* __restrict__ layers, int update_u) {
  extern __shared__ float sm[];
  __shared__ float red[33];
  const hm_sn_layer L = layers[blockIdx.x];
  const int n = L.n, m = L.m;
  float* su = sm;          // [n]  u0, later b
  float* sv = sm + n;      // [m]  a, later v
  float* st = L.stash;
  for (int i = threadIdx.x; i < n; i += kThreads) { su[i] = L.u[i]; st[i] = su[i]; }
  __syncthreads();
  matvec_t(L.W, n, m, su, sv);
  __syncthreads();
  float p = 0.f;
  for (int j = threadIdx.x; j < m; j += kThreads) p += sv[j] * sv[j];
  const float na = sqrtf(block_sum(p, red));
  const float inv_a = 1.f / (na + kEps);
  for (int j = threadIdx.x; j < m; j += kThreads) { sv[j] *= inv_a; st[2 * n + j] = sv[j]; }
  __syncthreads();
  matvec(L.W, n, m, sv, su);
  __syncthreads();
  p = 0.f;
  for (int i = threadIdx.x; i < n; i += kThreads) p += su[i] * su[i];
  const float nb2 = block_sum(p, red);
  const float nb = sqrtf(nb2);
  const float inv_b = 1.f / (nb + kEps);
  for (int i = threadIdx.x; i < n; i += kThreads) {
    st[n + i] = su[i];
    if (update_u) L.u[i] = su[i] * inv_b;
  }
  if (threadIdx.x == 0) {
    const float sigma = nb2 * inv_b;  // u'^T (W v) = b.b / (|b| + eps)
    st[2 * n + m + 0] = sigma;
    st[2 * n + m + 1] = 1.f / sigma;
    st[2 * n + m + 2] = na;
    st[2 * n + m + 3] = nb;
  }
}

__global__ void __launch_bounds__(kThreads) sn_grad_kernel(const hm_sn_layer* __restrict__ layers) {
  extern __shared__ float sm[];
  __shared__ float red[33];
  const hm_sn_layer L = layers[blockIdx.x];
  const int n = L.n, m = L.m;
  const float* st = L.stash;
  float* G = L.grad;
  float* gb = sm;            // [n]
  float* su0 = sm + n;       // [n]
  float* sv = sm + 2 * n;    // [m]
  float* ga = sm + 2 * n + m;  // [m]  g_v, then g_a
  const float sigma = st[2 * n + m], inv_sigma = st[2 * n + m + 1], na = st[2 * n + m + 2], nb = st[2 * n + m + 3];
  const float s_a = na + kEps, s_b = nb + kEps;
  const float cb = 1.f / s_b + kEps / (s_b * s_b);
  for (int i = threadIdx.x; i < n; i += kThreads) { gb[i] = st[n + i] * cb; su0[i] = st[i]; }
  for (int j = threadIdx.x; j < m; j += kThreads) sv[j] = st[2 * n + j];
  // c = <G, W> / sigma = <G, W_bar>
  float p = 0.f;
  const size_t total = size_t(n) * m;
  for (size_t e = threadIdx.x; e < total; e += kThreads) p = fmaf(G[e], __ldg(L.W + e), p);
  const float c = block_sum(p, red) * inv_sigma;
  matvec_t(L.W, n, m, gb, ga);  // g_v
  __syncthreads();
  p = 0.f;
  for (int j = threadIdx.x; j < m; j += kThreads) p = fmaf(sv[j] * s_a, ga[j], p);  // a . g_v
  const float t = block_sum(p, red);
  const float k2 = (na > 0.f) ? t / (na * s_a * s_a) : 0.f;
  for (int j = threadIdx.x; j < m; j += kThreads) ga[j] = ga[j] / s_a - sv[j] * s_a * k2;
  __syncthreads();
  for (size_t e = threadIdx.x; e < total; e += kThreads) {
    const int i = int(e / m), j = int(e - size_t(i) * m);
    G[e] = (G[e] - c * (gb[i] * sv[j] + su0[i] * ga[j])) * inv_sigma;
  }
  (void)sigma;
}

}  // namespace

extern "C" {

size_t hm_sn_stash_floats(int n, int m) { return size_t(2) * n + m + 8; }

int hm_sn_power_iteration(const hm_sn_layer* layers_dev, int n_layers, int max_n, int max_m, int update_u, void* stream) {
  if (!layers_dev || n_layers <= 0 || max_n <= 0 || max_m <= 0) return HM_ERR_INVALID;
  const size_t smem = size_t(max_n + max_m) * sizeof(float);
  if (smem > 200 * 1024) return HM_ERR_INVALID;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(sn_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
      return HM_ERR_LAUNCH;
    configured = smem;
  }
  sn_power_kernel<<<n_layers, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(layers_dev, update_u);
  return cudaGetLastError() == cudaSuccess ? HM_OK : HM_ERR_LAUNCH;
}

int hm_sn_weight_grad(const hm_sn_layer* layers_dev, int n_layers, int max_n, int max_m, void* stream) {
  if (!layers_dev || n_layers <= 0 || max_n <= 0 || max_m <= 0) return HM_ERR_INVALID;
  const size_t smem = size_t(2 * max_n + 2 * max_m) * sizeof(float);
  if (smem > 200 * 1024) return HM_ERR_INVALID;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(sn_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
      return HM_ERR_LAUNCH;
    configured = smem;
  }
  sn_grad_kernel<<<n_layers, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(layers_dev);
  return cudaGetLastError() == cudaSuccess ? HM_OK : HM_ERR_LAUNCH;
}

}  // extern "C"
