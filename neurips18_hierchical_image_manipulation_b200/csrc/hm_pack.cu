// hm_pack.cu -- weight packing (reference OIHW / IOHW fp32 -> bf16 hi/lo tap slabs) and the inverse
// scatter of weight-gradient workspaces back into the reference layout.
#include "../../include/hm_b200.h"
#include "hm_ptx.cuh"

namespace {

__global__ void pack_weight_kernel(const float* __restrict__ src, int rows, int kk, int taps, long s_row, long s_k,
                                   long s_tap, int rows_pad, int k_pad, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo) {
  const long total = long(taps) * rows_pad * k_pad;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int k = int(i % k_pad);
    const long rt = i / k_pad;
    const int r = int(rt % rows_pad);
    const int t = int(rt / rows_pad);
    float v = 0.f;
    if (r < rows && k < kk) v = __ldg(src + r * s_row + k * s_k + t * s_tap);
    __nv_bfloat16 h, l;
    hm::split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// G[(t*cp_pad + p)][q] (ld = cq_pad) -> dst[q][p][t]
__global__ void wgrad_unpack_kernel(const float* __restrict__ G, int taps, int cp, int cq, int cp_pad, int cq_pad,
                                    float* __restrict__ dst, int accumulate) {
  const long total = long(cq) * cp * taps;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int t = int(i % taps);
    const long qp = i / taps;
    const int pch = int(qp % cp);
    const int q = int(qp / cp);
    const float v = __ldg(G + (long(t) * cp_pad + pch) * cq_pad + q);
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

}  // namespace

extern "C" {

int hm_pack_weight(const float* src, int rows, int k, int taps, long s_row, long s_k, long s_tap, void* dst_hi,
                   void* dst_lo, void* stream) {
  if (!src || !dst_hi || rows <= 0 || k <= 0 || taps <= 0) return HM_ERR_INVALID;
  const int rows_pad = hm_rows_pad(rows), k_pad = hm_k_pad(k);
  const long total = long(taps) * rows_pad * k_pad;
  const int block = 256;
  const int grid = int(std::min<long>((total + block - 1) / block, 148L * 16));
  pack_weight_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      src, rows, k, taps, s_row, s_k, s_tap, rows_pad, k_pad, static_cast<__nv_bfloat16*>(dst_hi),
      static_cast<__nv_bfloat16*>(dst_lo));
  return cudaGetLastError() == cudaSuccess ? HM_OK : HM_ERR_LAUNCH;
}

int hm_wgrad_unpack(const float* G_ws, int KH, int KW, int cp, int cq, float* dst, int accumulate, void* stream) {
  if (!G_ws || !dst) return HM_ERR_INVALID;
  const int taps = KH * KW;
  const int cp_pad = (cp + 63) / 64 * 64, cq_pad = (cq + 63) / 64 * 64;
  const long total = long(cq) * cp * taps;
  const int block = 256;
  const int grid = int(std::min<long>((total + block - 1) / block, 148L * 16));
  wgrad_unpack_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(G_ws, taps, cp, cq, cp_pad, cq_pad, dst,
                                                                            accumulate);
  return cudaGetLastError() == cudaSuccess ? HM_OK : HM_ERR_LAUNCH;
}

}  // extern "C"
