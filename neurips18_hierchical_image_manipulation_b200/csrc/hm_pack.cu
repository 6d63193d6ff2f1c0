// hm_pack.cu -- weight packing (reference OIHW / IOHW fp32 -> bf16 hi/lo tap slabs) and the inverse
// scatter of weight-gradient workspaces back into the reference layout.
#include "../../include/hm_b200.h"
#include "hm_ptx.cuh"

#include <algorithm>

namespace {

// One thread converts TWO adjacent contraction indices (k, k+1) of one row for ALL taps: the fp32 source is read once
// (the taps of a filter are contiguous in OIHW / IOHW, so the per-tap loop walks the sectors the thread already
// fetched) and each tap plane is written with 4-byte bf16x2 stores that are contiguous across the warp.
__global__ void pack_weight_kernel(const float* __restrict__ src, int rows, int kk, int taps, long s_row, long s_k,
                                   long s_tap, int rows_pad, int k_pad, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo) {
  const int kh2 = k_pad >> 1;
  const long total = long(rows_pad) * kh2;
  const long plane = long(rows_pad) * k_pad;
  for (long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
    const int k = int(i % kh2) * 2;
    const int r = int(i / kh2);
    const bool ok0 = r < rows && k < kk, ok1 = r < rows && k + 1 < kk;
    const float* p0 = src + r * s_row + k * s_k;
    const float* p1 = p0 + s_k;
    const long o = long(r) * k_pad + k;
    for (int t = 0; t < taps; ++t) {
      const float v0 = ok0 ? __ldg(p0 + t * s_tap) : 0.f;
      const float v1 = ok1 ? __ldg(p1 + t * s_tap) : 0.f;
      __nv_bfloat16 h0, l0, h1, l1;
      hm::split_bf16(v0, h0, l0);
      hm::split_bf16(v1, h1, l1);
      __nv_bfloat162 hh; hh.x = h0; hh.y = h1;
      *reinterpret_cast<__nv_bfloat162*>(hi + t * plane + o) = hh;
      if (lo) {
        __nv_bfloat162 ll; ll.x = l0; ll.y = l1;
        *reinterpret_cast<__nv_bfloat162*>(lo + t * plane + o) = ll;
      }
    }
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

// Tiled variant (32 x 32 tile of (row, k), all taps of a chunk staged in shared memory) of the pack for sources whose ROW
// index is the contiguous one (the data-gradient role: rows = ci of W[co][ci][t]): the simple kernel reads 36-byte
// fragments 36 KB apart (ncu r01: 31 % of the HBM copy bandwidth for the whole pack class; r02: 1.48 -> 1.07 ms per step).
// (A tiled UNPACK was tried and measured 2.4x slower than the gather above: the workspace G was just written by the
// weight-gradient engine and is L2 resident, so the 4-byte gathers are cheap.)
constexpr int kTileTaps = 9;

__global__ void __launch_bounds__(256) pack_weight_rowfast_kernel(const float* __restrict__ src, int rows, int kk, int taps,
                                                                   long s_row, long s_k, long s_tap, int rows_pad, int k_pad,
                                                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[kTileTaps][32][33];          // [tap][k][row]
  const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;   // 8 warps
  const long plane = long(rows_pad) * k_pad;
  for (int t0 = 0; t0 < taps; t0 += kTileTaps) {
    const int nt = min(kTileTaps, taps - t0);
    // load: lanes along rows (contiguous-ish in the source: stride s_row), warps along k
    for (int kl = wy; kl < 32; kl += 8) {
      const int r = r0 + lane, k = k0 + kl;
      const bool ok = r < rows && k < kk;
      const float* p = src + r * s_row + k * s_k + long(t0) * s_tap;
      for (int t = 0; t < nt; ++t) tile[t][kl][lane] = ok ? __ldg(p + t * s_tap) : 0.f;
    }
    __syncthreads();
    // store: lanes along k (contiguous in the slab), warps along rows
    for (int rl = wy; rl < 32; rl += 8) {
      const int r = r0 + rl, k = k0 + lane;
      if (r < rows_pad && k < k_pad) {
        for (int t = 0; t < nt; ++t) {
          __nv_bfloat16 h, l;
          hm::split_bf16(tile[t][lane][rl], h, l);
          const long o = long(t0 + t) * plane + long(r) * k_pad + k;
          hi[o] = h;
          if (lo) lo[o] = l;
        }
      }
    }
    __syncthreads();
  }
}

// Both engine roles of one weight tensor from ONE read of the fp32 source W[A][B][taps] (OIHW: A = co, B = ci; IOHW:
// A = ci, B = co):  P1[t][a][b] (rows = A, contraction = B) and P2[t][b][a] (rows = B, contraction = A).  A 32 x 32 x taps
// tile goes through shared memory; every lane converts 8 consecutive contraction indices and writes them with one
// 16-byte store per plane (the two separate packs read the source twice and store 2-4 bytes per lane: 43 % of the HBM
// copy bandwidth for the class in round 2).
__global__ void __launch_bounds__(256) pack_weight_pair_kernel(const float* __restrict__ src, int A, int B, int taps,
                                                                int rp1, int kp1, int rp2, int kp2,
                                                                __nv_bfloat16* __restrict__ p1h, __nv_bfloat16* __restrict__ p1l,
                                                                __nv_bfloat16* __restrict__ p2h, __nv_bfloat16* __restrict__ p2l) {
  // every element is split ONCE while it is staged; the tile holds (hi | lo << 16) words
  constexpr int TS = 32 * 33 + 4;                    // tap stride = 4 banks: the staging stores of different taps spread out
  __shared__ uint32_t tile_[kTileTaps * TS];         // [tap][a][b], rows padded to 33
#define TILE(t, a, b) tile_[(t) * TS + (a) * 33 + (b)]
  const int a0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;   // 8 warps
  const long plane1 = long(rp1) * kp1, plane2 = long(rp2) * kp2;
  for (int t0 = 0; t0 < taps; t0 += kTileTaps) {
    const int nt = min(kTileTaps, taps - t0);
    // stage: lanes along b (stride `taps` floats), the tap loop walks the sectors the warp already fetched
    const bool bok = b0 + lane < B;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int al = wy + 8 * i, a = a0 + al;
      const float* p = src + (long(a) * B + b0 + lane) * taps + t0;
      const bool ok = bok && a < A;
      for (int t = 0; t < nt; ++t) {
        __nv_bfloat16 h, l;
        hm::split_bf16(ok ? __ldg(p + t) : 0.f, h, l);
        TILE(t, al, lane) = uint32_t(__bfloat16_as_ushort(h)) | (uint32_t(__bfloat16_as_ushort(l)) << 16);
      }
    }
    __syncthreads();
    for (int w = threadIdx.x; w < nt * 128; w += 256) {
      const int t = w >> 7, r = (w & 127) >> 2, q = (w & 3) * 8;
      if (a0 + r < rp1 && b0 + q < kp1) {              // P1: row a0 + r, contraction b0 + q .. +7
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = TILE(t, r, q + j);
        const long o = long(t0 + t) * plane1 + long(a0 + r) * kp1 + b0 + q;
        uint4 hh, ll;
        hh.x = __byte_perm(v[0], v[1], 0x5410); hh.y = __byte_perm(v[2], v[3], 0x5410);
        hh.z = __byte_perm(v[4], v[5], 0x5410); hh.w = __byte_perm(v[6], v[7], 0x5410);
        *reinterpret_cast<uint4*>(p1h + o) = hh;
        if (p1l) {
          ll.x = __byte_perm(v[0], v[1], 0x7632); ll.y = __byte_perm(v[2], v[3], 0x7632);
          ll.z = __byte_perm(v[4], v[5], 0x7632); ll.w = __byte_perm(v[6], v[7], 0x7632);
          *reinterpret_cast<uint4*>(p1l + o) = ll;
        }
      }
      if (b0 + r < rp2 && a0 + q < kp2) {              // P2: row b0 + r, contraction a0 + q .. +7
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = TILE(t, q + j, r);
        const long o = long(t0 + t) * plane2 + long(b0 + r) * kp2 + a0 + q;
        uint4 hh, ll;
        hh.x = __byte_perm(v[0], v[1], 0x5410); hh.y = __byte_perm(v[2], v[3], 0x5410);
        hh.z = __byte_perm(v[4], v[5], 0x5410); hh.w = __byte_perm(v[6], v[7], 0x5410);
        *reinterpret_cast<uint4*>(p2h + o) = hh;
        if (p2l) {
          ll.x = __byte_perm(v[0], v[1], 0x7632); ll.y = __byte_perm(v[2], v[3], 0x7632);
          ll.z = __byte_perm(v[4], v[5], 0x7632); ll.w = __byte_perm(v[6], v[7], 0x7632);
          *reinterpret_cast<uint4*>(p2l + o) = ll;
        }
      }
    }
    __syncthreads();
  }
#undef TILE
}

}  // namespace

extern "C" {

int hm_pack_weight_pair(const float* src, int A, int B, int taps, void* p1_hi, void* p1_lo, void* p2_hi, void* p2_lo,
                        void* stream) {
  if (!src || !p1_hi || !p2_hi || A <= 0 || B <= 0 || taps <= 0) return HM_ERR_INVALID;
  const int rp1 = hm_rows_pad(A), kp1 = hm_k_pad(B), rp2 = hm_rows_pad(B), kp2 = hm_k_pad(A);
  dim3 grid((std::max(rp1, kp2) + 31) / 32, (std::max(kp1, rp2) + 31) / 32);
  pack_weight_pair_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, A, B, taps, rp1, kp1, rp2, kp2, static_cast<__nv_bfloat16*>(p1_hi), static_cast<__nv_bfloat16*>(p1_lo),
      static_cast<__nv_bfloat16*>(p2_hi), static_cast<__nv_bfloat16*>(p2_lo));
  return cudaGetLastError() == cudaSuccess ? HM_OK : HM_ERR_LAUNCH;
}

int hm_pack_weight(const float* src, int rows, int k, int taps, long s_row, long s_k, long s_tap, void* dst_hi,
                   void* dst_lo, void* stream) {
  if (!src || !dst_hi || rows <= 0 || k <= 0 || taps <= 0) return HM_ERR_INVALID;
  const int rows_pad = hm_rows_pad(rows), k_pad = hm_k_pad(k);
  if (s_row < s_k && rows >= 32 && k >= 32) {     // row index contiguous in the source: transpose through shared memory
    dim3 grid((rows_pad + 31) / 32, (k_pad + 31) / 32);
    pack_weight_rowfast_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, rows, k, taps, s_row, s_k, s_tap, rows_pad, k_pad, static_cast<__nv_bfloat16*>(dst_hi),
        static_cast<__nv_bfloat16*>(dst_lo));
    return cudaGetLastError() == cudaSuccess ? HM_OK : HM_ERR_LAUNCH;
  }
  const long total = long(rows_pad) * (k_pad / 2);
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
