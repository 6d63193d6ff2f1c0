// hm_engine.cuh -- the two tcgen05 "tap-GEMM" engines every convolution on the mask2image hot path
// is lowered to (forward, data-gradient, weight-gradient; plain, strided and transposed).
//
//   K-engine  (operands K-major)  : out[pixel, co] = sum_{tap} sum_{c} A_tap[pixel, c] * B_tap[co, c]
//        A_tap is a shifted (and optionally strided) TMA box of a bf16 NHWC tensor -- im2col never exists;
//        B_tap is a [Cout_pad x Cin_pad] slab of the packed weights.  Used for fprop / dgrad / convT.
//   MN-engine (operands MN-major) : G[(tap,c_m), c_n]  = sum_{pixel} P_tap[pixel, c_m] * Q[pixel, c_n]
//        both operands are NHWC boxes whose *pixels* are the contraction index.  Used for wgrad.
//
// Both are persistent, warp-specialised kernels: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer
// (+ TMEM owner), warps 2..5 = epilogue (TMEM -> registers -> global).  fp32 accumulators live in
// TMEM and are double-buffered so the epilogue of tile i overlaps the main loop of tile i+1.
// Precision: operands are bf16; "bf16x3" (hi/lo split, 3 MMAs) is expressed purely as extra K-loop
// entries that point at the lo planes, so the same kernel serves the fp32-parity and the bf16 mode.
#pragma once
#include "hm_ptx.cuh"

namespace hm {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2, ACT_TANH = 3 };

constexpr int kMaxEntries = 160;  // 49 taps x 3 split products = 147
constexpr int kEngineThreads = 192;

struct __align__(8) KEntry {
  int8_t a_plane, b_plane;  // which tensor map (0 = hi, 1 = lo)
  int16_t dw, dh;           // A box origin offset (input pixels)
  int16_t pad_;
  int32_t b_row;            // first row of this tap's weight slab
};

struct __align__(64) KParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];
  int n_entries, chunks;  // K loop = entries x (Cin_pad / 64)
  int tiles_w, tiles_h, n_img, n_tiles_n;
  int tw_log2, th;        // M tile = th x tw output pixels, tw*th = 128
  int in_stride;          // A origin = tile origin * in_stride + (dw, dh)
  int a_c, a_lo_c0;       // valid channels of A; channels [0, a_lo_c0) have an all-zero lo plane (see hm_operand.lo_c0)
  int cout;               // valid output channels
  int valid_h, valid_w;   // valid extent of the tile space
  int out_sh, out_sw, out_oh, out_ow;  // tile-space pixel -> output pixel (h*out_sh + out_oh, ...)
  float* o32;  int o32_H, o32_W, o32_C, o32_hoff, o32_woff, o32_coff;
  __nv_bfloat16* ohi; __nv_bfloat16* olo; int o16_H, o16_W, o16_C, o16_hoff, o16_woff, o16_coff;
  const float* bias;
  int act; float slope;
  int* err;
  // stream-K tail of the CTA-pair kernel (hm_engine2.cuh): pair tiles [0, sk_full) are processed whole, the k-steps of
  // the remaining sk_rem tiles are dealt out evenly, sk_per per cluster (0 = plain persistent tile loop)
  int sk_full, sk_rem, sk_per;
  float* sk_ws;           // partial accumulators: [cluster][2 segments][256 rows][256 cols] fp32
  int* sk_cnt;            // arrival counters [sk_rem][2 CTA ranks]: zero at registration, reset by the last arriver
  KEntry entries[kMaxEntries];
};

struct __align__(64) MNParams {
  CUtensorMap tmP[2];  // M-side operand (hi, lo)
  CUtensorMap tmQ[2];  // N-side operand (hi, lo)
  int n_pairs; int8_t pairP[4], pairQ[4];
  int tiles_w, tiles_h, n_img;   // pixel tiling of the base space, 64 pixels per k-tile
  int tw_log2, th;               // tw*th = 64
  int sP, sQ;                    // coordinate multipliers of the two operands
  int m_tapped;                  // 1: taps index the M side, 0: the N side
  int upt_m, upt_n;              // 64-channel units per tap (tapped side) / total (other side)
  int m_units, n_units;          // G is [m_units*64][n_units*64]
  int n_m_tiles, n_n_tiles, splits, ktiles;
  int dwP0, dhP0, dwQ0, dhQ0;    // constant origin offsets of the untapped side
  float* G; int ldG; int use_atomic;
  int* err;
  int16_t tap_dw[64], tap_dh[64];
};

template <int BN>
struct KCfg {
  static constexpr int A_BYTES = 128 * 128;  // 128 pixels x 64 bf16
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int CH = (BN >= 32) ? 32 : 16;
};

// Fused-split variant of the K-engine (bf16x3 only, N tile <= 128): ONE pipeline stage holds the hi AND lo boxes of both
// operands for a (tap, 64-channel chunk) and the issuer fires all three products (hi*hi, lo*hi, hi*lo) from it.  The
// per-product stages of the plain kernel load A_hi and B_hi twice: 3 x (A + B) per tap against 2 x (A + B) here.  The
// narrow-N layers are bound by L2 -> SM delivery (ncu r02: the first PatchGAN layer moves 10.3 GB at the 11 TB/s fabric
// limit) or by the barrier round trip per 4 MMAs, so a third less traffic and 12 MMAs per wait is time saved.
template <int BN>
struct K3Cfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN >= 128) ? 3 : (BN >= 64 ? 4 : (BN >= 32 ? 5 : 6));
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int CH = (BN >= 32) ? 32 : 16;
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == ACT_TANH) return tanhf(v);
  return v;
}


// Epilogue of one TMEM chunk (CH consecutive output channels of this thread's pixel): + bias -> activation ->
// fp32 NHWC store and/or bf16 hi/lo operand store.  The bias vector is fetched with 16 B loads and the activation is
// selected ONCE per chunk: a per-element `if (bias) ... switch (act)` serialised 32 dependent __ldg + branches per
// chunk (~4.7 k cycles per 32 columns measured, ncu r01), which made every full-resolution 64/128-channel layer
// epilogue-bound.  P is KParams or RParams (same output fields).
template <int CH, class P>
__device__ __forceinline__ void epilogue_chunk(const P& p, const uint32_t (&raw)[CH], int cg, bool v32, bool v16,
                                               bool vb, size_t off32, size_t off16) {
  float v[CH];
  const bool full_chunk = (cg + CH <= p.cout);
#pragma unroll
  for (int i = 0; i < CH; ++i) v[i] = __uint_as_float(raw[i]);
  if (p.bias) {
    if (vb && full_chunk) {
#pragma unroll
      for (int i = 0; i < CH; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cg + i));
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i) if (cg + i < p.cout) v[i] += __ldg(p.bias + cg + i);
    }
  }
  if (p.act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (p.act == ACT_LRELU) {
    const float sl = p.slope;
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * sl;
  } else if (p.act == ACT_TANH) {
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = tanhf(v[i]);
  }
  if (p.o32) {
    float* dst = p.o32 + off32 + cg;
    if (v32 && full_chunk) {
#pragma unroll
      for (int i = 0; i < CH; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i) if (cg + i < p.cout) dst[i] = v[i];
    }
  }
  if (p.ohi) {
    __nv_bfloat16 hi[CH], lo[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) split_bf16(v[i], hi[i], lo[i]);
    __nv_bfloat16* dh = p.ohi + off16 + cg;
    __nv_bfloat16* dl = p.olo ? p.olo + off16 + cg : nullptr;
    if (v16 && full_chunk) {
#pragma unroll
      for (int i = 0; i < CH; i += 8) {
        *reinterpret_cast<uint4*>(dh + i) = *reinterpret_cast<const uint4*>(hi + i);
        if (dl) *reinterpret_cast<uint4*>(dl + i) = *reinterpret_cast<const uint4*>(lo + i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (cg + i < p.cout) { dh[i] = hi[i]; if (dl) dl[i] = lo[i]; }
    }
  }
}

// Which 16-channel groups (one K = 16 MMA each) of 64-channel chunk `chunk` carry data: channels >= a_c are zero padding,
// and the lo plane is all zero below a_lo_c0 (one-hot label channels), so those MMAs can be skipped: the 38-channel stem
// issues 7 instead of 12 MMAs per tap in bf16x3.  Used by the TRIM variant of the multi-row engine only (measured, r02:
// the generic K-engine's thin layers are bound by L2 -> SM traffic, not by MMA count, and ANY extra work in the single
// issuing thread's loop slows the full-chunk layers down by 15-50 %, so the other engines keep the plain 4-MMA loop).
__device__ __forceinline__ void k_groups(int a_c, int a_lo_c0, int a_plane, int chunk, int& j0, int& j1) {
  const int cv = min(64, a_c - chunk * 64);
  j1 = (cv + 15) >> 4;
  j0 = a_plane ? (max(0, min(a_lo_c0 - chunk * 64, 64)) >> 4) : 0;
}

// ------------------------------------------------------------------------------------------------
// K-engine
// ------------------------------------------------------------------------------------------------
template <bool B, class T, class F> struct hm_cond { using type = T; };
template <class T, class F> struct hm_cond<false, T, F> { using type = F; };

template <int BN, bool FUSED3 = false>
__global__ void __launch_bounds__(kEngineThreads, 1) hm_kgemm_kernel(const __grid_constant__ KParams p) {
  using C = typename hm_cond<FUSED3, K3Cfg<BN>, KCfg<BN>>::type;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                      // [STAGES]
  uint64_t* empty = bars + C::STAGES;         // [STAGES]
  uint64_t* tfull = bars + 2 * C::STAGES;     // [ACC]
  uint64_t* tempty = tfull + C::ACC;          // [ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int num_m_tiles = p.tiles_w * p.tiles_h * p.n_img;
  const int num_tiles = num_m_tiles * p.n_tiles_n;
  const int ksteps = p.n_entries * p.chunks;
  AbortCtl ab{abort_flag, p.err};

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < C::ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (warp-wide loop, one elected lane issues) =====================
    if (elect_one_sync()) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB[0]); }
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles_n;
      int mt = tile / p.n_tiles_n;
      const int twi = mt % p.tiles_w; mt /= p.tiles_w;
      const int thi = mt % p.tiles_h;
      const int n = mt / p.tiles_h;
      const int w0 = (twi << p.tw_log2) * p.in_stride;
      const int h0 = thi * p.th * p.in_stride;
      for (int e = 0; e < p.n_entries; ++e) {
        const KEntry en = p.entries[e];
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(&empty[s], ph ^ 1, ab, 101);
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
            if constexpr (FUSED3) {     // [A_hi | A_lo | B_hi | B_lo] of this (tap, chunk)
              tma_load_4d(&p.tmA[0], &full[s], sa, c * 64, w0 + en.dw, h0 + en.dh, n);
              tma_load_4d(&p.tmA[1], &full[s], sa + C::A_BYTES, c * 64, w0 + en.dw, h0 + en.dh, n);
              tma_load_2d(&p.tmB[0], &full[s], sa + 2 * C::A_BYTES, c * 64, en.b_row + nt * BN);
              tma_load_2d(&p.tmB[1], &full[s], sa + 2 * C::A_BYTES + C::B_BYTES, c * 64, en.b_row + nt * BN);
            } else {
              tma_load_4d(&p.tmA[en.a_plane], &full[s], sa, c * 64, w0 + en.dw, h0 + en.dh, n);
              tma_load_2d(&p.tmB[en.b_plane], &full[s], sa + C::A_BYTES, c * 64, en.b_row + nt * BN);
            }
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-wide loop, one elected lane issues) =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    int s = 0; uint32_t ph = 0; int a = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[a], aph ^ 1, ab, 102);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * BN;
      for (int k = 0; k < ksteps; ++k) {
        mbar_wait(&full[s], ph, ab, 103);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
        if constexpr (FUSED3) {
          const uint64_t ah = umma_smem_desc(sa, 16, 1024), al = umma_smem_desc(sa + C::A_BYTES, 16, 1024);
          const uint64_t bh = umma_smem_desc(sa + 2 * C::A_BYTES, 16, 1024);
          const uint64_t bl = umma_smem_desc(sa + 2 * C::A_BYTES + C::B_BYTES, 16, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, ah + 2 * j, bh + 2 * j, idesc, (k | j) != 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, al + 2 * j, bh + 2 * j, idesc, 1u);
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, ah + 2 * j, bl + 2 * j, idesc, 1u);
            umma_commit(&empty[s]);
          }
        } else {
          const uint64_t adesc = umma_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = umma_smem_desc(sa + C::A_BYTES, 16, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < 4; ++j)  // 4 x (K = 16 bf16 = 32 B) inside the 128 B swizzle row
              umma_bf16(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (k | j) != 0);
            umma_commit(&empty[s]);
          }
        }
        if (++s == C::STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one_sync()) umma_commit(&tfull[a]);
      if (++a == C::ACC) { a = 0; aph ^= 1; }
    }
  } else {
    // ===================== epilogue (4 warps, one TMEM sub-partition each) =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;  // accumulator row == TMEM lane == pixel inside the tile
    int a = 0; uint32_t aph = 0;
    const bool v32 = p.o32 && ((p.o32_C & 3) == 0) && ((p.o32_coff & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.o32) & 15) == 0);
    const bool v16 = p.ohi && ((p.o16_C & 7) == 0) && ((p.o16_coff & 7) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.ohi) & 15) == 0) &&
                     (!p.olo || (reinterpret_cast<uintptr_t>(p.olo) & 15) == 0);
    const bool vb = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles_n;
      int mt = tile / p.n_tiles_n;
      const int twi = mt % p.tiles_w; mt /= p.tiles_w;
      const int thi = mt % p.tiles_h;
      const int n = mt / p.tiles_h;
      const int ht = thi * p.th + (m >> p.tw_log2);
      const int wt = (twi << p.tw_log2) + (m & ((1 << p.tw_log2) - 1));
      const bool valid = (ht < p.valid_h) && (wt < p.valid_w);
      const int oh = ht * p.out_sh + p.out_oh, ow = wt * p.out_sw + p.out_ow;
      size_t off32 = 0, off16 = 0;
      if (p.o32) off32 = ((size_t(n) * p.o32_H + oh + p.o32_hoff) * p.o32_W + ow + p.o32_woff) * p.o32_C + p.o32_coff;
      if (p.ohi) off16 = ((size_t(n) * p.o16_H + oh + p.o16_hoff) * p.o16_W + ow + p.o16_woff) * p.o16_C + p.o16_coff;

      mbar_wait(&tfull[a], aph, ab, 104);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += C::CH) {
        uint32_t raw[C::CH];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + a * BN + c0;
        if constexpr (C::CH == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
        tmem_ld_wait();
        const int cg = nt * BN + c0;  // first output channel of this chunk
        if (valid && cg < p.cout) epilogue_chunk<C::CH>(p, raw, cg, v32, v16, vb, off32, off16);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      if (++a == C::ACC) { a = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// MN-engine (weight gradients).  M tile = 2 boxes x 64 channels (P side), N tile = NB boxes x 64 (Q side),
// K = pixels in tiles of 64 (th x tw rectangle of the base space).
// ------------------------------------------------------------------------------------------------
template <int NB>
struct MNCfg {
  static constexpr int BN = NB * 64;
  static constexpr int BOX_BYTES = 64 * 128;  // 64 pixels x 64 bf16
  static constexpr int A_BYTES = 2 * BOX_BYTES;
  static constexpr int B_BYTES = NB * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (NB >= 4) ? 4 : (NB >= 2 ? 6 : 8);
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// fused-split variant (see K3Cfg): one stage = hi + lo boxes of both operands, three products per k-tile
template <int NB>
struct MN3Cfg {
  static constexpr int BN = NB * 64;
  static constexpr int BOX_BYTES = 64 * 128;
  static constexpr int A_BYTES = 2 * BOX_BYTES;      // per plane
  static constexpr int B_BYTES = NB * BOX_BYTES;     // per plane
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (NB >= 2) ? 3 : 4;
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int NB, bool FUSED3 = false>
__global__ void __launch_bounds__(kEngineThreads, 1) hm_mngemm_kernel(const __grid_constant__ MNParams p) {
  using C = typename hm_cond<FUSED3, MN3Cfg<NB>, MNCfg<NB>>::type;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + C::ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int num_tiles = p.n_m_tiles * p.n_n_tiles * p.splits;
  AbortCtl ab{abort_flag, p.err};

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < C::ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (split, mt, nt); split slowest so that concurrently running CTAs share operand boxes in L2
  auto k_range = [&](int split, int& k0, int& k1) {
    const int per = (p.ktiles + p.splits - 1) / p.splits;
    k0 = split * per; k1 = min(p.ktiles, k0 + per);
    if (k1 < k0) k1 = k0;
  };

  if (warp == 0) {
    {
      if (elect_one_sync()) { tma_prefetch_desc(&p.tmP[0]); tma_prefetch_desc(&p.tmQ[0]); }
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_n_tiles;
        const int mt = (tile / p.n_n_tiles) % p.n_m_tiles;
        const int split = tile / (p.n_n_tiles * p.n_m_tiles);
        int k0, k1; k_range(split, k0, k1);
        // per-box channel offsets and pixel offsets
        int mc[2], mdw[2], mdh[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          int u = min(mt * 2 + r, p.m_units - 1);
          if (p.m_tapped) { int t = u / p.upt_m; mc[r] = (u % p.upt_m) * 64; mdw[r] = p.tap_dw[t]; mdh[r] = p.tap_dh[t]; }
          else { mc[r] = u * 64; mdw[r] = p.dwP0; mdh[r] = p.dhP0; }
        }
        int nc[NB], ndw[NB], ndh[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) {
          int u = min(nt * NB + r, p.n_units - 1);
          if (!p.m_tapped) { int t = u / p.upt_n; nc[r] = (u % p.upt_n) * 64; ndw[r] = p.tap_dw[t]; ndh[r] = p.tap_dh[t]; }
          else { nc[r] = u * 64; ndw[r] = p.dwQ0; ndh[r] = p.dhQ0; }
        }
        for (int pr = 0; pr < (FUSED3 ? 1 : p.n_pairs); ++pr) {
          const CUtensorMap* mp = &p.tmP[p.pairP[pr]];
          const CUtensorMap* mq = &p.tmQ[p.pairQ[pr]];
          for (int kt = k0; kt < k1; ++kt) {
            int t = kt;
            const int twi = t % p.tiles_w; t /= p.tiles_w;
            const int thi = t % p.tiles_h;
            const int n = t / p.tiles_h;
            const int w0 = twi << p.tw_log2, h0 = thi * p.th;
            mbar_wait(&empty[s], ph ^ 1, ab, 201);
            uint8_t* sa = smem + s * C::STAGE_BYTES;
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&full[s], C::STAGE_BYTES);
              if constexpr (FUSED3) {   // [P_hi | P_lo | Q_hi | Q_lo]
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                  for (int r = 0; r < 2; ++r)
                    tma_load_4d(&p.tmP[pl], &full[s], sa + pl * C::A_BYTES + r * C::BOX_BYTES, mc[r], w0 * p.sP + mdw[r],
                                h0 * p.sP + mdh[r], n);
#pragma unroll
                  for (int r = 0; r < NB; ++r)
                    tma_load_4d(&p.tmQ[pl], &full[s], sa + 2 * C::A_BYTES + pl * C::B_BYTES + r * C::BOX_BYTES, nc[r],
                                w0 * p.sQ + ndw[r], h0 * p.sQ + ndh[r], n);
                }
              } else {
#pragma unroll
                for (int r = 0; r < 2; ++r)
                  tma_load_4d(mp, &full[s], sa + r * C::BOX_BYTES, mc[r], w0 * p.sP + mdw[r], h0 * p.sP + mdh[r], n);
#pragma unroll
                for (int r = 0; r < NB; ++r)
                  tma_load_4d(mq, &full[s], sa + C::A_BYTES + r * C::BOX_BYTES, nc[r], w0 * p.sQ + ndw[r],
                              h0 * p.sQ + ndh[r], n);
              }
            }
            if (++s == C::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);
      int s = 0; uint32_t ph = 0; int a = 0; uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile / (p.n_n_tiles * p.n_m_tiles);
        int k0, k1; k_range(split, k0, k1);
        const int ksteps = (k1 - k0) * (FUSED3 ? 1 : p.n_pairs);
        mbar_wait(&tempty[a], aph ^ 1, ab, 202);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int k = 0; k < ksteps; ++k) {
          mbar_wait(&full[s], ph, ab, 203);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
          // MN-major SW128: LBO = distance between 64-channel groups (one TMA box), SBO = 8 K-rows = 1024 B
          if constexpr (FUSED3) {
            const uint64_t ph_ = umma_smem_desc(sa, C::BOX_BYTES, 1024), pl_ = umma_smem_desc(sa + C::A_BYTES, C::BOX_BYTES, 1024);
            const uint64_t qh_ = umma_smem_desc(sa + 2 * C::A_BYTES, C::BOX_BYTES, 1024);
            const uint64_t ql_ = umma_smem_desc(sa + 2 * C::A_BYTES + C::B_BYTES, C::BOX_BYTES, 1024);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, ph_ + j * (2048 >> 4), qh_ + j * (2048 >> 4), idesc, (k | j) != 0);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, pl_ + j * (2048 >> 4), qh_ + j * (2048 >> 4), idesc, 1u);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, ph_ + j * (2048 >> 4), ql_ + j * (2048 >> 4), idesc, 1u);
              umma_commit(&empty[s]);
            }
          } else {
            const uint64_t adesc = umma_smem_desc(sa, C::BOX_BYTES, 1024);
            const uint64_t bdesc = umma_smem_desc(sa + C::A_BYTES, C::BOX_BYTES, 1024);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 4; ++j)  // 16 pixels (K) per MMA = 16 rows x 128 B = 2048 B
                umma_bf16(d_tmem, adesc + j * (2048 >> 4), bdesc + j * (2048 >> 4), idesc, (k | j) != 0);
              umma_commit(&empty[s]);
            }
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one_sync()) umma_commit(&tfull[a]);
        if (++a == C::ACC) { a = 0; aph ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int a = 0; uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_n_tiles;
      const int mt = (tile / p.n_n_tiles) % p.n_m_tiles;
      const int split = tile / (p.n_n_tiles * p.n_m_tiles);
      int k0, k1; k_range(split, k0, k1);
      const int u = mt * 2 + (m >> 6);
      const bool row_ok = (u < p.m_units) && (k1 > k0);
      float* grow = p.G + size_t(u * 64 + (m & 63)) * p.ldG;
      const int ncols = p.n_units * 64;
      mbar_wait(&tfull[a], aph, ab, 204);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + a * BN + c0, raw);
        tmem_ld_wait();
        const int cg = nt * BN + c0;
        if (row_ok && cg < ncols) {
          if (p.use_atomic) {
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(grow + cg + i, __uint_as_float(raw[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(grow + cg + i) =
                  make_float4(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]), __uint_as_float(raw[i + 2]),
                              __uint_as_float(raw[i + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      if (++a == C::ACC) { a = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace hm
