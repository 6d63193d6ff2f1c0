// hm_engine_rows.cuh -- row-streaming variant of the K-engine for stride-1 convolutions on wide images.
//
// Why: with a narrow N tile (Cout <= 128) one 64-deep k-step of the generic K-engine is only 32..64 tensor cycles
// of work per MMA but costs the single producer / issuer threads an mbarrier round trip, two TMA issues and a commit
// (~700 cycles measured per tap, independent of the tile width): the full-resolution, few-channel layers (7x7 stem
// and head, VGG conv1_x, the D layer-0 gradients) were issue-latency bound at 5-20 % tensor-pipe utilisation, and
// they re-read every input pixel KH*KW times.
// Here the pipeline stage is a whole FILTER ROW: the M tile is one output-row segment of 128 pixels; per filter row
// and 64-channel chunk the producer issues exactly two TMA loads -- ONE box of 128+KW-1 input pixels and ONE 3-D box
// holding the KW weight slabs of that row -- and the issuer fires KW x 4 tcgen05.mma per barrier wait.  Each tap reads
// its shifted 128-pixel window straight out of the row box by offsetting the UMMA shared-memory descriptor by whole
// 128-byte pixel rows (the 128B swizzle is a function of the shared-memory address bits, so a window that starts
// inside a swizzle atom stays consistent with what TMA wrote; verified on hardware, descriptor base_offset = 0).
// bf16x3 is three row stages per filter row (hi*hi, lo*hi, hi*lo), exactly like the generic engine's entries.
#pragma once
#include "hm_engine.cuh"

namespace hm {

constexpr int kRowsMaxKW = 9;
constexpr int kRowsMaxEntries = 32;   // KH (<= 9) x 3 products

struct __align__(8) REntry {
  int8_t a_plane, b_plane;
  int16_t dh;          // input row = output row + dh
  int32_t tap0;        // first weight slab (tap index) of this filter row
  int8_t a_off[kRowsMaxKW + 3];  // window start of tile j inside the row box, in pixels
};

struct __align__(64) RParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];     // 3-D: {k_pad, rows_pad, taps}
  int n_entries, chunks;
  int a_c, a_lo_c0;       // valid channels of A; channels [0, a_lo_c0) have an all-zero lo plane (hm_operand.lo_c0)
  int kw;                 // taps per filter row (tiles per B box)
  int n_stages;           // pipeline depth (runtime: depends on kw and BN)
  int stage_bytes;        // A_PLANE + kw * BN * 128
  int tiles_w, rows_h, n_img, n_tiles_n;   // M tiles = tiles_w * rows_h * n_img
  int dw0;                // row-box origin = tile origin + dw0
  int box_w;              // pixels per row box (128 + kw - 1)
  int cout;
  int valid_w;
  float* o32;  int o32_H, o32_W, o32_C, o32_hoff, o32_woff, o32_coff;
  __nv_bfloat16* ohi; __nv_bfloat16* olo; int o16_H, o16_W, o16_C, o16_hoff, o16_woff, o16_coff;
  const float* bias;
  int act; float slope;
  int* err;
  REntry entries[kRowsMaxEntries];
};

struct RCfgCommon {
  static constexpr int A_PLANE = 136 * 128;   // up to 136 pixel rows of 128 B (1024-aligned)
  static constexpr int MAX_STAGES = 8;
  static constexpr int SMEM_BUDGET = 225 * 1024;
};

template <int BN>
struct RCfg : RCfgCommon {
  static constexpr int B_TILE = BN * 128;
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int CH = (BN >= 32) ? 32 : 16;
};

template <int BN>
__global__ void __launch_bounds__(kEngineThreads, 1) hm_krows_kernel(const __grid_constant__ RParams p) {
  using C = RCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.n_stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = full + C::MAX_STAGES;
  uint64_t* tfull = empty + C::MAX_STAGES;
  uint64_t* tempty = tfull + C::ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int num_m_tiles = p.tiles_w * p.rows_h * p.n_img;
  const int num_tiles = num_m_tiles * p.n_tiles_n;
  const int nst = p.n_stages;
  AbortCtl ab{abort_flag, p.err};

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < C::ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: 2 loads per filter-row stage =====================
    {
      if (elect_one_sync()) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB[0]); }
      int s = 0; uint32_t ph = 0;
      const uint32_t tx_bytes = uint32_t(p.box_w) * 128u + uint32_t(p.kw) * C::B_TILE;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int mt = tile / p.n_tiles_n;
        const int twi = mt % p.tiles_w; mt /= p.tiles_w;
        const int h = mt % p.rows_h;
        const int n = mt / p.rows_h;
        const int w0 = twi * 128 + p.dw0;
        for (int c = 0; c < p.chunks; ++c) {
          for (int e = 0; e < p.n_entries; ++e) {
            const int a_plane = p.entries[e].a_plane, b_plane = p.entries[e].b_plane;
            const int dh = p.entries[e].dh, tap0 = p.entries[e].tap0;
            mbar_wait(&empty[s], ph ^ 1, ab, 301);
            uint8_t* dst = smem + s * p.stage_bytes;
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&full[s], tx_bytes);
              tma_load_4d(&p.tmA[a_plane], &full[s], dst, c * 64, w0, h + dh, n);
              tma_load_3d(&p.tmB[b_plane], &full[s], dst + C::A_PLANE, c * 64, nt * BN, tap0);
            }
            if (++s == nst) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: kw x 4 MMAs per barrier wait =====================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
      int s = 0; uint32_t ph = 0; int a = 0; uint32_t aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[a], aph ^ 1, ab, 303);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        uint32_t acc = 0;
        for (int c = 0; c < p.chunks; ++c) {
          for (int e = 0; e < p.n_entries; ++e) {
            mbar_wait(&full[s], ph, ab, 304);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem + s * p.stage_bytes);
            const uint32_t b_base = a_base + C::A_PLANE;
            if (elect_one_sync()) {
              uint32_t acc_j = acc;
              for (int j = 0; j < p.kw; ++j) {
                const uint64_t ad = umma_smem_desc(a_base + uint32_t(p.entries[e].a_off[j]) * 128u, 16, 1024);
                const uint64_t bd = umma_smem_desc(b_base + uint32_t(j) * C::B_TILE, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) { umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, acc_j); acc_j = 1; }
              }
              umma_commit(&empty[s]);
            }
            acc = 1;
            if (++s == nst) { s = 0; ph ^= 1; }
          }
        }
        if (elect_one_sync()) umma_commit(&tfull[a]);
        if (++a == C::ACC) { a = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
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
      const int oh = mt % p.rows_h;
      const int n = mt / p.rows_h;
      const int ow = twi * 128 + m;
      const bool valid = ow < p.valid_w;
      size_t off32 = 0, off16 = 0;
      if (p.o32) off32 = ((size_t(n) * p.o32_H + oh + p.o32_hoff) * p.o32_W + ow + p.o32_woff) * p.o32_C + p.o32_coff;
      if (p.ohi) off16 = ((size_t(n) * p.o16_H + oh + p.o16_hoff) * p.o16_W + ow + p.o16_woff) * p.o16_C + p.o16_coff;
      mbar_wait(&tfull[a], aph, ab, 306);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += C::CH) {
        uint32_t raw[C::CH];
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + a * BN + c0;
        if constexpr (C::CH == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
        tmem_ld_wait();
        const int cg = nt * BN + c0;
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
// Multi-row variant: one CTA tile = MT consecutive output rows x 128 pixels, with SEPARATE shared-memory rings for the
// weight slabs of a filter row (B, 2 stages) and the input row boxes (A).  Why: in hm_krows_kernel every pipeline stage
// re-loads the KW weight slabs of its filter row (57 KB of the 74 KB stage for the 7x7 stem) -- ncu (r01) shows the
// stem at 11.9 TB/s of L2 -> SM traffic, i.e. bound by the L2 fabric, not by the tensor pipe (53 %).  Here the slabs of
// one (filter row, product) are loaded once per MT output rows and the MT row boxes stream through the A ring; the MT
// accumulators live side by side in TMEM (MT x BN columns, double buffered).
// ------------------------------------------------------------------------------------------------
template <int BN>
struct R2Cfg : RCfgCommon {
  static constexpr int MT = (BN <= 64) ? 4 : 2;
  static constexpr int B_TILE = BN * 128;
  static constexpr int B_STAGES = 2;
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = (2 * MT * BN < 32) ? 32 : 2 * MT * BN;
  static constexpr int CH = (BN >= 32) ? 32 : 16;
  static constexpr int MAX_A_STAGES = 8;
};

template <int BN, bool TRIM>
__global__ void __launch_bounds__(kEngineThreads, 1) hm_krows2_kernel(const __grid_constant__ RParams p) {
  using C = R2Cfg<BN>;
  constexpr int MT = C::MT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_stage_bytes = p.kw * C::B_TILE;          // multiple of 1024 (BN >= 16 -> B_TILE >= 2048)
  const int na = p.n_stages;                           // A ring depth (host: what fits next to the B ring)
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + C::B_STAGES * b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + na * C::A_PLANE);
  uint64_t* afull = bars;
  uint64_t* aempty = afull + C::MAX_A_STAGES;
  uint64_t* bfull = aempty + C::MAX_A_STAGES;
  uint64_t* bempty = bfull + C::B_STAGES;
  uint64_t* tfull = bempty + C::B_STAGES;
  uint64_t* tempty = tfull + C::ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int groups_h = (p.rows_h + MT - 1) / MT;
  const int num_m_tiles = p.tiles_w * groups_h * p.n_img;
  const int num_tiles = num_m_tiles * p.n_tiles_n;
  AbortCtl ab{abort_flag, p.err};

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < na; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
    for (int s = 0; s < C::B_STAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
    for (int a = 0; a < C::ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB[0]); }
    int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
    const uint32_t a_bytes = uint32_t(p.box_w) * 128u, b_bytes = uint32_t(b_stage_bytes);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles_n;
      int mt = tile / p.n_tiles_n;
      const int twi = mt % p.tiles_w; mt /= p.tiles_w;
      const int h0 = (mt % groups_h) * MT;
      const int n = mt / groups_h;
      const int w0 = twi * 128 + p.dw0;
      for (int c = 0; c < p.chunks; ++c) {
        for (int e = 0; e < p.n_entries; ++e) {
          const int a_plane = p.entries[e].a_plane, b_plane = p.entries[e].b_plane;
          const int dh = p.entries[e].dh, tap0 = p.entries[e].tap0;
          mbar_wait(&bempty[bs], bph ^ 1, ab, 701);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&bfull[bs], b_bytes);
            tma_load_3d(&p.tmB[b_plane], &bfull[bs], smem_b + bs * b_stage_bytes, c * 64, nt * BN, tap0);
          }
          if (++bs == C::B_STAGES) { bs = 0; bph ^= 1; }
          for (int r = 0; r < MT; ++r) {
            mbar_wait(&aempty[as], aph ^ 1, ab, 702);
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&afull[as], a_bytes);
              // rows past the image bottom are out of bounds for TMA: zero filled, their accumulators are never stored
              tma_load_4d(&p.tmA[a_plane], &afull[as], smem_a + as * C::A_PLANE, c * 64, w0, h0 + r + dh, n);
            }
            if (++as == na) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, 0);
    int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int a = 0; uint32_t tph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[a], tph ^ 1, ab, 703);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * (MT * BN);
      uint32_t acc = 0;
      for (int c = 0; c < p.chunks; ++c) {
        for (int e = 0; e < p.n_entries; ++e) {
          int k0 = 0, k1 = 4;
          if constexpr (TRIM) k_groups(p.a_c, p.a_lo_c0, p.entries[e].a_plane, c, k0, k1);
          mbar_wait(&bfull[bs], bph, ab, 704);
          const uint32_t b_base = smem_u32(smem_b + bs * b_stage_bytes);
          for (int r = 0; r < MT; ++r) {
            mbar_wait(&afull[as], aph, ab, 705);
            tc_fence_after();
            const uint32_t a_base = smem_u32(smem_a + as * C::A_PLANE);
            if (elect_one_sync()) {
              uint32_t acc_j = acc;
              for (int j = 0; j < p.kw; ++j) {
                const uint64_t ad = umma_smem_desc(a_base + uint32_t(p.entries[e].a_off[j]) * 128u, 16, 1024);
                const uint64_t bd = umma_smem_desc(b_base + uint32_t(j) * C::B_TILE, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if constexpr (TRIM) {
                    if (k >= k0 && k < k1) { umma_bf16(d_tmem + r * BN, ad + 2 * k, bd + 2 * k, idesc, acc_j); acc_j = 1; }
                  } else {
                    umma_bf16(d_tmem + r * BN, ad + 2 * k, bd + 2 * k, idesc, acc_j); acc_j = 1;
                  }
                }
              }
              umma_commit(&aempty[as]);
            }
            if (++as == na) { as = 0; aph ^= 1; }
          }
          if (elect_one_sync()) umma_commit(&bempty[bs]);
          if (++bs == C::B_STAGES) { bs = 0; bph ^= 1; }
          if (k1 > k0) acc = 1;
        }
      }
      if (elect_one_sync()) umma_commit(&tfull[a]);
      if (++a == C::ACC) { a = 0; tph ^= 1; }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int a = 0; uint32_t tph = 0;
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
      const int h0 = (mt % groups_h) * MT;
      const int n = mt / groups_h;
      const int ow = twi * 128 + m;
      mbar_wait(&tfull[a], tph, ab, 706);
      tc_fence_after();
#pragma unroll 1
      for (int r = 0; r < MT; ++r) {
        const int oh = h0 + r;
        const bool valid = (ow < p.valid_w) && (oh < p.rows_h);
        size_t off32 = 0, off16 = 0;
        if (p.o32) off32 = ((size_t(n) * p.o32_H + oh + p.o32_hoff) * p.o32_W + ow + p.o32_woff) * p.o32_C + p.o32_coff;
        if (p.ohi) off16 = ((size_t(n) * p.o16_H + oh + p.o16_hoff) * p.o16_W + ow + p.o16_woff) * p.o16_C + p.o16_coff;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += C::CH) {
          uint32_t raw[C::CH];
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + a * (MT * BN) + r * BN + c0;
          if constexpr (C::CH == 32) tmem_ld32(taddr, raw); else tmem_ld16(taddr, raw);
          tmem_ld_wait();
          const int cg = nt * BN + c0;
          if (valid && cg < p.cout) epilogue_chunk<C::CH>(p, raw, cg, v32, v16, vb, off32, off16);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      if (++a == C::ACC) { a = 0; tph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace hm
