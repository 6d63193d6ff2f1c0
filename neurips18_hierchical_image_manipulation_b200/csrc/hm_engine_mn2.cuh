// hm_engine_mn2.cuh -- CTA-pair (cta_group::2) variant of the MN-engine for wide weight gradients (>= 4 N units, i.e.
// Cout >= 256: the residual blocks, the deep G / D layers).
//
// Why: the single-CTA MN-engine loads 2 P boxes + 4 Q boxes (48 KB) per 64-pixel k-tile for four M=128 x N=256 MMAs:
// 96 shared-memory cycles per 64 tensor cycles, and ~3x the bytes per FLOP of the forward pair kernel through the
// L2 -> SM fabric (ncu r01: 65 % tensor pipe at 11.8 TB/s xbar traffic).  Here two CTAs of a cluster share one
// 256 x 256 tile of G: each loads ITS 2 P boxes (its 128 accumulator rows) and HALF of the N side (2 of the 4 Q
// boxes); one tcgen05.mma.cta_group::2 (M = 256, both operands MN-major) per 16 pixels consumes both halves, so per
// SM and k-tile 32 KB are loaded for 512 tensor cycles of work.  Barrier protocol identical to hm_engine2.cuh.
#pragma once
#include "hm_engine2.cuh"

namespace hm {

struct MN2Cfg {
  static constexpr int BN = 256;
  static constexpr int BOX_BYTES = 64 * 128;          // 64 pixels x 64 bf16
  static constexpr int A_BYTES = 2 * BOX_BYTES;       // my 128 M rows
  static constexpr int B_BYTES = 2 * BOX_BYTES;       // my 128 of the 256 N columns
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// fused-split variant (bf16x3): one stage = [P_hi | P_lo | Q_hi | Q_lo] boxes of a k-tile, three products per stage
struct MN2FCfg {
  static constexpr int BN = 256;
  static constexpr int BOX_BYTES = 64 * 128;
  static constexpr int A_BYTES = 2 * BOX_BYTES;       // per plane
  static constexpr int B_BYTES = 2 * BOX_BYTES;       // per plane
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// MNParams is reused: n_m_tiles counts 4-unit (256-row) M tiles, n_n_tiles 4-unit (256-column) N tiles.
template <bool FUSED3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kEngineThreads, 1)
hm_mngemm2_kernel(const __grid_constant__ MNParams p) {
  using C = typename hm_cond<FUSED3, MN2FCfg, MN2Cfg>::type;
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
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = p.n_m_tiles * p.n_n_tiles * p.splits;
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  AbortCtl ab{abort_flag, p.err};

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < C::ACC; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto k_range = [&](int split, int& k0, int& k1) {
    const int per = (p.ktiles + p.splits - 1) / p.splits;
    k0 = split * per; k1 = min(p.ktiles, k0 + per);
    if (k1 < k0) k1 = k0;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one_sync()) { tma_prefetch_desc(&p.tmP[0]); tma_prefetch_desc(&p.tmQ[0]); }
    int s = 0; uint32_t ph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int nt = tile % p.n_n_tiles;
      const int mt = (tile / p.n_n_tiles) % p.n_m_tiles;
      const int split = tile / (p.n_n_tiles * p.n_m_tiles);
      int k0, k1; k_range(split, k0, k1);
      int mc[2], mdw[2], mdh[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int u = min(mt * 4 + int(rank) * 2 + r, p.m_units - 1);
        if (p.m_tapped) { const int t = u / p.upt_m; mc[r] = (u % p.upt_m) * 64; mdw[r] = p.tap_dw[t]; mdh[r] = p.tap_dh[t]; }
        else { mc[r] = u * 64; mdw[r] = p.dwP0; mdh[r] = p.dhP0; }
      }
      int nc[2], ndw[2], ndh[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int u = min(nt * 4 + int(rank) * 2 + r, p.n_units - 1);
        if (!p.m_tapped) { const int t = u / p.upt_n; nc[r] = (u % p.upt_n) * 64; ndw[r] = p.tap_dw[t]; ndh[r] = p.tap_dh[t]; }
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
          mbar_wait(&empty[s], ph ^ 1, ab, 601);
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          if (elect_one_sync()) {
            if (leader) mbar_arrive_expect_tx(&full[s], 2 * C::STAGE_BYTES);
            if constexpr (FUSED3) {
#pragma unroll
              for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
                  tma_load_4d_2sm(&p.tmP[pl], &full[s], sa + pl * C::A_BYTES + r * C::BOX_BYTES, mc[r], w0 * p.sP + mdw[r],
                                  h0 * p.sP + mdh[r], n);
#pragma unroll
                for (int r = 0; r < 2; ++r)
                  tma_load_4d_2sm(&p.tmQ[pl], &full[s], sa + 2 * C::A_BYTES + pl * C::B_BYTES + r * C::BOX_BYTES, nc[r],
                                  w0 * p.sQ + ndw[r], h0 * p.sQ + ndh[r], n);
              }
            } else {
#pragma unroll
              for (int r = 0; r < 2; ++r)
                tma_load_4d_2sm(mp, &full[s], sa + r * C::BOX_BYTES, mc[r], w0 * p.sP + mdw[r], h0 * p.sP + mdh[r], n);
#pragma unroll
              for (int r = 0; r < 2; ++r)
                tma_load_4d_2sm(mq, &full[s], sa + C::A_BYTES + r * C::BOX_BYTES, nc[r], w0 * p.sQ + ndw[r],
                                h0 * p.sQ + ndh[r], n);
            }
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 1, 1);
      int s = 0; uint32_t ph = 0; int a = 0; uint32_t aph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int split = tile / (p.n_n_tiles * p.n_m_tiles);
        int k0, k1; k_range(split, k0, k1);
        const int ksteps = (k1 - k0) * (FUSED3 ? 1 : p.n_pairs);
        mbar_wait(&tempty[a], aph ^ 1, ab, 602);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int k = 0; k < ksteps; ++k) {
          mbar_wait(&full[s], ph, ab, 603);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
          // MN-major SW128: LBO = distance between 64-channel groups (one TMA box), SBO = 8 K-rows = 1024 B
          if constexpr (FUSED3) {
            const uint64_t ph_ = umma_smem_desc(sa, C::BOX_BYTES, 1024), pl_ = umma_smem_desc(sa + C::A_BYTES, C::BOX_BYTES, 1024);
            const uint64_t qh_ = umma_smem_desc(sa + 2 * C::A_BYTES, C::BOX_BYTES, 1024);
            const uint64_t ql_ = umma_smem_desc(sa + 2 * C::A_BYTES + C::B_BYTES, C::BOX_BYTES, 1024);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16_2sm(d_tmem, ph_ + j * (2048 >> 4), qh_ + j * (2048 >> 4), idesc, (k | j) != 0);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16_2sm(d_tmem, pl_ + j * (2048 >> 4), qh_ + j * (2048 >> 4), idesc, 1u);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16_2sm(d_tmem, ph_ + j * (2048 >> 4), ql_ + j * (2048 >> 4), idesc, 1u);
              umma_commit_2sm_mc(&empty[s], 3);
            }
          } else {
          const uint64_t adesc = umma_smem_desc(sa, C::BOX_BYTES, 1024);
          const uint64_t bdesc = umma_smem_desc(sa + C::A_BYTES, C::BOX_BYTES, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int j = 0; j < 4; ++j)  // 16 pixels (K) per MMA = 16 rows x 128 B = 2048 B
              umma_bf16_2sm(d_tmem, adesc + j * (2048 >> 4), bdesc + j * (2048 >> 4), idesc, (k | j) != 0);
            umma_commit_2sm_mc(&empty[s], 3);
          }
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one_sync()) umma_commit_2sm_mc(&tfull[a], 3);
        if (++a == C::ACC) { a = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (both CTAs; each drains its own 128 accumulator rows) =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int a = 0; uint32_t aph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int nt = tile % p.n_n_tiles;
      const int mt = (tile / p.n_n_tiles) % p.n_m_tiles;
      const int split = tile / (p.n_n_tiles * p.n_m_tiles);
      int k0, k1; k_range(split, k0, k1);
      const int u = mt * 4 + int(rank) * 2 + (m >> 6);
      const bool row_ok = (u < p.m_units) && (k1 > k0);
      float* grow = p.G + size_t(u * 64 + (m & 63)) * p.ldG;
      const int ncols = p.n_units * 64;
      mbar_wait(&tfull[a], aph, ab, 604);
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
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty[a]);
        else mbar_arrive_remote(mapa_u32(smem_u32(&tempty[a]), 0));
      }
      if (++a == C::ACC) { a = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
}

}  // namespace hm
