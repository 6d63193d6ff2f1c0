// hm_engine2.cuh -- CTA-pair (cta_group::2) variant of the K-engine for wide N tiles (Cout >= 256).
//
// Why: with one CTA per SM a 128x256x64 k-step makes the tensor core READ 48 KB of shared memory (A 16 KB + B 32 KB)
// while TMA WRITES another 48 KB for a later stage -- 96 KB per 512 tensor cycles is 1.5x the 128 B/cycle shared
// memory port, which caps the single-CTA kernel at ~70 % tensor-pipe utilisation (ncu, profiles/r01_ncu_k1_*).
// Two CTAs of a cluster (the two SMs of a TPC) instead share one 256x256 tile: each loads its own 128-pixel A box and
// only HALF of the weight slab (128 of the 256 rows); one tcgen05.mma.cta_group::2 (M = 256) issued by the leader CTA
// consumes both halves through the pair's shared-memory view, so per SM and k-step 32 KB are read and 32 KB written.
// Pipeline protocol (same as CUTLASS / DeepGEMM 2-SM kernels): both producers' TMA loads complete on the LEADER's
// full barrier (cta_group::2 loads, peer bit cleared); the leader's commits are multicast to both CTAs' empty /
// tmem-full barriers; both epilogues arrive on the leader's tmem-empty barrier.
#pragma once
#include "hm_engine.cuh"

namespace hm {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // address of the same object in the even (leader) CTA of the pair
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

struct K2Cfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = 128 * 128;       // my 128 pixels x 64 bf16
  static constexpr int B_BYTES = 128 * 128;       // my half of the weight slab: 128 rows x 64 bf16
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 6;
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};
// fused-split variant (bf16x3): one stage = [A_hi | A_lo | B_hi | B_lo] of a (tap, chunk), three products per stage.  K1
// moves 3.6 GB from L2 to the SMs per launch in the per-product layout (ncu r02: 11.8 TB/s, the fabric limit) because
// A_hi and B_hi are staged twice; here a third of that traffic (and of the shared-memory fill) disappears.
struct K2FCfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = 128 * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int ACC = 2;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

// ------------------------------------------------------------------------------------------------
// Deterministic stream-K tail.  A persistent tile loop leaves the last wave partly empty: the residual-block conv
// (K1: 128 pair tiles on 74 clusters = 1.73 waves) kept the tensor pipe busy for only 78 % of the elapsed cycles.  The
// host therefore splits the work in two phases: pair tiles [0, sk_full) (whole waves) are processed as before; the
// k-steps of the remaining sk_rem tiles are dealt out as ONE flattened range, sk_per consecutive k-steps per cluster,
// so every cluster finishes at the same time.  A cluster's range covers the end of one tile and / or the start of the
// next; such a SEGMENT's raw fp32 accumulator goes to a workspace slot, an arrival counter per (tile, CTA rank) is
// bumped, and whoever arrives LAST sums the segments of the tile IN SEGMENT ORDER (not arrival order) and runs the
// normal epilogue -- the result is bit-identical from run to run, and nobody ever waits for another cluster.
// ------------------------------------------------------------------------------------------------
struct SKWork {
  int tile;         // pair tile
  int k0, k1;       // k-step range of this work item
  int partial;      // 1: a segment of a split tile (fix-up path)
  int seg, nseg;    // position among the tile's segments
  int slot;         // workspace slot of this segment (cluster * 2 + {0, 1})
  int r, c_first;   // remainder-tile index and first cluster touching it
};

struct SKIter {
  int cid, G, full, ksteps, per, total_k;
  int t1, kb, ke, nth;
  __device__ __forceinline__ void init(const KParams& p, int cluster_id, int n_clusters, int num_tiles, int ksteps_) {
    cid = cluster_id; G = n_clusters; ksteps = ksteps_;
    if (p.sk_per > 0) { full = p.sk_full; per = p.sk_per; total_k = p.sk_rem * ksteps_; }
    else { full = num_tiles; per = 0; total_k = 0; }
    t1 = cid;
    kb = min(cid * per, total_k); ke = min(kb + per, total_k);
    nth = 0;
  }
  __device__ __forceinline__ bool next(SKWork& w) {
    if (t1 < full) {
      w.tile = t1; w.k0 = 0; w.k1 = ksteps; w.partial = 0; w.seg = 0; w.nseg = 1; w.slot = 0; w.r = 0; w.c_first = 0;
      t1 += G;
      return true;
    }
    if (kb >= ke) return false;
    const int r = kb / ksteps;
    const int k0 = kb - r * ksteps;
    const int len = min(ksteps - k0, ke - kb);
    const int c_first = (r * ksteps) / per;
    const int c_last = ((r + 1) * ksteps - 1) / per;
    w.tile = full + r; w.k0 = k0; w.k1 = k0 + len;
    w.nseg = c_last - c_first + 1; w.partial = w.nseg > 1; w.seg = cid - c_first;
    w.slot = cid * 2 + nth; w.r = r; w.c_first = c_first;
    kb += len; ++nth;
    return true;
  }
};

// workspace slot of segment s of remainder tile r: the cluster's first segment uses slot 0, its second slot 1
__device__ __forceinline__ int sk_slot_of(int r, int c_first, int s, int per, int ksteps) {
  const int c = c_first + s;
  const int first_tile_of_c = (c * per) / ksteps;
  return c * 2 + (first_tile_of_c == r ? 0 : 1);
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// KParams is reused; tmB must have been encoded with a 128-row box.
template <int BN, bool FUSED3 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kEngineThreads, 1)
hm_kgemm2_kernel(const __grid_constant__ KParams p) {
  using C = typename hm_cond<FUSED3, K2FCfg, K2Cfg>::type;
  static_assert(BN == C::BN, "the CTA-pair kernel is built for 256-wide N tiles");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* tfull = bars + 2 * C::STAGES;
  uint64_t* tempty = tfull + C::ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ACC);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  volatile int* last_flag = abort_flag + 1;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_m_tiles = p.tiles_w * p.tiles_h * p.n_img;
  const int num_m2 = (num_m_tiles + 1) >> 1;
  const int num_tiles = num_m2 * p.n_tiles_n;           // pair tiles
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int ksteps = p.n_entries * p.chunks;
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

  SKIter it;
  it.init(p, cluster_id, n_clusters, num_tiles, ksteps);
  SKWork w;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one_sync()) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB[0]); }
    int s = 0; uint32_t ph = 0;
    while (it.next(w)) {
      const int nt = w.tile % p.n_tiles_n;
      const int mt2 = w.tile / p.n_tiles_n;
      int mt = min(mt2 * 2 + int(rank), num_m_tiles - 1);
      const int twi = mt % p.tiles_w; mt /= p.tiles_w;
      const int thi = mt % p.tiles_h;
      const int n = mt / p.tiles_h;
      const int w0 = (twi << p.tw_log2) * p.in_stride;
      const int h0 = thi * p.th * p.in_stride;
      int e = w.k0 / p.chunks, c = w.k0 - e * p.chunks;
      for (int k = w.k0; k < w.k1; ++k) {
        const KEntry en = p.entries[e];
        mbar_wait(&empty[s], ph ^ 1, ab, 501);
        uint8_t* sa = smem + s * C::STAGE_BYTES;
        // only the leader arms its full barrier, for the bytes of BOTH CTAs; the peer's loads cannot run ahead of
        // it by a phase because they are gated by the leader's multicast commit on empty[s]
        if (elect_one_sync()) {
          if (leader) mbar_arrive_expect_tx(&full[s], 2 * C::STAGE_BYTES);
          if constexpr (FUSED3) {
            tma_load_4d_2sm(&p.tmA[0], &full[s], sa, c * 64, w0 + en.dw, h0 + en.dh, n);
            tma_load_4d_2sm(&p.tmA[1], &full[s], sa + C::A_BYTES, c * 64, w0 + en.dw, h0 + en.dh, n);
            tma_load_2d_2sm(&p.tmB[0], &full[s], sa + 2 * C::A_BYTES, c * 64, en.b_row + nt * BN + int(rank) * 128);
            tma_load_2d_2sm(&p.tmB[1], &full[s], sa + 2 * C::A_BYTES + C::B_BYTES, c * 64, en.b_row + nt * BN + int(rank) * 128);
          } else {
            tma_load_4d_2sm(&p.tmA[en.a_plane], &full[s], sa, c * 64, w0 + en.dw, h0 + en.dh, n);
            tma_load_2d_2sm(&p.tmB[en.b_plane], &full[s], sa + C::A_BYTES, c * 64, en.b_row + nt * BN + int(rank) * 128);
          }
        }
        if (++s == C::STAGES) { s = 0; ph ^= 1; }
        if (++c == p.chunks) { c = 0; ++e; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
      int s = 0; uint32_t ph = 0; int a = 0; uint32_t aph = 0;
      while (it.next(w)) {
        mbar_wait(&tempty[a], aph ^ 1, ab, 502);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int k = w.k0; k < w.k1; ++k) {
          mbar_wait(&full[s], ph, ab, 503);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
          if constexpr (FUSED3) {
            const uint64_t ah = umma_smem_desc(sa, 16, 1024), al = umma_smem_desc(sa + C::A_BYTES, 16, 1024);
            const uint64_t bh = umma_smem_desc(sa + 2 * C::A_BYTES, 16, 1024);
            const uint64_t bl = umma_smem_desc(sa + 2 * C::A_BYTES + C::B_BYTES, 16, 1024);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16_2sm(d_tmem, ah + 2 * j, bh + 2 * j, idesc, uint32_t((k != w.k0) | (j != 0)));
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16_2sm(d_tmem, al + 2 * j, bh + 2 * j, idesc, 1u);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16_2sm(d_tmem, ah + 2 * j, bl + 2 * j, idesc, 1u);
              umma_commit_2sm_mc(&empty[s], 3);
            }
          } else {
            const uint64_t adesc = umma_smem_desc(sa, 16, 1024);
            const uint64_t bdesc = umma_smem_desc(sa + C::A_BYTES, 16, 1024);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_bf16_2sm(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, uint32_t((k != w.k0) | (j != 0)));
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
    const bool v32 = p.o32 && ((p.o32_C & 3) == 0) && ((p.o32_coff & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.o32) & 15) == 0);
    const bool v16 = p.ohi && ((p.o16_C & 7) == 0) && ((p.o16_coff & 7) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.ohi) & 15) == 0) &&
                     (!p.olo || (reinterpret_cast<uintptr_t>(p.olo) & 15) == 0);
    const bool vb = p.bias && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    constexpr size_t kSlotFloats = size_t(256) * BN;      // one pair-tile accumulator
    while (it.next(w)) {
      const int nt = w.tile % p.n_tiles_n;
      const int mt2 = w.tile / p.n_tiles_n;
      int mt = mt2 * 2 + int(rank);
      const bool tile_ok = mt < num_m_tiles;
      mt = min(mt, num_m_tiles - 1);
      const int twi = mt % p.tiles_w; mt /= p.tiles_w;
      const int thi = mt % p.tiles_h;
      const int n = mt / p.tiles_h;
      const int ht = thi * p.th + (m >> p.tw_log2);
      const int wt = (twi << p.tw_log2) + (m & ((1 << p.tw_log2) - 1));
      const bool valid = tile_ok && (ht < p.valid_h) && (wt < p.valid_w);
      const int oh = ht * p.out_sh + p.out_oh, ow = wt * p.out_sw + p.out_ow;
      size_t off32 = 0, off16 = 0;
      if (p.o32) off32 = ((size_t(n) * p.o32_H + oh + p.o32_hoff) * p.o32_W + ow + p.o32_woff) * p.o32_C + p.o32_coff;
      if (p.ohi) off16 = ((size_t(n) * p.o16_H + oh + p.o16_hoff) * p.o16_W + ow + p.o16_woff) * p.o16_C + p.o16_coff;
      mbar_wait(&tfull[a], aph, ab, 504);
      tc_fence_after();
      if (!w.partial) {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + a * BN + c0, raw);
          tmem_ld_wait();
          const int cg = nt * BN + c0;
          if (valid && cg < p.cout) epilogue_chunk<32>(p, raw, cg, v32, v16, vb, off32, off16);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tempty[a]);
          else mbar_arrive_remote(mapa_u32(smem_u32(&tempty[a]), 0));
        }
      } else {
        // ---- split tile: park the raw accumulator rows of this segment, release TMEM, count the arrival ----
        // slot layout (private to this kernel): float4 index ((c0 / 32) * 8 + i / 4) * 128 + m, i.e. the 128 epilogue
        // threads of a CTA write / read consecutive 16 B words -> fully coalesced (a row-per-thread layout costs 8x the
        // L2 transactions and made the tail slower than the wave quantisation it removes)
        float4* mine = reinterpret_cast<float4*>(p.sk_ws + (size_t(w.slot) * 2 + rank) * 128 * BN) + m;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + a * BN + c0, raw);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            __stcg(mine + ((c0 >> 5) * 8 + (i >> 2)) * 128,
                   make_float4(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]), __uint_as_float(raw[i + 2]),
                               __uint_as_float(raw[i + 3])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&tempty[a]);
          else mbar_arrive_remote(mapa_u32(smem_u32(&tempty[a]), 0));
        }
        __threadfence();
        epi_bar_sync();
        if (m == 0) {
          int* cnt = p.sk_cnt + w.r * 2 + int(rank);
          const int old = atomicAdd(cnt, 1);
          const int is_last = (old == w.nseg - 1) ? 1 : 0;
          if (is_last) *cnt = 0;      // self-resetting: the counters are zero again when the kernel ends (no memset per launch)
          *last_flag = is_last;
        }
        epi_bar_sync();
        const bool last = *last_flag != 0;
        epi_bar_sync();            // everybody has read the flag before the next split tile may overwrite it
        if (last) {
          __threadfence();
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
            for (int sgm = 0; sgm < w.nseg; ++sgm) {     // fixed order: bit-identical whoever arrives last
              const float4* src = reinterpret_cast<const float4*>(
                  p.sk_ws + (size_t(sk_slot_of(w.r, w.c_first, sgm, it.per, ksteps)) * 2 + rank) * 128 * BN) + m;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 v = __ldcg(src + ((c0 >> 5) * 8 + (i >> 2)) * 128);
                acc[i] += v.x; acc[i + 1] += v.y; acc[i + 2] += v.z; acc[i + 3] += v.w;
              }
            }
            uint32_t raw[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(acc[i]);
            const int cg = nt * BN + c0;
            if (valid && cg < p.cout) epilogue_chunk<32>(p, raw, cg, v32, v16, vb, off32, off16);
          }
        }
      }
      if (++a == C::ACC) { a = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
}

}  // namespace hm
