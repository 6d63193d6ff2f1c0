// hm_engine_mnrows.cuh -- row-streaming variant of the MN-engine: weight gradients of stride-1 convolutions on wide
// images (7x7 stem / head at full resolution, LocalEnhancer's full-resolution 3x3 blocks).
//
// The generic MN-engine spends one pipeline stage (3 TMA loads, one mbarrier round trip) on 4 narrow MMAs and
// re-reads the tapped operand once per tap.  Here one CTA tile owns a whole filter ROW (kh) of one 64-channel unit:
// per 128-pixel k-tile the producer loads ONE box of 128+KW pixels of the tapped operand P and the Q box(es) once,
// and the issuer fires ceil(KW/2) x 8 MMAs per stage.  The two 64-row halves of an M=128 MMA are the SAME P box
// shifted by one pixel: in the MN-major shared-memory descriptor the second 64-element group sits LBO = 128 bytes
// (one pixel row) after the first, so a single MMA produces the gradients of taps kw and kw+1.  ceil(KW/2)
// accumulators (one per tap pair) live side by side in TMEM.
#pragma once
#include "hm_engine.cuh"
#include "hm_engine_rows.cuh"

namespace hm {

struct __align__(64) MRParams {
  CUtensorMap tmP[2];   // tapped operand (hi, lo): box {64 ch, box_w px}
  CUtensorMap tmQ[2];   // base-space operand (hi, lo): box {64 ch, 128 px}
  int n_prod; int8_t prodP[4], prodQ[4];
  int kh, kw, pairs;    // pairs = ceil(kw / 2)
  int tiles_w, rows_h, n_img, ktiles;
  int m_units, cp_pad;  // 64-channel units of P; rows of G per tap
  int n_n_tiles, splits;
  int dw0; int16_t dh[kRowsMaxKW + 3];
  int box_w;
  int n_stages, stage_bytes;
  float* G; int ldG; int n_cols; int use_atomic;
  int* err;
};

template <int NB>
struct MRCfg : RCfgCommon {
  static constexpr int BN = NB * 64;
  static constexpr int Q_BOX = 128 * 128;   // 128 pixels x 64 channels
};

template <int NB>
__global__ void __launch_bounds__(kEngineThreads, 1) hm_mnrows_kernel(const __grid_constant__ MRParams p) {
  using C = MRCfg<NB>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.n_stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = full + C::MAX_STAGES;
  uint64_t* tfull = empty + C::MAX_STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int base_tiles = p.kh * p.m_units * p.n_n_tiles;
  const int num_tiles = base_tiles * p.splits;
  const int nst = p.n_stages;
  AbortCtl ab{abort_flag, p.err};
  uint32_t tmem_cols = 32;
  while (tmem_cols < uint32_t(p.pairs * BN)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    *abort_flag = 0;
    for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto k_range = [&](int split, int& k0, int& k1) {
    const int per = (p.ktiles + p.splits - 1) / p.splits;
    k0 = split * per; k1 = min(p.ktiles, k0 + per);
    if (k1 < k0) k1 = k0;
  };
  auto decode = [&](int tile, int& split, int& kh, int& mu, int& nt) {
    nt = tile % p.n_n_tiles; tile /= p.n_n_tiles;
    mu = tile % p.m_units; tile /= p.m_units;
    kh = tile % p.kh;
    split = tile / p.kh;
  };

  if (warp == 0) {
    {
      if (elect_one_sync()) { tma_prefetch_desc(&p.tmP[0]); tma_prefetch_desc(&p.tmQ[0]); }
      int s = 0; uint32_t ph = 0;
      const uint32_t tx_bytes = uint32_t(p.box_w) * 128u + NB * C::Q_BOX;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int split, kh, mu, nt; decode(tile, split, kh, mu, nt);
        int k0, k1; k_range(split, k0, k1);
        for (int pr = 0; pr < p.n_prod; ++pr) {
          const CUtensorMap* mp = &p.tmP[p.prodP[pr]];
          const CUtensorMap* mq = &p.tmQ[p.prodQ[pr]];
          for (int kt = k0; kt < k1; ++kt) {
            int t = kt;
            const int twi = t % p.tiles_w; t /= p.tiles_w;
            const int h = t % p.rows_h;
            const int n = t / p.rows_h;
            const int w0 = twi * 128;
            mbar_wait(&empty[s], ph ^ 1, ab, 401);
            uint8_t* dst = smem + s * p.stage_bytes;
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&full[s], tx_bytes);
              tma_load_4d(mp, &full[s], dst, mu * 64, w0 + p.dw0, h + p.dh[kh], n);
#pragma unroll
              for (int r = 0; r < NB; ++r)
                tma_load_4d(mq, &full[s], dst + C::A_PLANE + r * C::Q_BOX, (nt * NB + r) * 64, w0, h, n);
            }
            if (++s == nst) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 1, 1);
      int s = 0; uint32_t ph = 0; uint32_t tph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int split, kh, mu, nt; decode(tile, split, kh, mu, nt);
        int k0, k1; k_range(split, k0, k1);
        const int ksteps = (k1 - k0) * p.n_prod;
        mbar_wait(tempty, tph ^ 1, ab, 402);   // single-buffered accumulators: wait for the previous epilogue
        tc_fence_after();
        for (int k = 0; k < ksteps; ++k) {
          mbar_wait(&full[s], ph, ab, 403);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + s * p.stage_bytes);
          const uint32_t b_base = a_base + C::A_PLANE;
          if (elect_one_sync()) {
            for (int pi = 0; pi < p.pairs; ++pi) {
              // A: taps (2pi, 2pi+1) = two 64-channel groups one pixel row (128 B) apart; 8 K-rows every 1024 B
              const uint64_t ad = umma_smem_desc(a_base + uint32_t(2 * pi) * 128u, 128, 1024);
              const uint64_t bd = umma_smem_desc(b_base, C::Q_BOX, 1024);
#pragma unroll
              for (int j = 0; j < 8; ++j)   // 16 pixels (K) per MMA = 2048 B
                umma_bf16(tmem_base + pi * BN, ad + j * (2048 >> 4), bd + j * (2048 >> 4), idesc, (k | j) != 0);
            }
            umma_commit(&empty[s]);
          }
          if (++s == nst) { s = 0; ph ^= 1; }
        }
        if (elect_one_sync()) umma_commit(tfull);
        tph ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    uint32_t tph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int split, kh, mu, nt; decode(tile, split, kh, mu, nt);
      int k0, k1; k_range(split, k0, k1);
      mbar_wait(tfull, tph, ab, 404);
      tc_fence_after();
      for (int pi = 0; pi < p.pairs; ++pi) {
        const int kwi = 2 * pi + (m >> 6);
        const bool row_ok = (kwi < p.kw) && (k1 > k0);
        float* grow = p.G + (size_t(kh * p.kw + kwi) * p.cp_pad + mu * 64 + (m & 63)) * p.ldG;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t raw[32];
          tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + pi * BN + c0, raw);
          tmem_ld_wait();
          const int cg = nt * BN + c0;
          if (row_ok && cg < p.n_cols) {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      tph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace hm
