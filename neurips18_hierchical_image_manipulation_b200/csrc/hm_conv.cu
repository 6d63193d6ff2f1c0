// hm_conv.cu -- host-side planner + C ABI for the tcgen05 convolution engines (hm_engine.cuh).
// Every conv of the mask2image hot path (reference: models/Pix2Pix_NET.py:63-101, models/layer_util.py:333-411,
// models/Discriminator_NET.py:61-118) is lowered here to a tap table + TMA tensor maps; nothing is im2col'ed.
#include "../../include/hm_b200.h"
#include "hm_engine.cuh"
#include "hm_engine_rows.cuh"
#include "hm_engine_mnrows.cuh"
#include "hm_engine2.cuh"
#include "hm_engine_mn2.cuh"

#include <cstdlib>

#include <algorithm>
#include <cstring>
#include <mutex>

namespace {

thread_local int g_last_cuda_error = 0;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    else g_last_cuda_error = int(e);
  });
  return fn;
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// bf16 NHWC tensor [n][h][w][cs] (c valid) -> 4-D map {c, w, h, n}; box = 64 channels x (bw x bh) pixels
// visited with element stride es (so the box spans bw*es x bh*es input pixels).
int make_tmap_nhwc(CUtensorMap* m, const void* base, int n, int h, int w, int c, int cs, int bw, int bh, int es) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return HM_ERR_DRIVER;
  if ((cs & 7) || (reinterpret_cast<uintptr_t>(base) & 15) || bw * es > 256 || bh * es > 256) return HM_ERR_INVALID;
  cuuint64_t dims[4] = {cuuint64_t(c), cuuint64_t(w), cuuint64_t(h), cuuint64_t(n)};
  cuuint64_t strides[3] = {cuuint64_t(cs) * 2, cuuint64_t(w) * cs * 2, cuuint64_t(h) * w * cs * 2};
  cuuint32_t box[4] = {64, cuuint32_t(bw * es), cuuint32_t(bh * es), 1};
  cuuint32_t estr[4] = {1, cuuint32_t(es), cuuint32_t(es), 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HM_OK : HM_ERR_TENSORMAP;
}

// packed weights [rows_total][k_pad] bf16 -> 2-D map, box = 64 (k) x bn rows
int make_tmap_weight(CUtensorMap* m, const void* base, int rows_total, int k_pad, int bn) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return HM_ERR_DRIVER;
  if ((k_pad & 63) || (reinterpret_cast<uintptr_t>(base) & 15)) return HM_ERR_INVALID;
  cuuint64_t dims[2] = {cuuint64_t(k_pad), cuuint64_t(rows_total)};
  cuuint64_t strides[1] = {cuuint64_t(k_pad) * 2};
  cuuint32_t box[2] = {64, cuuint32_t(bn)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HM_OK : HM_ERR_TENSORMAP;
}

int g_sm_limit = 0;            // hm_set_sm_limit: leave SMs free for a concurrent collective (0 = use all)
int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return (g_sm_limit > 0 && g_sm_limit < n) ? g_sm_limit : n;
}

// choose the th x tw = `pixels` rectangle that wastes the least area on a (vh x vw) tile space
void pick_rect(int pixels, int vh, int vw, int max_side, int* tw, int* th) {
  long best = -1;
  for (int w = pixels; w >= 8; w >>= 1) {
    int h = pixels / w;
    if (w > max_side || h > max_side) continue;
    long area = long(round_up(vw, w)) * round_up(vh, h);
    if (best < 0 || area < best) { best = area; *tw = w; *th = h; }
  }
}

template <int BN>
int launch_k(const hm::KParams& p, int num_tiles, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_kgemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         hm::KCfg<BN>::SMEM_BYTES);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = true;
  }
  int grid = std::min(num_tiles, sm_count());
  hm::hm_kgemm_kernel<BN><<<grid, hm::kEngineThreads, hm::KCfg<BN>::SMEM_BYTES, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}

// scratch registered by the caller (hm_set_scratch): stream-K partial accumulators + arrival counters
int g_use_streamk = -1;       // -1: take HM_STREAMK from the environment on first use; hm_set_streamk overrides
void* g_scratch = nullptr;
size_t g_scratch_bytes = 0;
constexpr size_t kSkSlotBytes = size_t(256) * 256 * sizeof(float);
constexpr size_t kSkCounterBytes = 4096;

int launch_k2(hm::KParams& p, int num_m_tiles, int n_tiles_n, bool fused3, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_kgemm2_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         hm::K2Cfg::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(hm::hm_kgemm2_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               hm::K2FCfg::SMEM_BYTES);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = true;
  }
  const int pair_tiles = ((num_m_tiles + 1) / 2) * n_tiles_n;
  const int max_clusters = sm_count() / 2;
  int clusters = std::min(pair_tiles, max_clusters);
  // stream-K tail (hm_engine2.cuh): whole waves as tiles, the k-steps of the last partial wave dealt out evenly
  // Measured on B200 (tools/bench_k1_variants.py, profiles/r02_k1_streamk_variants.txt): K1 0.294 ms with the tail vs
  // 0.297 ms with whole tiles in bf16x3, and 3-20 % SLOWER on the other pair-engine shapes and in bf16 -- parking and
  // re-reading the partial accumulators (2 x 33 MB through L2 for K1) plus the un-overlapped fix-up epilogue cost what the
  // better balance buys.  Hence opt-in (HM_STREAMK=1); the default stays the whole-tile schedule.
  const int use_sk = g_use_streamk < 0 ? (g_use_streamk = env_int("HM_STREAMK", 0)) : g_use_streamk;
  const int ksteps = p.n_entries * p.chunks;
  const int waves = pair_tiles / max_clusters;
  const int rem = pair_tiles - waves * max_clusters;
  p.sk_full = p.sk_rem = p.sk_per = 0;
  p.sk_ws = nullptr; p.sk_cnt = nullptr;
  const size_t need = kSkCounterBytes + size_t(max_clusters) * 2 * kSkSlotBytes;
  if (use_sk && g_scratch && g_scratch_bytes >= need && rem > 0 && size_t(rem) * 2 * sizeof(int) <= kSkCounterBytes &&
      rem * 10 <= max_clusters * 9 && (long(rem) * ksteps) / max_clusters >= 24) {
    clusters = max_clusters;
    p.sk_full = waves * max_clusters;
    p.sk_rem = rem;
    p.sk_per = int((long(rem) * ksteps + clusters - 1) / clusters);
    p.sk_cnt = static_cast<int*>(g_scratch);
    p.sk_ws = reinterpret_cast<float*>(static_cast<char*>(g_scratch) + kSkCounterBytes);
  }
  if (fused3) hm::hm_kgemm2_kernel<256, true><<<2 * clusters, hm::kEngineThreads, hm::K2FCfg::SMEM_BYTES, st>>>(p);
  else hm::hm_kgemm2_kernel<256, false><<<2 * clusters, hm::kEngineThreads, hm::K2Cfg::SMEM_BYTES, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}

template <int BN>
int launch_k3(const hm::KParams& p, int num_tiles, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_kgemm_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         hm::K3Cfg<BN>::SMEM_BYTES);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = true;
  }
  int grid = std::min(num_tiles, sm_count());
  hm::hm_kgemm_kernel<BN, true><<<grid, hm::kEngineThreads, hm::K3Cfg<BN>::SMEM_BYTES, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}

int launch_k3_bn(int bn, const hm::KParams& p, int num_tiles, cudaStream_t st) {
  switch (bn) {
    case 16: return launch_k3<16>(p, num_tiles, st);
    case 32: return launch_k3<32>(p, num_tiles, st);
    case 64: return launch_k3<64>(p, num_tiles, st);
    case 128: return launch_k3<128>(p, num_tiles, st);
  }
  return HM_ERR_INVALID;
}

int launch_k_bn(int bn, const hm::KParams& p, int num_tiles, cudaStream_t st) {
  switch (bn) {
    case 16: return launch_k<16>(p, num_tiles, st);
    case 32: return launch_k<32>(p, num_tiles, st);
    case 64: return launch_k<64>(p, num_tiles, st);
    case 128: return launch_k<128>(p, num_tiles, st);
    case 256: return launch_k<256>(p, num_tiles, st);
  }
  return HM_ERR_INVALID;
}

template <int NB>
int launch_mn(const hm::MNParams& p, int num_tiles, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_mngemm_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         hm::MNCfg<NB>::SMEM_BYTES);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = true;
  }
  int grid = std::min(num_tiles, sm_count());
  hm::hm_mngemm_kernel<NB><<<grid, hm::kEngineThreads, hm::MNCfg<NB>::SMEM_BYTES, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}

template <int NB>
int launch_mn3(const hm::MNParams& p, int num_tiles, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_mngemm_kernel<NB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         hm::MN3Cfg<NB>::SMEM_BYTES);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = true;
  }
  int grid = std::min(num_tiles, sm_count());
  hm::hm_mngemm_kernel<NB, true><<<grid, hm::kEngineThreads, hm::MN3Cfg<NB>::SMEM_BYTES, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}

int launch_mn2(const hm::MNParams& p, int num_tiles, bool fused3, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_mngemm2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         hm::MN2Cfg::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(hm::hm_mngemm2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               hm::MN2FCfg::SMEM_BYTES);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = true;
  }
  const int clusters = std::min(num_tiles, sm_count() / 2);
  if (fused3) hm::hm_mngemm2_kernel<true><<<2 * clusters, hm::kEngineThreads, hm::MN2FCfg::SMEM_BYTES, st>>>(p);
  else hm::hm_mngemm2_kernel<false><<<2 * clusters, hm::kEngineThreads, hm::MN2Cfg::SMEM_BYTES, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}

template <int BN>
int launch_rows(const hm::RParams& p, int num_tiles, int smem_bytes, cudaStream_t st) {
  static int configured = 0;
  if (configured < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_krows_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = smem_bytes;
  }
  int grid = std::min(num_tiles, sm_count());
  hm::hm_krows_kernel<BN><<<grid, hm::kEngineThreads, smem_bytes, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}

template <int BN, bool TRIM>
int launch_rows2_t(const hm::RParams& p, int num_tiles, int smem_bytes, cudaStream_t st) {
  static int configured = 0;
  if (configured < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_krows2_kernel<BN, TRIM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = smem_bytes;
  }
  int grid = std::min(num_tiles, sm_count());
  hm::hm_krows2_kernel<BN, TRIM><<<grid, hm::kEngineThreads, smem_bytes, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}
// TRIM variant (skips all-zero 16-channel groups) only when the operand has any: partial last chunk or exact low channels
template <int BN>
int launch_rows2(const hm::RParams& p, int num_tiles, int smem_bytes, cudaStream_t st) {
  const bool trim = (p.a_c & 63) != 0 && (((p.a_c & 63) + 15) >> 4) < 4 || p.a_lo_c0 >= 16;
  return trim ? launch_rows2_t<BN, true>(p, num_tiles, smem_bytes, st) : launch_rows2_t<BN, false>(p, num_tiles, smem_bytes, st);
}

// packed weights [taps][rows_pad][k_pad] bf16 -> 3-D map, box = 64 (k) x bn rows x kw taps
int make_tmap_weight3(CUtensorMap* m, const void* base, int taps, int rows_pad, int k_pad, int bn, int kw) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return HM_ERR_DRIVER;
  if ((k_pad & 63) || (reinterpret_cast<uintptr_t>(base) & 15)) return HM_ERR_INVALID;
  cuuint64_t dims[3] = {cuuint64_t(k_pad), cuuint64_t(rows_pad), cuuint64_t(taps)};
  cuuint64_t strides[2] = {cuuint64_t(k_pad) * 2, cuuint64_t(rows_pad) * k_pad * 2};
  cuuint32_t box[3] = {64, cuuint32_t(bn), cuuint32_t(kw)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HM_OK : HM_ERR_TENSORMAP;
}

struct Tap { int dw, dh, slab; };

// Row-streaming engine (hm_engine_rows.cuh): stride-1 tap grids on wide images with a narrow N tile.
// Dilated convolutions (hm_conv_*_dil with dilation > 1) have gaps between their horizontal taps: the row-streaming
// engines, which read all taps of a filter row out of one contiguous row box, are not eligible for them.
thread_local bool g_dilated = false;

bool rows_eligible(const Tap* taps, int n_taps, int in_stride, int out_sh, int out_sw, int valid_w, int bn) {
  static const int enabled = env_int("HM_ROWS", 1);
  if (g_dilated) return false;
  if (!enabled || in_stride != 1 || out_sh != 1 || out_sw != 1 || bn > 128 || valid_w < 96 || n_taps > 64) return false;
  int dw_min = taps[0].dw, dw_max = taps[0].dw;
  for (int t = 0; t < n_taps; ++t) { dw_min = std::min(dw_min, taps[t].dw); dw_max = std::max(dw_max, taps[t].dw); }
  const int kw = dw_max - dw_min + 1;
  if (kw < 2 || kw > hm::kRowsMaxKW || n_taps % kw) return false;
  return 2 * (hm::RCfgCommon::A_PLANE + kw * bn * 128) <= hm::RCfgCommon::SMEM_BUDGET;
}

int run_rows_engine(const hm_operand* x, const void* w_hi, const void* w_lo, int k_pad, int rows_pad, const float* bias,
                    const Tap* taps, int n_taps, int valid_h, int valid_w, int Cout, int act, float slope,
                    const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, cudaStream_t st) {
  const int bn = hm_pick_bn(Cout);
  hm::RParams p;
  std::memset(&p, 0, sizeof(p));
  int dw_min = taps[0].dw, dw_max = taps[0].dw, max_slab = 0;
  for (int t = 0; t < n_taps; ++t) {
    dw_min = std::min(dw_min, taps[t].dw); dw_max = std::max(dw_max, taps[t].dw);
    max_slab = std::max(max_slab, taps[t].slab);
  }
  const int kw = dw_max - dw_min + 1;
  p.kw = kw;
  p.box_w = 128 + kw - 1;
  p.dw0 = dw_min;
  const bool a_lo = x->lo != nullptr, b_lo = w_lo != nullptr;
  int rc;
  if ((rc = make_tmap_nhwc(&p.tmA[0], x->hi, x->n, x->h, x->w, x->c, x->cs, p.box_w, 1, 1))) return rc;
  if (a_lo && (rc = make_tmap_nhwc(&p.tmA[1], x->lo, x->n, x->h, x->w, x->c, x->cs, p.box_w, 1, 1))) return rc;
  if ((rc = make_tmap_weight3(&p.tmB[0], w_hi, max_slab + 1, rows_pad, k_pad, bn, kw))) return rc;
  if (b_lo && (rc = make_tmap_weight3(&p.tmB[1], w_lo, max_slab + 1, rows_pad, k_pad, bn, kw))) return rc;
  // one entry per (filter row, product): the row's taps must be kw consecutive slabs (true for the tap grids
  // hm_conv_fprop / hm_conv_dgrad build: slab = kh*KW + kw')
  bool used[64] = {false};
  int ne = 0;
  for (;;) {
    int first = -1;
    for (int t = 0; t < n_taps; ++t) if (!used[t]) { first = t; break; }
    if (first < 0) break;
    const int dh = taps[first].dh;
    int slab_min = 1 << 30, cnt = 0;
    for (int t = 0; t < n_taps; ++t) if (!used[t] && taps[t].dh == dh) { slab_min = std::min(slab_min, taps[t].slab); ++cnt; }
    if (cnt != kw) return HM_ERR_INVALID;
    int8_t a_off[hm::kRowsMaxKW + 3] = {0};
    for (int t = 0; t < n_taps; ++t) {
      if (used[t] || taps[t].dh != dh) continue;
      used[t] = true;
      const int j = taps[t].slab - slab_min;
      if (j < 0 || j >= kw) return HM_ERR_INVALID;
      a_off[j] = int8_t(taps[t].dw - dw_min);
    }
    const int pa[3] = {0, 1, 0}, pb[3] = {0, 0, 1};
    for (int q = 0; q < 3; ++q) {
      if ((q == 1 && !a_lo) || (q == 2 && !b_lo)) continue;
      if (ne >= hm::kRowsMaxEntries) return HM_ERR_INVALID;
      hm::REntry& e = p.entries[ne++];
      e.a_plane = int8_t(pa[q]); e.b_plane = int8_t(pb[q]); e.dh = int16_t(dh); e.tap0 = slab_min;
      std::memcpy(e.a_off, a_off, sizeof(a_off));
    }
  }
  p.n_entries = ne;
  p.chunks = k_pad / 64;
  p.a_c = x->c;
  p.a_lo_c0 = std::max(0, std::min(x->lo_c0, x->c));
  p.stage_bytes = hm::RCfgCommon::A_PLANE + kw * bn * 128;
  p.n_stages = std::min(hm::RCfgCommon::MAX_STAGES, hm::RCfgCommon::SMEM_BUDGET / p.stage_bytes);
  if (p.n_stages < 2) return HM_ERR_INVALID;
  p.tiles_w = (valid_w + 127) / 128;
  p.rows_h = valid_h;
  p.n_img = x->n;
  p.n_tiles_n = rows_pad / bn;
  p.cout = Cout;
  p.valid_w = valid_w;
  if (out32 && out32->ptr) {
    p.o32 = out32->ptr; p.o32_H = out32->H; p.o32_W = out32->W; p.o32_C = out32->C;
    p.o32_hoff = out32->h_off; p.o32_woff = out32->w_off; p.o32_coff = out32->c_off;
  }
  if (out16 && out16->hi) {
    p.ohi = static_cast<__nv_bfloat16*>(out16->hi); p.olo = static_cast<__nv_bfloat16*>(out16->lo);
    p.o16_H = out16->H; p.o16_W = out16->W; p.o16_C = out16->C;
    p.o16_hoff = out16->h_off; p.o16_woff = out16->w_off; p.o16_coff = out16->c_off;
  }
  p.bias = bias; p.act = act; p.slope = slope; p.err = err_flag;
  // multi-row variant (weight slabs loaded once per MT output rows): separate B (2 stages) and A rings
  static const int use_rows2 = env_int("HM_ROWS2", 1);
  {
    const int mt = bn <= 64 ? 4 : 2;
    const int b_ring = 2 * kw * bn * 128;
    const int a_stages = std::min(8, (hm::RCfgCommon::SMEM_BUDGET - b_ring) / hm::RCfgCommon::A_PLANE);
    if (use_rows2 && a_stages >= 3 && valid_h >= 2 * mt) {
      p.n_stages = a_stages;
      const int groups_h = (p.rows_h + mt - 1) / mt;
      const int tiles2 = p.tiles_w * groups_h * p.n_img * p.n_tiles_n;
      const int smem2 = b_ring + a_stages * hm::RCfgCommon::A_PLANE + 1024 + 512;
      switch (bn) {
        case 16: return launch_rows2<16>(p, tiles2, smem2, st);
        case 32: return launch_rows2<32>(p, tiles2, smem2, st);
        case 64: return launch_rows2<64>(p, tiles2, smem2, st);
        case 128: return launch_rows2<128>(p, tiles2, smem2, st);
      }
    }
  }
  const int num_tiles = p.tiles_w * p.rows_h * p.n_img * p.n_tiles_n;
  const int smem_bytes = p.n_stages * p.stage_bytes + 1024 + 512;
  switch (bn) {
    case 16: return launch_rows<16>(p, num_tiles, smem_bytes, st);
    case 32: return launch_rows<32>(p, num_tiles, smem_bytes, st);
    case 64: return launch_rows<64>(p, num_tiles, smem_bytes, st);
    case 128: return launch_rows<128>(p, num_tiles, smem_bytes, st);
  }
  return HM_ERR_INVALID;
}


template <int NB>
int launch_mnrows(const hm::MRParams& p, int num_tiles, int smem_bytes, cudaStream_t st) {
  static int configured = 0;
  if (configured < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(hm::hm_mnrows_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
    configured = smem_bytes;
  }
  int grid = std::min(num_tiles, sm_count());
  hm::hm_mnrows_kernel<NB><<<grid, hm::kEngineThreads, smem_bytes, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  return HM_OK;
}

// Row-streaming weight-gradient engine (hm_engine_mnrows.cuh): stride-1, wide base space.
int run_mnrows(const hm_operand* P, const hm_operand* Q, int KH, int KW, int pad, float* G_ws, int* err_flag,
               cudaStream_t st) {
  hm::MRParams p;
  std::memset(&p, 0, sizeof(p));
  const int n_units = round_up(Q->c, 64) / 64;
  const int nb = n_units >= 2 ? 2 : 1;
  p.kh = KH; p.kw = KW; p.pairs = (KW + 1) / 2;
  p.box_w = 128 + 2 * p.pairs - 1;
  int rc;
  if ((rc = make_tmap_nhwc(&p.tmP[0], P->hi, P->n, P->h, P->w, P->c, P->cs, p.box_w, 1, 1))) return rc;
  if (P->lo && (rc = make_tmap_nhwc(&p.tmP[1], P->lo, P->n, P->h, P->w, P->c, P->cs, p.box_w, 1, 1))) return rc;
  if ((rc = make_tmap_nhwc(&p.tmQ[0], Q->hi, Q->n, Q->h, Q->w, Q->c, Q->cs, 128, 1, 1))) return rc;
  if (Q->lo && (rc = make_tmap_nhwc(&p.tmQ[1], Q->lo, Q->n, Q->h, Q->w, Q->c, Q->cs, 128, 1, 1))) return rc;
  int np = 0;
  p.prodP[np] = 0; p.prodQ[np] = 0; ++np;
  if (P->lo) { p.prodP[np] = 1; p.prodQ[np] = 0; ++np; }
  if (Q->lo) { p.prodP[np] = 0; p.prodQ[np] = 1; ++np; }
  p.n_prod = np;
  p.tiles_w = (Q->w + 127) / 128;
  p.rows_h = Q->h;
  p.n_img = Q->n;
  p.ktiles = p.tiles_w * p.rows_h * p.n_img;
  p.cp_pad = round_up(P->c, 64);
  p.m_units = p.cp_pad / 64;
  p.n_n_tiles = (n_units + nb - 1) / nb;
  const int base_tiles = KH * p.m_units * p.n_n_tiles;
  int splits = 1;
  if (base_tiles < 2 * sm_count()) splits = (2 * sm_count() + base_tiles - 1) / base_tiles;
  splits = std::max(1, std::min(splits, p.ktiles / 4 > 0 ? p.ktiles / 4 : 1));
  { int per = (p.ktiles + splits - 1) / splits; splits = (p.ktiles + per - 1) / per; }
  p.splits = splits;
  p.dw0 = -pad;
  for (int kh = 0; kh < KH; ++kh) p.dh[kh] = int16_t(kh - pad);
  p.stage_bytes = hm::RCfgCommon::A_PLANE + nb * 128 * 128;
  p.n_stages = std::min(hm::RCfgCommon::MAX_STAGES, hm::RCfgCommon::SMEM_BUDGET / p.stage_bytes);
  p.G = G_ws;
  p.ldG = n_units * 64;
  p.n_cols = n_units * 64;
  p.use_atomic = splits > 1;
  p.err = err_flag;
  if (p.use_atomic) {
    cudaError_t e = cudaMemsetAsync(G_ws, 0, size_t(KH) * KW * p.cp_pad * p.ldG * sizeof(float), st);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  }
  const int num_tiles = base_tiles * splits;
  const int smem_bytes = p.n_stages * p.stage_bytes + 1024 + 512;
  return nb == 1 ? launch_mnrows<1>(p, num_tiles, smem_bytes, st) : launch_mnrows<2>(p, num_tiles, smem_bytes, st);
}

bool mnrows_eligible(const hm_operand* P, const hm_operand* Q, int KW, int stride) {
  static const int enabled = env_int("HM_ROWS", 1);
  if (g_dilated) return false;
  if (!enabled || stride != 1 || KW < 2 || KW > 8 || Q->w < 96) return false;
  // wide layers (>= 4 units on both sides: e.g. the stride-1 256 -> 512 PatchGAN layer at full resolution) belong to the
  // CTA-pair MN-engine: 85 % tensor pipe there against ~40 % for the row-streaming kernel with its 128 x 128 tiles
  if (round_up(Q->c, 64) / 64 >= 4 && KW * ((P->c + 63) / 64) >= 4) return false;
  const int nb = round_up(Q->c, 64) / 64 >= 2 ? 2 : 1;
  return ((KW + 1) / 2) * nb * 64 <= 512;
}

// Common back end of fprop / dgrad: one K-engine launch per output parity class.
int run_k_engine(const hm_operand* x, const void* w_hi, const void* w_lo, int k_pad, int rows_pad, const float* bias,
                 const Tap* taps, int n_taps, int in_stride, int valid_h, int valid_w, int out_sh, int out_sw,
                 int out_oh, int out_ow, int Cout, int act, float slope, const hm_out_f32* out32,
                 const hm_out_bf16* out16, int* err_flag, cudaStream_t st) {
  if (n_taps <= 0 || valid_h <= 0 || valid_w <= 0) return HM_OK;
  const int bn = hm_pick_bn(Cout);
  if (rows_pad % bn || rows_pad < Cout || (k_pad & 63)) return HM_ERR_INVALID;
  if (rows_eligible(taps, n_taps, in_stride, out_sh, out_sw, valid_w, bn))
    return run_rows_engine(x, w_hi, w_lo, k_pad, rows_pad, bias, taps, n_taps, valid_h, valid_w, Cout, act, slope, out32,
                           out16, err_flag, st);
  const bool a_lo = x->lo != nullptr, b_lo = w_lo != nullptr;
  const int prods = 1 + (a_lo ? 1 : 0) + (b_lo ? 1 : 0);
  if (n_taps * prods > hm::kMaxEntries) return HM_ERR_INVALID;

  hm::KParams p;
  std::memset(&p, 0, sizeof(p));
  int tw = 128, th = 1;
  pick_rect(128, valid_h, valid_w, 256 / in_stride, &tw, &th);
  int rc;
  if ((rc = make_tmap_nhwc(&p.tmA[0], x->hi, x->n, x->h, x->w, x->c, x->cs, tw, th, in_stride))) return rc;
  if (a_lo && (rc = make_tmap_nhwc(&p.tmA[1], x->lo, x->n, x->h, x->w, x->c, x->cs, tw, th, in_stride))) return rc;
  // total rows of the weight matrix: one rows_pad slab per tap slot referenced
  int max_slab = 0;
  for (int t = 0; t < n_taps; ++t) max_slab = std::max(max_slab, taps[t].slab);
  const int rows_total = (max_slab + 1) * rows_pad;
  static const int use_2cta = env_int("HM_2CTA", 1);
  const bool pair = use_2cta && bn == 256;   // CTA-pair kernel: each CTA loads half (128 rows) of the weight slab
  if ((rc = make_tmap_weight(&p.tmB[0], w_hi, rows_total, k_pad, pair ? 128 : bn))) return rc;
  if (b_lo && (rc = make_tmap_weight(&p.tmB[1], w_lo, rows_total, k_pad, pair ? 128 : bn))) return rc;

  // fused-split kernel (one stage = hi + lo boxes of both operands, three products per stage): bf16x3 with N tile <= 128
  static const int use_fused3 = env_int("HM_FUSED3", 1);
  static const int use_fused3_pair = env_int("HM_FUSED3_PAIR", 1);
  const bool fused3 = use_fused3 && a_lo && b_lo && (pair ? use_fused3_pair != 0 : bn <= 128);
  int ne = 0;
  for (int t = 0; t < n_taps; ++t) {
    const int pa[3] = {0, 1, 0}, pb[3] = {0, 0, 1};
    for (int q = 0; q < (fused3 ? 1 : 3); ++q) {
      if (q == 1 && !a_lo) continue;
      if (q == 2 && !b_lo) continue;
      hm::KEntry& e = p.entries[ne++];
      e.a_plane = int8_t(pa[q]); e.b_plane = int8_t(pb[q]);
      e.dw = int16_t(taps[t].dw); e.dh = int16_t(taps[t].dh); e.pad_ = 0;
      e.b_row = taps[t].slab * rows_pad;
    }
  }
  p.n_entries = ne;
  p.chunks = k_pad / 64;
  p.a_c = x->c;
  p.a_lo_c0 = std::max(0, std::min(x->lo_c0, x->c));
  p.tiles_w = (valid_w + tw - 1) / tw;
  p.tiles_h = (valid_h + th - 1) / th;
  p.n_img = x->n;
  p.n_tiles_n = rows_pad / bn;
  p.tw_log2 = ilog2(tw);
  p.th = th;
  p.in_stride = in_stride;
  p.cout = Cout;
  p.valid_h = valid_h; p.valid_w = valid_w;
  p.out_sh = out_sh; p.out_sw = out_sw; p.out_oh = out_oh; p.out_ow = out_ow;
  if (out32 && out32->ptr) {
    p.o32 = out32->ptr; p.o32_H = out32->H; p.o32_W = out32->W; p.o32_C = out32->C;
    p.o32_hoff = out32->h_off; p.o32_woff = out32->w_off; p.o32_coff = out32->c_off;
  }
  if (out16 && out16->hi) {
    p.ohi = static_cast<__nv_bfloat16*>(out16->hi); p.olo = static_cast<__nv_bfloat16*>(out16->lo);
    p.o16_H = out16->H; p.o16_W = out16->W; p.o16_C = out16->C;
    p.o16_hoff = out16->h_off; p.o16_woff = out16->w_off; p.o16_coff = out16->c_off;
  }
  p.bias = bias; p.act = act; p.slope = slope; p.err = err_flag;
  const int num_tiles = p.tiles_w * p.tiles_h * p.n_img * p.n_tiles_n;
  if (pair) return launch_k2(p, p.tiles_w * p.tiles_h * p.n_img, p.n_tiles_n, fused3, st);
  if (fused3) return launch_k3_bn(bn, p, num_tiles, st);
  return launch_k_bn(bn, p, num_tiles, st);
}

}  // namespace

extern "C" {

const char* hm_version(void) { return "hm_b200 0.2 (sm_100a, tcgen05+TMA)"; }

size_t hm_scratch_bytes(void) {
  const int keep = g_sm_limit;
  g_sm_limit = 0;
  const size_t b = kSkCounterBytes + size_t(sm_count() / 2) * 2 * kSkSlotBytes;
  g_sm_limit = keep;
  return b;
}
int hm_set_sm_limit(int n) { g_sm_limit = n > 0 ? (n & ~1) : 0; return HM_OK; }
int hm_set_streamk(int on) { g_use_streamk = on ? 1 : 0; return HM_OK; }
int hm_set_scratch(void* ptr, size_t bytes) {
  if (ptr && (bytes < hm_scratch_bytes() || (reinterpret_cast<uintptr_t>(ptr) & 255))) return HM_ERR_INVALID;
  g_scratch = ptr;
  g_scratch_bytes = ptr ? bytes : 0;
  return HM_OK;
}
int hm_last_cuda_error(void) { return g_last_cuda_error; }

int hm_pick_bn(int rows) {
  if (rows > 128) return 256;
  if (rows > 64) return 128;
  if (rows > 32) return 64;
  if (rows > 16) return 32;
  return 16;
}
int hm_rows_pad(int rows) { return round_up(rows, hm_pick_bn(rows)); }
int hm_k_pad(int k) { return round_up(k, 64); }

int hm_conv_fprop(const hm_operand* x, const void* w_hi, const void* w_lo, int k_pad, int rows_pad,
                  const float* bias, int KH, int KW, int stride, int pad, int Hout, int Wout, int Cout, int act,
                  float slope, const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, void* stream) {
  return hm_conv_fprop_dil(x, w_hi, w_lo, k_pad, rows_pad, bias, KH, KW, stride, pad, 1, Hout, Wout, Cout, act, slope, out32,
                           out16, err_flag, stream);
}

int hm_conv_fprop_dil(const hm_operand* x, const void* w_hi, const void* w_lo, int k_pad, int rows_pad,
                      const float* bias, int KH, int KW, int stride, int pad, int dilation, int Hout, int Wout, int Cout,
                      int act, float slope, const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, void* stream) {
  if (!x || !x->hi || !w_hi || KH * KW > 64 || (stride != 1 && stride != 2) || dilation < 1) return HM_ERR_INVALID;
  Tap taps[64];
  int nt = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) taps[nt++] = Tap{kw * dilation - pad, kh * dilation - pad, kh * KW + kw};
  g_dilated = dilation > 1;
  const int rc = run_k_engine(x, w_hi, w_lo, k_pad, rows_pad, bias, taps, nt, stride, Hout, Wout, 1, 1, 0, 0, Cout, act,
                              slope, out32, out16, err_flag, static_cast<cudaStream_t>(stream));
  g_dilated = false;
  return rc;
}

int hm_conv_dgrad(const hm_operand* dy, const void* w_hi, const void* w_lo, int k_pad, int rows_pad,
                  const float* bias, int KH, int KW, int stride, int pad, int Hout, int Wout, int Cout, int act,
                  float slope, const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, void* stream) {
  return hm_conv_dgrad_dil(dy, w_hi, w_lo, k_pad, rows_pad, bias, KH, KW, stride, pad, 1, Hout, Wout, Cout, act, slope, out32,
                           out16, err_flag, stream);
}

int hm_conv_dgrad_dil(const hm_operand* dy, const void* w_hi, const void* w_lo, int k_pad, int rows_pad,
                      const float* bias, int KH, int KW, int stride, int pad, int dilation, int Hout, int Wout, int Cout,
                      int act, float slope, const hm_out_f32* out32, const hm_out_bf16* out16, int* err_flag, void* stream) {
  if (!dy || !dy->hi || !w_hi || KH * KW > 64 || (stride != 1 && stride != 2) || dilation < 1 || (dilation > 1 && stride != 1))
    return HM_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Tap taps[64];
  if (stride == 1) {
    int nt = 0;
    for (int kh = 0; kh < KH; ++kh)
      for (int kw = 0; kw < KW; ++kw) taps[nt++] = Tap{pad - kw * dilation, pad - kh * dilation, kh * KW + kw};
    g_dilated = dilation > 1;
    const int rc = run_k_engine(dy, w_hi, w_lo, k_pad, rows_pad, bias, taps, nt, 1, Hout, Wout, 1, 1, 0, 0, Cout, act, slope,
                                out32, out16, err_flag, st);
    g_dilated = false;
    return rc;
  }
  // stride 2: four output parity classes, each a dense unit-stride tap-GEMM over its own tap subset
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      int nt = 0;
      for (int kh = 0; kh < KH; ++kh) {
        if ((ph + pad - kh) & 1) continue;
        for (int kw = 0; kw < KW; ++kw) {
          if ((pw + pad - kw) & 1) continue;
          taps[nt++] = Tap{(pw + pad - kw) / 2, (ph + pad - kh) / 2, kh * KW + kw};
        }
      }
      // a class without taps (1x1 stride-2 convolutions: only the even / even positions receive gradient) produces
      // nothing: the caller passes a zero-filled destination for such kernels
      if (nt == 0) { if (KH * KW > 1) return HM_ERR_INVALID; continue; }
      const int vh = (Hout - ph + 1) / 2, vw = (Wout - pw + 1) / 2;
      // NOTE: all slabs must be addressable -> rows_total is sized from the largest slab id used by this class;
      // the caller's buffer always holds KH*KW slabs.
      int rc = run_k_engine(dy, w_hi, w_lo, k_pad, rows_pad, bias, taps, nt, 1, vh, vw, 2, 2, ph, pw, Cout, act, slope,
                            out32, out16, err_flag, st);
      if (rc) return rc;
    }
  return HM_OK;
}

size_t hm_wgrad_ws_bytes(int KH, int KW, int cp, int cq) {
  return size_t(KH) * KW * round_up(cp, 64) * size_t(round_up(cq, 64)) * sizeof(float);
}

int hm_conv_wgrad(const hm_operand* P, const hm_operand* Q, int KH, int KW, int stride, int pad, float* G_ws,
                  int* err_flag, void* stream) {
  return hm_conv_wgrad_dil(P, Q, KH, KW, stride, pad, 1, G_ws, err_flag, stream);
}

int hm_conv_wgrad_dil(const hm_operand* P, const hm_operand* Q, int KH, int KW, int stride, int pad, int dilation, float* G_ws,
                      int* err_flag, void* stream) {
  if (!P || !Q || !P->hi || !Q->hi || !G_ws || KH * KW > 64 || (stride != 1 && stride != 2) || dilation < 1 ||
      (dilation > 1 && stride != 1))
    return HM_ERR_INVALID;
  if (P->n != Q->n) return HM_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  g_dilated = dilation > 1;
  const bool rows = mnrows_eligible(P, Q, KW, stride);
  g_dilated = false;
  if (rows) return run_mnrows(P, Q, KH, KW, pad, G_ws, err_flag, st);
  hm::MNParams p;
  std::memset(&p, 0, sizeof(p));
  int tw = 64, th = 1;
  pick_rect(64, Q->h, Q->w, 256 / stride, &tw, &th);
  int rc;
  if ((rc = make_tmap_nhwc(&p.tmP[0], P->hi, P->n, P->h, P->w, P->c, P->cs, tw, th, stride))) return rc;
  if (P->lo && (rc = make_tmap_nhwc(&p.tmP[1], P->lo, P->n, P->h, P->w, P->c, P->cs, tw, th, stride))) return rc;
  if ((rc = make_tmap_nhwc(&p.tmQ[0], Q->hi, Q->n, Q->h, Q->w, Q->c, Q->cs, tw, th, 1))) return rc;
  if (Q->lo && (rc = make_tmap_nhwc(&p.tmQ[1], Q->lo, Q->n, Q->h, Q->w, Q->c, Q->cs, tw, th, 1))) return rc;
  int np = 0;
  p.pairP[np] = 0; p.pairQ[np] = 0; ++np;
  if (P->lo) { p.pairP[np] = 1; p.pairQ[np] = 0; ++np; }
  if (Q->lo) { p.pairP[np] = 0; p.pairQ[np] = 1; ++np; }
  p.n_pairs = np;
  p.tiles_w = (Q->w + tw - 1) / tw;
  p.tiles_h = (Q->h + th - 1) / th;
  p.n_img = Q->n;
  p.tw_log2 = ilog2(tw);
  p.th = th;
  p.sP = stride; p.sQ = 1;
  p.m_tapped = 1;
  p.upt_m = round_up(P->c, 64) / 64;
  p.n_units = round_up(Q->c, 64) / 64;
  p.upt_n = p.n_units;
  p.m_units = KH * KW * p.upt_m;
  const int nb = p.n_units >= 3 ? 4 : p.n_units;
  static const int use_mn2 = env_int("HM_MN2", 1);
  const bool pair = use_mn2 && p.n_units >= 4 && p.m_units >= 4;   // CTA-pair engine: 256 x 256 tiles of G
  p.n_m_tiles = pair ? (p.m_units + 3) / 4 : (p.m_units + 1) / 2;
  p.n_n_tiles = (p.n_units + nb - 1) / nb;
  p.ktiles = p.tiles_w * p.tiles_h * p.n_img;
  const int base_tiles = p.n_m_tiles * p.n_n_tiles;
  const int workers = pair ? sm_count() / 2 : sm_count();
  int splits = 1;
  if (base_tiles < 2 * workers) splits = (2 * workers + base_tiles - 1) / base_tiles;
  splits = std::max(1, std::min(splits, p.ktiles / 4 > 0 ? p.ktiles / 4 : 1));
  // make every split non-empty
  { int per = (p.ktiles + splits - 1) / splits; splits = (p.ktiles + per - 1) / per; }
  p.splits = splits;
  p.dwP0 = p.dhP0 = p.dwQ0 = p.dhQ0 = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      p.tap_dw[kh * KW + kw] = int16_t(kw * dilation - pad);
      p.tap_dh[kh * KW + kw] = int16_t(kh * dilation - pad);
    }
  p.G = G_ws;
  p.ldG = p.n_units * 64;
  p.use_atomic = splits > 1;
  p.err = err_flag;
  if (p.use_atomic) {
    cudaError_t e = cudaMemsetAsync(G_ws, 0, size_t(p.m_units) * 64 * p.ldG * sizeof(float), st);
    if (e != cudaSuccess) { g_last_cuda_error = int(e); return HM_ERR_LAUNCH; }
  }
  const int num_tiles = base_tiles * splits;
  static const int use_fused3 = env_int("HM_FUSED3", 1);
  static const int use_fused3_pair = env_int("HM_FUSED3_PAIR", 1);
  if (pair) return launch_mn2(p, num_tiles, use_fused3 && use_fused3_pair && P->lo && Q->lo, st);
  if (use_fused3 && P->lo && Q->lo && nb <= 2) return nb == 1 ? launch_mn3<1>(p, num_tiles, st) : launch_mn3<2>(p, num_tiles, st);
  switch (nb) {
    case 1: return launch_mn<1>(p, num_tiles, st);
    case 2: return launch_mn<2>(p, num_tiles, st);
    default: return launch_mn<4>(p, num_tiles, st);
  }
}

}  // extern "C"
