// hm_debug.cu -- micro-benchmarks used while tuning the engines (NOT part of libhm_b200.so: tools/mma_issue_bench.py
// compiles this file into tools/libhm_debug.so on demand).
#include "../neurips18_hierchical_image_manipulation_b200/csrc/hm_ptx.cuh"
#include "../neurips18_hierchical_image_manipulation_b200/csrc/hm_engine2.cuh"

namespace {
// one CTA issues `count` back-to-back tcgen05.mma (M=128, N, K=16) on arbitrary smem and reports
// cycles until the last one completed and cycles the issuing thread spent issuing
template <int N>
__global__ void mma_issue_kernel(int count, int a_off_rows, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { hm::mbar_init(&bar, 1); hm::fence_barrier_init(); }
  if (threadIdx.x < 32) hm::tmem_alloc(&slot, 256);
  hm::tc_fence_before();
  __syncthreads();
  hm::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = hm::umma_idesc_bf16(128, N, 0, 0);
    const uint64_t ad = hm::umma_smem_desc(hm::smem_u32(smem) + a_off_rows * 128, 16, 1024);
    const uint64_t bd = hm::umma_smem_desc(hm::smem_u32(smem + 32768), 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) hm::umma_bf16(tm, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1u);
    long long t1 = clock64();
    hm::umma_commit(&bar);
    while (!hm::mbar_try_wait(&bar, 0)) {}
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  hm::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) hm::tmem_dealloc(tm, 256);
}
// MN-major variant (both operands pixel-major boxes as the weight-gradient engines use them): M=128, N, K=16
template <int N>
__global__ void mma_issue_mn_kernel(int count, int lbo_bytes, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { hm::mbar_init(&bar, 1); hm::fence_barrier_init(); }
  if (threadIdx.x < 32) hm::tmem_alloc(&slot, 256);
  hm::tc_fence_before();
  __syncthreads();
  hm::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = hm::umma_idesc_bf16(128, N, 1, 1);
    const uint64_t ad = hm::umma_smem_desc(hm::smem_u32(smem), lbo_bytes, 1024);
    const uint64_t bd = hm::umma_smem_desc(hm::smem_u32(smem + 40960), 8192, 1024);
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) hm::umma_bf16(tm, ad + (i & 3) * (2048 >> 4), bd + (i & 3) * (2048 >> 4), idesc, 1u);
    long long t1 = clock64();
    hm::umma_commit(&bar);
    while (!hm::mbar_try_wait(&bar, 0)) {}
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  hm::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) hm::tmem_dealloc(tm, 256);
}
// CTA-pair variant: leader issues `count` tcgen05.mma.cta_group::2 (M=256, N=256, K=16)
__global__ void __cluster_dims__(2, 1, 1) mma2_issue_kernel(int count, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const bool leader = hm::cluster_ctarank() == 0;
  if (threadIdx.x == 0) { hm::mbar_init(&bar, 1); hm::fence_barrier_init(); }
  if (threadIdx.x < 32) hm::tmem_alloc_2sm(&slot, 256);
  hm::tc_fence_before();
  hm::cluster_sync_all();
  hm::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0 && leader) {
    const uint32_t idesc = hm::umma_idesc_bf16(256, 256, 0, 0);
    const uint64_t ad = hm::umma_smem_desc(hm::smem_u32(smem), 16, 1024);
    const uint64_t bd = hm::umma_smem_desc(hm::smem_u32(smem + 32768), 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < count; ++i) hm::umma_bf16_2sm(tm, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1u);
    long long t1 = clock64();
    hm::umma_commit_2sm_mc(&bar, 1);
    while (!hm::mbar_try_wait(&bar, 0)) {}
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  hm::tc_fence_before();
  hm::cluster_sync_all();
  if (threadIdx.x < 32) hm::tmem_dealloc_2sm(tm, 256);
}
}  // namespace

extern "C" int hm_debug_mma2_issue(int count, long long* out_dev, void* stream) {
  const int smem = 96 * 1024;
  cudaFuncSetAttribute(mma2_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma2_issue_kernel<<<2, 64, smem, static_cast<cudaStream_t>(stream)>>>(count, out_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int hm_debug_mma_issue(int n, int count, int a_off_rows, long long* out_dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int smem = 96 * 1024;
  switch (n) {
    case 16:
      cudaFuncSetAttribute(mma_issue_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      mma_issue_kernel<16><<<1, 64, smem, st>>>(count, a_off_rows, out_dev); break;
    case 64:
      cudaFuncSetAttribute(mma_issue_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      mma_issue_kernel<64><<<1, 64, smem, st>>>(count, a_off_rows, out_dev); break;
    case 256:
      cudaFuncSetAttribute(mma_issue_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      mma_issue_kernel<256><<<1, 64, smem, st>>>(count, a_off_rows, out_dev); break;
    default: return -1;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

extern "C" int hm_debug_mma_issue_mn(int n, int count, int lbo_bytes, long long* out_dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int smem = 112 * 1024;
  switch (n) {
    case 64:
      cudaFuncSetAttribute(mma_issue_mn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      mma_issue_mn_kernel<64><<<1, 64, smem, st>>>(count, lbo_bytes, out_dev); break;
    case 128:
      cudaFuncSetAttribute(mma_issue_mn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      mma_issue_mn_kernel<128><<<1, 64, smem, st>>>(count, lbo_bytes, out_dev); break;
    case 256:
      cudaFuncSetAttribute(mma_issue_mn_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      mma_issue_mn_kernel<256><<<1, 64, smem, st>>>(count, lbo_bytes, out_dev); break;
    default: return -1;
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
