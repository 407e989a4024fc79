// Diagnostic micro-benchmarks of tcgen05.mma issue / commit costs (not part of the product path; used to
// derive the pipeline structure of rnn_tc.cu / gemm_tc.cu / conv_tc.cu -- see DESIGN.md section 5).
#include "tc_common.cuh"

namespace dsb {
namespace tc {

// variant bit 0: tcgen05.commit (to a scratch mbarrier) after every `group` MMAs
// variant bit 1: tcgen05.fence::after_thread_sync before every group
// variant bit 2: mbarrier.try_wait on an already-completed barrier before every group
// variant bit 3: warp-uniform control flow with elect_one_sync instead of a divergent `lane == 0` region
template <int M, int N>
__global__ void __launch_bounds__(128, 1) mma_issue_bench_kernel(int n_mma, int group, int variant, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);   // scratch: receives the per-group commits (never waited on)
    mbar_init(&bars[1], 1);   // done barrier
    mbar_init(&bars[2], 1);   // pre-completed barrier for the try_wait variant
    fence_mbar_init();
    mbar_arrive(&bars[2]);
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  fence_proxy_async();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1 && (variant & 8)) {
    // warp-uniform control flow + elect_one_sync (the structure the kernels use)
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t idesc = make_idesc_bf16(M, N);
    const uint64_t adesc = make_smem_desc(smem_u32(smem), 16, 1024, 2);
    const uint64_t bdesc = make_smem_desc(smem_u32(smem + 16384), 16, 1024, 2);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += group) {
      if (variant & 4) mbar_wait(&bars[2], 0);
      if (variant & 2) tc_fence_after();
      if (elect_one_sync()) {
        for (int k = 0; k < group; ++k)
          umma_bf16(tb, adesc + (uint64_t)((k & 3) * 2), bdesc + (uint64_t)((k & 3) * 2), idesc, 1);
        if (variant & 1) umma_commit(&bars[0]);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one_sync()) umma_commit(&bars[1]);
    __syncwarp();
    mbar_wait(&bars[1], 0);
    const long long t2 = clock64();
    if (lane == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N);
    const uint64_t adesc = make_smem_desc(smem_u32(smem), 16, 1024, 2);
    const uint64_t bdesc = make_smem_desc(smem_u32(smem + 16384), 16, 1024, 2);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += group) {
      if (variant & 4) mbar_wait(&bars[2], 0);
      if (variant & 2) tc_fence_after();
      for (int k = 0; k < group; ++k)
        umma_bf16(tmem_base, adesc + (uint64_t)((k & 3) * 2), bdesc + (uint64_t)((k & 3) * 2), idesc, 1);
      if (variant & 1) umma_commit(&bars[0]);
    }
    const long long t1 = clock64();
    umma_commit(&bars[1]);
    mbar_wait(&bars[1], 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;   // issue time
    out[1] = t2 - t0;   // until all MMAs completed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace tc
}  // namespace dsb

extern "C" int dsb_debug_mma_bench(int M, int N, int n_mma, int group, int variant, long long* host_out) {
  using namespace dsb;
  long long* d = nullptr;
  DSB_CUDA(cudaMalloc(&d, 2 * sizeof(long long)));
  const int smem = 65536 + 1024;
#define RUN(MM, NN)                                                                                     \
  DSB_CUDA(cudaFuncSetAttribute(tc::mma_issue_bench_kernel<MM, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
  tc::mma_issue_bench_kernel<MM, NN><<<1, 128, smem>>>(n_mma, group, variant, d);
  if (M == 64 && N == 64) { RUN(64, 64) }
  else if (M == 128 && N == 64) { RUN(128, 64) }
  else if (M == 128 && N == 32) { RUN(128, 32) }
  else if (M == 128 && N == 96) { RUN(128, 96) }
  else if (M == 128 && N == 128) { RUN(128, 128) }
  else if (M == 128 && N == 240) { RUN(128, 240) }
  else if (M == 128 && N == 256) { RUN(128, 256) }
  else { cudaFree(d); return set_error(DSB_ERR_UNSUPPORTED, "dsb_debug_mma_bench: shape"); }
#undef RUN
  DSB_CUDA(cudaDeviceSynchronize());
  DSB_CUDA(cudaMemcpy(host_out, d, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}
