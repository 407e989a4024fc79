// Diagnostic micro-benchmarks of tcgen05.mma issue / commit costs (not part of the product path; used to
// derive the pipeline structure of rnn_tc.cu / gemm_tc.cu / conv_tc.cu -- see DESIGN.md section 5).
#include "tc_common.cuh"   // from danspeech_b200/csrc (-I)

namespace dsb {
namespace tc {

// variant bit 0: tcgen05.commit (to a scratch mbarrier) after every `group` MMAs
// variant bit 1: tcgen05.fence::after_thread_sync before every group
// variant bit 2: mbarrier.try_wait on an already-completed barrier before every group
// variant bit 3: warp-uniform control flow with elect_one_sync instead of a divergent `lane == 0` region
// variant bit 4: (with bit 3) the A operand comes from tensor memory (TS form) instead of a shared-memory descriptor
// D[tmem] (+)= A[tmem] * B[smem desc]: the TS form of tcgen05.mma (A: M lanes x 8 columns of packed bf16 pairs per K = 16)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

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
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
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
        if (variant & 16) {
          for (int k = 0; k < group; ++k)
            umma_bf16_ts(tb, tb + 256u + (uint32_t)((k & 15) * 8), bdesc + (uint64_t)((k & 3) * 2), idesc, 1);
        } else {
          for (int k = 0; k < group; ++k)
            umma_bf16(tb, adesc + (uint64_t)((k & 3) * 2), bdesc + (uint64_t)((k & 3) * 2), idesc, 1);
        }
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
    tmem_dealloc<512>(tmem_base);
  }
}


// ---- TMEM accumulator layout probe: one tcgen05.mma with A[m][0] = m+1, A[m][1] = 1, B[n][0] = 1,
// B[n][1] = 256 (n+1), so that D[m][n] = (m+1) + 256 (n+1) identifies (m, n); every CTA dumps its 128 TMEM lanes
// x 128 columns.  Used to derive the epilogue mapping of the cta_group::2 M=128 recurrence variant.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <int CG>
__global__ void __launch_bounds__(128, 1) layout_probe_kernel(int M, int N, float* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  unsigned char* sA = smem;
  unsigned char* sB = smem + 16384;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  const int rows_a = M / CG, rows_b = N / CG;
  for (int r = threadIdx.x; r < rows_a; r += blockDim.x) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(sA + r * 128 + ((0 ^ (r & 7)) << 4));
    p[0] = __float2bfloat16_rn((float)(rank * rows_a + r + 1));
    p[1] = __float2bfloat16_rn(1.0f);
  }
  for (int r = threadIdx.x; r < rows_b; r += blockDim.x) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(sB + r * 128 + ((0 ^ (r & 7)) << 4));
    p[0] = __float2bfloat16_rn(1.0f);
    p[1] = __float2bfloat16_rn(256.0f * (float)(rank * rows_b + r + 1));
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    if (CG == 2) tmem_alloc_2cta<128>(&tmem_slot);
    else tmem_alloc<128>(&tmem_slot);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  {  // sentinel in every lane / column
    uint32_t neg[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) neg[j] = __float_as_uint(-1.0f);
    for (int c0 = 0; c0 < 128; c0 += 32) tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, neg);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24));
    const uint64_t adesc = make_smem_desc(smem_u32(sA), 16, 1024, 2);
    const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 1024, 2);
    if (elect_one_sync()) {
      if (CG == 2) {
        umma_bf16_2cta(tmem_base, adesc, bdesc, idesc, 0);
        umma_commit_2cta(&bar, 3);
      } else {
        umma_bf16(tmem_base, adesc, bdesc, idesc, 0);
        umma_commit(&bar);
      }
    }
    __syncwarp();
  }
  mbar_wait_trap(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[((size_t)rank * 128 + warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2cta<128>(tmem_base);
    else tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace tc
}  // namespace dsb

// host_out: [cg][128 lanes][128 columns] floats; -1 = never written by the MMA
extern "C" int dsb_debug_layout_probe(int cg, int M, int N, float* host_out) {
  using namespace dsb;
  if ((cg != 1 && cg != 2) || N > 128 || N % 16 || (M != 64 && M != 128 && M != 256)) return set_error(DSB_ERR_UNSUPPORTED, "dsb_debug_layout_probe: shape");
  float* d = nullptr;
  const size_t n = (size_t)cg * 128 * 128;
  DSB_CUDA(cudaMalloc(&d, n * sizeof(float)));
  DSB_CUDA(cudaMemset(d, 0, n * sizeof(float)));
  const int smem = 32768 + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cg);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const void* fn = cg == 2 ? (const void*)tc::layout_probe_kernel<2> : (const void*)tc::layout_probe_kernel<1>;
  DSB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  void* args[] = {(void*)&M, (void*)&N, (void*)&d};
  DSB_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  DSB_CUDA(cudaDeviceSynchronize());
  DSB_CUDA(cudaMemcpy(host_out, d, n * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

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
