// bf16 tensor-core GEMM for sm_100a: C[M,N] = A[M,K] * W[N,K]^T + bias[N]   (fp32 accumulate/output).
//
// This is the BatchRNN input projection (x_t * W_ih^T + b_ih for every time step at once -- the
// input half of torch.nn.GRU/LSTM/RNN behind model.py:107-108,118, with the eval BatchNorm1d of
// model.py:106,115-116 folded into W_ih/b_ih) on the 5th-generation tensor cores:
//   * persistent CTAs (one per SM) walking a static tile schedule,
//   * warp 0: TMA producer (cp.async.bulk.tensor, 128B-swizzled K-major tiles, 4-stage mbarrier ring),
//   * warp 1: single-thread tcgen05.mma issuer, fp32 accumulators double-buffered in TMEM,
//   * warps 2-5: epilogue, tcgen05.ld -> +bias -> vectorised global stores, overlapped with the
//     next tile's MMAs.
// Tensor-pipe bound: 2*M*N*K flop per call.
#include "tc_common.cuh"
#include "model_types.cuh"
#include <mutex>
#include <cstdlib>

namespace dsb {
namespace tc {

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return set_error(DSB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(DSB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

constexpr int BM = 128, BK = 64, STAGES = 4;
constexpr int GEMM_THREADS = 192;


// Epilogue of one 128-row accumulator tile: this thread owns one row (TMEM lane) and walks the BN columns in
// chunks of 16, with the next tcgen05.ld in flight while the current chunk is stored.  Rows are written with
// 256-bit stores (one full 32-byte sector per thread per instruction) when the output is 32-byte aligned.
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
// bias_row: the bias belongs to the ROWS of C (operands swapped: C^T = W * A^T, see gemm_bias_rows_tc).  The kernels
// then also walk the ROW blocks fastest: the rows are the weights (small, L2-resident), the columns the activations,
// which should stream from HBM once and not once per row block.
__device__ __forceinline__ void store_chunk16(const uint32_t (&r)[16], float* __restrict__ C, int64_t ldc, int row, int M,
                                              int N, int col, const float* __restrict__ bias, bool vec_ok, bool bias_row) {
  if (row >= M) return;
  float* dst = C + (int64_t)row * ldc + col;
  const float rb = (bias && bias_row) ? __ldg(bias + row) : 0.f;
  const float* cb = (bias && !bias_row) ? bias : nullptr;
  if (vec_ok && col + 16 <= N) {
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[j + i]) + (cb ? __ldg(cb + col + j + i) : rb);
      st_global_v8(dst + j, v);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col + j < N) dst[j] = __uint_as_float(r[j]) + (cb ? __ldg(cb + col + j) : rb);
  }
}
template <int BN>
__device__ __forceinline__ void epilogue_tile(uint32_t t_addr, float* __restrict__ C, int64_t ldc, int row, int M, int N,
                                              int n0, const float* __restrict__ bias, bool vec_ok, bool bias_row) {
  static_assert(BN % 16 == 0, "BN must be a multiple of 16");
  uint32_t ra[16], rb[16];
  tmem_ld16(t_addr, ra);
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    tmem_ld_wait_dep(ra);
    if (c0 + 16 < BN) tmem_ld16(t_addr + c0 + 16, rb);
    store_chunk16(ra, C, ldc, row, M, N, n0 + c0, bias, vec_ok, bias_row);
    if (c0 + 16 < BN) {
      tmem_ld_wait_dep(rb);
      if (c0 + 32 < BN) tmem_ld16(t_addr + c0 + 32, ra);
      store_chunk16(rb, C, ldc, row, M, N, n0 + c0 + 16, bias, vec_ok, bias_row);
    }
  }
}

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int B_STRIDE = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int BAR_OFF = STAGES * (A_BYTES + B_STRIDE);
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;   // barriers + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N, int K, int bias_row) {
  using S = GemmSmem<BN>;
  constexpr int TMEM_COLS = 512;
  constexpr int ACC_STRIDE = 256;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + STAGES * S::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
  const int n_tiles = m_blocks * n_blocks;
  const int nkb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // whole warp, warp-uniform control flow; one elected lane issues (see elect_one_sync)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mb = bias_row ? tile % m_blocks : tile / n_blocks, nb = bias_row ? tile / m_blocks : tile % n_blocks;
      const int m0 = mb * BM, n0 = nb * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&full[stage], S::A_BYTES + S::B_BYTES);
          tma_load_2d(sA + stage * S::A_BYTES, &tmap_a, &full[stage], kb * BK, m0);
          tma_load_2d(sB + stage * S::B_STRIDE, &tmap_b, &full[stage], kb * BK, n0);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(smem_u32(sA + stage * S::A_BYTES), 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(smem_u32(sB + stage * S::B_STRIDE), 16, 1024, 2);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (kb == nkb - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(C) & 31) == 0 && (BN & 7) == 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int mb = bias_row ? tile % m_blocks : tile / n_blocks, nb = bias_row ? tile / m_blocks : tile % n_blocks;
      const int m0 = mb * BM, n0 = nb * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_addr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      epilogue_tile<BN>(t_addr, C, ldc, row, M, N, n0, bias, vec_ok, bias_row != 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes a 256 x BN tile.  Each CTA stages its own
// 128 rows of A and BN/2 rows of W, the leader issues one tcgen05.mma per K=16 slice for both SMs, each
// CTA's TMEM holds its 128 accumulator rows.  The SS-mode operand fetch (~64 B/clk/SM) per slice drops
// from (128 + BN) to (128 + BN/2) rows, which is what bounds the 1-CTA kernel.
// ------------------------------------------------------------------------------------------------
template <int BN>
struct Gemm2Smem {
  static constexpr int STAGES2 = 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int B_STRIDE = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int BAR_OFF = STAGES2 * (A_BYTES + B_STRIDE);
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N, int K, int bias_row) {
  using S = Gemm2Smem<BN>;
  constexpr int STAGES2 = S::STAGES2;
  constexpr int TMEM_COLS = 512;
  constexpr int ACC_STRIDE = 256;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + STAGES2 * S::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // used in the leader only
  uint64_t* empty = full + STAGES2;                                  // per CTA (multicast commit)
  uint64_t* tfull = empty + STAGES2;                                 // per CTA (multicast commit)
  uint64_t* tempty = tfull + 2;                                      // leader only, 8 arrivals (4 warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_blocks = (M + 2 * BM - 1) / (2 * BM), n_blocks = (N + BN - 1) / BN;
  const int n_tiles = m_blocks * n_blocks;
  const int nkb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // producer of this CTA's halves; completion is signalled on the LEADER's full barrier
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int mb = bias_row ? tile % m_blocks : tile / n_blocks, nb = bias_row ? tile / m_blocks : tile % n_blocks;
      const int m0 = mb * 2 * BM + rank * BM, n0 = nb * BN + rank * (BN / 2);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_trap(&empty[stage], phase ^ 1);
        if (elect_one_sync()) {
          const uint32_t lead_bar = mapa_u32(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * (S::A_BYTES + S::B_BYTES));
          tma_load_2d_2cta(sA + stage * S::A_BYTES, &tmap_a, lead_bar, kb * BK, m0);
          tma_load_2d_2cta(sB + stage * S::B_STRIDE, &tmap_b, lead_bar, kb * BK, n0);
        }
        __syncwarp();
        if (++stage == STAGES2) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // leader: issues the MMAs of the pair
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        mbar_wait_trap(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_trap(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(sA + stage * S::A_BYTES), 16, 1024, 2);
          const uint64_t bdesc = make_smem_desc(smem_u32(sB + stage * S::B_STRIDE), 16, 1024, 2);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2cta(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            umma_commit_2cta(&empty[stage], 3);
            if (kb == nkb - 1) umma_commit_2cta(&tfull[acc], 3);
          }
          __syncwarp();
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(C) & 31) == 0 && (BN & 7) == 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int mb = bias_row ? tile % m_blocks : tile / n_blocks, nb = bias_row ? tile / m_blocks : tile % n_blocks;
      const int m0 = mb * 2 * BM + rank * BM, n0 = nb * BN;
      mbar_wait_trap(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const uint32_t t_addr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      epilogue_tile<BN>(t_addr, C, ldc, row, M, N, n0, bias, vec_ok, bias_row != 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));   // the leader's barrier
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's shared memory / TMEM stay valid until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<TMEM_COLS>(tmem_base);
  }
}

template <int BN>
static int launch_gemm2(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias,
                        float* C, int64_t ldc, int M, int N, int K, cudaStream_t st, int bias_row = 0) {
  CUtensorMap ta, tb;
  uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}, strA[2] = {2, (uint64_t)lda * 2};
  uint32_t boxA[2] = {BK, BM};
  if (int e = make_tmap_bf16(&ta, A, 2, dimsA, strA, boxA, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}, strB[2] = {2, (uint64_t)ldw * 2};
  uint32_t boxB[2] = {BK, BN / 2};
  if (int e = make_tmap_bf16(&tb, W, 2, dimsB, strB, boxB, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  // per device (and a cheap host-side call): set on every launch, not once per process
  DSB_CUDA(cudaFuncSetAttribute(gemm_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem<BN>::TOTAL));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = cdiv(M, 2 * BM) * cdiv(N, BN);
  int pairs = sms / 2;
  if (tiles < pairs) pairs = tiles;
  gemm_tc2_kernel<BN><<<2 * pairs, GEMM_THREADS, Gemm2Smem<BN>::TOTAL, st>>>(ta, tb, bias, C, ldc, M, N, K, bias_row);
  DSB_CHECK_LAUNCH();
  return 0;
}

template <int BN>
static int launch_gemm(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias,
                       float* C, int64_t ldc, int M, int N, int K, cudaStream_t st, int bias_row = 0) {
  CUtensorMap ta, tb;
  uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}, strA[2] = {2, (uint64_t)lda * 2};
  uint32_t boxA[2] = {BK, BM};
  if (int e = make_tmap_bf16(&ta, A, 2, dimsA, strA, boxA, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}, strB[2] = {2, (uint64_t)ldw * 2};
  uint32_t boxB[2] = {BK, BN};
  if (int e = make_tmap_bf16(&tb, W, 2, dimsB, strB, boxB, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  // per device (and a cheap host-side call): set on every launch, not once per process
  DSB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::TOTAL));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = cdiv(M, BM) * cdiv(N, BN);
  const int grid = tiles < sms ? tiles : sms;
  gemm_tc_kernel<BN><<<grid, GEMM_THREADS, GemmSmem<BN>::TOTAL, st>>>(ta, tb, bias, C, ldc, M, N, K, bias_row);
  DSB_CHECK_LAUNCH();
  return 0;
}

}  // namespace tc

// A [M, lda] bf16 (K valid columns), W [N, ldw] bf16, C [M, ldc] fp32.  lda/ldw multiples of 8 elements.
int gemm_bias_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias,
                 float* C, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  if ((lda & 7) || (ldw & 7)) return set_error(DSB_ERR_INVALID, "gemm_bias_tc: lda/ldw must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return set_error(DSB_ERR_INVALID, "gemm_bias_tc: operands must be 16-byte aligned");
  // pick the widest tile that divides N evenly into few blocks (7200 = 30 x 240; 2400 = 10 x 240; ...)
  static const bool two_cta = !(getenv("DSB_GEMM_2CTA") && atoi(getenv("DSB_GEMM_2CTA")) == 0);
  if (N % 240 == 0 && M >= 256 && two_cta) return tc::launch_gemm2<240>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
  if (N % 240 == 0) return tc::launch_gemm<240>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
  if (N >= 256 && N % 256 == 0) return tc::launch_gemm<256>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
  if (N % 192 == 0) return tc::launch_gemm<192>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
  if (N > 128) return tc::launch_gemm<256>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
  if (N > 64) return tc::launch_gemm<128>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
  return tc::launch_gemm<64>(A, lda, W, ldw, bias, C, ldc, M, N, K, st);
}

// Ct [N, ldc] = (A * W^T + bias)^T: the same product with the operands swapped on the tensor cores (W rows become the
// accumulator rows), so that the output is batch-minor -- what the CTA-pair recurrence (rnn_pair.cu) reads coalesced:
// its epilogue threads own one sequence each, a warp owns 32 consecutive (t, b) rows of one gate column.
int gemm_bias_rows_tc(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, const float* bias,
                      float* Ct, int64_t ldc, int M, int N, int K, cudaStream_t st) {
  if ((lda & 7) || (ldw & 7)) return set_error(DSB_ERR_INVALID, "gemm_bias_rows_tc: lda/ldw must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return set_error(DSB_ERR_INVALID, "gemm_bias_rows_tc: operands must be 16-byte aligned");
  if (N >= 256 && M >= 256) return tc::launch_gemm2<256>(W, ldw, A, lda, bias, Ct, ldc, N, M, K, st, 1);
  if (M > 128) return tc::launch_gemm<256>(W, ldw, A, lda, bias, Ct, ldc, N, M, K, st, 1);
  if (M > 64) return tc::launch_gemm<128>(W, ldw, A, lda, bias, Ct, ldc, N, M, K, st, 1);
  return tc::launch_gemm<64>(W, ldw, A, lda, bias, Ct, ldc, N, M, K, st, 1);
}

}  // namespace dsb

// Diagnostic / general entry point: C = A * W^T + bias on the tensor cores.
extern "C" int dsb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, float* C,
                             int64_t ldc, int M, int N, int K, void* stream) {
  DSB_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0, "dsb_gemm_bf16: bad argument");
  static const bool rows = getenv("DSB_GEMM_ROWS") != nullptr;   // diagnostic: C is then [N, ldc] (the transposed product)
  if (rows)
    return dsb::gemm_bias_rows_tc(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(W),
                                  ldw, bias, C, ldc, M, N, K, (cudaStream_t)stream);
  return dsb::gemm_bias_tc(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(W),
                           ldw, bias, C, ldc, M, N, K, (cudaStream_t)stream);
}
