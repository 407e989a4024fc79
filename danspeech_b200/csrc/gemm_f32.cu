// fp32 (CUDA-core) GEMM: C[M,N] = A[M,K] * W[N,K]^T + bias[N].
//
// Used by DSB_PREC_FP32 for the BatchRNN input projections (x_t * W_ih^T + b_ih for all t at once;
// the torch.nn.GRU/LSTM/RNN input half behind model.py:107-108,118) and for the SequenceWise FC
// (model.py:414-420).  The bf16 tensor-core path is gemm_tc.cu.
// 128x128x8 tiles, 256 threads, 8x8 register micro-tiles, register-prefetched double buffering.
#include "model_types.cuh"

namespace dsb {

constexpr int GM = 128, GN = 128, GK = 8;

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                float* __restrict__ C, int64_t M, int N, int K) {
  __shared__ __align__(16) float As[2][GK][GM + 4];
  __shared__ __align__(16) float Bs[2][GK][GN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * GM;
  const int n0 = blockIdx.x * GN;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;
  const bool vec = (K & 3) == 0;

  auto load = [&](const float* __restrict__ P, int64_t row, int64_t rows, int k0, float (&r)[4]) {
    r[0] = r[1] = r[2] = r[3] = 0.0f;
    if (row < rows) {
      const float* p = P + row * K + k0 + lk;
      if (vec && k0 + lk + 3 < K) {
        float4 v = __ldg(reinterpret_cast<const float4*>(p));
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (k0 + lk + i < K) r[i] = __ldg(p + i);
      }
    }
  };

  float ra[4], rb[4];
  load(A, m0 + lrow, M, 0, ra);
  load(W, n0 + lrow, N, 0, rb);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    As[0][lk + i][lrow] = ra[i];
    Bs[0][lk + i][lrow] = rb[i];
  }
  __syncthreads();

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  const int ty = tid >> 4, tx = tid & 15;
  const int nk = (K + GK - 1) / GK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load(A, m0 + lrow, M, (kt + 1) * GK, ra);
      load(W, n0 + lrow, N, (kt + 1) * GK, rb);
    }
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nxt = cur ^ 1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        As[nxt][lk + i][lrow] = ra[i];
        Bs[nxt][lk + i][lrow] = rb[i];
      }
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < N) C[m * N + n] = acc[i][j] + (bias ? bias[n] : 0.0f);
    }
  }
}

int gemm_bias_f32(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K,
                  cudaStream_t st) {
  dim3 grid(cdiv(N, GN), (unsigned)cdiv64(M, GM));
  gemm_f32_kernel<<<grid, 256, 0, st>>>(A, W, bias, C, M, N, K);
  DSB_CHECK_LAUNCH();
  return 0;
}

}  // namespace dsb
