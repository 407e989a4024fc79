// bf16 tensor-core MaskConv blocks for sm_100a: implicit-GEMM Conv2d + folded BatchNorm2d +
// Hardtanh(0,20) + length mask (danspeech/deepspeech/model.py:357-392 and MaskConv.forward :65-81).
//
// Activations are channels-last bf16 [B][D][T'][C].  An output tile is 128 consecutive frames of one
// (utterance, output frequency row) x all output channels:  D[128, Cout] = sum over (kh, kw) of
// A_{kh,kw}[128, Cin] * W_{kh,kw}[Cout, Cin]^T.  There is no im2col buffer and no boundary code: the zero padding
// in time and frequency is the TMA out-of-bounds fill.
//   * Blocks 2/3: per kh ONE box of 128 + 10 frames (coordinate t0 - 5) and one box with the 11 weight taps;
//     the A operand of tap kw is the same shared-memory block read from row kw on (descriptor start address
//     + kw rows), so the 11 overlapping windows cost one load instead of eleven.
//   * Block 1 has Cin = 1, which no MMA can use directly; its input is expanded once in time
//     ("x1[b,d,t,j] = spect[b,d,2t+j-5]", 11 taps padded to 16) so that it becomes a kH x 1 convolution over 16
//     channels.  A TMA box with 32-byte rows moves one sector per row and was the bottleneck, so the expansion
//     kernel writes x1 as ready-made operand tiles [b][128-frame block][d][128 x 32 B, 32B-swizzled]: four
//     consecutive kh taps (= input rows) are ONE contiguous 16 KB bulk copy (cp.async.bulk, no tensor map).
// A TMA instruction costs the producer ~240 cycles, more than the MMAs of one (kh,kw) step take: with one step
// per instruction the kernel was producer-bound; now it is bound by the MMA rate (~50-65 cycles per K=16 slice at
// these tile shapes, scripts/mma_microbench.py).
// Warp roles as in gemm_tc.cu: TMA producer / single-thread tcgen05.mma issuer / 4 epilogue warps, with the
// accumulator double-buffered in TMEM so that the epilogue of a tile overlaps the next tile's MMAs.
// PACK = 2 (the 32-channel blocks 1 and 2): a tile covers TWO neighbouring output frequency rows d0 = 2*dp, d0 + 1.
// With N = 32 a tcgen05.mma costs ~53 cycles whatever little it computes (the small-MMA floor), so the 32-channel
// blocks were bound by their instruction count.  Input row f = sd*d0 - pd + r feeds output row d0 through tap kh = r
// and output row d0 + 1 through tap kh = r - sd: with the weights packed as [r][kw][W[kh=r] ; W[kh=r-sd]] (zero blocks
// where a tap does not exist) ONE MMA of N = 64 (~59 cycles) does both, over KH + sd input rows instead of 2 * KH.
// Algorithmic flops = 2*T'*Cout*Dout*Cin*kH*kW per utterance.
#include "tc_common.cuh"
#include "model_types.cuh"

namespace dsb {
namespace tc {

constexpr int CV_BM = 128;
constexpr int CV_THREADS = 192;
constexpr int CV_G1 = 4;        // block 1: kh taps per box / per elected issue region

struct ConvTcParams {
  const float* bias;          // [NOUT] folded
  const int32_t* lens;        // [B] output frames per utterance (nullptr: no mask)
  const __nv_bfloat16* x_tiles;   // block 1 only: operand tiles [B][t_blocks][Din][128 frames x 16]
  __nv_bfloat16* out;
  int B, Tp, Din, Dout, KH, KW, sd, pd, pt;
  int rnn_layout;             // 0: [B][Dout][Tp][NOUT]   1: [(t*B+b)][Dout*NOUT] (feature = d*NOUT + co)
  int64_t out_ld;             // row stride of the rnn layout
  // Segmented time axis (streaming, B == 1): the Tp frames are nb segments of seg_len frames, one per stream,
  // each [seg_off zeros | seg_valid frames | zeros]; only the seg_valid frames of a segment are outputs and
  // frame u of segment s is written as (b = s, t = u - seg_off).  seg_len == 0: plain layout.
  int seg_len, seg_off, seg_valid, nb;
};

// One ring slot = one "group": block 1: CV_G1 kh taps (A tiles + weight tiles); blocks 2/3: one kh = a block of
// 128 + KW - 1 frames and the KW weight taps.
template <int NOUT, int CIN, int PACK>
struct ConvSmem {
  static constexpr bool FIRST = CIN == 16;
  static constexpr int ROW = CIN * 2;
  static constexpr int A_TILE = CV_BM * ROW;                                     // one 128-frame operand tile
  static constexpr int A_ROWS = FIRST ? CV_BM : CV_BM + kConvKW - 1;             // frames per box
  static constexpr int A_SLOT = FIRST ? CV_G1 * A_TILE : (A_ROWS * ROW + 1023) / 1024 * 1024;
  static constexpr int A_TX = FIRST ? CV_G1 * A_TILE : A_ROWS * ROW;             // bytes one A box delivers
  static constexpr int NP = NOUT * PACK;                                          // accumulator columns of a tile
  static constexpr int B_TILE = NP * ROW;
  static constexpr int B_SLOT = (FIRST ? CV_G1 : kConvKW) * B_TILE;
  static_assert(A_TILE % 1024 == 0 && B_TILE % 1024 == 0, "tiles must keep the swizzle phase");
  static constexpr int GROUPS = FIRST ? 4 : (NP <= 64 ? 4 : 2);                  // ring depth
  static constexpr int BAR_OFF = GROUPS * (A_SLOT + B_SLOT);
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <int NOUT, int CIN, int PACK>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const ConvTcParams p) {
  using S = ConvSmem<NOUT, CIN, PACK>;
  constexpr bool FIRST = S::FIRST;
  constexpr int NP = S::NP;
  constexpr int TMEM_COLS = NP <= 32 ? 64 : 256;
  constexpr int ACC_STRIDE = NP <= 32 ? 32 : 128;
  constexpr uint32_t SWZ = CIN == 32 ? 4u : 6u;          // SWIZZLE_64B : SWIZZLE_32B
  constexpr uint32_t SBO = 8 * S::ROW;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + S::GROUPS * S::A_SLOT;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);   // one per ring slot
  uint64_t* empty = full + S::GROUPS;
  uint64_t* tfull = empty + S::GROUPS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_blocks = (p.Tp + CV_BM - 1) / CV_BM;
  const int DoutP = (p.Dout + PACK - 1) / PACK;   // tiles along the output frequency axis
  const int KHP = p.KH + (PACK - 1) * p.sd;       // input rows per tile
  const int n_tiles = p.B * DoutP * t_blocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    for (int i = 0; i < S::GROUPS; ++i) {
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

  // tile -> (b, d = first output row of the tile, t0); frames fastest so that neighbouring CTAs share input rows in L2
  auto decode = [&](int tile, int& b, int& d, int& t0) {
    const int tb = tile % t_blocks;
    const int r = tile / t_blocks;
    d = (r % DoutP) * PACK;
    b = r / DoutP;
    t0 = tb * CV_BM;
  };
  // input rows r (of KHP) that fall inside the input for the tile starting at output row d -- rows that would only
  // feed an output row past Dout are left out too; number of ring slots ("groups") the tile uses
  auto tile_shape = [&](int d, int& kh_lo, int& n_kh, int& n_grp) {
    kh_lo = max(0, p.pd - p.sd * d);
    const int r_max = (d + PACK - 1 < p.Dout) ? KHP - 1 : p.KH - 1 + (p.Dout - 1 - d) * p.sd;
    n_kh = min(r_max, p.Din - 1 + p.pd - p.sd * d) - kh_lo + 1;
    n_grp = FIRST ? (n_kh + CV_G1 - 1) / CV_G1 : n_kh;
  };

  if (warp == 0) {
    // whole warp, warp-uniform control flow; one elected lane issues the two boxes of a group (see elect_one_sync)
    int grp = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int b, d, t0, kh_lo, n_kh, n_grp;
      decode(tile, b, d, t0);
      tile_shape(d, kh_lo, n_kh, n_grp);
      for (int g = 0; g < n_grp; ++g) {
        mbar_wait(&empty[grp], phase ^ 1);
        if (elect_one_sync()) {
          if (FIRST) {
            const int kh = kh_lo + g * CV_G1;
            const int n = min(CV_G1, n_kh - g * CV_G1);            // rows kh .. kh+n-1 are inside the input
            const int row = p.sd * d + kh - p.pd;
            mbar_arrive_expect_tx(&full[grp], (uint32_t)(n * S::A_TILE + S::B_SLOT));
            bulk_load(sA + grp * S::A_SLOT,
                      p.x_tiles + (((size_t)b * t_blocks + t0 / CV_BM) * p.Din + row) * (size_t)(CV_BM * CIN),
                      (uint32_t)(n * S::A_TILE), &full[grp]);
            tma_load_3d(sB + grp * S::B_SLOT, &tmap_w, &full[grp], 0, 0, kh);   // taps past KH are zero fill
          } else {
            mbar_arrive_expect_tx(&full[grp], (uint32_t)(S::A_TX + S::B_SLOT));   // full boxes (out-of-range frames / rows are zero fill)
            const int kh = kh_lo + g;
            tma_load_4d(sA + grp * S::A_SLOT, &tmap_x, &full[grp], 0, t0 - p.pt, p.sd * d + kh - p.pd, b);
            tma_load_3d(sB + grp * S::B_SLOT, &tmap_w, &full[grp], 0, 0, kh * p.KW);
          }
        }
        __syncwarp();
        if (++grp == S::GROUPS) { grp = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(CV_BM, NP);
    const uint64_t desc0 = make_smem_desc(0, 16, SBO, SWZ);
    const uint32_t a_lo = smem_u32(sA) >> 4, b_lo = smem_u32(sB) >> 4;
    int grp = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int b, d, t0, kh_lo, n_kh, n_grp;
      decode(tile, b, d, t0);
      tile_shape(d, kh_lo, n_kh, n_grp);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
      for (int g = 0; g < n_grp; ++g) {
        mbar_wait(&full[grp], phase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a0 = a_lo + (uint32_t)grp * (S::A_SLOT >> 4);
          const uint32_t b0 = b_lo + (uint32_t)grp * (S::B_SLOT >> 4);
          if (FIRST) {
            const int n = min(CV_G1, n_kh - g * CV_G1);
#pragma unroll
            for (int j = 0; j < CV_G1; ++j)
              if (j < n)
                umma_bf16(d_tmem, desc0 + (uint64_t)(a0 + (uint32_t)j * (S::A_TILE >> 4)),
                          desc0 + (uint64_t)(b0 + (uint32_t)j * (S::B_TILE >> 4)), idesc, j ? 1u : (uint32_t)(g != 0));
          } else {
            // tap kw reads the block from row kw on: start address + kw * ROW bytes.  The swizzle XOR is a function
            // of the shared-memory address bits, which the TMA write used as well, so a start that is not aligned
            // to the 8-row pattern needs neither a re-layout nor the descriptor's base-offset field (measured:
            // parity holds with the field left 0 and breaks when it is set to the row phase).
#pragma unroll
            for (int kw = 0; kw < kConvKW; ++kw) {
              const uint32_t as = a0 + (uint32_t)kw * (S::ROW >> 4);
              const uint64_t adesc = desc0 + (uint64_t)as;
              const uint64_t bdesc = desc0 + (uint64_t)(b0 + (uint32_t)kw * (S::B_TILE >> 4));
#pragma unroll
              for (int k = 0; k < CIN / 16; ++k)
                umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                          (kw | k) ? 1u : (uint32_t)(g != 0));
            }
          }
          umma_commit(&empty[grp]);
          if (g == n_grp - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++grp == S::GROUPS) { grp = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int b, d, t0;
      decode(tile, b, d, t0);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      int t = t0 + q * 32 + lane;
      bool valid = t < p.Tp;
      int nb = p.B;
      if (p.seg_len) {
        b = t / p.seg_len;
        t -= b * p.seg_len + p.seg_off;
        valid = valid && t >= 0 && t < p.seg_valid;
        nb = p.nb;
      }
      const bool live = valid && (p.lens == nullptr || t < p.lens[b]);
      const uint32_t t_addr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < NP; c0 += 32) {
        const int dd = d + c0 / NOUT, cc = c0 % NOUT;   // output row and first channel of these 32 accumulator columns
        uint32_t r[32];
        tmem_ld32(t_addr + c0, r);
        tmem_ld_wait();
        if (valid && dd < p.Dout) {
          __nv_bfloat16* dst = p.rnn_layout ? p.out + ((int64_t)t * nb + b) * p.out_ld + (int64_t)dd * NOUT
                                            : p.out + (((int64_t)b * p.Dout + dd) * p.Tp + t) * NOUT;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t w[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              float v0 = 0.f, v1 = 0.f;
              if (live) {
                v0 = fminf(fmaxf(__uint_as_float(r[j + 2 * h]) + __ldg(p.bias + cc + j + 2 * h), 0.f), 20.f);
                v1 = fminf(fmaxf(__uint_as_float(r[j + 2 * h + 1]) + __ldg(p.bias + cc + j + 2 * h + 1), 0.f), 20.f);
              }
              __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
              w[h] = *reinterpret_cast<uint32_t*>(&pk);
            }
            *reinterpret_cast<uint4*>(dst + cc + j) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// Block-1 operand tiles: x1[b][d][t][j] = spect[b][d][2t + j - 5] (j < 11, zero outside [0,T) and for j >= 11),
// stored as [b][t / 128][d][tile], tile = 128 rows (frames) x 32 bytes in the 32-byte-swizzled K-major order the
// MMA reads (16-byte chunk c of row r at r*32 + ((c ^ ((r >> 2) & 1)) * 16); frames >= Tp of the last block are 0.
__global__ void im2col_time_kernel(const float* __restrict__ spect, __nv_bfloat16* __restrict__ x1, int B, int D, int T,
                                   int Tp, int t_blocks) {
  const int Tpad = t_blocks * CV_BM;
  const int64_t total = (int64_t)B * D * Tpad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % Tpad);
    const int64_t bd = i / Tpad;
    const int d = (int)(bd % D);
    const int64_t b = bd / D;
    const float* src = spect + bd * T;
    uint32_t w[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int j0 = 2 * h, j1 = 2 * h + 1;
      const int s0 = 2 * t + j0 - 5, s1 = 2 * t + j1 - 5;
      const float v0 = (t < Tp && j0 < kConvKW && s0 >= 0 && s0 < T) ? src[s0] : 0.f;
      const float v1 = (t < Tp && j1 < kConvKW && s1 >= 0 && s1 < T) ? src[s1] : 0.f;
      __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
      w[h] = *reinterpret_cast<uint32_t*>(&pk);
    }
    const int tb = t / CV_BM, tl = t % CV_BM, sw = (tl >> 2) & 1;
    uint4* dst = reinterpret_cast<uint4*>(x1 + (((b * t_blocks + tb) * D + d) * (int64_t)CV_BM + tl) * 16);
    dst[sw] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[sw ^ 1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// folded fp32 weights [kh][cin][11][cout] -> bf16 [(r*KW + kw)][pack*cout][cin_pad]: block q of the rows holds tap
// kh = r - q*sd (zeros where that tap does not exist), r < KH + (pack-1)*sd; pack = 1: plain [(kh*KW + kw)][cout][cin_pad]
__global__ void pack_conv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int KH, int cin,
                                   int cout, int first, int pack, int sd) {
  const int KW = first ? 1 : kConvKW, CP = first ? 16 : cin;
  const int KHP = KH + (pack - 1) * sd, NP = pack * cout;
  const int64_t total = (int64_t)KHP * KW * NP * CP;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % CP);
    int64_t r = i / CP;
    const int n = (int)(r % NP);
    r /= NP;
    const int kw = (int)(r % KW), rr = (int)(r / KW);
    const int qq = n / cout, co = n % cout, kh = rr - qq * sd;
    float v = 0.f;
    if (kh >= 0 && kh < KH) {
      if (first) v = ci < kConvKW ? w[(((int64_t)kh * cin + 0) * kConvKW + ci) * cout + co] : 0.f;
      else v = w[(((int64_t)kh * cin + ci) * kConvKW + kw) * cout + co];
    }
    out[i] = __float2bfloat16_rn(v);
  }
}

template <int NOUT, int CIN, int PACK>
static int launch_conv(const __nv_bfloat16* x, const ConvLayer& L, const ConvTcParams& p, cudaStream_t st) {
  using S = ConvSmem<NOUT, CIN, PACK>;
  CUtensorMap tx, tw;
  const CUtensorMapSwizzle swz = CIN == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  const uint64_t frame = (uint64_t)CIN * 2;
  if (!S::FIRST) {   // block 1 reads ready-made tiles with bulk copies (p.x_tiles)
    uint64_t dx[4] = {(uint64_t)CIN, (uint64_t)p.Tp, (uint64_t)p.Din, (uint64_t)p.B};
    uint64_t sx[4] = {2, frame, (uint64_t)p.Tp * frame, (uint64_t)p.Din * p.Tp * frame};
    uint32_t bx[4] = {CIN, (uint32_t)S::A_ROWS, 1, 1};
    if (int e = make_tmap_bf16(&tx, x, 4, dx, sx, bx, swz)) return e;
  }
  // weights [(r*KW + kw)][NOUT*PACK][CIN]; rows past the last one are out-of-bounds zero fill
  const int KHP = p.KH + (PACK - 1) * p.sd;
  uint64_t dw[3] = {(uint64_t)CIN, (uint64_t)S::NP, (uint64_t)KHP * p.KW};
  uint64_t sw[3] = {2, frame, (uint64_t)S::NP * frame};
  uint32_t bw[3] = {CIN, (uint32_t)S::NP, S::FIRST ? (uint32_t)CV_G1 : (uint32_t)kConvKW};
  if (int e = make_tmap_bf16(&tw, L.w_tc, 3, dw, sw, bw, swz)) return e;
  // per device (and a cheap host-side call): set on every launch, not once per process
  DSB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NOUT, CIN, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.B * cdiv(p.Dout, PACK) * cdiv(p.Tp, CV_BM);
  conv_tc_kernel<NOUT, CIN, PACK><<<tiles < sms ? tiles : sms, CV_THREADS, S::TOTAL, st>>>(S::FIRST ? tw : tx, tw, p);
  DSB_CHECK_LAUNCH();
  return 0;
}

}  // namespace tc

size_t conv1_tiles_elems(int B, int Tp) { return (size_t)B * cdiv(Tp, tc::CV_BM) * kFreqBins * tc::CV_BM * 16; }

int im2col_time_tc(const float* spect, __nv_bfloat16* x1, int B, int T, int Tp, cudaStream_t st) {
  const int t_blocks = cdiv(Tp, tc::CV_BM);
  const int64_t total = (int64_t)B * kFreqBins * t_blocks * tc::CV_BM;
  tc::im2col_time_kernel<<<(int)(cdiv64(total, 256) < 148 * 16 ? cdiv64(total, 256) : 148 * 16), 256, 0, st>>>(
      spect, x1, B, kFreqBins, T, Tp, t_blocks);
  DSB_CHECK_LAUNCH();
  return 0;
}

// output rows per tile (see the header): block 1 (Cin = 1, bound by re-reading its expanded input rows from L2) four,
// the other 32-channel block two, the 96-channel block one
static int conv_pack(const ConvLayer& L, bool first) { return L.cout == 32 ? (first ? 4 : 2) : 1; }

size_t conv_w_tc_elems(const ConvLayer& L, bool first) {
  const int pack = conv_pack(L, first);
  return (size_t)(L.kh + (pack - 1) * L.sd) * (first ? 1 : kConvKW) * pack * L.cout * (first ? 16 : L.cin);
}

int pack_conv_w_tc(const ConvLayer& L, bool first, __nv_bfloat16* out, cudaStream_t st) {
  const int64_t total = (int64_t)conv_w_tc_elems(L, first);
  tc::pack_conv_w_kernel<<<(int)(cdiv64(total, 256) < 2048 ? cdiv64(total, 256) : 2048), 256, 0, st>>>(
      L.w, out, L.kh, L.cin, L.cout, first ? 1 : 0, conv_pack(L, first), L.sd);
  DSB_CHECK_LAUNCH();
  return 0;
}

// x: block 0 -> operand tiles of the time-expanded spectrogram (im2col_time_tc); later blocks -> [B][Din][Tp][32]
// d_len == nullptr: no length mask.  seg != nullptr: segmented time axis {seg_len, seg_off, seg_valid, n_streams}.
int conv_block_tc(const __nv_bfloat16* x, const ConvLayer& L, bool first, const int32_t* d_len, int B, int Tp,
                  __nv_bfloat16* out, bool rnn_layout, int64_t out_ld, cudaStream_t st, const int* seg) {
  tc::ConvTcParams p{};
  if (seg) {
    if (B != 1 || !rnn_layout) return set_error(DSB_ERR_INVALID, "conv_block_tc: segmented mode needs B == 1 and the rnn layout");
    p.seg_len = seg[0]; p.seg_off = seg[1]; p.seg_valid = seg[2]; p.nb = seg[3];
  }
  p.bias = L.bias;
  p.lens = d_len;
  p.x_tiles = first ? x : nullptr;
  p.out = out;
  p.B = B; p.Tp = Tp; p.Din = L.din; p.Dout = L.dout; p.KH = L.kh;
  p.KW = first ? 1 : kConvKW;
  p.sd = L.sd; p.pd = L.pd;
  p.pt = first ? 0 : kConvPT;
  p.rnn_layout = rnn_layout ? 1 : 0;
  p.out_ld = out_ld;
  if (first && L.cout == 32) return tc::launch_conv<32, 16, 4>(x, L, p, st);
  if (!first && L.cin == 32 && L.cout == 32) return tc::launch_conv<32, 32, 2>(x, L, p, st);
  if (!first && L.cin == 32 && L.cout == 96) return tc::launch_conv<96, 32, 1>(x, L, p, st);
  return set_error(DSB_ERR_UNSUPPORTED, "conv_block_tc: unsupported block (cin=%d, cout=%d)", L.cin, L.cout);
}

}  // namespace dsb
