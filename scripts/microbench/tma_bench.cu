// Diagnostic: how fast does one SM take in the h exchange buffer of the recurrence through TMA?
// Every CTA streams the same [64 rows x HP] bf16 tile (row stride HP*2 bytes, like rnn_tc.cu's hbuf) into a ring of
// shared-memory slots, n_iter times, with `depth` boxes in flight, in one of four ways:
//   mode 0  3-D tensor boxes {64 k, 64 rows, gsz chunks}              (what rnn_tc.cu does)
//   mode 1  1-D bulk copies of gsz*8 KB from a chunk-major, pre-tiled copy of the buffer
//   mode 2  as mode 0 in clusters of two CTAs: each CTA issues every other box and multicasts it to both
//   mode 3  as mode 0 issued by two warps (even / odd boxes)
// Prints cycles per box and bytes per clock per SM.  Not part of the product library.
#include "tc_common.cuh"   // from danspeech_b200/csrc (-I)
#include <vector>

namespace dsb {
namespace tc {

__global__ void __launch_bounds__(128, 1)
tma_bench_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* tiled, int mode, int gsz, int depth,
                 int n_box, int nkc, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], sfree[8];   // sfree (multicast): every CTA of the cluster has taken the slot's previous box
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot_bytes = gsz * 8192;
  const int CL = (int)cluster_nctarank(), crank = (int)cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&sfree[i], (uint32_t)CL);
    fence_mbar_init();
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  const int groups_per_pass = (nkc + gsz - 1) / gsz;
  long long t0 = clock64();
  if (warp == 0 || (mode == 3 && warp == 1)) {
    // box b uses slot b % depth; a slot is reused once its previous box has landed (no consumer: pure intake rate)
    uint32_t phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < n_box; ++b) {
      const int slot = b % depth;
      const bool mine_to_wait = mode != 3 || (b & 1) == warp;
      if (b >= depth && mine_to_wait) {
        while (!mbar_try_wait(&full[slot], phase[slot])) {}
        if (mode == 2) {
          // tell every CTA of the cluster that this CTA's slot is free again, then wait until all of them are
          if (lane < CL) mbar_arrive_cluster(mapa_u32(smem_u32(&sfree[slot]), (uint32_t)lane));
          while (!mbar_try_wait(&sfree[slot], phase[slot])) {}
        }
        phase[slot] ^= 1;
      } else if (b >= depth) {
        phase[slot] ^= 1;   // the other warp waits for it; keep the parity in step
      }
      if (mode == 3 && (b & 1) != warp) continue;
      const int g = b % groups_per_pass;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[slot], (uint32_t)slot_bytes);
        if (mode == 0 || mode == 3) {
          tma_load_3d(smem + slot * slot_bytes, &tmap, &full[slot], 0, 0, g * gsz);
        } else if (mode == 1) {
          bulk_load(smem + slot * slot_bytes, tiled + (size_t)g * gsz * 4096, (uint32_t)slot_bytes, &full[slot]);
        } else {
          if ((b % CL) == crank) tma_load_3d_mcast(smem + slot * slot_bytes, &tmap, &full[slot], 0, 0, g * gsz, (uint16_t)((1u << CL) - 1u));
        }
      }
      __syncwarp();
    }
    // drain
    if (mode != 3 || warp == 0)
      for (int b = max(0, n_box - depth); b < n_box; ++b) {
        const int slot = b % depth;
        while (!mbar_try_wait(&full[slot], phase[slot])) {}
        phase[slot] ^= 1;
      }
  }
  __syncthreads();
  long long t1 = clock64();
  if (CL > 1) cluster_sync_all();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  (void)lane;
}

}  // namespace tc
}  // namespace dsb

// mode, gsz (chunks of 64 k per box), depth (boxes in flight), grid (CTAs), n_box; returns avg / max cycles
extern "C" int dsb_debug_tma_bench(int mode, int gsz, int depth, int grid, int n_box, long long* host_out) {
  using namespace dsb;
  using namespace dsb::tc;
  const int H = 1200, HP = 1216, nkc = 19, rows = 64;
  if (gsz < 1 || gsz * depth * 8192 > 200 * 1024 || depth > 8) return set_error(DSB_ERR_INVALID, "tma_bench: ring too large");
  __nv_bfloat16 *buf = nullptr, *tiled = nullptr;
  long long* d = nullptr;
  DSB_CUDA(cudaMalloc(&buf, (size_t)rows * HP * 2));
  DSB_CUDA(cudaMalloc(&tiled, (size_t)(nkc + 8) * rows * 64 * 2));
  DSB_CUDA(cudaMemset(buf, 0, (size_t)rows * HP * 2));
  DSB_CUDA(cudaMemset(tiled, 0, (size_t)(nkc + 8) * rows * 64 * 2));
  DSB_CUDA(cudaMalloc(&d, sizeof(long long) * grid));
  CUtensorMap th;
  uint64_t dh[3] = {64, (uint64_t)rows, (uint64_t)nkc};
  uint64_t sh[3] = {2, (uint64_t)HP * 2, 128};
  uint32_t bh[3] = {64, (uint32_t)rows, (uint32_t)gsz};
  if (int e = make_tmap_bf16(&th, buf, 3, dh, sh, bh, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  const int smem = gsz * depth * 8192 + 1024;
  DSB_CUDA(cudaFuncSetAttribute(tma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = mode == 2 ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const __nv_bfloat16* tl = tiled;
  int nk = nkc;
  void* args[] = {(void*)&th, (void*)&tl, (void*)&mode, (void*)&gsz, (void*)&depth, (void*)&n_box, (void*)&nk, (void*)&d};
  for (int rep = 0; rep < 2; ++rep) {
    DSB_CUDA(cudaLaunchKernelExC(&cfg, (const void*)tma_bench_kernel, args));
    DSB_CUDA(cudaDeviceSynchronize());
  }
  std::vector<long long> h(grid);
  DSB_CUDA(cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  long long sum = 0, mx = 0;
  for (long long v : h) { sum += v; mx = v > mx ? v : mx; }
  host_out[0] = sum / grid;
  host_out[1] = mx;
  cudaFree(buf); cudaFree(tiled); cudaFree(d);
  (void)H;
  return 0;
}
