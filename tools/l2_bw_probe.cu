// Probe: aggregate L2 -> SM bandwidth when every CTA streams the SAME 512 KB of weights through TMA
// (what a fused multi-layer kernel does with the layer weights).  nvcc -arch=sm_100a.
#include <stdio.h>
#include <stdlib.h>
#include "../implicit_image_compression_b200/csrc/ptx.cuh"
#include "../implicit_image_compression_b200/csrc/tmap.h"
using namespace sb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

template <int STAGES, int ROWS>
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, int iters, int nblk) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * ROWS * 128);
  if (threadIdx.x == 0) { for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    // keep STAGES loads in flight; wait for the oldest, reissue
    for (int i = 0; i < iters + STAGES; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) mbar_wait(&full[s], ((i / STAGES) - 1) & 1);
      if (i < iters) {
        mbar_expect_tx(&full[s], ROWS * 128);
        const int blk = (i + blockIdx.x) % nblk;   // 64-col x ROWS-row box index
        tma_load_2d(smem + s * ROWS * 128, &tm, &full[s], (blk % 4) * 64, (blk / 4) * ROWS);
      }
    }
  }
}

int main() {
  const int W = 256, L = 4;                  // 4 layers of 256x256 fp16 = 512 KB
  void* d; CK(cudaMalloc(&d, size_t(L) * W * W * 2)); CK(cudaMemset(d, 0, size_t(L) * W * W * 2));
  CUtensorMap tm;
  if (make_tmap_16bit(&tm, d, uint64_t(L) * W, W, 256, false)) { printf("tmap fail\n"); return 1; }
  int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  constexpr int ST = 4, ROWS = 256;
  const int smem_bytes = ST * ROWS * 128 + 64 + 1024;
  auto k = probe<ST, ROWS>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int iters = 2000, nblk = L * 4;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int grid : {nsm, nsm / 2, 32}) {
    k<<<grid, 128, smem_bytes>>>(tm, 200, nblk); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); k<<<grid, 128, smem_bytes>>>(tm, iters, nblk); CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = double(grid) * iters * ROWS * 128;
    printf("grid %3d: %.3f ms  L2->SM %.2f TB/s  (%.1f B/clk/SM at 1.965 GHz)\n", grid, ms, bytes / ms * 1e-9,
           bytes / ms * 1e-6 / grid / 1.965e3 * 1e0 / 1e0);
  }
  return 0;
}
