// Standalone self-test + micro-benchmark for the tcgen05 kernels (runs on a B200 only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_selftest tools/tc_selftest.cu
// Checks rowgemm (FWD / DX) and colgemm against a double-precision CPU evaluation of the
// same fp16/bf16-rounded inputs, then times them at the c2 problem size (393216 x 256).
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../implicit_image_compression_b200/csrc/tc_kernels.cuh"
#include "../implicit_image_compression_b200/csrc/tmap.h"

using namespace sb;

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                    \
    }                                                                             \
  } while (0)

static uint32_t rng_state = 12345;
static float frand() {  // U(-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) * (1.0f / 8388608.0f)) - 1.0f;
}
static uint16_t f2h(float f) { __half h = __float2half_rn(f); return *reinterpret_cast<uint16_t*>(&h); }
static float h2f(uint16_t u) { __half h = *reinterpret_cast<__half*>(&u); return __half2float(h); }
static uint16_t f2b(float f) { __nv_bfloat16 h = __float2bfloat16_rn(f); return *reinterpret_cast<uint16_t*>(&h); }
static float b2f(uint16_t u) { __nv_bfloat16 h = *reinterpret_cast<__nv_bfloat16*>(&u); return __bfloat162float(h); }

template <int W>
static int test_rowgemm(int rows, bool timing) {
  const int K = W, N = W;
  const int tiles = (rows + 127) / 128;
  const int rows_pad = tiles * 128;
  const float omega = 30.0f;
  std::vector<uint16_t> hA(size_t(rows_pad) * K), hB(size_t(N) * K), hE(size_t(rows_pad) * N);
  std::vector<float> hbias(N);
  for (auto& v : hA) v = f2h(frand());
  const float wb = sqrtf(6.0f / K) / omega;
  for (auto& v : hB) v = f2h(frand() * wb);
  for (auto& v : hbias) v = frand() / sqrtf(float(K));
  for (auto& v : hE) v = f2h(frand());

  uint16_t *dA, *dB, *dE, *dO;
  float* dbias;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dE, hE.size() * 2));
  CK(cudaMalloc(&dO, size_t(rows_pad) * N * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dE, hE.data(), hE.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));

  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  int fails = 0;

  for (int mode = 0; mode < 2; ++mode) {
    // mode 0: FWD (f16 x f16 -> signed-half), mode 1: DX f16 in -> f16 out,
    // mode 2: DX with bf16 A operand x f16 B -> bf16 out (mixed operand formats)
    const bool a_bf16 = (mode == 2);
    if (a_bf16) {
      for (auto& v : hA) v = f2b(frand());
      CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    }
    CUtensorMap tmA, tmB, tmE, tmO;
    if (make_tmap_16bit(&tmA, dA, rows_pad, K, 128, a_bf16) ||
        make_tmap_16bit(&tmB, dB, N, K, N, false) ||
        make_tmap_16bit(&tmE, dE, rows_pad, N, 128, false) ||
        make_tmap_16bit(&tmO, dO, rows_pad, N, 128, mode == 2)) {
      printf("tensor map encode failed\n");
      return 1;
    }
    RowGemmArgs args{};
    args.num_tiles = tiles;
    args.a_row0 = args.e_row0 = args.o_row0 = 0;
    args.valid_rows = rows;
    args.omega = omega;
    args.bias = dbias;
    const uint32_t idesc = umma_idesc(128, N, a_bf16 ? 1 : 0, 0, 0, 0);
    const int grid = tiles < nsm ? tiles : nsm;
    CK(cudaMemset(dO, 0xFF, size_t(rows_pad) * N * 2));
    auto launch = [&]() {
      if (mode == 0) {
        using C = RowGemmCfg<W, W, MODE_FWD>;
        auto kfn = rowgemm_kernel<W, W, MODE_FWD, false>;
        CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        kfn<<<grid, rowgemm_threads(MODE_FWD, false, false), C::SMEM_BYTES>>>(tmA, tmB, tmE, tmO, args, idesc);
      } else if (mode == 1) {
        using C = RowGemmCfg<W, W, MODE_DX>;
        auto kfn = rowgemm_kernel<W, W, MODE_DX, false>;
        CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        kfn<<<grid, rowgemm_threads(MODE_DX, false, false), C::SMEM_BYTES>>>(tmA, tmB, tmE, tmO, args, idesc);
      } else {
        using C = RowGemmCfg<W, W, MODE_DX>;
        auto kfn = rowgemm_kernel<W, W, MODE_DX, true>;
        CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        kfn<<<grid, rowgemm_threads(MODE_DX, false, false), C::SMEM_BYTES>>>(tmA, tmB, tmE, tmO, args, idesc);
      }
    };
    launch();
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());

    if (timing) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      for (int i = 0; i < 3; ++i) launch();
      CK(cudaEventRecord(e0));
      const int reps = 20;
      for (int i = 0; i < reps; ++i) launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= reps;
      const double flops = 2.0 * rows * double(K) * N;
      const double bytes = double(rows) * N * 2 * (mode == 0 ? 2 : 3);
      printf("  rowgemm W=%d mode=%d rows=%d: %.1f us  %.1f TFLOP/s  %.0f GB/s (algorithmic)\n", W,
             mode, rows, ms * 1e3, flops / ms * 1e-9, bytes / ms * 1e-6);
    }

    std::vector<uint16_t> hO(size_t(rows_pad) * N);
    CK(cudaMemcpy(hO.data(), dO, hO.size() * 2, cudaMemcpyDeviceToHost));
    // verify a subset of rows (all rows of first/last tiles + strided sample)
    double max_err = 0;
    long sign_bad = 0, checked = 0;
    for (int r = 0; r < rows; ++r) {
      const bool pick = r < 256 || r >= rows - 256 || (r % 97) == 0;
      if (!pick) continue;
      for (int n = 0; n < N; ++n) {
        double acc = 0;
        for (int k = 0; k < K; ++k) {
          const double a = a_bf16 ? b2f(hA[size_t(r) * K + k]) : h2f(hA[size_t(r) * K + k]);
          acc += a * h2f(hB[size_t(n) * K + k]);
        }
        const uint16_t got = hO[size_t(r) * N + n];
        double ref, gv;
        if (mode == 0) {
          const double t = omega * (acc + hbias[n]);
          ref = sin(t);
          gv = h2f(got);
          const bool cneg = cos(t) < 0;
          if (fabs(cos(t)) > 1e-3 && cneg != bool(got & 1)) ++sign_bad;
        } else {
          const uint16_t e = hE[size_t(r) * N + n];
          const double a = h2f(e);
          double c = sqrt(fmax(0.0, 1.0 - a * a));
          if (e & 1) c = -c;
          ref = acc * c;
          gv = (mode == 2) ? b2f(got) : h2f(got);
        }
        const double err = fabs(gv - ref);
        const double tol_scale = (mode == 2) ? 1.0 / 128 : 1.0 / 512;
        const double nerr = err / (fabs(ref) * tol_scale + 2e-3);
        if (nerr > max_err) max_err = nerr;
        ++checked;
      }
    }
    // rows beyond `rows` inside the last tile must not have been written past the tensor
    const bool ok = max_err < 1.0 && sign_bad == 0;
    printf("rowgemm W=%d mode=%d rows=%d: checked %ld  max_norm_err %.3f  sign_bad %ld  %s\n", W,
           mode, rows, checked, max_err, sign_bad, ok ? "PASS" : "FAIL");
    fails += !ok;
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dE); cudaFree(dO); cudaFree(dbias);
  return fails;
}

template <int W>
static int test_colgemm(int rows, bool timing) {
  const int NX = W, NY = W;
  const int tiles = (rows + 127) / 128;
  const int rows_pad = tiles * 128;
  const int nprob = 2;  // two stacked problems to exercise row offsets
  int fails = 0;
  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  for (int xfmt = 0; xfmt < 1; ++xfmt) {  // X fp16 (mixed f16/bf16 operands are an illegal instruction on sm_100a)
    std::vector<uint16_t> hX(size_t(nprob) * rows_pad * NX, 0), hY(size_t(nprob) * rows_pad * NY, 0);
    for (int p = 0; p < nprob; ++p)
      for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < NX; ++c) {
          const float v = frand() * 0.25f;
          hX[(size_t(p) * rows_pad + r) * NX + c] = xfmt ? f2b(v) : f2h(v);
        }
        for (int c = 0; c < NY; ++c) hY[(size_t(p) * rows_pad + r) * NY + c] = f2h(frand());
      }
    uint16_t *dX, *dY;
    CK(cudaMalloc(&dX, hX.size() * 2));
    CK(cudaMalloc(&dY, hY.size() * 2));
    CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dY, hY.data(), hY.size() * 2, cudaMemcpyHostToDevice));
    ColGemmJobs jobs{};
    jobs.num_problems = nprob;
    jobs.mblocks = NX / 128;
    jobs.nparts = 1;
    jobs.ny_total = NY;
    jobs.prob0 = 0;
    jobs.prob_total = nprob;
    jobs.interleave = 0;
    jobs.splits = nsm / (nprob * jobs.mblocks);
    if (jobs.splits > tiles) jobs.splits = tiles;
    jobs.tiles_total = tiles;
    jobs.tiles_per_split = (tiles + jobs.splits - 1) / jobs.splits;
    for (int p = 0; p < nprob; ++p) jobs.x_row0[p] = jobs.y_row0[p] = p * rows_pad;
    jobs.nx = NX;
    CK(cudaMalloc(&jobs.dw_partial, size_t(jobs.splits) * nprob * NX * NY * 4));
    CK(cudaMalloc(&jobs.db_partial, size_t(jobs.splits) * nprob * NX * 4));
    CUtensorMap tmX, tmY;
    if (xfmt == 1) {
      printf("colgemm: bf16 X operands are no longer built (fp16 only)\n");
      return 1;
    }
    if (make_tmap_16bit_chunks(&tmX, dX, uint64_t(nprob) * rows_pad, NX, 128, 2) ||
        make_tmap_16bit_chunks(&tmY, dY, uint64_t(nprob) * rows_pad, NY, 128, NY / 64)) {
      printf("tensor map encode failed\n");
      return 1;
    }
    const uint32_t idm = umma_idesc(128, NY, xfmt, 0, 1, 1);
    const uint32_t ido = umma_idesc(128, 16, xfmt, 0, 1, 1);
    using C = ColGemmCfg<W>;
    auto kfn = colgemm_kernel<W>;
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    const int grid = nprob * jobs.mblocks * jobs.splits;
    kfn<<<grid, 256, C::SMEM_BYTES>>>(tmX, tmY, jobs, idm, ido);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (timing) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      for (int i = 0; i < 3; ++i) kfn<<<grid, 256, C::SMEM_BYTES>>>(tmX, tmY, jobs, idm, ido);
      CK(cudaEventRecord(e0));
      const int reps = 20;
      for (int i = 0; i < reps; ++i) kfn<<<grid, 256, C::SMEM_BYTES>>>(tmX, tmY, jobs, idm, ido);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= reps;
      const double flops = 2.0 * nprob * rows * double(NX) * NY;
      const double bytes = double(nprob) * rows * (NX + NY) * 2;
      printf("  colgemm W=%d xfmt=%d rows=%d stages=%d grid=%d: %.1f us  %.1f TFLOP/s  %.0f GB/s\n",
             W, xfmt, rows, C::STAGES, grid, ms * 1e3, flops / ms * 1e-9, bytes / ms * 1e-6);
    }
    std::vector<float> hdw(size_t(jobs.splits) * nprob * NX * NY), hdb(size_t(jobs.splits) * nprob * NX);
    CK(cudaMemcpy(hdw.data(), jobs.dw_partial, hdw.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdb.data(), jobs.db_partial, hdb.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0;
    const int check_rows = rows > 4096 ? 4096 : rows;  // CPU cost bound: compare on a prefix...
    (void)check_rows;
    for (int p = 0; p < nprob; ++p) {
      // full reference only for a sample of (m, n) pairs
      for (int sidx = 0; sidx < 400; ++sidx) {
        const int m = (sidx * 37 + p * 11) % NX, n = (sidx * 101 + 7) % NY;
        double ref = 0, refb = 0;
        for (int r = 0; r < rows; ++r) {
          const uint16_t xr = hX[(size_t(p) * rows_pad + r) * NX + m];
          const double x = xfmt ? b2f(xr) : h2f(xr);
          ref += x * h2f(hY[(size_t(p) * rows_pad + r) * NY + n]);
          refb += x;
        }
        double got = 0, gotb = 0;
        for (int s = 0; s < jobs.splits; ++s) {
          got += hdw[((size_t(s) * nprob + p) * NX + m) * NY + n];
          gotb += hdb[(size_t(s) * nprob + p) * NX + m];
        }
        const double scale = sqrt(double(rows)) * 0.25 * 0.6;
        const double e1 = fabs(got - ref) / (scale * 1e-3 + 1e-4);
        const double e2 = fabs(gotb - refb) / (scale * 1e-3 + 1e-4);
        if (e1 > max_err) max_err = e1;
        if (e2 > max_err) max_err = e2;
      }
    }
    const bool ok = max_err < 1.0;
    printf("colgemm W=%d xfmt=%d rows=%d splits=%d: max_norm_err %.4f %s\n", W, xfmt, rows,
           jobs.splits, max_err, ok ? "PASS" : "FAIL");
    fails += !ok;
    cudaFree(dX); cudaFree(dY); cudaFree(jobs.dw_partial); cudaFree(jobs.db_partial);
  }
  return fails;
}

int main(int argc, char** argv) {
  const bool big = argc > 1 && atoi(argv[1]) != 0;
  int fails = 0;
  fails += test_rowgemm<256>(128 * 5 + 77, false);
  fails += test_rowgemm<128>(128 * 3 + 5, false);
  fails += test_colgemm<256>(128 * 41 + 19, false);
  fails += test_colgemm<128>(128 * 9 + 3, false);
  if (big) {
    fails += test_rowgemm<256>(393216, true);
    fails += test_colgemm<256>(393216, true);
  }
  printf("%s (%d failures)\n", fails ? "SELFTEST FAILED" : "SELFTEST OK", fails);
  return fails ? 1 : 0;
}
