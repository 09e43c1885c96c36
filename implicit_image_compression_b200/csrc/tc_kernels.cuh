// tcgen05 / TMEM / TMA kernels for the SIREN hidden layers (sm_100a only).
//
//   rowgemm_kernel  : C[p, n] = f( sum_k A[p, k] * B[n, k] )   p = pixel rows (M = 128 per tile)
//       MODE_FWD    : f = sin(omega * (acc + bias[n]))  -> "signed-half" activation
//                     (reference: SineLayer.forward, implicit_image/models/siren.py:56-68)
//       MODE_DX     : f = acc * sqrt(1 - a^2) * sgn      -> dZ of the previous layer, where
//                     a = stashed activation of that layer (autograd of siren.py:66:
//                     d sin(w z) = w cos(w z); w is pre-folded into B)
//   colgemm_kernel  : dW[m, n] = sum_p X[p, m] * Y[p, n],  db[m] = sum_p X[p, m]
//                     (the pixel-dimension reduction of autograd's Linear backward)
//
// "signed-half": an IEEE fp16 value of sin(t) whose mantissa LSB is replaced by the sign
// of cos(t) (1 = negative).  The backward pass rebuilds cos(t) = +-sqrt(1 - a^2) from it,
// so only ONE 2-byte tensor per layer is stashed between forward and backward.
//
// Shared-memory operand tiles use the canonical UMMA 128-byte-swizzle layouts, which are
// exactly what a TMA box of {64 fp16, rows} with CU_TENSOR_MAP_SWIZZLE_128B produces:
//   K-major  : row r of the tile = 128 contiguous bytes (64 K-elements), 8-row atoms of 1 KiB
//   MN-major : row k of the tile = 128 contiguous bytes (64 M/N-elements), 8-row atoms of 1 KiB
#pragma once
#include "ptx.cuh"

namespace sb {

constexpr int kRowsPerTile = 128;
constexpr uint32_t kChunkBytes = 128 * 128;  // one {64 x 128-row} fp16 box = 16 KiB
constexpr float kInvPi = 0.318309886183790671538f;
constexpr float kRoundMagic = 12582912.0f;  // 1.5 * 2^23: fp32 add rounds to integer, even LSB

enum RowGemmMode { MODE_FWD = 0, MODE_DX = 1 };

__host__ __device__ constexpr uint32_t tmem_cols_pow2(uint32_t n) {
  return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512;
}

// pack two fp32 -> f16x2 (lo = a, hi = b)
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// sin(t) as signed-half pair.  t0/t1 are the sine arguments.
__device__ __forceinline__ uint32_t sine_signed_half2(float t0, float t1) {
  const float s0 = __sinf(t0), s1 = __sinf(t1);
  // LSB of (t/pi + magic) = parity of rint(t/pi) = [cos(t) < 0]
  const uint32_t m0 = __float_as_uint(fmaf(t0, kInvPi, kRoundMagic));
  const uint32_t m1 = __float_as_uint(fmaf(t1, kInvPi, kRoundMagic));
  const uint32_t h = pack_f16x2(s0, s1);
  const uint32_t p = __byte_perm(m0, m1, 0x4400);  // byte0 <- m0.b0, byte2 <- m1.b0
  return (h & 0xFFFEFFFEu) | (p & 0x00010001u);
}

// cos factor from a signed-half: +-sqrt(1 - a^2)
__device__ __forceinline__ float cos_from_signed_half(uint32_t h16) {
  const float a = __half2float(__ushort_as_half(static_cast<unsigned short>(h16 & 0xFFFFu)));
  const float c = fast_sqrt(__saturatef(fmaf(-a, a, 1.0f)));
  return __uint_as_float(__float_as_uint(c) ^ ((h16 & 1u) << 31));
}

// ------------------------------------------------------------------------------------------
// rowgemm
// ------------------------------------------------------------------------------------------
template <int KDIM, int NDIM, int MODE>
struct RowGemmCfg {
  static_assert(KDIM % 64 == 0 && NDIM % 64 == 0, "hidden size must be a multiple of 64");
  static_assert(NDIM >= 16 && NDIM <= 256, "UMMA N range");
  static constexpr int KB = KDIM / 64;  // 64-wide K blocks
  static constexpr int NB = NDIM / 64;  // 64-wide output chunks
  static constexpr int SA = 3;          // A-tile ring depth (16 KiB stages)
  static constexpr int SEO = (MODE == MODE_DX) ? 3 : 2;  // epilogue in/out ring depth
  static constexpr uint32_t B_KB_BYTES = NDIM * 128;
  static constexpr uint32_t OFF_B = 0;
  static constexpr uint32_t OFF_A = OFF_B + KB * B_KB_BYTES;
  static constexpr uint32_t OFF_EO = OFF_A + SA * kChunkBytes;
  static constexpr uint32_t OFF_CONST = OFF_EO + SEO * kChunkBytes;
  static constexpr uint32_t CONST_BYTES = (MODE == MODE_FWD) ? NDIM * 4 : 0;
  static constexpr uint32_t OFF_BAR = OFF_CONST + CONST_BYTES;
  static constexpr int NUM_BARS = 2 * SA + 1 + 2 * SEO + 4;
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;  // + align slack
  static constexpr uint32_t TMEM_COLS = tmem_cols_pow2(2 * NDIM);
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of shared memory");
};

struct RowGemmArgs {
  int num_tiles;      // 128-row tiles in this launch
  int a_row0;         // first row of this launch inside the A tensor map
  int e_row0;         // ... inside the epilogue-input tensor map (MODE_DX)
  int o_row0;         // ... inside the output tensor map
  int valid_rows;     // rows >= valid_rows (relative to the launch) are written as zero (MODE_DX)
  float omega;        // MODE_FWD: sine frequency
  const float* bias;  // MODE_FWD: fp32 bias[NDIM]
};

template <int KDIM, int NDIM, int MODE, bool OUT_BF16>
__global__ void __launch_bounds__(256, 1)
rowgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmO,
               const RowGemmArgs args, const uint32_t idesc) {
  using C = RowGemmCfg<KDIM, NDIM, MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::SA;
  uint64_t* b_full = a_empty + C::SA;
  uint64_t* eo_full = b_full + 1;
  uint64_t* eo_empty = eo_full + C::SEO;
  uint64_t* tm_full = eo_empty + C::SEO;
  uint64_t* tm_empty = tm_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::SA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(b_full, 1);
    for (int i = 0; i < C::SEO; ++i) {
      mbar_init(&eo_full[i], 1);
      mbar_init(&eo_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tm_full[i], 1);
      mbar_init(&tm_empty[i], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (MODE == MODE_DX) tma_prefetch_desc(&tmE);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  if (MODE == MODE_FWD && warp >= 4) {
    float* cst = reinterpret_cast<float*>(smem + C::OFF_CONST);
    for (int i = threadIdx.x - 128; i < NDIM; i += 128) cst[i] = args.omega * args.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: B once, then A k-blocks =====================
    if (lane == 0) {
      mbar_expect_tx(b_full, C::KB * C::B_KB_BYTES);
      for (int kb = 0; kb < C::KB; ++kb)
        tma_load_2d(smem + C::OFF_B + kb * C::B_KB_BYTES, &tmB, b_full, kb * 64, 0);
      uint32_t ia = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
        const int row = args.a_row0 + t * kRowsPerTile;
        for (int kb = 0; kb < C::KB; ++kb, ++ia) {
          const uint32_t s = ia % C::SA, ph = (ia / C::SA) & 1u;
          mbar_wait(&a_empty[s], ph ^ 1u);
          mbar_expect_tx(&a_full[s], kChunkBytes);
          tma_load_2d(smem + C::OFF_A + s * kChunkBytes, &tmA, &a_full[s], kb * 64, row);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_wait(b_full, 0);
      tc_fence_after();
      uint32_t ia = 0, it = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
        mbar_wait(&tm_empty[acc], aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * NDIM;
        for (int kb = 0; kb < C::KB; ++kb, ++ia) {
          const uint32_t s = ia % C::SA, ph = (ia / C::SA) & 1u;
          mbar_wait(&a_full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + C::OFF_A + s * kChunkBytes);
          const uint32_t b_addr = smem_u32(smem + C::OFF_B + kb * C::B_KB_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * 32, 0, 1024, 2);
            const uint64_t db = umma_smem_desc(b_addr + k * 32, 0, 1024, 2);
            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&tm_full[acc]);
      }
    }
  } else if (warp == 2) {
    // ===================== epilogue-input producer (MODE_DX only) =====================
    if (MODE == MODE_DX && lane == 0) {
      uint32_t ic = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
        const int row = args.e_row0 + t * kRowsPerTile;
        for (int nb = 0; nb < C::NB; ++nb, ++ic) {
          const uint32_t s = ic % C::SEO, ph = (ic / C::SEO) & 1u;
          mbar_wait(&eo_empty[s], ph ^ 1u);
          mbar_expect_tx(&eo_full[s], kChunkBytes);
          tma_load_2d(smem + C::OFF_EO + s * kChunkBytes, &tmE, &eo_full[s], nb * 64, row);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> f() -> smem -> TMA store =====================
    const int q = warp & 3;
    const int r_in_tile = q * 32 + lane;
    const bool issuer = (threadIdx.x == 128);
    const float* cst = reinterpret_cast<const float*>(smem + C::OFF_CONST);
    uint32_t it = 0, ic = 0;
    for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
      const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
      mbar_wait(&tm_full[acc], aph);
      tc_fence_after();
      const bool row_valid = (t * kRowsPerTile + r_in_tile) < args.valid_rows;
      for (int nb = 0; nb < C::NB; ++nb, ++ic) {
        const uint32_t s = ic % C::SEO, ph = (ic / C::SEO) & 1u;
        const uint32_t buf = smem_u32(smem + C::OFF_EO + s * kChunkBytes);
        const uint32_t row_addr = buf + r_in_tile * 128;
        if (MODE == MODE_DX)
          mbar_wait(&eo_full[s], ph);
        else
          mbar_wait(&eo_empty[s], ph ^ 1u);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + acc * NDIM + nb * 64 + hb * 32, v);
          tmem_ld_wait();
          uint32_t o[16];
          if (MODE == MODE_FWD) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = nb * 64 + hb * 32 + 2 * j;
              const float t0 = fmaf(__uint_as_float(v[2 * j]), args.omega, cst[col]);
              const float t1 = fmaf(__uint_as_float(v[2 * j + 1]), args.omega, cst[col + 1]);
              o[j] = sine_signed_half2(t0, t1);
            }
          } else {
            uint32_t e[16];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const uint32_t chunk = uint32_t(hb * 4 + c4) ^ uint32_t(r_in_tile & 7);
              const uint4 ld = ld_shared_v4(row_addr + (chunk << 4));
              e[4 * c4 + 0] = ld.x;
              e[4 * c4 + 1] = ld.y;
              e[4 * c4 + 2] = ld.z;
              e[4 * c4 + 3] = ld.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float g0 = __uint_as_float(v[2 * j]) * cos_from_signed_half(e[j] & 0xFFFFu);
              float g1 = __uint_as_float(v[2 * j + 1]) * cos_from_signed_half(e[j] >> 16);
              if (!row_valid) g0 = g1 = 0.0f;
              o[j] = OUT_BF16 ? pack_bf16x2(g0, g1) : pack_f16x2(g0, g1);
            }
          }
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const uint32_t chunk = uint32_t(hb * 4 + c4) ^ uint32_t(r_in_tile & 7);
            st_shared_v4(row_addr + (chunk << 4), o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2],
                         o[4 * c4 + 3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer) {
          tma_store_2d(&tmO, smem + C::OFF_EO + s * kChunkBytes, nb * 64,
                       args.o_row0 + t * kRowsPerTile);
          tma_store_commit();
          if (ic > 0) {
            // all but the newest store have finished reading shared memory
            tma_store_wait_read<1>();
            mbar_arrive(&eo_empty[(ic - 1) % C::SEO]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tm_empty[acc]);
    }
    if (issuer) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// colgemm: weight-gradient reduction over the pixel dimension (split-K over pixel tiles)
// ------------------------------------------------------------------------------------------
template <int NY>
struct ColGemmCfg {
  static_assert(NY % 64 == 0 && NY >= 64 && NY <= 256, "operand width");
  static constexpr int XC = 2;        // X chunks per stage (M = 128 output rows)
  static constexpr int YC = NY / 64;  // Y chunks per stage
  static constexpr uint32_t STAGE_BYTES = (XC + YC) * kChunkBytes;
  static constexpr int STAGES = (2 * STAGE_BYTES + 8192 <= 232448) ? ((3 * STAGE_BYTES + 8192 <= 232448) ? 3 : 2) : 1;
  static constexpr uint32_t OFF_ONES = STAGES * STAGE_BYTES;
  static constexpr uint32_t ONES_BYTES = 1024;
  static constexpr uint32_t OFF_BAR = OFF_ONES + ONES_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 1;
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
  static constexpr uint32_t TMEM_COLS = tmem_cols_pow2(NY + 16);
  static_assert(STAGES >= 2, "need a double-buffered pipeline");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of shared memory");
};

// One job = one (problem, 128-row output block, pixel split).
struct ColGemmJobs {
  int num_problems;     // e.g. hidden layers 1..D-2
  int mblocks;          // output row blocks per problem (NX / 128)
  int splits;           // pixel splits per (problem, mblock)
  int tile0;            // first 128-pixel tile of this launch (row chunks)
  int tiles_total;      // 128-pixel tiles in the launch's row range
  int tiles_per_split;  // ceil(tiles_total / splits)
  int accumulate;       // add to the existing partials instead of overwriting them
  int x_row0[8];        // first row of problem p in the X (dZ) tensor map
  int y_row0[8];        // first row of problem p in the Y (activation) tensor map
  float* dw_partial;    // [splits][num_problems][NX][NY] fp32
  float* db_partial;    // [splits][num_problems][NX] fp32
  int nx;               // rows of dW per problem (= X width)
};

template <int NY>
__global__ void __launch_bounds__(256, 1)
colgemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
               const ColGemmJobs jobs, const uint32_t idesc_main, const uint32_t idesc_ones) {
  using C = ColGemmCfg<NY>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* done = empty + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // job decode
  const int job = blockIdx.x;
  const int split = job % jobs.splits;
  const int mb = (job / jobs.splits) % jobs.mblocks;
  const int prob = job / (jobs.splits * jobs.mblocks);
  const int tile_begin = split * jobs.tiles_per_split;
  int tile_end = tile_begin + jobs.tiles_per_split;
  if (tile_end > jobs.tiles_total) tile_end = jobs.tiles_total;
  const int ntiles = tile_end > tile_begin ? tile_end - tile_begin : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 4) {
    // 16x16 tile of fp16 ones: B operand of the bias-gradient (column-sum) MMA
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + C::OFF_ONES);
    for (int i = threadIdx.x - 128; i < int(C::ONES_BYTES / 4); i += 128) ones[i] = 0x3C003C00u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < ntiles; ++i) {
        const uint32_t s = i % C::STAGES, ph = (i / C::STAGES) & 1u;
        const int prow = (jobs.tile0 + tile_begin + i) * kRowsPerTile;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], C::STAGE_BYTES);
        uint8_t* st = smem + s * C::STAGE_BYTES;
        for (int c = 0; c < C::XC; ++c)
          tma_load_2d(st + c * kChunkBytes, &tmX, &full[s], mb * 128 + c * 64,
                      jobs.x_row0[prob] + prow);
        for (int c = 0; c < C::YC; ++c)
          tma_load_2d(st + (C::XC + c) * kChunkBytes, &tmY, &full[s], c * 64,
                      jobs.y_row0[prob] + prow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t ones_addr = smem_u32(smem + C::OFF_ONES);
      const uint64_t d_ones = umma_smem_desc(ones_addr, 128, 256, 0);
      for (int i = 0; i < ntiles; ++i) {
        const uint32_t s = i % C::STAGES, ph = (i / C::STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t x_addr = smem_u32(smem + s * C::STAGE_BYTES);
        const uint32_t y_addr = x_addr + C::XC * kChunkBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 8 x 16 pixels
          const uint64_t dx = umma_smem_desc(x_addr + k * 2048, kChunkBytes, 1024, 2);
          const uint64_t dy = umma_smem_desc(y_addr + k * 2048, kChunkBytes, 1024, 2);
          const uint32_t accum = (i | k) != 0 ? 1u : 0u;
          umma_f16(tmem_base, dx, dy, idesc_main, accum);
          umma_f16(tmem_base + NY, dx, d_ones, idesc_ones, accum);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(done);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int m = mb * 128 + q * 32 + lane;  // output row of this thread
    float* dw = jobs.dw_partial +
                ((size_t(split) * jobs.num_problems + prob) * jobs.nx + m) * size_t(NY);
    float* dbp = jobs.db_partial + (size_t(split) * jobs.num_problems + prob) * jobs.nx + m;
    if (ntiles > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      for (int cb = 0; cb < NY / 32; ++cb) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + cb * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          float4* dst = reinterpret_cast<float4*>(dw + cb * 32) + j;
          if (jobs.accumulate) {
            const float4 old = *dst;
            o.x += old.x;
            o.y += old.y;
            o.z += old.z;
            o.w += old.w;
          }
          *dst = o;
        }
      }
      uint32_t b8[8];
      tmem_ld_32x8(tmem_base + (uint32_t(q * 32) << 16) + NY, b8);
      tmem_ld_wait();
      *dbp = __uint_as_float(b8[0]) + (jobs.accumulate ? *dbp : 0.f);
    } else if (!jobs.accumulate) {
      for (int j = 0; j < NY / 4; ++j) reinterpret_cast<uint4*>(dw)[j] = make_uint4(0, 0, 0, 0);
      *dbp = 0.0f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace sb
