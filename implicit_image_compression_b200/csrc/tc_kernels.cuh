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
constexpr int kMaxOutTc = 4;                // output channels handled by the tensor-core last layer

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

// dZ pair of the layer below: {acc0 * cos0, acc1 * cos1} as packed 16-bit floats, cos = +-sqrt(1 - a^2) rebuilt from
// the signed-half pair `e`.  The sign goes on AFTER rounding, as one XOR on the packed pair (rounding is symmetric,
// so the bits equal pack(acc * cos_from_signed_half(.))): one instruction per element less than two float XORs.
template <bool OUT_BF16>
__device__ __forceinline__ uint32_t dz_pair_from_signed_half(float acc0, float acc1, uint32_t e) {
  const float a0 = __half2float(__ushort_as_half(static_cast<unsigned short>(e & 0xFFFFu)));
  const float a1 = __half2float(__ushort_as_half(static_cast<unsigned short>(e >> 16)));
  const float c0 = fast_sqrt(__saturatef(fmaf(-a0, a0, 1.0f)));
  const float c1 = fast_sqrt(__saturatef(fmaf(-a1, a1, 1.0f)));
  const uint32_t mag = OUT_BF16 ? pack_bf16x2(acc0 * c0, acc1 * c1) : pack_f16x2(acc0 * c0, acc1 * c1);
  return mag ^ ((e << 15) & 0x80008000u);
}

// Where the input coordinates of pixel p (local index inside this handle's rows) come from.
struct CoordSrc {
  const float* lin_h;   // [H] or null
  const float* lin_w;   // [W] or null
  const float* coords;  // [npix, 2] or null
  int width;            // image width
  int row_begin;        // first image row of this handle
  int64_t p_offset;     // pixel offset of the current launch inside the handle's rows (row chunks)
};

// siren.py:125-128: x = (grid - 0.5) * 2, features ordered (h, w) (data.py:82-86, 'ij' meshgrid)
__device__ __forceinline__ void load_xy(const CoordSrc& c, int64_t p, float& xh, float& xw) {
  float gh, gw;
  p += c.p_offset;
  if (c.coords) {
    const float2 v = reinterpret_cast<const float2*>(c.coords)[p];
    gh = v.x;
    gw = v.y;
  } else {
    const unsigned pu = unsigned(p);  // npix < 2^31 (checked at create)
    const int r = int(pu / unsigned(c.width)), col = int(pu - unsigned(r) * unsigned(c.width));
    gh = __ldg(c.lin_h + c.row_begin + r);
    gw = __ldg(c.lin_w + col);
  }
  xh = (gh - 0.5f) * 2.0f;
  xw = (gw - 0.5f) * 2.0f;
}

// Pace hint between two roles of ONE launch that sweep the same tensors (the dX GEMM and the weight-gradient
// reduction of a layer both read dz[l] and act[l-1]): each role counts the tiles whose loads it has issued and
// holds its own loads back while it is more than `window` tiles ahead of the other, so that whichever role
// touches a tile second finds it in the 126 MB L2 instead of HBM.  It is a hint, not a dependency: the wait is
// bounded.
struct PaceCtx {
  unsigned int* mine;         // tiles (x my_per_tile) this role has issued; null = no pacing
  const unsigned int* other;  // the other role's counter
  int window;                 // this role may run `window` tiles AHEAD of the other's frontier; negative: it stays
                              // |window| tiles BEHIND it (it then finds every tile in L2)
  int my_per_tile, other_per_tile;
  int total;                  // tiles of the launch: a follower's target is clamped to the leader's last tile
};
// `seen`: the caller's copy of the other role's frontier (monotonic): the counter - one L2 line that every CTA of
// the launch updates, ~1 k cycles per read - is only read again when the copy no longer covers the request.
__device__ __forceinline__ void pace_wait(const PaceCtx& pc, int tile, int& seen) {
  if (!pc.mine) return;
  int need = tile - pc.window;  // the other role's frontier must have reached this tile
  if (need > pc.total) need = pc.total;
  if (need <= seen) return;
  // the other role's frontier ends at the last tile once all its CTAs are done, so this can only wait on a role
  // that is still running; bounded anyway (~0.1 ms) so that a pacing mistake costs time, never a hang
  for (int spins = 0; spins < 1024; ++spins) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(pc.other) : "memory");
    seen = int(v / unsigned(pc.other_per_tile));
    if (need <= seen) return;
    __nanosleep(64);
  }
}
__device__ __forceinline__ void pace_post(const PaceCtx& pc, unsigned int n = 1u) {
  if (pc.mine) atomicAdd(pc.mine, n);
}

// Debug stall accounting: cycles spent inside a wait, accumulated per single-thread role (no cost when off
// beyond a predicated branch).
#define SB_WAIT_TIMED(st, acc, stmt)            \
  do {                                          \
    if (st) {                                   \
      const long long _t0 = clock64();          \
      stmt;                                     \
      (acc) += clock64() - _t0;                 \
    } else {                                    \
      stmt;                                     \
    }                                           \
  } while (0)

// ------------------------------------------------------------------------------------------
// rowgemm
// ------------------------------------------------------------------------------------------
// NDIM = output columns per work item (<= 256, one UMMA N); NPARTS = output width / NDIM.  When the whole
// B operand (NDIM x KDIM fp16) fits next to the rings it stays RESIDENT in shared memory for the kernel's
// lifetime (hidden <= 256); otherwise (hidden 512) its 64-wide k-blocks are streamed through the same
// ring stage as the A k-blocks and every 128-row tile is visited once per output part.
template <int KDIM, int NDIM, int MODE, int NPARTS = 1, bool RED = false>
struct RowGemmCfg {
  static_assert(KDIM % 64 == 0 && NDIM % 64 == 0, "hidden size must be a multiple of 64");
  static_assert(NDIM >= 16 && NDIM <= 256, "UMMA N range");
  static constexpr int KB = KDIM / 64;  // 64-wide K blocks
  static constexpr int NB = NDIM / 64;  // 64-wide output chunks
  // ring depth: 4 stages when the (resident-B, forward) configuration has the shared memory for it
#ifndef SB_FWD_SA
#define SB_FWD_SA 4
#define SB_FWD_SEO 2
#endif
  // RED (resident B): the reducer warps hold every staging slot ~2 k cycles longer than the store does, and the MMA
  // (2 k cycles per tile) is far from the limit there: one A stage less buys a fourth staging slot
#ifndef SB_RED_SA
#define SB_RED_SA 2
#define SB_RED_SEO 4
#endif
  static constexpr bool RED_RING = RED && MODE == MODE_DX && KDIM * NDIM * 2 <= 131072;
  // SB_DX_STREAM (experiment): the plain dX GEMM of hidden <= 256 streams its B k-blocks from L2 like hidden 512
  // does instead of keeping the 128 KiB operand resident, which buys two more in-place staging slots
#ifndef SB_DX_STREAM
#define SB_DX_STREAM 0
#endif
  static constexpr bool DX_STREAMED = SB_DX_STREAM && !RED && MODE == MODE_DX && KDIM * NDIM * 2 <= 131072;
#ifndef SB_DX_SA
#define SB_DX_SA 3
#define SB_DX_SEO 3
#endif
  static constexpr bool DX_RING = !RED && MODE == MODE_DX && KDIM * NDIM * 2 <= 131072 && !DX_STREAMED;
  // streamed-B forward (hidden 512): the MMA issuer waited for operands 38 % of the time with three 48 KiB stages
  // (one k-block = 512 cycles of MMA against ~2 k cycles of load latency); the bias table moves out of shared memory
  // (read through L1 instead) to make room for a fourth
#ifndef SB_WIDE_FWD_SA
#define SB_WIDE_FWD_SA 4
#endif
  static constexpr bool WIDE_FWD = MODE == MODE_FWD && KDIM * NDIM * 2 > 131072;
#ifndef SB_WIDE_DX_SA
#define SB_WIDE_DX_SA 3
#define SB_WIDE_DX_SEO 3
#endif
  static constexpr bool WIDE_DX = MODE == MODE_DX && KDIM * NDIM * 2 > 131072;
  static constexpr int SA = RED_RING ? SB_RED_SA
                            : DX_RING ? SB_DX_SA
                                      : WIDE_FWD ? SB_WIDE_FWD_SA
                                      : DX_STREAMED ? 3
                                      : WIDE_DX ? SB_WIDE_DX_SA
                                                 : (MODE == MODE_FWD && KDIM * NDIM * 2 <= 131072) ? SB_FWD_SA : 3;
  static constexpr int SEO = RED_RING ? SB_RED_SEO
                             : DX_RING ? SB_DX_SEO
                             : DX_STREAMED ? 5
                             : WIDE_DX ? SB_WIDE_DX_SEO
                                       : (MODE == MODE_DX) ? 3 : ((KDIM * NDIM * 2 <= 131072) ? SB_FWD_SEO : 2);  // epilogue in/out ring depth
  static constexpr uint32_t B_KB_BYTES = NDIM * 128;
  static constexpr bool STREAM_B = (uint32_t(KB) * B_KB_BYTES > 131072u) || DX_STREAMED;
  static constexpr uint32_t A_STAGE = kChunkBytes + (STREAM_B ? B_KB_BYTES : 0u);
  static constexpr uint32_t OFF_B = 0;
  static constexpr uint32_t OFF_A = OFF_B + (STREAM_B ? 0u : KB * B_KB_BYTES);
  static constexpr uint32_t OFF_EO = OFF_A + SA * A_STAGE;
  static constexpr uint32_t OFF_CONST = OFF_EO + SEO * kChunkBytes;
  static constexpr uint32_t CONST_BYTES = (MODE == MODE_FWD && !(WIDE_FWD && SA > 3)) ? NDIM * NPARTS * 4 : 0;
  static constexpr uint32_t OFF_BAR = OFF_CONST + CONST_BYTES;
  static constexpr int NUM_BARS = 2 * SA + 1 + 3 * SEO + 4 + 4;
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
  // 1: the CTAs sweep the tiles back to front (work item i = tile num_tiles - 1 - i).  Consecutive launches alternate
  // the direction, so a launch starts with the tiles its predecessor wrote LAST - the part of a 201 MB tensor that is
  // still in the 126 MB L2
  int reverse;
  // 1 (MODE_FWD): the A operand (the previous layer's activation: not read again before the backward pass) is loaded
  // with an L2 evict_first hint, and so is the GEN variant's stash store of layer 0's activation, so that what stays
  // in L2 is this launch's OUTPUT - the next launch's input, swept back to front
  int l2_hints;
  float omega;        // MODE_FWD: sine frequency
  const float* bias;  // MODE_FWD: fp32 bias[NDIM]
  const float* bias_w;  // MODE_FWD, streamed B: omega * bias (device table, 16-byte aligned), read through L1
  // GEN (first hidden layer): the A operand is not loaded but GENERATED — layer 0 of the network,
  // a0 = sin(w0 (W0 x + b0)) from in-kernel coordinates (siren.py:62,66 with is_first) — by four extra
  // warps straight into the A ring, and stored to the activation stash from there.
  // RED (dX of the first hidden layer): extra warps (SB_RED_RW per 64-column chunk) read every finished dz[0] chunk back from the
  // staging buffer and accumulate layer 0's weight / bias gradient (dW0 = dz0^T x, db0 = sum dz0; autograd of
  // siren.py:62 for the is_first layer), so dz[0] is never re-read from HBM.  red_part: [kRedWarpsPerChunk * gridDim.x][3 * hidden]
  // = per-CTA partials {dW0 [NDIM, 2], db0 [NDIM]}.
  float* red_part;
  CoordSrc gen_coord;   // GEN and RED: coordinates of this launch's pixels
  const float* gen_w0;  // [KDIM, 2] fp32
  const float* gen_b0;  // [KDIM]
  float gen_omega;
  // 1: the resident B operand was written at least two kernels ago, so it may be fetched BEFORE waiting for
  // the predecessor kernel (programmatic dependent launch); 0: fetch it after the wait
  int b_early;
  // SIRENB200_STALLS (debug): cycles each single-thread role of a CTA spent waiting, stall[cta * 16 + k]
  long long* stall;
  long long* gen_tl;  // SIRENB200_TIMELINE: clock64 stamps of block 0 (debug)
};

// timeline slots of the GEN variant: tl[(role * 8 + tile) * 8 + k], block 0, first 8 tiles;
// roles 0..3 = generator warps, 4 = MMA issuer, 5 = epilogue
#define SB_DBG_G(role, tile_i, k)                                                   \
  do {                                                                              \
    if (GEN && args.gen_tl && cta == 0 && (tile_i) < 8)                      \
      args.gen_tl[((role) * 8 + (tile_i)) * 8 + (k)] = clock64();                   \
  } while (0)

#ifndef SB_DX_EPW
#define SB_DX_EPW 16
#endif
// RED (the first hidden layer's dX GEMM also reduces layer 0's gradient): SB_RED_RW reducer warps per 64-column
// chunk (each takes 128 / SB_RED_RW of the tile's pixel rows) next to SB_RED_EPW epilogue warps
#ifndef SB_RED_RW
#define SB_RED_RW 2
#define SB_RED_EPW 16
#endif
constexpr int kRedWarpsPerChunk = SB_RED_RW;
// `wide`: the streamed-B configuration (hidden 512).  Its forward GEMM is not HBM-bound like hidden 256's: two sin
// epilogues per 128-pixel tile against 8 k cycles of MMA - but 16 epilogue warps measured no faster than 8 there
#ifndef SB_FWD_WIDE_EPW
#define SB_FWD_WIDE_EPW 8
#endif
__host__ __device__ constexpr int rowgemm_epi_warps(int mode, bool gen, bool red, bool wide = false) {
  return (mode == MODE_DX && !gen) ? (red ? SB_RED_EPW : SB_DX_EPW) : ((wide && !gen) ? SB_FWD_WIDE_EPW : 8);
}
__host__ __device__ constexpr int rowgemm_threads(int mode, bool gen, bool red, bool wide = false) {
  return gen ? 576 : 32 * (4 + rowgemm_epi_warps(mode, gen, red, wide) + (red ? 4 * SB_RED_RW : 0));
}

template <int KDIM, int NDIM, int MODE, bool OUT_BF16, int NPARTS = 1, bool GEN = false, bool RED = false>
// warps: 0 = TMA producer, 1 = MMA issuer, 2 = epilogue-input producer (MODE_DX), 3 = store warp,
// 4..4+EPW-1 = epilogue (EPW / 4 warps per TMEM lane quadrant; each takes 64 / (EPW / 4) of the 64 columns of
// every output chunk; EPW = 16 for the dX GEMMs, 8 otherwise);
// GEN: warps 0, 2, 3 and 12..16 = A-operand generators (network layer 0), two per 64-wide k-block; the
// weight load moves to warp 1 and the store warp is warp 17;  RED: 4 * SB_RED_RW more warps = layer-0 gradient
// reducers (SB_RED_RW per 64-column chunk).
// `cta` / `ncta`: index of this CTA among the CTAs of the role and their number (== cta / ncta when the
// whole grid runs this body).
__device__ __forceinline__ void
rowgemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmE, const CUtensorMap& tmO,
             const RowGemmArgs& args, const uint32_t idesc, const int cta, const int ncta, const PaceCtx pace) {
  using C = RowGemmCfg<KDIM, NDIM, MODE, NPARTS, RED>;
  // epilogue warps: 8 (two per TMEM lane quadrant, 32 of a chunk's 64 columns each), or 16 for the plain dX GEMM
  // (four per quadrant, 16 columns each): its cvt -> FFMA -> MUFU.SQRT -> FMUL -> F2FP chains need more than two
  // warps per scheduler to hide their latency (stall accounting: the 8-warp epilogue was busy 5.1 k cycles per tile)
  constexpr int EPW = rowgemm_epi_warps(MODE, GEN, RED, C::STREAM_B);
  constexpr int CPW = 64 / (EPW / 4);
  const int num_items = args.num_tiles * NPARTS;  // item = (tile, output part)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // SIRENB200_STALLS: lifetime of the CTA in SM cycles and in globaltimer ns (-> effective SM clock, set-up and drain
  // time outside the role loops, gaps between consecutive launches)
  long long life_c0 = 0;
  unsigned long long life_g0 = 0;
  if (args.stall && threadIdx.x == 0) {
    life_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(life_g0));
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::SA;
  uint64_t* b_full = a_empty + C::SA;
  uint64_t* eo_full = b_full + 1;
  uint64_t* eo_empty = eo_full + C::SEO;
  uint64_t* red_full = eo_empty + C::SEO;
  uint64_t* tm_full = red_full + 4;  // red_full: one per 64-column chunk index (RED)
  uint64_t* tm_empty = tm_full + 2;
  uint64_t* o_ready = tm_empty + 2;  // [SEO] every epilogue warp has written (and fenced) its part of the chunk
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // The finished chunks are taken to HBM by a dedicated store warp (the spare warp 3; with GEN, where warp 3
  // generates, the last warp): the epilogue warps hand a chunk over through o_ready and go straight on to the next
  // one, so neither a CTA-wide barrier nor the TMA issue / wait_group latency sits on their path.
  constexpr int STW = GEN ? 17 : 3;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::SEO; ++i) mbar_init(&o_ready[i], EPW);
    for (int i = 0; i < C::SA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(b_full, 1);
    for (int i = 0; i < C::SEO; ++i) {
      mbar_init(&eo_full[i], 1);
      mbar_init(&eo_empty[i], RED ? 1 + kRedWarpsPerChunk : 1);  // RED: the store has read the chunk AND its reducer warps have
    }
    for (int i = 0; i < 4; ++i) mbar_init(&red_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tm_full[i], 1);
      mbar_init(&tm_empty[i], EPW);
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
  if (MODE == MODE_FWD && C::CONST_BYTES > 0 && warp >= 4) {
    float* cst = reinterpret_cast<float*>(smem + C::OFF_CONST);
    for (int i = threadIdx.x - 128; i < NDIM * NPARTS; i += 32 * EPW) cst[i] = args.omega * args.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // resident B: issued before the dependency wait when it does not depend on the predecessor kernel
  const bool b_loader = !C::STREAM_B && lane == 0 && warp == (GEN ? 1 : 0);
  if (b_loader && args.b_early) {
    mbar_expect_tx(b_full, C::KB * C::B_KB_BYTES);
    for (int kb = 0; kb < C::KB; ++kb)
      tma_load_2d(smem + C::OFF_B + kb * C::B_KB_BYTES, &tmB, b_full, kb * 64, 0);
  }
  pdl_wait();
  pdl_launch_dependents();
  if (b_loader && !args.b_early) {
    mbar_expect_tx(b_full, C::KB * C::B_KB_BYTES);
    for (int kb = 0; kb < C::KB; ++kb)
      tma_load_2d(smem + C::OFF_B + kb * C::B_KB_BYTES, &tmB, b_full, kb * 64, 0);
  }

  // generator index (GEN only): warps 0, 2, 3, 12, 13, 14, 15, 16 -> 0..7
  const int gi = !GEN ? -1 : (warp == 0 ? 0 : (warp == 2 || warp == 3) ? warp - 1 : ((warp >= 12 && warp <= 16) ? warp - 9 : -1));
  if (GEN && gi >= 0) {
    // ===================== A generator: layer 0 of the network -> A ring (+ stash) =====================
    // Two warps per 64-wide k-block (16 of the 32 four-row iterations each): the 8 column groups x 3
    // layer-0 parameters of a lane stay in registers for the kernel's lifetime; lane = (column group,
    // row group).
    static_assert(!GEN || (NPARTS == 1 && !C::STREAM_B && MODE == MODE_FWD && C::KB <= 4 && C::SA % C::KB == 0),
                  "GEN: resident-B forward only");
    const int kb = GEN ? (gi >> 1) : 0, half = gi & 1;  // (GEN ? :) keeps the dead non-GEN instantiation warning-free
    if (kb < C::KB) {
      const int cg = lane & 7;   // 16-byte column group inside the k-block
      const int rg = lane >> 3;  // rows rg + 4 i
      const bool issuer_g = (half == 0 && lane == 0);
      const CoordSrc& cs = args.gen_coord;
      float wa[8], wb[8], bb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = kb * 64 + cg * 8 + j;
        wa[j] = __ldg(args.gen_w0 + col * 2 + 0) * args.gen_omega;
        wb[j] = __ldg(args.gen_w0 + col * 2 + 1) * args.gen_omega;
        bb[j] = __ldg(args.gen_b0 + col) * args.gen_omega;
      }
      uint32_t ig = uint32_t(kb);
      for (int t = cta; t < args.num_tiles; t += ncta, ig += C::KB) {
        const uint32_t s = ig % C::SA, ph = (ig / C::SA) & 1u;
        const uint32_t stage = smem_u32(smem + C::OFF_A + s * C::A_STAGE);
        const int r_first = rg + 64 * half;
        // Coordinates once per tile: lane L owns rows 64*half + L and 64*half + 32 + L of the tile (already
        // mapped to [-1, 1]); the row loop below fetches them with two shuffles instead of recomputing
        // (row, column) and reloading the tables for every row.
        float own_h[2], own_w[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int r = t * kRowsPerTile + 64 * half + 32 * q + lane;
          own_h[q] = own_w[q] = 0.f;
          if (r < args.valid_rows) load_xy(cs, r, own_h[q], own_w[q]);
        }
        // rows >= valid_rows exist only in the last tile: their activations are written as zero
        const bool full_tile = (t + 1) * kRowsPerTile <= args.valid_rows;
        // every lane waits (no divergent region in front of the arithmetic); only the issuer owns bulk groups
        if (issuer_g) SB_DBG_G(kb, ig / C::KB, 0);
        mbar_wait(&a_empty[s], ph ^ 1u);                   // the MMAs that read this stage are done
        tma_store_wait_read<C::SA / C::KB - 1>();          // ... and so is the stash store issued from it
        named_bar_sync(2 + kb, 64);
        if (issuer_g) SB_DBG_G(kb, ig / C::KB, 1);
        // batches of 8 rows: the arithmetic of a batch, then its stores (the shared stores are asm volatile
        // with a memory clobber; interleaving them with the arithmetic would serialise it)
#pragma unroll
        for (int ib = 0; ib < 16; ib += 8) {
          uint32_t o[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // row rg + 4 (ib + i) of this half: owner lane (rg + 4 (ib+i)) & 31, slot (ib + i) >> 3 (rg < 4)
            const int src = (rg + 4 * (ib + i)) & 31;
            const float x0 = __shfl_sync(0xffffffffu, own_h[(ib + i) >> 3], src);
            const float x1 = __shfl_sync(0xffffffffu, own_w[(ib + i) >> 3], src);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float t0 = fmaf(x0, wa[2 * j], fmaf(x1, wb[2 * j], bb[2 * j]));
              const float t1 = fmaf(x0, wa[2 * j + 1], fmaf(x1, wb[2 * j + 1], bb[2 * j + 1]));
              o[i][j] = sine_signed_half2(t0, t1);
            }
          }
          if (!full_tile) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t keep = (t * kRowsPerTile + r_first + 4 * (ib + i) < args.valid_rows) ? 0xFFFFFFFFu : 0u;
#pragma unroll
              for (int j = 0; j < 4; ++j) o[i][j] &= keep;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rt = r_first + 4 * (ib + i);
            st_shared_v4(stage + rt * 128 + ((uint32_t(cg) ^ uint32_t(rt & 7)) << 4), o[i][0], o[i][1],
                         o[i][2], o[i][3]);
          }
        }
        if (issuer_g) SB_DBG_G(kb, ig / C::KB, 2);
        fence_proxy_async_smem();
        named_bar_sync(2 + kb, 64);
        if (issuer_g) {
          SB_DBG_G(kb, ig / C::KB, 3);
          if (args.l2_hints)
            tma_store_2d_hint(&tmA, smem + C::OFF_A + s * C::A_STAGE, kb * 64, args.a_row0 + t * kRowsPerTile,
                              l2_policy_evict_first());
          else
            tma_store_2d(&tmA, smem + C::OFF_A + s * C::A_STAGE, kb * 64, args.a_row0 + t * kRowsPerTile);
          tma_store_commit();
          mbar_arrive(&a_full[s]);
          SB_DBG_G(kb, ig / C::KB, 4);
        }
        __syncwarp();
      }
      if (issuer_g) tma_store_wait_all<0>();
    }
  } else if (warp == 0) {
    // ===================== TMA producer: B once, then A k-blocks =====================
    if (lane == 0) {
      uint32_t ia = 0;
      long long st_pace = 0, st_aempty = 0;
      int pace_seen = 0;
      const uint64_t pol_first = l2_policy_evict_first();
      const long long st_begin = args.stall ? clock64() : 0;
      for (int it = cta; !GEN && it < num_items; it += ncta) {
        const int tl = it / NPARTS, part = it % NPARTS;  // tl: position in the sweep (what the pace hint counts)
        const int t = args.reverse ? args.num_tiles - 1 - tl : tl;
        const int row = args.a_row0 + t * kRowsPerTile;
        SB_WAIT_TIMED(args.stall, st_pace, pace_wait(pace, tl, pace_seen));
        pace_post(pace);
        for (int kb = 0; kb < C::KB; ++kb, ++ia) {
          const uint32_t s = ia % C::SA, ph = (ia / C::SA) & 1u;
          SB_WAIT_TIMED(args.stall, st_aempty, mbar_wait(&a_empty[s], ph ^ 1u));
          mbar_expect_tx(&a_full[s], C::A_STAGE);
          if (MODE == MODE_FWD && args.l2_hints)
            tma_load_2d_hint(smem + C::OFF_A + s * C::A_STAGE, &tmA, &a_full[s], kb * 64, row, pol_first);
          else
            tma_load_2d(smem + C::OFF_A + s * C::A_STAGE, &tmA, &a_full[s], kb * 64, row);
          if (C::STREAM_B)
            tma_load_2d(smem + C::OFF_A + s * C::A_STAGE + kChunkBytes, &tmB, &a_full[s], kb * 64,
                        part * NDIM);
        }
      }
      if (args.stall && !GEN) {
        args.stall[cta * 16 + 0] = st_pace;
        args.stall[cta * 16 + 1] = st_aempty;
        args.stall[cta * 16 + 2] = clock64() - st_begin;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      if (!C::STREAM_B) mbar_wait(b_full, 0);
      tc_fence_after();
      uint32_t ia = 0, it = 0;
      long long st_tmempty = 0, st_afull = 0;
      const long long st_begin = args.stall ? clock64() : 0;
      for (int item = cta; item < num_items; item += ncta, ++it) {
        const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
        SB_DBG_G(4, it, 0);
        SB_WAIT_TIMED(args.stall, st_tmempty, mbar_wait(&tm_empty[acc], aph ^ 1u));
        SB_DBG_G(4, it, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * NDIM;
        for (int kb = 0; kb < C::KB; ++kb, ++ia) {
          const uint32_t s = ia % C::SA, ph = (ia / C::SA) & 1u;
          SB_WAIT_TIMED(args.stall, st_afull, mbar_wait(&a_full[s], ph));
          SB_DBG_G(4, it, 2 + (kb & 3));
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + C::OFF_A + s * C::A_STAGE);
          const uint32_t b_addr = C::STREAM_B ? a_addr + kChunkBytes
                                              : smem_u32(smem + C::OFF_B + kb * C::B_KB_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * 32, 0, 1024, 2);
            const uint64_t db = umma_smem_desc(b_addr + k * 32, 0, 1024, 2);
            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&tm_full[acc]);
        SB_DBG_G(4, it, 6);
      }
      if (args.stall) {
        args.stall[cta * 16 + 3] = st_tmempty;
        args.stall[cta * 16 + 4] = st_afull;
        args.stall[cta * 16 + 5] = clock64() - st_begin;
      }
    }
  } else if (warp == 2) {
    // ===================== epilogue-input producer (MODE_DX only) =====================
    if (MODE == MODE_DX && lane == 0) {
      uint32_t ic = 0;
      long long st_eoempty = 0;
      for (int item = cta; item < num_items; item += ncta) {
        const int t = args.reverse ? args.num_tiles - 1 - item / NPARTS : item / NPARTS, part = item % NPARTS;
        const int row = args.e_row0 + t * kRowsPerTile;
        for (int nb = 0; nb < C::NB; ++nb, ++ic) {
          const uint32_t s = ic % C::SEO, ph = (ic / C::SEO) & 1u;
          SB_WAIT_TIMED(args.stall, st_eoempty, mbar_wait(&eo_empty[s], ph ^ 1u));
          mbar_expect_tx(&eo_full[s], kChunkBytes);
          tma_load_2d(smem + C::OFF_EO + s * kChunkBytes, &tmE, &eo_full[s], part * NDIM + nb * 64,
                      row);
        }
      }
      if (args.stall) args.stall[cta * 16 + 6] = st_eoempty;
    }
  } else if (RED && warp >= 4 + EPW) {
    // ===================== layer-0 gradient: reduce each finished dz[0] chunk over its 128 pixels =========
    static_assert(!RED || (MODE == MODE_DX && NPARTS <= 2 && C::NB <= 4), "RED: dX of the first hidden layer");
    // RW warps per 64-column chunk (128 / RW pixel rows each); lane -> columns part*NDIM + nb*64 + 2*lane, +1.
    // With two output parts (hidden 512) every tile is visited twice, once per part; each part has its own
    // accumulators.
    constexpr int RW = kRedWarpsPerChunk;
    constexpr int QN = (128 / RW + 31) / 32;  // coordinate registers per lane (rows lane + 32 q of the warp's range)
    const int nb = (warp - 4 - EPW) / RW, sub = (warp - 4 - EPW) % RW;
    const int row_lo = sub * 128 / RW, row_hi = (sub + 1) * 128 / RW;
    if (nb < C::NB) {
      const CoordSrc& cs = args.gen_coord;
      float acc[NPARTS][6];
#pragma unroll
      for (int pp = 0; pp < NPARTS; ++pp)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[pp][j] = 0.f;
      const uint32_t lane_off = (uint32_t(lane) & 3u) * 4u;
      // red_full[nb] belongs to this chunk index alone and is waited on once per work item, so its parity
      // cannot alias (a barrier shared between the chunk indices could be probed more than one phase ahead)
      uint32_t ic = uint32_t(nb), il = 0;
      for (int item = cta; item < num_items; item += ncta, ic += C::NB, ++il) {
        const int t = args.reverse ? args.num_tiles - 1 - item / NPARTS : item / NPARTS, part = item % NPARTS;
        // coordinates of rows row_lo + lane + 32 q (zero for the padding rows: their dz is zero anyway)
        float xh[QN], xw[QN];
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          const int rl = row_lo + lane + 32 * q;
          const int r = t * kRowsPerTile + rl;
          xh[q] = xw[q] = 0.f;
          if (rl < row_hi && r < args.valid_rows) load_xy(cs, r, xh[q], xw[q]);
        }
        const uint32_t s = ic % C::SEO;
        const uint32_t buf = smem_u32(smem + C::OFF_EO + s * kChunkBytes);
        mbar_wait(&red_full[nb], il & 1u);
        float sh0 = 0.f, sw0 = 0.f, sb0 = 0.f, sh1 = 0.f, sw1 = 0.f, sb1 = 0.f;
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          const int nrow = (row_hi - row_lo - 32 * q) < 32 ? (row_hi - row_lo - 32 * q) : 32;
#pragma unroll 8
          for (int rr = 0; rr < nrow; ++rr) {
            const int r = row_lo + q * 32 + rr;
            const uint32_t addr = buf + r * 128 + (((uint32_t(lane) >> 2) ^ uint32_t(r & 7)) << 4) + lane_off;
            const uint32_t hv = ld_shared_u32(addr);
            const float v0 = __half2float(__ushort_as_half(static_cast<unsigned short>(hv & 0xFFFFu)));
            const float v1 = __half2float(__ushort_as_half(static_cast<unsigned short>(hv >> 16)));
            const float ch = __shfl_sync(0xffffffffu, xh[q], rr);
            const float cw = __shfl_sync(0xffffffffu, xw[q], rr);
            sh0 = fmaf(v0, ch, sh0);
            sw0 = fmaf(v0, cw, sw0);
            sb0 += v0;
            sh1 = fmaf(v1, ch, sh1);
            sw1 = fmaf(v1, cw, sw1);
            sb1 += v1;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&eo_empty[s]);
#pragma unroll
        for (int pp = 0; pp < NPARTS; ++pp)
          if (pp == part) {
            acc[pp][0] += sh0;
            acc[pp][1] += sw0;
            acc[pp][2] += sb0;
            acc[pp][3] += sh1;
            acc[pp][4] += sw1;
            acc[pp][5] += sb1;
          }
      }
      constexpr int WFULL = NDIM * NPARTS;
      float* part_out = args.red_part + (size_t(cta) * RW + sub) * 3 * WFULL;
#pragma unroll
      for (int pp = 0; pp < NPARTS; ++pp) {
        const int c0 = pp * NDIM + nb * 64 + 2 * lane;
        part_out[c0 * 2 + 0] = acc[pp][0];
        part_out[c0 * 2 + 1] = acc[pp][1];
        part_out[c0 * 2 + 2] = acc[pp][3];
        part_out[c0 * 2 + 3] = acc[pp][4];
        part_out[2 * WFULL + c0] = acc[pp][2];
        part_out[2 * WFULL + c0 + 1] = acc[pp][5];
      }
    }
  } else if (warp >= 4 && warp < 4 + EPW) {
    // ===================== epilogue: TMEM -> f() -> smem -> TMA store =====================
    const int q = warp & 3;
    const int hb = (warp - 4) >> 2;  // which CPW of the 64 columns of a chunk this warp handles
    const int r_in_tile = q * 32 + lane;
    const bool issuer = (threadIdx.x == 128);
    const float* cst = reinterpret_cast<const float*>(smem + C::OFF_CONST);
    uint32_t it = 0, ic = 0;
    long long st_tmfull = 0, st_eo = 0;
    const long long st_begin = args.stall ? clock64() : 0;
    for (int item = cta; item < num_items; item += ncta, ++it) {
      const int t = args.reverse ? args.num_tiles - 1 - item / NPARTS : item / NPARTS, part = item % NPARTS;
      const uint32_t acc = it & 1u, aph = (it >> 1) & 1u;
      if (issuer) SB_DBG_G(5, it, 0);
      SB_WAIT_TIMED(args.stall, st_tmfull, mbar_wait(&tm_full[acc], aph));
      if (issuer) SB_DBG_G(5, it, 1);
      tc_fence_after();
      const bool row_valid = (t * kRowsPerTile + r_in_tile) < args.valid_rows;
      for (int nb = 0; nb < C::NB; ++nb, ++ic) {
        const uint32_t s = ic % C::SEO, ph = (ic / C::SEO) & 1u;
        const uint32_t buf = smem_u32(smem + C::OFF_EO + s * kChunkBytes);
        const uint32_t row_addr = buf + r_in_tile * 128;
        {
          uint32_t v[CPW];
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * NDIM + nb * 64 + hb * CPW;
          if constexpr (CPW == 32) {
            tmem_ld_32x32(taddr, v);
          } else {
            tmem_ld_32x16(taddr, v);
          }
          if (MODE == MODE_DX)
            SB_WAIT_TIMED(args.stall, st_eo, mbar_wait(&eo_full[s], ph));
          else
            SB_WAIT_TIMED(args.stall, st_eo, mbar_wait(&eo_empty[s], ph ^ 1u));
          uint32_t o[CPW / 2];
          if (MODE == MODE_FWD) {
            const int col0 = part * NDIM + nb * 64 + hb * CPW;
            float4 bw[CPW / 4];
            if constexpr (C::CONST_BYTES == 0) {
#pragma unroll
              for (int j = 0; j < CPW / 4; ++j) bw[j] = __ldg(reinterpret_cast<const float4*>(args.bias_w + col0) + j);
            } else {
#pragma unroll
              for (int j = 0; j < CPW / 4; ++j) bw[j] = *reinterpret_cast<const float4*>(cst + col0 + 4 * j);
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CPW / 4; ++j) {
              const float t0 = fmaf(__uint_as_float(v[4 * j]), args.omega, bw[j].x);
              const float t1 = fmaf(__uint_as_float(v[4 * j + 1]), args.omega, bw[j].y);
              const float t2 = fmaf(__uint_as_float(v[4 * j + 2]), args.omega, bw[j].z);
              const float t3 = fmaf(__uint_as_float(v[4 * j + 3]), args.omega, bw[j].w);
              o[2 * j] = sine_signed_half2(t0, t1);
              o[2 * j + 1] = sine_signed_half2(t2, t3);
            }
          } else {
            uint32_t e[CPW / 2];
#pragma unroll
            for (int c4 = 0; c4 < CPW / 8; ++c4) {
              const uint32_t chunk = uint32_t(hb * (CPW / 8) + c4) ^ uint32_t(r_in_tile & 7);
              const uint4 ld = ld_shared_v4(row_addr + (chunk << 4));
              e[4 * c4 + 0] = ld.x;
              e[4 * c4 + 1] = ld.y;
              e[4 * c4 + 2] = ld.z;
              e[4 * c4 + 3] = ld.w;
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CPW / 2; ++j) {
              const uint32_t g = dz_pair_from_signed_half<OUT_BF16>(__uint_as_float(v[2 * j]),
                                                                    __uint_as_float(v[2 * j + 1]), e[j]);
              o[j] = row_valid ? g : 0u;
            }
          }
#pragma unroll
          for (int c4 = 0; c4 < CPW / 8; ++c4) {
            const uint32_t chunk = uint32_t(hb * (CPW / 8) + c4) ^ uint32_t(r_in_tile & 7);
            st_shared_v4(row_addr + (chunk << 4), o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2],
                         o[4 * c4 + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_ready[s]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tm_empty[acc]);
      if (issuer) SB_DBG_G(5, it, 2);
    }
    if (issuer && args.stall) {
      args.stall[cta * 16 + 7] = st_tmfull;
      args.stall[cta * 16 + 8] = st_eo;
      args.stall[cta * 16 + 11] = clock64() - st_begin;
    }
  } else if (warp == STW) {
    // ===================== store warp: finished chunks -> HBM, staging buffers back to their producer ==========
    if (lane == 0) {
      uint32_t ic = 0;
      long long st_ready = 0, st_rd = 0;
      for (int item = cta; item < num_items; item += ncta) {
        const int t = args.reverse ? args.num_tiles - 1 - item / NPARTS : item / NPARTS, part = item % NPARTS;
        for (int nb = 0; nb < C::NB; ++nb, ++ic) {
          const uint32_t s = ic % C::SEO, ph = (ic / C::SEO) & 1u;
          SB_WAIT_TIMED(args.stall, st_ready, mbar_wait(&o_ready[s], ph));
          if (RED) {
            // dz[0] has no reader but the reducer warps of this CTA (nothing lies below layer 0): it never goes
            // to HBM - 2 bytes per pixel and feature that the step used to write for nobody
            mbar_arrive(&red_full[nb]);
          } else {
            tma_store_2d(&tmO, smem + C::OFF_EO + s * kChunkBytes, part * NDIM + nb * 64,
                         args.o_row0 + t * kRowsPerTile);
            tma_store_commit();
            // this warp has nothing else to do: wait for the store to have read the buffer, hand it straight back
            SB_WAIT_TIMED(args.stall, st_rd, tma_store_wait_read<0>());
          }
          mbar_arrive(&eo_empty[s]);
        }
      }
      if (args.stall) {
        args.stall[cta * 16 + 9] = st_ready;
        args.stall[cta * 16 + 10] = st_rd;
      }
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
  if (args.stall && threadIdx.x == 0) {
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    args.stall[cta * 16 + 12] = clock64() - life_c0;
    args.stall[cta * 16 + 13] = (long long)(g1 - life_g0);
    args.stall[cta * 16 + 14] = (long long)life_g0;
    args.stall[cta * 16 + 15] = (long long)g1;
  }
}

template <int KDIM, int NDIM, int MODE, bool OUT_BF16, int NPARTS = 1, bool GEN = false, bool RED = false>
__global__ void __launch_bounds__(rowgemm_threads(MODE, GEN, RED, RowGemmCfg<KDIM, NDIM, MODE, NPARTS, RED>::STREAM_B), 1)
rowgemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmO,
               const RowGemmArgs args, const uint32_t idesc) {
  rowgemm_body<KDIM, NDIM, MODE, OUT_BF16, NPARTS, GEN, RED>(tmA, tmB, tmE, tmO, args, idesc, int(blockIdx.x),
                                                             int(gridDim.x), PaceCtx{});
}

// ------------------------------------------------------------------------------------------
// colgemm: weight-gradient reduction over the pixel dimension (split-K over pixel tiles)
// ------------------------------------------------------------------------------------------
// NY = dW columns per job (<= 256, one UMMA N); wider layers (hidden 512) are split into column parts.
template <int NY>
struct ColGemmCfg {
  static_assert(NY % 64 == 0 && NY >= 64 && NY <= 256, "operand width");
  static constexpr int PXS = 128;                       // pixels per pipeline stage (one tile)
  static constexpr int XC = 2;        // X chunks per stage (M = 128 output rows)
  static constexpr int YC = NY / 64;  // Y chunks per stage
  static constexpr int SUB = 128 / PXS;                 // stages per 128-pixel tile
  static constexpr uint32_t CHUNK = PXS * 128;          // one {64 x PXS-row} fp16 box
  static constexpr uint32_t STAGE_BYTES = (XC + YC) * CHUNK;
  static constexpr int STAGES = (2 * STAGE_BYTES + 8192 <= 232448)
                                    ? ((4 * STAGE_BYTES + 8192 <= 232448) ? 4 : ((3 * STAGE_BYTES + 8192 <= 232448) ? 3 : 2))
                                    : 1;
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
  int nparts;           // output column parts per problem (width / NY)
  int splits;           // pixel splits per (problem, mblock, part)
  int tile0;            // first 128-pixel tile of this launch (row chunks)
  int tiles_total;      // 128-pixel tiles in the launch's row range
  int tiles_per_split;  // ceil(tiles_total / splits)
  int accumulate;       // add to the existing partials instead of overwriting them
  int x_row0[8];        // first row of problem p in the X (dZ) tensor map
  int y_row0[8];        // first row of problem p in the Y (activation) tensor map
  float* dw_partial;    // [splits][num_problems][NX][ny_total] fp32
  float* db_partial;    // [splits][num_problems][NX] fp32
  int nx;               // rows of dW per problem (= X width)
  int ny_total;         // columns of dW per problem (= Y width = NY * nparts)
  int prob0;            // index of this launch's first problem inside the partial buffers
  int prob_total;       // problems per split slab in the partial buffers
  int interleave;       // 1: split s visits tiles s, s+splits, ... (sweeps the image front to back,
                        //    in step with a concurrently running rowgemm); 0: contiguous tile ranges
  long long* stall;     // SIRENB200_STALLS (debug): stall[job * 16 + k]
  int reverse;          // 1: sweep position x = tile tiles_total - 1 - x (see RowGemmArgs::reverse)
};

template <int NY>
__device__ __forceinline__ void
colgemm_body(const CUtensorMap& tmX, const CUtensorMap& tmY, const ColGemmJobs& jobs, const uint32_t idesc_main,
             const uint32_t idesc_ones, const int job, const PaceCtx pace) {
  using C = ColGemmCfg<NY>;
  constexpr int PXS = C::PXS;
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
  const int split = job % jobs.splits;
  const int part = (job / jobs.splits) % jobs.nparts;
  const int mb = (job / (jobs.splits * jobs.nparts)) % jobs.mblocks;
  const int prob = job / (jobs.splits * jobs.nparts * jobs.mblocks);
  int tile_begin, tile_step, ntiles;
  if (jobs.interleave) {
    tile_begin = split;
    tile_step = jobs.splits;
    ntiles = split < jobs.tiles_total ? (jobs.tiles_total - split + jobs.splits - 1) / jobs.splits : 0;
  } else {
    tile_begin = split * jobs.tiles_per_split;
    tile_step = 1;
    int tile_end = tile_begin + jobs.tiles_per_split;
    if (tile_end > jobs.tiles_total) tile_end = jobs.tiles_total;
    ntiles = tile_end > tile_begin ? tile_end - tile_begin : 0;
  }

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
  pdl_wait();  // everything below reads what the preceding kernels of the step produced
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      long long st_pace = 0, st_empty = 0;
      int pace_seen = 0;
      const long long st_begin = jobs.stall ? clock64() : 0;
      for (int ii = 0; ii < ntiles * C::SUB; ++ii) {
        const int i = ii / C::SUB, sub = ii % C::SUB;
        const uint32_t s = ii % C::STAGES, ph = (ii / C::STAGES) & 1u;
        const int tpos = tile_begin + i * tile_step;
        const int prow = (jobs.tile0 + (jobs.reverse ? jobs.tiles_total - 1 - tpos : tpos)) * kRowsPerTile + sub * PXS;
        if (sub == 0) {
          SB_WAIT_TIMED(jobs.stall, st_pace, pace_wait(pace, tile_begin + i * tile_step, pace_seen));
          pace_post(pace);
        }
        SB_WAIT_TIMED(jobs.stall, st_empty, mbar_wait(&empty[s], ph ^ 1u));
        mbar_expect_tx(&full[s], C::STAGE_BYTES);
        uint8_t* st = smem + s * C::STAGE_BYTES;
        // one box per operand: {64 columns, PXS pixel rows, XC / YC chunks} (make_tmap_16bit_chunks) - a single
        // thread pays ~100 cycles per TMA instruction, six 2-D boxes per stage cost half the MMA time of a stage
        tma_load_3d(st, &tmX, &full[s], 0, jobs.x_row0[prob] + prow, mb * C::XC);
        tma_load_3d(st + C::XC * C::CHUNK, &tmY, &full[s], 0, jobs.y_row0[prob] + prow, part * C::YC);
      }
      if (jobs.stall) {
        jobs.stall[job * 16 + 0] = st_pace;
        jobs.stall[job * 16 + 1] = st_empty;
        jobs.stall[job * 16 + 2] = clock64() - st_begin;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t ones_addr = smem_u32(smem + C::OFF_ONES);
      const uint64_t d_ones = umma_smem_desc(ones_addr, 128, 256, 0);
      long long st_full = 0;
      const long long st_begin = jobs.stall ? clock64() : 0;
      for (int i = 0; i < ntiles * C::SUB; ++i) {
        const uint32_t s = i % C::STAGES, ph = (i / C::STAGES) & 1u;
        SB_WAIT_TIMED(jobs.stall, st_full, mbar_wait(&full[s], ph));
        tc_fence_after();
        const uint32_t x_addr = smem_u32(smem + s * C::STAGE_BYTES);
        const uint32_t y_addr = x_addr + C::XC * C::CHUNK;
#pragma unroll
        for (int k = 0; k < PXS / 16; ++k) {  // 16 pixels per MMA
          const uint64_t dx = umma_smem_desc(x_addr + k * 2048, C::CHUNK, 1024, 2);
          const uint64_t dy = umma_smem_desc(y_addr + k * 2048, C::CHUNK, 1024, 2);
          const uint32_t accum = (i | k) != 0 ? 1u : 0u;
          umma_f16(tmem_base, dx, dy, idesc_main, accum);
          umma_f16(tmem_base + NY, dx, d_ones, idesc_ones, accum);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(done);
      if (jobs.stall) {
        jobs.stall[job * 16 + 4] = st_full;
        jobs.stall[job * 16 + 5] = clock64() - st_begin;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp & 3;
    const int m = mb * 128 + q * 32 + lane;  // output row of this thread
    const size_t slab = size_t(split) * jobs.prob_total + jobs.prob0 + prob;
    float* dw = jobs.dw_partial + (slab * jobs.nx + m) * size_t(jobs.ny_total) + part * NY;
    float* dbp = jobs.db_partial + slab * jobs.nx + m;
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
      if (part == 0) {  // the bias gradient does not depend on the column part
        uint32_t b8[8];
        tmem_ld_32x8(tmem_base + (uint32_t(q * 32) << 16) + NY, b8);
        tmem_ld_wait();
        *dbp = __uint_as_float(b8[0]) + (jobs.accumulate ? *dbp : 0.f);
      }
    } else if (!jobs.accumulate) {
      for (int j = 0; j < NY / 4; ++j) reinterpret_cast<uint4*>(dw)[j] = make_uint4(0, 0, 0, 0);
      if (part == 0) *dbp = 0.0f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}



// ------------------------------------------------------------------------------------------
// The same reduction on CTA PAIRS (hidden 256): the two CTAs that used to compute the two 128-row blocks of dW from
// the same pixel tiles now issue ONE cta_group::2 MMA (M = 256 = all dz features, N = 256).  Each CTA loads only
// its 128 dz features AND only its half of the activation columns (64 KiB per 128-pixel stage instead of 96), reads
// a third less operand data out of shared memory per MMA, and the smaller stage buys a third ring stage - the
// one-CTA version waited for its loads a third of the time with two.  job = (pixel split, CTA rank).
// ------------------------------------------------------------------------------------------
struct ColGemm2Cfg {
  static constexpr int XC = 2, YC = 2;
  static constexpr uint32_t STAGE_BYTES = (XC + YC) * kChunkBytes;  // per CTA
  static constexpr int STAGES = 3;
  static constexpr uint32_t OFF_ONES = STAGES * STAGE_BYTES;
  static constexpr uint32_t OFF_BAR = OFF_ONES + 1024;
  static constexpr int NUM_BARS = 2 * STAGES + 1;
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
  static constexpr uint32_t TMEM_COLS = 512;  // 256 (dW rows of this CTA) + 16 (db), power of two
};

// W = 256: one pair per pixel split.  W = 512: dW is four 256 x 256 blocks (feature half fp x column half ch), one pair
// each, so a pixel split is 8 CTAs as with single CTAs - but every CTA loads 64 KiB per 128-pixel stage instead of 96
// (its 128 dz features + 128 of the block's activation columns) and issues twice the MMA work per loaded byte.
// job = (((problem * FP + fp) * CH + ch) * splits + split) * 2 + CTA rank.
template <int W>
__device__ __forceinline__ void
colgemm2_body(const CUtensorMap& tmX, const CUtensorMap& tmY, const ColGemmJobs& jobs, const int job,
              const PaceCtx pace) {
  static_assert(W == 256 || W == 512, "CTA-pair reduction: hidden 256 or 512");
  using C = ColGemm2Cfg;
  constexpr int FP = W / 256, CH = W / 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* full = bars;               // used in the leader: bytes of BOTH CTAs' loads
  uint64_t* empty = bars + C::STAGES;  // each CTA's own; the leader's commits arrive on both
  uint64_t* done = empty + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // == job & 1: which 128 dz features / which half of the columns
  const bool leader = rank == 0;
  int jd = job >> 1;
  const int split = jd % jobs.splits;
  jd /= jobs.splits;
  const int ch = jd % CH;
  jd /= CH;
  const int fp = jd % FP;
  const int prob = jd / FP;
  int tile_begin, tile_step, ntiles;
  if (jobs.interleave) {
    tile_begin = split;
    tile_step = jobs.splits;
    ntiles = split < jobs.tiles_total ? (jobs.tiles_total - split + jobs.splits - 1) / jobs.splits : 0;
  } else {
    tile_begin = split * jobs.tiles_per_split;
    tile_step = 1;
    int tile_end = tile_begin + jobs.tiles_per_split;
    if (tile_end > jobs.tiles_total) tile_end = jobs.tiles_total;
    ntiles = tile_end > tile_begin ? tile_end - tile_begin : 0;
  }

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
    tmem_alloc_2sm(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  if (warp >= 4) {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + C::OFF_ONES);
    for (int i = threadIdx.x - 128; i < 256; i += int(blockDim.x) - 128) ones[i] = 0x3C003C00u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers exist before anything of ours can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      long long st_pace = 0, st_empty = 0;
      int pace_seen = 0;
      const long long st_begin = jobs.stall ? clock64() : 0;
      for (int i = 0; i < ntiles; ++i) {
        const uint32_t s = i % C::STAGES, ph = (i / C::STAGES) & 1u;
        const int tpos = tile_begin + i * tile_step;
        const int prow = (jobs.tile0 + (jobs.reverse ? jobs.tiles_total - 1 - tpos : tpos)) * kRowsPerTile;
        SB_WAIT_TIMED(jobs.stall, st_pace, pace_wait(pace, tile_begin + i * tile_step, pace_seen));
        pace_post(pace);
        SB_WAIT_TIMED(jobs.stall, st_empty, mbar_wait(&empty[s], ph ^ 1u));
        if (leader) mbar_expect_tx(&full[s], 2 * C::STAGE_BYTES);
        uint8_t* st = smem + s * C::STAGE_BYTES;
        tma_load_3d_2sm(st, &tmX, &full[s], 0, jobs.x_row0[prob] + prow, (fp * 2 + int(rank)) * C::XC);
        tma_load_3d_2sm(st + C::XC * kChunkBytes, &tmY, &full[s], 0, jobs.y_row0[prob] + prow,
                        (ch * 2 + int(rank)) * C::YC);
      }
      if (jobs.stall) {
        jobs.stall[job * 16 + 0] = st_pace;
        jobs.stall[job * 16 + 1] = st_empty;
        jobs.stall[job * 16 + 2] = clock64() - st_begin;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t idesc_main = umma_idesc(256, 256, 0, 0, 1, 1);
      const uint32_t idesc_ones = umma_idesc(256, 16, 0, 0, 1, 1);
      const uint64_t d_ones = umma_smem_desc(smem_u32(smem + C::OFF_ONES), 128, 256, 0);
      long long st_full = 0;
      const long long st_begin = jobs.stall ? clock64() : 0;
      for (int i = 0; i < ntiles; ++i) {
        const uint32_t s = i % C::STAGES, ph = (i / C::STAGES) & 1u;
        SB_WAIT_TIMED(jobs.stall, st_full, mbar_wait(&full[s], ph));
        tc_fence_after();
        const uint32_t x_addr = smem_u32(smem + s * C::STAGE_BYTES);
        const uint32_t y_addr = x_addr + C::XC * kChunkBytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 8 x 16 pixels
          const uint64_t dx = umma_smem_desc(x_addr + k * 2048, kChunkBytes, 1024, 2);
          const uint64_t dy = umma_smem_desc(y_addr + k * 2048, kChunkBytes, 1024, 2);
          const uint32_t accum = (i | k) != 0 ? 1u : 0u;
          umma_f16_2sm(tmem_base, dx, dy, idesc_main, accum);
          if (ch == 0) umma_f16_2sm(tmem_base + 256, dx, d_ones, idesc_ones, accum);  // db: once per feature half
        }
        umma_commit_2sm(&empty[s]);
      }
      umma_commit_2sm(done);
      if (jobs.stall) {
        jobs.stall[job * 16 + 4] = st_full;
        jobs.stall[job * 16 + 5] = clock64() - st_begin;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp & 3;
    const int m = fp * 256 + int(rank) * 128 + q * 32 + lane;  // dW row (dz feature) of this thread
    const size_t slab = size_t(split) * jobs.prob_total + jobs.prob0 + prob;
    float* dw = jobs.dw_partial + (slab * jobs.nx + m) * size_t(jobs.ny_total) + ch * 256;
    float* dbp = jobs.db_partial + slab * jobs.nx + m;
    if (ntiles > 0) {
      mbar_wait(done, 0);
      tc_fence_after();
      for (int cb = 0; cb < 256 / 32; ++cb) {
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
      if (ch == 0) {  // the bias gradient does not depend on the column half
        uint32_t b8[8];
        tmem_ld_32x8(tmem_base + (uint32_t(q * 32) << 16) + 256, b8);
        tmem_ld_wait();
        *dbp = __uint_as_float(b8[0]) + (jobs.accumulate ? *dbp : 0.f);
      }
    } else if (!jobs.accumulate) {
      for (int j = 0; j < 256 / 4; ++j) reinterpret_cast<uint4*>(dw)[j] = make_uint4(0, 0, 0, 0);
      if (ch == 0) *dbp = 0.0f;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs are done with the pair's tensor memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
  }
}

template <int NY>
__global__ void __launch_bounds__(256, 1)
colgemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
               const ColGemmJobs jobs, const uint32_t idesc_main, const uint32_t idesc_ones) {
  colgemm_body<NY>(tmX, tmY, jobs, idesc_main, idesc_ones, int(blockIdx.x), PaceCtx{});
}

// stand-alone launch of the pair reduction (2-CTA clusters): all hidden layers' weight gradients of hidden 512
template <int W>
__global__ void __launch_bounds__(256, 1)
colgemm2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                const ColGemmJobs jobs) {
  colgemm2_body<W>(tmX, tmY, jobs, int(blockIdx.x), PaceCtx{});
}

// ------------------------------------------------------------------------------------------
// One backward layer in ONE launch: CTAs [0, dx_ctas) run the dX GEMM of layer l (rowgemm, MODE_DX), the others
// its weight-gradient reduction (colgemm).  Both roles read dz[l] and act[l-1]; a pace hint (PaceCtx) keeps their
// sweeps within a window of each other, so every tile comes from HBM once and from L2 the second time — the
// reduction's 2 x 201 MB per layer never reach HBM (autograd of nn.Linear: grad_input and grad_weight from ONE
// pass over grad_output, siren.py:62).  pace_counters: two zeroed uint32 (dX, dW) owned by this launch.
// ------------------------------------------------------------------------------------------
// PAIR: launched as 2-CTA clusters; the reduction role runs on CTA pairs (colgemm2_body, hidden 256); the dX role's
// CTAs ignore their cluster.
template <int W, bool RED, bool PAIR = false>
__global__ void __launch_bounds__(rowgemm_threads(MODE_DX, false, RED), 1)
bwd_merged_kernel(const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmWt,
                  const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmDzR,
                  const __grid_constant__ CUtensorMap tmActR, const RowGemmArgs rargs, const uint32_t idesc_row,
                  const ColGemmJobs jobs, const uint32_t idesc_main, const uint32_t idesc_ones, const int dx_ctas,
                  unsigned int* pace_counters, const int window, const int window_dx) {
  constexpr int NT = W < 256 ? W : 256;
  constexpr int NPARTS = W / NT;
  if (int(blockIdx.x) < dx_ctas) {
    const PaceCtx pc{pace_counters, pace_counters + 1, window_dx, NPARTS, jobs.mblocks * jobs.nparts, rargs.num_tiles};
    rowgemm_body<W, NT, MODE_DX, false, NPARTS, false, RED>(tmDz, tmWt, tmAct, tmDz, rargs, idesc_row,
                                                            int(blockIdx.x), dx_ctas, pc);
  } else {
    const PaceCtx pc{pace_counters + 1, pace_counters, window, jobs.mblocks * jobs.nparts, NPARTS, jobs.tiles_total};
    // (tmDzR / tmActR: the reduction role's views of dz / act: one box per operand and stage)
    if constexpr (PAIR) {
      static_assert(!PAIR || W == 256, "CTA-pair reduction: hidden 256");
      colgemm2_body<W>(tmDzR, tmActR, jobs, int(blockIdx.x) - dx_ctas, pc);
    } else {
      colgemm_body<NT>(tmDzR, tmActR, jobs, idesc_main, idesc_ones, int(blockIdx.x) - dx_ctas, pc);
    }
  }
}

// ------------------------------------------------------------------------------------------
// last layer on tensor cores (hidden <= 256, out <= 4): per 128-pixel tile, from ONE read of act[D-2]:
//   y  = act . W_last^T (+ b)            UMMA M=128 px, N=16, K=W       -> pred, squared error, seed g
//   dA = g . (omega W_last)              UMMA M=128 px, N=W,  K=16      -> dz[D-2] = dA (*) +-sqrt(1-a^2)
//   dW_last^T += act^T . g               UMMA M=128 feat, N=16, K=128 px (act tile read MN-major)
// The seed tile g (128 x 16 fp16, 4 KiB, no-swizzle core-matrix layout) is written once by the epilogue
// and read twice: K-major as the A operand of dA and MN-major as the B operand of dW_last.
// (reference: siren.py:64,110-118,131 last SineLayer + /2+0.5; train_helper.py:151-161 mse + backward)
// ------------------------------------------------------------------------------------------
struct LastTcArgs {
  int num_tiles;
  int act_row0;          // first row of act[D-2] inside the activation tensor map
  int dz_row0;           // first row of dz[D-2] inside the dz tensor map
  int64_t npix;          // valid pixels
  const float* b;        // [C] last-layer bias
  const float* img;      // mode 1: target [npix, C]; mode 2: dpred [npix, C]
  float* pred;           // [npix, C] or null
  float* part;           // per-CTA partials [grid][C*W + C + 1] (dW_last, db_last, sum sq err)
  const float* gscale;   // device seed scale G
  int C;
  int mode;              // 0 forward only, 1 MSE, 2 external dpred
  int outermost_linear;
  float omega_last;
  long long* dbg;        // optional timeline capture (block 0): dbg[tile * 16 + k], first 12 tiles
};

#define SB_DBG_L(tile_i, k)                                                 \
  do {                                                                      \
    if (args.dbg && blockIdx.x == 0 && (tile_i) < 12)                       \
      args.dbg[(tile_i) * 16 + (k)] = clock64();                            \
  } while (0)

template <int W>
struct LastTcCfg {
  static_assert(W == 128 || W == 256 || W == 512, "tensor-core last layer: hidden 128, 256 or 512");
  static constexpr int NCH = W / 64;
  // hidden 512: one 128 KiB activation tile at a time, and dA goes through TMEM in two N = 256 halves
  static constexpr int STAGES = W <= 256 ? 3 : 1;
  static constexpr int DA_N = W <= 256 ? W : 256;   // columns of one dA MMA
  static constexpr int DA_HALVES = W / DA_N;
  static constexpr int CH_PER_HALF = DA_N / 64;
  static constexpr uint32_t STAGE_BYTES = NCH * kChunkBytes;       // act tile
  static constexpr uint32_t OFF_ACT = 0;
  static constexpr uint32_t OFF_WL = STAGES * STAGE_BYTES;          // [16 x W] fp16, NCH k-blocks of 2 KiB
  static constexpr uint32_t OFF_WLT = OFF_WL + NCH * 2048;          // [W x 16] fp16, core-matrix order
  static constexpr uint32_t WLT_BYTES = W * 32;
  static constexpr uint32_t OFF_G = OFF_WLT + WLT_BYTES;            // seed tile 128 x 16 fp16
  static constexpr uint32_t OFF_RED = OFF_G + 4096;                 // block reduction scratch
  static constexpr uint32_t OFF_BAR = OFF_RED + 8 * 8 * 4;
  static constexpr int NUM_BARS = 2 * STAGES + 6;
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
  static constexpr uint32_t TM_Y = 0, TM_DW = 32, TM_DA = 32 + (W / 128) * 16;  // TMEM columns
  static constexpr uint32_t TMEM_COLS = 512;
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of shared memory");
};

template <int W>
__global__ void __launch_bounds__(384, 1)
last_layer_tc_kernel(const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmDz,
                     const __grid_constant__ CUtensorMap tmWl, const __half* __restrict__ wlt_interleaved,
                     const LastTcArgs args) {
  using C = LastTcCfg<W>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* act_full = bars;                   // [STAGES]
  uint64_t* act_empty = act_full + C::STAGES;  // [STAGES]
  uint64_t* w_full = act_empty + C::STAGES;
  uint64_t* y_full = w_full + 1;   // y accumulator complete
  uint64_t* g_ready = y_full + 1;  // seed tile written (4 warps)
  uint64_t* mma_done = g_ready + 1;  // dA and dW MMAs of the tile retired
  uint64_t* fin_done = mma_done + 1;
  uint64_t* da_free = fin_done + 1;  // hidden 512: the first dA half has been read out of TMEM (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool train = args.mode != 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&act_full[i], 1);
      mbar_init(&act_empty[i], 1);
    }
    mbar_init(w_full, 1);
    mbar_init(y_full, 1);
    mbar_init(g_ready, 4);
    mbar_init(mma_done, 1);
    mbar_init(fin_done, 1);
    mbar_init(da_free, 8);
    fence_barrier_init();
    tma_prefetch_desc(&tmAct);
    tma_prefetch_desc(&tmDz);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything below reads what the preceding kernels of the step produced
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, C::NCH * 2048 + (train ? C::WLT_BYTES : 0));
      for (int kb = 0; kb < C::NCH; ++kb)
        tma_load_2d(smem + C::OFF_WL + kb * 2048, &tmWl, w_full, kb * 64, 0);
      // omega W_last^T arrives already in the no-swizzle core-matrix order (prep kernel), 1-D bulk copy
      if (train) bulk_load_1d(smem + C::OFF_WLT, wlt_interleaved, C::WLT_BYTES, w_full);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1u;
        mbar_wait(&act_empty[s], ph ^ 1u);
        mbar_expect_tx(&act_full[s], C::STAGE_BYTES);
        for (int c = 0; c < C::NCH; ++c)
          tma_load_2d(smem + C::OFF_ACT + s * C::STAGE_BYTES + c * kChunkBytes, &tmAct, &act_full[s],
                      c * 64, args.act_row0 + t * kRowsPerTile);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_y = umma_idesc(128, 16, 0, 0, 0, 0);
      const uint32_t id_da = umma_idesc(128, C::DA_N, 0, 0, 0, 0);
      const uint32_t id_dw = umma_idesc(128, 16, 0, 0, 1, 1);
      const uint32_t wl_addr = smem_u32(smem + C::OFF_WL);
      const uint32_t wlt_addr = smem_u32(smem + C::OFF_WLT);
      const uint32_t g_addr = smem_u32(smem + C::OFF_G);
      mbar_wait(w_full, 0);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1u;
        const uint32_t a_addr = smem_u32(smem + C::OFF_ACT + s * C::STAGE_BYTES);
        mbar_wait(&act_full[s], ph);
        SB_DBG_L(it, 6);
        tc_fence_after();
        // y = act . W_last^T  (the epilogue of the previous tile has drained TM_Y: it waited mma_done,
        // which was committed after that tile's y was consumed)
#pragma unroll
        for (int kb = 0; kb < C::NCH; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + kb * kChunkBytes + k * 32, 0, 1024, 2);
            const uint64_t db = umma_smem_desc(wl_addr + kb * 2048 + k * 32, 0, 1024, 2);
            umma_f16(tmem_base + C::TM_Y, da, db, id_y, (kb | k) != 0 ? 1u : 0u);
          }
        umma_commit(y_full);
        SB_DBG_L(it, 7);
        // the epilogue has read y (and, when training, written the seed tile)
        mbar_wait(g_ready, it & 1u);
        SB_DBG_L(it, 8);
        tc_fence_after();
        if (train) {
          // dA = g (K-major, no swizzle: LBO 128 between the two 8-wide K halves, SBO 256 between
          // 8-row groups) . omega W_last^T (same core-matrix layout, 256 rows x K 16)
          umma_f16(tmem_base + C::TM_DA, umma_smem_desc(g_addr, 128, 256, 0),
                   umma_smem_desc(wlt_addr, 128, 256, 0), id_da, 0u);
          // dW_last^T[feature, channel] += act^T . g : A = act tile MN-major, B = g MN-major
          // (same bytes: LBO 256 between 8-pixel groups, SBO 128 between the two 8-channel halves)
#pragma unroll
          for (int mb = 0; mb < W / 128; ++mb)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t da =
                  umma_smem_desc(a_addr + 2 * mb * kChunkBytes + k * 2048, kChunkBytes, 1024, 2);
              const uint64_t db = umma_smem_desc(g_addr + k * 512, 256, 128, 0);
              umma_f16(tmem_base + C::TM_DW + mb * 16, da, db, id_dw, (it | uint32_t(k)) != 0 ? 1u : 0u);
            }
          umma_commit(mma_done);
          SB_DBG_L(it, 9);
          if (C::DA_HALVES == 2) {
            // second half of dA into the same TMEM region once the epilogue has read the first one out
            mbar_wait(da_free, it & 1u);
            tc_fence_after();
            umma_f16(tmem_base + C::TM_DA, umma_smem_desc(g_addr, 128, 256, 0),
                     umma_smem_desc(wlt_addr + C::DA_N * 32, 128, 256, 0), id_da, 0u);
            umma_commit(mma_done);
          }
        }
      }
      umma_commit(fin_done);
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    const int q = warp & 3;
    const int hb = (warp - 4) >> 2;
    const int r_in_tile = q * 32 + lane;
    const bool issuer = (threadIdx.x == 128);
    const float G = train ? *args.gscale : 1.f;
    float bias[kMaxOutTc];
#pragma unroll
    for (int c = 0; c < kMaxOutTc; ++c) bias[c] = c < args.C ? args.b[c] : 0.f;
    float sse = 0.f, dbs[kMaxOutTc] = {};
    // the target (or upstream gradient) of this thread's row is fetched one tile ahead
    float tgt_next[kMaxOutTc] = {};
    auto fetch_target = [&](int tile) {
      const int64_t pn = int64_t(tile) * kRowsPerTile + r_in_tile;
#pragma unroll
      for (int c = 0; c < kMaxOutTc; ++c)
        tgt_next[c] = (train && hb == 0 && tile < args.num_tiles && pn < args.npix && c < args.C)
                          ? args.img[pn * args.C + c]
                          : 0.f;
    };
    fetch_target(blockIdx.x);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
      const uint32_t s = it % C::STAGES, ph = (it / C::STAGES) & 1u;
      const int64_t p = int64_t(t) * kRowsPerTile + r_in_tile;
      const bool row_valid = p < args.npix;
      float tgt[kMaxOutTc];
#pragma unroll
      for (int c = 0; c < kMaxOutTc; ++c) tgt[c] = tgt_next[c];
      fetch_target(t + gridDim.x);
      const bool dbgw = (threadIdx.x == 128);
      if (dbgw) SB_DBG_L(it, 0);
      mbar_wait(y_full, it & 1u);
      if (dbgw) SB_DBG_L(it, 1);
      tc_fence_after();
      if (hb == 0) {
        uint32_t yv[8];
        tmem_ld_32x8(tmem_base + (uint32_t(q * 32) << 16) + C::TM_Y, yv);
        tmem_ld_wait();
        float g[kMaxOutTc];
#pragma unroll
        for (int c = 0; c < kMaxOutTc; ++c) {
          g[c] = 0.f;
          if (c < args.C && row_valid) {
            const float z = __uint_as_float(yv[c]) + bias[c];
            const float o = args.outermost_linear ? z : sinf(z * args.omega_last);
            const float pr = o / 2 + 0.5f;
            if (args.pred) args.pred[p * args.C + c] = pr;
            if (args.mode == 1) {
              const float d = pr - tgt[c];
              sse += d * d;
              g[c] = d * G;
            } else if (args.mode == 2) {
              g[c] = 0.5f * tgt[c] * G;
            }
            if (!args.outermost_linear) g[c] *= args.omega_last * cosf(z * args.omega_last);
            dbs[c] += g[c];
          }
        }
        if (train) {
          // seed row -> core-matrix layout: 8 channels (16 B) at (r%8)*16 + (r/8)*256, next 8 at +128
          const uint32_t ga = smem_u32(smem + C::OFF_G) + (r_in_tile & 7) * 16 + (r_in_tile >> 3) * 256;
          st_shared_v4(ga, pack_f16x2(g[0], g[1]), pack_f16x2(g[2], g[3]), 0u, 0u);
          st_shared_v4(ga + 128, 0u, 0u, 0u, 0u);
          fence_proxy_async_smem();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(g_ready);
      }
      if (!train) {
        // forward only: the stage can be reused as soon as y has been read by every warp
        named_bar_sync(1, 256);
        if (issuer) mbar_arrive(&act_empty[s]);
        continue;
      }
      if (dbgw) SB_DBG_L(it, 2);
      mbar_wait(&act_full[s], ph);  // the act tile read below was written by TMA
      mbar_wait(mma_done, (it * C::DA_HALVES) & 1u);
      if (dbgw) SB_DBG_L(it, 3);
      tc_fence_after();
      const uint32_t tile_addr = smem_u32(smem + C::OFF_ACT + s * C::STAGE_BYTES);
#pragma unroll 1
      for (int nb = 0; nb < C::NCH; ++nb) {
        if (C::DA_HALVES == 2 && nb == C::CH_PER_HALF) {
          mbar_wait(mma_done, (it * C::DA_HALVES + 1) & 1u);
          tc_fence_after();
        }
        const uint32_t row_addr = tile_addr + nb * kChunkBytes + r_in_tile * 128;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + C::TM_DA + (nb % C::CH_PER_HALF) * 64 + hb * 32, v);
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
        tmem_ld_wait();
        if (C::DA_HALVES == 2 && nb == C::CH_PER_HALF - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(da_free);
        }
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t g = dz_pair_from_signed_half<false>(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]),
                                                             e[j]);
          o[j] = row_valid ? g : 0u;
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t chunk = uint32_t(hb * 4 + c4) ^ uint32_t(r_in_tile & 7);
          st_shared_v4(row_addr + (chunk << 4), o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
        }
      }
      if (dbgw) SB_DBG_L(it, 4);
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (dbgw) SB_DBG_L(it, 5);
      if (issuer) {
        for (int nb = 0; nb < C::NCH; ++nb)
          tma_store_2d(&tmDz, smem + C::OFF_ACT + s * C::STAGE_BYTES + nb * kChunkBytes, nb * 64,
                       args.dz_row0 + t * kRowsPerTile);
        tma_store_commit();
        if (C::STAGES == 1) {
          // single stage (hidden 512): hand it back as soon as this tile's own stores have been read
          tma_store_wait_read<0>();
          mbar_arrive(&act_empty[0]);
        } else if (it > 0) {
          // the previous tile's stores have been read out of their stage: hand it back to the producer
          tma_store_wait_read<1>();
          mbar_arrive(&act_empty[(it - 1) % C::STAGES]);
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
    // ---- per-CTA partials: db_last / squared error (block reduction), dW_last from TMEM ----
    float* out = args.part + int64_t(blockIdx.x) * (args.C * W + args.C + 1);
    float* red = reinterpret_cast<float*>(smem + C::OFF_RED);  // [8 warps][8]
    if (train) {
      float vals[kMaxOutTc + 1];
#pragma unroll
      for (int c = 0; c < kMaxOutTc; ++c) vals[c] = dbs[c];
      vals[kMaxOutTc] = sse;
#pragma unroll
      for (int k = 0; k <= kMaxOutTc; ++k) {
        float x = vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[(warp - 4) * 8 + k] = x;
      }
      named_bar_sync(1, 256);
      if (threadIdx.x - 128 <= unsigned(args.C)) {
        const int k = int(threadIdx.x) - 128;
        const int idx = k < args.C ? k : kMaxOutTc;
        float x = 0.f;
        for (int w8 = 0; w8 < 8; ++w8) x += red[w8 * 8 + idx];
        out[args.C * W + k] = x;  // db_last[0..C-1], then the squared error
      }
      if (hb == 0) {
        const bool any = int(blockIdx.x) < args.num_tiles;
        if (any) {
          mbar_wait(fin_done, 0);
          tc_fence_after();
        }
#pragma unroll
        for (int mb = 0; mb < W / 128; ++mb) {
          uint32_t dv[8] = {};
          if (any) {
            tmem_ld_32x8(tmem_base + (uint32_t(q * 32) << 16) + C::TM_DW + mb * 16, dv);
            tmem_ld_wait();
          }
          for (int c = 0; c < args.C; ++c) out[c * W + mb * 128 + r_in_tile] = __uint_as_float(dv[c]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// tail kernel (training step, hidden 128 / 256, at least two hidden GEMM layers): the LAST hidden layer's
// forward GEMM fused with everything last_layer_tc_kernel does, so that the last hidden activation is never
// written to (or re-read from) HBM.  Per 128-pixel tile:
//   z  = act[D-3] . W^T            UMMA M=128, N=W, K=W (A and W k-blocks streamed through a 2-stage ring)
//   a  = sin(omega (z + b))        epilogue 1 -> signed-half tile T in shared memory (never stored)
//   y  = a . W_last^T              UMMA N=16 -> pred, squared error, seed g           (epilogue 2)
//   dW_last^T += a^T . g           UMMA M=128 features, N=16, K=128 pixels (T read MN-major)
//   dA = g . (omega W_last)        UMMA K=16, two N=W/2 halves through one TMEM region
//   dz[D-2] = dA (*) +-sqrt(1-a^2) epilogue 3, in place over T, TMA-stored chunk by chunk
// One thread issues every MMA from a small event loop, so the next tile's GEMM k-blocks are issued
// between the short dependent steps of the current tile and overlap its epilogues.
// (reference: siren.py:56-68,110-118,131; train_helper.py:151-161 mse + backward)
// ------------------------------------------------------------------------------------------
struct TailArgs {
  int num_tiles;
  int a_row0;            // first row of act[D-3] (GEMM input) inside the activation tensor map
  int dz_row0;           // first row of dz[D-2] inside the dz tensor map
  int64_t npix;
  float omega;           // omega of the last hidden layer
  const float* bias;     // [W] bias of the last hidden layer
  const float* b_last;   // [C]
  const float* img;      // target [npix, C]
  float* pred;           // optional
  float* part;           // per-CTA partials [grid][C*W + C + 1]
  const float* gscale;
  int C;
  int outermost_linear;
  float omega_last;
  long long* dbg;        // optional timeline (block 0): dbg[tile * 16 + k], first 12 tiles
  int reverse;           // 1: sweep the tiles back to front (see RowGemmArgs::reverse)
  int l2_hints;          // 1: the activation k-blocks are loaded with an L2 evict_first hint (see RowGemmArgs::l2_hints)
};

#define SB_DBG_T(tile_i, k)                                                 \
  do {                                                                      \
    if (args.dbg && blockIdx.x == 0 && (tile_i) < 12)                       \
      args.dbg[(tile_i) * 16 + (k)] = clock64();                            \
  } while (0)

template <int W>
struct TailCfg {
  static_assert(W == 128 || W == 256, "tail kernel: hidden 128 or 256");
  static constexpr int NCH = W / 64;   // 64-wide chunks of the tile == k-blocks of the GEMM
  static constexpr int CPH = NCH / 2;  // chunks per half of T (dA and the dz stores go half by half)
  // Two rings: the activation k-blocks come from HBM (one tile ahead = NCH stages hides that latency), the weight
  // k-blocks from L2 (two stages are enough).  With a shared two-stage ring of {A, W} pairs the GEMM of a tile took
  // ~6 k cycles - one HBM latency per two k-blocks - and did not fit behind the previous tile's epilogues.
  static constexpr int SA = NCH;       // activation ring stages (16 KiB each)
  static constexpr int SW = 2;         // weight ring stages (W * 128 bytes each)
  static constexpr uint32_t B_KB_BYTES = W * 128;
  static constexpr uint32_t OFF_T = 0;
  static constexpr uint32_t OFF_A = NCH * kChunkBytes;
  static constexpr uint32_t OFF_W = OFF_A + SA * kChunkBytes;
  static constexpr uint32_t OFF_WL = OFF_W + SW * B_KB_BYTES;
  static constexpr uint32_t OFF_WLT = OFF_WL + NCH * 2048;
  static constexpr uint32_t WLT_BYTES = W * 32;
  static constexpr uint32_t OFF_G = OFF_WLT + WLT_BYTES;
  static constexpr uint32_t OFF_CONST = OFF_G + 4096;  // omega * bias [W] fp32
  static constexpr uint32_t OFF_RED = OFF_CONST + W * 4;
  static constexpr uint32_t OFF_BAR = OFF_RED + 16 * 8 * 4;
  static constexpr int NUM_BARS = 2 * SA + 2 * SW + 8 + 4 + 4;
  static constexpr uint32_t SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
  static constexpr uint32_t TM_ACC = 0, TM_DA = W, TM_Y = W + W / 2, TM_DW = W + W / 2 + 16;
  static constexpr uint32_t TMEM_COLS = 512;
  static_assert(TM_DW + (W / 128) * 16 <= 512, "TMEM columns");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KiB of shared memory");
};

template <int N>
__device__ __forceinline__ void tma_store_wait_read_dyn(int n) {
  // cp.async.bulk.wait_group.read takes an immediate: n in [0, N]
  if constexpr (N > 0) {
    if (n == N) {
      tma_store_wait_read<N>();
      return;
    }
    tma_store_wait_read_dyn<N - 1>(n);
  } else {
    tma_store_wait_read<0>();
  }
}

template <int W>
__global__ void __launch_bounds__(640, 1)
tail_tc_kernel(const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmDz, const __grid_constant__ CUtensorMap tmWl,
               const __half* __restrict__ wlt_interleaved, const TailArgs args) {
  using C = TailCfg<W>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* a_full = bars;                 // [SA]
  uint64_t* a_empty = a_full + C::SA;      // [SA]
  uint64_t* wk_full = a_empty + C::SA;     // [SW]
  uint64_t* wk_empty = wk_full + C::SW;    // [SW]
  uint64_t* w_full = wk_empty + C::SW;
  uint64_t* acc_full = w_full + 1;   // GEMM of the tile retired
  uint64_t* acc_free = acc_full + 1; // epilogue 1 has drained the accumulator (8 warps)
  uint64_t* t_ready = acc_free + 1;  // [4, two used] half of T written and fenced (16 warps each)
  uint64_t* y_full = t_ready + 4;
  uint64_t* g_ready = y_full + 1;    // seed tile written (4 warps)
  uint64_t* da_full = g_ready + 1;   // two phases per tile: dA half 0 (+ dW_last), dA half 1
  uint64_t* da_free = da_full + 1;   // half 0 has been read out of TMEM (8 warps)
  uint64_t* fin_done = da_free + 1;
  uint64_t* dz_done = fin_done + 1;  // [2] half of dz written over T and fenced (16 warps each)
  uint64_t* t_free = dz_done + 2;    // [2] the dz store of that half has been read out of T (store warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + C::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::SA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < C::SW; ++i) {
      mbar_init(&wk_full[i], 1);
      mbar_init(&wk_empty[i], 1);
    }
    mbar_init(w_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_free, 16);
    for (int i = 0; i < 4; ++i) mbar_init(&t_ready[i], 16);
    mbar_init(y_full, 1);
    mbar_init(g_ready, 4);
    mbar_init(da_full, 1);
    mbar_init(da_free, 16);
    mbar_init(fin_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&dz_done[i], 16);
      mbar_init(&t_free[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmAct);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmDz);
    tma_prefetch_desc(&tmWl);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 4) {
    float* cst = reinterpret_cast<float*>(smem + C::OFF_CONST);
    for (int i = threadIdx.x - 128; i < W; i += 512) cst[i] = args.omega * args.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the small weight operands were staged at the start of the step: fetch them before the dependency wait
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(w_full, C::NCH * 2048 + C::WLT_BYTES);
    for (int kb = 0; kb < C::NCH; ++kb)
      tma_load_2d(smem + C::OFF_WL + kb * 2048, &tmWl, w_full, kb * 64, 0);
    bulk_load_1d(smem + C::OFF_WLT, wlt_interleaved, C::WLT_BYTES, w_full);
  }
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== producer: activation k-blocks (HBM) =====================
    if (lane == 0) {
      uint32_t ia = 0;
      const uint64_t pol_first = l2_policy_evict_first();
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
        for (int kb = 0; kb < C::NCH; ++kb, ++ia) {
          const uint32_t s = ia % C::SA, ph = (ia / C::SA) & 1u;
          mbar_wait(&a_empty[s], ph ^ 1u);
          mbar_expect_tx(&a_full[s], kChunkBytes);
          const int arow = args.a_row0 + (args.reverse ? args.num_tiles - 1 - t : t) * kRowsPerTile;
          if (args.l2_hints)
            tma_load_2d_hint(smem + C::OFF_A + s * kChunkBytes, &tmAct, &a_full[s], kb * 64, arow, pol_first);
          else
            tma_load_2d(smem + C::OFF_A + s * kChunkBytes, &tmAct, &a_full[s], kb * 64, arow);
        }
      }
    }
  } else if (warp == 3) {
    // ===================== producer: weight k-blocks (L2), the same four for every tile =====================
    if (lane == 0) {
      uint32_t iw = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) {
        for (int kb = 0; kb < C::NCH; ++kb, ++iw) {
          const uint32_t s = iw % C::SW, ph = (iw / C::SW) & 1u;
          mbar_wait(&wk_empty[s], ph ^ 1u);
          mbar_expect_tx(&wk_full[s], C::B_KB_BYTES);
          tma_load_2d(smem + C::OFF_W + s * C::B_KB_BYTES, &tmW, &wk_full[s], kb * 64, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: event loop over {tile chain, next GEMM k-block} =====================
    if (lane == 0) {
      const uint32_t id_gemm = umma_idesc(128, W, 0, 0, 0, 0);
      const uint32_t id_y = umma_idesc(128, 16, 0, 0, 0, 0);
      const uint32_t id_da = umma_idesc(128, W / 2, 0, 0, 0, 0);
      const uint32_t id_dw = umma_idesc(128, 16, 0, 0, 1, 1);
      const uint32_t t_addr = smem_u32(smem + C::OFF_T);
      const uint32_t wl_addr = smem_u32(smem + C::OFF_WL);
      const uint32_t wlt_addr = smem_u32(smem + C::OFF_WLT);
      const uint32_t g_addr = smem_u32(smem + C::OFF_G);
      int my_tiles = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x) ++my_tiles;
      mbar_wait(w_full, 0);
      uint32_t ig = 0, kbg = 0, ia = 0;  // GEMM: tile index (local), next k-block, ring counter
      uint32_t ic = 0, cst = 0, ykb = 0;  // chain: tile index (local), state 0 y (per chunk) / 1 dW+dA0 / 2 dA1
      uint32_t idle = 0;
      while (ic < uint32_t(my_tiles)) {
        bool did = false;
        if (cst == 0) {
          // y accumulates half by half, as soon as epilogue 1 has written each half of T
          if (mbar_try_wait(&t_ready[ykb / C::CPH], ic & 1u)) {
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < C::CPH; ++cc, ++ykb)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_smem_desc(t_addr + ykb * kChunkBytes + k * 32, 0, 1024, 2);
                const uint64_t db = umma_smem_desc(wl_addr + ykb * 2048 + k * 32, 0, 1024, 2);
                umma_f16(tmem_base + C::TM_Y, da, db, id_y, (ykb | uint32_t(k)) != 0 ? 1u : 0u);
              }
            if (ykb == uint32_t(C::NCH)) {
              umma_commit(y_full);
              ykb = 0;
              cst = 1;
            }
            did = true;
          }
        } else if (cst == 1) {
          if (mbar_try_wait(g_ready, ic & 1u)) {
            tc_fence_after();
            // dW_last^T[feature, channel] += T^T . g (T read MN-major).  Epilogue 3 overwrites T in place half
            // by half, so only the feature blocks living in the first half of T have to retire before dA
            // half 0 is handed over; the others ride with dA half 1.
            auto dw_block = [&](int mb) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint64_t da = umma_smem_desc(t_addr + 2 * mb * kChunkBytes + k * 2048, kChunkBytes, 1024, 2);
                const uint64_t db = umma_smem_desc(g_addr + k * 512, 256, 128, 0);
                umma_f16(tmem_base + C::TM_DW + mb * 16, da, db, id_dw, (ic | uint32_t(k)) != 0 ? 1u : 0u);
              }
            };
            if (W == 128) dw_block(0);            // one feature block spans both halves of T
            if (W == 256) dw_block(0);            // chunks 0,1 = half 0
            umma_f16(tmem_base + C::TM_DA, umma_smem_desc(g_addr, 128, 256, 0),
                     umma_smem_desc(wlt_addr, 128, 256, 0), id_da, 0u);
            umma_commit(da_full);
            if (W == 256) dw_block(1);            // chunks 2,3 = half 1: retires before the next commit
            cst = 2;
            did = true;
          }
        } else {
          if (mbar_try_wait(da_free, ic & 1u)) {
            tc_fence_after();
            umma_f16(tmem_base + C::TM_DA, umma_smem_desc(g_addr, 128, 256, 0),
                     umma_smem_desc(wlt_addr + W * 16, 128, 256, 0), id_da, 0u);
            umma_commit(da_full);
            cst = 0;
            ++ic;
            did = true;
          }
        }
        if (!did && ig < uint32_t(my_tiles)) {
          // next GEMM k-block: the accumulator must have been drained by epilogue 1 of the previous tile
          const uint32_t s = ia % C::SA, ph = (ia / C::SA) & 1u;
          const uint32_t sw = ia % C::SW, phw = (ia / C::SW) & 1u;
          if ((kbg > 0 || mbar_try_wait(acc_free, (ig & 1u) ^ 1u)) && mbar_try_wait(&a_full[s], ph) &&
              mbar_try_wait(&wk_full[sw], phw)) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + C::OFF_A + s * kChunkBytes);
            const uint32_t b_addr = smem_u32(smem + C::OFF_W + sw * C::B_KB_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = umma_smem_desc(a_addr + k * 32, 0, 1024, 2);
              const uint64_t db = umma_smem_desc(b_addr + k * 32, 0, 1024, 2);
              umma_f16(tmem_base + C::TM_ACC, da, db, id_gemm, (kbg | uint32_t(k)) != 0 ? 1u : 0u);
            }
            umma_commit(&a_empty[s]);
            umma_commit(&wk_empty[sw]);
            ++ia;
            if (kbg == 0) SB_DBG_T(ig, 9);
            if (++kbg == uint32_t(C::NCH)) {
              umma_commit(acc_full);
              SB_DBG_T(ig, 10);
              kbg = 0;
              ++ig;
            }
            did = true;
          }
        }
        if (did) {
          idle = 0;
        } else if (++idle > (1u << 24)) {
          printf("sirenb200: tail kernel MMA loop stalled block %d (gemm tile %u kb %u, chain tile %u state %u)\n",
                 blockIdx.x, ig, kbg, ic, cst);
          __trap();
        }
      }
      umma_commit(fin_done);
    }
  } else if (warp == 2) {
    // ===================== store warp: dz halves -> HBM, then hand the half of T back =====================
    // (a dedicated thread instead of CTA-wide named barriers: the epilogue warps never wait for each other)
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&dz_done[half], it & 1u);
          for (int nbl = 0; nbl < C::CPH; ++nbl) {
            const int nb = half * C::CPH + nbl;
            tma_store_2d(&tmDz, smem + C::OFF_T + nb * kChunkBytes, nb * 64,
                         args.dz_row0 + (args.reverse ? args.num_tiles - 1 - t : t) * kRowsPerTile);
          }
          tma_store_commit();
          SB_DBG_T(it, 6 + 2 * half);
          tma_store_wait_read<0>();
          mbar_arrive(&t_free[half]);
        }
      }
      tma_store_wait_all<0>();
    }
  } else if (warp >= 4) {
    // ===================== epilogues (16 warps: four per TMEM lane quadrant, 16 of the 64 columns of a
    // chunk each - enough warps per scheduler to hide the FFMA -> MUFU -> F2FP chains) =====================
    const int q = warp & 3;
    const int hb = (warp - 4) >> 2;  // 0..3
    const int r_in_tile = q * 32 + lane;
    const bool issuer = (threadIdx.x == 128);
    const float G = *args.gscale;
    const float* cst = reinterpret_cast<const float*>(smem + C::OFF_CONST);
    const uint32_t t_addr = smem_u32(smem + C::OFF_T);
    const uint32_t lane_tm = uint32_t(q * 32) << 16;
    float bias[kMaxOutTc];
#pragma unroll
    for (int c = 0; c < kMaxOutTc; ++c) bias[c] = c < args.C ? args.b_last[c] : 0.f;
    float sse = 0.f, dbs[kMaxOutTc] = {};
    float tgt_next[kMaxOutTc] = {};
    auto fetch_target = [&](int tile) {  // tile: position in the sweep
      const int64_t pn = int64_t(args.reverse ? args.num_tiles - 1 - tile : tile) * kRowsPerTile + r_in_tile;
#pragma unroll
      for (int c = 0; c < kMaxOutTc; ++c)
        tgt_next[c] = (hb == 0 && tile < args.num_tiles && pn < args.npix && c < args.C) ? args.img[pn * args.C + c]
                                                                                       : 0.f;
    };
    fetch_target(blockIdx.x);
    uint32_t it = 0;
    for (int t = blockIdx.x; t < args.num_tiles; t += gridDim.x, ++it) {
      const int64_t p = int64_t(args.reverse ? args.num_tiles - 1 - t : t) * kRowsPerTile + r_in_tile;
      const bool row_valid = p < args.npix;
      float tgt[kMaxOutTc];
#pragma unroll
      for (int c = 0; c < kMaxOutTc; ++c) tgt[c] = tgt_next[c];
      fetch_target(t + gridDim.x);
      // ---- epilogue 1: accumulator -> sin -> T.  The TMEM load of the next chunk is in flight while this
      // one is computed; T is handed to the y MMA half by half; a half of T may only be overwritten once
      // the previous tile's dz store out of it has been read (one bulk group per half).
      if (issuer) SB_DBG_T(it, 0);
      mbar_wait(acc_full, it & 1u);
      if (issuer) SB_DBG_T(it, 1);
      tc_fence_after();
      {
        uint32_t v[2][16];
        tmem_ld_32x16(tmem_base + lane_tm + C::TM_ACC + hb * 16, v[0]);
#pragma unroll
        for (int nb = 0; nb < C::NCH; ++nb) {
          if (it > 0 && nb % C::CPH == 0) mbar_wait(&t_free[nb / C::CPH], (it - 1) & 1u);
          tmem_ld_wait();
          if (nb + 1 < C::NCH)
            tmem_ld_32x16(tmem_base + lane_tm + C::TM_ACC + (nb + 1) * 64 + hb * 16, v[(nb + 1) & 1]);
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = nb * 64 + hb * 16 + 2 * j;
            const float t0 = fmaf(__uint_as_float(v[nb & 1][2 * j]), args.omega, cst[col]);
            const float t1 = fmaf(__uint_as_float(v[nb & 1][2 * j + 1]), args.omega, cst[col + 1]);
            o[j] = sine_signed_half2(t0, t1);
          }
          const uint32_t row_addr = t_addr + nb * kChunkBytes + r_in_tile * 128;
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const uint32_t chunk = uint32_t(hb * 2 + c2) ^ uint32_t(r_in_tile & 7);
            st_shared_v4(row_addr + (chunk << 4), o[4 * c2], o[4 * c2 + 1], o[4 * c2 + 2], o[4 * c2 + 3]);
          }
          if (nb % C::CPH == C::CPH - 1) {
            fence_proxy_async_smem();
            if (nb == C::NCH - 1) tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (nb == C::NCH - 1) mbar_arrive(acc_free);
              mbar_arrive(&t_ready[nb / C::CPH]);
            }
          }
        }
      }
      if (issuer) SB_DBG_T(it, 2);
      // ---- epilogue 2: y -> pred, squared error, seed ----
      if (hb == 0) {
        mbar_wait(y_full, it & 1u);
        if (issuer) SB_DBG_T(it, 3);
        tc_fence_after();
        uint32_t yv[8];
        tmem_ld_32x8(tmem_base + lane_tm + C::TM_Y, yv);
        tmem_ld_wait();
        float g[kMaxOutTc];
#pragma unroll
        for (int c = 0; c < kMaxOutTc; ++c) {
          g[c] = 0.f;
          if (c < args.C && row_valid) {
            const float z = __uint_as_float(yv[c]) + bias[c];
            const float o = args.outermost_linear ? z : sinf(z * args.omega_last);
            const float pr = o / 2 + 0.5f;
            if (args.pred) args.pred[p * args.C + c] = pr;
            const float d = pr - tgt[c];
            sse += d * d;
            g[c] = d * G;
            if (!args.outermost_linear) g[c] *= args.omega_last * cosf(z * args.omega_last);
            dbs[c] += g[c];
          }
        }
        const uint32_t ga = smem_u32(smem + C::OFF_G) + (r_in_tile & 7) * 16 + (r_in_tile >> 3) * 256;
        st_shared_v4(ga, pack_f16x2(g[0], g[1]), pack_f16x2(g[2], g[3]), 0u, 0u);
        st_shared_v4(ga + 128, 0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(g_ready);
        if (issuer) SB_DBG_T(it, 4);
      }
      // ---- epilogue 3: dz = dA (*) cos, in place over T, stored half by half ----
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        mbar_wait(da_full, uint32_t(half));
        if (issuer) SB_DBG_T(it, 5 + 2 * half);
        tc_fence_after();
        uint32_t v[2][16];
        tmem_ld_32x16(tmem_base + lane_tm + C::TM_DA + hb * 16, v[0]);
#pragma unroll
        for (int nbl = 0; nbl < C::CPH; ++nbl) {
          const int nb = half * C::CPH + nbl;
          const uint32_t row_addr = t_addr + nb * kChunkBytes + r_in_tile * 128;
          uint32_t e[8];
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const uint32_t chunk = uint32_t(hb * 2 + c2) ^ uint32_t(r_in_tile & 7);
            const uint4 ld = ld_shared_v4(row_addr + (chunk << 4));
            e[4 * c2 + 0] = ld.x;
            e[4 * c2 + 1] = ld.y;
            e[4 * c2 + 2] = ld.z;
            e[4 * c2 + 3] = ld.w;
          }
          tmem_ld_wait();
          if (nbl + 1 < C::CPH)
            tmem_ld_32x16(tmem_base + lane_tm + C::TM_DA + (nbl + 1) * 64 + hb * 16, v[(nbl + 1) & 1]);
          if (half == 0 && nbl == C::CPH - 1) {
            // half 0 is out of TMEM: the second half of dA may overwrite the region while we do the math
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(da_free);
          }
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] = dz_pair_from_signed_half<false>(__uint_as_float(v[nbl & 1][2 * j]),
                                                   __uint_as_float(v[nbl & 1][2 * j + 1]), e[j]);
          }
          if (!row_valid) {  // padding rows of the last tile (their seed is zero, but keep dz exactly zero)
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0u;
          }
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const uint32_t chunk = uint32_t(hb * 2 + c2) ^ uint32_t(r_in_tile & 7);
            st_shared_v4(row_addr + (chunk << 4), o[4 * c2], o[4 * c2 + 1], o[4 * c2 + 2], o[4 * c2 + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dz_done[half]);
      }
      tc_fence_before();
    }
    // ---- per-CTA partials: db_last / squared error (block reduction), dW_last from TMEM ----
    float* out = args.part + int64_t(blockIdx.x) * (args.C * W + args.C + 1);
    float* red = reinterpret_cast<float*>(smem + C::OFF_RED);  // [16 warps][8]
    float vals[kMaxOutTc + 1];
#pragma unroll
    for (int c = 0; c < kMaxOutTc; ++c) vals[c] = dbs[c];
    vals[kMaxOutTc] = sse;
#pragma unroll
    for (int k = 0; k <= kMaxOutTc; ++k) {
      float x = vals[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) red[(warp - 4) * 8 + k] = x;
    }
    named_bar_sync(1, 512);
    if (threadIdx.x - 128 <= unsigned(args.C)) {
      const int k = int(threadIdx.x) - 128;
      const int idx = k < args.C ? k : kMaxOutTc;
      float x = 0.f;
      for (int w8 = 0; w8 < 16; ++w8) x += red[w8 * 8 + idx];
      out[args.C * W + k] = x;  // db_last[0..C-1], then the squared error
    }
    if (hb == 0) {
      const bool any = int(blockIdx.x) < args.num_tiles;
      if (any) {
        mbar_wait(fin_done, 0);
        tc_fence_after();
      }
#pragma unroll
      for (int mb = 0; mb < W / 128; ++mb) {
        uint32_t dv[8] = {};
        if (any) {
          tmem_ld_32x8(tmem_base + lane_tm + C::TM_DW + mb * 16, dv);
          tmem_ld_wait();
        }
        for (int c = 0; c < args.C; ++c) out[c * W + mb * 128 + r_in_tile] = __uint_as_float(dv[c]);
      }
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
