// libsirenb200.so — C ABI (include/siren_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>
#include <utility>
#include <vector>

#include "../../include/siren_b200.h"
#include "simt_kernels.cuh"
#include "tc_kernels.cuh"
#include "tmap.h"

using namespace sb;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fail(SIRENB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                     \
  } while (0)

#define LAUNCH_CHECK()                                                                     \
  do {                                                                                     \
    ++g_launches;                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess)                                                                \
      return fail(SIRENB200_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                  \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                            \
  } while (0)

inline int cdiv(int64_t a, int64_t b) { return int((a + b - 1) / b); }

}  // namespace

constexpr int kTimelineSlots = 3 * 4 * 8 * 16 + 12 * 16;
constexpr int kStallLaunches = 12;  // slots: 0..5 forward layer l, 6..11 backward layer l

struct sirenb200_plan {
  sirenb200_config_t cfg;
  int device = 0;
  int nsm = 0;
  int D = 0, W = 0, C = 0;  // W = kernel width (tensor-core path: the model width padded to 128 / 256 / 512)
  int Wm = 0;               // the model's hidden width
  int rows = 0;
  int64_t npix = 0, npix_pad = 0;
  int ntiles = 0;
  float inv_count = 0.f;  // 1 / (H*W*C)
  CoordSrc coord{};
  bool have_fwd = false;
  int64_t bytes = 0;

  // device scalars / small buffers
  float* gstate = nullptr;      // [0] seed scale G, [1] cap
  float* loss_part = nullptr;   // loss partials
  double* eval_acc = nullptr;

  // ---- fp32 path ----
  float* x32 = nullptr;        // [npix, 2]
  float* z32 = nullptr;        // [(D-1)][npix, W] pre-activations
  float* a32 = nullptr;        // [(D-1)][npix, W] activations
  float* y32 = nullptr;        // [npix, C]
  float* g32 = nullptr;        // [npix, C]
  float* dz32[2] = {nullptr, nullptr};  // [npix, W] ping-pong
  float* part32 = nullptr;     // split-K partials for the largest tensor (+ bias sums)
  int simt_splits = 1;
  int simt_split_len = 0;

  // ---- tensor-core path ----
  __half* act = nullptr;  // [(D-1)][npix_pad, W] signed-half activations
  __half* dz = nullptr;   // [(D-1)][npix_pad, W] fp16 gradients (seed units)
  __half* wh = nullptr;   // [(D-2)][W, W]
  __half* wth = nullptr;  // [(D-2)][W, W]
  float4* tab0 = nullptr; // [W] layer-0 epilogue table (fused forward)
  float* bias_w = nullptr;  // [(D-2)][W] omega * bias
  float* bias_raw = nullptr;  // [(D-2)][W] hidden-layer biases, zero-padded to the kernel width
  float* w0p = nullptr;       // [W, 2] layer-0 weight, zero-padded
  float* b0p = nullptr;       // [W]    layer-0 bias, zero-padded
  CUtensorMap tm_act{}, tm_dz{};
  CUtensorMap tm_act_c{}, tm_dz_c{};  // chunked 3-D views (one box = a whole reduction operand of a 128-pixel stage)
  CUtensorMap tm_act_c2{};            // ... of one CTA of a pair (two activation chunks)
  bool bwd_pair = true;               // merged backward launch: reduction role on CTA pairs (SIRENB200_PAIR=0: off)
  __half* wl16 = nullptr;   // [16, W] last-layer weights for the tensor-core last layer
  __half* wlt16 = nullptr;  // [W, 64]
  CUtensorMap tm_wl{}, tm_wlt{};
  bool last_tc = false;
  bool tail_fused = true;   // last hidden GEMM + output layer + loss + dZ in one kernel (SIRENB200_TAIL=0: two kernels)
  bool pdl = true;          // programmatic dependent launch of the GEMM kernels (SIRENB200_PDL=0: off)
  int alt_sweep = 7;        // consecutive launches sweep the tiles in alternating directions, so each starts with what
                            // its predecessor wrote last (still in L2); SIRENB200_ALT_SWEEP bits: 1 forward GEMMs,
                            // 2 tail kernel, 4 merged backward launches
  int l2_hints = 0;         // forward GEMMs / tail kernel load their input with an L2 evict_first hint (SIRENB200_L2_HINTS)
  bool tail_ran = false;    // the last forward pass ended with the tail kernel ...
  int tail_rev = 0;         // ... sweeping in this direction (1 = back to front)
  bool fuse_l0 = true;      // layer-0 gradient reduced inside the dX GEMM of the first hidden layer (SIRENB200_FUSE_L0=0: own kernel)
  int l0_used = 0;          // partial rows of l0_part written by the last backward
  int last_rowgemm_grid = 0;
  bool gen_first = true;    // layer 0 generated inside the first hidden GEMM (SIRENB200_GEN_FIRST=0: own kernel)
  long long* dbg_stall = nullptr;     // SIRENB200_STALLS=1: wait-cycle counters, [launch slot][160][16] (debug)
  long long* dbg_timeline = nullptr;  // SIRENB200_TIMELINE=1: 3*4*8*16 clock64 slots (debug)
  std::vector<CUtensorMap> tm_w, tm_wt;
  float* dw_part = nullptr;  // [splits][D-2][W][W]
  float* db_part = nullptr;  // [splits][D-2][W]
  int col_splits = 1;
  float* last_part = nullptr;  // [last_grid][C*W + C + 1]
  int last_grid = 0;
  float* l0_part = nullptr;    // [nchunks][l0_grid][3*W]
  int l0_grid = 0;
  int chunk_tiles = 0;         // 128-row tiles per L2-resident chunk (0 = whole shard)
  int nchunks = 1;
  int active_splits = 1;       // split slabs the dW GEMM actually writes (<= col_splits)
  unsigned int* pace = nullptr;  // [2 * kMaxLayers] pace counters (merged dX + dW launches)
  bool bwd_merged = true;      // dX GEMM and weight-gradient reduction of a layer in ONE launch (SIRENB200_BWD_MERGED=0: separate)
  int dw_ctas = 0;             // CTAs of that launch that run the reduction
  int merged_splits = 0;       // pixel splits of the reduction role
  int merged_splits_l0 = 0;    // ... in the first hidden layer's launch (its dX role also reduces layer 0's gradient and
                               // is the slower role there, so it gets more of the SMs); 0: same as merged_splits
  int l1_splits = 0;           // split slabs the first hidden layer's reduction wrote in the last backward
  int pace_window = 192;       // tiles the reduction role may run ahead of the dX role
  int pace_window_dx = 192;    // ... the dX role ahead of the reduction role (negative: it follows that far behind)
  int stall_slot = 0;
  // QAT activation fake-quant (fp32 handles): observer state is caller-owned, masks / partials live here
  float* actq_state = nullptr;          // [D][4] {running min, running max, scale, zero point} (device, caller's)
  bool actq_on = false;
  int actq_training = 0, actq_qmin = 0, actq_qmax = 127;
  float actq_avg = 0.01f;
  unsigned char* actq_mask = nullptr;   // [(D-1) * npix * W + npix * C] straight-through masks
  float* actq_partial = nullptr;        // [2 * kActqBlocks]
  // model family on the fp32 path: 0 = Siren, 1 = FourierNet (models/fourier.py: Fourier-feature encoding, ReLU, sigmoid)
  int model_kind = 0;
  int nlin = 0;                // linear layers (Siren: depth; FourierNet: depth - 1)
  int k0 = 2;                  // input features of the first linear layer (FourierNet: map_size)
  const float* encB = nullptr; // FourierNet: encoding.B [2, map_size / 2] (device, caller's)
  float* enc32 = nullptr;      // FourierNet: [npix, map_size] encoded coordinates
  bool defer_reduce = false;   // transient: tc_run leaves the partial reduction to the fused step-end kernel
  unsigned long long* bar = nullptr;  // grid-barrier counter of step_end_kernel

  // ---- optional per-kernel timing (cudaEvent pairs recorded around tagged launches) ----
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;   // start/stop pairs
  std::vector<int> prof_kind;         // kind of pair i
  size_t prof_used = 0;
};

enum ProfKind {
  PK_PREP = 0, PK_FIRST = 1, PK_FWD_GEMM = 2, PK_LAST = 3, PK_DX_GEMM = 4, PK_DW_GEMM = 5,
  PK_L0_GRAD = 6, PK_REDUCE = 7, PK_FINALIZE = 8, PK_COUNT = 9
};

namespace {
struct ProfScope {
  sirenb200_plan* p;
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  ProfScope(sirenb200_plan* plan, int kind, cudaStream_t s) : p(plan), st(s) {
    if (!p->prof_on) return;
    if (p->prof_used + 2 > p->prof_ev.size()) {
      for (int i = 0; i < 512; ++i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        p->prof_ev.push_back(e);
      }
    }
    cudaEvent_t start = p->prof_ev[p->prof_used];
    stop = p->prof_ev[p->prof_used + 1];
    p->prof_used += 2;
    p->prof_kind.push_back(kind);
    cudaEventRecord(start, st);
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};
}  // namespace

namespace {

template <typename T>
int dev_alloc(sirenb200_plan* p, T** ptr, int64_t count) {
  if (count <= 0) count = 1;
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(ptr), size_t(count) * sizeof(T)));
  p->bytes += count * int64_t(sizeof(T));
  return 0;
}

bool tc_supported(int W) { return W == 128 || W == 256 || W == 512; }

float omega_of(const sirenb200_plan* p, int layer) {
  return layer == 0 ? p->cfg.first_omega : p->cfg.hidden_omega;
}

int check_ready(const sirenb200_plan* p) {
  if (!p) return fail(SIRENB200_ERR_INVALID, "null handle");
  if (!p->coord.coords && !(p->coord.lin_h && p->coord.lin_w))
    return fail(SIRENB200_ERR_STATE, "no grid bound: call sirenb200_set_grid_lut/coords first");
  return 0;
}

// ---------------------------------------------------------------------------------------
// rowgemm / colgemm launch helpers
// ---------------------------------------------------------------------------------------
// Launch with (pdl) or without programmatic stream serialisation.  A kernel launched with it may begin while
// its predecessor is still draining; every such kernel calls griddepcontrol.wait before it touches anything
// the predecessor produced.
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ... as clusters of two CTAs (the two SMs of a TPC): CTA pairs for cta_group::2 MMAs
template <typename... KArgs, typename... Args>
cudaError_t launch_pairs(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <int W, int MODE, bool GEN = false, bool RED = false>
int launch_rowgemm(sirenb200_plan* p, const CUtensorMap& tmA, const CUtensorMap& tmB,
                   const CUtensorMap& tmE, const CUtensorMap& tmO, const RowGemmArgs& args,
                   cudaStream_t st) {
  constexpr int NT = W < 256 ? W : 256;  // output columns per work item
  constexpr int NPARTS = W / NT;
  using Cfg = RowGemmCfg<W, NT, MODE, NPARTS, RED>;
  auto kfn = rowgemm_kernel<W, NT, MODE, false, NPARTS, GEN, RED>;
  static bool attr_set[64] = {};
  if (!attr_set[p->device & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(Cfg::SMEM_BYTES)));
    attr_set[p->device & 63] = true;
  }
  const int items = args.num_tiles * NPARTS;
  int grid = items < p->nsm ? items : p->nsm;
  const uint32_t idesc = umma_idesc(128, NT, 0, 0, 0, 0);
  RowGemmArgs ra = args;
  ra.stall = p->dbg_stall ? p->dbg_stall + int64_t(p->stall_slot % kStallLaunches) * 160 * 16 : nullptr;
  {
    ProfScope ps(p, MODE == MODE_FWD ? PK_FWD_GEMM : PK_DX_GEMM, st);
    launch_ex(kfn, dim3(grid), dim3(rowgemm_threads(MODE, GEN, RED, Cfg::STREAM_B)), Cfg::SMEM_BYTES, st, p->pdl && !p->prof_on, tmA,
              tmB, tmE, tmO, ra, idesc);
  }
  LAUNCH_CHECK();
  p->last_rowgemm_grid = grid;
  return 0;
}

template <int W>
int launch_colgemm(sirenb200_plan* p, const ColGemmJobs& jobs, cudaStream_t st) {
  constexpr int NT = W < 256 ? W : 256;
  using Cfg = ColGemmCfg<NT>;
  auto kfn = colgemm_kernel<NT>;
  static bool attr_set[64] = {};
  if (!attr_set[p->device & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(Cfg::SMEM_BYTES)));
    attr_set[p->device & 63] = true;
  }
  const int grid = jobs.num_problems * jobs.mblocks * jobs.nparts * jobs.splits;
  {
    ProfScope ps(p, PK_DW_GEMM, st);
    launch_ex(kfn, dim3(grid), dim3(256), Cfg::SMEM_BYTES, st, p->pdl && !p->prof_on, p->tm_dz_c, p->tm_act_c, jobs,
              umma_idesc(128, NT, 0, 0, 1, 1), umma_idesc(128, 16, 0, 0, 1, 1));
  }
  LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------
// tensor-core path.  The pixel rows are processed in chunks of `chunk_tiles` 128-row tiles: for each
// chunk the whole layer chain runs forward and then backward before the next chunk starts, so the
// activations / gradients of a chunk are re-read from L2 instead of HBM (the 126 MB L2 holds one
// chunk's working set); weight-gradient partials accumulate across chunks.
// ---------------------------------------------------------------------------------------
struct Chunk {
  int index;
  int t0, ntiles;          // 128-row tiles
  int64_t p0;              // first pixel
  int64_t npix, npix_pad;  // valid / padded pixels of this chunk
};

std::vector<Chunk> make_chunks(const sirenb200_plan* p) {
  std::vector<Chunk> v;
  const int ct = p->chunk_tiles > 0 ? p->chunk_tiles : p->ntiles;
  for (int t0 = 0, i = 0; t0 < p->ntiles; t0 += ct, ++i) {
    Chunk c;
    c.index = i;
    c.t0 = t0;
    c.ntiles = (t0 + ct <= p->ntiles) ? ct : p->ntiles - t0;
    c.p0 = int64_t(t0) * kRowsPerTile;
    c.npix_pad = int64_t(c.ntiles) * kRowsPerTile;
    c.npix = (c.p0 + c.npix_pad <= p->npix) ? c.npix_pad : p->npix - c.p0;
    v.push_back(c);
  }
  return v;
}

template <int W>
int tc_prep(sirenb200_plan* p, const float* const* prm, cudaStream_t st, float* stats_to_zero) {
  const int nh = p->D - 2;  // hidden (W x W) layers
  if (nh > 0) {
    PrepArgs pa{};
    for (int l = 1; l <= nh; ++l) {
      pa.w[l - 1] = prm[2 * l];
      pa.omega_prev[l - 1] = omega_of(p, l - 1);
    }
    pa.nlayers = nh;
    pa.W = W;
    pa.wh = p->wh;
    pa.wth = p->wth;
    pa.stats = stats_to_zero;
    for (int l = 1; l <= nh; ++l) pa.bias[l - 1] = prm[2 * l + 1];
    pa.w0 = prm[0];
    pa.b0 = prm[1];
    pa.omega0 = omega_of(p, 0);
    pa.omega_h = p->cfg.hidden_omega;
    pa.tab0 = p->tab0;
    pa.bias_w = p->bias_w;
    pa.w_last = prm[2 * (p->D - 1)];
    pa.C = p->C;
    pa.omega_prev_last = omega_of(p, p->D - 2);
    pa.wl16 = p->last_tc ? p->wl16 : nullptr;
    pa.wlt16 = p->wlt16;
    pa.bias_raw = p->bias_raw;
    pa.pace = p->pace;
    pa.Wm = p->Wm;
    pa.w0p = p->w0p;
    pa.b0p = p->b0p;
    {
      ProfScope ps(p, PK_PREP, st);
      tc_prep_weights_kernel<<<dim3(W / 32, W / 32, nh), 256, 0, st>>>(pa);
    }
    LAUNCH_CHECK();
  } else if (stats_to_zero) {
    CUDA_TRY(cudaMemsetAsync(stats_to_zero, 0, 4 * sizeof(float), st));
  }
  return 0;
}

template <int W>
int tc_forward_chunk(sirenb200_plan* p, const float* const* prm, const Chunk& ch, cudaStream_t st,
                     bool skip_last_hidden = false) {
  const int nh = p->D - 2;
  // Layer 0 runs inside the first hidden layer's GEMM (its A operand is generated in the kernel) when
  // that kernel keeps its weights resident; otherwise as its own CUDA-core kernel.
  constexpr bool kCanGen = (W <= 256);
  const bool gen_first = kCanGen && p->gen_first && nh >= 1;
  CoordSrc cs = p->coord;
  cs.p_offset = ch.p0;
  if (!gen_first) {
    int grid = p->nsm * 8;
    const int need = cdiv(ch.npix_pad, 256 / (W / 8));
    if (grid > need) grid = need;
    ProfScope ps(p, PK_FIRST, st);
    tc_first_layer_kernel<W><<<grid, 256, 0, st>>>(cs, p->w0p, p->b0p, omega_of(p, 0),
                                                  p->act + ch.p0 * W, ch.npix, ch.npix_pad);
  }
  LAUNCH_CHECK();
  const int l_end = skip_last_hidden ? nh - 1 : nh;
  for (int l = 1; l <= l_end; ++l) {
    RowGemmArgs ra{};
    ra.num_tiles = ch.ntiles;
    ra.a_row0 = int((l - 1) * p->npix_pad + ch.p0);
    ra.e_row0 = 0;
    ra.o_row0 = int(l * p->npix_pad + ch.p0);
    ra.valid_rows = int(ch.npix);
    ra.omega = omega_of(p, l);
    ra.bias = p->bias_raw + size_t(l - 1) * W;  // staged (zero-padded) copy of prm[2 * l + 1]
    ra.bias_w = p->bias_w + size_t(l - 1) * W;  // omega * bias (the streamed-B kernel reads it through L1)
    // layer 1 sweeps front to back (its input is generated / freshly written front to back), layer 2 back to front, ...
    ra.reverse = ((p->alt_sweep & 1) && p->nchunks == 1 && (l % 2 == 0)) ? 1 : 0;
    ra.l2_hints = p->l2_hints;
    ra.b_early = (l >= 2) ? 1 : 0;  // layer 1 follows the weight-staging kernel directly
    p->stall_slot = l;
    int rc;
    if constexpr (kCanGen) {
      if (l == 1 && gen_first) {
        ra.gen_coord = cs;
        ra.gen_w0 = p->w0p;
        ra.gen_b0 = p->b0p;
        ra.gen_omega = omega_of(p, 0);
        ra.gen_tl = p->dbg_timeline;
        rc = launch_rowgemm<W, MODE_FWD, true>(p, p->tm_act, p->tm_w[0], p->tm_act, p->tm_act, ra, st);
        if (rc) return rc;
        continue;
      }
    }
    rc = launch_rowgemm<W, MODE_FWD>(p, p->tm_act, p->tm_w[l - 1], p->tm_act, p->tm_act, ra, st);
    if (rc) return rc;
  }
  return 0;
}

template <int W>
int launch_last_tc(sirenb200_plan* p, const float* const* prm, int mode, const float* img_or_dpred,
                   float* pred, const Chunk& ch, cudaStream_t st) {
  if constexpr (W == 128 || W == 256 || W == 512) {
    using Cfg = LastTcCfg<W>;
    auto kfn = last_layer_tc_kernel<W>;
    static bool attr_set[64] = {};
    if (!attr_set[p->device & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    int(Cfg::SMEM_BYTES)));
      attr_set[p->device & 63] = true;
    }
    const int D = p->D, C = p->C;
    LastTcArgs la{};
    la.num_tiles = ch.ntiles;
    la.act_row0 = int((D - 2) * p->npix_pad + ch.p0);
    la.dz_row0 = int((D - 2) * p->npix_pad + ch.p0);
    la.npix = ch.npix;
    la.b = prm[2 * (D - 1) + 1];
    la.img = img_or_dpred ? img_or_dpred + ch.p0 * C : nullptr;
    la.pred = pred ? pred + ch.p0 * C : nullptr;
    la.part = p->last_part + size_t(ch.index) * p->last_grid * (C * W + C + 1);
    la.gscale = p->gstate;
    la.C = C;
    la.mode = mode;
    la.outermost_linear = p->cfg.outermost_linear;
    la.omega_last = omega_of(p, D - 1);
    la.dbg = p->dbg_timeline ? p->dbg_timeline + 3 * 4 * 8 * 16 : nullptr;
    {
      ProfScope ps(p, PK_LAST, st);
      launch_ex(kfn, dim3(p->last_grid), dim3(384), Cfg::SMEM_BYTES, st, p->pdl && !p->prof_on, p->tm_act, p->tm_dz,
                p->tm_wl, static_cast<const __half*>(p->wlt16), la);
    }
    LAUNCH_CHECK();
    return 0;
  } else {
    return fail(SIRENB200_ERR_INVALID, "tensor-core last layer needs hidden 128 or 256");
  }
}

// training step: last hidden layer's GEMM + output layer + loss + dZ in one kernel (tail_tc_kernel)
template <int W>
int launch_tail(sirenb200_plan* p, const float* const* prm, const float* img, float* pred, const Chunk& ch,
                cudaStream_t st) {
  if constexpr (W == 128 || W == 256) {
    using Cfg = TailCfg<W>;
    auto kfn = tail_tc_kernel<W>;
    static bool attr_set[64] = {};
    if (!attr_set[p->device & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    int(Cfg::SMEM_BYTES)));
      attr_set[p->device & 63] = true;
    }
    const int D = p->D, C = p->C, nh = D - 2;
    TailArgs ta{};
    ta.num_tiles = ch.ntiles;
    ta.a_row0 = int((nh - 1) * p->npix_pad + ch.p0);
    ta.dz_row0 = int(nh * p->npix_pad + ch.p0);
    ta.npix = ch.npix;
    ta.omega = omega_of(p, nh);
    ta.bias = p->bias_raw + size_t(nh - 1) * W;
    ta.reverse = ((p->alt_sweep & 2) && p->nchunks == 1 && (nh % 2 == 0)) ? 1 : 0;  // layer nh: as the forward GEMMs alternate
    p->tail_rev = ta.reverse;
    ta.l2_hints = p->l2_hints;
    ta.b_last = prm[2 * (D - 1) + 1];
    ta.img = img + ch.p0 * C;
    ta.pred = pred ? pred + ch.p0 * C : nullptr;
    ta.part = p->last_part + size_t(ch.index) * p->last_grid * (C * W + C + 1);
    ta.gscale = p->gstate;
    ta.C = C;
    ta.outermost_linear = p->cfg.outermost_linear;
    ta.omega_last = omega_of(p, D - 1);
    ta.dbg = p->dbg_timeline ? p->dbg_timeline + 3 * 4 * 8 * 16 : nullptr;
    {
      ProfScope ps(p, PK_LAST, st);
      launch_ex(kfn, dim3(p->last_grid), dim3(640), Cfg::SMEM_BYTES, st, p->pdl && !p->prof_on, p->tm_act,
                p->tm_w[nh - 1], p->tm_dz, p->tm_wl, static_cast<const __half*>(p->wlt16), ta);
    }
    LAUNCH_CHECK();
    return 0;
  } else {
    return fail(SIRENB200_ERR_INVALID, "tail kernel needs hidden 128 or 256");
  }
}

template <int W>
int tc_last_chunk(sirenb200_plan* p, const float* const* prm, int mode, const float* img_or_dpred,
                  float* pred, const Chunk& ch, cudaStream_t st) {
  if (p->last_tc) return launch_last_tc<W>(p, prm, mode, img_or_dpred, pred, ch, st);
  const int D = p->D, C = p->C;
  LastArgs la{};
  la.act = p->act + (size_t(D - 2) * p->npix_pad + ch.p0) * W;
  la.dz = p->dz + (size_t(D - 2) * p->npix_pad + ch.p0) * W;
  la.w = prm[2 * (D - 1)];
  la.b = prm[2 * (D - 1) + 1];
  la.img = img_or_dpred ? img_or_dpred + ch.p0 * C : nullptr;
  la.pred = pred ? pred + ch.p0 * C : nullptr;
  la.part = p->last_part + size_t(ch.index) * p->last_grid * (C * W + C + 1);
  la.gscale = p->gstate;
  la.npix = ch.npix;
  la.npix_pad = ch.npix_pad;
  la.C = C;
  la.mode = mode;
  la.outermost_linear = p->cfg.outermost_linear;
  la.omega_last = omega_of(p, D - 1);
  la.omega_prev = omega_of(p, D - 2);
  {
    ProfScope ps(p, PK_LAST, st);
    if (C == 3)
      tc_last_layer_kernel<W, 3><<<p->last_grid, 256, 0, st>>>(la);
    else
      tc_last_layer_kernel<W, kMaxOut><<<p->last_grid, 256, 0, st>>>(la);
  }
  LAUNCH_CHECK();
  return 0;
}

// hidden 512: the weight-gradient reduction of all hidden layers on CTA pairs (2-CTA clusters), one launch
template <int W>
int launch_colgemm2(sirenb200_plan* p, const ColGemmJobs& jobs, cudaStream_t st) {
  auto kfn = colgemm2_kernel<W>;
  static bool attr_set[64] = {};
  if (!attr_set[p->device & 63]) {
    CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(ColGemm2Cfg::SMEM_BYTES)));
    attr_set[p->device & 63] = true;
  }
  const int grid = jobs.num_problems * (W / 256) * (W / 256) * jobs.splits * 2;
  {
    ProfScope ps(p, PK_DW_GEMM, st);
    ColGemmJobs cj = jobs;
    cj.stall = p->dbg_stall ? p->dbg_stall + int64_t(11 % kStallLaunches) * 160 * 16 : nullptr;  // slot 11
    CUDA_TRY(launch_pairs(kfn, dim3(grid), dim3(256), ColGemm2Cfg::SMEM_BYTES, st, p->pdl && !p->prof_on, p->tm_dz_c,
                          p->tm_act_c2, cj));
  }
  LAUNCH_CHECK();
  return 0;
}

// one backward layer: dX GEMM of layer l (CTAs [0, dx)) + its weight-gradient reduction (the rest), one launch
template <int W, bool RED>
int launch_bwd_merged(sirenb200_plan* p, const RowGemmArgs& ra, const ColGemmJobs& jobs, int l, cudaStream_t st) {
  constexpr int NT = W < 256 ? W : 256;
  constexpr int NPARTS = W / NT;
  using RCfg = RowGemmCfg<W, NT, MODE_DX, NPARTS, RED>;
  using CCfg = ColGemmCfg<NT>;
  constexpr uint32_t SMEM1 = RCfg::SMEM_BYTES > CCfg::SMEM_BYTES ? RCfg::SMEM_BYTES : CCfg::SMEM_BYTES;
  constexpr uint32_t SMEM = SMEM1 > ColGemm2Cfg::SMEM_BYTES ? SMEM1 : ColGemm2Cfg::SMEM_BYTES;
  const int dw = jobs.num_problems * jobs.mblocks * jobs.nparts * jobs.splits;
  int dx = p->nsm - dw;
  if (dx > ra.num_tiles * NPARTS) dx = ra.num_tiles * NPARTS;
  // CTA pairs for the reduction role: hidden 256 (one pair = both 128-row blocks of dW), both roles' CTA counts even
  const bool pair = W == 256 && p->bwd_pair && jobs.mblocks == 2 && jobs.nparts == 1 && dx % 2 == 0 && dw % 2 == 0 &&
                    jobs.interleave == 1;
  RowGemmArgs rb = ra;
  ColGemmJobs cj = jobs;
  rb.stall = p->dbg_stall ? p->dbg_stall + int64_t(p->stall_slot % kStallLaunches) * 160 * 16 : nullptr;
  cj.stall = rb.stall ? rb.stall + int64_t(dx) * 16 : nullptr;
  const dim3 block(rowgemm_threads(MODE_DX, false, RED));
  const bool pdl = p->pdl && !p->prof_on;
  static bool attr_set[64][2] = {};
  {
    ProfScope ps(p, PK_DX_GEMM, st);
    if constexpr (W == 256) {
      if (pair) {
        auto kfn = bwd_merged_kernel<W, RED, true>;
        if (!attr_set[p->device & 63][1]) {
          CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
          attr_set[p->device & 63][1] = true;
        }
        launch_pairs(kfn, dim3(dx + dw), block, SMEM, st, pdl, p->tm_dz, p->tm_wt[l - 1], p->tm_act, p->tm_dz_c,
                     p->tm_act_c2, rb, umma_idesc(128, NT, 0, 0, 0, 0), cj, 0u, 0u, dx, p->pace + 2 * l,
                     p->pace_window, p->pace_window_dx);
      }
    }
    if (!pair) {
      auto kfn = bwd_merged_kernel<W, RED, false>;
      if (!attr_set[p->device & 63][0]) {
        CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM)));
        attr_set[p->device & 63][0] = true;
      }
      launch_ex(kfn, dim3(dx + dw), block, SMEM, st, pdl, p->tm_dz, p->tm_wt[l - 1], p->tm_act,
                p->tm_dz_c, p->tm_act_c, rb,
                umma_idesc(128, NT, 0, 0, 0, 0), cj, umma_idesc(128, NT, 0, 0, 1, 1),
                umma_idesc(128, 16, 0, 0, 1, 1), dx, p->pace + 2 * l, p->pace_window, p->pace_window_dx);
    }
  }
  LAUNCH_CHECK();
  p->last_rowgemm_grid = dx;
  return 0;
}

template <int W>
int tc_backward_chunk(sirenb200_plan* p, const float* const* prm, const Chunk& ch, cudaStream_t st) {
  const int nh = p->D - 2;
  constexpr int NPARTS = W / (W < 256 ? W : 256);
  auto fill_jobs = [&](ColGemmJobs& jobs, int l_first, int nprob, int splits, int interleave) {
    jobs.num_problems = nprob;
    jobs.mblocks = W / 128;
    jobs.nparts = NPARTS;
    jobs.splits = splits;
    jobs.tile0 = ch.t0;
    jobs.tiles_total = ch.ntiles;
    jobs.tiles_per_split = cdiv(ch.ntiles, splits);
    jobs.accumulate = ch.index > 0 ? 1 : 0;
    for (int i = 0; i < nprob; ++i) {
      jobs.x_row0[i] = int((l_first + i) * p->npix_pad);
      jobs.y_row0[i] = int((l_first + i - 1) * p->npix_pad);
    }
    jobs.dw_partial = p->dw_part;
    jobs.db_partial = p->db_part;
    jobs.nx = W;
    jobs.ny_total = W;
    jobs.prob0 = l_first - 1;
    jobs.prob_total = nh;
    jobs.interleave = interleave;
  };
  const bool fuse_l0 = p->fuse_l0 && nh >= 1 && p->nchunks == 1;
  const bool merged = p->bwd_merged && p->dw_ctas > 0 && p->nchunks == 1;
  // dZ chain: dz[l-1] = (dz[l] * omega_{l-1} W_l) .* cos(...)
  for (int l = nh; l >= 1; --l) {
    RowGemmArgs ra{};
    ra.num_tiles = ch.ntiles;
    ra.a_row0 = int(l * p->npix_pad + ch.p0);
    ra.e_row0 = int((l - 1) * p->npix_pad + ch.p0);
    ra.o_row0 = int((l - 1) * p->npix_pad + ch.p0);
    ra.valid_rows = int(ch.npix);
    ra.b_early = 1;  // omega W^T was staged at the start of the step, at least two kernels ago
    p->stall_slot = 6 + l;
    int rc;
    bool done = false;
    if (merged) {
      ColGemmJobs jobs{};
      const bool own_split = l == 1 && fuse_l0 && p->merged_splits_l0 > 0;
      fill_jobs(jobs, l, 1, own_split ? p->merged_splits_l0 : p->active_splits, 1);
      // the tail kernel wrote dz[nh] in direction tail_rev: layer nh's launch sweeps the other way, layer nh - 1's
      // back again, ... (both roles of a launch sweep the same way: the pace hint counts sweep positions)
      if ((p->alt_sweep & 4) && p->tail_ran) ra.reverse = jobs.reverse = p->tail_rev ^ (((nh - l) % 2 == 0) ? 1 : 0);
      if (l == 1) p->l1_splits = jobs.splits;
      if (l == 1 && fuse_l0) {
        ra.gen_coord = p->coord;
        ra.gen_coord.p_offset = ch.p0;
        ra.red_part = p->l0_part;
        rc = launch_bwd_merged<W, true>(p, ra, jobs, l, st);
        p->l0_used = kRedWarpsPerChunk * p->last_rowgemm_grid;
      } else {
        rc = launch_bwd_merged<W, false>(p, ra, jobs, l, st);
      }
      if (rc) return rc;
      continue;
    }
    {
      if (l == 1 && fuse_l0) {
        ra.gen_coord = p->coord;
        ra.gen_coord.p_offset = ch.p0;
        ra.red_part = p->l0_part;
        rc = launch_rowgemm<W, MODE_DX, false, true>(p, p->tm_dz, p->tm_wt[0], p->tm_act, p->tm_dz, ra, st);
        p->l0_used = kRedWarpsPerChunk * p->last_rowgemm_grid;  // one partial row per reducer warp of a chunk
        done = true;
      }
    }
    if (!done) rc = launch_rowgemm<W, MODE_DX>(p, p->tm_dz, p->tm_wt[l - 1], p->tm_act, p->tm_dz, ra, st);
    if (rc) return rc;
  }
  // hidden-layer weight / bias gradients: one split-K launch over all layers
  if (nh > 0 && !merged) {
    ColGemmJobs jobs{};
    static const int col_interleave = getenv("SIRENB200_COL_INTERLEAVE") ? atoi(getenv("SIRENB200_COL_INTERLEAVE")) : 0;
    fill_jobs(jobs, 1, nh, p->col_splits, col_interleave);
    int rc;
    if constexpr (W == 512) {
      rc = p->bwd_pair ? launch_colgemm2<W>(p, jobs, st) : launch_colgemm<W>(p, jobs, st);
    } else {
      rc = launch_colgemm<W>(p, jobs, st);
    }
    if (rc) return rc;
  }
  if (!fuse_l0) {
    CoordSrc cs = p->coord;
    cs.p_offset = ch.p0;
    int grid = p->l0_grid;
    p->l0_used = p->l0_grid * p->nchunks;
    ProfScope ps(p, PK_L0_GRAD, st);
    static bool l0_attr[64] = {};
    if (!l0_attr[p->device & 63]) {
      CUDA_TRY(cudaFuncSetAttribute(tc_layer0_grad_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    int(L0GradCfg<W>::SMEM_BYTES)));
      l0_attr[p->device & 63] = true;
    }
    tc_layer0_grad_kernel<W><<<grid, 256, L0GradCfg<W>::SMEM_BYTES, st>>>(cs, p->dz + ch.p0 * W,
                                                  p->l0_part + size_t(ch.index) * p->l0_grid * 3 * W,
                                                  ch.npix);
  }
  LAUNCH_CHECK();
  return 0;
}

// descriptors of every partial buffer -> the caller's gradient tensors (order = model.parameters())
void tc_build_reduce(const sirenb200_plan* p, float* const* grads, float scale, float* stats, int nchunks,
                     ReduceArgs& ra) {
  const int D = p->D, C = p->C, W = p->W, Wm = p->Wm, nh = D - 2;
  const bool padded = Wm != W;
  int nd = 0;
  // n gradient elements; cols > 0: the gradient is a [n / cols, cols] matrix inside a [*, W]-pitched partial slab
  auto add = [&](float* dst, const float* src, int n, int nsplit, int64_t stride, int cols) {
    ra.d[nd].dst = dst;
    ra.d[nd].src = src;
    ra.d[nd].n = n;
    ra.d[nd].nsplit = nsplit;
    ra.d[nd].split_stride = stride;
    // weight-gradient partials (few splits, many elements, 16-byte aligned rows): vectorised path
    ra.d[nd].vec = (!padded && n >= 4096 && n % 4 == 0 && stride % 4 == 0 && nsplit <= 64 &&
                    (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0)
                       ? 1 : 0;
    ra.d[nd].cols = padded ? cols : 0;
    ra.d[nd].cols_pad = padded ? W : 0;
    ++nd;
  };
  add(grads[0], p->l0_part, 2 * Wm, p->l0_used, 3 * W, 0);   // [W, 2] block: the model's rows come first
  add(grads[1], p->l0_part + 2 * W, Wm, p->l0_used, 3 * W, 0);
  for (int l = 1; l <= nh; ++l) {
    const int ns = (l == 1 && p->l1_splits > 0) ? p->l1_splits : p->active_splits;
    add(grads[2 * l], p->dw_part + size_t(l - 1) * W * W, Wm * Wm, ns, int64_t(nh) * W * W, Wm);
    add(grads[2 * l + 1], p->db_part + size_t(l - 1) * W, Wm, ns, int64_t(nh) * W, 0);
  }
  const int64_t lstride = int64_t(C) * W + C + 1;
  add(grads[2 * (D - 1)], p->last_part, C * Wm, p->last_grid * nchunks, lstride, Wm);
  add(grads[2 * (D - 1) + 1], p->last_part + C * W, C, p->last_grid * nchunks, lstride, 0);
  ra.ndesc = nd;
  int chunks = 0;
  for (int i = 0; i < nd; ++i) {
    ra.chunk_begin[i] = chunks;
    chunks += cdiv(ra.d[i].n, ra.d[i].vec ? 1024 : 32);
  }
  ra.chunk_begin[nd] = chunks;
  ra.scale = scale;
  ra.gscale = p->gstate;
  ra.stats = stats;
}

// reduce every partial buffer into the caller's gradient tensors
int tc_reduce(sirenb200_plan* p, float* const* grads, float scale, float* stats, int nchunks,
              cudaStream_t st) {
  ReduceArgs ra{};
  tc_build_reduce(p, grads, scale, stats, nchunks, ra);
  {
    ProfScope ps(p, PK_REDUCE, st);
    reduce_partials_kernel<<<ra.chunk_begin[ra.ndesc], 256, 0, st>>>(ra);
  }
  LAUNCH_CHECK();
  return 0;
}

// forward only (pred), forward + MSE + backward (mode 1), or backward from dpred (mode 2, no forward)
template <int W>
int tc_run(sirenb200_plan* p, const float* const* prm, int mode, const float* img_or_dpred,
           float* pred, float* const* grads, float scale, float* stats, cudaStream_t st) {
  const std::vector<Chunk> chunks = make_chunks(p);
  int rc = 0;
  if (mode != 2) rc = tc_prep<W>(p, prm, st, mode == 1 ? stats : nullptr);
  for (const Chunk& ch : chunks) {
    if (rc) return rc;
    // training step with at least two hidden GEMM layers: the last one is fused with the output layer
    const bool tail = W <= 256 && mode == 1 && p->tail_fused && p->last_tc && p->D - 2 >= 2 && p->nchunks == 1 &&
                      img_or_dpred;
    p->tail_ran = tail;
    if (mode != 2) rc = tc_forward_chunk<W>(p, prm, ch, st, tail);
    if (!rc) rc = tail ? launch_tail<W>(p, prm, img_or_dpred, pred, ch, st)
                       : tc_last_chunk<W>(p, prm, mode, img_or_dpred, pred, ch, st);
    if (!rc && mode != 0) rc = tc_backward_chunk<W>(p, prm, ch, st);
  }
  if (!rc && mode != 0 && !p->defer_reduce) rc = tc_reduce(p, grads, scale, stats, int(chunks.size()), st);
  return rc;
}

int tc_dispatch(sirenb200_plan* p, const float* const* prm, int mode, const float* img_or_dpred,
                float* pred, float* const* grads, float scale, float* stats, cudaStream_t st) {
  switch (p->W) {
    case 128: return tc_run<128>(p, prm, mode, img_or_dpred, pred, grads, scale, stats, st);
    case 256: return tc_run<256>(p, prm, mode, img_or_dpred, pred, grads, scale, stats, st);
    case 512: return tc_run<512>(p, prm, mode, img_or_dpred, pred, grads, scale, stats, st);
    default: return fail(SIRENB200_ERR_INVALID, "no tensor-core kernels for hidden %d", p->W);
  }
}

// ---------------------------------------------------------------------------------------
// fp32 path
// ---------------------------------------------------------------------------------------
template <int OP>
int launch_simt(const SimtGemmArgs& a, int splits, cudaStream_t st) {
  dim3 grid(cdiv(a.N, 64), cdiv(a.M, 64), splits);
  simt_gemm_kernel<OP><<<grid, 256, 0, st>>>(a);
  LAUNCH_CHECK();
  return 0;
}

// FusedMovingAvgObsFakeQuantize on one tensor: x <- fake_quant(x) in place, optional act = sin(omega x), mask
int actq_run(sirenb200_plan* p, float* x, int64_t n, float* state, unsigned char* mask, float* act, float omega,
             cudaStream_t st) {
  if (p->actq_training) {
    actq_minmax_kernel<<<kActqBlocks, 256, 0, st>>>(x, n, p->actq_partial);
    LAUNCH_CHECK();
  }
  actq_update_kernel<<<1, 32, 0, st>>>(p->actq_partial, kActqBlocks, state, p->actq_training, p->actq_avg, p->actq_qmin,
                                       p->actq_qmax);
  LAUNCH_CHECK();
  int grid = cdiv(n, 256 * 4);
  if (grid > p->nsm * 8) grid = p->nsm * 8;
  actq_apply_kernel<<<grid, 256, 0, st>>>(x, n, state, p->actq_qmin, p->actq_qmax, x, mask, act, omega);
  LAUNCH_CHECK();
  return 0;
}

int f32_forward(sirenb200_plan* p, const float* const* prm, int mode, const float* img_or_dpred,
                float* pred, cudaStream_t st) {
  const int L = p->nlin, W = p->W, C = p->C;
  const bool fourier = p->model_kind == 1;
  const int64_t n = p->npix;
  const float* in;
  int in_dim = p->k0;
  if (fourier) {
    if (!p->encB) return fail(SIRENB200_ERR_STATE, "FourierNet: call sirenb200_set_fourier_encoding first");
    const int half = p->k0 / 2;
    fourier_encode_kernel<<<cdiv(n * half, 256), 256, 0, st>>>(p->coord, p->encB, half, p->enc32, n);
    LAUNCH_CHECK();
    in = p->enc32;
  } else {
    simt_coords_kernel<<<cdiv(n, 256), 256, 0, st>>>(p->coord, p->x32, n);
    LAUNCH_CHECK();
    in = p->x32;
  }
  for (int l = 0; l < L - 1; ++l) {
    SimtGemmArgs a{};
    a.A = in;
    a.B = prm[2 * l];
    a.bias = prm[2 * l + 1];
    a.Z = p->z32 + size_t(l) * n * W;
    a.Out = p->a32 + size_t(l) * n * W;
    a.M = int(n);
    a.N = W;
    a.K = in_dim;
    a.lda = in_dim;
    a.ldb = in_dim;
    a.ldo = W;
    a.omega = omega_of(p, l);
    int rc = fourier ? launch_simt<OP_NT_RELU>(a, 1, st) : launch_simt<OP_NT_SINE>(a, 1, st);
    if (rc) return rc;
    if (p->actq_on) {  // z <- fake_quant(z), a <- sin(omega z_q)   (activation_post_process of the qat Linear)
      rc = actq_run(p, a.Z, n * W, p->actq_state + 4 * l, p->actq_mask + size_t(l) * n * W, a.Out, a.omega, st);
      if (rc) return rc;
    }
    in = a.Out;
    in_dim = W;
  }
  {
    SimtGemmArgs a{};
    a.A = in;
    a.B = prm[2 * (L - 1)];
    a.bias = prm[2 * (L - 1) + 1];
    a.Out = p->y32;
    a.M = int(n);
    a.N = C;
    a.K = in_dim;
    a.lda = in_dim;
    a.ldb = in_dim;
    a.ldo = C;
    int rc = launch_simt<OP_NT_LIN>(a, 1, st);
    if (rc) return rc;
    if (p->actq_on) {
      rc = actq_run(p, p->y32, n * C, p->actq_state + 4 * (L - 1), p->actq_mask + size_t(L - 1) * n * W, nullptr, 0.f,
                    st);
      if (rc) return rc;
    }
  }
  LossArgs la{};
  la.y = p->y32;
  la.img = img_or_dpred;
  la.pred = pred;
  la.g = p->g32;
  la.loss_partial = p->loss_part;
  la.n = n * C;
  la.mode = mode;
  la.outermost_linear = p->cfg.outermost_linear;
  la.omega = omega_of(p, L - 1);
  la.out_kind = fourier ? 1 : 0;
  simt_loss_kernel<<<256, 256, 0, st>>>(la);
  LAUNCH_CHECK();
  return 0;
}

int f32_backward(sirenb200_plan* p, const float* const* prm, float* const* grads, float scale,
                 float* stats, cudaStream_t st) {
  const int D = p->nlin, W = p->W, C = p->C;  // D = number of linear layers here
  const bool fourier = p->model_kind == 1;
  const int64_t n = p->npix;
  const float* g = p->g32;  // dL/dz of layer l (seed units)
  int gdim = C;
  int pp = 0;
  auto mask_grad = [&](float* gbuf, int layer, int64_t count) -> int {
    // straight-through estimator of the activation fake-quant: no gradient where the value was clipped
    int grid = cdiv(count, 256 * 4);
    if (grid > p->nsm * 8) grid = p->nsm * 8;
    mul_mask_u8_kernel<<<grid, 256, 0, st>>>(gbuf, p->actq_mask + size_t(layer) * n * W, count);
    LAUNCH_CHECK();
    return 0;
  };
  if (p->actq_on) {
    int rc = mask_grad(p->g32, D - 1, n * C);
    if (rc) return rc;
  }
  for (int l = D - 1; l >= 0; --l) {
    const float* xin = (l == 0) ? (fourier ? p->enc32 : p->x32) : p->a32 + size_t(l - 1) * n * W;
    const int xdim = (l == 0) ? p->k0 : W;
    // dW_l = g^T xin (split over pixels), db_l = column sums of g
    SimtGemmArgs a{};
    a.A = g;
    a.B = xin;
    a.Out = p->part32;
    a.ColSum = p->part32 + size_t(p->simt_splits) * gdim * xdim;
    a.M = gdim;
    a.N = xdim;
    a.K = int(n);
    a.lda = gdim;
    a.ldb = xdim;
    a.ldo = xdim;
    a.ksplit_len = p->simt_split_len;
    int rc = launch_simt<OP_TN_PART>(a, p->simt_splits, st);
    if (rc) return rc;
    ReduceArgs ra{};
    ra.d[0] = {grads[2 * l], p->part32, gdim * xdim, p->simt_splits, int64_t(gdim) * xdim, 0};
    ra.d[1] = {grads[2 * l + 1], a.ColSum, gdim, p->simt_splits, int64_t(gdim), 0};
    ra.ndesc = 2;
    ra.chunk_begin[0] = 0;
    ra.chunk_begin[1] = cdiv(gdim * xdim, 32);
    ra.chunk_begin[2] = ra.chunk_begin[1] + cdiv(gdim, 32);
    ra.scale = scale;
    ra.gscale = nullptr;
    ra.stats = stats;
    reduce_partials_kernel<<<ra.chunk_begin[2], 256, 0, st>>>(ra);
    LAUNCH_CHECK();
    if (l > 0) {
      // g_{l-1} = (g_l W_l) .* omega cos(omega z_{l-1})
      SimtGemmArgs b{};
      b.A = g;
      b.B = prm[2 * l];
      b.Z = p->z32 + size_t(l - 1) * n * W;
      b.Out = p->dz32[pp];
      b.M = int(n);
      b.N = W;
      b.K = gdim;
      b.lda = gdim;
      b.ldb = W;
      b.ldo = W;
      b.omega = omega_of(p, l - 1);
      rc = fourier ? launch_simt<OP_NN_DRELU>(b, 1, st) : launch_simt<OP_NN_DCOS>(b, 1, st);
      if (rc) return rc;
      if (p->actq_on) {
        rc = mask_grad(p->dz32[pp], l - 1, n * W);
        if (rc) return rc;
      }
      g = p->dz32[pp];
      gdim = W;
      pp ^= 1;
    }
  }
  return 0;
}

int finalize(sirenb200_plan* p, const float* parts, int nparts, int64_t stride, float* stats,
             bool tc, cudaStream_t st) {
  FinalizeArgs fa{};
  fa.loss_partial = parts;
  fa.nparts = nparts;
  fa.part_stride = stride;
  fa.inv_count = p->inv_count;
  fa.stats = stats;
  fa.gstate = tc ? p->gstate : nullptr;
  {
    ProfScope ps(p, PK_FINALIZE, st);
    finalize_loss_kernel<<<1, 256, 0, st>>>(fa);
  }
  LAUNCH_CHECK();
  return 0;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
// Scratch for the stateless entry points comes from the device's default stream-ordered pool.  Its default
// release threshold is 0 (memory goes back to the driver at every synchronisation, so each call pays a
// fresh allocation of milliseconds); keep what was allocated.
static void keep_pool_memory() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[dev] = true;
}

extern "C" {

int sirenb200_version(void) { return 100; }
const char* sirenb200_last_error(void) { return g_err; }
int64_t sirenb200_launch_count(void) { return g_launches.load(); }

int sirenb200_create(const sirenb200_config_t* cfg, sirenb200_handle_t* out) {
  if (!cfg || !out) return fail(SIRENB200_ERR_INVALID, "null argument");
  if (cfg->depth < 2 || cfg->depth > kMaxLayers)
    return fail(SIRENB200_ERR_INVALID, "depth %d not in [2, %d]", cfg->depth, kMaxLayers);
  if (cfg->hidden < 1 || cfg->in_features != 2 || cfg->out_features < 1 ||
      cfg->out_features > kMaxOut)
    return fail(SIRENB200_ERR_INVALID, "unsupported layer sizes (hidden %d, in %d, out %d)",
                cfg->hidden, cfg->in_features, cfg->out_features);
  if (cfg->height < 1 || cfg->width < 1 || cfg->row_begin < 0 || cfg->row_end > cfg->height ||
      cfg->row_begin >= cfg->row_end)
    return fail(SIRENB200_ERR_INVALID, "bad image geometry %dx%d rows [%d,%d)", cfg->height,
                cfg->width, cfg->row_begin, cfg->row_end);
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(SIRENB200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is sm_100a only", dev,
                prop.major, prop.minor);
  if (cfg->precision == SIRENB200_PREC_F16TC && (cfg->hidden > 512 || (!tc_supported(cfg->hidden) && cfg->depth < 3)))
    return fail(SIRENB200_ERR_INVALID,
                "tensor-core path supports hidden <= 512 (any width with depth >= 3: zero-padded to 128 / 256 / 512); "
                "got hidden %d depth %d (use SIRENB200_PREC_FP32)", cfg->hidden, cfg->depth);
  if (cfg->precision != SIRENB200_PREC_F16TC && cfg->precision != SIRENB200_PREC_FP32)
    return fail(SIRENB200_ERR_INVALID, "unknown precision %d", cfg->precision);

  const int model_kind = cfg->reserved[0], map_size = cfg->reserved[1];
  if (model_kind != 0 && model_kind != 1) return fail(SIRENB200_ERR_INVALID, "unknown model kind %d", model_kind);
  if (model_kind == 1) {  // FourierNet (models/fourier.py)
    if (cfg->precision != SIRENB200_PREC_FP32)
      return fail(SIRENB200_ERR_INVALID, "FourierNet runs on the fp32 path (SIRENB200_PREC_FP32)");
    if (map_size < 2 || (map_size & 1)) return fail(SIRENB200_ERR_INVALID, "FourierNet: map_size %d must be even", map_size);
    if (cfg->depth < 3) return fail(SIRENB200_ERR_INVALID, "FourierNet: depth %d < 3", cfg->depth);
  }
  sirenb200_plan* p = new sirenb200_plan();
  p->cfg = *cfg;
  p->device = dev;
  p->nsm = prop.multiProcessorCount;
  p->model_kind = model_kind;
  p->nlin = model_kind == 1 ? cfg->depth - 1 : cfg->depth;
  p->k0 = model_kind == 1 ? map_size : cfg->in_features;
  p->D = p->nlin;  // every per-layer loop below counts LINEAR layers
  p->Wm = cfg->hidden;
  p->W = cfg->hidden;
  if (cfg->precision == SIRENB200_PREC_F16TC)  // kernel width: the tcgen05 kernels exist for 128 / 256 / 512 columns
    p->W = cfg->hidden <= 128 ? 128 : (cfg->hidden <= 256 ? 256 : 512);
  p->C = cfg->out_features;
  p->rows = cfg->row_end - cfg->row_begin;
  p->npix = int64_t(p->rows) * cfg->width;
  p->ntiles = cdiv(p->npix, kRowsPerTile);
  p->npix_pad = int64_t(p->ntiles) * kRowsPerTile;
  p->inv_count = float(1.0 / (double(cfg->height) * cfg->width * cfg->out_features));
  p->coord.width = cfg->width;
  p->coord.row_begin = cfg->row_begin;
  const int D = p->D, W = p->W, C = p->C;
  int rc = 0;
#define ALLOC(ptr, count)                 \
  if ((rc = dev_alloc(p, &(ptr), (count)))) { \
    sirenb200_destroy(p);                 \
    return rc;                            \
  }
  ALLOC(p->gstate, 4);
  ALLOC(p->eval_acc, 2);
  ALLOC(p->bar, 1);
  ALLOC(p->pace, 2 * kMaxLayers);
  cudaMemset(p->pace, 0, 2 * kMaxLayers * sizeof(unsigned int));
  if (cudaMemset(p->bar, 0, sizeof(*p->bar)) != cudaSuccess) {
    sirenb200_destroy(p);
    return fail(SIRENB200_ERR_CUDA, "cudaMemset failed");
  }
  {
    const float init[4] = {1.f, 4096.f, 0.f, 0.f};
    cudaError_t e = cudaMemcpy(p->gstate, init, sizeof(init), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      sirenb200_destroy(p);
      return fail(SIRENB200_ERR_CUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
    }
  }
  if (cfg->precision == SIRENB200_PREC_FP32) {
    ALLOC(p->loss_part, 256);
    ALLOC(p->x32, p->npix * 2);
    ALLOC(p->z32, int64_t(D - 1) * p->npix * W);
    ALLOC(p->a32, int64_t(D - 1) * p->npix * W);
    ALLOC(p->y32, p->npix * C);
    ALLOC(p->g32, p->npix * C);
    ALLOC(p->dz32[0], p->npix * W);
    ALLOC(p->dz32[1], p->npix * W);
    int splits = int((p->npix + 4095) / 4096);
    if (splits > 64) splits = 64;
    if (splits < 1) splits = 1;
    p->simt_splits = splits;
    p->simt_split_len = cdiv(cdiv(p->npix, splits), 16) * 16;
    if (p->model_kind == 1) ALLOC(p->enc32, p->npix * p->k0);
    int64_t big = int64_t(W) * W;
    if (big < int64_t(p->k0) * W) big = int64_t(p->k0) * W;
    if (big < int64_t(C) * W) big = int64_t(C) * W;
    ALLOC(p->part32, int64_t(splits) * (big + W + 16));
  } else {
    const int nh = D - 2;
    if (int64_t(D - 1) * p->npix_pad >= (int64_t(1) << 31)) {
      sirenb200_destroy(p);
      return fail(SIRENB200_ERR_INVALID, "shard too large for 32-bit TMA row coordinates");
    }
    ALLOC(p->act, int64_t(D - 1) * p->npix_pad * W);
    ALLOC(p->dz, int64_t(D - 1) * p->npix_pad * W);
    ALLOC(p->wh, int64_t(nh > 0 ? nh : 1) * W * W);
    ALLOC(p->wth, int64_t(nh > 0 ? nh : 1) * W * W);
    ALLOC(p->tab0, W);
    ALLOC(p->wl16, 16 * W);
    ALLOC(p->wlt16, int64_t(W) * 64);
    ALLOC(p->bias_w, int64_t(nh > 0 ? nh : 1) * W);
    ALLOC(p->bias_raw, int64_t(nh > 0 ? nh : 1) * W);
    ALLOC(p->w0p, 2 * W);
    ALLOC(p->b0p, W);
    // pixel splits of the weight-gradient GEMM: one CTA per (layer, 128-row block, column part, split).
    // One wave of CTAs when that fills the machine (>= 93 % of the SMs), otherwise two balanced waves
    // (hidden 512, depth 6: 32 x 4 = 128 jobs would leave 20 SMs idle; 32 x 9 = 288 = 2 x 144 does not).
    int splits = 1;
    if (nh > 0) {
      const int base = nh * (W / 128) * (W / (W < 256 ? W : 256));
      splits = p->nsm / base;
      if (splits < 1) splits = 1;
      const int two = (2 * p->nsm) / base;
      if (base * splits * 100 < p->nsm * 93 && two > splits && base * two * 100 >= 2 * p->nsm * 93) splits = two;
    }
    if (splits < 1) splits = 1;
    if (splits > p->ntiles) splits = p->ntiles;
    p->col_splits = splits;  // (re-clamped to the chunk size below)
    {
      // merged dX + dW launches: the reduction role gets ~35 % of the SMs (52 of 148: its MMAs are busy ~1.1 k cycles
      // per tile and CTA against ~3.7 k of epilogue per tile on a dX CTA; A/B at c2: 44 / 52 / 60 CTAs = 1037 / 1045 /
      // 1030 steps/s), as (W/128 row blocks) x (column parts) x splits
      const char* env = getenv("SIRENB200_BWD_MERGED");
      // (hidden 512 keeps the separate launches: measured at config 3, 18.1 steps/s separate vs 15.1 merged — its
      // reduction role needs 8 CTAs per pixel split and streams twice the operand bytes per tile)
      p->bwd_merged = nh > 0 && (env ? atoi(env) != 0 : W <= 256);
      const int per_split = (W / 128) * (W / (W < 256 ? W : 256));
      int want = (p->nsm * 35 + 50) / 100;
      env = getenv("SIRENB200_DW_CTAS");
      if (env && atoi(env) > 0) want = atoi(env);
      int ms = want / per_split;
      if (ms < 1) ms = 1;
      if (ms > p->ntiles) ms = p->ntiles;
      p->dw_ctas = ms * per_split;
      if (p->dw_ctas >= p->nsm) p->bwd_merged = false;
      env = getenv("SIRENB200_PACE_WINDOW");
      if (env && atoi(env) > 0) p->pace_window = p->pace_window_dx = atoi(env);
      env = getenv("SIRENB200_PACE_DX");
      if (env) p->pace_window_dx = atoi(env);
      env = getenv("SIRENB200_PAIR");
      p->bwd_pair = !(env && atoi(env) == 0);
      if (p->bwd_merged) p->merged_splits = ms;
      // the first hidden layer's launch: its dX role also reduces layer 0's gradient (~6.8 k cycles per tile against
      // ~5.5 k) and writes no dz, so the reduction role gets 30 % of the SMs there (A/B at c2, dX-class time per step:
      // 52 / 44 / 36 / 28 CTAs = 535 / 518 / 546 / 585 us)
      {
        int want0 = (p->nsm * 30 + 50) / 100;
        env = getenv("SIRENB200_DW_CTAS_L0");
        if (env && atoi(env) > 0) want0 = atoi(env);
        int m0 = want0 / per_split;
        if (m0 < 1) m0 = 1;
        if (m0 > ms) m0 = ms;  // the partial slabs are sized for merged_splits
        if (p->bwd_merged) p->merged_splits_l0 = m0;
      }
    }
    const int slabs = (p->bwd_merged && p->merged_splits > splits) ? p->merged_splits : splits;
    ALLOC(p->dw_part, int64_t(slabs) * (nh > 0 ? nh : 1) * W * W);
    ALLOC(p->db_part, int64_t(slabs) * (nh > 0 ? nh : 1) * W);
    p->chunk_tiles = p->ntiles;  // the whole shard is one chunk (row chunking lost to launch overheads)
    p->nchunks = 1;
    if (p->col_splits > p->chunk_tiles) p->col_splits = p->chunk_tiles;
    p->active_splits = p->bwd_merged ? p->merged_splits : p->col_splits;
    const int64_t chunk_pad = int64_t(p->chunk_tiles) * kRowsPerTile;
    {
      const char* env = getenv("SIRENB200_LAST_TC");
      p->last_tc = (W == 128 || W == 256 || W == 512) && nh > 0 && !(env && atoi(env) == 0);
      env = getenv("SIRENB200_GEN_FIRST");
      p->gen_first = !(env && atoi(env) == 0);
      env = getenv("SIRENB200_TAIL");
      p->tail_fused = !(env && atoi(env) == 0);
      env = getenv("SIRENB200_PDL");
      p->pdl = !(env && atoi(env) == 0);
      env = getenv("SIRENB200_ALT_SWEEP");
      if (env) p->alt_sweep = atoi(env);
      env = getenv("SIRENB200_L2_HINTS");
      if (env) p->l2_hints = atoi(env);
      env = getenv("SIRENB200_FUSE_L0");
      p->fuse_l0 = !(env && atoi(env) == 0);
    }
    p->last_grid = p->nsm * 2;
    if (int64_t(p->last_grid) * 8 > chunk_pad) p->last_grid = cdiv(chunk_pad, 8);
    if (p->last_tc) {  // persistent tensor-core kernel: one CTA per SM (or per tile)
      p->last_grid = p->nsm;
      if (p->last_grid > p->chunk_tiles) p->last_grid = p->chunk_tiles;
    }
    ALLOC(p->last_part, int64_t(p->nchunks) * p->last_grid * (C * W + C + 1));
    p->l0_grid = p->nsm * 2;
    if (p->l0_grid > p->chunk_tiles) p->l0_grid = p->chunk_tiles;
    {  // room for the stand-alone kernel's l0_grid rows or the dX-fused reducer's 2 rows per CTA
      int rows = kRedWarpsPerChunk * p->nsm;
      if (rows < p->l0_grid) rows = p->l0_grid;
      ALLOC(p->l0_part, int64_t(p->nchunks) * rows * 3 * W);
    }
    cudaError_t e = cudaMemset(p->dz, 0, size_t(D - 1) * p->npix_pad * W * sizeof(__half));
    if (e == cudaSuccess) e = cudaMemset(p->act, 0, size_t(D - 1) * p->npix_pad * W * sizeof(__half));
    if (e != cudaSuccess) {
      sirenb200_destroy(p);
      return fail(SIRENB200_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    int trc = make_tmap_16bit(&p->tm_act, p->act, uint64_t(D - 1) * p->npix_pad, W, 128, false);
    trc |= make_tmap_16bit(&p->tm_dz, p->dz, uint64_t(D - 1) * p->npix_pad, W, 128, false);
    if (nh > 0) {
      // weight-gradient reduction: one box per operand and stage = {64 columns, pixel rows, chunks}; dz chunks
      // = one 128-row block of dW, activation chunks = one column part of dW
      const uint32_t ychunks = (W < 256 ? W : 256) / 64;
      trc |= make_tmap_16bit_chunks(&p->tm_dz_c, p->dz, uint64_t(D - 1) * p->npix_pad, W, 128, 2);
      trc |= make_tmap_16bit_chunks(&p->tm_act_c, p->act, uint64_t(D - 1) * p->npix_pad, W, 128, ychunks);
      // (CTA pairs: each CTA loads half of the activation columns)
      trc |= make_tmap_16bit_chunks(&p->tm_act_c2, p->act, uint64_t(D - 1) * p->npix_pad, W, 128, 2);
    }
    p->tm_w.resize(nh > 0 ? nh : 0);
    p->tm_wt.resize(nh > 0 ? nh : 0);
    for (int l = 0; l < nh; ++l) {
      const uint32_t brows = W < 256 ? W : 256;  // B box rows = output columns per work item
      trc |= make_tmap_16bit(&p->tm_w[l], p->wh + size_t(l) * W * W, W, W, brows, false);
      trc |= make_tmap_16bit(&p->tm_wt[l], p->wth + size_t(l) * W * W, W, W, brows, false);
    }
    if (p->last_tc) {
      trc |= make_tmap_16bit(&p->tm_wl, p->wl16, 16, W, 16, false);
    }
    if (trc) {
      sirenb200_destroy(p);
      return fail(SIRENB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", trc);
    }
    if (getenv("SIRENB200_TIMELINE") || getenv("SIRENB200_STALLS")) {
      // timeline slots, then (SIRENB200_STALLS) kStallLaunches x 160 CTAs x 16 wait-cycle counters
      const int64_t n = kTimelineSlots + (getenv("SIRENB200_STALLS") ? int64_t(kStallLaunches) * 160 * 16 : 0);
      ALLOC(p->dbg_timeline, n);
      cudaMemset(p->dbg_timeline, 0, n * sizeof(long long));
      if (getenv("SIRENB200_STALLS")) p->dbg_stall = p->dbg_timeline + kTimelineSlots;
    }
  }
#undef ALLOC
  *out = p;
  return 0;
}

int sirenb200_destroy(sirenb200_handle_t p) {
  if (!p) return 0;
  void* ptrs[] = {p->enc32, p->actq_mask, p->actq_partial, p->bar, p->pace, p->gstate, p->loss_part, p->eval_acc, p->x32,     p->z32,     p->a32,
                  p->y32,    p->g32,       p->dz32[0],  p->dz32[1], p->part32,  p->act,
                  p->dz,     p->wh,        p->wth,      p->dw_part, p->db_part, p->last_part,
                  p->l0_part, p->tab0,     p->bias_w,   p->bias_raw, p->dbg_timeline, p->w0p, p->b0p,
                  p->wl16,   p->wlt16};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  for (cudaEvent_t e : p->prof_ev) cudaEventDestroy(e);
  delete p;
  return 0;
}

int64_t sirenb200_workspace_bytes(sirenb200_handle_t h) { return h ? h->bytes : 0; }

int sirenb200_debug_timeline(sirenb200_handle_t h, long long* h_out, int32_t n) {
  if (!h || !h->dbg_timeline) return fail(SIRENB200_ERR_STATE, "timeline capture not enabled");
  CUDA_TRY(cudaMemcpy(h_out, h->dbg_timeline, size_t(n) * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

int sirenb200_profile_enable(sirenb200_handle_t h, int32_t enable) {
  if (!h) return fail(SIRENB200_ERR_INVALID, "null handle");
  h->prof_on = enable != 0;
  h->prof_used = 0;
  h->prof_kind.clear();
  return 0;
}

int sirenb200_profile_read(sirenb200_handle_t h, float* h_total_ms, int32_t* h_count, int32_t n_kinds) {
  if (!h || !h_total_ms || !h_count) return fail(SIRENB200_ERR_INVALID, "null argument");
  for (int k = 0; k < n_kinds; ++k) {
    h_total_ms[k] = 0.f;
    h_count[k] = 0;
  }
  for (size_t i = 0; i < h->prof_kind.size(); ++i) {
    cudaEvent_t a = h->prof_ev[2 * i], b = h->prof_ev[2 * i + 1];
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    const int k = h->prof_kind[i];
    if (k < n_kinds) {
      h_total_ms[k] += ms;
      h_count[k] += 1;
    }
  }
  return 0;
}

int sirenb200_set_grid_lut(sirenb200_handle_t h, const float* lin_h, const float* lin_w) {
  if (!h || !lin_h || !lin_w) return fail(SIRENB200_ERR_INVALID, "null argument");
  h->coord.lin_h = lin_h;
  h->coord.lin_w = lin_w;
  h->coord.coords = nullptr;
  return 0;
}

int sirenb200_set_grid_coords(sirenb200_handle_t h, const float* coords) {
  if (!h || !coords) return fail(SIRENB200_ERR_INVALID, "null argument");
  h->coord.coords = coords;
  return 0;
}

int sirenb200_forward(sirenb200_handle_t h, const float* const* prm, float* pred,
                      sirenb200_stream_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!prm) return fail(SIRENB200_ERR_INVALID, "null parameter table");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (h->cfg.precision == SIRENB200_PREC_FP32) {
    rc = f32_forward(h, prm, 0, nullptr, pred, st);
  } else {
    rc = tc_dispatch(h, prm, 0, nullptr, pred, nullptr, 0.f, nullptr, st);
  }
  h->have_fwd = (rc == 0);
  return rc;
}

int sirenb200_forward_backward(sirenb200_handle_t h, const float* const* prm, const float* img,
                               float loss_scale, float* const* grads, float* stats,
                               sirenb200_stream_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!prm || !img || !grads || !stats) return fail(SIRENB200_ERR_INVALID, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float scale = loss_scale * h->inv_count;
  if (h->cfg.precision == SIRENB200_PREC_FP32) {
    CUDA_TRY(cudaMemsetAsync(stats, 0, 4 * sizeof(float), st));
    rc = f32_forward(h, prm, 1, img, nullptr, st);
    if (!rc) rc = f32_backward(h, prm, grads, scale, stats, st);
    if (!rc) rc = finalize(h, h->loss_part, 256, 1, stats, false, st);
  } else {
    rc = tc_dispatch(h, prm, 1, img, nullptr, grads, scale, stats, st);
    if (!rc)
      rc = finalize(h, h->last_part + h->C * h->W + h->C, h->last_grid * h->nchunks,
                    int64_t(h->C) * h->W + h->C + 1, stats, true, st);
  }
  h->have_fwd = (rc == 0);
  return rc;
}

int sirenb200_backward(sirenb200_handle_t h, const float* const* prm, const float* dpred,
                       float* const* grads, sirenb200_stream_t stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  if (!prm || !dpred || !grads) return fail(SIRENB200_ERR_INVALID, "null argument");
  if (!h->have_fwd)
    return fail(SIRENB200_ERR_STATE, "sirenb200_backward needs a preceding sirenb200_forward");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* stats = h->gstate;  // scratch flag lives in gstate[2] (index 2 of a 4-float block)
  if (h->cfg.precision == SIRENB200_PREC_FP32) {
    // re-run the loss stage on the stashed last-layer output to seed the gradient
    LossArgs la{};
    la.y = h->y32;
    la.img = dpred;
    la.g = h->g32;
    la.n = h->npix * h->C;
    la.mode = 2;
    la.outermost_linear = h->cfg.outermost_linear;
    la.omega = omega_of(h, h->D - 1);
    la.out_kind = h->model_kind == 1 ? 1 : 0;
    simt_loss_kernel<<<256, 256, 0, st>>>(la);
    LAUNCH_CHECK();
    rc = f32_backward(h, prm, grads, 1.0f, stats, st);
  } else {
    absmax_scale_kernel<<<1, 1024, 0, st>>>(dpred, h->npix * h->C, h->gstate);
    LAUNCH_CHECK();
    rc = tc_dispatch(h, prm, 2, dpred, nullptr, grads, 1.0f, stats, st);
  }
  return rc;
}

int sirenb200_set_fourier_encoding(sirenb200_handle_t h, const float* B) {
  if (!h || !B) return fail(SIRENB200_ERR_INVALID, "null argument");
  if (h->model_kind != 1) return fail(SIRENB200_ERR_STATE, "not a FourierNet handle");
  h->encB = B;
  return 0;
}

int sirenb200_set_act_quant(sirenb200_handle_t h, float* state, int32_t enable, int32_t training, float averaging_const,
                            int32_t qmin, int32_t qmax) {
  if (!h) return fail(SIRENB200_ERR_INVALID, "null handle");
  if (!enable) {
    h->actq_on = false;
    return 0;
  }
  if (h->cfg.precision != SIRENB200_PREC_FP32)
    return fail(SIRENB200_ERR_INVALID, "activation fake-quant runs on the fp32 path (create the handle with SIRENB200_PREC_FP32)");
  if (!state || qmax <= qmin) return fail(SIRENB200_ERR_INVALID, "set_act_quant: bad argument");
  if (!h->actq_mask) {
    const int64_t bytes = int64_t(h->D - 1) * h->npix * h->W + h->npix * h->C;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->actq_mask), size_t(bytes)));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->actq_partial), 2 * kActqBlocks * sizeof(float)));
    h->bytes += bytes + 2 * kActqBlocks * int64_t(sizeof(float));
  }
  h->actq_state = state;
  h->actq_on = true;
  h->actq_training = training ? 1 : 0;
  h->actq_avg = averaging_const;
  h->actq_qmin = qmin;
  h->actq_qmax = qmax;
  return 0;
}

int sirenb200_fakequant_per_tensor(const float* x, int64_t n, float* state, int32_t training, float averaging_const,
                                   int32_t qmin, int32_t qmax, float* out, uint8_t* mask, sirenb200_stream_t stream) {
  if (!x || !state || !out || n < 1 || qmax <= qmin) return fail(SIRENB200_ERR_INVALID, "fakequant_per_tensor: bad argument");
  keep_pool_memory();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* partial = nullptr;
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&partial), 2 * kActqBlocks * sizeof(float), st));
  if (training) {
    actq_minmax_kernel<<<kActqBlocks, 256, 0, st>>>(x, n, partial);
    LAUNCH_CHECK();
  }
  actq_update_kernel<<<1, 32, 0, st>>>(partial, kActqBlocks, state, training ? 1 : 0, averaging_const, qmin, qmax);
  LAUNCH_CHECK();
  int grid = cdiv(n, 256 * 4);
  if (grid > 148 * 8) grid = 148 * 8;
  actq_apply_kernel<<<grid, 256, 0, st>>>(x, n, state, qmin, qmax, out, mask, nullptr, 0.f);
  LAUNCH_CHECK();
  CUDA_TRY(cudaFreeAsync(partial, st));
  return 0;
}

int sirenb200_pack_stream(int32_t n_items, const void* const* h_src, void* const* h_dst, const int32_t* h_kind,
                          const int64_t* h_count, const int64_t* h_offset, const int64_t* h_aux, uint8_t* stream_buf,
                          sirenb200_stream_t stream) {
  if (n_items < 1 || n_items > kPackMaxItems || !h_kind || !h_count || !h_offset || !stream_buf)
    return fail(SIRENB200_ERR_INVALID, "pack_stream: bad argument (1 <= items <= %d)", kPackMaxItems);
  PackArgs a{};
  long long most = 1;
  for (int i = 0; i < n_items; ++i) {
    const int k = h_kind[i];
    if (k < PACK_F32_TO_F16 || k > UNPACK_GATHER_U16) return fail(SIRENB200_ERR_INVALID, "pack_stream: kind %d", k);
    const bool unpack = k >= UNPACK_F16_TO_F32;
    if ((unpack && (!h_dst || !h_dst[i])) || (!unpack && (!h_src || !h_src[i])) || h_count[i] < 0 || h_offset[i] < 0)
      return fail(SIRENB200_ERR_INVALID, "pack_stream: item %d is incomplete", i);
    a.item[i].src = h_src ? h_src[i] : nullptr;
    a.item[i].dst = h_dst ? h_dst[i] : nullptr;
    a.item[i].offset = h_offset[i];
    a.item[i].count = h_count[i];
    a.item[i].aux = h_aux ? h_aux[i] : 0;
    a.item[i].kind = k;
    if (h_count[i] > most) most = h_count[i];
  }
  a.nitems = n_items;
  a.stream = stream_buf;
  int gx = cdiv(most, 256 * 4);
  if (gx > 1024) gx = 1024;
  if (gx < 1) gx = 1;
  pack_stream_kernel<<<dim3(gx, n_items), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  LAUNCH_CHECK();
  return 0;
}

int sirenb200_eval_metrics(const float* pred, const float* img, int64_t n, float* out,
                           sirenb200_stream_t stream) {
  if (!pred || !img || !out || n <= 0) return fail(SIRENB200_ERR_INVALID, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* acc = nullptr;
  keep_pool_memory();
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&acc), 2 * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(acc, 0, 2 * sizeof(double), st));
  int grid = cdiv(n, 256 * 8);
  if (grid > 1184) grid = 1184;
  eval_metrics_kernel<<<grid, 256, 0, st>>>(pred, img, n, acc);
  LAUNCH_CHECK();
  eval_metrics_finish_kernel<<<1, 1, 0, st>>>(acc, n, out);
  LAUNCH_CHECK();
  CUDA_TRY(cudaFreeAsync(acc, st));
  return 0;
}

int sirenb200_adam_step(int32_t nt, float* const* prm, float* const* grads, float* const* m,
                        float* const* v, const float* const* mask, const int64_t* numel, float lr,
                        float beta1, float beta2, float eps, int32_t step, float inv_scale,
                        const float* skip_flag, int32_t zero_grad, sirenb200_stream_t stream) {
  if (nt < 1 || nt > kMaxTensors) return fail(SIRENB200_ERR_INVALID, "tensor count %d", nt);
  if (!prm || !grads || !m || !v || !numel || step < 1)
    return fail(SIRENB200_ERR_INVALID, "bad argument");
  AdamArgs a{};
  int chunks = 0;
  for (int i = 0; i < nt; ++i) {
    a.p[i] = prm[i];
    a.g[i] = grads[i];
    a.m[i] = m[i];
    a.v[i] = v[i];
    a.mask[i] = mask ? mask[i] : nullptr;
    if (numel[i] < 0 || numel[i] > 0x7fffffff) return fail(SIRENB200_ERR_INVALID, "numel");
    a.n[i] = int(numel[i]);
    a.chunk_begin[i] = chunks;
    chunks += cdiv(numel[i], 1024);
  }
  a.chunk_begin[nt] = chunks;
  a.ntensors = nt;
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = eps;
  a.omb1 = float(1.0 - double(beta1));
  a.omb2 = float(1.0 - double(beta2));
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  a.step_size = float(double(lr) / bc1);
  a.bc2_sqrt = float(sqrt(bc2));
  a.inv_scale = inv_scale;
  a.skip_flag = skip_flag;
  a.zero_grad = zero_grad;
  a.dev_sched = nullptr;
  if (chunks == 0) return 0;
  adam_multi_kernel<<<chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  LAUNCH_CHECK();
  return 0;
}

int sirenb200_sched_step(double* state, const float* stats, float inv_count, float* loss_ring,
                         int32_t ring_len, float* loss_host, sirenb200_stream_t stream) {
  if (!state || !stats) return fail(SIRENB200_ERR_INVALID, "null argument");
  SchedArgs a{state, stats, inv_count, loss_ring, ring_len > 0 ? ring_len : 1, loss_host};
  sched_step_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  LAUNCH_CHECK();
  return 0;
}

int sirenb200_adam_step_dev(int32_t nt, float* const* prm, float* const* grads, float* const* m,
                            float* const* v, const float* const* mask, const int64_t* numel,
                            float beta1, float beta2, float eps, const double* sched_state,
                            float inv_scale, const float* skip_flag, int32_t zero_grad,
                            sirenb200_stream_t stream) {
  if (nt < 1 || nt > kMaxTensors) return fail(SIRENB200_ERR_INVALID, "tensor count %d", nt);
  if (!prm || !grads || !m || !v || !numel || !sched_state)
    return fail(SIRENB200_ERR_INVALID, "bad argument");
  AdamArgs a{};
  int chunks = 0;
  for (int i = 0; i < nt; ++i) {
    a.p[i] = prm[i];
    a.g[i] = grads[i];
    a.m[i] = m[i];
    a.v[i] = v[i];
    a.mask[i] = mask ? mask[i] : nullptr;
    if (numel[i] < 0 || numel[i] > 0x7fffffff) return fail(SIRENB200_ERR_INVALID, "numel");
    a.n[i] = int(numel[i]);
    a.chunk_begin[i] = chunks;
    chunks += cdiv(numel[i], 1024);
  }
  a.chunk_begin[nt] = chunks;
  a.ntensors = nt;
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = eps;
  a.omb1 = float(1.0 - double(beta1));
  a.omb2 = float(1.0 - double(beta2));
  a.inv_scale = inv_scale;
  a.skip_flag = skip_flag;
  a.zero_grad = zero_grad;
  a.dev_sched = sched_state + 6;  // [6] step_size, [7] bc2_sqrt
  if (chunks == 0) return 0;
  adam_multi_kernel<<<chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  LAUNCH_CHECK();
  return 0;
}

int sirenb200_apply_mask(float* w, const float* mask, int64_t n, sirenb200_stream_t stream) {
  if (!w || !mask || n < 0) return fail(SIRENB200_ERR_INVALID, "bad argument");
  if (n == 0) return 0;
  apply_mask_kernel<<<cdiv(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w, mask, n);
  LAUNCH_CHECK();
  return 0;
}

int sirenb200_prune_threshold_search(const float* sorted_mags, int64_t n, int64_t nonzero_total, int64_t tokill,
                                     double tolerance, double* state, double* result, sirenb200_stream_t stream) {
  if (!sorted_mags || n < 1 || !state || !result || tokill < 1 || !(tolerance >= 0.0))
    return fail(SIRENB200_ERR_INVALID, "prune_threshold_search: bad argument");
  prune_threshold_search_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      sorted_mags, static_cast<long long>(n), static_cast<long long>(nonzero_total), static_cast<long long>(tokill),
      tolerance, state, result);
  LAUNCH_CHECK();
  return 0;
}

int sirenb200_fakequant_per_channel(const float* w, int32_t rows, int32_t cols,
                                    const float* row_min, const float* row_max, float neg_div,
                                    float pos_div, int8_t* codes, float* scales, float* w_out,
                                    sirenb200_stream_t stream) {
  if (!w || rows < 1 || cols < 1 || !(neg_div > 0.f) || !(pos_div > 0.f))
    return fail(SIRENB200_ERR_INVALID, "bad argument");
  fakequant_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w, rows, cols, row_min, row_max, neg_div, pos_div, codes, scales, w_out);
  LAUNCH_CHECK();
  return 0;
}

int sirenb200_kmeans_quantize(const float* w, int64_t n, int32_t bits, int32_t iter_limit, float tol,
                              const float* init_centers, float* centroids, int32_t* n_centroids, int64_t* labels, float* w_out,
                              sirenb200_stream_t stream) {
  if (!w || n < 1 || bits < 1 || bits > 10 || !centroids || !n_centroids)
    return fail(SIRENB200_ERR_INVALID, "bad argument (bits must be in [1, 10])");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int k = (1 << bits) - 1;
  // scratch: centres[k] | minmax[2] | shift | sums[k] | cnts[k] | k_cur | status | done | label16[npad]
  const int64_t npad = (n + 255) / 256 * 256;
  const size_t fbytes = (size_t(k) * 2 + 8) * sizeof(float);
  const size_t ibytes = (size_t(k) + 8) * sizeof(int);
  char* scratch = nullptr;
  keep_pool_memory();
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&scratch), fbytes + ibytes + 16 + size_t(npad) * 2, st));
  float* cent = reinterpret_cast<float*>(scratch);
  float* mm = cent + k;
  float* shift = mm + 2;
  float* sums = mm + 8;
  unsigned int* cnts = reinterpret_cast<unsigned int*>(scratch + fbytes);
  int* k_cur = reinterpret_cast<int*>(cnts + k);
  int* status = k_cur + 1;
  int* done = k_cur + 2;
  // 16-byte aligned: fbytes + ibytes = (3k + 16) * 4 with k odd -> round the label offset up
  uint16_t* label16 = reinterpret_cast<uint16_t*>(scratch + (fbytes + ibytes + 15) / 16 * 16);
  const float inf_init[2] = {INFINITY, -INFINITY};
  const int int_init[3] = {k, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(mm, inf_init, sizeof(inf_init), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(k_cur, int_init, sizeof(int_init), cudaMemcpyHostToDevice, st));
  int grid = cdiv(n, 256);
  if (grid > 592) grid = 592;
  int rc = 0;
  auto cleanup = [&](int code) {
    cudaFreeAsync(scratch, st);
    return code;
  };
  kmeans_minmax_kernel<<<grid, 256, 0, st>>>(w, n, mm);
  ++g_launches;
  if (init_centers) {
    CUDA_TRY(cudaMemcpyAsync(cent, init_centers, k * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    kmeans_linspace_kernel<<<cdiv(k, 256), 256, 0, st>>>(mm, k, cent);
    ++g_launches;
  }
  // The whole Lloyd loop is enqueued at once: the convergence test (center_shift ** 2 < tolerance) runs on
  // the device and turns the remaining iterations into no-ops.
  for (int it = 0; it < iter_limit; ++it) {
    kmeans_label_kernel<<<grid, 256, k * sizeof(float), st>>>(w, n, npad, cent, k_cur, done, label16);
    kmeans_cluster_sum_kernel<<<cdiv(k, 8), 256, 0, st>>>(w, npad, label16, k_cur, done, sums, cnts);
    KmeansUpdateArgs ua{cent, sums, cnts, k_cur, shift, status, done, tol};
    kmeans_update_kernel<<<1, 1024, 0, st>>>(ua);
    g_launches += 3;
  }
  kmeans_codebook_kernel<<<1, 1024, 0, st>>>(cent, k_cur, centroids, n_centroids);
  ++g_launches;
  kmeans_predict_kernel<<<grid, 256, (k + 1) * sizeof(float), st>>>(w, n, centroids, n_centroids,
                                                                   labels, w_out);
  ++g_launches;
  int hstat = 0;
  if (cudaMemcpyAsync(&hstat, status, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    rc = fail(SIRENB200_ERR_CUDA, "k-means failed: %s", cudaGetErrorString(cudaGetLastError()));
    return cleanup(rc);
  }
  if (hstat != 0)
    return cleanup(fail(SIRENB200_ERR_INVALID,
                        "k-means: top clusters are empty (the reference raises a shape mismatch here, "
                        "quant/kmeans_helper.py:91)"));
  return cleanup(0);
}

// ---------------------------------------------------------------------------------------
// gradient exchange over peer memory
// ---------------------------------------------------------------------------------------
}  // extern "C" (the comm struct is C++)

// Wall-clock bound of a rank's wait for its peers inside the exchange kernel: SIRENB200_EXCHANGE_TIMEOUT_S
// (default 120 s; 0 = wait forever).
static uint64_t exchange_timeout_ns() {
  static const uint64_t v = [] {
    const char* env = getenv("SIRENB200_EXCHANGE_TIMEOUT_S");
    const double s = env ? atof(env) : 120.0;
    return s <= 0.0 ? uint64_t(0) : uint64_t(s * 1e9);
  }();
  return v;
}

struct sirenb200_comm {
  int rank = 0, world = 1, device = 0;
  int64_t max_floats = 0;
  char* base = nullptr;          // local region: signals | epochs | data
  void* peer_base[kCommMaxRanks] = {};
  bool connected = false;
  CommPeers peers{};
  uint32_t* epoch_b = nullptr;
};
static constexpr size_t kCommSigBytes = size_t(2) * kCommBlocks * kCommMaxRanks * sizeof(uint32_t);
static constexpr size_t kCommEpochBytes = kCommBlocks * sizeof(uint32_t);
static constexpr size_t kCommHeader = (kCommSigBytes + kCommEpochBytes + 255) / 256 * 256;

extern "C" {

int sirenb200_comm_create(int32_t rank, int32_t world, int64_t max_floats, sirenb200_comm_t* out) {
  if (!out || world < 1 || world > kCommMaxRanks || rank < 0 || rank >= world || max_floats < 1)
    return fail(SIRENB200_ERR_INVALID, "comm_create: bad argument (1 <= world <= %d)", kCommMaxRanks);
  auto* c = new (std::nothrow) sirenb200_comm();
  if (!c) return fail(SIRENB200_ERR_CUDA, "out of host memory");
  c->rank = rank;
  c->world = world;
  c->max_floats = (max_floats + 3) / 4 * 4;
  cudaGetDevice(&c->device);
  const size_t bytes = kCommHeader + size_t(2) * c->max_floats * sizeof(float);
  if (cudaMalloc(reinterpret_cast<void**>(&c->base), bytes) != cudaSuccess ||
      cudaMemset(c->base, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    const int rc = fail(SIRENB200_ERR_CUDA, "comm_create: %s", cudaGetErrorString(cudaGetLastError()));
    if (c->base) cudaFree(c->base);
    delete c;
    return rc;
  }
  c->epoch_b = reinterpret_cast<uint32_t*>(c->base + kCommSigBytes);
  c->peer_base[rank] = c->base;
  if (world == 1) c->connected = true;
  *out = c;
  return 0;
}

int sirenb200_comm_handle(sirenb200_comm_t c, void* handle_out) {
  if (!c || !handle_out) return fail(SIRENB200_ERR_INVALID, "comm_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, c->base));
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

int sirenb200_comm_connect(sirenb200_comm_t c, const void* handles) {
  if (!c || !handles) return fail(SIRENB200_ERR_INVALID, "comm_connect: null argument");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + size_t(r) * sizeof(h), sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&c->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(SIRENB200_ERR_CUDA, "comm_connect: cannot map rank %d's buffer (%s) - is peer access available?",
                  r, cudaGetErrorString(e));
    }
  }
  for (int r = 0; r < c->world; ++r) {
    char* b = static_cast<char*>(c->peer_base[r]);
    c->peers.sig[r] = reinterpret_cast<uint32_t*>(b);
    c->peers.data[r] = reinterpret_cast<float*>(b + kCommHeader);
  }
  c->connected = true;
  return 0;
}

int sirenb200_comm_allreduce(sirenb200_comm_t c, float* data, int64_t n, sirenb200_stream_t stream) {
  if (!c || !data || n < 1) return fail(SIRENB200_ERR_INVALID, "comm_allreduce: bad argument");
  if (!c->connected) return fail(SIRENB200_ERR_STATE, "comm_allreduce before comm_connect");
  if (n > c->max_floats) return fail(SIRENB200_ERR_INVALID, "comm_allreduce: n exceeds max_floats");
  if (reinterpret_cast<uintptr_t>(data) & 15u) return fail(SIRENB200_ERR_INVALID, "comm_allreduce: data must be 16-byte aligned");
  if (c->world == 1) return 0;
  if (c->peers.data[c->rank] == nullptr) {  // world > 1 needs connect() even though peer_base[rank] is set
    return fail(SIRENB200_ERR_STATE, "comm_allreduce before comm_connect");
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  p2p_allreduce_kernel<<<kCommBlocks, kCommThreads, 0, st>>>(c->peers, c->epoch_b, data, n, c->max_floats, c->rank,
                                                    c->world, exchange_timeout_ns());
  LAUNCH_CHECK();
  return 0;
}

// One whole fit step.  Tensor-core handles: forward + MSE + backward leave their partials in the workspace and
// ONE kernel (step_end_kernel) reduces them, exchanges them over NVLink when fit->comm is set, finalises the loss,
// advances the schedule and applies Adam.  fp32 handles: the same sequence as separate launches.
int sirenb200_fit_step(sirenb200_handle_t h, const float* img, const sirenb200_fit_t* f, sirenb200_stream_t stream) {
  if (!h || !img || !f) return fail(SIRENB200_ERR_INVALID, "fit_step: bad argument");
  if (!f->h_params || !f->h_grads || !f->h_exp_avg || !f->h_exp_avg_sq || !f->h_numel || !f->sched_state ||
      !f->stats)
    return fail(SIRENB200_ERR_INVALID, "fit_step: null field");
  if (f->comm && (!f->flat || f->flat_n < 1)) return fail(SIRENB200_ERR_INVALID, "fit_step: comm without flat");
  const bool fused = h->cfg.precision == SIRENB200_PREC_F16TC && f->n_tensors == 2 * h->D &&
                     !(getenv("SIRENB200_STEP_END") && atoi(getenv("SIRENB200_STEP_END")) == 0);
  if (!fused) {
    int rc = sirenb200_forward_backward(h, f->h_params, img, 1.0f, f->h_grads, f->stats, stream);
    if (!rc && f->comm) rc = sirenb200_comm_allreduce(f->comm, f->flat, f->flat_n, stream);
    if (!rc)
      rc = sirenb200_sched_step(f->sched_state, f->stats, f->comm ? f->inv_count : 0.0f, f->loss_ring,
                                f->ring_len, f->loss_host, stream);
    if (!rc)
      rc = sirenb200_adam_step_dev(f->n_tensors, f->h_params, f->h_grads, f->h_exp_avg, f->h_exp_avg_sq, f->h_mask,
                                   f->h_numel, f->beta1, f->beta2, f->eps, f->sched_state, 1.0f, f->stats + 2, 0,
                                   stream);
    return rc;
  }
  int rc = check_ready(h);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  sirenb200_comm_t c = f->comm;
  const bool comm = c && c->world > 1;
  if (comm) {
    if (!c->connected || c->peers.data[c->rank] == nullptr) return fail(SIRENB200_ERR_STATE, "fit_step: comm not connected");
    if (f->flat_n > c->max_floats) return fail(SIRENB200_ERR_INVALID, "fit_step: flat buffer exceeds the comm region");
    if (reinterpret_cast<uintptr_t>(f->flat) & 15u) return fail(SIRENB200_ERR_INVALID, "fit_step: flat must be 16-byte aligned");
    if (f->stats < f->flat || f->stats + 4 > f->flat + f->flat_n)
      return fail(SIRENB200_ERR_INVALID, "fit_step: stats must be 4 floats inside the flat buffer");
  }
  h->defer_reduce = true;
  rc = tc_dispatch(h, f->h_params, 1, img, nullptr, f->h_grads, h->inv_count, f->stats, st);
  h->defer_reduce = false;
  h->have_fwd = (rc == 0);
  if (rc) return rc;
  StepEndArgs a{};
  tc_build_reduce(h, f->h_grads, h->inv_count, f->stats, h->nchunks, a.red);
  for (int i = 0; i < f->n_tensors; ++i) {
    if (f->h_numel[i] != a.red.d[i].n) return fail(SIRENB200_ERR_INVALID, "fit_step: tensor %d has %lld elements, the model has %d", i, (long long)f->h_numel[i], a.red.d[i].n);
    a.p[i] = f->h_params[i];
    a.m[i] = f->h_exp_avg[i];
    a.v[i] = f->h_exp_avg_sq[i];
    a.mask[i] = f->h_mask ? f->h_mask[i] : nullptr;
    if (comm) {
      const int64_t off = f->h_grads[i] - f->flat;
      if (off < 0 || off + a.red.d[i].n > f->flat_n || (a.red.d[i].vec && (off & 3)))
        return fail(SIRENB200_ERR_INVALID, "fit_step: gradient %d is not an (aligned) view of the flat buffer", i);
      a.flat_off[i] = off;
    }
  }
  a.beta2 = f->beta2;
  a.eps = f->eps;
  a.omb1 = float(1.0 - double(f->beta1));
  a.omb2 = float(1.0 - double(f->beta2));
  a.loss_partial = h->last_part + h->C * h->W + h->C;
  a.loss_nparts = h->last_grid * h->nchunks;
  a.loss_stride = int64_t(h->C) * h->W + h->C + 1;
  a.inv_count = h->inv_count;
  a.gstate = h->gstate;
  a.sched = f->sched_state;
  a.loss_ring = f->loss_ring;
  a.ring_len = f->ring_len > 0 ? f->ring_len : 1;
  a.loss_host = f->loss_host;
  a.bar = h->bar;
  a.world = 1;
  int grid = h->nsm < a.red.chunk_begin[a.red.ndesc] ? h->nsm : a.red.chunk_begin[a.red.ndesc];
  if (comm) {
    a.cp = c->peers;
    a.epoch_b = c->epoch_b;
    a.rank = c->rank;
    a.world = c->world;
    a.max_floats = c->max_floats;
    a.stats_off = f->stats - f->flat;
    a.timeout_ns = exchange_timeout_ns();
    grid = kCommBlocks;  // every rank runs the same blocks: block b pairs with block b of each peer
  }
  {
    ProfScope ps(h, PK_REDUCE, st);
    step_end_kernel<<<grid, kStepEndThreads, 0, st>>>(a);
  }
  LAUNCH_CHECK();
  return 0;
}

int sirenb200_fit_steps(sirenb200_handle_t h, int32_t k, const float* img, const sirenb200_fit_t* f,
                        sirenb200_stream_t stream) {
  if (k < 0) return fail(SIRENB200_ERR_INVALID, "fit_steps: bad argument");
  for (int i = 0; i < k; ++i) {
    int rc = sirenb200_fit_step(h, img, f, stream);
    if (rc) return rc;
  }
  return 0;
}

int sirenb200_comm_destroy(sirenb200_comm_t c) {
  if (!c) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->base) cudaFree(c->base);
  delete c;
  return 0;
}

}  // extern "C"
