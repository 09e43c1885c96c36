// CUDA-core kernels: the fp32 path (any shape), and the bandwidth-bound pieces shared with the
// tensor-core path (first layer with in-kernel coordinates, last layer + MSE loss, layer-0 gradient,
// partial reductions, multi-tensor Adam, masks, k-means and int8 fake quantisation).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tc_kernels.cuh"

namespace sb {

constexpr int kMaxLayers = 16;
constexpr int kMaxTensors = 2 * kMaxLayers;
constexpr int kMaxOut = 4;

// ------------------------------------------------------------------------------------------
// generic fp32 tiled GEMM (64x64x16 tiles, 4x4 per thread) with fused epilogues
// ------------------------------------------------------------------------------------------
enum SimtOp {
  OP_NT_SINE = 0,  // C = A[M,K] * B[N,K]^T + bias ; Z <- C, Out <- sin(omega*C)     (forward)
  OP_NT_LIN = 1,   // C = A * B^T + bias ; Out <- C                                   (last layer)
  OP_NN_DCOS = 2,  // C = A[M,K] * B[K,N] ; Out <- C * omega*cos(omega*Z)              (dX + dsine)
  OP_TN_PART = 3,  // C = A[K,M]^T * B[K,N] over k in this split ; Out[split] <- C     (dW partial)
  OP_NT_RELU = 4,  // C = A * B^T + bias ; Z <- C, Out <- max(C, 0)                     (FourierNet forward, fourier.py:44-52)
  OP_NN_DRELU = 5  // C = A[M,K] * B[K,N] ; Out <- C * [Z > 0]                          (dX + dReLU)
};

struct SimtGemmArgs {
  const float* A;
  const float* B;
  const float* bias;
  float* Z;    // pre-activation (written by NT_SINE, read by NN_DCOS)
  float* Out;
  float* ColSum;  // TN_PART: per-split column sums of A (bias gradient) or null
  int M, N, K;
  int lda, ldb, ldo;
  float omega;
  int ksplit_len;  // TN_PART: rows of K per split (gridDim.z splits)
};

template <int OP>
__global__ void __launch_bounds__(256) simt_gemm_kernel(const SimtGemmArgs a) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  int k_begin = 0, k_end = a.K;
  if (OP == OP_TN_PART) {
    k_begin = blockIdx.z * a.ksplit_len;
    k_end = min(a.K, k_begin + a.ksplit_len);
  }
  float acc[4][4] = {};
  float csum[4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    // load A tile -> As[k][m], B tile -> Bs[k][n]
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      int kk, mm;
      float v = 0.f;
      if (OP == OP_TN_PART) {  // A is [K, M] row-major
        kk = i >> 6;
        mm = i & 63;
        if (k0 + kk < k_end && m0 + mm < a.M) v = a.A[int64_t(k0 + kk) * a.lda + m0 + mm];
      } else {  // A is [M, K] row-major
        mm = i >> 4;
        kk = i & 15;
        if (k0 + kk < k_end && m0 + mm < a.M) v = a.A[int64_t(m0 + mm) * a.lda + k0 + kk];
      }
      As[kk][mm] = v;
    }
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      int kk, nn;
      float v = 0.f;
      if (OP == OP_NT_SINE || OP == OP_NT_LIN || OP == OP_NT_RELU) {  // B is [N, K]
        nn = i >> 4;
        kk = i & 15;
        if (k0 + kk < k_end && n0 + nn < a.N) v = a.B[int64_t(n0 + nn) * a.ldb + k0 + kk];
      } else {  // B is [K, N]
        kk = i >> 6;
        nn = i & 63;
        if (k0 + kk < k_end && n0 + nn < a.N) v = a.B[int64_t(k0 + kk) * a.ldb + n0 + nn];
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (OP == OP_TN_PART) {
#pragma unroll
        for (int i = 0; i < 4; ++i) csum[i] += av[i];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float c = acc[i][j];
      if (OP == OP_NT_SINE) {
        c += a.bias[n];
        a.Z[int64_t(m) * a.ldo + n] = c;
        a.Out[int64_t(m) * a.ldo + n] = sinf(c * a.omega);
      } else if (OP == OP_NT_LIN) {
        a.Out[int64_t(m) * a.ldo + n] = c + a.bias[n];
      } else if (OP == OP_NN_DCOS) {
        const float z = a.Z[int64_t(m) * a.ldo + n];
        a.Out[int64_t(m) * a.ldo + n] = c * (a.omega * cosf(z * a.omega));
      } else if (OP == OP_NT_RELU) {
        c += a.bias[n];
        a.Z[int64_t(m) * a.ldo + n] = c;
        a.Out[int64_t(m) * a.ldo + n] = fmaxf(c, 0.f);
      } else if (OP == OP_NN_DRELU) {
        a.Out[int64_t(m) * a.ldo + n] = a.Z[int64_t(m) * a.ldo + n] > 0.f ? c : 0.f;
      } else {
        a.Out[(int64_t(blockIdx.z) * a.M + m) * a.ldo + n] = c;
      }
    }
    if (OP == OP_TN_PART && a.ColSum && blockIdx.x == 0 && tx == 0)
      a.ColSum[int64_t(blockIdx.z) * a.M + m] = csum[i];
  }
}

// FourierNet's input encoding (fourier.py:20-25): xp = (2 pi x) @ B, features = [sin(xp) | cos(xp)], x = the RAW grid
// coordinates in [0, 1] (h, w).  enc: [npix, 2 * half] fp32.
__global__ void fourier_encode_kernel(CoordSrc c, const float* B, int half, float* enc, int64_t npix) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= npix * half) return;
  const int64_t p = idx / half;
  const int j = int(idx - p * half);
  float gh, gw;
  const int64_t pp = p + c.p_offset;
  if (c.coords) {
    const float2 v = reinterpret_cast<const float2*>(c.coords)[pp];
    gh = v.x;
    gw = v.y;
  } else {
    const unsigned pu = unsigned(pp);
    const int r = int(pu / unsigned(c.width)), col = int(pu - unsigned(r) * unsigned(c.width));
    gh = __ldg(c.lin_h + c.row_begin + r);
    gw = __ldg(c.lin_w + col);
  }
  const float two_pi = 6.283185307179586f;
  const float xp = fmaf(__fmul_rn(two_pi, gw), B[half + j], __fmul_rn(__fmul_rn(two_pi, gh), B[j]));
  enc[p * (2 * half) + j] = sinf(xp);
  enc[p * (2 * half) + half + j] = cosf(xp);
}

// materialise x = (grid - 0.5) * 2 as [npix, 2] fp32 (fp32 path only)
__global__ void simt_coords_kernel(CoordSrc c, float* x, int64_t npix) {
  const int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (p >= npix) return;
  float xh, xw;
  load_xy(c, p, xh, xw);
  reinterpret_cast<float2*>(x)[p] = make_float2(xh, xw);
}

// fp32 path: pred / loss / dL/dy from the last layer's output y [npix, C].
//   mode 0: forward only (pred).  mode 1: MSE loss, g = (pred-img) (seed, unscaled).
//   mode 2: g = 0.5 * dpred.
// With outermost_linear == 0 the last layer is a sine layer: y holds z, pred = sin(w z)/2+.5.
struct LossArgs {
  const float* y;
  const float* img;    // mode 1: target; mode 2: dpred
  float* pred;         // may be null
  float* g;            // dL/dz of the last layer (seed units), may be null in mode 0
  float* loss_partial; // [gridDim.x]
  int64_t n;           // npix * C
  int mode;
  int outermost_linear;
  float omega;
  int out_kind;        // 0: pred = o / 2 + 0.5 (siren.py:131); 1: pred = sigmoid(y) (fourier.py:55-56)
};
__global__ void __launch_bounds__(256) simt_loss_kernel(const LossArgs a) {
  float lsum = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < a.n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float z = a.y[i];
    if (a.out_kind == 1) {
      const float pred = 1.0f / (1.0f + expf(-z));
      if (a.pred) a.pred[i] = pred;
      if (a.mode != 0) {
        float g;
        if (a.mode == 1) {
          const float d = pred - a.img[i];
          lsum += d * d;
          g = 2.0f * d;  // the caller's scale is 1 / (H W C): d(mean sq err)/d(pred) = 2 d / (H W C)
        } else {
          g = a.img[i];
        }
        a.g[i] = g * pred * (1.0f - pred);
      }
      continue;
    }
    const float o = a.outermost_linear ? z : sinf(z * a.omega);
    const float pred = o / 2 + 0.5f;
    if (a.pred) a.pred[i] = pred;
    if (a.mode != 0) {
      float g;
      if (a.mode == 1) {
        const float d = pred - a.img[i];
        lsum += d * d;
        g = d;
      } else {
        g = 0.5f * a.img[i];
      }
      if (!a.outermost_linear) g *= a.omega * cosf(z * a.omega);
      a.g[i] = g;
    }
  }
  __shared__ float red[256];
  red[threadIdx.x] = lsum;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && a.loss_partial) a.loss_partial[blockIdx.x] = red[0];
}

// ------------------------------------------------------------------------------------------
// tensor-core path, layer 0: a0 = sin(w0 * (W0 x + b0)) with in-kernel coordinates -> signed-half
// (siren.py:62,66 for the is_first layer; K = 2 is not a tensor-core shape)
// ------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) tc_first_layer_kernel(CoordSrc c, const float* __restrict__ w0,
                                                             const float* __restrict__ b0,
                                                             float omega, __half* __restrict__ act0,
                                                             int64_t npix, int64_t npix_pad) {
  // each thread: 8 consecutive columns of one pixel
  constexpr int TPR = W / 8;         // threads per row
  constexpr int RPB = 256 / TPR;     // rows per block iteration
  const int tc = threadIdx.x % TPR, tr = threadIdx.x / TPR;
  float wa[8], wb[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    wa[j] = w0[(tc * 8 + j) * 2 + 0] * omega;
    wb[j] = w0[(tc * 8 + j) * 2 + 1] * omega;
    bb[j] = b0[tc * 8 + j] * omega;
  }
  // row / column of the pixel are tracked incrementally (one division per thread, not per pixel)
  const int64_t pfirst = int64_t(blockIdx.x) * RPB + tr;
  const unsigned step = gridDim.x * RPB;
  const unsigned step_r = step / unsigned(c.width), step_c = step % unsigned(c.width);
  const uint64_t g0 = uint64_t(pfirst + c.p_offset);
  unsigned row = unsigned(g0 / unsigned(c.width)), col = unsigned(g0 % unsigned(c.width));
  for (int64_t p = pfirst; p < npix_pad; p += step) {
    uint32_t o[4] = {0, 0, 0, 0};
    if (p < npix) {
      float xh, xw;
      if (c.coords) {
        load_xy(c, p, xh, xw);
      } else {
        xh = (__ldg(c.lin_h + c.row_begin + row) - 0.5f) * 2.0f;
        xw = (__ldg(c.lin_w + col) - 0.5f) * 2.0f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float t0 = fmaf(xh, wa[2 * j], fmaf(xw, wb[2 * j], bb[2 * j]));
        const float t1 = fmaf(xh, wa[2 * j + 1], fmaf(xw, wb[2 * j + 1], bb[2 * j + 1]));
        o[j] = sine_signed_half2(t0, t1);
      }
    }
    row += step_r;
    col += step_c;
    if (col >= unsigned(c.width)) {
      col -= unsigned(c.width);
      ++row;
    }
    reinterpret_cast<uint4*>(act0 + p * W)[tc] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------
// tensor-core path, last layer (W -> C <= 4, CUDA cores) fused with y/2+0.5, the MSE loss, the
// gradient seed, dA = g W_last, dZ of the last hidden layer (cos rebuilt from the signed-half)
// and the last layer's weight / bias gradients.  One warp per pixel row.
//   (siren.py:64,110-118,131; train_helper.py:151-154 F.mse_loss and its backward)
// ------------------------------------------------------------------------------------------
struct LastArgs {
  const __half* act;     // [npix_pad, W] signed-half, last hidden activation
  __half* dz;            // [npix_pad, W] out (mode != 0)
  const float* w;        // [C, W]
  const float* b;        // [C]
  const float* img;      // mode 1: target [npix, C]; mode 2: dpred [npix, C]
  float* pred;           // [npix, C] or null
  float* part;           // per-block partials: [grid][C*W + C + 1] (dW, db, sum sq err)
  const float* gscale;   // device scalar: seed scale G
  int64_t npix, npix_pad;
  int C;
  int mode;              // 0 forward only, 1 MSE, 2 external dpred
  int outermost_linear;
  float omega_last;      // omega of the last layer (only when it is a sine layer)
  float omega_prev;      // omega of the layer that produced `act`
};

// CT = compile-time bound on the output channels (3 for RGB; 4 = generic up to CT)
template <int W, int CT>
__global__ void __launch_bounds__(256) tc_last_layer_kernel(const LastArgs a) {
  constexpr int NCH = (W + 255) / 256;            // 8-column chunks per lane
  constexpr int ACTIVE = (W / 8) < 32 ? (W / 8) : 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool on = lane < ACTIVE;
  float wr[CT][NCH * 8];
  float dwacc[CT][NCH * 8];
  float dbacc[CT] = {};
  float lsum = 0.f;
#pragma unroll
  for (int cc = 0; cc < CT; ++cc)
#pragma unroll
    for (int j = 0; j < NCH * 8; ++j) {
      const int col = (j / 8) * 256 + lane * 8 + (j % 8);
      wr[cc][j] = (on && cc < a.C) ? a.w[cc * W + col] : 0.f;
      dwacc[cc][j] = 0.f;
    }
  const float G = (a.mode != 0) ? *a.gscale : 1.f;
  const int64_t nwarps = int64_t(gridDim.x) * 8;
  // software pipeline: the activation rows of the next PF iterations are already in flight
  constexpr int PF = 3;
  uint4 q[PF][NCH];
  float qi[PF];  // lane cc < C prefetches img[row, cc] (target or upstream gradient)
  auto fetch = [&](int64_t row, uint4 (&dst)[NCH], float& dimg) {
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      dst[ch] = make_uint4(0, 0, 0, 0);
      if (on && row < a.npix) dst[ch] = reinterpret_cast<const uint4*>(a.act + row * W)[ch * 32 + lane];
    }
    dimg = 0.f;
    if (a.mode != 0 && lane < a.C && row < a.npix) dimg = a.img[row * a.C + lane];
  };
  const int64_t pstart = int64_t(blockIdx.x) * 8 + warp;
#pragma unroll
  for (int i = 0; i < PF; ++i) fetch(pstart + i * nwarps, q[i], qi[i]);
  for (int64_t p = pstart; p < a.npix_pad; p += nwarps) {
    uint32_t raw[NCH * 4];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      raw[ch * 4 + 0] = q[0][ch].x;
      raw[ch * 4 + 1] = q[0][ch].y;
      raw[ch * 4 + 2] = q[0][ch].z;
      raw[ch * 4 + 3] = q[0][ch].w;
    }
    const float img_lane = qi[0];
#pragma unroll
    for (int i = 0; i + 1 < PF; ++i) {
      qi[i] = qi[i + 1];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) q[i][ch] = q[i + 1][ch];
    }
    fetch(p + PF * nwarps, q[PF - 1], qi[PF - 1]);
    if (p >= a.npix) {  // padding rows: zero gradient so the dW reduction ignores them
      if (a.mode != 0 && on)
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
          reinterpret_cast<uint4*>(a.dz + p * W)[ch * 32 + lane] = make_uint4(0, 0, 0, 0);
      continue;
    }
    float av[NCH * 8];
#pragma unroll
    for (int j = 0; j < NCH * 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&raw[j]));
      av[2 * j] = f.x;
      av[2 * j + 1] = f.y;
    }
    float y[CT];
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NCH * 8; ++j) s = fmaf(av[j], wr[cc][j], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      y[cc] = s;
    }
    float g[CT];
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) {
      g[cc] = 0.f;
      if (cc < a.C) {
        const float z = y[cc] + a.b[cc];
        const float o = a.outermost_linear ? z : sinf(z * a.omega_last);
        const float pred = o / 2 + 0.5f;
        if (a.pred && lane == cc) a.pred[p * a.C + cc] = pred;
        const float tgt = __shfl_sync(0xffffffffu, img_lane, cc);
        if (a.mode == 1) {
          const float d = pred - tgt;
          if (lane == 0) lsum += d * d;
          g[cc] = d * G;
        } else if (a.mode == 2) {
          g[cc] = 0.5f * tgt * G;
        }
        if (!a.outermost_linear) g[cc] *= a.omega_last * cosf(z * a.omega_last);
      }
    }
    if (a.mode != 0) {
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) {
        if (lane == 0) dbacc[cc] += g[cc];
#pragma unroll
        for (int j = 0; j < NCH * 8; ++j) dwacc[cc][j] = fmaf(g[cc], av[j], dwacc[cc][j]);
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        uint32_t o4[4];
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          float d0 = 0.f, d1 = 0.f;
#pragma unroll
          for (int cc = 0; cc < CT; ++cc) {
            d0 = fmaf(g[cc], wr[cc][ch * 8 + 2 * j2], d0);
            d1 = fmaf(g[cc], wr[cc][ch * 8 + 2 * j2 + 1], d1);
          }
          const uint32_t r = raw[ch * 4 + j2];
          d0 *= a.omega_prev * cos_from_signed_half(r & 0xFFFFu);
          d1 *= a.omega_prev * cos_from_signed_half(r >> 16);
          o4[j2] = pack_f16x2(d0, d1);
        }
        if (on)
          reinterpret_cast<uint4*>(a.dz + p * W)[ch * 32 + lane] =
              make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
  if (a.mode == 0) return;
  // block reduction of the per-warp partial sums (8 warps) through shared memory
  __shared__ float red[8][CT + 1];
  __shared__ float buf[8][32][CT];
  float* out = a.part + int64_t(blockIdx.x) * (a.C * W + a.C + 1);
#pragma unroll
  for (int j = 0; j < NCH * 8; ++j) {
    __syncthreads();
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) buf[warp][lane][cc] = dwacc[cc][j];
    __syncthreads();
    if (warp == 0 && on) {
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) {
        if (cc < a.C) {
          float s = 0.f;
          for (int w8 = 0; w8 < 8; ++w8) s += buf[w8][lane][cc];
          const int col = (j / 8) * 256 + lane * 8 + (j % 8);
          out[cc * W + col] = s;
        }
      }
    }
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) red[warp][cc] = dbacc[cc];
    red[warp][CT] = lsum;
  }
  __syncthreads();
  if (threadIdx.x <= a.C) {
    const int idx = threadIdx.x < a.C ? threadIdx.x : CT;
    float s = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][idx];
    out[a.C * W + threadIdx.x] = s;  // db[0..C-1], then sum of squared error
  }
}

// ------------------------------------------------------------------------------------------
// tensor-core path, layer-0 gradients: dW0[j, {h,w}] = sum_p dz0[p, j] * x[p], db0[j] = sum_p dz0
// ------------------------------------------------------------------------------------------
// Streaming version: each block owns a contiguous pixel range and pulls it through a ring of
// shared-memory stages with 1-D bulk copies (dozens of KB in flight per SM without register cost).
template <int W>
struct L0GradCfg {
  static constexpr int ROWS = 32;                       // pixels per stage
  static constexpr int STAGES = 6;
  static constexpr uint32_t STAGE_BYTES = ROWS * W * 2;
  static constexpr uint32_t OFF_XY = STAGES * STAGE_BYTES + STAGES * 8 + 16;  // float2 xy[2][ROWS]
  static constexpr uint32_t SMEM_BYTES = OFF_XY + 2 * ROWS * 8;
};

template <int W>
__global__ void __launch_bounds__(256) tc_layer0_grad_kernel(CoordSrc c, const __half* __restrict__ dz0,
                                                             float* __restrict__ part, int64_t npix) {
  using Cfg = L0GradCfg<W>;
  constexpr int TPR = W / 8;      // threads per row (8 columns = 16 bytes each)
  constexpr int RL = 256 / TPR;   // row lanes
  extern __shared__ __align__(128) uint8_t l0_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(l0_smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  const int tc = threadIdx.x % TPR, tr = threadIdx.x / TPR;
  const int64_t per_block = ((npix + gridDim.x - 1) / gridDim.x + Cfg::ROWS - 1) / Cfg::ROWS * Cfg::ROWS;
  const int64_t p0 = blockIdx.x * per_block;
  const int64_t p1 = min(npix, p0 + per_block);
  const int niter = p1 > p0 ? int((p1 - p0 + Cfg::ROWS - 1) / Cfg::ROWS) : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int it) {
    const int s = it % Cfg::STAGES;
    const int64_t q0 = p0 + int64_t(it) * Cfg::ROWS;
    const int64_t left = p1 - q0;
    const int rows = left < Cfg::ROWS ? int(left) : Cfg::ROWS;
    mbar_expect_tx(&full[s], uint32_t(rows) * W * 2);
    bulk_load_1d(l0_smem + s * Cfg::STAGE_BYTES, dz0 + q0 * W, uint32_t(rows) * W * 2, &full[s]);
  };
  if (threadIdx.x == 0)
    for (int it = 0; it < Cfg::STAGES && it < niter; ++it) issue(it);
  // coordinates of a stage's rows are computed once (32 threads) one iteration ahead
  float2* xy = reinterpret_cast<float2*>(l0_smem + Cfg::OFF_XY);
  auto stage_coords = [&](int it) {
    if (threadIdx.x < Cfg::ROWS && it < niter) {
      const int64_t q = p0 + int64_t(it) * Cfg::ROWS + threadIdx.x;
      float xh = 0.f, xw = 0.f;
      if (q < p1) load_xy(c, q, xh, xw);
      xy[(it & 1) * Cfg::ROWS + threadIdx.x] = make_float2(xh, xw);
    }
  };
  stage_coords(0);
  __syncthreads();
  float ah[8] = {}, aw[8] = {}, ab[8] = {};
  for (int it = 0; it < niter; ++it) {
    const int s = it % Cfg::STAGES;
    stage_coords(it + 1);
    mbar_wait(&full[s], (it / Cfg::STAGES) & 1);
    const int64_t q0 = p0 + int64_t(it) * Cfg::ROWS;
    const uint8_t* st = l0_smem + s * Cfg::STAGE_BYTES;
#pragma unroll
    for (int r = tr; r < Cfg::ROWS; r += RL) {
      if (q0 + r < p1) {
        const float2 cxy = xy[(it & 1) * Cfg::ROWS + r];
        const float xh = cxy.x, xw = cxy.y;
        const uint4 v = *reinterpret_cast<const uint4*>(st + (r * W + tc * 8) * 2);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 g = __half22float2(*reinterpret_cast<const __half2*>(&w4[j]));
          ah[2 * j] = fmaf(g.x, xh, ah[2 * j]);
          aw[2 * j] = fmaf(g.x, xw, aw[2 * j]);
          ab[2 * j] += g.x;
          ah[2 * j + 1] = fmaf(g.y, xh, ah[2 * j + 1]);
          aw[2 * j + 1] = fmaf(g.y, xw, aw[2 * j + 1]);
          ab[2 * j + 1] += g.y;
        }
      }
    }
    __syncthreads();  // everyone is done with stage s
    if (threadIdx.x == 0 && it + Cfg::STAGES < niter) issue(it + Cfg::STAGES);
  }
  // block reduction over the row lanes, reusing stage memory: red[RL][3][W]
  float* red = reinterpret_cast<float*>(l0_smem);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(tr * 3 + 0) * W + tc * 8 + j] = ah[j];
    red[(tr * 3 + 1) * W + tc * 8 + j] = aw[j];
    red[(tr * 3 + 2) * W + tc * 8 + j] = ab[j];
  }
  __syncthreads();
  float* out = part + int64_t(blockIdx.x) * (3 * W);
  for (int i = threadIdx.x; i < 3 * W; i += 256) {
    const int k = i / W, col = i % W;
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < RL; ++r) sum += red[(r * 3 + k) * W + col];
    if (k < 2)
      out[col * 2 + k] = sum;  // dW0[col, {h, w}]
    else
      out[2 * W + col] = sum;  // db0[col]
  }
}

// ------------------------------------------------------------------------------------------
// weight staging for the tensor-core path ("weight load"): fp32 master -> fp16 operands
//   wh[l]  = fp16(W_l)                [out, in]   forward B operand (K-major)
//   wth[l] = fp16(omega_{l-1} W_l^T)  [in, out]   dX B operand (omega of the cos factor folded in)
// ------------------------------------------------------------------------------------------
struct PrepArgs {
  const float* w[kMaxLayers];
  float omega_prev[kMaxLayers];
  int nlayers;  // hidden layers staged: l = 1 .. nlayers
  int W;
  __half* wh;   // [nlayers][W][W]
  __half* wth;  // [nlayers][W][W]
  float* stats; // zeroed here (start of a step); may be null
  // epilogue constants of the fused forward kernel
  const float* bias[kMaxLayers];  // hidden-layer biases
  const float* w0;                // layer 0 weight [W, 2] and bias [W]
  const float* b0;
  float omega0, omega_h;
  // operands of the tensor-core last layer
  const float* w_last;            // [C, W]
  int C;
  float omega_prev_last;          // omega of the layer feeding the last layer
  __half* wl16;                   // [16, W]: rows < C = W_last, rest 0
  __half* wlt16;                  // [W x 16] in UMMA no-swizzle K-major core-matrix order:
                                  // element (n, k) at (n%8)*8 + (n/8)*128 + (k/8)*64 + (k%8) halves,
                                  // value omega_prev * W_last[k, n] for k < C, else 0
  float4* tab0;                   // [W] (omega0*w_h, omega0*w_w, omega0*b0, 0)
  float* bias_w;                  // [nlayers][W] omega_h * bias
  float* bias_raw;                // [nlayers][W] bias
  unsigned int* pace;             // [2 * kMaxLayers] pace counters of the merged backward launches (zeroed here)
  int Wm;                         // the MODEL's hidden width (<= W): parameters are [Wm, *]; everything staged here is
                                  // zero-padded to the kernel width W (siren.py:88 makes widths like 114)
  float* w0p;                     // [W, 2] layer-0 weight, padded
  float* b0p;                     // [W]    layer-0 bias, padded
};
__global__ void __launch_bounds__(256) tc_prep_weights_kernel(const PrepArgs a) {
  __shared__ float tile[32][33];
  if (a.stats && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x < 4)
    a.stats[threadIdx.x] = 0.f;
  if (a.pace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x < 2 * kMaxLayers)
    a.pace[threadIdx.x] = 0u;
  const int l = blockIdx.z;
  const float* w = a.w[l];
  const int W = a.W, Wm = a.Wm;
  if (blockIdx.x == 1 % gridDim.x && blockIdx.y == 0 && l == 0 && a.wl16) {
    for (int i = threadIdx.x; i < 16 * W; i += 256) {
      const int c = i / W, n = i % W;
      a.wl16[i] = __float2half_rn((c < a.C && n < Wm) ? a.w_last[c * Wm + n] : 0.f);
    }
    for (int i = threadIdx.x; i < W * 16; i += 256) {
      const int n = i / 16, c = i % 16;
      const int off = (n % 8) * 8 + (n / 8) * 128 + (c / 8) * 64 + (c % 8);
      a.wlt16[off] = __float2half_rn((c < a.C && n < Wm) ? a.omega_prev_last * a.w_last[c * Wm + n] : 0.f);
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int i = threadIdx.x; i < W; i += 256) {
      const float bv = i < Wm ? a.bias[l][i] : 0.f;
      a.bias_w[l * W + i] = a.omega_h * bv;
      a.bias_raw[l * W + i] = bv;
      if (l == 0) {
        const float wh = i < Wm ? a.w0[2 * i] : 0.f, ww = i < Wm ? a.w0[2 * i + 1] : 0.f, b0 = i < Wm ? a.b0[i] : 0.f;
        a.tab0[i] = make_float4(a.omega0 * wh, a.omega0 * ww, a.omega0 * b0, 0.f);
        a.w0p[2 * i] = wh;
        a.w0p[2 * i + 1] = ww;
        a.b0p[i] = b0;
      }
    }
  }
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const float v = (by + r < Wm && bx + tx < Wm) ? w[(by + r) * Wm + bx + tx] : 0.f;
    tile[r][tx] = v;
    a.wh[(size_t(l) * W + by + r) * W + bx + tx] = __float2half_rn(v);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    a.wth[(size_t(l) * W + bx + r) * W + by + tx] = __float2half_rn(tile[tx][r] * a.omega_prev[l]);
}

// ------------------------------------------------------------------------------------------
// split-K / per-block partial reduction into the gradient tensors (+ non-finite detection)
// ------------------------------------------------------------------------------------------
struct ReduceDesc {
  float* dst;
  const float* src;
  int n;             // elements
  int nsplit;
  int64_t split_stride;
  int vec;           // 1: few splits of a large tensor -> a thread owns 4 consecutive elements (n % 4 == 0)
  int cols;          // > 0: the partial slab is a [rows, cols_pad] matrix of which [rows, cols] are gradient elements
  int cols_pad;
};
struct ReduceArgs {
  ReduceDesc d[kMaxTensors];
  int chunk_begin[kMaxTensors + 1];  // prefix sums of ceil(n / 32) (vec: ceil(n / 1024))
  int ndesc;
  float scale;               // host-known factor
  const float* gscale;       // device seed scale G (grads are divided by it) or null
  float* stats;              // stats[2] <- 1 if any non-finite
};
__device__ __forceinline__ int64_t reduce_src_index(const ReduceDesc& d, int e) {
  // gradient element e of a [rows, cols] tensor lives at (e / cols) * cols_pad + e % cols of the (width-padded)
  // partial slab; cols == 0: same layout
  return d.cols ? int64_t(e / d.cols) * d.cols_pad + (e % d.cols) : int64_t(e);
}
__global__ void __launch_bounds__(256) reduce_partials_kernel(const ReduceArgs a) {
  // block = 32 consecutive elements x 8 split lanes (one warp per lane: 128-byte coalesced rows)
  int t = 0;
  while (t + 1 < a.ndesc && int(blockIdx.x) >= a.chunk_begin[t + 1]) ++t;
  const ReduceDesc d = a.d[t];
  const float scale = a.gscale ? a.scale / *a.gscale : a.scale;
  if (d.vec) {
    // split-K partials of a weight gradient: every thread sums its float4 over the splits in order,
    // all loads independent
    const int e4 = (blockIdx.x - a.chunk_begin[t]) * 1024 + threadIdx.x * 4;
    bool bad4 = false;
    if (e4 < d.n) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* src = d.src + reduce_src_index(d, e4);
#pragma unroll 8
      for (int sp = 0; sp < d.nsplit; ++sp) {
        const float4 v = *reinterpret_cast<const float4*>(src + int64_t(sp) * d.split_stride);
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
      acc.x *= scale;
      acc.y *= scale;
      acc.z *= scale;
      acc.w *= scale;
      *reinterpret_cast<float4*>(d.dst + e4) = acc;
      bad4 = !(isfinite(acc.x) && isfinite(acc.y) && isfinite(acc.z) && isfinite(acc.w));
    }
    if (__syncthreads_or(bad4) && threadIdx.x == 0) a.stats[2] = 1.0f;
    return;
  }
  const int e = (blockIdx.x - a.chunk_begin[t]) * 32 + (threadIdx.x & 31);
  const int lane = threadIdx.x >> 5;
  float s = 0.f;
  if (e < d.n) {
    const float* src = d.src + reduce_src_index(d, e);
    int sp = lane;
    for (; sp + 24 < d.nsplit; sp += 32) {  // 4 independent loads in flight
      const float v0 = src[int64_t(sp) * d.split_stride];
      const float v1 = src[int64_t(sp + 8) * d.split_stride];
      const float v2 = src[int64_t(sp + 16) * d.split_stride];
      const float v3 = src[int64_t(sp + 24) * d.split_stride];
      s += (v0 + v1) + (v2 + v3);
    }
    for (; sp < d.nsplit; sp += 8) s += src[int64_t(sp) * d.split_stride];
  }
  __shared__ float red[8][32];
  red[lane][threadIdx.x & 31] = s;
  __syncthreads();
  bool bad = false;
  if (lane == 0 && e < d.n) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += red[k][threadIdx.x];
    r *= scale;
    d.dst[e] = r;
    bad = !isfinite(r);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) a.stats[2] = 1.0f;
}

// loss finalisation + seed-scale policy (one block)
struct FinalizeArgs {
  const float* loss_partial;
  int nparts;
  int64_t part_stride;  // distance between consecutive partial values
  float inv_count;      // 1 / (H*W*C)
  float* stats;         // [0] sum sq err, [1] loss, [2] non-finite flag
  float* gstate;        // [0] G, [1] G cap  (tensor-core path) or null
  float* loss_hist;     // optional ring of per-step losses
  const int* step;      // device step counter (index into loss_hist) or null
  int hist_len;
};
__global__ void __launch_bounds__(256) finalize_loss_kernel(const FinalizeArgs a) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < a.nparts; i += 256) s += a.loss_partial[i * a.part_stride];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float sse = red[0];
    const float loss = sse * a.inv_count;
    a.stats[0] = sse;
    a.stats[1] = loss;
    if (a.loss_hist && a.step) a.loss_hist[*a.step % a.hist_len] = loss;
    if (a.gstate) {
      float cap = a.gstate[1];
      if (a.stats[2] != 0.f) cap = fmaxf(1.f, a.gstate[0] * (1.f / 16.f));
      // next seed scale: power of two near 0.125 / rmse, clamped to [1, cap]
      float g = 1.f;
      if (loss > 0.f && isfinite(loss)) g = exp2f(floorf(log2f(0.125f * rsqrtf(loss))));
      g = fminf(fmaxf(g, 1.f), cap);
      a.gstate[0] = g;
      a.gstate[1] = cap;
    }
  }
}

// max |x| -> power-of-two seed scale for an external dpred (one block, grid-stride)
__global__ void __launch_bounds__(1024) absmax_scale_kernel(const float* x, int64_t n, float* gstate) {
  __shared__ float red[1024];
  float m = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) m = fmaxf(m, fabsf(x[i]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int k = 512; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + k]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mx = red[0];
    float g = 1.f;
    if (mx > 0.f && isfinite(mx)) g = exp2f(floorf(log2f(4.0f / mx)));
    gstate[0] = fminf(fmaxf(g, 1.f), 1.0e30f);
  }
}

// ------------------------------------------------------------------------------------------
// eval_epoch metrics (train_helper.py:48-57): mse and mse of the (x*255).int() images
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eval_metrics_kernel(const float* pred, const float* img,
                                                           int64_t n, double* acc) {
  double s = 0, s8 = 0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n;
       i += int64_t(gridDim.x) * blockDim.x) {
    const float p = pred[i], t = img[i];
    const float d = p - t;
    s += double(d) * d;
    const int d8 = int(t * 255.0f) - int(p * 255.0f);  // .int() truncates toward zero
    s8 += double(d8 * d8);
  }
  __shared__ double r0[256], r1[256];
  r0[threadIdx.x] = s;
  r1[threadIdx.x] = s8;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) {
      r0[threadIdx.x] += r0[threadIdx.x + k];
      r1[threadIdx.x] += r1[threadIdx.x + k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(&acc[0], r0[0]);
    atomicAdd(&acc[1], r1[0]);
  }
}
__global__ void eval_metrics_finish_kernel(const double* acc, int64_t n, float* out) {
  out[0] = float(acc[0] / double(n));
  out[1] = float(acc[1] / double(n));
}

// ------------------------------------------------------------------------------------------
// multi-tensor Adam (+ unscale, skip-on-inf, mask, zero-grad)
// ------------------------------------------------------------------------------------------
struct AdamArgs {
  float* p[kMaxTensors];
  float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  const float* mask[kMaxTensors];
  int n[kMaxTensors];
  int chunk_begin[kMaxTensors + 1];
  int ntensors;
  float beta1, beta2, eps;
  float omb1, omb2;     // float(1 - beta) evaluated in double on the host, as torch does
  float step_size;     // lr / (1 - beta1^t)
  float bc2_sqrt;      // sqrt(1 - beta2^t)
  float inv_scale;
  const float* skip_flag;
  int zero_grad;
  // device-driven schedule (graph replay): when non-null, step_size / bc2_sqrt are read from here
  const double* dev_sched;  // [0] step_size, [1] bc2_sqrt
};
__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamArgs a) {
  int t = 0;
  while (t + 1 < a.ntensors && int(blockIdx.x) >= a.chunk_begin[t + 1]) ++t;
  const int base = (blockIdx.x - a.chunk_begin[t]) * 1024;
  const bool skip = a.skip_flag && (*a.skip_flag != 0.f);
  const float step_size = a.dev_sched ? float(a.dev_sched[0]) : a.step_size;
  const float bc2_sqrt = a.dev_sched ? float(a.dev_sched[1]) : a.bc2_sqrt;
  float* P = a.p[t];
  float* G = a.g[t];
  float* M = a.m[t];
  float* V = a.v[t];
  const float* K = a.mask[t];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = base + k * 256 + threadIdx.x;
    if (i >= a.n[t]) continue;
    float p = P[i];
    if (!skip) {
      const float g = G[i] * a.inv_scale;
      float m = M[i], v = V[i];
      m = m + (g - m) * a.omb1;                                 // exp_avg.lerp_(grad, 1-beta1)
      v = v * a.beta2 + (a.omb2 * g) * g;                     // mul_(beta2).addcmul_(g, g, 1-beta2)
      const float denom = sqrtf(v) / bc2_sqrt + a.eps;
      p = p + (-step_size * m) / denom;                      // addcdiv_(m, denom, -step_size)
      M[i] = m;
      V[i] = v;
    }
    if (K) p = p * K[i];                                     // Masking.apply_mask
    P[i] = p;
    if (a.zero_grad) G[i] = 0.f;
  }
}

// Device-side optimiser schedule (CUDA-graph replay: nothing that changes per step may be a launch
// argument).  One thread: computes this step's Adam step size / bias correction from the device step
// counter, records the loss of the step that just ran, advances the counter.
//   state (doubles, so the host's double-precision schedule arithmetic is reproduced exactly):
//   [0] step, [1] lr0, [2] gamma, [3] period, [4] beta1, [5] beta2, [6] out step_size, [7] out bc2_sqrt
struct SchedArgs {
  double* state;
  const float* stats;   // [0] sum sq err, [1] loss
  float inv_count;      // >0: loss = stats[0] * inv_count (pixel-sharded fits), else stats[1]
  float* loss_ring;
  int ring_len;
  float* loss_host;     // optional: host-mapped (pinned) float[2]: {loss, number of steps completed}
};
__global__ void sched_step_kernel(const SchedArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int step = int(a.state[0]);  // 0-based index of the optimiser step about to run
  // (the eager path hands lr to the library as a C float: reproduce that rounding)
  const double lr = double(float(a.state[1] * pow(a.state[2], double(step / int(a.state[3])))));
  const double bc1 = 1.0 - pow(double(float(a.state[4])), double(step + 1));  // betas arrive as C floats
  const double bc2 = 1.0 - pow(double(float(a.state[5])), double(step + 1));  // on the eager path too
  a.state[6] = lr / bc1;
  a.state[7] = sqrt(bc2);
  const float loss = a.inv_count > 0.f ? a.stats[0] * a.inv_count : a.stats[1];
  if (a.loss_ring) a.loss_ring[step % a.ring_len] = loss;
  if (a.loss_host) {
    // the host may poll [1] instead of synchronising the stream: loss first, then the step count
    a.loss_host[0] = loss;
    __threadfence_system();
    a.loss_host[1] = float(step + 1);
  }
  a.state[0] = double(step + 1);
}

__global__ void apply_mask_kernel(float* w, const float* mask, int64_t n) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i < n) w[i] = w[i] * mask[i];
}

// ------------------------------------------------------------------------------------------
// int8 per-channel symmetric fake quantisation (QAT weights)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fakequant_rows_kernel(const float* w, int rows, int cols,
                                                             const float* row_min,
                                                             const float* row_max, float neg_div,
                                                             float pos_div, int8_t* codes,
                                                             float* scales, float* w_out) {
  const int r = blockIdx.x;
  __shared__ float rlo[256], rhi[256];
  float lo, hi;
  if (row_min && row_max) {
    lo = row_min[r];
    hi = row_max[r];
  } else {
    lo = INFINITY;
    hi = -INFINITY;
    for (int c = threadIdx.x; c < cols; c += 256) {
      const float x = w[int64_t(r) * cols + c];
      lo = fminf(lo, x);
      hi = fmaxf(hi, x);
    }
    rlo[threadIdx.x] = lo;
    rhi[threadIdx.x] = hi;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
      if (threadIdx.x < k) {
        rlo[threadIdx.x] = fminf(rlo[threadIdx.x], rlo[threadIdx.x + k]);
        rhi[threadIdx.x] = fmaxf(rhi[threadIdx.x], rhi[threadIdx.x + k]);
      }
      __syncthreads();
    }
    lo = rlo[0];
    hi = rhi[0];
  }
  // symmetric qparams: scale = max(-min(lo,0)/neg_div, max(hi,0)/pos_div), clamped to eps
  const float s_neg = -fminf(lo, 0.f) / neg_div, s_pos = fmaxf(hi, 0.f) / pos_div;
  const float scale = fmaxf(fmaxf(s_neg, s_pos), 1.1920928955078125e-07f);
  if (threadIdx.x == 0 && scales) scales[r] = scale;
  // torch's fake-quant kernel multiplies by the reciprocal scale: nearbyint(x * (1/scale))
  const float inv_scale = 1.0f / scale;
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float x = w[int64_t(r) * cols + c];
    float q = nearbyintf(x * inv_scale);  // round half to even
    q = fminf(fmaxf(q, -128.f), 127.f);
    if (codes) codes[int64_t(r) * cols + c] = int8_t(q);
    if (w_out) w_out[int64_t(r) * cols + c] = q * scale;
  }
}

// ------------------------------------------------------------------------------------------
// 1-D k-means (Deep-Compression weight sharing, quant/kmeans.py + kmeans_helper.py)
// ------------------------------------------------------------------------------------------
// first-min argmin of (x - c_k)^2 over k (torch.argmin returns the first minimum)
__device__ __forceinline__ int kmeans_nearest(float x, const float* cent, int k) {
  float best = INFINITY;
  int bi = 0;
  for (int j = 0; j < k; ++j) {
    const float d = x - cent[j];
    const float dd = d * d;  // (a - b) ** 2.0
    if (dd < best) {
      best = dd;
      bi = j;
    }
  }
  return bi;
}

// min / max / count of non-zero weights
__global__ void __launch_bounds__(256) kmeans_minmax_kernel(const float* w, int64_t n, float* mm) {
  __shared__ float smin[256], smax[256];
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = blockIdx.x * int64_t(256) + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) {
    const float x = w[i];
    if (x != 0.f) {
      lo = fminf(lo, x);
      hi = fmaxf(hi, x);
    }
  }
  smin[threadIdx.x] = lo;
  smax[threadIdx.x] = hi;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) {
      smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + k]);
      smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + k]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // float atomics via int punning (values may be negative): use CAS loops
    int* imin = reinterpret_cast<int*>(&mm[0]);
    int* imax = reinterpret_cast<int*>(&mm[1]);
    int old = *imin, assumed;
    do {
      assumed = old;
      if (__int_as_float(assumed) <= smin[0]) break;
      old = atomicCAS(imin, assumed, __float_as_int(smin[0]));
    } while (old != assumed);
    old = *imax;
    do {
      assumed = old;
      if (__int_as_float(assumed) >= smax[0]) break;
      old = atomicCAS(imax, assumed, __float_as_int(smax[0]));
    } while (old != assumed);
  }
}

// Lloyd assignment step for the non-zero weights: label16[i] = nearest centre, 0xFFFF for zeros and for the
// padding up to a multiple of 256 entries.  `done` (device flag) turns the remaining iterations into no-ops
// once the centre shift has converged, so the host never synchronises inside the Lloyd loop.
__global__ void __launch_bounds__(256) kmeans_label_kernel(const float* w, int64_t n, int64_t npad,
                                                           const float* cent, const int* k_cur,
                                                           const int* done, uint16_t* label16) {
  extern __shared__ float sc[];
  if (*done) return;
  const int k = *k_cur;
  for (int j = threadIdx.x; j < k; j += 256) sc[j] = cent[j];
  __syncthreads();
  for (int64_t i = blockIdx.x * int64_t(256) + threadIdx.x; i < npad; i += int64_t(gridDim.x) * 256) {
    const float x = (i < n) ? w[i] : 0.f;
    label16[i] = (x == 0.f) ? uint16_t(0xFFFF) : uint16_t(kmeans_nearest(x, sc, k));
  }
}

// Per-cluster sum and count, one warp per cluster, accumulating in INDEX ORDER in fp32 so the result
// is bit-identical to the reference's sequential index_add_ (scatter_mean restatement) on the CPU.
// A lane scans 8 labels per 16-byte load (256 labels per warp iteration, next load in flight); only
// the lanes that hold a hit fetch the weight.
__global__ void __launch_bounds__(256) kmeans_cluster_sum_kernel(const float* w, int64_t npad,
                                                                 const uint16_t* label16,
                                                                 const int* k_cur, const int* done,
                                                                 float* sums, unsigned int* cnts) {
  if (*done) return;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= *k_cur) return;
  const uint32_t pat = uint32_t(c) | (uint32_t(c) << 16);
  const uint4* lv = reinterpret_cast<const uint4*>(label16);
  const int64_t ngroups = npad / 8;  // npad is a multiple of 256 -> ngroups a multiple of 32
  float sum = 0.f;
  unsigned int cnt = 0;
  uint4 cur = lv[lane];
  for (int64_t g0 = 0; g0 < ngroups; g0 += 32) {
    uint4 nxt = cur;
    if (g0 + 32 < ngroups) nxt = lv[g0 + 32 + lane];
    const uint32_t e[4] = {cur.x ^ pat, cur.y ^ pat, cur.z ^ pat, cur.w ^ pat};
    unsigned int m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      m |= ((e[q] & 0xFFFFu) == 0u ? 1u : 0u) << (2 * q);
      m |= ((e[q] >> 16) == 0u ? 1u : 0u) << (2 * q + 1);
    }
    unsigned int bal = __ballot_sync(0xffffffffu, m != 0u);
    while (bal) {
      const int b = __ffs(bal) - 1;
      bal &= bal - 1;
      unsigned int mb = __shfl_sync(0xffffffffu, m, b);
      cnt += __popc(mb);
      while (mb) {
        const int j = __ffs(mb) - 1;
        mb &= mb - 1;
        float x = 0.f;
        if (lane == b) x = w[(g0 + b) * 8 + j];
        sum += __shfl_sync(0xffffffffu, x, b);
      }
    }
    cur = nxt;
  }
  if (lane == 0) {
    sums[c] = sum;
    cnts[c] = cnt;
  }
}

// one block: new centres (scatter_mean semantics), centre shift, convergence flag
struct KmeansUpdateArgs {
  float* cent;        // [k] in/out
  float* sums;
  unsigned int* cnts;
  int* k_cur;         // current number of centres
  float* shift;       // out: (sum_i |c_i - c'_i|)
  int* status;        // out: 1 = shape mismatch (reference would raise)
  int* done;          // in/out: set once shift**2 < tol (kmeans_helper.py:99-100) or on a shape mismatch
  float tol;
};
__global__ void __launch_bounds__(1024) kmeans_update_kernel(const KmeansUpdateArgs a) {
  __shared__ float red[1024];
  __shared__ int smax;
  if (*a.done) return;
  const int k = *a.k_cur;
  if (threadIdx.x == 0) smax = -1;
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += 1024)
    if (a.cnts[j] > 0) atomicMax(&smax, j);
  __syncthreads();
  const int nout = smax + 1;  // scatter_mean output length = labels.max() + 1
  float sh = 0.f;
  if (nout == k) {
    for (int j = threadIdx.x; j < k; j += 1024) {
      const unsigned int c = a.cnts[j];
      const float nc = a.sums[j] / float(c > 0 ? c : 1u);
      const float d = a.cent[j] - nc;
      sh += sqrtf(d * d);
      a.cent[j] = nc;
    }
  }
  red[threadIdx.x] = sh;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *a.shift = red[0];
    if (nout != k) {
      *a.status = 1;
      *a.done = 1;
    } else if (red[0] * red[0] < a.tol) {
      *a.done = 1;
    }
  }
}

// one block: prepend 0, unique, sort by |c| (stable w.r.t. the ascending unique order)
__global__ void __launch_bounds__(1024) kmeans_codebook_kernel(const float* cent, const int* k_cur,
                                                               float* codebook, int* n_codes) {
  __shared__ float v[1024];
  __shared__ float u[1024];
  __shared__ int nu;
  const int k = *k_cur + 1;  // with the prepended zero
  for (int j = threadIdx.x; j < 1024; j += 1024) v[j] = (j == 0) ? 0.f : (j < k ? cent[j - 1] : INFINITY);
  __syncthreads();
  // rank sort ascending (k <= 1024): position = #smaller + #equal-before
  float mine = v[threadIdx.x];
  int pos = 0;
  if (threadIdx.x < k) {
    for (int j = 0; j < k; ++j) {
      const float o = v[j];
      pos += (o < mine) || (o == mine && j < int(threadIdx.x));
    }
  }
  __syncthreads();
  if (threadIdx.x < k) u[pos] = mine;
  __syncthreads();
  // unique (sorted): keep first of each run; 0.0 == -0.0 collapses like torch.unique
  if (threadIdx.x == 0) {
    int m = 0;
    for (int j = 0; j < k; ++j)
      if (j == 0 || u[j] != u[j - 1]) v[m++] = u[j];
    nu = m;
  }
  __syncthreads();
  const int m = nu;
  // stable sort by |c|
  if (threadIdx.x < m) {
    mine = v[threadIdx.x];
    const float am = fabsf(mine);
    pos = 0;
    for (int j = 0; j < m; ++j) {
      const float ao = fabsf(v[j]);
      pos += (ao < am) || (ao == am && j < int(threadIdx.x));
    }
    codebook[pos] = mine;
  }
  if (threadIdx.x == 0) *n_codes = m;
}

// assign ALL weights (zeros included) to the codebook and rebuild the weights
__global__ void __launch_bounds__(256) kmeans_predict_kernel(const float* w, int64_t n,
                                                             const float* codebook,
                                                             const int* n_codes, int64_t* labels,
                                                             float* w_out) {
  extern __shared__ float sc[];
  const int k = *n_codes;
  for (int j = threadIdx.x; j < k; j += 256) sc[j] = codebook[j];
  __syncthreads();
  for (int64_t i = blockIdx.x * int64_t(256) + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) {
    const int b = kmeans_nearest(w[i], sc, k);
    if (labels) labels[i] = b;
    if (w_out) w_out[i] = sc[b];
  }
}

__global__ void kmeans_linspace_kernel(const float* mm, int k, float* cent) {
  // torch.linspace(lo, hi, k): symmetric evaluation, step = (hi - lo) / (k - 1)
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const float lo = mm[0], hi = mm[1];
  const float step = (hi - lo) / float(k - 1);
  cent[j] = (j < k / 2) ? lo + step * float(j) : hi - step * float(k - 1 - j);
}

// ------------------------------------------------------------------------------------------
// QAT activation fake-quant: torch's FusedMovingAvgObsFakeQuantize on an nn.Linear output (what
// torch.quantization.prepare_qat attaches for quant/context.py:35-47; per-tensor affine quint8, reduce_range ->
// [0, 127], MovingAverageMinMaxObserver with averaging constant 0.01).  state (device, 4 floats):
// {running min, running max, scale, zero point}; running min/max start at +inf / -inf.
//   1. actq_minmax_kernel   : per-block min / max of the tensor
//   2. actq_update_kernel   : observer update (training) + ChooseQuantizationParams (fbgemm's rule, double
//                             arithmetic, exactly aten/native/quantized/cpu/fused_obs_fake_quant.cpp)
//   3. actq_apply_kernel    : q = nearbyint(x * (1/scale)) + zp; out = (clamp(q) - zp) * scale; mask = q in range
//                             (the straight-through gradient mask), optionally a = sin(omega * out)
// ------------------------------------------------------------------------------------------
constexpr int kActqBlocks = 256;
__global__ void __launch_bounds__(256) actq_minmax_kernel(const float* x, int64_t n, float* partial) {
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float v = x[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  __shared__ float slo[256], shi[256];
  slo[threadIdx.x] = lo;
  shi[threadIdx.x] = hi;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) {
      slo[threadIdx.x] = fminf(slo[threadIdx.x], slo[threadIdx.x + k]);
      shi[threadIdx.x] = fmaxf(shi[threadIdx.x], shi[threadIdx.x + k]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = slo[0];
    partial[2 * blockIdx.x + 1] = shi[0];
  }
}
__global__ void actq_update_kernel(const float* partial, int nblocks, float* state, int training, float avg_const,
                                   int qmin, int qmax) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float rmin = state[0], rmax = state[1];
  if (training) {
    float cmin = INFINITY, cmax = -INFINITY;
    for (int i = 0; i < nblocks; ++i) {
      cmin = fminf(cmin, partial[2 * i]);
      cmax = fmaxf(cmax, partial[2 * i + 1]);
    }
    if (isinf(rmin) || isinf(rmax)) {
      rmin = cmin;
      rmax = cmax;
    } else {
      rmin = __fadd_rn(rmin, __fmul_rn(avg_const, __fsub_rn(cmin, rmin)));
      rmax = __fadd_rn(rmax, __fmul_rn(avg_const, __fsub_rn(cmax, rmax)));
    }
    state[0] = rmin;
    state[1] = rmax;
  }
  // quant_utils::ChooseQuantizationParams(min, max, qmin, qmax)
  double mn = fmin(double(rmin), 0.0), mx = fmax(double(rmax), 0.0);
  double scale = (mx - mn) / double(qmax - qmin);
  if (float(scale) == 0.0f || isinf(1.0f / float(scale))) scale = 0.1;
  const double kSmall = 6.1e-5;
  if (scale < kSmall) {
    const float org = float(scale);
    scale = kSmall;
    if (mn == 0.0) {
      mx = kSmall * (qmax - qmin);
    } else if (mx == 0.0) {
      mn = -kSmall * (qmax - qmin);
    } else {
      const float amp = float(kSmall / org);
      mn *= amp;
      mx *= amp;
    }
  }
  const double zmin = qmin - mn / scale, zmax = qmax - mx / scale;
  const double emin = fabs(double(qmin)) - fabs(mn / scale), emax = fabs(double(qmax)) - fabs(mx / scale);
  const double z0 = emin < emax ? zmin : zmax;
  int zp;
  if (z0 < qmin) zp = qmin;
  else if (z0 > qmax) zp = qmax;
  else zp = int(nearbyint(z0));
  state[2] = float(scale);
  state[3] = float(zp);
}
__global__ void __launch_bounds__(256) actq_apply_kernel(const float* x, int64_t n, const float* state, int qmin,
                                                         int qmax, float* out, unsigned char* mask, float* act,
                                                         float omega) {
  const float scale = state[2], zp = state[3];
  const float inv = __fdiv_rn(1.0f, scale);
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const float q = __fadd_rn(rintf(__fmul_rn(x[i], inv)), zp);
    const bool in = q >= float(qmin) && q <= float(qmax);
    const float qc = fminf(fmaxf(q, float(qmin)), float(qmax));
    const float v = __fmul_rn(__fsub_rn(qc, zp), scale);
    out[i] = v;
    if (mask) mask[i] = in ? 1 : 0;
    if (act) act[i] = sinf(v * omega);
  }
}
__global__ void __launch_bounds__(256) mul_mask_u8_kernel(float* g, const unsigned char* mask, int64_t n) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
    if (!mask[i]) g[i] = 0.f;
}

// ------------------------------------------------------------------------------------------
// Device-side packing for the entropy coder (pipeline/entropy_coding/__init__.py:15-41,70-120): the byte stream
// the reference builds on the host — `model.half()` tensors and uint8 k-means codes written back to back in
// state_dict order — is assembled by ONE kernel into one contiguous device buffer (one D2H copy feeds zstd), and
// taken apart again by one kernel (`decompress_state_dict`: fp16 -> fp32, weight = centroids[labels]).
// ------------------------------------------------------------------------------------------
constexpr int kPackMaxItems = 96;
enum PackKind {
  PACK_F32_TO_F16 = 0,   // float -> IEEE half (round to nearest even, as tensor.half())
  PACK_I64_TO_U8 = 1,    // int64 labels -> uint8
  PACK_I64_TO_U16 = 2,   // int64 labels -> uint16
  PACK_RAW = 3,          // byte copy (count = bytes)
  UNPACK_F16_TO_F32 = 4, // half -> float
  UNPACK_GATHER_U8 = 5,  // weight[i] = float(centroids_f16[codes_u8[i]])   (aux = stream offset of the centroids)
  UNPACK_GATHER_U16 = 6
};
struct PackItem {
  const void* src;     // pack: the tensor; unpack: unused
  void* dst;           // unpack: the destination tensor; pack: unused
  long long offset;    // byte offset inside the stream
  long long count;     // elements (PACK_RAW: bytes)
  long long aux;       // UNPACK_GATHER_*: byte offset of the fp16 code book inside the stream
  int kind;
  int pad_;
};
struct PackArgs {
  PackItem item[kPackMaxItems];
  int nitems;
  unsigned char* stream;
};
__global__ void __launch_bounds__(256) pack_stream_kernel(const PackArgs a) {
  const PackItem it = a.item[blockIdx.y];
  unsigned char* base = a.stream + it.offset;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < it.count; i += (long long)gridDim.x * blockDim.x) {
    switch (it.kind) {
      case PACK_F32_TO_F16:
        reinterpret_cast<__half*>(base)[i] = __float2half_rn(static_cast<const float*>(it.src)[i]);
        break;
      case PACK_I64_TO_U8:
        base[i] = static_cast<unsigned char>(static_cast<const long long*>(it.src)[i]);
        break;
      case PACK_I64_TO_U16:
        reinterpret_cast<unsigned short*>(base)[i] = static_cast<unsigned short>(static_cast<const long long*>(it.src)[i]);
        break;
      case PACK_RAW:
        base[i] = static_cast<const unsigned char*>(it.src)[i];
        break;
      case UNPACK_F16_TO_F32:
        static_cast<float*>(it.dst)[i] = __half2float(reinterpret_cast<const __half*>(base)[i]);
        break;
      case UNPACK_GATHER_U8:
        static_cast<float*>(it.dst)[i] = __half2float(reinterpret_cast<const __half*>(a.stream + it.aux)[base[i]]);
        break;
      case UNPACK_GATHER_U16:
        static_cast<float*>(it.dst)[i] =
            __half2float(reinterpret_cast<const __half*>(a.stream + it.aux)[reinterpret_cast<const unsigned short*>(base)[i]]);
        break;
    }
  }
}

// ------------------------------------------------------------------------------------------
// step_end_kernel: everything of a fit step that follows the last GEMM, in ONE launch.
//   phase 1  every block reduces ITS chunks of the split-K / per-CTA gradient partials (x scale / G);
//            block 0 also sums the squared-error partials (F.mse_loss, train_helper.py:151-154)
//   exchange (pixel-sharded fits, world > 1) the block's reduced chunks go to this rank's
//            peer-visible buffer, block b of every rank signals block b of every peer (system-scope flags), then
//            loads the same chunks from EVERY rank's buffer over NVLink and sums them in rank order — bit-identical
//            results on all ranks, so the replicated weights never drift.  No grid-wide step is needed for it.
//   barrier  one grid-wide arrive/wait (all blocks are co-resident: grid <= #SMs): publishes the non-finite flag
//            (GradScaler's found_inf) and the loss
//   phase 2  StepLR + Adam bias corrections from the device-side step counter (same double arithmetic as
//            sched_step_kernel), then torch.optim.Adam.step + Masking.apply_mask (train_helper.py:72-84,177;
//            core.py:272-279,687) on the block's own chunks — the gradients it just wrote.
// It replaces reduce_partials + finalize_loss + p2p_allreduce + sched_step + adam_multi (5 launches).
// ------------------------------------------------------------------------------------------
constexpr int kCommMaxRanks = 8;
constexpr int kCommBlocks = 144;   // one CTA per SM (<= 148): a 1 MB buffer moves in a single pass of float4s
constexpr int kCommThreads = 512;
struct CommPeers {
  float* data[kCommMaxRanks];      // rank r's data region: [2][max_floats]
  uint32_t* sig[kCommMaxRanks];    // rank r's signal region: [2][kCommBlocks][kCommMaxRanks]
};
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_volatile_f1(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
// Wait until *flag == epoch (written by a peer GPU) with exponential back-off; the bound is wall-clock
// (globaltimer): a rank may legitimately be seconds late (rank-0-only evaluation, first-use graph capture).
__device__ __forceinline__ void wait_peer_flag(const uint32_t* flag, uint32_t epoch, uint64_t timeout_ns, int rank,
                                               int b, int r) {
  uint32_t ns = 32;
  uint64_t t_start = 0;
  while (ld_acquire_sys(flag) != epoch) {
    __nanosleep(ns);
    if (ns < 4096) ns <<= 1;
    if (timeout_ns != 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t_start == 0) t_start = now;
      if (now - t_start > timeout_ns) {
        printf("sirenb200: gradient exchange timeout: rank %d block %d waiting for rank %d epoch %u\n", rank, b, r,
               epoch);
        __trap();
      }
    }
  }
}

struct StepEndArgs {
  ReduceArgs red;                     // partials -> gradient tensors (dst = the caller's h_grads[t])
  float* p[kMaxTensors];              // Adam tensors, aligned with red.d[]
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  const float* mask[kMaxTensors];
  int64_t flat_off[kMaxTensors];      // exchange: offset (floats) of tensor t's gradient inside the flat buffer
  float beta2, eps, omb1, omb2;
  const float* loss_partial;          // squared-error partials
  int loss_nparts;
  int64_t loss_stride;
  float inv_count;                    // 1 / (H*W*C) of the FULL image
  float* gstate;                      // [0] seed scale G, [1] cap (tensor-core path) or null
  double* sched;                      // 8 doubles, see sched_step_kernel
  float* loss_ring;
  int ring_len;
  float* loss_host;
  unsigned long long* bar;            // grid barrier counter (monotonic)
  CommPeers cp;                       // exchange (world > 1)
  uint32_t* epoch_b;
  int rank, world;
  int64_t max_floats;
  int64_t stats_off;                  // offset of the 4 stats floats inside the flat buffer
  unsigned long long timeout_ns;
};

// The Adam inputs of the block's first kStepEndPre chunks are loaded next to the partials (phase 1) and the whole
// update is computed while the grid barrier fills; what follows the barrier is the skip decision and the stores.
constexpr int kStepEndPre = 3;
struct StepEndPre {
  float g[4], p[4], m[4], v[4], k[4];
  int t, e, n;  // descriptor, first element, elements of this thread (0: nothing)
};

// 8 worker warps + 1 warp whose lane 0 evaluates the schedule (three double-precision pow: ~3 us on one thread,
// which used to sit between the grid barrier and the Adam arithmetic of every block)
constexpr int kStepEndThreads = 288;
__global__ void __launch_bounds__(kStepEndThreads) step_end_kernel(const StepEndArgs a) {
  __shared__ float red_s[8][32];
  __shared__ double sh_sched[2];
  const int b = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
  const bool worker = tid < 256;
  const ReduceArgs& r = a.red;
  const int nchunks = r.chunk_begin[r.ndesc];
  const float scale = r.gscale ? r.scale / *r.gscale : r.scale;  // the OLD seed scale: read before the barrier
  const int step = int(a.sched[0]);                              // 0-based index of this optimiser step
  const bool comm = a.world > 1;
  uint32_t epoch = 0, par = 0;
  float* mine = nullptr;
  if (comm) {
    epoch = a.epoch_b[b] + 1u;
    par = epoch & 1u;
    mine = a.cp.data[a.rank] + int64_t(par) * a.max_floats;
  }
  bool bad = false;
  if (tid == 256) {
    // schedule of the step about to be applied (sched_step_kernel's arithmetic, bit for bit); the spare
    // warp computes it while the workers reduce
    const double lr = double(float(a.sched[1] * pow(a.sched[2], double(step / int(a.sched[3])))));
    const double bc1 = 1.0 - pow(double(float(a.sched[4])), double(step + 1));
    const double bc2 = 1.0 - pow(double(float(a.sched[5])), double(step + 1));
    sh_sched[0] = lr / bc1;
    sh_sched[1] = sqrt(bc2);
  }

  // ---- squared error (block 0) ----
  float sse = 0.f;
  if (b == 0) {
    float s = 0.f;
    for (int i = tid; worker && i < a.loss_nparts; i += 256) s += a.loss_partial[i * a.loss_stride];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (worker && (tid & 31) == 0) red_s[0][tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      for (int k = 0; k < 8; ++k) sse += red_s[0][k];
    }
    __syncthreads();
  }

  StepEndPre pre[kStepEndPre];
#pragma unroll
  for (int i = 0; i < kStepEndPre; ++i) pre[i].n = 0;
  // Adam inputs of (descriptor t, elements e .. e + n): issued early, consumed after the gradients are final
  auto load_state = [&](StepEndPre& q, int t, int e, int n) {
    q.t = t;
    q.e = e;
    q.n = a.p[t] ? n : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < q.n) {
        q.p[j] = a.p[t][e + j];
        q.m[j] = a.m[t][e + j];
        q.v[j] = a.v[t][e + j];
        q.k[j] = a.mask[t] ? a.mask[t][e + j] : 1.0f;
      }
  };

  // ---- phase 1: reduce this block's chunks over the splits ----
  int t = 0;
  auto phase1_chunk = [&](int c, StepEndPre* q) {
    while (t + 1 < r.ndesc && c >= r.chunk_begin[t + 1]) ++t;
    const ReduceDesc d = r.d[t];
    float* out = comm ? mine + a.flat_off[t] : d.dst;
    if (d.vec) {
      const int e4 = (c - r.chunk_begin[t]) * 1024 + tid * 4;
      if (worker && e4 < d.n) {
        if (q) load_state(*q, t, e4, 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* src = d.src + reduce_src_index(d, e4);
#pragma unroll 8
        for (int sp = 0; sp < d.nsplit; ++sp) {
          const float4 v = *reinterpret_cast<const float4*>(src + int64_t(sp) * d.split_stride);
          acc.x += v.x;
          acc.y += v.y;
          acc.z += v.z;
          acc.w += v.w;
        }
        acc.x *= scale;
        acc.y *= scale;
        acc.z *= scale;
        acc.w *= scale;
        *reinterpret_cast<float4*>(out + e4) = acc;
        if (!comm) {
          bad |= !(isfinite(acc.x) && isfinite(acc.y) && isfinite(acc.z) && isfinite(acc.w));
          if (q) {
            q->g[0] = acc.x;
            q->g[1] = acc.y;
            q->g[2] = acc.z;
            q->g[3] = acc.w;
          }
        }
      }
    } else {
      const int e = (c - r.chunk_begin[t]) * 32 + (tid & 31);
      const int lane = tid >> 5;
      float s = 0.f;
      if (worker && e < d.n) {
        if (q && lane == 0) load_state(*q, t, e, 1);
        const float* src = d.src + reduce_src_index(d, e);
        int sp = lane;
        for (; sp + 24 < d.nsplit; sp += 32) {  // 4 independent loads in flight
          const float v0 = src[int64_t(sp) * d.split_stride];
          const float v1 = src[int64_t(sp + 8) * d.split_stride];
          const float v2 = src[int64_t(sp + 16) * d.split_stride];
          const float v3 = src[int64_t(sp + 24) * d.split_stride];
          s += (v0 + v1) + (v2 + v3);
        }
        for (; sp < d.nsplit; sp += 8) s += src[int64_t(sp) * d.split_stride];
      }
      if (worker) red_s[lane][tid & 31] = s;
      __syncthreads();
      if (lane == 0 && e < d.n) {
        float x = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) x += red_s[k][tid];
        x *= scale;
        out[e] = x;
        if (!comm) {
          bad |= !isfinite(x);
          if (q) q->g[0] = x;
        }
      }
      __syncthreads();
    }
  };
#pragma unroll
  for (int i = 0; i < kStepEndPre; ++i) {
    const int c = b + i * G;
    if (c < nchunks) phase1_chunk(c, &pre[i]);
  }
  for (int c = b + kStepEndPre * G; c < nchunks; c += G) phase1_chunk(c, nullptr);

  // ---- exchange over NVLink peer memory ----
  if (comm) {
    if (b == 0 && tid == 0) mine[a.stats_off] = sse;  // (the 4 stats floats need not be 16-byte aligned)
    __syncthreads();
    if (tid < a.world) {
      const int q = tid;
      const int64_t slot = (int64_t(par) * kCommBlocks + b) * kCommMaxRanks;
      __threadfence_system();
      st_release_sys(a.cp.sig[q] + slot + a.rank, epoch);
      wait_peer_flag(a.cp.sig[a.rank] + slot + q, epoch, a.timeout_ns, a.rank, b, q);
    }
    __syncthreads();
    const int64_t poff = int64_t(par) * a.max_floats;
    t = 0;
    auto exchange_chunk = [&](int c, StepEndPre* qp) {
      while (t + 1 < r.ndesc && c >= r.chunk_begin[t + 1]) ++t;
      const ReduceDesc d = r.d[t];
      if (d.vec) {
        const int e4 = (c - r.chunk_begin[t]) * 1024 + tid * 4;
        if (worker && e4 < d.n) {
          float4 v[kCommMaxRanks];
#pragma unroll
          for (int q = 0; q < kCommMaxRanks; ++q)
            if (q < a.world) v[q] = ld_volatile_f4(a.cp.data[q] + poff + a.flat_off[t] + e4);
          float4 acc = v[0];
#pragma unroll
          for (int q = 1; q < kCommMaxRanks; ++q)
            if (q < a.world) {
              acc.x += v[q].x;
              acc.y += v[q].y;
              acc.z += v[q].z;
              acc.w += v[q].w;
            }
          *reinterpret_cast<float4*>(d.dst + e4) = acc;
          bad |= !(isfinite(acc.x) && isfinite(acc.y) && isfinite(acc.z) && isfinite(acc.w));
          if (qp) {
            qp->g[0] = acc.x;
            qp->g[1] = acc.y;
            qp->g[2] = acc.z;
            qp->g[3] = acc.w;
          }
        }
      } else {
        const int e = (c - r.chunk_begin[t]) * 32 + tid;
        if (tid < 32 && e < d.n) {
          float v[kCommMaxRanks];
#pragma unroll
          for (int q = 0; q < kCommMaxRanks; ++q)
            if (q < a.world) v[q] = ld_volatile_f1(a.cp.data[q] + poff + a.flat_off[t] + e);
          float acc = v[0];
#pragma unroll
          for (int q = 1; q < kCommMaxRanks; ++q)
            if (q < a.world) acc += v[q];
          d.dst[e] = acc;
          bad |= !isfinite(acc);
          if (qp) qp->g[0] = acc;
        }
      }
    };
#pragma unroll
    for (int i = 0; i < kStepEndPre; ++i) {
      const int c = b + i * G;
      if (c < nchunks) exchange_chunk(c, &pre[i]);
    }
    for (int c = b + kStepEndPre * G; c < nchunks; c += G) exchange_chunk(c, nullptr);
    if (b == 0 && tid == 0) {
      sse = 0.f;
      for (int q = 0; q < a.world; ++q) sse += ld_volatile_f1(a.cp.data[q] + poff + a.stats_off);
    }
  }
  if (b == 0 && tid == 0) {
    r.stats[0] = sse;
    r.stats[1] = sse * a.inv_count;
  }
  if (__syncthreads_or(bad) && tid == 0) r.stats[2] = 1.0f;

  // ---- grid barrier: arrive, do the Adam arithmetic of the preloaded chunks, then wait ----
  unsigned long long target = 0;
  if (tid == 0) {
    __threadfence();
    const unsigned long long old = atomicAdd(a.bar, 1ull);
    target = (old / G + 1ull) * G;
  }
  const float step_size = float(sh_sched[0]), bc2_sqrt = float(sh_sched[1]);  // (written before the barrier above)
  // torch.optim.Adam.step + Masking.apply_mask (train_helper.py:72-84,177; core.py:272-279,687): pnew if the step
  // is taken, pskip (mask only) if a non-finite gradient anywhere makes every rank skip it
  auto adam_math = [&](float g, float& pv, float& mv, float& vv, float k, float& pskip) {
    mv = mv + (g - mv) * a.omb1;                  // exp_avg.lerp_(grad, 1-beta1)
    vv = vv * a.beta2 + (a.omb2 * g) * g;         // mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + a.eps;
    const float pn = pv + (-step_size * mv) / denom;  // addcdiv_(m, denom, -step_size)
    pskip = pv;
    pv = pn;
  };
  float pskip[kStepEndPre][4];
#pragma unroll
  for (int i = 0; i < kStepEndPre; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < pre[i].n) adam_math(pre[i].g[j], pre[i].p[j], pre[i].m[j], pre[i].v[j], pre[i].k[j], pskip[i][j]);
  if (tid == 0) {
    uint32_t ns = 16;
    while (true) {
      unsigned long long cur;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(cur) : "l"(a.bar) : "memory");
      if (cur >= target) break;
      __nanosleep(ns);
      if (ns < 256) ns <<= 1;
    }
    if (b == 0) {
      const float loss = ld_volatile_f1(r.stats + 1);
      const float flag = ld_volatile_f1(r.stats + 2);
      if (a.gstate) {
        float cap = a.gstate[1];
        if (flag != 0.f) cap = fmaxf(1.f, a.gstate[0] * (1.f / 16.f));
        // next seed scale: power of two near 0.125 / rmse, clamped to [1, cap]
        float g = 1.f;
        if (loss > 0.f && isfinite(loss)) g = exp2f(floorf(log2f(0.125f * rsqrtf(loss))));
        g = fminf(fmaxf(g, 1.f), cap);
        a.gstate[0] = g;
        a.gstate[1] = cap;
      }
      a.sched[6] = sh_sched[0];
      a.sched[7] = sh_sched[1];
      if (a.loss_ring) a.loss_ring[step % a.ring_len] = loss;
      if (a.loss_host) {
        a.loss_host[0] = loss;
        __threadfence_system();
        a.loss_host[1] = float(step + 1);
      }
      a.sched[0] = double(step + 1);
    }
  }
  __syncthreads();
  const bool skip = ld_volatile_f1(r.stats + 2) != 0.f;

  // ---- phase 2: the stores of the preloaded chunks ----
#pragma unroll
  for (int i = 0; i < kStepEndPre; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < pre[i].n) {
        const int tt = pre[i].t, e = pre[i].e + j;
        const bool masked = a.mask[tt] != nullptr;
        float pv = skip ? pskip[i][j] : pre[i].p[j];
        if (masked) pv = pv * pre[i].k[j];          // Masking.apply_mask
        a.p[tt][e] = pv;
        if (!skip) {
          a.m[tt][e] = pre[i].m[j];
          a.v[tt][e] = pre[i].v[j];
        }
      }
  // ---- ... and Adam (+ mask) on the block's remaining chunks (models with more than kStepEndPre chunks per block) ----
  auto adam1 = [&](float* P, float* M, float* V, const float* K, int i, float g) {
    float pv = P[i];
    if (!skip) {
      float mv = M[i], vv = V[i];
      mv = mv + (g - mv) * a.omb1;
      vv = vv * a.beta2 + (a.omb2 * g) * g;
      const float denom = sqrtf(vv) / bc2_sqrt + a.eps;
      pv = pv + (-step_size * mv) / denom;
      M[i] = mv;
      V[i] = vv;
    }
    if (K) pv = pv * K[i];
    P[i] = pv;
  };
  t = 0;
  for (int c = b + kStepEndPre * G; c < nchunks; c += G) {
    while (t + 1 < r.ndesc && c >= r.chunk_begin[t + 1]) ++t;
    const ReduceDesc d = r.d[t];
    if (a.p[t] == nullptr) continue;
    if (d.vec) {
      const int e4 = (c - r.chunk_begin[t]) * 1024 + tid * 4;
      if (worker && e4 < d.n) {
        const float4 g = *reinterpret_cast<const float4*>(d.dst + e4);
        adam1(a.p[t], a.m[t], a.v[t], a.mask[t], e4 + 0, g.x);
        adam1(a.p[t], a.m[t], a.v[t], a.mask[t], e4 + 1, g.y);
        adam1(a.p[t], a.m[t], a.v[t], a.mask[t], e4 + 2, g.z);
        adam1(a.p[t], a.m[t], a.v[t], a.mask[t], e4 + 3, g.w);
      }
    } else {
      const int e = (c - r.chunk_begin[t]) * 32 + tid;
      if (tid < 32 && e < d.n) adam1(a.p[t], a.m[t], a.v[t], a.mask[t], e, d.dst[e]);
    }
  }
  if (comm) {
    __syncthreads();
    if (tid == 0) a.epoch_b[b] = epoch;
  }
}

// ------------------------------------------------------------------------------------------
// Gradient exchange of a pixel-sharded fit over NVLink peer memory: in-place sum of a flat
// fp32 buffer over the ranks of ONE node.  Every rank owns a peer-mapped region {signals, two data
// buffers}.  Block b of every rank: (1) copies its slice of `io` into the local data buffer of this epoch's
// parity, (2) tells block b of every peer that the slice is there and waits for theirs (one system-scope
// flag per (parity, block, source rank)), (3) sums the slice over the ranks' buffers IN RANK ORDER — the
// same order on every rank, so the replicated weights stay bit-identical — and writes it back to `io`.
// No trailing barrier: the next exchange uses the other parity, and a rank can only be one exchange ahead
// of a peer because step (2) needs that peer's flag for the new epoch.  Epochs live in device memory
// (per block), so the launch is identical every step and can be captured in a CUDA graph.
// ------------------------------------------------------------------------------------------
// (constants, peer tables and the system-scope flag helpers are declared with step_end_kernel above)
__global__ void __launch_bounds__(kCommThreads) p2p_allreduce_kernel(const CommPeers cp, uint32_t* epoch_b, float* io,
                                                            int64_t n, int64_t max_floats, int rank,
                                                            int world, uint64_t timeout_ns) {
  const int b = blockIdx.x;
  const uint32_t epoch = epoch_b[b] + 1u;
  const uint32_t par = epoch & 1u;
  // slice of this block in units of 4 floats (n is padded to a multiple of 4 by the caller's buffer)
  const int64_t n4 = (n + 3) / 4;
  const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const int64_t lo = int64_t(b) * per, hi = (lo + per < n4) ? lo + per : n4;
  float* mine = cp.data[rank] + int64_t(par) * max_floats;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
    reinterpret_cast<float4*>(mine)[i] = reinterpret_cast<const float4*>(io)[i];
  __syncthreads();
  if (int(threadIdx.x) < world) {
    const int r = threadIdx.x;
    const int64_t slot = (int64_t(par) * kCommBlocks + b) * kCommMaxRanks;
    __threadfence_system();
    st_release_sys(cp.sig[r] + slot + rank, epoch);
    const uint32_t* flag = cp.sig[rank] + slot + r;
    wait_peer_flag(flag, epoch, timeout_ns, rank, b, r);
  }
  __syncthreads();
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    // all peers' loads in flight together (NVLink round trips overlap), then the rank-ordered sum
    float4 v[kCommMaxRanks];
#pragma unroll
    for (int r = 0; r < kCommMaxRanks; ++r)
      if (r < world) v[r] = ld_volatile_f4(cp.data[r] + int64_t(par) * max_floats + 4 * i);
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < kCommMaxRanks; ++r)
      if (r < world) {
        acc.x += v[r].x;
        acc.y += v[r].y;
        acc.z += v[r].z;
        acc.w += v[r].w;
      }
    reinterpret_cast<float4*>(io)[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) epoch_b[b] = epoch;
}


// ------------------------------------------------------------------------------------------
// Global-magnitude prune threshold (reference: pipeline/masking/funcs/prune.py:54-104).  The reference multiplies a
// persistent threshold up or down until the number of weights it removes is within `tolerance` of the target, each
// probe being `(|w| > threshold).sum()` per layer.  Here the magnitudes of all masked layers are sorted ONCE (device
// sort by the caller); a probe is then n_valid - upper_bound(threshold), and the whole search runs in ONE warp on
// the device.  The arithmetic is the reference's: Python floats are IEEE doubles, the comparison with the fp32
// weights rounds the threshold to fp32 (as torch does when it compares an fp32 tensor with a Python float), and
// the products are kept unfused (__dmul_rn) so that the threshold trajectory is bit-identical.
//   state[0] = threshold (in/out), state[1] = increment (in), result[0] = total_removed (out, as double)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ long long warp_upper_bound(const float* a, long long n, float t, int lane) {
  // number of elements <= t in the ascending, NaN-free array a[0, n): 32-ary search, one probe per lane.
  // invariant: a[i] <= t for i < lo, a[i] > t for i >= hi
  long long lo = 0, hi = n;
  while (hi - lo > 32) {
    const long long step = (hi - lo + 31) / 32;
    long long pos = lo + (lane + 1) * step - 1;
    if (pos > hi - 1) pos = hi - 1;
    const int k = __popc(__ballot_sync(0xffffffffu, a[pos] <= t));  // probes 0..k-1 hold (positions ascend)
    if (k == 32) return hi;  // the last probe is a[hi - 1]
    long long phi = lo + (k + 1) * step - 1;  // probe k: a[phi] > t
    if (phi > hi - 1) phi = hi - 1;
    lo += k * step;  // probe k-1 was not clamped (else all later probes would hold too)
    hi = phi;
  }
  const long long pos = lo + lane;
  return lo + __popc(__ballot_sync(0xffffffffu, pos < hi && a[pos] <= t));
}

__global__ void prune_threshold_search_kernel(const float* __restrict__ sorted_mags, long long n, long long nonzero_total,
                                              long long tokill, double tolerance, double* state, double* result) {
  const int lane = threadIdx.x & 31;
  // NaNs sort last: n_valid = index of the first NaN
  long long lo = 0, hi = n;
  while (lo < hi) {  // plain binary search, uniform across the warp
    const long long mid = (lo + hi) >> 1;
    const float v = sorted_mags[mid];
    if (v != v) hi = mid; else lo = mid + 1;
  }
  const long long n_valid = lo;
  double threshold = state[0], increment = state[1];
  long long total_removed = 0, prev_removed = 0;
  int tries = 0;
  const double tk = double(tokill);
  const double upper = __dmul_rn(tk, __dadd_rn(1.0, tolerance)), lower = __dmul_rn(tk, __dsub_rn(1.0, tolerance));
  const double band = __dmul_rn(tk, tolerance);
  for (int guard = 0; guard < 1000000; ++guard) {
    if (!(fabs(double(total_removed - tokill)) > band)) break;
    const long long le = warp_upper_bound(sorted_mags, n_valid, __double2float_rn(threshold), lane);
    total_removed = nonzero_total - (n_valid - le);
    if (prev_removed == total_removed) {
      if (++tries == 10) break;
    } else {
      tries = 0;
    }
    prev_removed = total_removed;
    if (double(total_removed) > upper) {
      threshold = __dmul_rn(threshold, __dsub_rn(1.0, increment));
      increment = __dmul_rn(increment, 0.99);
    } else if (double(total_removed) < lower) {
      threshold = __dmul_rn(threshold, __dadd_rn(1.0, increment));
      increment = __dmul_rn(increment, 0.99);
    }
  }
  if (lane == 0) {
    state[0] = threshold;
    result[0] = double(total_removed);
  }
}

}  // namespace sb
