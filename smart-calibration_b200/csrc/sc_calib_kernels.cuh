// Kernels of the calibration conv-SAC update (calibration/calib_sac.py): stride-2 5x5 convolutions (forward, input
// gradient, weight gradient), training-mode BatchNorm2d (batch statistics, running-statistics update, backward),
// the small dense layers around the trunk, the policy head and Adam.  FP32 throughout; the convolutions are direct
// register-blocked FFMA2 (fma.rn.f32x2) kernels.  Layouts are torch's: activations [B][C][H][W], conv weights
// [Cout][Cin][5][5], Linear weights [out][in].  Included by sc_calib.cu only.
#pragma once
#include "sc_common.cuh"

namespace sc {
namespace calib {

constexpr int KS = 5;          // kernel size (calib_sac.py:101,103,105)
constexpr int TAPS = 25;
enum { ACT_RELU = 0, ACT_ELU = 1 };

__device__ __forceinline__ float act_fwd(float u, int act) { return act == ACT_RELU ? fmaxf(u, 0.f) : elu(u); }
// d act / d u expressed through the output a
__device__ __forceinline__ float act_grad(float a, int act) { return act == ACT_RELU ? (a > 0.f ? 1.f : 0.f) : elu_grad_from_out(a); }

__device__ __forceinline__ double block_sum_double(double v, double* sh) {      // 256 threads; result valid in every thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += sh[w];
  return s;
}
__device__ __forceinline__ float block_sum_float(float v, float* sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += sh[w];
  return s;
}

// =====================================================================================================================
// conv forward: z[b][co][oy][ox] = bias[co] + sum_{ci,ky,kx} w[co][ci][ky][kx] * x[b][ci][2 oy + ky][2 ox + kx]
// A thread owns 2 horizontally adjacent output pixels x 8 output channels; a block owns RPB output rows of one image.
// Weights of 8 input channels at a time are staged transposed in shared memory ([ci][tap][co]) so that the 8 channel
// weights of a tap are two broadcast 16-byte reads; the 7 input values a pixel pair needs per kernel row come from L1.
// =====================================================================================================================
template <int CIN, int COUT, int HIN, int HOUT>
struct ConvGeom {
  static constexpr int PP = (HOUT + 1) / 2;            // pixel pairs per output row
  static constexpr int CG = COUT / 8;                  // channel groups
  static constexpr int RPB = 256 / (PP * CG);          // output rows per block
  static constexpr int BLOCKS_PER_IMG = (HOUT + RPB - 1) / RPB;
  static constexpr int CIC = CIN < 8 ? CIN : 8;        // input channels per weight stage
};

template <int CIN, int COUT, int HIN, int HOUT>
__global__ void __launch_bounds__(256) k_conv_fwd(const float* __restrict__ x, size_t x_img_stride, const float* __restrict__ w,
                                                   const float* __restrict__ bias, float* __restrict__ z) {
  using G = ConvGeom<CIN, COUT, HIN, HOUT>;
  __shared__ __align__(16) float sw[G::CIC * TAPS * COUT];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int cg = tid / (G::RPB * G::PP);
  const int rem = tid - cg * (G::RPB * G::PP);
  const int r = rem / G::PP, pair = rem - r * G::PP;
  const int oy = blockIdx.x * G::RPB + r;
  const bool active = cg < G::CG && oy < HOUT;
  const int ox0 = 2 * pair;
  const bool px1 = ox0 + 1 < HOUT;
  float2 acc[2][4];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[p][j] = make_float2(0.f, 0.f);
  const float* xb = x + (size_t)b * x_img_stride;
  for (int c0 = 0; c0 < CIN; c0 += G::CIC) {
    __syncthreads();
    for (int i = tid; i < G::CIC * TAPS * COUT; i += 256) {
      const int co = i % COUT, ct = i / COUT;            // ct = ci_local * 25 + tap
      const int cl = ct / TAPS, tap = ct - cl * TAPS;
      sw[i] = __ldg(w + ((size_t)co * CIN + (c0 + cl)) * TAPS + tap);
    }
    __syncthreads();
    if (active) {
#pragma unroll 1
      for (int cl = 0; cl < G::CIC; ++cl) {
        const float* xc = xb + (size_t)(c0 + cl) * HIN * HIN + (size_t)(2 * oy) * HIN + 2 * ox0;
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
          float xv[7];
#pragma unroll
          for (int i = 0; i < 7; ++i) xv[i] = (2 * ox0 + i < HIN) ? __ldg(xc + ky * HIN + i) : 0.f;
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            const float4* wp = reinterpret_cast<const float4*>(sw + ((cl * TAPS) + ky * KS + kx) * COUT + cg * 8);
            const float4 w0 = wp[0], w1 = wp[1];
            const float2 wq[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
            const float2 a0 = make_float2(xv[kx], xv[kx]), a1 = make_float2(xv[kx + 2], xv[kx + 2]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[0][j] = __ffma2_rn(a0, wq[j], acc[0][j]);
              acc[1][j] = __ffma2_rn(a1, wq[j], acc[1][j]);
            }
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = cg * 8 + 2 * j;
      const float b0 = __ldg(bias + co), b1 = __ldg(bias + co + 1);
      float* z0 = z + (((size_t)b * COUT + co) * HOUT + oy) * HOUT + ox0;
      float* z1 = z0 + (size_t)HOUT * HOUT;
      z0[0] = acc[0][j].x + b0;
      z1[0] = acc[0][j].y + b1;
      if (px1) {
        z0[1] = acc[1][j].x + b0;
        z1[1] = acc[1][j].y + b1;
      }
    }
  }
}

// =====================================================================================================================
// conv input gradient: da[b][ci][iy][ix] = sum_{co} sum_{ky = iy mod 2 (2), kx = ix mod 2 (2)} dz[b][co][(iy-ky)/2][(ix-kx)/2] w[co][ci][ky][kx]
// A thread owns the pixels (ix, ix+2) of one input row (same kx parity, neighbouring output columns) x 8 input channels.
// Weights of 8 output channels at a time are staged as [co][tap][ci].
// =====================================================================================================================
template <int CIN, int COUT, int HIN, int HOUT>
struct DgradGeom {
  static constexpr int TPR = ((HIN + 3) / 4) * 2;      // threads per input row (pixel pairs)
  static constexpr int CG = CIN / 8;
  static constexpr int RPB = 256 / (TPR * CG);
  static constexpr int BLOCKS_PER_IMG = (HIN + RPB - 1) / RPB;
};
template <int CIN, int COUT, int HIN, int HOUT>
__global__ void __launch_bounds__(256) k_conv_dgrad(const float* __restrict__ dz, const float* __restrict__ w, float* __restrict__ da) {
  using G = DgradGeom<CIN, COUT, HIN, HOUT>;
  __shared__ __align__(16) float sw[8 * TAPS * CIN];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int cg = tid / (G::RPB * G::TPR);
  const int rem = tid - cg * (G::RPB * G::TPR);
  const int r = rem / G::TPR, t = rem - r * G::TPR;
  const int iy = blockIdx.x * G::RPB + r;
  const int ix0 = 4 * (t >> 1) + (t & 1), ix1 = ix0 + 2;
  const bool active = cg < G::CG && iy < HIN && ix0 < HIN;
  const int py = iy & 1, px = ix0 & 1;
  float2 acc[2][4];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[p][j] = make_float2(0.f, 0.f);
  const float* dzb = dz + (size_t)b * COUT * HOUT * HOUT;
  for (int c0 = 0; c0 < COUT; c0 += 8) {
    __syncthreads();
    for (int i = tid; i < 8 * TAPS * CIN; i += 256) {
      const int ci = i % CIN, ct = i / CIN;              // ct = co_local * 25 + tap
      const int cl = ct / TAPS, tap = ct - cl * TAPS;
      sw[i] = __ldg(w + ((size_t)(c0 + cl) * CIN + ci) * TAPS + tap);
    }
    __syncthreads();
    if (active) {
#pragma unroll 1
      for (int cl = 0; cl < 8; ++cl) {
        const float* dzc = dzb + (size_t)(c0 + cl) * HOUT * HOUT;
#pragma unroll
        for (int jy = 0; jy < 3; ++jy) {
          const int ky = py + 2 * jy;
          const int oy = (iy - ky) >> 1;
          if (ky >= KS || iy < ky || oy >= HOUT) continue;
#pragma unroll
          for (int jx = 0; jx < 3; ++jx) {
            const int kx = px + 2 * jx;
            if (kx >= KS) continue;
            const int ox = (ix0 - kx) >> 1;              // output column feeding pixel ix0; pixel ix1 uses ox + 1
            const float d0 = (ix0 >= kx && ox < HOUT) ? __ldg(dzc + oy * HOUT + ox) : 0.f;
            const float d1 = (ix1 < HIN && ix1 >= kx && ox + 1 < HOUT) ? __ldg(dzc + oy * HOUT + ox + 1) : 0.f;
            const float4* wp = reinterpret_cast<const float4*>(sw + (cl * TAPS + ky * KS + kx) * CIN + cg * 8);
            const float4 w0 = wp[0], w1 = wp[1];
            const float2 wq[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
            const float2 a0 = make_float2(d0, d0), a1 = make_float2(d1, d1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[0][j] = __ffma2_rn(a0, wq[j], acc[0][j]);
              acc[1][j] = __ffma2_rn(a1, wq[j], acc[1][j]);
            }
          }
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = cg * 8 + 2 * j;
      float* o0 = da + (((size_t)b * CIN + ci) * HIN + iy) * HIN;
      float* o1 = o0 + (size_t)HIN * HIN;
      o0[ix0] = acc[0][j].x;
      o1[ix0] = acc[0][j].y;
      if (ix1 < HIN) {
        o0[ix1] = acc[1][j].x;
        o1[ix1] = acc[1][j].y;
      }
    }
  }
}

// =====================================================================================================================
// conv weight gradient (CIN >= 8): dw[co][ci][tap] = sum_{b,oy,ox} dz[b][co][oy][ox] * x[b][ci][2 oy + ky][2 ox + kx]
// Persistent blocks walk (image, output row) items; per item the 5 input rows of every channel and the dz row are staged
// in shared memory.  Thread (pixel slice ps, channel group of 4 co, ci) keeps 4 x 25 accumulators.  Every block writes
// its partial sum; k_sum_partials adds them in block order (deterministic).
// =====================================================================================================================
template <int CIN, int COUT, int HIN, int HOUT>
__global__ void __launch_bounds__(256, 1) k_conv_wgrad(const float* __restrict__ x, size_t x_img_stride, const float* __restrict__ dz, int B,
                                                       float* __restrict__ partial) {
  constexpr int NTH = (COUT / 4) * CIN;                // threads that tile (co group, ci)
  constexpr int PS = 256 / NTH;                        // pixel slices
  static_assert(NTH * PS == 256, "thread tiling");
  extern __shared__ __align__(16) float smw[];
  float* sx = smw;                                     // [CIN][5][HIN]
  float* sd = sx + CIN * KS * HIN;                     // [COUT][HOUT]
  const int tid = threadIdx.x;
  const int ps = tid / NTH, tl = tid - ps * NTH;
  const int cg = tl / CIN, ci = tl - cg * CIN;
  float acc[4][TAPS];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int t = 0; t < TAPS; ++t) acc[u][t] = 0.f;
  const int items = B * HOUT;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int b = it / HOUT, oy = it - b * HOUT;
    __syncthreads();
    const float* xb = x + (size_t)b * x_img_stride + (size_t)(2 * oy) * HIN;
    for (int i = tid; i < CIN * KS * HIN; i += 256) {
      const int c = i / (KS * HIN), rr = i - c * (KS * HIN);           // rr = ky * HIN + col
      sx[i] = __ldg(xb + (size_t)c * HIN * HIN + rr);
    }
    const float* dzb = dz + (size_t)b * COUT * HOUT * HOUT + (size_t)oy * HOUT;
    for (int i = tid; i < COUT * HOUT; i += 256) {
      const int c = i / HOUT, o = i - c * HOUT;
      sd[i] = __ldg(dzb + (size_t)c * HOUT * HOUT + o);
    }
    __syncthreads();
    const float* sxc = sx + ci * KS * HIN;
#pragma unroll 1
    for (int ox = ps; ox < HOUT; ox += PS) {
      float d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = sd[(cg * 4 + u) * HOUT + ox];
#pragma unroll
      for (int ky = 0; ky < KS; ++ky)
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float xv = sxc[ky * HIN + 2 * ox + kx];
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u][ky * KS + kx] = fmaf(d[u], xv, acc[u][ky * KS + kx]);
        }
    }
  }
  // pixel slices are added in order through shared memory, then the block's partial goes out
  __syncthreads();
  float* red = smw;                                    // [NTH][100]
  for (int s = 0; s < PS; ++s) {
    if (ps == s) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
          const int o = tl * 100 + u * TAPS + t;
          red[o] = (s == 0 ? 0.f : red[o]) + acc[u][t];
        }
    }
    __syncthreads();
  }
  float* outp = partial + (size_t)blockIdx.x * COUT * CIN * TAPS;
  for (int i = tid; i < NTH * 100; i += 256) {
    const int tl2 = i / 100, rr = i - tl2 * 100, u = rr / TAPS, t = rr - u * TAPS;
    const int cg2 = tl2 / CIN, ci2 = tl2 - cg2 * CIN;
    outp[((size_t)(cg2 * 4 + u) * CIN + ci2) * TAPS + t] = red[i];
  }
}

// first layer (CIN = 1, COUT = 16, 128 -> 62): thread (pixel slice ps of 16, co) keeps 25 accumulators
__global__ void __launch_bounds__(256, 2) k_conv_wgrad_l1(const float* __restrict__ x, size_t x_img_stride, const float* __restrict__ dz, int B,
                                                          float* __restrict__ partial) {
  constexpr int HIN = 128, HOUT = 62, COUT = 16;
  __shared__ float sx[KS * HIN];
  __shared__ float sd[COUT * HOUT];
  __shared__ float red[COUT * TAPS];
  const int tid = threadIdx.x, co = tid & 15, ps = tid >> 4;
  float acc[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) acc[t] = 0.f;
  const int items = B * HOUT;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int b = it / HOUT, oy = it - b * HOUT;
    __syncthreads();
    const float* xb = x + (size_t)b * x_img_stride + (size_t)(2 * oy) * HIN;
    for (int i = tid; i < KS * HIN; i += 256) sx[i] = __ldg(xb + i);
    const float* dzb = dz + (size_t)b * COUT * HOUT * HOUT + (size_t)oy * HOUT;
    for (int i = tid; i < COUT * HOUT; i += 256) {
      const int c = i / HOUT, o = i - c * HOUT;
      sd[i] = __ldg(dzb + (size_t)c * HOUT * HOUT + o);
    }
    __syncthreads();
#pragma unroll 1
    for (int ox = ps; ox < HOUT; ox += 16) {
      const float d = sd[co * HOUT + ox];
#pragma unroll
      for (int ky = 0; ky < KS; ++ky)
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) acc[ky * KS + kx] = fmaf(d, sx[ky * HIN + 2 * ox + kx], acc[ky * KS + kx]);
    }
  }
  __syncthreads();
  for (int s = 0; s < 16; ++s) {
    if (ps == s) {
#pragma unroll
      for (int t = 0; t < TAPS; ++t) red[co * TAPS + t] = (s == 0 ? 0.f : red[co * TAPS + t]) + acc[t];
    }
    __syncthreads();
  }
  for (int i = tid; i < COUT * TAPS; i += 256) partial[(size_t)blockIdx.x * COUT * TAPS + i] = red[i];
}

// out[i] = sum_{p < np} partial[p][i]   (fixed order)
__global__ void k_sum_partials(const float* __restrict__ partial, int np, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < np; ++p) s += partial[(size_t)p * n + i];
  out[i] = s;
}

// =====================================================================================================================
// per-channel reductions over (B, H, W) for BatchNorm: grid (C, NCH); double partials [C][NCH][2]
//   mode 0: (sum z, sum z^2)                      forward statistics
//   mode 1: (sum du, sum du * xhat)               backward, du = dA * act'(a), xhat = (z - mean) * rstd
//   mode 2: (sum dz, 0)                           conv bias gradient
// =====================================================================================================================
struct ChanReduceP {
  const float* z; const float* a; const float* dA;      // z: [B][C][HW] contiguous; a, dA: image strides below
  size_t a_img_stride, dA_img_stride;
  const float* mean; const float* rstd;
  int B, C, HW, act, mode;
  double* partial;                                      // [C][gridDim.y][2]
};
__global__ void __launch_bounds__(256) k_chan_reduce(const ChanReduceP p) {
  __shared__ double sh[8];
  const int c = blockIdx.x, ch = blockIdx.y, nch = gridDim.y;
  const int ipb = (p.B + nch - 1) / nch;
  const int b0 = ch * ipb, b1 = min(p.B, b0 + ipb);
  double s1 = 0.0, s2 = 0.0;
  const float mean = p.mode == 1 ? p.mean[c] : 0.f, rstd = p.mode == 1 ? p.rstd[c] : 0.f;
  for (int b = b0; b < b1; ++b) {
    const float* zc = p.z + ((size_t)b * p.C + c) * p.HW;
    const float* ac = p.a ? p.a + (size_t)b * p.a_img_stride + (size_t)c * p.HW : nullptr;
    const float* dc = p.dA ? p.dA + (size_t)b * p.dA_img_stride + (size_t)c * p.HW : nullptr;
    float f1 = 0.f, f2 = 0.f;
    for (int i = threadIdx.x; i < p.HW; i += 256) {
      if (p.mode == 0) {
        const float v = zc[i];
        f1 += v;
        f2 = fmaf(v, v, f2);
      } else if (p.mode == 1) {
        const float du = dc[i] * act_grad(ac[i], p.act);
        f1 += du;
        f2 = fmaf(du, (zc[i] - mean) * rstd, f2);
      } else {
        f1 += dc[i];
      }
    }
    s1 += (double)f1;
    s2 += (double)f2;
  }
  s1 = block_sum_double(s1, sh);
  s2 = block_sum_double(s2, sh);
  if (threadIdx.x == 0) {
    p.partial[((size_t)c * nch + ch) * 2] = s1;
    p.partial[((size_t)c * nch + ch) * 2 + 1] = s2;
  }
}

struct BnFinalizeP {
  const double* partial; int nch;
  const float* gamma; const float* beta;
  float* rmean; float* rvar; long long* nbt;            // running statistics (updated when train)
  float* mean; float* rstd; float* scale; float* shift; // per-channel outputs
  double n;                                             // B * H * W
  int train;
};
// one thread per channel: batch statistics (train) or running statistics (eval), nn.BatchNorm2d semantics
// (momentum 0.1, eps 1e-5, biased variance for normalisation, unbiased for the running estimate)
__global__ void k_bn_finalize(const BnFinalizeP p, int C) {
  const int c = threadIdx.x;
  if (c >= C) return;
  double mean, var;
  if (p.train) {
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < p.nch; ++i) {
      s1 += p.partial[((size_t)c * p.nch + i) * 2];
      s2 += p.partial[((size_t)c * p.nch + i) * 2 + 1];
    }
    mean = s1 / p.n;
    var = fmax(s2 / p.n - mean * mean, 0.0);
    p.rmean[c] = (float)(0.9 * (double)p.rmean[c] + 0.1 * mean);
    p.rvar[c] = (float)(0.9 * (double)p.rvar[c] + 0.1 * var * (p.n / (p.n - 1.0)));
    if (c == 0) p.nbt[0] += 1;
  } else {
    mean = (double)p.rmean[c];
    var = (double)p.rvar[c];
  }
  const float rstd = (float)(1.0 / sqrt(var + 1e-5));
  const float sc = p.gamma[c] * rstd;
  p.mean[c] = (float)mean;
  p.rstd[c] = rstd;
  p.scale[c] = sc;
  p.shift[c] = p.beta[c] - (float)mean * sc;
}

// a[b][c][i] = act(z[b][c][i] * scale[c] + shift[c]); `a` may have a different image stride (the flatten + concat buffer)
__global__ void k_bn_act(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift, int C, int HW,
                         size_t total, int act, float* __restrict__ a, size_t a_img_stride) {
  const size_t chw = (size_t)C * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / chw, r = i - b * chw;
    const int c = (int)(r / HW);
    a[b * a_img_stride + r] = act_fwd(fmaf(z[i], scale[c], shift[c]), act);
  }
}

struct BnBwdFinalizeP {
  const double* partial; int nch;
  float* dgamma; float* dbeta; float* c1; float* c2;
  double n;
};
__global__ void k_bn_bwd_finalize(const BnBwdFinalizeP p, int C) {
  const int c = threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < p.nch; ++i) {
    s1 += p.partial[((size_t)c * p.nch + i) * 2];
    s2 += p.partial[((size_t)c * p.nch + i) * 2 + 1];
  }
  p.dbeta[c] = (float)s1;
  p.dgamma[c] = (float)s2;
  p.c1[c] = (float)(s1 / p.n);
  p.c2[c] = (float)(s2 / p.n);
}
// dz = gamma * rstd * (du - mean(du) - xhat * mean(du * xhat)), du = dA * act'(a); written contiguous [B][C][HW]
__global__ void k_bn_bwd_apply(const float* __restrict__ dA, size_t dA_img_stride, const float* __restrict__ a, size_t a_img_stride,
                               const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ mean,
                               const float* __restrict__ rstd, const float* __restrict__ c1, const float* __restrict__ c2, int C, int HW,
                               size_t total, int act, float* __restrict__ dz) {
  const size_t chw = (size_t)C * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / chw, r = i - b * chw;
    const int c = (int)(r / HW);
    const float du = dA[b * dA_img_stride + r] * act_grad(a[b * a_img_stride + r], act);
    const float xh = (z[i] - mean[c]) * rstd[c];
    dz[i] = scale[c] * (du - c1[c] - xh * c2[c]);
  }
}
__global__ void k_bias_grad_finalize(const double* __restrict__ partial, int nch, int C, float* __restrict__ db) {
  const int c = threadIdx.x;
  if (c >= C) return;
  double s = 0.0;
  for (int i = 0; i < nch; ++i) s += partial[((size_t)c * nch + i) * 2];
  db[c] = (float)s;
}

// =====================================================================================================================
// generic strided FP32 GEMM for the dense layers: C[m][n] (+)= sum_k A(m,k) B(k,n), A(m,k) = A[m*sam + k*sak],
// B(k,n) = B[k*sbk + n*sbn].  64 x 64 tiles, 16-wide k slabs, 4 x 4 outputs per thread.
// =====================================================================================================================
// With gridDim.z > 1 the reduction axis is cut into gridDim.z slices of `kslice` (a multiple of 16) and slice z writes its
// partial product to Cm + z * M * ldc (no bias / activation); k_splitk_reduce adds the slices in order.
__global__ void __launch_bounds__(256) k_sgemm(const float* __restrict__ A, long sam, long sak, const float* __restrict__ Bm, long sbk, long sbn,
                                               float* __restrict__ Cm, long ldc, int M, int N, int K, const float* __restrict__ bias_n, int act,
                                               int kslice) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int kbeg = blockIdx.z * kslice, kend = min(K, kbeg + kslice);
  Cm += (size_t)blockIdx.z * M * ldc;
  K = kend;
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
    for (int i = tid; i < 16 * 64; i += 256) {
      int kk, mm;
      if (sak == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }     // walk the contiguous axis fastest
      const int m = m0 + mm, k = k0 + kk;
      sA[kk][mm] = (m < M && k < K) ? __ldg(A + (long)m * sam + (long)k * sak) : 0.f;
    }
    for (int i = tid; i < 16 * 64; i += 256) {
      int kk, nn;
      if (sbk == 1) { kk = i & 15; nn = i >> 4; } else { nn = i & 63; kk = i >> 6; }
      const int n = n0 + nn, k = k0 + kk;
      sB[kk][nn] = (n < N && k < K) ? __ldg(Bm + (long)k * sbk + (long)n * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias_n ? __ldg(bias_n + n) : 0.f);
      if (act >= 0) v = act_fwd(v, act);
      Cm[(long)m * ldc + n] = v;
    }
  }
}

// C[m][n] = act(bias[n] + sum_z partial[z][m][n])
__global__ void k_splitk_reduce(const float* __restrict__ partial, int nz, int M, int N, const float* __restrict__ bias_n, int act,
                                float* __restrict__ Cm, long ldc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float s = 0.f;
  for (int z = 0; z < nz; ++z) s += partial[((size_t)z * M + m) * N + n];
  s += bias_n ? __ldg(bias_n + n) : 0.f;
  if (act >= 0) s = act_fwd(s, act);
  Cm[(long)m * ldc + n] = s;
}

// out[n] = sum_b w[b] * X[b][n]  (w == nullptr: plain column sum); one thread per column, fixed order
// out[n] = sum_b wrow[b] * X[b][n]: 32 columns per block, 8 row groups (rows rg, rg + 8, ...) combined in fixed order
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ X, long ldx, const float* __restrict__ wrow, int B, int N, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int c = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  float s = 0.f;
  if (n < N)
    for (int b = rg; b < B; b += 8) s = fmaf(wrow ? wrow[b] : 1.f, X[(long)b * ldx + n], s);
  part[rg][c] = s;
  __syncthreads();
  if (rg == 0 && n < N) {
    float t = part[0][c];
#pragma unroll
    for (int r = 1; r < 8; ++r) t += part[r][c];
    out[n] = t;
  }
}

// =====================================================================================================================
// Adam over a flat parameter block (torch.optim.Adam defaults; same roundings as sc_wgrad.cu)
// =====================================================================================================================
struct AdamP { float lr_neg_step, bc2_sqrt, one_minus_b1, b2, one_minus_b2, eps; };
// (the step-dependent scalars come from device memory, written by the one small host->device copy at the head of a step, so
//  that the launch itself is identical every step and a learn() can be replayed from a CUDA graph)
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n, const AdamP* __restrict__ hp) {
  const AdamP h = *hp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    float mi = m[i], vi = v[i];
    mi = fmaf(h.one_minus_b1, gi - mi, mi);
    vi = __fadd_rn(__fmul_rn(vi, h.b2), __fmul_rn(__fmul_rn(h.one_minus_b2, gi), gi));
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), h.bc2_sqrt), h.eps);
    p[i] = __fadd_rn(p[i], __fdiv_rn(__fmul_rn(h.lr_neg_step, mi), denom));
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace calib
}  // namespace sc
