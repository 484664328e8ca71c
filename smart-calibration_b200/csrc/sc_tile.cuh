// FP32 register-tiled GEMM pieces used by the cluster row-phase kernel and the wgrad kernel.
//
// All three contractions of the MLP step are served from row-major global layouts (PyTorch-native W[N][K]):
//   forward  Z[m][n]  = sum_k A[m][k]  * W[n][k]     "nt": both operands reduction-contiguous in smem
//   dgrad    dA[m][k] = sum_n dZ[m][n] * W[n][k]     "nn": B operand is a [n][k-slice] panel
//   wgrad    dW[n][k] = sum_b dZ[b][n] * A[b][k]     "tn": both operands are [b][slice] panels
// A warp owns a 32 x (4*TN) output tile; the 8 warps of a CTA split the REDUCTION axis and their partial
// tiles are summed in fixed warp order through shared memory (deterministic).
#pragma once
#include "sc_common.cuh"

namespace sc {

// ---- forward: rows rg+8i (i<4), cols cg+4j (j<TN); operands read as float4 along k --------------------
// sA: [32][pa] activations, sW: [4*TN][pw] weight slice; pa, pw = smem_pitch(K) -> conflict-free LDS.128.
template <int TN>
__device__ __forceinline__ void mma_nt(const float* __restrict__ sA, int pa, const float* __restrict__ sW, int pw,
                                       int q0, int q1, float (&acc)[4][TN]) {
  const int lane = threadIdx.x & 31, rg = lane >> 2, cg = lane & 3;
  const float* a0 = sA + rg * pa;
  const float* b0 = sW + cg * pw;
#pragma unroll 2
  for (int q = q0; q < q1; ++q) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + (8 * i) * pa + 4 * q);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const float4 b = *reinterpret_cast<const float4*>(b0 + (4 * j) * pw + 4 * q);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][j] = dot4(a[i], b, acc[i][j]);
    }
  }
}
__device__ __forceinline__ int nt_row(int lane, int i) { return (lane >> 2) + 8 * i; }
__device__ __forceinline__ int nt_col(int lane, int j) { return (lane & 3) + 4 * j; }

// ---- dgrad: rows rg+8i, cols cg*TN+j; A read as float4 along n, W panel sW[n][4*TN] -------------------
template <int TN>
__device__ __forceinline__ void mma_nn(const float* __restrict__ sA, int pa, const float* __restrict__ sW, int pw,
                                       int q0, int q1, float (&acc)[4][TN]) {
  const int lane = threadIdx.x & 31, rg = lane >> 2, cg = lane & 3;
  const float* a0 = sA + rg * pa;
  const float* b0 = sW + cg * TN;
#pragma unroll 1
  for (int q = q0; q < q1; ++q) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + (8 * i) * pa + 4 * q);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      float b[TN];
      const float* bp = b0 + (4 * q + s) * pw;
      if constexpr (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(bp + j);
          b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = bp[j];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = s == 0 ? a[i].x : s == 1 ? a[i].y : s == 2 ? a[i].z : a[i].w;
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
      }
    }
  }
}
template <int TN> __device__ __forceinline__ int nn_col(int lane, int j) { return (lane & 3) * TN + j; }

// ---- wgrad: out rows (n) rg*4+i, out cols (k) cg*8+j; one reduction index b per step -------------------
// sD: [b][32] dZ panel, sX: [b][32] input panel (pitch 32).
__device__ __forceinline__ void mma_tn(const float* __restrict__ sD, const float* __restrict__ sX, int b0, int b1,
                                       float (&acc)[4][8]) {
  const int lane = threadIdx.x & 31, rg = lane >> 2, cg = lane & 3;
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const float4 d = *reinterpret_cast<const float4*>(sD + b * 32 + rg * 4);
    const float4 x0 = *reinterpret_cast<const float4*>(sX + b * 32 + cg * 8);
    const float4 x1 = *reinterpret_cast<const float4*>(sX + b * 32 + cg * 8 + 4);
    const float dv[4] = {d.x, d.y, d.z, d.w};
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
  }
}

// ---- split-reduction combine: every warp parks its partial tile, then 256 threads add the 8 partials -----
// red: [8][32][NT+4] floats.  Result (plus optional bias over columns) lands in sZ[32][NT] (pitch NT).
template <int TN, bool NN>
__device__ __forceinline__ void park_partial(float* red, const float (&acc)[4][TN]) {
  constexpr int NT = 4 * TN, P = NT + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* base = red + warp * 32 * P;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = NN ? nn_col<TN>(lane, j) : nt_col(lane, j);
      base[nt_row(lane, i) * P + c] = acc[i][j];
    }
}
template <int TN>
__device__ __forceinline__ void combine_partials(const float* red, float* sZ, const float* __restrict__ bias_slice,
                                                 int ncv) {
  constexpr int NT = 4 * TN, P = NT + 4;
  for (int e = threadIdx.x; e < 32 * NT; e += SC_THREADS) {
    const int r = e / NT, c = e % NT;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < SC_WARPS; ++w) s += red[(w * 32 + r) * P + c];
    if (bias_slice != nullptr && c < ncv) s += ld_w(bias_slice + c);
    sZ[r * NT + c] = s;
  }
}

// ---- staging: copy `nrows_valid` rows of `k4` float4 into a [32][pitch] smem tile, zero-filling the rest ----
// row_ptr(r) returns the global address of row r.  Warp w serves rows w, w+8, w+16, w+24 (no integer division).
template <bool ACT, typename RowPtr>
__device__ __forceinline__ void stage_rows(float* sm, int pitch, int nrows_total, int nrows_valid, int k4, RowPtr row_ptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < nrows_total; r += SC_WARPS) {
    float* dst = sm + r * pitch;
    if (r < nrows_valid) {
      const float* src = row_ptr(r);
      for (int c = lane; c < k4; c += 32) {
        const float4 v = ACT ? ld_act4(src + 4 * c) : ld_w4(src + 4 * c);
        *reinterpret_cast<float4*>(dst + 4 * c) = v;
      }
    } else {
      for (int c = lane; c < k4; c += 32) *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

}  // namespace sc
