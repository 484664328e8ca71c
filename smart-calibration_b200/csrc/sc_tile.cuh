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

// ---- K-chunked operand ring ------------------------------------------------------------------------------------
// Both GEMM operands stream through a NSTAGE-deep ring of shared-memory slots.  A dedicated producer warp fills a
// slot with 2-D tiled TMA loads (cp.async.bulk.tensor, one box of 32 rows x 32 floats per op, 128-byte swizzle)
// that complete on the slot's "full" mbarrier; the 8 consumer warps release it through the "empty" mbarrier.
// Slot layout: A part [8 k-blocks][32 rows][32 floats], W part the same (dgrad: [256 n][32 cols] as two 128-row panels).
// With SWIZZLE_128B the 16-byte chunk c of row r sits at chunk position c ^ (r & 7), which is exactly what makes the
// register-tile reads below conflict-free: lanes that read rows r, r+1, .., r+7 at the same k hit 8 different chunks.
constexpr int BOXF = 32 * 32;      // floats per TMA box (32 rows x 128 bytes)
// (slot geometry -- KC, NSTAGE, SLOT_A/W -- is a per-configuration constant of sc_rowphase_impl.cuh)

// ---- forward: rows rg+8i (i<TM), cols cg+4j (j<TN); operands read as float4 along k -------------------------
template <int TM, int TN>
__device__ __forceinline__ void mma_nt(const float* __restrict__ sA, const float* __restrict__ sW, int q0, int q1,
                                       float (&acc)[TM][TN]) {
  const int lane = threadIdx.x & 31, rg = lane >> 2, cg = lane & 3;
#pragma unroll 2
  for (int q = q0; q < q1; ++q) {
    const int kb = q >> 3, c = q & 7;
    const float* ab = sA + kb * BOXF + rg * 32 + ((c ^ rg) << 2);              // (row & 7) == rg for rows rg + 8i
    const float* bb = sW + kb * BOXF + cg * 32;
    const int x0 = (c ^ cg) << 2, x1 = (c ^ (cg + 4)) << 2;                    // (row & 7) for rows cg + 4j
    float4 a[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(ab + 256 * i);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const float4 b = *reinterpret_cast<const float4*>(bb + 128 * j + ((j & 1) ? x1 : x0));
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[i][j] = dot4(a[i], b, acc[i][j]);
    }
  }
}
__device__ __forceinline__ int nt_row(int lane, int i) { return (lane >> 2) + 8 * i; }
__device__ __forceinline__ int nt_col(int lane, int j) { return (lane & 3) + 4 * j; }

// ---- dgrad: rows rg+8i, cols cg*TN+j; A read as float4 along n, W panel sW[n][32 cols] (swizzled rows) -----------
template <int TM, int TN>
__device__ __forceinline__ void mma_nn(const float* __restrict__ sA, const float* __restrict__ sW, int q0, int q1,
                                       float (&acc)[TM][TN]) {
  const int lane = threadIdx.x & 31, rg = lane >> 2, cg = lane & 3;
#pragma unroll 1
  for (int q = q0; q < q1; ++q) {
    const int kb = q >> 3, c = q & 7;
    const float* ab = sA + kb * BOXF + rg * 32 + ((c ^ rg) << 2);
    float4 a[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(ab + 256 * i);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int n = 4 * q + s;                       // panel row (reduction index inside the chunk)
      const float* bp = sW + n * 32;
      const int sw = n & 7;
      float b[TN];
      if constexpr (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(bp + ((((cg * TN + j) >> 2) ^ sw) << 2));
          b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int col = cg * TN + j;
          b[j] = bp[(((col >> 2) ^ sw) << 2) + (col & 3)];
        }
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
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
// red: [8][8*TM][NT+4] floats.  Result (plus optional bias over columns) lands in sZ[8*TM][NT] (pitch NT).
template <int TM, int TN, bool NN>
__device__ __forceinline__ void park_partial(float* red, const float (&acc)[TM][TN]) {
  constexpr int NT = 4 * TN, P = NT + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* base = red + warp * (8 * TM) * P;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = NN ? nn_col<TN>(lane, j) : nt_col(lane, j);
      base[nt_row(lane, i) * P + c] = acc[i][j];
    }
}
template <int TM, int TN>
__device__ __forceinline__ void combine_partials(const float* red, float* sZ, const float* __restrict__ bias_slice,
                                                 int ncv) {
  constexpr int NT = 4 * TN, P = NT + 4;
  for (int e = threadIdx.x; e < 8 * TM * NT; e += SC_THREADS) {   // consumer threads only (tid < 256)
    const int r = e / NT, c = e % NT;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < SC_WARPS; ++w) s += red[(w * (8 * TM) + r) * P + c];
    if (bias_slice != nullptr && c < ncv) s += ld_w(bias_slice + c);
    sZ[r * NT + c] = s;
  }
}

}  // namespace sc
