// FP32 register-tiled GEMM pieces used by the cluster row-phase kernel and the wgrad kernel.
//
// All three contractions of the MLP step are served from row-major global layouts (PyTorch-native W[N][K]):
//   forward  Z[m][n]  = sum_k A[m][k]  * W[n][k]     "nt": both operands reduction-contiguous in smem
//   dgrad    dA[m][k] = sum_n dZ[m][n] * W[n][k]     "nn": B operand is a [n][k-slice] panel
//   wgrad    dW[n][k] = sum_b dZ[b][n] * A[b][k]     "tn": both operands are [b][slice] panels (sc_wgrad.cu)
// forward and dgrad run on the tensor cores (mma.sync TF32 with error compensation, below); the wgrad kernel keeps a
// register-tiled FP32 FFMA loop.  The 8 warps of a CTA split the reduction axis and their partial tiles are summed in
// fixed order through shared memory (deterministic).
#pragma once
#include "sc_common.cuh"

namespace sc {

// ---- K-chunked operand ring ------------------------------------------------------------------------------------
// Both GEMM operands stream through a NSTAGE-deep ring of shared-memory slots.  A dedicated producer warp fills a
// slot with 2-D tiled TMA loads (cp.async.bulk.tensor, one box of 32 rows x 32 floats per op, 128-byte swizzle)
// that complete on the slot's "full" mbarrier; the 8 consumer warps release it through the "empty" mbarrier.
// Slot layout: X part [k-blocks][tile rows][32 floats], W part [k-blocks][<=64 features][32 floats] (dgrad: panels of
// [reduction rows][32 cols]).  With SWIZZLE_128B the 16-byte chunk c of row r sits at chunk position c ^ (r & 7), which is
// exactly what makes the fragment reads below conflict-free: 8 rows read at the same k hit 8 different chunks.
constexpr int BOXF = 32 * 32;      // floats per TMA box (32 rows x 128 bytes)
// (slot geometry -- KC, NSTAGE, SLOT_A/W -- is a per-configuration constant of sc_rowphase_impl.cuh)

// ---- tensor-core inner loops: mma.sync m16n8k8 TF32 with error compensation ("3xTF32") ------------------------------
// The row-phase GEMMs have only 8..32 batch rows per CTA, so they are bound by how often a weight element has to come
// out of shared memory, not by FMA rate.  They are therefore issued "swapped": the 16-wide M side of the MMA walks the
// output FEATURES (weights = A fragment, loaded once per k-step with one ldmatrix.x4 and reused by every row tile),
// the 8-wide N side walks the batch rows.  Each fp32 operand x is split into big = tf32(x) (low 13 mantissa bits cleared)
// and small = x - big (exact in fp32); D += small_a*big_b + big_a*small_b + big_a*big_b keeps ~21 mantissa bits per
// product with fp32 accumulation, which is what the fp32 parity tolerances of tests/ are written against.
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& big, uint32_t& small) {
  big = x & 0xffffe000u;
  small = __float_as_uint(__uint_as_float(x) - __uint_as_float(big)) & 0xffffe000u;
}
__device__ __forceinline__ void mma_m16n8k8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// three independent accumulator chains (big*big, small*big, big*small): consecutive k-steps do not wait for each other's
// MMA latency three times over, and the small correction terms are summed among themselves before they meet the big sum
struct Acc3 { float b[4], s[4], t[4]; };
__device__ __forceinline__ void acc3_zero(Acc3& a) {
#pragma unroll
  for (int j = 0; j < 4; ++j) a.b[j] = a.s[j] = a.t[j] = 0.f;
}
__device__ __forceinline__ void acc3_fold(const Acc3& a, float (&d)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) d[j] = a.b[j] + (a.s[j] + a.t[j]);
}
__device__ __forceinline__ void mma3(Acc3& d, const uint32_t (&ab)[4], const uint32_t (&as)[4], const uint32_t (&bb)[2],
                                     const uint32_t (&bs)[2]) {
  mma_m16n8k8(d.s, as, bb);
  mma_m16n8k8(d.t, ab, bs);
  mma_m16n8k8(d.b, ab, bb);
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

// forward  Z^T[n][m] = sum_k W[n][k] X[m][k]: feature tile `mt` (16 rows of the W slot), NT8 row tiles of 8, k-steps
// [ks0, ks1) of 8 inside the slot.  Slot layouts: X [k-block][xrows][32 floats], W [k-block][wrows][32 floats], both
// written by TMA with the 128-byte swizzle (16-byte chunk c of row r at chunk position c ^ (r & 7)); every ldmatrix
// phase reads 8 rows at the same logical chunk, i.e. 8 different physical chunks = all 32 banks.
// acc[i][0..3] = D(feature g, row 8i+2t), D(g, 8i+2t+1), D(g+8, 8i+2t), D(g+8, 8i+2t+1) with g = lane/4, t = lane%4.
template <int NT8>
__device__ __forceinline__ void mma3_nt(const float* __restrict__ sX, int xrows, const float* __restrict__ sW, int wrows, int mt,
                                        int ks0, int ks1, Acc3 (&acc)[NT8]) {
  const int lane = threadIdx.x & 31, j = lane >> 3, r = lane & 7;
  const uint32_t wrow = smem_u32(sW) + (uint32_t)((mt * 16 + r + 8 * (j & 1)) << 7);
  const uint32_t xrow = smem_u32(sX) + (uint32_t)(r << 7);
  const uint32_t wkb = (uint32_t)wrows << 7, xkb = (uint32_t)xrows << 7;
#pragma unroll 2
  for (int ks = ks0; ks < ks1; ++ks) {
    const uint32_t kb = (uint32_t)ks >> 2, c0 = ((uint32_t)ks & 3u) << 1;
    uint32_t a[4], ab[4], as[4];
    ldsm_x4(a, wrow + kb * wkb + (((c0 + (uint32_t)(j >> 1)) ^ (uint32_t)r) << 4));
#pragma unroll
    for (int q = 0; q < 4; ++q) split_tf32(a[q], ab[q], as[q]);
#pragma unroll
    for (int i = 0; i < NT8; ++i) {
      uint32_t b[2], bb[2], bs[2];
      ldsm_x2(b, xrow + kb * xkb + (uint32_t)(i << 10) + (((c0 + (uint32_t)(j & 1)) ^ (uint32_t)r) << 4));
      split_tf32(b[0], bb[0], bs[0]);
      split_tf32(b[1], bb[1], bs[1]);
      mma3(acc[i], ab, as, bb, bs);
    }
  }
}

// dgrad  dX^T[c][m] = sum_n W[n][c] dZ[m][n]: output-column tile `mt` (16 columns of the CTA's slice), reduction over the
// panel rows n.  Slot layouts: dZ [k-block][xrows][32 floats] (as above), W panels [n/prow][c/32][prow][32 floats]
// (TMA boxes of 32 columns x prow rows, swizzled).  The A fragment is a transposed read, so it uses scalar loads; the
// 8 reduction indices of a k-step are assigned to the fragment slots as (0,2,5,7 | 1,3,4,6), which makes both the panel
// reads (4 different rows -> 4 different chunk pairs) and the dZ reads conflict-free.
template <int NT8, int prow>
__device__ __forceinline__ void mma3_nn(const float* __restrict__ sX, int xrows, const float* __restrict__ sW, int nch,
                                        int mt, int ks0, int ks1, Acc3 (&acc)[NT8]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int klo = (t == 0) ? 0 : (t == 1) ? 2 : (t == 2) ? 5 : 7;
  const int khi = (t == 0) ? 1 : (t == 1) ? 3 : (t == 2) ? 4 : 6;
  const int col = mt * 16 + g;                       // and col + 8 (same 32-column panel: 16 | 32)
  const int cc = col >> 5, cl = col & 31;
  const uint32_t* xw = reinterpret_cast<const uint32_t*>(sX);
  const uint32_t* ww = reinterpret_cast<const uint32_t*>(sW);
#pragma unroll 2
  for (int ks = ks0; ks < ks1; ++ks) {
    const int n_lo = ks * 8 + klo, n_hi = ks * 8 + khi;
    uint32_t a[4], ab[4], as[4];
    {
      const int pl = n_lo / prow, rl = n_lo - pl * prow, ph = n_hi / prow, rh = n_hi - ph * prow;
      const uint32_t* bl = ww + ((size_t)(pl * nch + cc) * prow + rl) * 32;
      const uint32_t* bh = ww + ((size_t)(ph * nch + cc) * prow + rh) * 32;
      a[0] = bl[((((cl) >> 2) ^ (rl & 7)) << 2) + (cl & 3)];
      a[1] = bl[((((cl + 8) >> 2) ^ (rl & 7)) << 2) + (cl & 3)];
      a[2] = bh[((((cl) >> 2) ^ (rh & 7)) << 2) + (cl & 3)];
      a[3] = bh[((((cl + 8) >> 2) ^ (rh & 7)) << 2) + (cl & 3)];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) split_tf32(a[q], ab[q], as[q]);
    const int kb = ks >> 2, f_lo = (ks & 3) * 8 + klo, f_hi = (ks & 3) * 8 + khi;
#pragma unroll
    for (int i = 0; i < NT8; ++i) {
      const uint32_t* xr = xw + ((size_t)kb * xrows + i * 8 + g) * 32;
      uint32_t b[2], bb[2], bs[2];
      b[0] = xr[(((f_lo >> 2) ^ g) << 2) + (f_lo & 3)];
      b[1] = xr[(((f_hi >> 2) ^ g) << 2) + (f_hi & 3)];
      split_tf32(b[0], bb[0], bs[0]);
      split_tf32(b[1], bb[1], bs[1]);
      mma3(acc[i], ab, as, bb, bs);
    }
  }
}

// ---- split-reduction combine: the 8 warps of a CTA are (feature tile mt, reduction part kp); every warp parks its
// partial tile, then all consumer threads add the KS partials of each output in fixed order (deterministic).
// red: [KS][rows][NC + 4] floats.  Result (plus optional bias over columns) lands in sZ[rows][NC] (pitch NC).
template <int NT8>
__device__ __forceinline__ void park_partial(float* red, int kp, int mt, int NC, const float (&acc)[NT8][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, P = NC + 4;
  float* base = red + (size_t)kp * (8 * NT8) * P + mt * 16 + g;
#pragma unroll
  for (int i = 0; i < NT8; ++i) {
    float* r0 = base + (i * 8 + 2 * t) * P;
    r0[0] = acc[i][0];
    r0[P] = acc[i][1];
    r0[8] = acc[i][2];
    r0[P + 8] = acc[i][3];
  }
}
template <int NT8>
__device__ __forceinline__ void combine_partials(const float* red, int KS, int NC, float* sZ, const float* __restrict__ bias_slice,
                                                 int ncv) {
  const int P = NC + 4;
  for (int e = threadIdx.x; e < 8 * NT8 * NC; e += SC_THREADS) {   // consumer threads only (tid < 256)
    const int r = e / NC, c = e - r * NC;
    float s = 0.f;
    for (int w = 0; w < KS; ++w) s += red[((size_t)w * (8 * NT8) + r) * P + c];
    if (bias_slice != nullptr && c < ncv) s += ld_w(bias_slice + c);
    sZ[r * NC + c] = s;
  }
}

}  // namespace sc
