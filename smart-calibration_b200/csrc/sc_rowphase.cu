// Cluster "row-phase" kernel: everything of one SAC/TD3 update that is independent across batch rows.
//
// One thread-block cluster (CS = 8 or 16 CTAs) owns a slab of batch rows and walks the whole layer chain
// for them with activations staged through L2 and only cluster barriers in between (no grid-wide sync):
//   * wide layers (fc1/fc11: S->512, fc2/fc12: 512->256, and the matching dgrads) are "cluster-tiled":
//     CTA `rank` computes output columns [rank*Nc, rank*Nc+Nc) for all slab rows with the FP32 register-tiled
//     GEMM of sc_tile.cuh, LayerNorm statistics are exchanged as per-CTA (mean, M2) partials;
//   * the narrow tail (actor fc3 + heads + tanh-Gaussian sampling, critic action branch, fc3 -> Q, TD target,
//     TD error and the backward pass down to dz of the 256/128/64-wide layers) is "row-local": CTA `rank`
//     carries rows rank, rank+CS, ... through it as GEMVs with warp-shuffle reductions.
// Programs: PROG_CRITIC (target + critic forward + critic dgrad), PROG_ACTOR (actor forward + Q(s,pi(s)) +
// actor dgrad), PROG_ACT (choose_action, one row).  Replaces the torch-eager bodies of
// enet_sac.py:569-585,589-598 / enet_td3.py:247-288,303-308 (forward+backward) -- the weight gradients,
// Adam and the target soft update are in sc_wgrad.cu.
#include "sc_common.cuh"
#include "sc_params.h"
#include "sc_tile.cuh"

namespace sc {

constexpr int MAXOPS = 8;
constexpr int NPMAX = 16;          // max LayerNorm partials per row (CS * column chunks)
constexpr int MAXROWS = 128;       // max slab rows per cluster

struct DenseOp {
  const float *W, *bias, *gamma, *beta;
  int K, N, ldw;
  const float* X; int ldx; int gather;   // gather 1: X rows indexed by bidx (replay array)
  float *H, *ZH, *RS; int ldo;
};
struct DgradOp {
  const float* DZ; int Nl, lddz;
  const float* W; int ldw, Kin;
  const float *Hp, *ZHp, *RSp, *gp;      // layer that produced the input (width Kin)
  float *DU, *DZo; int ldo;
};

// dynamic shared memory map (floats)
constexpr int SM_A = SC_MT * 516;                 // activation tile [32][pitch<=516]
constexpr int SM_W = SC_MT * 516;                 // weight chunk [<=32][pitch] or dgrad panel [<=512][32]
constexpr int SM_RED = SC_WARPS * 32 * 36;        // split-reduction partial tiles
constexpr int SM_Z = 32 * 32;
constexpr int SM_STAT = MAXOPS * MAXROWS * 2;
constexpr int SM_TOTAL_FLOATS = SM_A + SM_W + SM_RED + SM_Z + SM_STAT;

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

// ======================================================================================================
// cluster-tiled forward stage
// ======================================================================================================
template <int TN>
__device__ __forceinline__ void dense_tile(const float* sA, int pa, const float* sW, int pw, int K4, float* red,
                                           float* sZ, const float* bias_slice, int ncv) {
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int warp = threadIdx.x >> 5;
  mma_nt<TN>(sA, pa, sW, pw, (K4 * warp) / SC_WARPS, (K4 * (warp + 1)) / SC_WARPS, acc);
  park_partial<TN, false>(red, acc);
  __syncthreads();
  combine_partials<TN>(red, sZ, bias_slice, ncv);
  __syncthreads();
}

template <int TN>
__device__ __forceinline__ void dgrad_tile(const float* sA, int pa, const float* sW, int Nq, float* red, float* sZ) {
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int warp = threadIdx.x >> 5;
  mma_nn<TN>(sA, pa, sW, 4 * TN, (Nq * warp) / SC_WARPS, (Nq * (warp + 1)) / SC_WARPS, acc);
  park_partial<TN, true>(red, acc);
  __syncthreads();
  combine_partials<TN>(red, sZ, nullptr, 0);
  __syncthreads();
}

__device__ __forceinline__ int part_count(int N, int Nc, int nchunks, int pidx) {
  const int rk = pidx / nchunks, ch = pidx - rk * nchunks;
  const int start = rk * Nc + ch * 32;
  const int width = min(32, Nc - ch * 32);
  return max(0, min(width, N - start));
}

template <int CS>
__device__ void dense_stage(const DenseOp* ops, int nops, int B, int row0, int nrows, const int* __restrict__ bidx,
                            float* part, float* smem) {
  float* sA = smem;
  float* sW = sA + SM_A;
  float* red = sW + SM_W;
  float* sZ = red + SM_RED;
  float* stat = sZ + SM_Z;
  const int rank = (int)cluster_ctarank();
  const int tid = threadIdx.x;
  const int nsub = (nrows + SC_MT - 1) / SC_MT;
  const float* curX = nullptr;
  int curSub = -1;

  for (int o = 0; o < nops; ++o) {
    const DenseOp op = ops[o];
    const int Nc = round4((op.N + CS - 1) / CS);
    const int nchunks = (Nc + 31) / 32;
    const int pk = smem_pitch(op.K);
    const int K4 = op.K >> 2;
    for (int ch = 0; ch < nchunks; ++ch) {
      const int n0 = rank * Nc + ch * 32;
      const int width = min(32, Nc - ch * 32);                 // multiple of 4
      const int ncv = max(0, min(width, op.N - n0));
      __syncthreads();                                          // previous users of sW / sA are done
      stage_rows<false>(sW, pk, width, ncv, K4, [&](int r) { return op.W + (size_t)(n0 + r) * op.ldw; });
      for (int sub = 0; sub < nsub; ++sub) {
        const int r0 = row0 + sub * SC_MT;
        const int rv = min(SC_MT, nrows - sub * SC_MT);
        if (!(curX == op.X && curSub == sub)) {
          if (op.gather)
            stage_rows<false>(sA, pk, SC_MT, rv, K4, [&](int r) { return op.X + (size_t)bidx[r0 + r] * op.ldx; });
          else
            stage_rows<true>(sA, pk, SC_MT, rv, K4, [&](int r) { return op.X + (size_t)(r0 + r) * op.ldx; });
          curX = op.X;
          curSub = sub;
        }
        __syncthreads();
        const float* bias_slice = op.bias ? op.bias + n0 : nullptr;
        switch (width >> 2) {
          case 8: dense_tile<8>(sA, pk, sW, pk, K4, red, sZ, bias_slice, ncv); break;
          case 4: dense_tile<4>(sA, pk, sW, pk, K4, red, sZ, bias_slice, ncv); break;
          case 2: dense_tile<2>(sA, pk, sW, pk, K4, red, sZ, bias_slice, ncv); break;
          default: dense_tile<1>(sA, pk, sW, pk, K4, red, sZ, bias_slice, ncv); break;
        }
        // raw z -> H (own column slice) and per-row (mean, M2) over the valid columns of this chunk
        {
          const int row = tid >> 3, sub8 = tid & 7;
          float s = 0.f;
          for (int c = sub8; c < ncv; c += 8) s += sZ[row * width + c];
          s = group8_sum(s);
          const float mean = ncv > 0 ? s / (float)ncv : 0.f;
          float m2 = 0.f;
          for (int c = sub8; c < ncv; c += 8) {
            const float d = sZ[row * width + c] - mean;
            m2 = fmaf(d, d, m2);
          }
          m2 = group8_sum(m2);
          if (row < rv) {
            float* hrow = op.H + (size_t)(r0 + row) * op.ldo + n0;
            for (int c = sub8; c < ncv; c += 8) hrow[c] = sZ[row * width + c];
            if (sub8 == 0) {
              float* pp = part + (((size_t)o * B + (r0 + row)) * NPMAX + (rank * nchunks + ch)) * 2;
              pp[0] = mean;
              pp[1] = m2;
            }
          }
        }
      }
    }
  }
  cluster_sync();
  // ---- combine the LayerNorm partials of every (op,row) in fixed order (Chan et al.) ----
  for (int i = tid; i < nops * nrows; i += SC_THREADS) {
    const int o = i / nrows, r = i - o * nrows;
    const DenseOp& op = ops[o];
    const int Nc = round4((op.N + CS - 1) / CS);
    const int nchunks = (Nc + 31) / 32;
    const int np = CS * nchunks;
    const float* pp = part + (((size_t)o * B + (row0 + r)) * NPMAX) * 2;
    float msum = 0.f;
    for (int p = 0; p < np; ++p) {
      const int cnt = part_count(op.N, Nc, nchunks, p);
      if (cnt > 0) msum = fmaf((float)cnt, ld_act(pp + 2 * p), msum);
    }
    const float mean = msum / (float)op.N;
    float m2 = 0.f;
    for (int p = 0; p < np; ++p) {
      const int cnt = part_count(op.N, Nc, nchunks, p);
      if (cnt > 0) {
        const float d = ld_act(pp + 2 * p) - mean;
        m2 += ld_act(pp + 2 * p + 1) + (float)cnt * d * d;
      }
    }
    const float rstd = 1.0f / sqrtf(m2 / (float)op.N + SC_LN_EPS);
    stat[2 * i] = mean;
    stat[2 * i + 1] = rstd;
    if (rank == 0 && op.RS != nullptr) op.RS[row0 + r] = rstd;
  }
  __syncthreads();
  for (int o = 0; o < nops; ++o) {
    const DenseOp& op = ops[o];
    const int Nc = round4((op.N + CS - 1) / CS);
    const int n0 = rank * Nc;
    for (int e = tid; e < nrows * Nc; e += SC_THREADS) {
      const int r = e / Nc, c = e - r * Nc, n = n0 + c;
      if (n < op.N) {
        const size_t off = (size_t)(row0 + r) * op.ldo + n;
        const float z = ld_act(op.H + off);
        const float zh = (z - stat[2 * (o * nrows + r)]) * stat[2 * (o * nrows + r) + 1];
        const float u = fmaf(zh, ld_w(op.gamma + n), ld_w(op.beta + n));
        op.H[off] = elu(u);
        if (op.ZH != nullptr) op.ZH[off] = zh;
      }
    }
  }
  cluster_sync();
}

// ======================================================================================================
// cluster-tiled dgrad stage: dH_in = DZ * W, then back through ELU and LayerNorm of the producing layer
// ======================================================================================================
template <int CS>
__device__ void dgrad_stage(const DgradOp* ops, int nops, int B, int row0, int nrows, float* part, float* smem) {
  float* sA = smem;
  float* sW = sA + SM_A;
  float* red = sW + SM_W;
  float* sZ = red + SM_RED;
  float* stat = sZ + SM_Z;
  const int rank = (int)cluster_ctarank();
  const int tid = threadIdx.x;
  const int nsub = (nrows + SC_MT - 1) / SC_MT;

  for (int o = 0; o < nops; ++o) {
    const DgradOp op = ops[o];
    const int Kc = round4((op.Kin + CS - 1) / CS);
    const int nchunks = (Kc + 31) / 32;
    const int pa = smem_pitch(op.Nl);
    for (int ch = 0; ch < nchunks; ++ch) {
      const int k0 = rank * Kc + ch * 32;
      const int width = min(32, Kc - ch * 32);
      const int kcv = max(0, min(width, op.Kin - k0));
      __syncthreads();
      // weight panel sW[n][width]: row n of W restricted to columns [k0, k0+width)
      {
        const int w4 = width >> 2;
        for (int e = tid; e < op.Nl * w4; e += SC_THREADS) {
          const int n = e / w4, c4 = e - n * w4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (4 * c4 < kcv) v = ld_w4(op.W + (size_t)n * op.ldw + k0 + 4 * c4);   // Kin % 4 == 0
          *reinterpret_cast<float4*>(sW + n * width + 4 * c4) = v;
        }
      }
      for (int sub = 0; sub < nsub; ++sub) {
        const int r0 = row0 + sub * SC_MT;
        const int rv = min(SC_MT, nrows - sub * SC_MT);
        if (sub > 0) __syncthreads();
        stage_rows<true>(sA, pa, SC_MT, rv, op.Nl >> 2, [&](int r) { return op.DZ + (size_t)(r0 + r) * op.lddz; });
        __syncthreads();
        switch (width >> 2) {
          case 8: dgrad_tile<8>(sA, pa, sW, op.Nl >> 2, red, sZ); break;
          case 4: dgrad_tile<4>(sA, pa, sW, op.Nl >> 2, red, sZ); break;
          case 2: dgrad_tile<2>(sA, pa, sW, op.Nl >> 2, red, sZ); break;
          default: dgrad_tile<1>(sA, pa, sW, op.Nl >> 2, red, sZ); break;
        }
        {
          const int row = tid >> 3, sub8 = tid & 7;
          float s1 = 0.f, s2 = 0.f;
          if (row < rv) {
            const size_t base = (size_t)(r0 + row) * op.ldo + k0;
            for (int c = sub8; c < kcv; c += 8) {
              const float dh = sZ[row * width + c];
              const float h = ld_act(op.Hp + base + c);
              const float zh = ld_act(op.ZHp + base + c);
              const float du = dh * elu_grad_from_out(h);
              const float t = du * ld_w(op.gp + k0 + c);
              s1 += t;
              s2 = fmaf(t, zh, s2);
              op.DU[base + c] = du;
              op.DZo[base + c] = t;
            }
          }
          s1 = group8_sum(s1);
          s2 = group8_sum(s2);
          if (row < rv && sub8 == 0) {
            float* pp = part + (((size_t)o * B + (r0 + row)) * NPMAX + (rank * nchunks + ch)) * 2;
            pp[0] = s1;
            pp[1] = s2;
          }
        }
      }
    }
  }
  cluster_sync();
  for (int i = tid; i < nops * nrows; i += SC_THREADS) {
    const int o = i / nrows, r = i - o * nrows;
    const DgradOp& op = ops[o];
    const int Kc = round4((op.Kin + CS - 1) / CS);
    const int nchunks = (Kc + 31) / 32;
    const float* pp = part + (((size_t)o * B + (row0 + r)) * NPMAX) * 2;
    float s1 = 0.f, s2 = 0.f;
    for (int p = 0; p < CS * nchunks; ++p) {
      if (part_count(op.Kin, Kc, nchunks, p) > 0) {
        s1 += ld_act(pp + 2 * p);
        s2 += ld_act(pp + 2 * p + 1);
      }
    }
    stat[2 * i] = s1 / (float)op.Kin;
    stat[2 * i + 1] = s2 / (float)op.Kin;
  }
  __syncthreads();
  for (int o = 0; o < nops; ++o) {
    const DgradOp& op = ops[o];
    const int Kc = round4((op.Kin + CS - 1) / CS);
    const int k0 = rank * Kc;
    for (int e = tid; e < nrows * Kc; e += SC_THREADS) {
      const int r = e / Kc, c = e - r * Kc, k = k0 + c;
      if (k < op.Kin) {
        const size_t off = (size_t)(row0 + r) * op.ldo + k;
        const float t = ld_act(op.DZo + off);
        const float zh = ld_act(op.ZHp + off);
        const float rs = ld_act(op.RSp + row0 + r);
        op.DZo[off] = rs * (t - stat[2 * (o * nrows + r)] - zh * stat[2 * (o * nrows + r) + 1]);
      }
    }
  }
  // no trailing cluster barrier: consumers are in the next kernel
}

// ======================================================================================================
// row-local helpers (RT rows at a time, vectors in shared memory)
// ======================================================================================================
// zout[t][n] = bias[n] + sum_k W[n][k] * xin[t][k]; one warp per output feature, lanes along k (coalesced W rows)
__device__ void gemv_rows(const float* __restrict__ W, int ldw, const float* __restrict__ bias, int N, int K,
                          const float* xin, int ldx, float* zout, int ldz) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K4 = K >> 2;
  for (int n0 = warp * 4; n0 < N; n0 += SC_WARPS * 4) {
    float acc[4][RT];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < RT; ++t) acc[u][t] = 0.f;
    for (int k4 = lane; k4 < K4; k4 += 32) {
      float4 x[RT];
#pragma unroll
      for (int t = 0; t < RT; ++t) x[t] = *reinterpret_cast<const float4*>(xin + t * ldx + 4 * k4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (n0 + u < N) {
          const float4 w = ld_w4(W + (size_t)(n0 + u) * ldw + 4 * k4);
#pragma unroll
          for (int t = 0; t < RT; ++t) acc[u][t] = dot4(w, x[t], acc[u][t]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int t = 0; t < RT; ++t) acc[u][t] = warp_sum(acc[u][t]);
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (n0 + u < N) {
          const float b = bias ? ld_w(bias + n0 + u) : 0.f;
#pragma unroll
          for (int t = 0; t < RT; ++t) zout[t * ldz + n0 + u] = acc[u][t] + b;
        }
    }
  }
}

// out[t][k] = sum_n din[t][n] * W[n][k]  (k < K, K % 4 == 0, K/4 <= 256); scratch: [256/K4][RT][K] floats.
// Ends with a __syncthreads(); `out` is complete afterwards.
__device__ void gemvT_rows(const float* __restrict__ W, int ldw, int N, int K, const float* din, int ldd, float* out,
                           int ldo, float* scratch) {
  const int K4 = K >> 2;
  const int parts = SC_THREADS / K4;
  const int tid = threadIdx.x;
  const int part = tid / K4, k4 = tid - part * K4;
  if (part < parts) {
    float4 acc[RT];
#pragma unroll
    for (int t = 0; t < RT; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int n = part; n < N; n += parts) {
      const float4 w = ld_w4(W + (size_t)n * ldw + 4 * k4);
#pragma unroll
      for (int t = 0; t < RT; ++t) {
        const float d = din[t * ldd + n];
        acc[t].x = fmaf(d, w.x, acc[t].x);
        acc[t].y = fmaf(d, w.y, acc[t].y);
        acc[t].z = fmaf(d, w.z, acc[t].z);
        acc[t].w = fmaf(d, w.w, acc[t].w);
      }
    }
#pragma unroll
    for (int t = 0; t < RT; ++t) *reinterpret_cast<float4*>(scratch + ((size_t)(part * RT + t)) * K + 4 * k4) = acc[t];
  }
  __syncthreads();
  for (int e = tid; e < RT * K; e += SC_THREADS) {
    const int t = e / K, k = e - t * K;
    float s = 0.f;
    for (int p = 0; p < parts; ++p) s += scratch[((size_t)(p * RT + t)) * K + k];
    out[t * ldo + k] = s;
  }
  __syncthreads();
}

// in place z[t][0:N] -> ELU(LN(z)); optionally keeps the normalised value and 1/std.  One warp per row.
__device__ void ln_elu_rows(float* z, int ld, int N, const float* __restrict__ gamma, const float* __restrict__ beta,
                            float* zh, float* rs, int rs_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < RT) {
    float* zr = z + warp * ld;
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s += zr[j];
    const float mean = warp_sum(s) / (float)N;
    float m2 = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float d = zr[j] - mean;
      m2 = fmaf(d, d, m2);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(m2) / (float)N + SC_LN_EPS);
    for (int j = lane; j < N; j += 32) {
      const float n = (zr[j] - mean) * rstd;
      if (zh != nullptr) zh[warp * ld + j] = n;
      zr[j] = elu(fmaf(n, ld_w(gamma + j), ld_w(beta + j)));
    }
    if (lane == 0 && rs != nullptr) rs[warp * rs_stride] = rstd;
  }
}

// dh -> du = dh * ELU'(.), dz = rstd * (du*gamma - mean(du*gamma) - zh * mean(du*gamma*zh)).  One warp per row.
__device__ void ln_bwd_rows(const float* dh, const float* h, const float* zh, int ld, int N,
                            const float* __restrict__ gamma, const float* rs, int rs_stride, float* du, float* dz) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < RT) {
    const int o = warp * ld;
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float g = dh[o + j] * elu_grad_from_out(h[o + j]);
      du[o + j] = g;
      const float t = g * ld_w(gamma + j);
      s1 += t;
      s2 = fmaf(t, zh[o + j], s2);
    }
    const float c1 = warp_sum(s1) / (float)N, c2 = warp_sum(s2) / (float)N;
    const float r = rs[warp * rs_stride];
    for (int j = lane; j < N; j += 32) dz[o + j] = r * (du[o + j] * ld_w(gamma + j) - c1 - zh[o + j] * c2);
  }
}

struct RowSet {          // the RT rows a CTA is carrying through the tail
  int row[RT];           // global batch row, or -1
};

__device__ __forceinline__ void load_rows_act(float* dst, int ld, const float* src, int lds, int N, const RowSet& rs) {
  for (int e = threadIdx.x; e < RT * (N >> 2); e += SC_THREADS) {
    const int t = e / (N >> 2), c4 = e - t * (N >> 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rs.row[t] >= 0) v = ld_act4(src + (size_t)rs.row[t] * lds + 4 * c4);
    *reinterpret_cast<float4*>(dst + t * ld + 4 * c4) = v;
  }
}
__device__ __forceinline__ void store_rows(float* dst, int ldd, const float* src, int ld, int N, const RowSet& rs) {
  if (dst == nullptr) return;
  for (int e = threadIdx.x; e < RT * (N >> 2); e += SC_THREADS) {
    const int t = e / (N >> 2), c4 = e - t * (N >> 2);
    if (rs.row[t] >= 0)
      *reinterpret_cast<float4*>(dst + (size_t)rs.row[t] * ldd + 4 * c4) = *reinterpret_cast<const float4*>(src + t * ld + 4 * c4);
  }
}

// shared-memory map of the tail (floats); lives in the same dynamic buffer as the GEMM tiles
struct TailSm {
  float *xh2, *zh2r;              // [RT][H2]   actor layer-2 output row (+ normalised, actor phase)
  float *x3, *zh3;                // [RT][H3]
  float *head;                    // [RT][2*AMAX] mu | raw logsigma
  float *an;                      // [RT][AMAX] action fed to the critics
  float *smp;                     // [RT][AMAX][4]: tanh(u), sigma, eps, clamp-pass mask
  float *sc;                      // [RT][16] per-row scalars
  float *rs;                      // [RT][8] 1/std scratch
  float *h12r, *zh12r;            // [RT][H2]
  float *v21, *zh21;              // [2][RT][HA1] (per critic)
  float *v22, *zh22;              // [2][RT][HA2]
  float *ga, *gb, *gc;            // [RT][H1] gradient vectors
  float *scratch;                 // [256*4*RT] gemvT partials
};
__device__ TailSm carve_tail(float* sm) {
  TailSm t;
  t.xh2 = sm; sm += RT * H2;
  t.zh2r = sm; sm += RT * H2;
  t.x3 = sm; sm += RT * H3;
  t.zh3 = sm; sm += RT * H3;
  t.head = sm; sm += RT * 2 * AMAX;
  t.an = sm; sm += RT * AMAX;
  t.smp = sm; sm += RT * AMAX * 4;
  t.sc = sm; sm += RT * 16;
  t.rs = sm; sm += RT * 8;
  t.h12r = sm; sm += RT * H2;
  t.zh12r = sm; sm += RT * H2;
  t.v21 = sm; sm += 2 * RT * HA1;
  t.zh21 = sm; sm += 2 * RT * HA1;
  t.v22 = sm; sm += 2 * RT * HA2;
  t.zh22 = sm; sm += 2 * RT * HA2;
  t.ga = sm; sm += RT * H1;
  t.gb = sm; sm += RT * H1;
  t.gc = sm; sm += RT * H1;
  t.scratch = sm; sm += SC_THREADS * 4 * RT;
  return t;
}

enum { SC_LOGP = 0, SC_QT1 = 1, SC_QT2 = 2, SC_Y = 3, SC_Q1 = 4, SC_Q2 = 5, SC_DQ1 = 6, SC_DQ2 = 7, SC_GLP = 8 };

// action branch of one critic on `an`: fc21 -> LN -> ELU -> fc22 -> LN -> ELU.  Leaves h21/zh21/h22/zh22 in the
// given smem rows and 1/std in rs[t*8 + {0,1}].  Ends synchronised.
__device__ void critic_action_branch(const CriticP& c, int A, int Ap, const float* an, float* v21, float* zh21,
                                     float* v22, float* zh22, float* rs) {
  for (int e = threadIdx.x; e < RT * HA1; e += SC_THREADS) {
    const int t = e / HA1, n = e - t * HA1;
    float acc = 0.f;
    for (int j = 0; j < A; ++j) acc = fmaf(ld_w(c.w21 + n * Ap + j), an[t * AMAX + j], acc);
    v21[e] = acc + ld_w(c.b21 + n);
  }
  __syncthreads();
  ln_elu_rows(v21, HA1, HA1, c.g21, c.be21, zh21, rs + 0, 8);
  __syncthreads();
  gemv_rows(c.w22, HA1, c.b22, HA2, HA1, v21, HA1, v22, HA2);
  __syncthreads();
  ln_elu_rows(v22, HA2, HA2, c.g22, c.be22, zh22, rs + 1, 8);
  __syncthreads();
}

// q[t] = b3 + w3[0:H2].h12row[t] + w3[H2:].h22[t]   (one warp per row)
__device__ void critic_q(const CriticP& c, const float* h12r, const float* v22, float* out, int out_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < RT) {
    float s = 0.f;
    for (int k = lane; k < H2; k += 32) s = fmaf(ld_w(c.w3 + k), h12r[warp * H2 + k], s);
    for (int k = lane; k < HA2; k += 32) s = fmaf(ld_w(c.w3 + H2 + k), v22[warp * HA2 + k], s);
    s = warp_sum(s);
    if (lane == 0) out[warp * out_stride] = s + ld_w(c.b3);
  }
}

// actor layer 3 + heads on the rows of `rs`; input rows come from `h2src` ([B][H2]).  Leaves h3/zh3 in t.x3/t.zh3,
// the raw head outputs in t.head, 1/std of layer 3 in t.rs[t*8+2].  Ends synchronised.
__device__ void actor_tail_forward(const RowPhaseParams& p, const TailSm& t, const float* h2src, const RowSet& rs) {
  const ActorP& a = p.actor;
  load_rows_act(t.xh2, H2, h2src, H2, H2, rs);
  __syncthreads();
  gemv_rows(a.w3, H2, a.b3, H3, H2, t.xh2, H2, t.x3, H3);
  __syncthreads();
  ln_elu_rows(t.x3, H3, H3, a.g3, a.be3, t.zh3, t.rs + 2, 8);
  __syncthreads();
  gemv_rows(a.wmu, H3, a.bmu, p.A, H3, t.x3, H3, t.head, 2 * AMAX);
  if (a.wls != nullptr) gemv_rows(a.wls, H3, a.bls, p.A, H3, t.x3, H3, t.head + AMAX, 2 * AMAX);
  __syncthreads();
}

// tanh-squashed Gaussian sample + log-prob (SAC, ref ActorNetwork.sample_normal enet_sac.py:446-466) or the
// deterministic TD3 head.  eps_row(t, j) supplies N(0,1).  Results: t.an (action), t.sc[SC_LOGP], t.smp.
template <typename EpsFn>
__device__ void policy_head(const RowPhaseParams& p, const TailSm& t, const RowSet& rs, bool td3_target, EpsFn eps_of) {
  const int tid = threadIdx.x;
  if (tid < RT * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    float* smp = t.smp + (r * AMAX + j) * 4;
    if (p.kind == 0) {
      const float mu = t.head[r * 2 * AMAX + j];
      const float raw = t.head[r * 2 * AMAX + AMAX + j];
      const float ls = fminf(fmaxf(raw, -20.f), 2.f);
      const float sg = expf(ls);
      const float e = rs.row[r] >= 0 ? eps_of(rs.row[r], j) : 0.f;
      const float u = __fadd_rn(mu, __fmul_rn(e, sg));                 // loc + eps*scale, two roundings like torch
      const float th = tanhf(u);
      const float dm = u - mu;
      // Normal.log_prob: -((v-loc)^2)/(2 var) - log(scale) - log(sqrt(2 pi)); then the tanh correction
      float lp = -(dm * dm) / (2.f * (sg * sg)) - logf(sg) - 0.91893853320467274178f;
      lp -= logf(p.maxa * (1.f - th * th) + 1e-6f);
      t.an[r * AMAX + j] = th * p.maxa;
      smp[0] = th; smp[1] = sg; smp[2] = e; smp[3] = (raw >= -20.f && raw <= 2.f) ? 1.f : 0.f;
      t.gc[r * AMAX + j] = lp;                                          // per-action log-prob, summed below
    } else {
      float act = tanhf(t.head[r * 2 * AMAX + j]);
      smp[0] = act;
      if (td3_target) act = fminf(fmaxf(act + p.td3_noise_c, -1.f), 1.f);   // enet_td3.py:248-251
      t.an[r * AMAX + j] = act;
    }
  }
  __syncthreads();
  if (p.kind == 0 && tid < RT) {
    float s = 0.f;
    for (int j = 0; j < p.A; ++j) s += t.gc[tid * AMAX + j];
    t.sc[tid * 16 + SC_LOGP] = s;
  }
  __syncthreads();
}

// ======================================================================================================
// PROG_CRITIC tail
// ======================================================================================================
__device__ void tail_critic(const RowPhaseParams& p, const TailSm& t, const RowSet& rs) {
  const int tid = threadIdx.x;
  // ---- policy on s' (SAC: online actor, sampled; TD3: target actor + clipped scalar noise) ----
  actor_tail_forward(p, t, p.aw.t_h2, rs);
  policy_head(p, t, rs, true, [&](int row, int j) { return ld_act(p.eps + (size_t)row * p.A + j); });
  // ---- target critics on (s', a') ----
  for (int k = 0; k < 2; ++k) {
    load_rows_act(t.h12r, H2, p.tw[k].h12, H2, H2, rs);
    critic_action_branch(p.targ[k], p.A, p.Ap, t.an, t.v21, t.zh21, t.v22, t.zh22, t.rs);
    critic_q(p.targ[k], t.h12r, t.v22, t.sc + SC_QT1 + k, 16);
    __syncthreads();
  }
  if (tid < RT && rs.row[tid] >= 0) {
    const int row = rs.row[tid];
    const int slot = p.bidx[row];
    float* s = t.sc + tid * 16;
    const bool done = p.rp.done[slot] != 0;
    float yv;
    if (p.kind == 0) {   // enet_sac.py:573-575
      float m = __fsub_rn(fminf(s[SC_QT1], s[SC_QT2]), __fmul_rn(p.alpha, s[SC_LOGP]));
      if (done) m = 0.f;
      yv = __fadd_rn(__fmul_rn(p.scale, ld_w(p.rp.reward + slot)), __fmul_rn(p.gamma, m));
    } else {             // enet_td3.py:255-260
      const float m = done ? 0.f : fminf(s[SC_QT1], s[SC_QT2]);
      yv = __fadd_rn(ld_w(p.rp.reward + slot), __fmul_rn(p.gamma, m));
    }
    s[SC_Y] = yv;
    p.y[row] = yv;
  }
  // ---- online critics on (s, a) with backward ----
  if (tid < RT * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    t.an[r * AMAX + j] = rs.row[r] >= 0 ? ld_w(p.rp.action + (size_t)p.bidx[rs.row[r]] * p.Ap + j) : 0.f;
  }
  __syncthreads();
  const float wsum = p.is_wsum ? ld_act(p.is_wsum) : (float)p.B;
  const float coef = 2.f * wsum / ((float)p.B * (float)p.B);      // d/dq of mean(w)*mean((q-y)^2)
  for (int k = 0; k < 2; ++k) {
    const CriticP& c = p.crit[k];
    const CriticWs& w = p.cw[k];
    load_rows_act(t.h12r, H2, w.h12, H2, H2, rs);
    load_rows_act(t.zh12r, H2, w.zh12, H2, H2, rs);
    critic_action_branch(c, p.A, p.Ap, t.an, t.v21, t.zh21, t.v22, t.zh22, t.rs);
    critic_q(c, t.h12r, t.v22, t.sc + SC_Q1 + k, 16);
    if (tid < RT) t.rs[tid * 8 + 3] = rs.row[tid] >= 0 ? ld_act(w.rs12 + rs.row[tid]) : 0.f;
    __syncthreads();
    if (tid < RT && rs.row[tid] >= 0) {
      float* s = t.sc + tid * 16;
      const int row = rs.row[tid];
      w.q[row] = s[SC_Q1 + k];
      const float dq = coef * (s[SC_Q1 + k] - s[SC_Y]);
      s[SC_DQ1 + k] = dq;
      w.dq[row] = dq;
      w.rs21[row] = t.rs[tid * 8 + 0];
      w.rs22[row] = t.rs[tid * 8 + 1];
    } else if (tid < RT) {
      t.sc[tid * 16 + SC_DQ1 + k] = 0.f;
    }
    store_rows(w.h21, HA1, t.v21, HA1, HA1, rs);
    store_rows(w.zh21, HA1, t.zh21, HA1, HA1, rs);
    store_rows(w.h22, HA2, t.v22, HA2, HA2, rs);
    store_rows(w.zh22, HA2, t.zh22, HA2, HA2, rs);
    __syncthreads();
    // d cat = dq * w3: state part -> ga[0:H2], action part -> gc[0:HA2]
    for (int e = tid; e < RT * HCAT; e += SC_THREADS) {
      const int r = e / HCAT, j = e - r * HCAT;
      const float g = t.sc[r * 16 + SC_DQ1 + k] * ld_w(c.w3 + j);
      if (j < H2) t.ga[r * H2 + j] = g; else t.gc[r * HA2 + (j - H2)] = g;
    }
    __syncthreads();
    ln_bwd_rows(t.ga, t.h12r, t.zh12r, H2, H2, c.g12, t.rs + 3, 8, t.ga, t.gb);
    __syncthreads();
    store_rows(w.du12, H2, t.ga, H2, H2, rs);
    store_rows(w.dz12, H2, t.gb, H2, H2, rs);
    __syncthreads();
    // action branch backward: layer 22 then 21
    ln_bwd_rows(t.gc, t.v22, t.zh22, HA2, HA2, c.g22, t.rs + 1, 8, t.gc, t.ga);
    __syncthreads();
    store_rows(w.du22, HA2, t.gc, HA2, HA2, rs);
    store_rows(w.dz22, HA2, t.ga, HA2, HA2, rs);
    gemvT_rows(c.w22, HA1, HA2, HA1, t.ga, HA2, t.gb, HA1, t.scratch);
    ln_bwd_rows(t.gb, t.v21, t.zh21, HA1, HA1, c.g21, t.rs + 0, 8, t.gb, t.gc);
    __syncthreads();
    store_rows(w.du21, HA1, t.gb, HA1, HA1, rs);
    store_rows(w.dz21, HA1, t.gc, HA1, HA1, rs);
    __syncthreads();
  }
  if (tid < RT && rs.row[tid] >= 0) {
    const float* s = t.sc + tid * 16;
    p.err[rs.row[tid]] = 0.5f * (fabsf(s[SC_Q1] - s[SC_Y]) + fabsf(s[SC_Q2] - s[SC_Y]));   // enet_td3.py:266-269
  }
  __syncthreads();
}

// ======================================================================================================
// PROG_ACTOR tail
// ======================================================================================================
__device__ void tail_actor(const RowPhaseParams& p, const TailSm& t, const RowSet& rs) {
  const int tid = threadIdx.x;
  const ActorP& a = p.actor;
  actor_tail_forward(p, t, p.aw.h2, rs);
  store_rows(p.aw.h3, H3, t.x3, H3, H3, rs);
  store_rows(p.aw.zh3, H3, t.zh3, H3, H3, rs);
  if (tid < RT && rs.row[tid] >= 0) p.aw.rs3[rs.row[tid]] = t.rs[tid * 8 + 2];
  policy_head(p, t, rs, false, [&](int row, int j) { return ld_act(p.eps + (size_t)row * p.A + j); });
  if (tid < RT * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    if (rs.row[r] >= 0) p.aw.pi[(size_t)rs.row[r] * p.A + j] = t.an[r * AMAX + j];
  }
  if (tid < RT && rs.row[tid] >= 0 && p.kind == 0) p.aw.logp[rs.row[tid]] = t.sc[tid * 16 + SC_LOGP];
  // ---- Q_k(s, pi(s)) with the freshly updated critics ----
  for (int k = 0; k < p.n_q; ++k) {
    load_rows_act(t.h12r, H2, p.cw[k].p_h12, H2, H2, rs);
    critic_action_branch(p.crit[k], p.A, p.Ap, t.an, t.v21 + k * RT * HA1, t.zh21 + k * RT * HA1,
                         t.v22 + k * RT * HA2, t.zh22 + k * RT * HA2, t.rs + 4 + 2 * k);
    critic_q(p.crit[k], t.h12r, t.v22 + k * RT * HA2, t.sc + SC_Q1 + k, 16);
    __syncthreads();
    // critic_action_branch wrote 1/std to rs[4+2k], rs[5+2k]
  }
  const float wsum = p.is_wsum ? ld_act(p.is_wsum) : (float)p.B;
  const float coef = wsum / ((float)p.B * (float)p.B);           // d/dx of mean(x*w) with the (B,B) broadcast
  if (tid < RT) {
    float* s = t.sc + tid * 16;
    float g1 = 0.f, g2 = 0.f;
    if (rs.row[tid] >= 0) {
      if (p.n_q == 2) {   // torch.min backward: gradient to the smaller, halves on ties
        if (s[SC_Q1] < s[SC_Q2]) g1 = -coef; else if (s[SC_Q2] < s[SC_Q1]) g2 = -coef; else { g1 = -0.5f * coef; g2 = -0.5f * coef; }
        p.aw.qmin[rs.row[tid]] = fminf(s[SC_Q1], s[SC_Q2]);       // critic_value (enet_sac.py:592)
      } else {
        g1 = -coef;
        p.aw.qmin[rs.row[tid]] = s[SC_Q1];
      }
    }
    s[SC_DQ1] = g1;
    s[SC_DQ2] = g2;
    s[SC_GLP] = rs.row[tid] >= 0 ? p.alpha * coef : 0.f;
  }
  for (int e = tid; e < RT * AMAX; e += SC_THREADS) t.gc[RT * H1 - RT * AMAX + e] = 0.f;   // d loss / d action accumulator
  float* dact = t.gc + RT * H1 - RT * AMAX;                                               // [RT][AMAX]
  __syncthreads();
  for (int k = 0; k < p.n_q; ++k) {
    const CriticP& c = p.crit[k];
    float* v21 = t.v21 + k * RT * HA1; float* zh21 = t.zh21 + k * RT * HA1;
    float* v22 = t.v22 + k * RT * HA2; float* zh22 = t.zh22 + k * RT * HA2;
    for (int e = tid; e < RT * HA2; e += SC_THREADS) {
      const int r = e / HA2, j = e - r * HA2;
      t.ga[r * HA2 + j] = t.sc[r * 16 + SC_DQ1 + k] * ld_w(c.w3 + H2 + j);
    }
    __syncthreads();
    ln_bwd_rows(t.ga, v22, zh22, HA2, HA2, c.g22, t.rs + 5 + 2 * k, 8, t.ga, t.gb);
    __syncthreads();
    gemvT_rows(c.w22, HA1, HA2, HA1, t.gb, HA2, t.ga, HA1, t.scratch);
    ln_bwd_rows(t.ga, v21, zh21, HA1, HA1, c.g21, t.rs + 4 + 2 * k, 8, t.ga, t.gb);
    __syncthreads();
    // d action[j] += sum_n dz21[n] * w21[n][j]   (one warp per (row, j))
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int o = warp; o < RT * p.A; o += SC_WARPS) {
        const int r = o / p.A, j = o - r * p.A;
        float s = 0.f;
        for (int n = lane; n < HA1; n += 32) s = fmaf(t.gb[r * HA1 + n], ld_w(c.w21 + n * p.Ap + j), s);
        s = warp_sum(s);
        if (lane == 0) dact[r * AMAX + j] += s;
      }
    }
    __syncthreads();
  }
  // ---- back through the policy head ----
  if (tid < RT * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    float da = dact[r * AMAX + j];
    const float* smp = t.smp + (r * AMAX + j) * 4;
    const int row = rs.row[r];
    if (p.hint_coef != nullptr && row >= 0)
      da = fmaf(ld_act(p.hint_coef), t.an[r * AMAX + j] - ld_w(p.rp.hint + (size_t)p.bidx[row] * p.Ap + j), da);
    if (p.kind == 0) {
      const float th = smp[0], sg = smp[1], e = smp[2];
      const float omt = 1.f - th * th;
      const float glp = t.sc[r * 16 + SC_GLP];
      const float du = da * p.maxa * omt + glp * (2.f * th * omt * p.maxa / (p.maxa * omt + 1e-6f));
      const float dl = (du * sg * e - glp) * smp[3];
      t.head[r * 2 * AMAX + j] = du;
      t.head[r * 2 * AMAX + AMAX + j] = dl;
      if (row >= 0) {
        p.aw.dhead[(size_t)row * 2 * p.A + j] = du;
        p.aw.dhead[(size_t)row * 2 * p.A + p.A + j] = dl;
      }
    } else {
      const float act = smp[0];
      const float dz = da * (1.f - act * act);
      t.head[r * 2 * AMAX + j] = dz;
      if (row >= 0) p.aw.dhead[(size_t)row * p.A + j] = dz;
    }
  }
  __syncthreads();
  for (int e = tid; e < RT * H3; e += SC_THREADS) {
    const int r = e / H3, k = e - r * H3;
    float s = 0.f;
    for (int j = 0; j < p.A; ++j) {
      s = fmaf(t.head[r * 2 * AMAX + j], ld_w(a.wmu + j * H3 + k), s);
      if (a.wls != nullptr) s = fmaf(t.head[r * 2 * AMAX + AMAX + j], ld_w(a.wls + j * H3 + k), s);
    }
    t.ga[r * H3 + k] = s;
  }
  __syncthreads();
  ln_bwd_rows(t.ga, t.x3, t.zh3, H3, H3, a.g3, t.rs + 2, 8, t.ga, t.gb);
  __syncthreads();
  store_rows(p.aw.du3, H3, t.ga, H3, H3, rs);
  store_rows(p.aw.dz3, H3, t.gb, H3, H3, rs);
  gemvT_rows(a.w3, H2, H3, H2, t.gb, H3, t.ga, H2, t.scratch);
  load_rows_act(t.zh2r, H2, p.aw.zh2, H2, H2, rs);
  if (tid < RT) t.rs[tid * 8 + 3] = rs.row[tid] >= 0 ? ld_act(p.aw.rs2 + rs.row[tid]) : 0.f;
  __syncthreads();
  ln_bwd_rows(t.ga, t.xh2, t.zh2r, H2, H2, a.g2, t.rs + 3, 8, t.ga, t.gb);
  __syncthreads();
  store_rows(p.aw.du2, H2, t.ga, H2, H2, rs);
  store_rows(p.aw.dz2, H2, t.gb, H2, H2, rs);
  __syncthreads();
}

// ======================================================================================================
// the kernel
// ======================================================================================================
template <int CS>
__global__ void __launch_bounds__(SC_THREADS, 1) k_rowphase(const __grid_constant__ RowPhaseParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ DenseOp dops[MAXOPS];
  __shared__ DgradOp gops[2];
  const int tid = threadIdx.x;
  const int rank = (int)cluster_ctarank();
  const int ncl = (int)num_clusters();
  const int cid = (int)cluster_idx();
  const int per = (p.B + ncl - 1) / ncl;
  const int row0 = cid * per;
  const int nrows = max(0, min(per, p.B - row0));
  // (clusters with nrows == 0 still walk the barriers; all their loops are empty)

  auto dense = [&](int i, const float* W, const float* b, const float* g, const float* be, int K, int N, int ldw,
                   const float* X, int ldx, int gather, float* H, float* ZH, float* RS) {
    if (tid == 0) dops[i] = DenseOp{W, b, g, be, K, N, ldw, X, ldx, gather, H, ZH, RS, N};
  };
  const TailSm ts = carve_tail(smem);

  if (p.program == PROG_CRITIC) {
    const ActorP& a = p.actor;
    // stage 1: first layers.  s' feeds the policy and both target critics, s feeds both online critics.
    dense(0, a.w1, a.b1, a.g1, a.be1, p.Sp, H1, p.Sp, p.rp.new_state, p.Sp, 1, p.aw.t_h1, nullptr, nullptr);
    dense(1, p.targ[0].w11, p.targ[0].b11, p.targ[0].g11, p.targ[0].be11, p.Sp, H1, p.Sp, p.rp.new_state, p.Sp, 1, p.tw[0].h11, nullptr, nullptr);
    dense(2, p.targ[1].w11, p.targ[1].b11, p.targ[1].g11, p.targ[1].be11, p.Sp, H1, p.Sp, p.rp.new_state, p.Sp, 1, p.tw[1].h11, nullptr, nullptr);
    dense(3, p.crit[0].w11, p.crit[0].b11, p.crit[0].g11, p.crit[0].be11, p.Sp, H1, p.Sp, p.rp.state, p.Sp, 1, p.cw[0].h11, p.cw[0].zh11, p.cw[0].rs11);
    dense(4, p.crit[1].w11, p.crit[1].b11, p.crit[1].g11, p.crit[1].be11, p.Sp, H1, p.Sp, p.rp.state, p.Sp, 1, p.cw[1].h11, p.cw[1].zh11, p.cw[1].rs11);
    __syncthreads();
    dense_stage<CS>(dops, 5, p.B, row0, nrows, p.bidx, p.part, smem);
    // stage 2: second layers
    dense(0, a.w2, a.b2, a.g2, a.be2, H1, H2, H1, p.aw.t_h1, H1, 0, p.aw.t_h2, nullptr, nullptr);
    dense(1, p.targ[0].w12, p.targ[0].b12, p.targ[0].g12, p.targ[0].be12, H1, H2, H1, p.tw[0].h11, H1, 0, p.tw[0].h12, nullptr, nullptr);
    dense(2, p.targ[1].w12, p.targ[1].b12, p.targ[1].g12, p.targ[1].be12, H1, H2, H1, p.tw[1].h11, H1, 0, p.tw[1].h12, nullptr, nullptr);
    dense(3, p.crit[0].w12, p.crit[0].b12, p.crit[0].g12, p.crit[0].be12, H1, H2, H1, p.cw[0].h11, H1, 0, p.cw[0].h12, p.cw[0].zh12, p.cw[0].rs12);
    dense(4, p.crit[1].w12, p.crit[1].b12, p.crit[1].g12, p.crit[1].be12, H1, H2, H1, p.cw[1].h11, H1, 0, p.cw[1].h12, p.cw[1].zh12, p.cw[1].rs12);
    __syncthreads();
    dense_stage<CS>(dops, 5, p.B, row0, nrows, p.bidx, p.part, smem);
    // stage 3: row-local tail
    for (int g = rank; g < nrows; g += CS * RT) {
      RowSet rs;
#pragma unroll
      for (int t = 0; t < RT; ++t) rs.row[t] = (g + t * CS < nrows) ? row0 + g + t * CS : -1;
      tail_critic(p, ts, rs);
    }
    cluster_sync();
    // stage 4: dz12 -> layer-11 gradients
    if (tid == 0) {
      for (int k = 0; k < 2; ++k)
        gops[k] = DgradOp{p.cw[k].dz12, H2, H2, p.crit[k].w12, H1, H1, p.cw[k].h11, p.cw[k].zh11, p.cw[k].rs11,
                          p.crit[k].g11, p.cw[k].du11, p.cw[k].dz11, H1};
    }
    __syncthreads();
    dgrad_stage<CS>(gops, 2, p.B, row0, nrows, p.part, smem);
  } else if (p.program == PROG_ACTOR) {
    const ActorP& a = p.actor;
    dense(0, a.w1, a.b1, a.g1, a.be1, p.Sp, H1, p.Sp, p.rp.state, p.Sp, 1, p.aw.h1, p.aw.zh1, p.aw.rs1);
    for (int k = 0; k < p.n_q; ++k)
      dense(1 + k, p.crit[k].w11, p.crit[k].b11, p.crit[k].g11, p.crit[k].be11, p.Sp, H1, p.Sp, p.rp.state, p.Sp, 1, p.cw[k].p_h11, nullptr, nullptr);
    __syncthreads();
    dense_stage<CS>(dops, 1 + p.n_q, p.B, row0, nrows, p.bidx, p.part, smem);
    dense(0, a.w2, a.b2, a.g2, a.be2, H1, H2, H1, p.aw.h1, H1, 0, p.aw.h2, p.aw.zh2, p.aw.rs2);
    for (int k = 0; k < p.n_q; ++k)
      dense(1 + k, p.crit[k].w12, p.crit[k].b12, p.crit[k].g12, p.crit[k].be12, H1, H2, H1, p.cw[k].p_h11, H1, 0, p.cw[k].p_h12, nullptr, nullptr);
    __syncthreads();
    dense_stage<CS>(dops, 1 + p.n_q, p.B, row0, nrows, p.bidx, p.part, smem);
    for (int g = rank; g < nrows; g += CS * RT) {
      RowSet rs;
#pragma unroll
      for (int t = 0; t < RT; ++t) rs.row[t] = (g + t * CS < nrows) ? row0 + g + t * CS : -1;
      tail_actor(p, ts, rs);
    }
    cluster_sync();
    if (tid == 0)
      gops[0] = DgradOp{p.aw.dz2, H2, H2, a.w2, H1, H1, p.aw.h1, p.aw.zh1, p.aw.rs1, a.g1, p.aw.du1, p.aw.dz1, H1};
    __syncthreads();
    dgrad_stage<CS>(gops, 1, p.B, row0, nrows, p.part, smem);
  } else {   // PROG_ACT: one observation through the actor (enet_sac.py:547-553)
    const ActorP& a = p.actor;
    dense(0, a.w1, a.b1, a.g1, a.be1, p.Sp, H1, p.Sp, p.obs, p.Sp, 0, p.act_aux, nullptr, nullptr);
    __syncthreads();
    dense_stage<CS>(dops, 1, 1, row0, nrows, nullptr, p.part, smem);
    dense(0, a.w2, a.b2, a.g2, a.be2, H1, H2, H1, p.act_aux, H1, 0, p.act_aux + H1, nullptr, nullptr);
    __syncthreads();
    dense_stage<CS>(dops, 1, 1, row0, nrows, nullptr, p.part, smem);
    if (rank == 0 && nrows > 0) {
      RowSet rs;
      rs.row[0] = 0;
#pragma unroll
      for (int t = 1; t < RT; ++t) rs.row[t] = -1;
      actor_tail_forward(p, ts, p.act_aux + H1, rs);
      policy_head(p, ts, rs, false, [&](int, int j) { return p.eps ? ld_act(p.eps + j) : 0.f; });
      if (tid < p.A) {
        float v = ts.an[tid];
        if (p.kind != 0 && p.eps != nullptr) v = fminf(fmaxf(v + ld_act(p.eps + tid), -1.f), 1.f);   // enet_td3.py:215-217
        p.act_out[tid] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------------
size_t rowphase_smem_bytes() { return (size_t)SM_TOTAL_FLOATS * sizeof(float); }

template <int CS>
static cudaError_t launch_cs(const RowPhaseParams& p, int nclusters, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_rowphase<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rowphase_smem_bytes());
    if (e != cudaSuccess) return e;
    if (CS > 8) {
      e = cudaFuncSetAttribute(k_rowphase<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) return e;
    }
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CS, 1, 1);
  cfg.blockDim = dim3(SC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = rowphase_smem_bytes();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_rowphase<CS>, p);
}

// cluster size to use: 16 when at least `want_clusters` 16-CTA clusters are co-resident, else 8 (portable), else 16
// if it fits at all; 0 if the device cannot host the kernel.
template <int CS>
static int max_active_clusters() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(8 * CS, 1, 1);
  cfg.blockDim = dim3(SC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = rowphase_smem_bytes();
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaFuncSetAttribute(k_rowphase<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rowphase_smem_bytes());
  if (e == cudaSuccess && CS > 8) e = cudaFuncSetAttribute(k_rowphase<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, k_rowphase<CS>, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int rowphase_pick_cluster_size(int want_clusters) {
  const int n16 = max_active_clusters<16>();
  if (n16 >= want_clusters) return 16;
  const int n8 = max_active_clusters<8>();
  if (n8 >= 1) return 8;
  return n16 >= 1 ? 16 : 0;
}

cudaError_t launch_rowphase(const RowPhaseParams& p, int cluster_size, int nclusters, cudaStream_t st) {
  if (cluster_size == 16) return launch_cs<16>(p, nclusters, st);
  return launch_cs<8>(p, nclusters, st);
}

}  // namespace sc
