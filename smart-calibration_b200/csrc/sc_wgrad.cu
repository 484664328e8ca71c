// Weight-gradient + Adam + target soft-update kernel (one launch per optimizer group).
//
// Work list = flat list of 32x32 output tiles of dW = dZ^T X for every Linear of the networks being stepped,
// plus 32-column tiles for the vector parameters (biases, LayerNorm weight/bias).  A CTA reduces its tile over
// the whole batch (warps split the batch, partial tiles summed in fixed warp order), then applies
// torch.optim.Adam's update to exactly those parameters and, when a target network is attached, the soft
// update  t <- tau*p + (1-tau)*t  on the same elements.  Every parameter / moment / target element is read and
// written once per step.  Replaces  loss.backward() (weight-gradient part), optimizer.step() and
// update_network_parameters() of enet_sac.py:585-587,597-599,523-542.
#include "sc_common.cuh"
#include "sc_params.h"
#include "sc_tile.cuh"

namespace sc {

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


constexpr int WG_BC = 256;                         // batch rows staged per pass
constexpr int WG_SM_D = WG_BC * 32;
constexpr int WG_SM_X = WG_BC * 32;
constexpr int WG_SM_RED = SC_WARPS * 32 * 36;
constexpr int WG_SM_FLOATS = WG_SM_D + WG_SM_X + WG_SM_RED;

// torch/optim/adam.py::_single_tensor_adam, evaluated with the same roundings as ATen's CPU kernels:
//   exp_avg.lerp_(g, 1-b1)                 -> fma(1-b1, g-m, m)
//   exp_avg_sq.mul_(b2).addcmul_(g,g,1-b2) -> v*b2 + ((1-b2)*g)*g
//   denom = sqrt(v)/sqrt(1-b2^t) + eps ;  p += (-(lr/(1-b1^t)) * m) / denom
__device__ __forceinline__ void adam_apply(float g, float* P, float* M, float* V, float* T, size_t idx,
                                           const AdamHyper& h, float tau, float omt, int do_soft) {
  float m = M[idx], v = V[idx], p = P[idx];
  m = fmaf(h.one_minus_b1, g - m, m);
  v = __fadd_rn(__fmul_rn(v, h.b2), __fmul_rn(__fmul_rn(h.one_minus_b2, g), g));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), h.bc2_sqrt), h.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(h.lr_neg_step, m), denom));
  M[idx] = m;
  V[idx] = v;
  P[idx] = p;
  if (T != nullptr && do_soft) T[idx] = __fadd_rn(__fmul_rn(tau, p), __fmul_rn(omt, T[idx]));
}

// stage rows [b0, b0+nb) x columns [c0, c0+32) of a [B][ld] matrix (ncols valid columns in total) into sm[b][32]
template <bool ACT>
__device__ __forceinline__ void stage_panel(float* sm, const float* src, int ld, int ncols, int c0, int b0, int nb,
                                            const int* __restrict__ gather) {
  const int tid = threadIdx.x;
  const bool vec = ((ld & 3) == 0) && ((ncols & 3) == 0);
  if (vec) {
    // 16-byte cp.async (LDGSTS through L2): every copy of the panel is in flight at once; the caller waits on the group
    for (int e = tid; e < nb * 8; e += SC_THREADS) {
      const int b = e >> 3, c4 = e & 7;
      const bool ok = c0 + 4 * c4 < ncols;
      const size_t row = ok ? (gather ? (size_t)gather[b0 + b] : (size_t)(b0 + b)) : 0;
      const unsigned d = (unsigned)__cvta_generic_to_shared(sm + b * 32 + 4 * c4);
      const float* g = ok ? src + row * ld + c0 + 4 * c4 : src;
      const int sz = ok ? 16 : 0;            // src-size 0 zero-fills the destination
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(g), "r"(sz) : "memory");
    }
  } else {
    // scalar path (layouts that rule out 16-byte copies: N = 1, K = 2, gathered replay rows).  Eight elements per thread are
    // kept in flight: a gathered element costs two dependent L2 round trips, and a panel is 32 elements per thread
    for (int e0 = tid; e0 < nb * 32; e0 += 8 * SC_THREADS) {
      size_t row[8];
      bool ok[8];
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * SC_THREADS, b = e >> 5, c = e & 31;
        ok[u] = e < nb * 32 && c0 + c < ncols;
        row[u] = ok[u] ? (gather ? (size_t)__ldg(gather + b0 + b) : (size_t)(b0 + b)) : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = (e0 + u * SC_THREADS) & 31;
        v[u] = ok[u] ? (ACT ? ld_act(src + row[u] * ld + c0 + c) : ld_w(src + row[u] * ld + c0 + c)) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * SC_THREADS;
        if (e < nb * 32) sm[e] = v[u];                    // sm[b * 32 + c] with e = b * 32 + c
      }
    }
  }
}

__global__ void __launch_bounds__(SC_THREADS, 2) k_wgrad_adam(const WgradParams* __restrict__ pp) {
  __shared__ WgradParams p_sm;
  const WgradParams& p = load_params(pp, &p_sm);              // blockIdx.y = agent of the ensemble
  extern __shared__ __align__(16) float smem[];
  float* sD = smem;
  float* sX = sD + WG_SM_D;
  float* red = sX + WG_SM_X;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int tile = blockIdx.x; tile <= p.total_tiles; tile += gridDim.x) {
    __syncthreads();
    if (tile == p.total_tiles) {
      // ---- loss / diagnostics readout (deterministic block reduction) ----
      if (p.stats_mode == 0) continue;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
      for (int b = tid; b < p.B; b += SC_THREADS) {
        if (p.stats_mode == 1) {
          const float y = ld_act(p.y + b), q1 = ld_act(p.q1 + b), q2 = ld_act(p.q2 + b);
          a0 = fmaf(q1 - y, q1 - y, a0);
          a1 = fmaf(q2 - y, q2 - y, a1);
          a2 += q1; a3 += q2; a4 += y;
        } else {
          const float lp = p.logp ? ld_act(p.logp + b) : 0.f, q = ld_act(p.pi_q + b);
          a0 += p.alpha * lp - q;
          a1 += lp;
        }
      }
      float vals[5] = {a0, a1, a2, a3, a4};
      for (int i = 0; i < 5; ++i) {
        const float w = warp_sum(vals[i]);
        if (lane == 0) red[i * SC_WARPS + warp] = w;
      }
      __syncthreads();
      if (tid < 5) {
        float s = 0.f;
        for (int w = 0; w < SC_WARPS; ++w) s += red[tid * SC_WARPS + w];
        const float invB = 1.f / (float)p.B;
        const float wmean = p.is_wsum ? ld_act(p.is_wsum) * invB : 1.f;
        if (p.stats_mode == 1) {
          if (tid < 2) p.stats[tid] = s * invB * wmean;            // critic_k_loss
          else p.stats[1 + tid] = s * invB;                        // [3] q1_mean [4] q2_mean [5] target_mean
        } else {
          if (tid == 0) p.stats[2] = s * invB * wmean;             // actor_loss
          if (tid == 1) p.stats[6] = s * invB;                     // logp_mean
        }
      }
      continue;
    }
    // ---- the op that owns this tile (host-built table) ----
    const int owner = __ldg(p.tile_map + tile);
    int oi = owner < 1000 ? owner : -1;
    if (oi >= 0) {
      const WgradOp op = p.ops[oi];
      const int lt = tile - op.tile_base;
      const int tn = lt / op.tiles_k, tk = lt - tn * op.tiles_k;
      const int n0 = tn * 32, k0 = tk * 32;
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      for (int b0 = 0; b0 < p.B; b0 += WG_BC) {
        const int nb = min(WG_BC, p.B - b0);
        if (b0 > 0) __syncthreads();
        stage_panel<true>(sD, op.D, op.ldd, op.N, n0, b0, nb, nullptr);
        if (op.gather) stage_panel<false>(sX, op.X, op.ldx, op.K, k0, b0, nb, p.bidx);
        else stage_panel<true>(sX, op.X, op.ldx, op.K, k0, b0, nb, nullptr);
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        mma_tn(sD, sX, (nb * warp) / SC_WARPS, (nb * (warp + 1)) / SC_WARPS, acc);
      }
      {
        const int rg = lane >> 2, cg = lane & 3;
        float* base = red + warp * 32 * 36;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) base[(rg * 4 + i) * 36 + cg * 8 + j] = acc[i][j];
      }
      __syncthreads();
      const AdamHyper h = p.hyper[op.opt];
      // each thread owns 4 elements (rows warp+8u, column lane): all parameter/moment/target loads are issued before
      // any arithmetic so that their HBM/L2 round trips overlap
      {
        const int k = lane;
        const bool soft = op.T != nullptr && p.do_soft;
        float g[4], m[4], v[4], pr[4], tg[4];
        size_t idx[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = warp + 8 * u;
          ok[u] = n0 + n < op.N && k0 + k < op.K;
          idx[u] = (size_t)(n0 + n) * op.ldw + k0 + k;
          if (ok[u]) {
            m[u] = op.M[idx[u]]; v[u] = op.V[idx[u]]; pr[u] = op.W[idx[u]];
            tg[u] = soft ? op.T[idx[u]] : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = warp + 8 * u;
          float s = 0.f;
#pragma unroll
          for (int w = 0; w < SC_WARPS; ++w) s += red[(w * 32 + n) * 36 + k];
          g[u] = s;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!ok[u]) continue;
          const float mm = fmaf(h.one_minus_b1, g[u] - m[u], m[u]);
          const float vv = __fadd_rn(__fmul_rn(v[u], h.b2), __fmul_rn(__fmul_rn(h.one_minus_b2, g[u]), g[u]));
          const float denom = __fadd_rn(__fdiv_rn(sqrtf(vv), h.bc2_sqrt), h.eps);
          const float pp = __fadd_rn(pr[u], __fdiv_rn(__fmul_rn(h.lr_neg_step, mm), denom));
          op.M[idx[u]] = mm;
          op.V[idx[u]] = vv;
          op.W[idx[u]] = pp;
          if (soft) op.T[idx[u]] = __fadd_rn(__fmul_rn(p.tau, pp), __fmul_rn(p.one_minus_tau, tg[u]));
        }
      }
      continue;
    }
    // ---- vector parameter tile ----
    oi = owner - 1000;
    if (oi < 0 || oi >= p.nvops) continue;
    const VecOp vo = p.vops[oi];
    const int n = (tile - vo.tile_base) * 32 + lane;
    float s = 0.f;
    if (n < vo.N) {
      const int b_lo = (p.B * warp) / SC_WARPS, b_hi = (p.B * (warp + 1)) / SC_WARPS;
#pragma unroll 8
      for (int b = b_lo; b < b_hi; ++b) {
        const float d = ld_act(vo.D + (size_t)b * vo.ld + n);
        s = vo.E ? fmaf(d, ld_act(vo.E + (size_t)b * vo.ld + n), s) : s + d;
      }
    }
    red[warp * 32 + lane] = s;
    __syncthreads();
    if (warp == 0 && n < vo.N) {
      float g = 0.f;
#pragma unroll
      for (int w = 0; w < SC_WARPS; ++w) g += red[w * 32 + lane];
      adam_apply(g, vo.P, vo.M, vo.V, vo.T, (size_t)n, p.hyper[vo.opt], p.tau, p.one_minus_tau, p.do_soft);
    }
  }
}

// plain soft update over a contiguous block (Agent.update_network_parameters called on its own:
// construction / load_models hard copy with tau = 1, enet_sac.py:519,645)
__global__ void k_soft_update(const float* __restrict__ src, float* __restrict__ dst, size_t n, float tau, float omt) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __fadd_rn(__fmul_rn(tau, src[i]), __fmul_rn(omt, dst[i]));
}

size_t wgrad_smem_bytes() { return (size_t)WG_SM_FLOATS * sizeof(float); }

cudaError_t launch_wgrad(const WgradParams* pp, int E, int grid, cudaStream_t st) {
  static bool configured[64];
  int d = 0;
  cudaGetDevice(&d);
  d &= 63;
  if (!configured[d]) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_adam, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_smem_bytes());
    if (e != cudaSuccess) return e;
    configured[d] = true;
  }
  k_wgrad_adam<<<dim3(grid, E), SC_THREADS, wgrad_smem_bytes(), st>>>(pp);
  return cudaGetLastError();
}

cudaError_t launch_soft_update(const float* src, float* dst, size_t n, float tau, cudaStream_t st) {
  const float omt = (float)(1.0 - (double)tau);
  int grid = (int)((n + 1023) / 1024);
  if (grid > 1184) grid = 1184;
  k_soft_update<<<grid, 256, 0, st>>>(src, dst, n, tau, omt);
  return cudaGetLastError();
}

}  // namespace sc
