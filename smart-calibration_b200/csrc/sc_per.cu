// HBM-resident replay store, uniform sampling, and the prioritised-replay sum-tree kernels.
//
// The sum tree is the reference's flat float64 heap (root 0, leaves [C-1, 2C-1), enet_sac.py:82-200).
// Bit-reproducibility rules followed here:
//   * all tree arithmetic is IEEE double with explicit _rn intrinsics (no FMA contraction);
//   * a leaf write propagates ONE `change` value to every ancestor (tree[a] += change), samples applied in
//     batch order.  k_per_update reproduces that left fold per ancestor: for stratified batches the tree
//     indices are non-decreasing, so every ancestor's contributors are a contiguous run of the batch and
//     the fold is done by the run's first sample; all (level, run) pairs proceed in parallel.
//     Unsorted index lists take an exact single-thread path.
//   * descent uses `v <= left ? left : (v -= left, right)` (enet_sac.py:188-192).
#include "sc_common.cuh"
#include "sc_params.h"

namespace sc {

// ---------------------------------------------------------------------------------------------------
// store one packed transition [s(S) | s2(S) | a(A) | hint(A) | r | done] into slot `slot`
// priority_mode: 0 none, 1 max-leaf (or clip when the tree is empty), 2 given (f64 change), 3 given (f32 change)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void replay_store_body(const ReplayP& rp, const float* __restrict__ tr, int slot, int priority_mode,
                                                  double priority, double clip) {
  const int tid = threadIdx.x;
  const int S = rp.S, A = rp.A;
  for (int i = tid; i < S; i += blockDim.x) {
    rp.state[(size_t)slot * rp.Sp + i] = tr[i];
    rp.new_state[(size_t)slot * rp.Sp + i] = tr[S + i];
  }
  for (int i = tid; i < A; i += blockDim.x) {
    rp.action[(size_t)slot * rp.Ap + i] = tr[2 * S + i];
    rp.hint[(size_t)slot * rp.Ap + i] = tr[2 * S + A + i];
  }
  if (tid == 0) {
    rp.reward[slot] = tr[2 * S + 2 * A];
    rp.done[slot] = tr[2 * S + 2 * A + 1] != 0.f ? 1 : 0;
  }
  if (priority_mode == 0 || rp.tree == nullptr) return;
  __shared__ double smax[256];
  double p = priority;
  if (priority_mode == 1) {     // np.amax over the leaves (enet_sac.py:239-240)
    double m = 0.0;
    for (int i = tid; i < rp.C; i += blockDim.x) m = fmax(m, rp.tree[rp.C - 1 + i]);
    smax[tid] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
      if (tid < s) smax[tid] = fmax(smax[tid], smax[tid + s]);
      __syncthreads();
    }
    p = smax[0];
    if (p == 0.0) p = clip;
  }
  if (tid == 0) {
    int node = slot + rp.C - 1;
    const double old = rp.tree[node];
    double change = __dsub_rn(p, old);
    if (priority_mode == 3) change = (double)__fsub_rn((float)p, (float)old);
    rp.tree[node] = p;
    while (node != 0) {
      node = (node - 1) >> 1;
      rp.tree[node] = __dadd_rn(rp.tree[node], change);
    }
  }
}
__global__ void k_replay_store(ReplayP rp, const float* __restrict__ tr, int slot, int priority_mode, double priority, double clip) {
  replay_store_body(rp, tr, slot, priority_mode, priority, clip);
}
// the same store as the first launch of a fused store + learn step: parameters (and the packed transition) come from the step tape
__global__ void k_replay_store_t(const StoreP* __restrict__ pp) {
  const StoreP p = pp[blockIdx.y];
  replay_store_body(p.rp, p.tr, p.slot, p.priority_mode, p.priority, p.clip);
}

// bulk fill of slots [0,n) from packed transitions already in HBM
__global__ void k_replay_fill(ReplayP rp, const float* __restrict__ tr, int n) {
  const int T = 2 * rp.S + 2 * rp.A + 2;
  for (int slot = blockIdx.x; slot < n; slot += gridDim.x) {
    const float* t = tr + (size_t)slot * T;
    for (int i = threadIdx.x; i < rp.S; i += blockDim.x) {
      rp.state[(size_t)slot * rp.Sp + i] = t[i];
      rp.new_state[(size_t)slot * rp.Sp + i] = t[rp.S + i];
    }
    for (int i = threadIdx.x; i < rp.A; i += blockDim.x) {
      rp.action[(size_t)slot * rp.Ap + i] = t[2 * rp.S + i];
      rp.hint[(size_t)slot * rp.Ap + i] = t[2 * rp.S + rp.A + i];
    }
    if (threadIdx.x == 0) {
      rp.reward[slot] = t[2 * rp.S + 2 * rp.A];
      rp.done[slot] = t[2 * rp.S + 2 * rp.A + 1] != 0.f ? 1 : 0;
    }
  }
}

// leaves <- priorities (n of them, rest 0), inner nodes <- exact pairwise sums bottom-up.  This is a FRESH build
// (not the reference's insertion-order rounding); tests that need reference bits upload the oracle's tree instead.
__global__ void k_tree_build(double* tree, int C, const double* __restrict__ pr, int n) {
  for (int i = threadIdx.x; i < C; i += blockDim.x) tree[C - 1 + i] = i < n ? pr[i] : 0.0;
  __syncthreads();
  for (int width = C >> 1; width >= 1; width >>= 1) {
    for (int i = threadIdx.x; i < width; i += blockDim.x) {
      const int node = width - 1 + i;
      tree[node] = __dadd_rn(tree[2 * node + 1], tree[2 * node + 2]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// stratified proportional sampling (PER.sample_buffer, enet_sac.py:270-312).  One block.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_per_sample(const PerSampleP* __restrict__ pp) {
  const PerSampleP P = pp[blockIdx.y];
  const double* __restrict__ tree = P.tree;
  const int C = P.C, B = P.B;
  const double* __restrict__ u01 = P.u01;
  const uint64_t seed = P.seed, ctr = P.ctr;
  const double beta = P.beta;
  int *tree_idx = P.tree_idx, *data_idx = P.data_idx;
  float *is_w = P.is_w, *is_wsum = P.is_wsum;
  extern __shared__ float sw[];                  // [B] raw weights
  __shared__ float sred[32];
  __shared__ double stop[511];                   // top 9 levels of the tree: one round of loads instead of 9 dependent ones
  const int tid = threadIdx.x;
  const int ntree = 2 * C - 1;
  const int ntop = ntree < 511 ? ntree : 511;
  for (int i = tid; i < ntop; i += blockDim.x) stop[i] = tree[i];
  __syncthreads();
  const double total = stop[0];
  const double seg = __ddiv_rn(total, (double)B);
  float lmax = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    const double u = u01 ? u01[i] : philox_uniform_double(seed, ctr, (uint64_t)i);
    const double a = __dmul_rn(seg, (double)i), b = __dmul_rn(seg, (double)(i + 1));
    double v = __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), u));
    int node = 0;
    while (true) {
      const int left = 2 * node + 1;
      if (left >= ntree) break;
      const double lv = left < ntop ? stop[left] : tree[left];
      if (v <= lv) node = left;
      else { v = __dsub_rn(v, lv); node = left + 1; }
    }
    const double pr = node < ntop ? stop[node] : tree[node];
    const double prob = __ddiv_rn(pr, total);
    const float w = (float)pow(__dmul_rn((double)B, prob), -beta);
    sw[i] = w;
    lmax = fmaxf(lmax, w);
    tree_idx[i] = node;
    data_idx[i] = node - (C - 1);
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) sred[tid >> 5] = lmax;
  __syncthreads();
  if (tid < 32) {
    float m = tid < (int)(blockDim.x >> 5) ? sred[tid] : 0.f;
    m = warp_max(m);
    if (tid == 0) sred[0] = m;
  }
  __syncthreads();
  const float wmax = sred[0];
  for (int i = tid; i < B; i += blockDim.x) {
    const float w = __fdiv_rn(sw[i], wmax);
    sw[i] = w;
    is_w[i] = w;
  }
  __syncthreads();
  {                          // deterministic (fixed-shape) sum -> mean(w) used by the (B,B)-broadcast losses
    float s = 0.f;
    for (int i = tid; i < B; i += blockDim.x) s += sw[i];
    s = warp_sum(s);
    if ((tid & 31) == 0) sred[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sred[w];
      *is_wsum = tot;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// priority refresh (PER.batch_update, enet_sac.py:314-323).  One block; dynamic smem: B doubles x2 + B ints.
// ---------------------------------------------------------------------------------------------------
// ps_in != nullptr: the float32 priorities were already evaluated by the critic row-phase kernel (err is ignored)
__global__ void __launch_bounds__(1024) k_per_update(const PerUpdateP* __restrict__ pp) {
  const PerUpdateP P = pp[blockIdx.y];
  double* tree = P.tree;
  const int C = P.C, B = P.B;
  const int* __restrict__ idx = P.idx;
  const float* __restrict__ err = P.err;
  const float clip = P.clip;
  float* ps_out = P.ps_out;
  const float* __restrict__ ps_in = P.ps_in;
  extern __shared__ double sd[];
  double* pv = sd;                 // new priorities (widened float32)
  double* chg = sd + B;            // change propagated by sample i
  int* leaf = reinterpret_cast<int*>(sd + 2 * B);
  const int tid = threadIdx.x;
  int ok = 1;
  for (int i = tid; i < B; i += blockDim.x) {
    float p;
    if (ps_in != nullptr) {
      p = __ldcg(ps_in + i);
    } else {
      const float e = fminf(__fadd_rn(err[i], 0.01f), clip);
      // float32 power with a float32 exponent, evaluated in double and rounded once
      p = (float)pow((double)e, (double)0.6f);
    }
    pv[i] = (double)p;
    leaf[i] = idx[i];
    if (ps_out) ps_out[i] = p;
  }
  __syncthreads();
  for (int i = tid; i < B; i += blockDim.x)
    if (i > 0 && leaf[i] < leaf[i - 1]) ok = 0;
  const int sorted = __syncthreads_and(ok);
  if (!sorted) {                   // exact sequential semantics for arbitrary index lists
    if (tid == 0) {
      for (int i = 0; i < B; ++i) {
        int node = leaf[i];
        const double change = __dsub_rn(pv[i], tree[node]);
        tree[node] = pv[i];
        while (node != 0) {
          node = (node - 1) >> 1;
          tree[node] = __dadd_rn(tree[node], change);
        }
      }
    }
    return;
  }
  for (int i = tid; i < B; i += blockDim.x) {
    const double old = (i > 0 && leaf[i - 1] == leaf[i]) ? pv[i - 1] : tree[leaf[i]];
    chg[i] = __dsub_rn(pv[i], old);
  }
  __syncthreads();
  for (int i = tid; i < B; i += blockDim.x)
    if (i == B - 1 || leaf[i + 1] != leaf[i]) tree[leaf[i]] = pv[i];
  int depth = 0;
  while ((1 << depth) < C) ++depth;
  // (level d, sample i): the first sample of each run of equal ancestors folds the run's changes in order.  The run end is
  // found by binary search (ancestors are non-decreasing in i), so the fold itself is a plain counted loop of dependent
  // double adds -- the only inherently sequential part (B adds at the root).
  for (int w = tid; w < depth * B; w += blockDim.x) {
    const int d = w / B + 1, i = w - (d - 1) * B;
    const int node = ((leaf[i] + 1) >> d) - 1;
    if (i > 0 && (((leaf[i - 1] + 1) >> d) - 1) == node) continue;
    int lo = i + 1, hi = B;                      // first j in (i, B] whose ancestor differs
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((((leaf[mid] + 1) >> d) - 1) == node) lo = mid + 1; else hi = mid;
    }
    double val = tree[node];
    int j = i;
#pragma unroll 1
    for (; j + 8 <= lo; j += 8) {
      const double c0 = chg[j], c1 = chg[j + 1], c2 = chg[j + 2], c3 = chg[j + 3], c4 = chg[j + 4], c5 = chg[j + 5], c6 = chg[j + 6],
                   c7 = chg[j + 7];
      val = __dadd_rn(val, c0); val = __dadd_rn(val, c1); val = __dadd_rn(val, c2); val = __dadd_rn(val, c3);
      val = __dadd_rn(val, c4); val = __dadd_rn(val, c5); val = __dadd_rn(val, c6); val = __dadd_rn(val, c7);
    }
    for (; j < lo; ++j) val = __dadd_rn(val, chg[j]);
    tree[node] = val;
  }
}

// ---------------------------------------------------------------------------------------------------
// uniform sampling without replacement on the device: Philox key per candidate, take the B smallest keys
// (n <= 8192: bitonic sort in shared memory; larger: Floyd's algorithm by one thread).  Replaces
// np.random.choice(max_mem, B, replace=False) (enet_sac.py:48) when the caller does not inject indices.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_sample_uniform(const UniformP* __restrict__ pp) {
  const UniformP P = pp[blockIdx.y];
  const int n = P.n, B = P.B;
  const uint64_t seed = P.seed, ctr = P.ctr;
  int* out = P.out;
  extern __shared__ unsigned long long keys[];
  const int tid = threadIdx.x;
  if (n <= 8192) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = tid; i < np2; i += blockDim.x) {
      unsigned long long k = ~0ull;
      if (i < n) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)i, 0u, (uint32_t)ctr, (uint32_t)(ctr >> 32)),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        k = ((unsigned long long)r.x << 32) | (unsigned long long)(uint32_t)i;
      }
      keys[i] = k;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += blockDim.x) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = keys[i], b = keys[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < B; i += blockDim.x) out[i] = (int)(keys[i] & 0xffffffffull);
  } else if (tid == 0) {
    // Floyd: for j = n-B .. n-1: t = U{0..j}; insert t unless already chosen, else j
    int* chosen = out;
    int cnt = 0;
    for (int j = n - B; j < n; ++j) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)j, 1u, (uint32_t)ctr, (uint32_t)(ctr >> 32)),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const int t = (int)((((unsigned long long)r.x << 32) | r.y) % (unsigned long long)(j + 1));
      bool dup = false;
      for (int q = 0; q < cnt; ++q) dup |= (chosen[q] == t);
      chosen[cnt++] = dup ? j : t;
    }
  }
}

// N(0,1) fill for the on-device noise path
__global__ void k_fill_normal(float* out, int n, uint64_t seed, uint64_t stream) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * i < n) {
    const float2 z = philox_normal2(seed, stream, (uint64_t)i);
    out[2 * i] = z.x;
    if (2 * i + 1 < n) out[2 * i + 1] = z.y;
  }
}

// the same fill with the Philox stream id read from device memory (graph-replayed steps: the launch arguments stay constant)
__global__ void k_fill_normal_p(float* out, int n, uint64_t seed, const unsigned long long* __restrict__ stream) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * i < n) {
    const float2 z = philox_normal2(seed, (uint64_t)*stream, (uint64_t)i);
    out[2 * i] = z.x;
    if (2 * i + 1 < n) out[2 * i + 1] = z.y;
  }
}

// hint penalty coefficient (SAC, enet_sac.py:601-602,616-617): m = mean((pi-hint)^2) over B*A,
// g = max(0, m-thr)^2, d(0.5*admm_rho*g^2 + rho*g)/d pi = (admm_rho*g + rho) * 2*max(0,m-thr) * 2(pi-hint)/(B*A).
// Writes out[0] = that scalar coefficient (without the (pi-hint) factor), out[1] = g.  One block.
__global__ void k_hint_coef(const HintCoefP* __restrict__ pp) {
  const HintCoefP P = pp[blockIdx.y];
  const float* __restrict__ pi = P.pi;
  const float* __restrict__ hint = P.hint;
  const int Ap = P.Ap, B = P.B, A = P.A;
  const int* __restrict__ bidx = P.bidx;
  const float thr = P.thr, admm_rho = P.admm_rho;
  const float* rho = P.rho;
  float *out = P.out, *rho_update = P.rho_update;
  __shared__ float sred[32];
  float s = 0.f;
  for (int e = threadIdx.x; e < B * A; e += blockDim.x) {
    const int b = e / A, j = e - b * A;
    const float d = pi[e] - hint[(size_t)bidx[b] * Ap + j];
    s = fmaf(d, d, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sred[w];
    const float m = tot / (float)(B * A);
    const float ex = fmaxf(0.f, m - thr);
    const float g = ex * ex;
    out[0] = (admm_rho * g + *rho) * 2.f * ex * 2.f / (float)(B * A);
    out[1] = g;
    out[2] = 0.5f * admm_rho * g * g + *rho * g;                               // penalty added to actor_loss (enet_sac.py:602)
    float rho_now = *rho;
    if (rho_update != nullptr) {                                               // rho += admm_rho*gfun (enet_sac.py:617)
      rho_now = *rho_update + admm_rho * g;
      *rho_update = rho_now;
    }
    if (P.stats_g != nullptr) { P.stats_g[0] = g; P.stats_g[1] = out[2]; }
    if (P.stats_rho != nullptr) *P.stats_rho = rho_now;                        // the multiplier after the (optional) dual update
  }
}

// TD3 ADMM bookkeeping after one inner actor step (enet_td3.py:331-359).  One block.
//   mode 0: start of the loop: lambda <- 0, rho <- rho0
//   mode 1: lambda += rho*(pi - hint); it == 0 additionally y0 <- pi, a0 <- pi (the reference seeds y0 with the actions);
//           it == 3 additionally the adaptive-penalty test with the three dot products
__global__ void k_admm_update(const AdmmP* __restrict__ pp) {
  const AdmmP P = pp[blockIdx.y];
  const int mode = P.mode, it = P.it, Ap = P.Ap, B = P.B, A = P.A;
  const float* __restrict__ pi = P.pi;
  const float* __restrict__ hint = P.hint;
  const int* __restrict__ bidx = P.bidx;
  float *lambda = P.lambda, *y0 = P.y0, *a0 = P.a0, *rho = P.rho;
  const float rho0 = P.rho0, corr_min = P.corr_min;
  __shared__ float sred[3][32];
  const int n = B * A, tid = threadIdx.x;
  if (mode == 0) {
    for (int e = tid; e < n; e += blockDim.x) lambda[e] = 0.f;
    if (tid == 0) *rho = rho0;
    return;
  }
  const float r = *rho;
  float d11 = 0.f, d12 = 0.f, d22 = 0.f;
  for (int e = tid; e < n; e += blockDim.x) {
    const int b = e / A, j = e - b * A;
    const float a = pi[e];
    const float diff = a - hint[(size_t)bidx[b] * Ap + j];
    const float lam = lambda[e] + r * diff;
    lambda[e] = lam;
    if (it == 0) {
      y0[e] = a;
      a0[e] = a;
    } else if (it == 3) {
      const float y1 = lam + r * diff;
      const float dy = y1 - y0[e], du = a - a0[e];
      d11 = fmaf(dy, dy, d11);
      d12 = fmaf(dy, du, d12);
      d22 = fmaf(du, du, d22);
      y0[e] = y1;
      a0[e] = a;
    }
  }
  if (it != 3) {
    if (tid == 0 && P.stats_rho != nullptr) *P.stats_rho = r;
    return;
  }
  d11 = warp_sum(d11); d12 = warp_sum(d12); d22 = warp_sum(d22);
  if ((tid & 31) == 0) { sred[0][tid >> 5] = d11; sred[1][tid >> 5] = d12; sred[2][tid >> 5] = d22; }
  __syncthreads();
  if (tid == 0) {
    float s11 = 0.f, s12 = 0.f, s22 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s11 += sred[0][w]; s12 += sred[1][w]; s22 += sred[2][w]; }
    float rho_now = r;
    if (s11 > 0.f && s12 > 0.f && s22 > 0.f) {
      const float alpha = s12 / sqrtf(s11 * s22), a_sd = s11 / s12, a_mg = s12 / s22;
      const float a_hat = 2.f * a_mg > a_sd ? a_mg : a_sd - 0.5f * a_mg;
      if (alpha > corr_min && a_hat < 10.f * rho0 && a_hat > 0.1f * rho0) {
        rho_now = a_hat;
        *rho = a_hat;
      }
    }
    if (P.stats_rho != nullptr) *P.stats_rho = rho_now;
  }
}

// ---- host launchers ------------------------------------------------------------------------------------
cudaError_t launch_replay_store(const ReplayP& rp, const float* tr, int slot, int mode, double priority, double clip, cudaStream_t st) {
  k_replay_store<<<1, 256, 0, st>>>(rp, tr, slot, mode, priority, clip);
  return cudaGetLastError();
}
cudaError_t launch_replay_fill(const ReplayP& rp, const float* tr, int n, cudaStream_t st) {
  k_replay_fill<<<n < 592 ? n : 592, 128, 0, st>>>(rp, tr, n);
  return cudaGetLastError();
}
cudaError_t launch_tree_build(double* tree, int C, const double* pr, int n, cudaStream_t st) {
  k_tree_build<<<1, 1024, 0, st>>>(tree, C, pr, n);
  return cudaGetLastError();
}
// per-device one-time kernel attribute set-up (an agent per GPU of one process each needs its own opt-in)
static bool first_use(bool (&done)[64]) {
  int d = 0;
  cudaGetDevice(&d);
  d &= 63;
  if (done[d]) return false;
  done[d] = true;
  return true;
}
cudaError_t launch_replay_store_t(const StoreP* pp, int E, cudaStream_t st) {
  k_replay_store_t<<<dim3(1, E), 256, 0, st>>>(pp);
  return cudaGetLastError();
}
cudaError_t launch_per_sample(const PerSampleP* pp, int E, int B, cudaStream_t st) {
  const int threads = B >= 1024 ? 1024 : ((B + 31) / 32) * 32;
  k_per_sample<<<dim3(1, E), threads, (size_t)B * sizeof(float), st>>>(pp);
  return cudaGetLastError();
}
cudaError_t launch_per_update(const PerUpdateP* pp, int E, int B, cudaStream_t st) {
  static bool configured[64];
  const size_t smem = (size_t)B * (2 * sizeof(double) + sizeof(int));
  if (first_use(configured)) cudaFuncSetAttribute(k_per_update, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_per_update<<<dim3(1, E), 1024, smem, st>>>(pp);
  return cudaGetLastError();
}
cudaError_t launch_sample_uniform(const UniformP* pp, int E, int n, cudaStream_t st) {
  static bool configured[64];
  if (first_use(configured)) cudaFuncSetAttribute(k_sample_uniform, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8);
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  const size_t smem = n <= 8192 ? (size_t)np2 * 8 : 0;
  k_sample_uniform<<<dim3(1, E), 1024, smem, st>>>(pp);
  return cudaGetLastError();
}
cudaError_t launch_fill_normal(float* out, int n, uint64_t seed, uint64_t stream, cudaStream_t st) {
  const int pairs = (n + 1) / 2;
  k_fill_normal<<<(pairs + 255) / 256, 256, 0, st>>>(out, n, seed, stream);
  return cudaGetLastError();
}
cudaError_t launch_fill_normal_p(float* out, int n, uint64_t seed, const unsigned long long* stream_dev, cudaStream_t st) {
  const int pairs = (n + 1) / 2;
  k_fill_normal_p<<<(pairs + 255) / 256, 256, 0, st>>>(out, n, seed, stream_dev);
  return cudaGetLastError();
}
cudaError_t launch_hint_coef(const HintCoefP* pp, int E, cudaStream_t st) {
  k_hint_coef<<<dim3(1, E), 256, 0, st>>>(pp);
  return cudaGetLastError();
}
cudaError_t launch_admm_update(const AdmmP* pp, int E, cudaStream_t st) {
  k_admm_update<<<dim3(1, E), 256, 0, st>>>(pp);
  return cudaGetLastError();
}

}  // namespace sc
