// Calibration conv-SAC agent (calibration/calib_sac.py) on the GPU: C ABI `sc_calib_*` of include/smartcal_b200.h.
// One learn() = eight trunk forward passes and three trunk backward passes, sequenced
// here out of the kernels of sc_calib_kernels.cuh plus the small per-row kernels below.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/smartcal_b200.h"
#include "sc_calib_kernels.cuh"
#include "sc_conv_tc.cuh"
#include "sc_params.h"

namespace sc {
cudaError_t launch_soft_update(const float* src, float* dst, size_t n, float tau, cudaStream_t st);
cudaError_t launch_fill_normal(float* out, int n, uint64_t seed, uint64_t stream, cudaStream_t st);
cudaError_t launch_fill_normal_p(float* out, int n, uint64_t seed, const unsigned long long* stream_dev, cudaStream_t st);
int encode_map_f32_2d(void* out, const void* base, int ld, int rows, int box_rows);      // sc_api.cu
cudaError_t launch_per_sample(const PerSampleP* pp, int E, int B, cudaStream_t st);        // sc_per.cu
cudaError_t launch_per_update(const PerUpdateP* pp, int E, int B, cudaStream_t st);
cudaError_t launch_admm_update(const AdmmP* pp, int E, cudaStream_t st);
}  // namespace sc

using namespace sc;
using namespace sc::calib;

extern "C" const char* sc_last_error(void);
extern "C" void sc_set_last_error_(const char* msg);     // sc_api.cu

static int failc(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  sc_set_last_error_(buf);
  return code;
}
#define CKC(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return failc(SC_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

// ---- geometry of the reference networks (calib_sac.py:100-117,163-183 with input_dims = [1,128,128]) -----------------
constexpr int IMG = 128, H1C = 62, H2C = 29, H3C = 13;
constexpr int C1 = 16, C2 = 32, C3 = 32;
constexpr int FLAT = C3 * H3C * H3C;      // 5408
constexpr int FEAT = FLAT + 16;           // 5424
constexpr int HID = 128, SIDE = 16;
constexpr int AMAXC = 128;                // max n_actions (demixing_fuzzy: 24 (K-1) + 8 = 128)
constexpr int SIDEMAX = 32;               // INNER_DIM of demixing_fuzzy (16 for calibration / demixing_rl)

// =====================================================================================================================
// small per-row kernels (one block per batch row unless noted)
// =====================================================================================================================
namespace {

// dst[b][:] = src[idx[b]][:]
__global__ void k_gather(const float* __restrict__ src, const int* __restrict__ idx, size_t W, float* __restrict__ dst) {
  const size_t b = blockIdx.y;
  const float* s = src + (size_t)idx[b] * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < W; i += (size_t)gridDim.x * blockDim.x) dst[b * W + i] = s[i];
}

// side network: x1 = [sky | action?] -> relu(fc_a) (128) -> relu(fc_b) (16) -> feat[b][5408:5424]; keeps x1, h1, o16
// act_first: x1 = [action | sky] (demix_td3.py:279 concatenates (action, metadata); demix_sac.py / calib_sac.py the other way round)
__global__ void __launch_bounds__(128) k_side_fwd(const float* __restrict__ sky, int SKY, const float* __restrict__ act, int A,
                                                  const float* __restrict__ wa, const float* __restrict__ ba, const float* __restrict__ wb,
                                                  const float* __restrict__ bb, float* __restrict__ x1, float* __restrict__ h1,
                                                  float* __restrict__ o16, float* __restrict__ feat, int side, int feat_ld, int flat,
                                                  int act_first = 0) {
  __shared__ float sx[192];
  __shared__ float sh[HID];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int K1 = SKY + (act ? A : 0);
  for (int i = tid; i < K1; i += 128) {
    float v;
    if (act && act_first) v = i < A ? act[(size_t)b * A + i] : sky[(size_t)b * SKY + (i - A)];
    else v = i < SKY ? sky[(size_t)b * SKY + i] : act[(size_t)b * A + (i - SKY)];
    sx[i] = v;
    x1[(size_t)b * K1 + i] = v;
  }
  __syncthreads();
  {
    const float* w = wa + (size_t)tid * K1;
    float s = ba[tid];
    for (int i = 0; i < K1; ++i) s = fmaf(w[i], sx[i], s);
    s = fmaxf(s, 0.f);
    sh[tid] = s;
    h1[(size_t)b * HID + tid] = s;
  }
  __syncthreads();
  if (tid < side) {
    const float* w = wb + (size_t)tid * HID;
    float s = bb[tid];
    for (int i = 0; i < HID; ++i) s = fmaf(w[i], sh[i], s);
    s = fmaxf(s, 0.f);
    o16[(size_t)b * side + tid] = s;
    feat[(size_t)b * feat_ld + flat + tid] = s;
  }
}

// q[b] = head_w . feat[b] + head_b
__global__ void __launch_bounds__(256) k_critic_head(const float* __restrict__ feat, const float* __restrict__ hw, const float* __restrict__ hb,
                                                     float* __restrict__ q, int feat_ld) {
  __shared__ float sh[8];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < feat_ld; i += 256) s = fmaf(feat[(size_t)b * feat_ld + i], hw[i], s);
  s = block_sum_float(s, sh);
  if (threadIdx.x == 0) q[b] = s + hb[0];
}

// policy head (calib_sac.py:207-231): mu, sigma = clamp(raw, 1e-6, 1), u = mu + eps*sigma, a = tanh(u),
// logp = sum_j [ -(u-mu)^2/(2 sigma^2) - log sigma - log sqrt(2 pi) - log(1 - a^2 + 1e-6) ]
__global__ void __launch_bounds__(128) k_actor_head(const float* __restrict__ x21, const float* __restrict__ wmu, const float* __restrict__ bmu,
                                                    const float* __restrict__ wsg, const float* __restrict__ bsg, const float* __restrict__ eps,
                                                    int A, float* __restrict__ act, float* __restrict__ sig, float* __restrict__ raw,
                                                    float* __restrict__ logp, float* __restrict__ mu_out, float* __restrict__ u_out, int sigma_mode) {
  __shared__ float sx[HID];
  __shared__ float slp[AMAXC];
  const int b = blockIdx.x, tid = threadIdx.x;
  sx[tid] = x21[(size_t)b * HID + tid];
  __syncthreads();
  if (tid < A) {
    float m = bmu[tid], r = bsg[tid];
    const float *w1 = wmu + (size_t)tid * HID, *w2 = wsg + (size_t)tid * HID;
    for (int i = 0; i < HID; ++i) {
      m = fmaf(w1[i], sx[i], m);
      r = fmaf(w2[i], sx[i], r);
    }
    // sigma_mode 0: sigma = clamp(raw, 1e-6, 1) (calib_sac.py:201-203); 1: sigma = exp(clamp(raw, -20, 2)) (demix_sac.py fc22logsigma)
    const float s = sigma_mode == 0 ? fminf(fmaxf(r, 1e-6f), 1.f) : expf(fminf(fmaxf(r, -20.f), 2.f));
    const float e = eps[(size_t)b * A + tid];
    const float u = __fadd_rn(m, __fmul_rn(e, s));
    const float a = tanhf(u);
    const float d = u - m;
    float lp = -(d * d) / (2.f * (s * s)) - logf(s) - 0.91893853320467274178f;
    lp -= logf(1.f - a * a + 1e-6f);
    act[(size_t)b * A + tid] = a;
    sig[(size_t)b * A + tid] = s;
    raw[(size_t)b * A + tid] = r;
    if (mu_out != nullptr) {
      mu_out[(size_t)b * A + tid] = m;
      u_out[(size_t)b * A + tid] = u;
    }
    slp[tid] = lp;
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int j = 0; j < A; ++j) s += slp[j];
    logp[b] = s;
  }
}

// y = r + gamma * (done ? 0 : min(qt1, qt2) - alpha * logp')   (calib_sac.py:320-323; no reward scale)
__global__ void k_target(const float* qt1, const float* qt2, const float* logp2, const float* r, const float* done, float gamma, float alpha,
                         int B, float* y) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = __fsub_rn(fminf(qt1[b], qt2[b]), __fmul_rn(alpha, logp2[b]));
  if (done[b] != 0.f) m = 0.f;
  y[b] = __fadd_rn(r[b], __fmul_rn(gamma, m));
}

// critic loss seed: dq = 2 (q - y) / B; dfeat[b][i] = dq[b] * head_w[i]; stats[slot] = mean (q - y)^2  (one block)
__global__ void __launch_bounds__(256) k_critic_seed(const float* __restrict__ q, const float* __restrict__ y, int B, float* __restrict__ dq,
                                                     float* __restrict__ stat) {
  __shared__ float sh[8];
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    const float d = q[b] - y[b];
    dq[b] = 2.f * d / (float)B;
    s = fmaf(d, d, s);
  }
  s = block_sum_float(s, sh);
  if (threadIdx.x == 0) stat[0] = s / (float)B;
}
__global__ void k_outer(const float* __restrict__ dq, const float* __restrict__ hw, int B, int N, float* __restrict__ out) {
  const size_t total = (size_t)B * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / N;
    out[i] = dq[b] * hw[i - b * N];
  }
}

// side network backward for one row: d16 = dfeat[b][5408:] * (o16 > 0); dh1 = (W_b^T d16) * (h1 > 0);
// optionally dact[b][j] (+)= sum_n dh1[n] * W_a[n][SKY + j]
__global__ void __launch_bounds__(128) k_side_bwd(const float* __restrict__ dfeat_side, long ld_dfeat, const float* __restrict__ o16,
                                                  const float* __restrict__ h1, const float* __restrict__ wb, const float* __restrict__ wa, int K1,
                                                  int SKY, int A, float* __restrict__ d16, float* __restrict__ dh1, float* __restrict__ dact,
                                                  int accumulate, int side, int act_first = 0) {
  __shared__ float sd[SIDEMAX];
  __shared__ float sdh[HID];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < side) {
    const float v = o16[(size_t)b * side + tid] > 0.f ? dfeat_side[(size_t)b * ld_dfeat + tid] : 0.f;
    sd[tid] = v;
    d16[(size_t)b * side + tid] = v;
  }
  __syncthreads();
  {
    float s = 0.f;
    for (int k = 0; k < side; ++k) s = fmaf(sd[k], wb[(size_t)k * HID + tid], s);
    s = h1[(size_t)b * HID + tid] > 0.f ? s : 0.f;
    sdh[tid] = s;
    dh1[(size_t)b * HID + tid] = s;
  }
  __syncthreads();
  if (dact != nullptr && tid < A) {
    float s = 0.f;
    const int col = (act_first ? 0 : SKY) + tid;
    for (int n = 0; n < HID; ++n) s = fmaf(sdh[n], wa[(size_t)n * K1 + col], s);
    dact[(size_t)b * A + tid] = (accumulate ? dact[(size_t)b * A + tid] : 0.f) + s;
  }
}

// actor loss seeds (calib_sac.py:344,355): L = mean(alpha*logp - min(q1,q2)); dq_k = -(1/B) [k attains the min] (halves on ties);
// dfeat_side_k[b][j] = dq_k[b] * head_w_k[5408 + j]; qmin for the loss readout
__global__ void k_actor_seed(const float* q1, const float* q2, const float* hw1, const float* hw2, int B, float* ds1, float* ds2, float* qmin, int side,
                             int flat) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float a = q1[b], c = q2[b];
  const float w = 1.f / (float)B;
  float g1 = 0.f, g2 = 0.f;
  if (a < c) g1 = -w; else if (c < a) g2 = -w; else { g1 = -0.5f * w; g2 = -0.5f * w; }
  qmin[b] = fminf(a, c);
  for (int j = 0; j < side; ++j) {
    ds1[(size_t)b * side + j] = g1 * hw1[flat + j];
    ds2[(size_t)b * side + j] = g2 * hw2[flat + j];
  }
}

// hint penalty (calib_sac.py:345-365): kld = h_m (log h_m - log a_m), a_m = clamp(0.5 a + 0.5 + EPS, EPS, 1) (same for h);
// g = max(0, mean(kld) - thr)^2.  out[0] = d(0.5 admm_rho g^2 + rho g)/d mean(kld) / (B A), out[1] = g, out[2] = penalty value;
// rho_update (every 10th learn): rho += admm_rho * g.   One block.
// kld_mode 1 (demix_sac.py kld_loss): no clamp of the mapped values
__global__ void __launch_bounds__(256) k_kld_coef(const float* __restrict__ act, const float* __restrict__ hint, int n, float thr, float admm_rho,
                                                  float* rho, float* out, int do_rho_update, int kld_mode) {
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    float am = 0.5f * act[i] + 0.5f + 1e-6f, hm = 0.5f * hint[i] + 0.5f + 1e-6f;
    if (kld_mode == 0) { am = fminf(fmaxf(am, 1e-6f), 1.f); hm = fminf(fmaxf(hm, 1e-6f), 1.f); }
    s += hm * (logf(hm) - logf(am));
  }
  s = block_sum_float(s, sh);
  if (threadIdx.x == 0) {
    const float mk = s / (float)n - thr;
    const float pos = fmaxf(mk, 0.f);
    const float g = pos * pos;
    const float r = rho[0];
    out[0] = (admm_rho * g + r) * 2.f * pos / (float)n;
    out[1] = g;
    out[2] = 0.5f * admm_rho * g * g + r * g;
    if (do_rho_update) rho[0] = r + admm_rho * g;
  }
}

// policy head backward for one row.  With ga = dL/da (critics + hint + the tanh correction of logp) and gt = ga (1 - a^2):
//   dL/dsigma = gt * eps - glp / sigma analytically (see below for the float32 order),   d raw = d sigma * [1e-6 <= raw <= 1]
//   dL/dmu    = gt analytically -- but the reference reaches it as (gv + gt) - gv, where gv = -glp (u - mu) / sigma^2 is
//   the derivative of the Gaussian term w.r.t. the sample, which autograd adds to the sample's gradient and subtracts
//   again through `loc` (Normal.log_prob + rsample, calib_sac.py:213-228).  For actions whose sigma sits at the 1e-6
//   clamp |gv| ~ 4e3 |eps| while |gt| ~ 1e-3, so the reference's float32 result is gt plus rounding noise of that order.
//   The same float32 operations are issued here in the same order so that the update matches the reference's.
//   dx21 = Wmu^T dmu + Wsg^T draw, then through the ELU of fc21.
__global__ void __launch_bounds__(128) k_actor_head_bwd(const float* __restrict__ da, const float* __restrict__ act, const float* __restrict__ sig,
                                                        const float* __restrict__ raw, const float* __restrict__ eps, const float* __restrict__ mu,
                                                        const float* __restrict__ uu, const float* __restrict__ hint,
                                                        const float* __restrict__ hint_coef, float glp, int A, const float* __restrict__ wmu,
                                                        const float* __restrict__ wsg, const float* __restrict__ x21, float* __restrict__ dmu,
                                                        float* __restrict__ draw, float* __restrict__ dpre, int sigma_mode, int kld_mode) {
  __shared__ float sm[AMAXC], sr[AMAXC];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < A) {
    const size_t o = (size_t)b * A + tid;
    const float a = act[o], s = sig[o], r = raw[o], e = eps[o];
    float g = da[o];
    if (hint_coef != nullptr) {
      const float arg = 0.5f * a + 0.5f + 1e-6f;
      if (kld_mode == 0) {
        const float am = fminf(fmaxf(arg, 1e-6f), 1.f);
        const float hm = fminf(fmaxf(0.5f * hint[o] + 0.5f + 1e-6f, 1e-6f), 1.f);
        if (arg >= 1e-6f && arg <= 1.f) g += hint_coef[0] * (-hm / am) * 0.5f;
      } else {
        g += hint_coef[0] * (-(0.5f * hint[o] + 0.5f + 1e-6f) / arg) * 0.5f;
      }
    }
    const float omt = __fsub_rn(1.f, __fmul_rn(a, a));
    g += glp * (2.f * a / (omt + 1e-6f));                      // d(-glp log(1 - a^2 + 1e-6))/da
    const float gt = __fmul_rn(g, omt);                        // tanh backward
    const float z = __fsub_rn(uu[o], mu[o]);
    const float d2 = __fmul_rn(2.f, __fmul_rn(s, s));
    const float gv = -__fmul_rn(__fdiv_rn(glp, d2), __fmul_rn(2.f, z));
    const float gu = __fadd_rn(gv, gt);
    const float du = __fsub_rn(gu, gv);                        // = gt up to the reference's cancellation noise
    // sigma collects three contributions in autograd's arrival order (log, var = sigma^2, then the sample mu + eps*sigma):
    //   g3 = -glp / sigma,  g2 = d(-(z^2)/(2 var))/d var * 2 sigma  (ATen: grad_other = -grad * ((self/other)/other)),  g1 = gu * eps.
    // Analytically g1 + g2 = gt * eps; in float32 z = fl(u - mu) differs from eps*sigma by ~ulp(mu), which for sigma near the
    // 1e-6 clamp is a 1e-4 relative effect on the dominant -glp/sigma term, so the reference's order is kept.
    const float nq = __fdiv_rn(-__fmul_rn(z, z), d2);
    const float gd2 = __fmul_rn(-glp, __fdiv_rn(nq, d2));
    const float g2 = __fmul_rn(__fmul_rn(gd2, 2.f), __fmul_rn(2.f, s));
    const float g3 = __fdiv_rn(-glp, s);
    const float g1 = __fmul_rn(gu, e);
    const float ds = __fadd_rn(__fadd_rn(g3, g2), g1);
    // sigma_mode 1: sigma = exp(clamp(raw, -20, 2)): exp backward multiplies by its output, the clamp masks
    const float dr = sigma_mode == 0 ? ((r >= 1e-6f && r <= 1.f) ? ds : 0.f) : ((r >= -20.f && r <= 2.f) ? __fmul_rn(ds, s) : 0.f);
    sm[tid] = du;
    sr[tid] = dr;
    dmu[o] = du;
    draw[o] = dr;
  }
  __syncthreads();
  float s = 0.f;
  for (int j = 0; j < A; ++j) s = fmaf(sm[j], wmu[(size_t)j * HID + tid], fmaf(sr[j], wsg[(size_t)j * HID + tid], s));
  dpre[(size_t)b * HID + tid] = s * elu_grad_from_out(x21[(size_t)b * HID + tid]);
}

// actor loss readout: mean(alpha*logp - qmin) (+ penalty)   (one block)
__global__ void __launch_bounds__(256) k_actor_loss(const float* __restrict__ logp, const float* __restrict__ qmin, int B, float alpha,
                                                    const float* __restrict__ pen, float* __restrict__ stat) {
  __shared__ float sh[8];
  float s = 0.f, l = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    s += alpha * logp[b] - qmin[b];
    l += logp[b];
  }
  s = block_sum_float(s, sh);
  l = block_sum_float(l, sh);
  if (threadIdx.x == 0) {
    stat[0] = s / (float)B + (pen ? pen[2] : 0.f);
    stat[1] = l / (float)B;
  }
}

// ---- demixing conv TD3 (demixing_rl/demix_td3.py) ---------------------------------------------------------------------
// deterministic policy head (:342): a = tanh(fc22 x21).  Target pass (:487-491): a = clamp(a + n, -1, 1) with ONE scalar n per
// step; `noise` holds it already clamped to [-0.5, 0.5] (host draw), or a raw N(0,1) Philox value when `noise_raw` (device RNG:
// scaled by 0.2 and clamped here).  choose_action (:448-453): per-action additive noise `anoise` (A values), clamp to [-1, 1].
__global__ void __launch_bounds__(128) k_td3_head(const float* __restrict__ x21, const float* __restrict__ w, const float* __restrict__ bias, int A,
                                                  const float* __restrict__ noise, int noise_raw, const float* __restrict__ anoise,
                                                  float* __restrict__ act) {
  __shared__ float sx[HID];
  const int b = blockIdx.x, tid = threadIdx.x;
  sx[tid] = x21[(size_t)b * HID + tid];
  __syncthreads();
  if (tid < A) {
    float m = bias[tid];
    const float* w1 = w + (size_t)tid * HID;
    for (int i = 0; i < HID; ++i) m = fmaf(w1[i], sx[i], m);
    float a = tanhf(m);
    if (noise != nullptr) {
      float n = noise[0];
      if (noise_raw) n = fminf(fmaxf(0.2f * n, -0.5f), 0.5f);
      a = fminf(fmaxf(__fadd_rn(a, n), -1.f), 1.f);
    }
    if (anoise != nullptr) a = fminf(fmaxf(__fadd_rn(a, anoise[tid]), -1.f), 1.f);
    act[(size_t)b * A + tid] = a;
  }
}

// y = r + gamma * (done ? 0 : min(q1', q2'))   (:498-504)
__global__ void k_td3_target(const float* qt1, const float* qt2, const float* r, const float* done, float gamma, int B, float* y) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float m = done[b] != 0.f ? 0.f : fminf(qt1[b], qt2[b]);
  y[b] = __fadd_rn(r[b], __fmul_rn(gamma, m));
}

// priority errors of the pre-pass (:506-511): 0.5 (|q1 - y| + |q2 - y|), float32 like the numpy expression
__global__ void k_td3_per_err(const float* q1, const float* q2, const float* y, int B, float* err) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  err[b] = __fmul_rn(0.5f, __fadd_rn(fabsf(__fsub_rn(q1[b], y[b])), fabsf(__fsub_rn(q2[b], y[b]))));
}

// PER.mse (:137-140) with is_weights (B,) against (B,1) errors: the (B,B) broadcast makes the loss mean(w) * mean((y - q)^2), so
// dL/dq = mean(w) * 2 (q - y) / B.  One block.
__global__ void __launch_bounds__(256) k_td3_critic_seed(const float* __restrict__ q, const float* __restrict__ y, const float* __restrict__ wsum,
                                                         int B, float* __restrict__ dq, float* __restrict__ stat) {
  __shared__ float sh[8];
  const float wm = wsum[0] / (float)B;
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    const float d = q[b] - y[b];
    dq[b] = wm * (2.f * d / (float)B);
    s = fmaf(d, d, s);
  }
  s = block_sum_float(s, sh);
  if (threadIdx.x == 0) stat[0] = wm * (s / (float)B);
}

// actor loss -mean(q1(s, pi(s)) * w) (:538 / :553) with the same broadcast = -mean(q) mean(w): dq = -mean(w) / B for every row;
// ds[b][j] = dq * head_w1[flat + j] seeds the backward pass through critic_1's side network.  stat[0] = that loss term.  One block.
__global__ void __launch_bounds__(256) k_td3_actor_seed(const float* __restrict__ q, const float* __restrict__ hw, const float* __restrict__ wsum,
                                                        int B, float* __restrict__ ds, int side, int flat, float* __restrict__ stat) {
  __shared__ float sh[8];
  const float wm = wsum[0] / (float)B;
  const float g = -wm / (float)B;
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) s += q[b];
  for (int i = threadIdx.x; i < B * side; i += 256) ds[i] = g * hw[flat + (i % side)];
  s = block_sum_float(s, sh);
  if (threadIdx.x == 0) stat[0] = -(s / (float)B) * wm;
}

// ADMM hint penalty of one inner actor step (:555-561): pen = mean(w) (lambda . d + rho / 2 mse(a, hint)) / n with d = a - hint
// flattened, n = B A.  dact += mean(w) / n (lambda + rho d / n); stat[0] += pen, stat[1] = pen.  One block.
__global__ void __launch_bounds__(256) k_td3_hint(const float* __restrict__ act, const float* __restrict__ hint, const float* __restrict__ lambda,
                                                  const float* __restrict__ rho, const float* __restrict__ wsum, int B, int n,
                                                  float* __restrict__ dact, float* __restrict__ stat) {
  __shared__ float sh[8];
  const float wm = wsum[0] / (float)B, r = rho[0], fn = (float)n;
  float dot = 0.f, sq = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = act[i] - hint[i];
    dot = fmaf(lambda[i], d, dot);
    sq = fmaf(d, d, sq);
    dact[i] += (wm / fn) * (lambda[i] + r * d / fn);
  }
  dot = block_sum_float(dot, sh);
  sq = block_sum_float(sq, sh);
  if (threadIdx.x == 0) {
    const float pen = (dot + 0.5f * r * (sq / fn)) * wm / fn;
    stat[0] += pen;
    stat[1] = pen;
  }
}

// backward of a = tanh(fc22 x21) for one row: dfc = dact (1 - a^2); dpre = (W22^T dfc) * ELU'(x21)
__global__ void __launch_bounds__(128) k_td3_head_bwd(const float* __restrict__ da, const float* __restrict__ act, int A, const float* __restrict__ w,
                                                      const float* __restrict__ x21, float* __restrict__ dfc, float* __restrict__ dpre) {
  __shared__ float sm[AMAXC];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < A) {
    const size_t o = (size_t)b * A + tid;
    const float a = act[o];
    const float g = __fmul_rn(da[o], __fsub_rn(1.f, __fmul_rn(a, a)));
    sm[tid] = g;
    dfc[o] = g;
  }
  __syncthreads();
  float s = 0.f;
  for (int j = 0; j < A; ++j) s = fmaf(sm[j], w[(size_t)j * HID + tid], s);
  dpre[(size_t)b * HID + tid] = s * elu_grad_from_out(x21[(size_t)b * HID + tid]);
}

// PER.store_transition with error=None (:60-66) + SumTree.add / update (demix_sac.py:66-110): the new leaf gets the largest leaf
// priority (absolute_error_upper while the tree is empty), the change is added to every ancestor in order.  One block.
__global__ void __launch_bounds__(256) k_tree_add(double* __restrict__ tree, int C, int slot, double clip) {
  __shared__ double sh[256];
  double m = 0.0;
  for (int i = threadIdx.x; i < C; i += 256) m = fmax(m, tree[C - 1 + i]);
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double p = sh[0];
    if (p == 0.0) p = clip;
    int node = C - 1 + slot;
    const double change = __dsub_rn(p, tree[node]);
    tree[node] = p;
    while (node != 0) {
      node = (node - 1) >> 1;
      tree[node] = __dadd_rn(tree[node], change);
    }
  }
}

// update_network_parameters over the BatchNorm num_batches_tracked entries (:413-418): tau * online + (1 - tau) * target in float32
// (int64 * python float promotes to float32), truncated back to int64 by load_state_dict's copy_
__global__ void k_nbt_blend(const long long* __restrict__ src, long long* __restrict__ dst, float tau, float omt) {
  const int i = threadIdx.x;
  if (i < 3) dst[i] = (long long)__fadd_rn(__fmul_rn(tau, (float)src[i]), __fmul_rn(omt, (float)dst[i]));
}

}  // namespace

// =====================================================================================================================
// host side
// =====================================================================================================================
struct CTensor { std::string name; int rows, cols; size_t off; };
struct CLayout {
  std::vector<CTensor> t;
  size_t size = 0;
  void add(const char* n, int r, int c) { t.push_back(CTensor{n, r, c, size}); size += ((size_t)r * c + 3) & ~(size_t)3; }
  size_t off(const char* n) const {
    for (auto& d : t) if (d.name == n) return d.off;
    return (size_t)-1;
  }
};
struct NetP {
  float *cw[3], *cb[3], *g[3], *be[3];
  float *fa_w, *fa_b, *fb_w, *fb_b, *h_w, *h_b, *mu_w, *mu_b, *sg_w, *sg_b;
};
struct TrunkWs {
  float *z[3], *a[3];                 // a[2] aliases feat (image stride FEAT)
  float* feat;                        // [B][5424]
  float *mean[3], *rstd[3], *scale[3], *shift[3];
  float *x1, *h1, *o16;               // side network activations
};
struct PiSide { float *x1, *h1, *o16; };

struct sc_calib {
  sc_calib_config cfg;
  int B, C, A, M, SKY, K1C;           // K1C: critic side input width (sky + action)
  // network family (sc_calib_config.variant): calibration (calib_sac.py), demixing_rl (demix_sac.py) or demixing_fuzzy MLP
  bool td3;                           // SC_CALIB_TD3: demixing_rl/demix_td3.py program (tanh policy, target actor, PER, state_dict soft update)
  int act_first;                      // critic side input is cat((action, metadata)) (demix_td3.py:279) instead of (metadata, action)
  int interval;                       // update_actor_interval
  bool trunk;                         // conv trunk present (use_influence); false: pure MLP, features = side network only
  int side, flat, feat;               // side-network width (INNER_DIM), flattened trunk width (0 without trunk), feat = flat + side
  int sigma_mode, kld_mode;           // 0 calibration, 1 demixing (log-sigma heads; unclamped KLD + resampled dual update)
  float *eps_dual;
  cudaStream_t st;
  bool own_arena;
  char* arena;
  size_t arena_bytes;
  CLayout La, Lc;
  // parameter-like blocks: [actor | critic_1 | critic_2]
  float *params, *grads, *adam_m, *adam_v, *targets;     // targets: [t1 | t2] in critic layout (+ [target actor] in actor layout, TD3)
  size_t off_net[3], n_params;
  float* bnbuf[6];                    // per net: rm1 rv1 rm2 rv2 rm3 rv3 (16,16,32,32,32,32 floats)
  // conv2 / conv3 on tcgen05 (sc_conv_tc.cuh): tf32 big / small copies of the two weight tensors of every network
  // ([2][32 x 400] then [2][32 x 800] floats) and their tensor maps [net][layer - 1][big, small]
  float* conv_hl[6];
  CUtensorMap conv_map[6][2][2];
  bool conv_tc;
  int conv_dbg;
  unsigned long long* conv_ts;        // debug timeline of k_conv_fwd_tc (SMARTCAL_STAMPS), "debug_ts"
  long long* nbt[6];                  // per net: 3 counters
  // TD3: prioritized replay (float64 sum tree exactly like the reference's flat heap), ADMM state of the actor loop, target noise
  double* tree;
  int* tree_idx;
  float *is_w, *is_wsum, *per_err, *per_ps;
  double* u01_dev;
  double per_beta;
  float *lambda, *admm_y0, *admm_a0, *admm_rho_dev, *tnoise, *anoise;
  AdmmP* admm_dev;                    // [6]: loop start, then the bookkeeping after inner step 0..4
  AdmmP admm_host[6];
  float* dq2;                         // second critic's loss seed
  // replay
  float *r_img, *r_img2, *r_sky, *r_sky2, *r_act, *r_hint, *r_rew, *r_done;
  int* bidx;
  // minibatch
  float *mb_img, *mb_img2, *mb_sky, *mb_sky2, *mb_act, *mb_hint, *mb_rew, *mb_done;
  float *eps_next, *eps_cur;
  TrunkWs ws[4];
  PiSide pside[2];
  // heads / scalars
  float *x21, *pre21;                 // actor fc21 output (post-ELU) [B][128]; scratch
  float *a2, *sig2, *raw2, *logp2;    // policy on s'
  float *pi, *sig, *raw, *logp, *mu, *u;   // policy on s
  float *qt[2], *q[2], *qpi[2], *y, *dq, *qmin;
  float *dfeat, *dz3, *dA2, *dA1;     // gradient workspace (dz2, dz1 written in place over dA2, dA1)
  float *d16, *dh1, *ds[2], *dact, *dmu, *draw, *dpre;
  float *bn_c1, *bn_c2;               // BN backward per-channel means
  double* red_partial;                // [32][NCH][2]
  float* wg_partial;
  float* splitk;                      // [32][B][128] split-K partials of the fc21 forward GEMM
  float *stats, *rho, *hint_out;
  float* h_stage;                     // pinned host staging
  float* h_stats;
  int* h_idx;
  double* h_u01;                      // pinned: B PER uniforms + the target-noise scalar of one TD3 learn()
  int64_t mem_cntr, learn_counter, t_actor, t_critic, launches, noise_ctr;
  bool batch_ready;
  float admm_rho;
  int wg_blocks, nch, sm_count;
  // per-step scalars in device memory (Adam bias corrections of the three optimizers, Philox stream ids of the noise draws): one
  // small copy per step; with them the launch sequence of learn() is identical from step to step and replays from a CUDA graph
  struct StepScalars {
    AdamP adam[3]; unsigned long long noise[4];
    AdamP adam_x[5];                 // TD3 + hints: the actor takes five Adam steps inside one learn() (demix_td3.py:548-563)
    PerSampleP per_s; PerUpdateP per_u;
  };
  StepScalars* step_dev;
  StepScalars* step_host;          // pinned, two slots
  cudaEvent_t ev_step[2];
  int step_k;
  cudaStream_t st_cap;             // library-owned capture stream (the caller's may be the legacy default stream)
  cudaGraphExec_t graph[2];        // [dual update step?]
  int graph_seen[2];
  int use_graph;
  int64_t graph_launches[2];
  std::map<std::string, sc_tensor_info> reg;
};

static void add_trunk(CLayout& L) {
  L.add("conv1.weight", C1, 1 * TAPS); L.add("conv1.bias", 1, C1); L.add("bn1.weight", 1, C1); L.add("bn1.bias", 1, C1);
  L.add("conv2.weight", C2, C1 * TAPS); L.add("conv2.bias", 1, C2); L.add("bn2.weight", 1, C2); L.add("bn2.bias", 1, C2);
  L.add("conv3.weight", C3, C2 * TAPS); L.add("conv3.bias", 1, C3); L.add("bn3.weight", 1, C3); L.add("bn3.bias", 1, C3);
}
static CLayout critic_layout_c(int K1, bool trunk, int side, int feat) {
  CLayout L;
  if (trunk) add_trunk(L);
  L.add("fc1.weight", HID, K1); L.add("fc1.bias", 1, HID);
  L.add("fc2.weight", side, HID); L.add("fc2.bias", 1, side);
  L.add("head.weight", 1, feat); L.add("head.bias", 1, 1);
  return L;
}
static CLayout actor_layout_c(int SKY, int A, bool trunk, int side, int feat, int sigma_mode, bool td3) {
  CLayout L;
  if (trunk) add_trunk(L);
  L.add("fc11.weight", HID, SKY); L.add("fc11.bias", 1, HID);
  L.add("fc12.weight", side, HID); L.add("fc12.bias", 1, side);
  L.add("fc21.weight", HID, feat); L.add("fc21.bias", 1, HID);
  if (td3) { L.add("fc22.weight", A, HID); L.add("fc22.bias", 1, A); return L; }      // demix_td3.py:318: one tanh head
  L.add("fc22mu.weight", A, HID); L.add("fc22mu.bias", 1, A);
  // calib_sac.py:185 names the second head fc22sigma, demix_sac.py:467 fc22logsigma
  if (sigma_mode == 0) { L.add("fc22sigma.weight", A, HID); L.add("fc22sigma.bias", 1, A); }
  else { L.add("fc22logsigma.weight", A, HID); L.add("fc22logsigma.bias", 1, A); }
  return L;
}
static NetP net_ptrs(float* b, const CLayout& L, bool actor) {
  NetP p;
  memset(&p, 0, sizeof(p));
  const char* cw[3] = {"conv1.weight", "conv2.weight", "conv3.weight"};
  const char* cb[3] = {"conv1.bias", "conv2.bias", "conv3.bias"};
  const char* g[3] = {"bn1.weight", "bn2.weight", "bn3.weight"};
  const char* be[3] = {"bn1.bias", "bn2.bias", "bn3.bias"};
  if (L.off(cw[0]) != (size_t)-1)
    for (int l = 0; l < 3; ++l) { p.cw[l] = b + L.off(cw[l]); p.cb[l] = b + L.off(cb[l]); p.g[l] = b + L.off(g[l]); p.be[l] = b + L.off(be[l]); }
  if (actor) {
    p.fa_w = b + L.off("fc11.weight"); p.fa_b = b + L.off("fc11.bias");
    p.fb_w = b + L.off("fc12.weight"); p.fb_b = b + L.off("fc12.bias");
    p.h_w = b + L.off("fc21.weight"); p.h_b = b + L.off("fc21.bias");
    if (L.off("fc22.weight") != (size_t)-1) {         // TD3: single head, kept in the mu slots
      p.mu_w = b + L.off("fc22.weight"); p.mu_b = b + L.off("fc22.bias");
      return p;
    }
    p.mu_w = b + L.off("fc22mu.weight"); p.mu_b = b + L.off("fc22mu.bias");
    const bool logsig = L.off("fc22sigma.weight") == (size_t)-1;
    p.sg_w = b + L.off(logsig ? "fc22logsigma.weight" : "fc22sigma.weight"); p.sg_b = b + L.off(logsig ? "fc22logsigma.bias" : "fc22sigma.bias");
  } else {
    p.fa_w = b + L.off("fc1.weight"); p.fa_b = b + L.off("fc1.bias");
    p.fb_w = b + L.off("fc2.weight"); p.fb_b = b + L.off("fc2.bias");
    p.h_w = b + L.off("head.weight"); p.h_b = b + L.off("head.bias");
  }
  return p;
}

struct CarverC {
  char* base; size_t off = 0;
  void* take(size_t bytes) { off = (off + 255) & ~(size_t)255; void* p = base ? base + off : nullptr; off += bytes; return p; }
  float* f(size_t n) { return (float*)take(n * sizeof(float)); }
};

static size_t carve_c(sc_calib* a, char* base) {
  CarverC cv{base};
  const size_t B = a->B, C = a->C, A = a->A, SKY = a->SKY;
  auto reg = [&](const char* name, void* p, int64_t rows, int64_t cols, int elem = 4, int is_int = 0) {
    if (base) a->reg[name] = sc_tensor_info{p, rows, cols, cols, elem, is_int};
  };
  a->off_net[0] = 0; a->off_net[1] = a->La.size; a->off_net[2] = a->La.size + a->Lc.size;
  a->n_params = a->La.size + 2 * a->Lc.size;
  a->params = cv.f(a->n_params); a->grads = cv.f(a->n_params); a->adam_m = cv.f(a->n_params); a->adam_v = cv.f(a->n_params);
  a->targets = cv.f(2 * a->Lc.size + (a->td3 ? a->La.size : 0));
  for (int n = 0; n < 6; ++n) { a->bnbuf[n] = cv.f(160); a->nbt[n] = (long long*)cv.take(4 * sizeof(long long)); }
  a->conv_ts = (unsigned long long*)cv.take(sizeof(unsigned long long) * 1024);
  reg("debug_ts", a->conv_ts, 1024, 1, 8);
  for (int n = 0; n < 6; ++n) a->conv_hl[n] = cv.f(a->trunk ? (size_t)2 * C2 * C1 * TAPS + (size_t)2 * C3 * C2 * TAPS : 4);
  const size_t NIMG = a->trunk ? (size_t)IMG * IMG : 0;       // no influence maps without the conv trunk
  const size_t FEATR = a->feat, SIDER = a->side;
  a->r_img = cv.f(C * NIMG); a->r_img2 = cv.f(C * NIMG);
  a->r_sky = cv.f(C * SKY); a->r_sky2 = cv.f(C * SKY); a->r_act = cv.f(C * A); a->r_hint = cv.f(C * A); a->r_rew = cv.f(C); a->r_done = cv.f(C);
  if (a->trunk) { reg("state_memory_img", a->r_img, C, IMG * IMG); reg("new_state_memory_img", a->r_img2, C, IMG * IMG); }
  reg("state_memory_sky", a->r_sky, C, SKY); reg("new_state_memory_sky", a->r_sky2, C, SKY);
  reg("action_memory", a->r_act, C, A); reg("hint_memory", a->r_hint, C, A); reg("reward_memory", a->r_rew, C, 1); reg("terminal_memory", a->r_done, C, 1);
  a->bidx = (int*)cv.take(B * sizeof(int)); reg("batch_idx", a->bidx, B, 1, 4, 1);
  a->mb_img = cv.f(B * NIMG); a->mb_img2 = cv.f(B * NIMG); a->mb_sky = cv.f(B * SKY); a->mb_sky2 = cv.f(B * SKY);
  a->mb_act = cv.f(B * A); a->mb_hint = cv.f(B * A); a->mb_rew = cv.f(B); a->mb_done = cv.f(B);
  a->eps_next = cv.f(B * A); a->eps_cur = cv.f(B * A); reg("eps_next", a->eps_next, B, A); reg("eps_cur", a->eps_cur, B, A);
  a->eps_dual = cv.f(B * A);
  for (int w = 0; w < 4; ++w) {
    TrunkWs& t = a->ws[w];
    const size_t TB = a->trunk ? B : 0;
    t.z[0] = cv.f(TB * C1 * H1C * H1C); t.a[0] = cv.f(TB * C1 * H1C * H1C);
    t.z[1] = cv.f(TB * C2 * H2C * H2C); t.a[1] = cv.f(TB * C2 * H2C * H2C);
    t.z[2] = cv.f(TB * C3 * H3C * H3C);
    t.feat = cv.f(B * FEATR); t.a[2] = t.feat;
    for (int l = 0; l < 3; ++l) { t.mean[l] = cv.f(32); t.rstd[l] = cv.f(32); t.scale[l] = cv.f(32); t.shift[l] = cv.f(32); }
    t.x1 = cv.f(B * (SKY + A)); t.h1 = cv.f(B * HID); t.o16 = cv.f(B * SIDER);
  }
  for (int k = 0; k < 2; ++k) { a->pside[k].x1 = cv.f(B * (SKY + A)); a->pside[k].h1 = cv.f(B * HID); a->pside[k].o16 = cv.f(B * SIDER); }
  a->x21 = cv.f(B * HID); a->pre21 = cv.f(B * HID);
  a->a2 = cv.f(B * A); a->sig2 = cv.f(B * A); a->raw2 = cv.f(B * A); a->logp2 = cv.f(B);
  a->pi = cv.f(B * A); a->sig = cv.f(B * A); a->raw = cv.f(B * A); a->logp = cv.f(B); a->mu = cv.f(B * A); a->u = cv.f(B * A);
  reg("pi", a->pi, B, A); reg("logp", a->logp, B, 1); reg("sigma", a->sig, B, A);
  for (int k = 0; k < 2; ++k) { a->qt[k] = cv.f(B); a->q[k] = cv.f(B); a->qpi[k] = cv.f(B); a->ds[k] = cv.f(B * SIDER); }
  a->y = cv.f(B); a->dq = cv.f(B); a->qmin = cv.f(B);
  reg("y", a->y, B, 1); reg("q1", a->q[0], B, 1); reg("q2", a->q[1], B, 1); reg("q1_pi", a->qpi[0], B, 1); reg("q2_pi", a->qpi[1], B, 1);
  const size_t TB2 = a->trunk ? B : 0;
  a->dfeat = cv.f(B * FEATR); a->dz3 = cv.f(TB2 * C3 * H3C * H3C); a->dA2 = cv.f(TB2 * C2 * H2C * H2C); a->dA1 = cv.f(TB2 * C1 * H1C * H1C);
  a->d16 = cv.f(B * SIDER); a->dh1 = cv.f(B * HID); a->dact = cv.f(B * A); reg("d_pi", a->dact, B, A); a->dmu = cv.f(B * A); a->draw = cv.f(B * A); reg("d_mu", a->dmu, B, A); reg("d_sigma_raw", a->draw, B, A); a->dpre = cv.f(B * HID);
  a->bn_c1 = cv.f(32); a->bn_c2 = cv.f(32);
  a->red_partial = (double*)cv.take(sizeof(double) * 32 * 64 * 2);
  a->wg_partial = cv.f(a->trunk ? (size_t)a->wg_blocks * C3 * C2 * TAPS : 4);
  a->splitk = cv.f((size_t)32 * B * HID);
  a->stats = cv.f(16); a->rho = cv.f(4); a->hint_out = cv.f(8);
  if (a->td3) {
    a->tree = (double*)cv.take(sizeof(double) * (2 * C - 1)); reg("tree", a->tree, 1, 2 * (int64_t)C - 1, 8);
    a->tree_idx = (int*)cv.take(B * sizeof(int)); reg("tree_idx", a->tree_idx, B, 1, 4, 1);
    a->is_w = cv.f(B); reg("is_weights", a->is_w, B, 1);
    a->is_wsum = cv.f(4); a->per_err = cv.f(B); a->per_ps = cv.f(B); reg("per_errors", a->per_err, B, 1); reg("per_priorities", a->per_ps, B, 1);
    a->u01_dev = (double*)cv.take(sizeof(double) * B);
    a->lambda = cv.f(B * A); a->admm_y0 = cv.f(B * A); a->admm_a0 = cv.f(B * A); a->admm_rho_dev = cv.f(4); reg("admm_rho", a->admm_rho_dev, 1, 1);
    reg("admm_lambda", a->lambda, B, A);
    a->tnoise = cv.f(4); a->anoise = cv.f(A);
    a->admm_dev = (AdmmP*)cv.take(6 * sizeof(AdmmP));
    a->dq2 = cv.f(B);
    reg("target_action", a->a2, B, A);
  }
  return cv.off;
}

static int check_cfg_c(const sc_calib_config* c) {
  if (!c) return failc(SC_ERR_ARG, "null config");
  if (c->abi_version != SC_ABI_VERSION) return failc(SC_ERR_ARG, "ABI version mismatch");
  const bool trunk = !(c->variant & SC_CALIB_NO_TRUNK);
  if (trunk && c->img != IMG) return failc(SC_ERR_ARG, "only input_dims = [1,128,128] is built (calibration/main_sac.py:17)");
  if (trunk && (c->variant & SC_CALIB_SIDE32)) return failc(SC_ERR_UNSUPPORTED, "INNER_DIM = 32 is built for the MLP (use_influence=False) family only");
  if (c->batch_size < 2 || c->mem_size < c->batch_size) return failc(SC_ERR_ARG, "need 2 <= batch_size <= mem_size");
  if (c->variant & SC_CALIB_TD3) {
    if (!trunk || !(c->variant & SC_CALIB_FLAT_META) || (c->variant & (SC_CALIB_LOGSIGMA | SC_CALIB_SIDE32)))
      return failc(SC_ERR_ARG, "SC_CALIB_TD3 is the demixing_rl/demix_td3.py agent: conv trunk, flat metadata, INNER_DIM 16, one tanh head");
    if (c->mem_size & (c->mem_size - 1)) return failc(SC_ERR_ARG, "TD3 replay is prioritized: max_mem_size must be a power of two (demix_sac.py:47)");
    if (c->update_actor_interval < 1) return failc(SC_ERR_ARG, "update_actor_interval must be >= 1");
    if (c->batch_size > 8192) return failc(SC_ERR_ARG, "batch_size <= 8192 for the prioritized sampler");
  }
  const int sky = (c->variant & SC_CALIB_FLAT_META) ? c->M : 7 * (c->M + 1);
  if (c->n_actions < 1 || c->n_actions > AMAXC || c->M < 1 || sky + c->n_actions > 192)
    return failc(SC_ERR_ARG, "n_actions must be in [1,%d] and metadata width + n_actions <= 192", AMAXC);
  return SC_OK;
}
static void init_dims_c(sc_calib* a, const sc_calib_config* c) {
  a->cfg = *c;
  a->B = c->batch_size; a->C = c->mem_size; a->A = c->n_actions; a->M = c->M;
  a->trunk = !(c->variant & SC_CALIB_NO_TRUNK);
  a->td3 = (c->variant & SC_CALIB_TD3) != 0;
  a->act_first = a->td3 ? 1 : 0;
  a->interval = a->td3 ? c->update_actor_interval : 1;
  a->side = (c->variant & SC_CALIB_SIDE32) ? 32 : SIDE;
  a->flat = a->trunk ? FLAT : 0;
  a->feat = a->flat + a->side;
  a->sigma_mode = (c->variant & SC_CALIB_LOGSIGMA) ? 1 : 0;
  a->kld_mode = (c->variant & SC_CALIB_DEMIX_HINT) ? 1 : 0;
  a->SKY = (c->variant & SC_CALIB_FLAT_META) ? c->M : 7 * (c->M + 1);
  a->K1C = a->SKY + a->A;
  a->La = actor_layout_c(a->SKY, a->A, a->trunk, a->side, a->feat, a->sigma_mode, a->td3);
  a->Lc = critic_layout_c(a->K1C, a->trunk, a->side, a->feat);
  a->wg_blocks = 296;
  a->sm_count = 148;
  a->nch = a->B < 32 ? a->B : 32;
}

// ---- launch helpers -------------------------------------------------------------------------------------------------
#define LAUNCH_OK(a) do { CKC(cudaGetLastError()); (a)->launches++; } while (0)
static inline int grid_for(size_t n, int block = 256, int cap = 148 * 16) {
  size_t g = (n + block - 1) / block;
  return (int)(g > (size_t)cap ? cap : (g ? g : 1));
}

// rep > 1 (training mode): the running statistics take `rep` momentum updates with the same batch statistics -- what `rep` forward
// passes over the same input with unchanged weights leave behind (demix_td3.py: the priority pre-pass :506-511 before the training
// pass :516-517 of the critics)
static int bn_forward(sc_calib* a, TrunkWs& t, const NetP& P, int net, int l, int C, int HW, int act, int train, int Bn, int rep = 1) {
  float* rb = a->bnbuf[net];
  float* rm = rb + (l == 0 ? 0 : l == 1 ? 32 : 96);
  float* rv = rm + (l == 0 ? 16 : 32);
  if (train) {
    ChanReduceP r;
    memset(&r, 0, sizeof(r));
    r.z = t.z[l]; r.B = Bn; r.C = C; r.HW = HW; r.mode = 0; r.partial = a->red_partial;
    const int nch = Bn < a->nch ? Bn : a->nch;
    k_chan_reduce<<<dim3(C, nch), 256, 0, a->st>>>(r);
    LAUNCH_OK(a);
    BnFinalizeP f{a->red_partial, nch, P.g[l], P.be[l], rm, rv, a->nbt[net] + l, t.mean[l], t.rstd[l], t.scale[l], t.shift[l], (double)Bn * HW, 1};
    for (int r = 0; r < rep; ++r) {
      k_bn_finalize<<<1, 32, 0, a->st>>>(f, C);
      LAUNCH_OK(a);
    }
  } else {
    BnFinalizeP f{nullptr, 0, P.g[l], P.be[l], rm, rv, a->nbt[net] + l, t.mean[l], t.rstd[l], t.scale[l], t.shift[l], 1.0, 0};
    k_bn_finalize<<<1, 32, 0, a->st>>>(f, C);
    LAUNCH_OK(a);
  }
  const size_t total = (size_t)Bn * C * HW;
  const size_t astride = l == 2 ? (size_t)FEAT : (size_t)C * HW;
  k_bn_act<<<grid_for(total), 256, 0, a->st>>>(t.z[l], t.scale[l], t.shift[l], C, HW, total, act, t.a[l], astride);
  LAUNCH_OK(a);
  return SC_OK;
}

// conv trunk forward of network `net` (parameters P) on images x [Bn][16384] into workspace t
static int trunk_forward(sc_calib* a, TrunkWs& t, const NetP& P, int net, const float* x, int act, int train, int Bn, int rep = 1) {
  using G1 = ConvGeom<1, C1, IMG, H1C>;
  using G2 = ConvGeom<C1, C2, H1C, H2C>;
  using G3 = ConvGeom<C2, C3, H2C, H3C>;
  k_conv_fwd<1, C1, IMG, H1C><<<dim3(G1::BLOCKS_PER_IMG, Bn), 256, 0, a->st>>>(x, (size_t)IMG * IMG, P.cw[0], P.cb[0], t.z[0]);
  LAUNCH_OK(a);
  int rc = bn_forward(a, t, P, net, 0, C1, H1C * H1C, act, train, Bn, rep);
  if (rc) return rc;
  if (a->conv_tc) {
    // conv2 / conv3 as implicit GEMMs on tcgen05: this network's weights as tf32 big / small parts first (they change every step)
    float* hl2 = a->conv_hl[net];
    float* hl3 = hl2 + (size_t)2 * C2 * C1 * TAPS;
    k_conv_w_split<<<38, 256, 0, a->st>>>(P.cw[1], hl2, C2 * C1 * TAPS, P.cw[2], hl3, C3 * C2 * TAPS);
    LAUNCH_OK(a);
    const int t2 = (Bn * H2C * H2C + CT_M - 1) / CT_M, t3 = (Bn * H3C * H3C + CT_M - 1) / CT_M;
    k_conv_fwd_tc<C1, H1C, H2C><<<t2 < a->sm_count ? t2 : a->sm_count, CT_THREADS, conv_tc_smem_bytes(), a->st>>>(
        ConvTcP{t.a[0], (size_t)C1 * H1C * H1C, P.cb[1], t.z[1], Bn, a->conv_ts, a->conv_dbg}, a->conv_map[net][0][0], a->conv_map[net][0][1]);
    LAUNCH_OK(a);
    if ((rc = bn_forward(a, t, P, net, 1, C2, H2C * H2C, act, train, Bn, rep))) return rc;
    k_conv_fwd_tc<C2, H2C, H3C><<<t3 < a->sm_count ? t3 : a->sm_count, CT_THREADS, conv_tc_smem_bytes(), a->st>>>(
        ConvTcP{t.a[1], (size_t)C2 * H2C * H2C, P.cb[2], t.z[2], Bn, nullptr, a->conv_dbg}, a->conv_map[net][1][0], a->conv_map[net][1][1]);
    LAUNCH_OK(a);
    return bn_forward(a, t, P, net, 2, C3, H3C * H3C, act, train, Bn, rep);
  }
  k_conv_fwd<C1, C2, H1C, H2C><<<dim3(G2::BLOCKS_PER_IMG, Bn), 256, 0, a->st>>>(t.a[0], (size_t)C1 * H1C * H1C, P.cw[1], P.cb[1], t.z[1]);
  LAUNCH_OK(a);
  if ((rc = bn_forward(a, t, P, net, 1, C2, H2C * H2C, act, train, Bn, rep))) return rc;
  k_conv_fwd<C2, C3, H2C, H3C><<<dim3(G3::BLOCKS_PER_IMG, Bn), 256, 0, a->st>>>(t.a[1], (size_t)C2 * H2C * H2C, P.cw[2], P.cb[2], t.z[2]);
  LAUNCH_OK(a);
  return bn_forward(a, t, P, net, 2, C3, H3C * H3C, act, train, Bn, rep);
}

static int sgemm(sc_calib* a, const float* A_, long sam, long sak, const float* B_, long sbk, long sbn, float* C_, long ldc, int M, int N, int K,
                 const float* bias, int act) {
  const int tiles = ((N + 63) / 64) * ((M + 63) / 64);
  if (tiles < 32 && K >= 2048 && (size_t)M * N <= (size_t)a->B * HID) {
    // skinny output, long reduction (fc21 forward: 256 x 128 x 5424): split K over the grid, then add the slices in order
    const int nz = 32, kslice = ((K + nz * 16 - 1) / (nz * 16)) * 16;
    k_sgemm<<<dim3((N + 63) / 64, (M + 63) / 64, nz), 256, 0, a->st>>>(A_, sam, sak, B_, sbk, sbn, a->splitk, N, M, N, K, nullptr, -1, kslice);
    LAUNCH_OK(a);
    k_splitk_reduce<<<(M * N + 255) / 256, 256, 0, a->st>>>(a->splitk, nz, M, N, bias, act, C_, ldc);
    LAUNCH_OK(a);
    return SC_OK;
  }
  if (tiles <= 16 && K >= 128 && (size_t)M * N <= (size_t)a->B * HID) {
    // few output tiles (the small GEMMs of the metadata / head networks and their weight gradients, reduction = batch): a handful of
    // CTAs would walk the whole reduction 16 elements at a time, two barriers per step -- cut it into 16-element-multiple slices instead
    const int nz = K >= 512 ? 16 : 8, kslice = ((K + nz * 16 - 1) / (nz * 16)) * 16;
    k_sgemm<<<dim3((N + 63) / 64, (M + 63) / 64, nz), 256, 0, a->st>>>(A_, sam, sak, B_, sbk, sbn, a->splitk, N, M, N, K, nullptr, -1, kslice);
    LAUNCH_OK(a);
    k_splitk_reduce<<<(M * N + 255) / 256, 256, 0, a->st>>>(a->splitk, nz, M, N, bias, act, C_, ldc);
    LAUNCH_OK(a);
    return SC_OK;
  }
  k_sgemm<<<dim3((N + 63) / 64, (M + 63) / 64), 256, 0, a->st>>>(A_, sam, sak, B_, sbk, sbn, C_, ldc, M, N, K, bias, act, K);
  LAUNCH_OK(a);
  return SC_OK;
}
static int colsum(sc_calib* a, const float* X, long ldx, const float* w, int Bn, int N, float* out) {
  k_colsum<<<(N + 31) / 32, 256, 0, a->st>>>(X, ldx, w, Bn, N, out);
  LAUNCH_OK(a);
  return SC_OK;
}

// BatchNorm + conv backward of layer l: dA (image stride dA_stride) is the gradient w.r.t. the layer's activation output.
// Leaves dz in `dz_out`, writes d gamma / d beta / d conv weight / d conv bias into G, and (l > 0) the input gradient into dA_prev.
static int layer_backward(sc_calib* a, TrunkWs& t, const NetP& P, const NetP& G, int l, const float* dA, size_t dA_stride, float* dz_out,
                          const float* x_in, size_t x_stride, float* dA_prev, int act) {
  const int Cn = l == 0 ? C1 : (l == 1 ? C2 : C3);
  const int Ho = l == 0 ? H1C : (l == 1 ? H2C : H3C);
  const int HW = Ho * Ho;
  const int Bn = a->B;
  const size_t astride = l == 2 ? (size_t)FEAT : (size_t)Cn * HW;
  ChanReduceP r;
  memset(&r, 0, sizeof(r));
  r.z = t.z[l]; r.a = t.a[l]; r.dA = dA; r.a_img_stride = astride; r.dA_img_stride = dA_stride; r.mean = t.mean[l]; r.rstd = t.rstd[l];
  r.B = Bn; r.C = Cn; r.HW = HW; r.act = act; r.mode = 1; r.partial = a->red_partial;
  const int nch = Bn < a->nch ? Bn : a->nch;
  k_chan_reduce<<<dim3(Cn, nch), 256, 0, a->st>>>(r);
  LAUNCH_OK(a);
  BnBwdFinalizeP f{a->red_partial, nch, G.g[l], G.be[l], a->bn_c1, a->bn_c2, (double)Bn * HW};
  k_bn_bwd_finalize<<<1, 32, 0, a->st>>>(f, Cn);
  LAUNCH_OK(a);
  const size_t total = (size_t)Bn * Cn * HW;
  k_bn_bwd_apply<<<grid_for(total), 256, 0, a->st>>>(dA, dA_stride, t.a[l], astride, t.z[l], t.scale[l], t.mean[l], t.rstd[l], a->bn_c1, a->bn_c2, Cn,
                                                      HW, total, act, dz_out);
  LAUNCH_OK(a);
  // conv bias gradient = sum dz
  memset(&r, 0, sizeof(r));
  r.z = dz_out; r.dA = dz_out; r.dA_img_stride = (size_t)Cn * HW; r.B = Bn; r.C = Cn; r.HW = HW; r.mode = 2; r.partial = a->red_partial;
  k_chan_reduce<<<dim3(Cn, nch), 256, 0, a->st>>>(r);
  LAUNCH_OK(a);
  k_bias_grad_finalize<<<1, 32, 0, a->st>>>(a->red_partial, nch, Cn, G.cb[l]);
  LAUNCH_OK(a);
  // conv weight gradient: partials per block, then a fixed-order sum
  const int nb = a->wg_blocks;
  if (l == 0) {
    k_conv_wgrad_l1<<<nb, 256, 0, a->st>>>(x_in, x_stride, dz_out, Bn, a->wg_partial);
    LAUNCH_OK(a);
    k_sum_partials<<<(C1 * TAPS + 255) / 256, 256, 0, a->st>>>(a->wg_partial, nb, C1 * TAPS, G.cw[0]);
  } else if (l == 1) {
    k_conv_wgrad<C1, C2, H1C, H2C><<<nb, 256, 52 * 1024, a->st>>>(x_in, x_stride, dz_out, Bn, a->wg_partial);
    LAUNCH_OK(a);
    k_sum_partials<<<(C2 * C1 * TAPS + 255) / 256, 256, 0, a->st>>>(a->wg_partial, nb, C2 * C1 * TAPS, G.cw[1]);
  } else {
    k_conv_wgrad<C2, C3, H2C, H3C><<<nb, 256, 101 * 1024, a->st>>>(x_in, x_stride, dz_out, Bn, a->wg_partial);
    LAUNCH_OK(a);
    k_sum_partials<<<(C3 * C2 * TAPS + 255) / 256, 256, 0, a->st>>>(a->wg_partial, nb, C3 * C2 * TAPS, G.cw[2]);
  }
  LAUNCH_OK(a);
  if (l == 2) {
    using G3 = DgradGeom<C2, C3, H2C, H3C>;
    k_conv_dgrad<C2, C3, H2C, H3C><<<dim3(G3::BLOCKS_PER_IMG, Bn), 256, 0, a->st>>>(dz_out, P.cw[2], dA_prev);
    LAUNCH_OK(a);
  } else if (l == 1) {
    using G2 = DgradGeom<C1, C2, H1C, H2C>;
    k_conv_dgrad<C1, C2, H1C, H2C><<<dim3(G2::BLOCKS_PER_IMG, Bn), 256, 0, a->st>>>(dz_out, P.cw[1], dA_prev);
    LAUNCH_OK(a);
  }
  return SC_OK;
}

// whole trunk backward from dfeat[:, 0:5408]
static int trunk_backward(sc_calib* a, TrunkWs& t, const NetP& P, const NetP& G, const float* img, int act) {
  int rc;
  if ((rc = layer_backward(a, t, P, G, 2, a->dfeat, FEAT, a->dz3, t.a[1], (size_t)C2 * H2C * H2C, a->dA2, act))) return rc;
  if ((rc = layer_backward(a, t, P, G, 1, a->dA2, (size_t)C2 * H2C * H2C, a->dA2, t.a[0], (size_t)C1 * H1C * H1C, a->dA1, act))) return rc;
  return layer_backward(a, t, P, G, 0, a->dA1, (size_t)C1 * H1C * H1C, a->dA1, img, (size_t)IMG * IMG, nullptr, act);
}

// side network weight gradients from d16 [B][16], dh1 [B][128], activations h1 and x1 [B][K1]
static int side_wgrads(sc_calib* a, const float* x1, const float* h1, int K1, const NetP& G) {
  int rc;
  if ((rc = sgemm(a, a->d16, 1, a->side, h1, HID, 1, G.fb_w, HID, a->side, HID, a->B, nullptr, -1))) return rc;
  if ((rc = colsum(a, a->d16, a->side, nullptr, a->B, a->side, G.fb_b))) return rc;
  if ((rc = sgemm(a, a->dh1, 1, HID, x1, K1, 1, G.fa_w, K1, HID, K1, a->B, nullptr, -1))) return rc;
  return colsum(a, a->dh1, HID, nullptr, a->B, HID, G.fa_b);
}

static AdamP adam_hyper_c(double lr, int64_t t) {
  const double b1 = 0.9, b2 = 0.999;
  AdamP h;
  const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
  h.lr_neg_step = (float)(-(lr / bc1));
  h.bc2_sqrt = (float)sqrt(bc2);
  h.one_minus_b1 = (float)(1.0 - b1);
  h.b2 = (float)b2;
  h.one_minus_b2 = (float)(1.0 - b2);
  h.eps = 1e-8f;
  return h;
}
static int adam_step(sc_calib* a, int net, const AdamP* hp = nullptr) {       // hyper-parameters of optimizer `net`: step_dev->adam[net] (written by learn())
  const size_t n = net == 0 ? a->La.size : a->Lc.size, off = a->off_net[net];
  k_adam<<<grid_for(n), 256, 0, a->st>>>(a->params + off, a->grads + off, a->adam_m + off, a->adam_v + off, n, hp ? hp : &a->step_dev->adam[net]);
  LAUNCH_OK(a);
  return SC_OK;
}

extern "C" {

size_t sc_calib_arena_bytes(const sc_calib_config* cfg) {
  if (check_cfg_c(cfg) != SC_OK) return 0;
  sc_calib tmp;
  init_dims_c(&tmp, cfg);
  return carve_c(&tmp, nullptr);
}

int sc_calib_create(const sc_calib_config* cfg, void* arena, size_t arena_bytes, void* cuda_stream, sc_calib** out) {
  if (!out) return failc(SC_ERR_ARG, "null out");
  *out = nullptr;
  int rc = check_cfg_c(cfg);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return failc(SC_ERR_CUDA, "no CUDA device: libsmartcal_b200 has no CPU fallback");
  }
  sc_calib* a = new sc_calib();
  init_dims_c(a, cfg);
  a->st = (cudaStream_t)cuda_stream;
  const size_t need = carve_c(a, nullptr);
  if (arena) {
    if (arena_bytes < need) { delete a; return failc(SC_ERR_ARG, "arena too small: %zu < %zu", arena_bytes, need); }
    a->arena = (char*)arena; a->own_arena = false;
  } else {
    if (cudaMalloc((void**)&a->arena, need) != cudaSuccess) { delete a; return failc(SC_ERR_CUDA, "cudaMalloc(%zu) failed", need); }
    a->own_arena = true;
  }
  a->arena_bytes = need;
  carve_c(a, a->arena);
  CKC(cudaMemsetAsync(a->arena, 0, need, a->st));
  // BatchNorm buffers start at running_mean = 0, running_var = 1 (nn.BatchNorm2d defaults)
  {
    std::vector<float> init(160, 0.f);
    for (int i = 16; i < 32; ++i) init[i] = 1.f;
    for (int i = 64; i < 96; ++i) init[i] = 1.f;
    for (int i = 128; i < 160; ++i) init[i] = 1.f;
    for (int n = 0; n < 6; ++n) CKC(cudaMemcpyAsync(a->bnbuf[n], init.data(), 160 * sizeof(float), cudaMemcpyHostToDevice, a->st));
    CKC(cudaStreamSynchronize(a->st));
  }
  const size_t stage = (size_t)2 * IMG * IMG + 2 * a->SKY + 2 * a->A + 8 + (size_t)2 * a->B * a->A;
  CKC(cudaHostAlloc((void**)&a->h_stage, stage * sizeof(float), cudaHostAllocDefault));
  CKC(cudaHostAlloc((void**)&a->h_stats, 16 * sizeof(float), cudaHostAllocDefault));
  CKC(cudaHostAlloc((void**)&a->h_idx, (size_t)a->B * sizeof(int), cudaHostAllocDefault));
  a->h_u01 = nullptr;
  a->per_beta = 0.4;                                                  // PER.beta (demix_td3.py:32)
  if (a->td3) {
    CKC(cudaHostAlloc((void**)&a->h_u01, ((size_t)a->B + 2) * sizeof(double), cudaHostAllocDefault));
    // ADMM bookkeeping blocks of the actor loop (k_admm_update, sc_per.cu): constant from step to step
    AdmmP* blk = a->admm_host;
    for (int i = 0; i < 6; ++i) {
      AdmmP& p = blk[i];
      memset(&p, 0, sizeof(p));
      p.mode = i == 0 ? 0 : 1; p.it = i - 1;
      p.pi = a->pi; p.hint = a->r_hint; p.Ap = a->A; p.bidx = a->bidx; p.B = a->B; p.A = a->A;
      p.lambda = a->lambda; p.y0 = a->admm_y0; p.a0 = a->admm_a0; p.rho = a->admm_rho_dev;
      p.rho0 = cfg->admm_rho; p.corr_min = 0.5f;                      // demix_td3.py:400,403
      p.stats_rho = nullptr;
    }
    CKC(cudaMemcpyAsync(a->admm_dev, blk, 6 * sizeof(AdmmP), cudaMemcpyHostToDevice, a->st));
    CKC(cudaStreamSynchronize(a->st));
  }
  CKC(cudaFuncSetAttribute(k_conv_wgrad<C1, C2, H1C, H2C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
  CKC(cudaFuncSetAttribute(k_conv_wgrad<C2, C3, H2C, H3C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 101 * 1024));
  {
    const char* e = getenv("SMARTCAL_CONV");                    // "ffma" keeps the direct FP32 convolutions (A/B runs)
    a->conv_tc = a->trunk && !(e && !strcmp(e, "ffma"));
    const char* d = getenv("SMARTCAL_CONV_DBG");
    a->conv_dbg = d ? atoi(d) : 0;
    if (!getenv("SMARTCAL_STAMPS")) a->conv_ts = nullptr;
    if (a->conv_tc) {
      for (int n = 0; n < 6; ++n) {
        float* hl2 = a->conv_hl[n];
        float* hl3 = hl2 + (size_t)2 * C2 * C1 * TAPS;
        for (int part = 0; part < 2; ++part) {
          if ((rc = encode_map_f32_2d(&a->conv_map[n][0][part], hl2 + (size_t)part * C2 * C1 * TAPS, C1 * TAPS, C2, 32))) return rc;
          if ((rc = encode_map_f32_2d(&a->conv_map[n][1][part], hl3 + (size_t)part * C3 * C2 * TAPS, C2 * TAPS, C3, 32))) return rc;
        }
      }
      CKC(conv_tc_upload_tables());
      CKC(cudaFuncSetAttribute(k_conv_fwd_tc<C1, H1C, H2C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_tc_smem_bytes()));
      CKC(cudaFuncSetAttribute(k_conv_fwd_tc<C2, H2C, H3C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_tc_smem_bytes()));
    }
  }
  CKC(cudaMalloc((void**)&a->step_dev, sizeof(sc_calib::StepScalars)));
  CKC(cudaHostAlloc((void**)&a->step_host, 2 * sizeof(sc_calib::StepScalars), cudaHostAllocDefault));
  for (auto& e : a->ev_step) CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CKC(cudaStreamCreateWithFlags(&a->st_cap, cudaStreamNonBlocking));
  a->step_k = 0;
  a->graph[0] = a->graph[1] = nullptr;
  a->graph_seen[0] = a->graph_seen[1] = 0;
  a->graph_launches[0] = a->graph_launches[1] = 0;
  {
    const char* e = getenv("SMARTCAL_GRAPH");
    a->use_graph = (e && e[0] == '0') ? 0 : 1;
  }
  a->mem_cntr = a->learn_counter = a->t_actor = a->t_critic = a->launches = a->noise_ctr = 0;
  a->batch_ready = false;
  a->admm_rho = cfg->admm_rho;
  *out = a;
  return SC_OK;
}

int sc_calib_destroy(sc_calib* a) {
  if (!a) return SC_OK;
  cudaStreamSynchronize(a->st);
  for (auto& g : a->graph) if (g) cudaGraphExecDestroy(g);
  if (a->st_cap) { cudaStreamSynchronize(a->st_cap); cudaStreamDestroy(a->st_cap); }
  for (auto& e : a->ev_step) if (e) cudaEventDestroy(e);
  if (a->step_dev) cudaFree(a->step_dev);
  if (a->step_host) cudaFreeHost(a->step_host);
  if (a->h_stage) cudaFreeHost(a->h_stage);
  if (a->h_stats) cudaFreeHost(a->h_stats);
  if (a->h_idx) cudaFreeHost(a->h_idx);
  if (a->h_u01) cudaFreeHost(a->h_u01);
  if (a->own_arena && a->arena) cudaFree(a->arena);
  delete a;
  return SC_OK;
}
int sc_calib_sync(sc_calib* a) {
  if (!a) return failc(SC_ERR_ARG, "null handle");
  CKC(cudaStreamSynchronize(a->st));
  return SC_OK;
}
int64_t sc_calib_launch_count(sc_calib* a) { return a ? a->launches : 0; }
int sc_calib_get_per_beta(sc_calib* a, double* beta_out) {
  if (!a || !beta_out) return failc(SC_ERR_ARG, "null argument");
  if (!a->td3) return failc(SC_ERR_STATE, "not a prioritized (SC_CALIB_TD3) agent");
  *beta_out = a->per_beta;
  return SC_OK;
}
int sc_calib_set_per_beta(sc_calib* a, double beta) {
  if (!a) return failc(SC_ERR_ARG, "null handle");
  if (!a->td3) return failc(SC_ERR_STATE, "not a prioritized (SC_CALIB_TD3) agent");
  a->per_beta = beta;
  return SC_OK;
}

int sc_calib_tensor(sc_calib* a, int net, int slot, const char* name, sc_tensor_info* out) {
  if (!a || !name || !out) return failc(SC_ERR_ARG, "null argument");
  if (net < 0 || net > (a->td3 ? 5 : 4)) return failc(SC_ERR_ARG, "bad net id %d", net);
  const bool actor = net == SC_NET_ACTOR || net == SC_NET_TARGET_ACTOR;
  const CLayout& L = actor ? a->La : a->Lc;
  // BatchNorm buffers
  const char* dot = strrchr(name, '.');
  if (dot && (!strcmp(dot, ".running_mean") || !strcmp(dot, ".running_var") || !strcmp(dot, ".num_batches_tracked"))) {
    if (strncmp(name, "bn", 2) != 0 || name[2] < '1' || name[2] > '3') return failc(SC_ERR_ARG, "unknown buffer %s", name);
    const int l = name[2] - '1';
    if (!strcmp(dot, ".num_batches_tracked")) { *out = sc_tensor_info{a->nbt[net] + l, 1, 1, 1, 8, 1}; return SC_OK; }
    float* rm = a->bnbuf[net] + (l == 0 ? 0 : l == 1 ? 32 : 96);
    const int C = l == 0 ? 16 : 32;
    *out = sc_tensor_info{!strcmp(dot, ".running_mean") ? rm : rm + C, 1, C, C, 4, 0};
    return SC_OK;
  }
  float* base;
  if (net <= 2) {
    float* blocks[3] = {a->params, a->adam_m, a->adam_v};
    if (slot < 0 || slot > 3) return failc(SC_ERR_ARG, "bad slot %d", slot);
    base = (slot == 3 ? a->grads : blocks[slot]) + a->off_net[net];
  } else {
    if (slot != SC_SLOT_PARAM) return failc(SC_ERR_ARG, "targets have no optimizer state");
    base = a->targets + (size_t)(net - 3) * a->Lc.size;       // [t1 | t2 | target actor]
  }
  if (!strcmp(name, "*")) { *out = sc_tensor_info{base, 1, (int64_t)L.size, (int64_t)L.size, 4, 0}; return SC_OK; }
  for (auto& d : L.t)
    if (d.name == name) { *out = sc_tensor_info{base + d.off, d.rows, d.cols, d.cols, 4, 0}; return SC_OK; }
  return failc(SC_ERR_ARG, "no tensor named %s", name);
}
int sc_calib_buffer(sc_calib* a, const char* name, sc_tensor_info* out) {
  if (!a || !name || !out) return failc(SC_ERR_ARG, "null argument");
  auto it = a->reg.find(name);
  if (it == a->reg.end()) return failc(SC_ERR_ARG, "no buffer named %s", name);
  *out = it->second;
  return SC_OK;
}
int sc_calib_get_counters(sc_calib* a, int64_t* mem_cntr, int64_t* learn_counter) {
  if (!a) return failc(SC_ERR_ARG, "null handle");
  if (mem_cntr) *mem_cntr = a->mem_cntr;
  if (learn_counter) *learn_counter = a->learn_counter;
  return SC_OK;
}
int sc_calib_set_counters(sc_calib* a, int64_t mem_cntr, int64_t learn_counter) {
  if (!a) return failc(SC_ERR_ARG, "null handle");
  a->mem_cntr = mem_cntr; a->learn_counter = learn_counter;
  return SC_OK;
}
int sc_calib_set_scalars(sc_calib* a, float rho, float admm_rho) {
  if (!a) return failc(SC_ERR_ARG, "null handle");
  a->admm_rho = admm_rho;
  CKC(cudaMemcpyAsync(a->rho, &rho, sizeof(float), cudaMemcpyHostToDevice, a->st));
  CKC(cudaStreamSynchronize(a->st));
  if (a->td3) {                     // agent.admm_rho (demix_td3.py:400): the value every learn() starts its ADMM loop from
    for (auto& p : a->admm_host) p.rho0 = admm_rho;
    CKC(cudaMemcpyAsync(a->admm_dev, a->admm_host, 6 * sizeof(AdmmP), cudaMemcpyHostToDevice, a->st));
    CKC(cudaStreamSynchronize(a->st));
  }
  return SC_OK;
}

int sc_calib_store_host(sc_calib* a, const float* img, const float* sky, const float* img2, const float* sky2, const float* action,
                        const float* hint, float reward, int done) {
  if (!a || !sky || !sky2 || !action || !hint || (a->trunk && (!img || !img2))) return failc(SC_ERR_ARG, "null argument");
  const size_t slot = (size_t)(a->mem_cntr % a->C), NI = a->trunk ? (size_t)IMG * IMG : 0;       // (no influence maps in the MLP family)
  CKC(cudaStreamSynchronize(a->st));                      // the staging buffer may still be in flight
  float* h = a->h_stage;
  if (NI) { memcpy(h, img, NI * 4); memcpy(h + NI, img2, NI * 4); }
  float* s = h + 2 * NI;
  memcpy(s, sky, a->SKY * 4); memcpy(s + a->SKY, sky2, a->SKY * 4);
  float* t = s + 2 * a->SKY;
  memcpy(t, action, a->A * 4); memcpy(t + a->A, hint, a->A * 4);
  t[2 * a->A] = reward; t[2 * a->A + 1] = done ? 1.f : 0.f;
  if (NI) {
    CKC(cudaMemcpyAsync(a->r_img + slot * NI, h, NI * 4, cudaMemcpyHostToDevice, a->st));
    CKC(cudaMemcpyAsync(a->r_img2 + slot * NI, h + NI, NI * 4, cudaMemcpyHostToDevice, a->st));
  }
  CKC(cudaMemcpyAsync(a->r_sky + slot * a->SKY, s, a->SKY * 4, cudaMemcpyHostToDevice, a->st));
  CKC(cudaMemcpyAsync(a->r_sky2 + slot * a->SKY, s + a->SKY, a->SKY * 4, cudaMemcpyHostToDevice, a->st));
  CKC(cudaMemcpyAsync(a->r_act + slot * a->A, t, a->A * 4, cudaMemcpyHostToDevice, a->st));
  CKC(cudaMemcpyAsync(a->r_hint + slot * a->A, t + a->A, a->A * 4, cudaMemcpyHostToDevice, a->st));
  CKC(cudaMemcpyAsync(a->r_rew + slot, t + 2 * a->A, 4, cudaMemcpyHostToDevice, a->st));
  CKC(cudaMemcpyAsync(a->r_done + slot, t + 2 * a->A + 1, 4, cudaMemcpyHostToDevice, a->st));
  if (a->td3) {                                   // PER.store_transition(error=None): max-leaf priority, tree.add (demix_td3.py:60-67)
    k_tree_add<<<1, 256, 0, a->st>>>(a->tree, a->C, (int)slot, 100.0);
    LAUNCH_OK(a);
  }
  a->mem_cntr++;
  return SC_OK;
}

int sc_calib_set_batch_host(sc_calib* a, const int32_t* idx, int n) {
  if (!a || !idx) return failc(SC_ERR_ARG, "null argument");
  if (n != a->B) return failc(SC_ERR_ARG, "need %d indices", a->B);
  const int64_t lim = a->mem_cntr < a->C ? a->mem_cntr : a->C;
  for (int i = 0; i < n; ++i)
    if (idx[i] < 0 || idx[i] >= lim) return failc(SC_ERR_ARG, "batch index %d out of range", idx[i]);
  CKC(cudaStreamSynchronize(a->st));
  memcpy(a->h_idx, idx, (size_t)n * sizeof(int));
  CKC(cudaMemcpyAsync(a->bidx, a->h_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, a->st));
  a->batch_ready = true;
  return SC_OK;
}

int sc_calib_soft_update(sc_calib* a, float tau) {
  if (!a) return failc(SC_ERR_ARG, "null handle");
  // calib_sac.py:271-276: parameters() only; the targets' BatchNorm buffers are never copied
  CKC(launch_soft_update(a->params + a->off_net[1], a->targets, 2 * a->Lc.size, tau, a->st));
  a->launches++;
  if (a->td3) {
    // demix_td3.py:408-432 walks state_dict(): the target actor moves too, and so do the BatchNorm buffers of all three targets
    CKC(launch_soft_update(a->params + a->off_net[0], a->targets + 2 * a->Lc.size, a->La.size, tau, a->st));
    a->launches++;
    const int src[3] = {0, 1, 2}, dst[3] = {5, 3, 4};
    const float omt = (float)(1.0 - (double)tau);
    for (int i = 0; i < 3; ++i) {
      CKC(launch_soft_update(a->bnbuf[src[i]], a->bnbuf[dst[i]], 160, tau, a->st));
      k_nbt_blend<<<1, 32, 0, a->st>>>(a->nbt[src[i]], a->nbt[dst[i]], tau, omt);
      LAUNCH_OK(a);
      a->launches++;
    }
  }
  return SC_OK;
}

// choose_action (calib_sac.py:285-295): actor in eval mode (running statistics), one observation
int sc_calib_act(sc_calib* a, const float* img, const float* sky, const float* eps, float* action_out) {
  if (!a || !sky || !action_out || (a->trunk && !img)) return failc(SC_ERR_ARG, "null argument");
  const size_t NI = a->trunk ? (size_t)IMG * IMG : 0;
  CKC(cudaStreamSynchronize(a->st));
  if (NI) memcpy(a->h_stage, img, NI * 4);
  memcpy(a->h_stage + NI, sky, a->SKY * 4);
  if (eps) memcpy(a->h_stage + NI + a->SKY, eps, a->A * 4);
  if (NI) CKC(cudaMemcpyAsync(a->mb_img2, a->h_stage, NI * 4, cudaMemcpyHostToDevice, a->st));
  CKC(cudaMemcpyAsync(a->mb_sky2, a->h_stage + NI, a->SKY * 4, cudaMemcpyHostToDevice, a->st));
  if (a->td3) {
    // demix_td3.py:444-455 after the warm-up: eval-mode actor, mu + noise (A values drawn by the caller, or none), clamp to [-1, 1]
    if (eps) CKC(cudaMemcpyAsync(a->anoise, a->h_stage + NI + a->SKY, a->A * 4, cudaMemcpyHostToDevice, a->st));
    const NetP P = net_ptrs(a->params + a->off_net[0], a->La, true);
    TrunkWs& t = a->ws[0];
    int rc = trunk_forward(a, t, P, 0, a->mb_img2, ACT_ELU, 0, 1);
    if (rc) return rc;
    k_side_fwd<<<1, 128, 0, a->st>>>(a->mb_sky2, a->SKY, nullptr, 0, P.fa_w, P.fa_b, P.fb_w, P.fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    if ((rc = sgemm(a, t.feat, a->feat, 1, P.h_w, 1, a->feat, a->x21, HID, 1, HID, a->feat, P.h_b, ACT_ELU))) return rc;
    k_td3_head<<<1, 128, 0, a->st>>>(a->x21, P.mu_w, P.mu_b, a->A, nullptr, 0, eps ? a->anoise : nullptr, a->a2);
    LAUNCH_OK(a);
    CKC(cudaMemcpyAsync(a->h_stage, a->a2, a->A * 4, cudaMemcpyDeviceToHost, a->st));
    CKC(cudaStreamSynchronize(a->st));
    memcpy(action_out, a->h_stage, a->A * 4);
    a->batch_ready = false;
    return SC_OK;
  }
  if (eps) CKC(cudaMemcpyAsync(a->eps_next, a->h_stage + NI + a->SKY, a->A * 4, cudaMemcpyHostToDevice, a->st));
  else { CKC(launch_fill_normal(a->eps_next, a->A, a->cfg.rng_seed, 0x9000u + (uint64_t)(a->noise_ctr++), a->st)); a->launches++; }
  const NetP P = net_ptrs(a->params + a->off_net[0], a->La, true);
  TrunkWs& t = a->ws[0];
  int rc = a->trunk ? trunk_forward(a, t, P, 0, a->mb_img2, ACT_ELU, 0, 1) : SC_OK;
  if (rc) return rc;
  k_side_fwd<<<1, 128, 0, a->st>>>(a->mb_sky2, a->SKY, nullptr, 0, P.fa_w, P.fa_b, P.fb_w, P.fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
  LAUNCH_OK(a);
  if ((rc = sgemm(a, t.feat, a->feat, 1, P.h_w, 1, a->feat, a->x21, HID, 1, HID, a->feat, P.h_b, ACT_ELU))) return rc;
  k_actor_head<<<1, 128, 0, a->st>>>(a->x21, P.mu_w, P.mu_b, P.sg_w, P.sg_b, a->eps_next, a->A, a->a2, a->sig2, a->raw2, a->logp2, nullptr, nullptr, a->sigma_mode);
  LAUNCH_OK(a);
  CKC(cudaMemcpyAsync(a->h_stage, a->a2, a->A * 4, cudaMemcpyDeviceToHost, a->st));
  CKC(cudaStreamSynchronize(a->st));
  memcpy(action_out, a->h_stage, a->A * 4);
  a->batch_ready = false;       // the minibatch staging buffers were used as scratch
  return SC_OK;
}

// every launch of one learn() (calib_sac.py:317-378 / demix_sac.py:594-682) on a->st; with device-generated noise the sequence
// is identical from step to step (step-dependent scalars live in step_dev) and is replayed from a CUDA graph
static int calib_body(sc_calib* a, bool host_noise, bool hint, int dual) {
  const int B = a->B, A = a->A, SKY = a->SKY;
  const size_t NI = (size_t)IMG * IMG;
  int rc;
  if (!host_noise) {
    CKC(launch_fill_normal_p(a->eps_next, B * A, a->cfg.rng_seed, &a->step_dev->noise[0], a->st));
    CKC(launch_fill_normal_p(a->eps_cur, B * A, a->cfg.rng_seed, &a->step_dev->noise[1], a->st));
    a->launches += 2;
  }
  // ---- minibatch gather (replaces the host fancy-indexing + H2D copies of calib_sac.py:304-315) ----
  if (a->trunk) {
    k_gather<<<dim3(16, B), 256, 0, a->st>>>(a->r_img, a->bidx, NI, a->mb_img);  LAUNCH_OK(a);
    k_gather<<<dim3(16, B), 256, 0, a->st>>>(a->r_img2, a->bidx, NI, a->mb_img2);  LAUNCH_OK(a);
  }
  k_gather<<<dim3(1, B), 128, 0, a->st>>>(a->r_sky, a->bidx, SKY, a->mb_sky);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 128, 0, a->st>>>(a->r_sky2, a->bidx, SKY, a->mb_sky2);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 128, 0, a->st>>>(a->r_act, a->bidx, A, a->mb_act);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 128, 0, a->st>>>(a->r_hint, a->bidx, A, a->mb_hint);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 32, 0, a->st>>>(a->r_rew, a->bidx, 1, a->mb_rew);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 32, 0, a->st>>>(a->r_done, a->bidx, 1, a->mb_done);  LAUNCH_OK(a);

  const NetP PA = net_ptrs(a->params + a->off_net[0], a->La, true), GA = net_ptrs(a->grads + a->off_net[0], a->La, true);
  NetP PC[2], GC[2], PT[2];
  for (int k = 0; k < 2; ++k) {
    PC[k] = net_ptrs(a->params + a->off_net[1 + k], a->Lc, false);
    GC[k] = net_ptrs(a->grads + a->off_net[1 + k], a->Lc, false);
    PT[k] = net_ptrs(a->targets + (size_t)k * a->Lc.size, a->Lc, false);
  }
  const float alpha = (float)a->cfg.alpha, gamma = (float)a->cfg.gamma;

  // ---- 1. policy on s' (no grad; BatchNorm in train mode like the reference) ----
  {
    TrunkWs& t = a->ws[0];
    if (a->trunk && (rc = trunk_forward(a, t, PA, 0, a->mb_img2, ACT_ELU, 1, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky2, SKY, nullptr, 0, PA.fa_w, PA.fa_b, PA.fb_w, PA.fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    if ((rc = sgemm(a, t.feat, a->feat, 1, PA.h_w, 1, a->feat, a->x21, HID, B, HID, a->feat, PA.h_b, ACT_ELU))) return rc;
    k_actor_head<<<B, 128, 0, a->st>>>(a->x21, PA.mu_w, PA.mu_b, PA.sg_w, PA.sg_b, a->eps_next, A, a->a2, a->sig2, a->raw2, a->logp2, nullptr, nullptr, a->sigma_mode);
    LAUNCH_OK(a);
  }
  // ---- 2. target critics on (s', sky', a') ----
  for (int k = 0; k < 2; ++k) {
    TrunkWs& t = a->ws[0];
    if (a->trunk && (rc = trunk_forward(a, t, PT[k], 3 + k, a->mb_img2, ACT_RELU, 1, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky2, SKY, a->a2, A, PT[k].fa_w, PT[k].fa_b, PT[k].fb_w, PT[k].fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    k_critic_head<<<B, 256, 0, a->st>>>(t.feat, PT[k].h_w, PT[k].h_b, a->qt[k], a->feat);
    LAUNCH_OK(a);
  }
  k_target<<<(B + 127) / 128, 128, 0, a->st>>>(a->qt[0], a->qt[1], a->logp2, a->mb_rew, a->mb_done, gamma, alpha, B, a->y);
  LAUNCH_OK(a);
  // ---- 3. critics: forward on (s, sky, a), MSE loss, backward, Adam ----
  for (int k = 0; k < 2; ++k) {
    TrunkWs& t = a->ws[1 + k];
    if (a->trunk && (rc = trunk_forward(a, t, PC[k], 1 + k, a->mb_img, ACT_RELU, 1, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky, SKY, a->mb_act, A, PC[k].fa_w, PC[k].fa_b, PC[k].fb_w, PC[k].fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    k_critic_head<<<B, 256, 0, a->st>>>(t.feat, PC[k].h_w, PC[k].h_b, a->q[k], a->feat);
    LAUNCH_OK(a);
  }
  for (int k = 0; k < 2; ++k) {
    TrunkWs& t = a->ws[1 + k];
    k_critic_seed<<<1, 256, 0, a->st>>>(a->q[k], a->y, B, a->dq, a->stats + k);
    LAUNCH_OK(a);
    k_outer<<<grid_for((size_t)B * a->feat), 256, 0, a->st>>>(a->dq, PC[k].h_w, B, a->feat, a->dfeat);
    LAUNCH_OK(a);
    if ((rc = colsum(a, t.feat, a->feat, a->dq, B, a->feat, GC[k].h_w))) return rc;       // d head.weight
    if ((rc = colsum(a, a->dq, 1, nullptr, B, 1, GC[k].h_b))) return rc;             // d head.bias
    k_side_bwd<<<B, 128, 0, a->st>>>(a->dfeat + a->flat, a->feat, t.o16, t.h1, PC[k].fb_w, PC[k].fa_w, a->K1C, SKY, A, a->d16, a->dh1, nullptr, 0, a->side);
    LAUNCH_OK(a);
    if ((rc = side_wgrads(a, t.x1, t.h1, a->K1C, GC[k]))) return rc;
    if (a->trunk && (rc = trunk_backward(a, t, PC[k], GC[k], a->mb_img, ACT_RELU))) return rc;
  }
  for (int k = 0; k < 2; ++k)
    if ((rc = adam_step(a, 1 + k))) return rc;
  // ---- 4. actor: policy on s (reparameterised), Q_k(s, pi) with the updated critics, backward, Adam ----
  {
    TrunkWs& t = a->ws[3];
    if (a->trunk && (rc = trunk_forward(a, t, PA, 0, a->mb_img, ACT_ELU, 1, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky, SKY, nullptr, 0, PA.fa_w, PA.fa_b, PA.fb_w, PA.fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    if ((rc = sgemm(a, t.feat, a->feat, 1, PA.h_w, 1, a->feat, a->x21, HID, B, HID, a->feat, PA.h_b, ACT_ELU))) return rc;
    k_actor_head<<<B, 128, 0, a->st>>>(a->x21, PA.mu_w, PA.mu_b, PA.sg_w, PA.sg_b, a->eps_cur, A, a->pi, a->sig, a->raw, a->logp, a->mu, a->u, a->sigma_mode);
    LAUNCH_OK(a);
  }
  for (int k = 0; k < 2; ++k) {
    TrunkWs& t = a->ws[0];
    if (a->trunk && (rc = trunk_forward(a, t, PC[k], 1 + k, a->mb_img, ACT_RELU, 1, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky, SKY, a->pi, A, PC[k].fa_w, PC[k].fa_b, PC[k].fb_w, PC[k].fb_b, a->pside[k].x1, a->pside[k].h1,
                                     a->pside[k].o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    k_critic_head<<<B, 256, 0, a->st>>>(t.feat, PC[k].h_w, PC[k].h_b, a->qpi[k], a->feat);
    LAUNCH_OK(a);
  }
  k_actor_seed<<<(B + 127) / 128, 128, 0, a->st>>>(a->qpi[0], a->qpi[1], PC[0].h_w, PC[1].h_w, B, a->ds[0], a->ds[1], a->qmin, a->side, a->flat);
  LAUNCH_OK(a);
  for (int k = 0; k < 2; ++k) {
    k_side_bwd<<<B, 128, 0, a->st>>>(a->ds[k], a->side, a->pside[k].o16, a->pside[k].h1, PC[k].fb_w, PC[k].fa_w, a->K1C, SKY, A, a->d16, a->dh1, a->dact, k, a->side);
    LAUNCH_OK(a);
  }
  if (hint) {
    k_kld_coef<<<1, 256, 0, a->st>>>(a->pi, a->mb_hint, B * A, a->cfg.hint_threshold, a->admm_rho, a->rho, a->hint_out, 0, a->kld_mode);
    LAUNCH_OK(a);
  }
  k_actor_loss<<<1, 256, 0, a->st>>>(a->logp, a->qmin, B, alpha, hint ? a->hint_out : nullptr, a->stats + 2);
  LAUNCH_OK(a);
  {
    TrunkWs& t = a->ws[3];
    k_actor_head_bwd<<<B, 128, 0, a->st>>>(a->dact, a->pi, a->sig, a->raw, a->eps_cur, a->mu, a->u, a->mb_hint, hint ? a->hint_out : nullptr, alpha / (float)B, A,
                                           PA.mu_w, PA.sg_w, a->x21, a->dmu, a->draw, a->dpre, a->sigma_mode, a->kld_mode);
    LAUNCH_OK(a);
    // heads
    if ((rc = sgemm(a, a->dmu, 1, A, a->x21, HID, 1, GA.mu_w, HID, A, HID, B, nullptr, -1))) return rc;
    if ((rc = colsum(a, a->dmu, A, nullptr, B, A, GA.mu_b))) return rc;
    if ((rc = sgemm(a, a->draw, 1, A, a->x21, HID, 1, GA.sg_w, HID, A, HID, B, nullptr, -1))) return rc;
    if ((rc = colsum(a, a->draw, A, nullptr, B, A, GA.sg_b))) return rc;
    // fc21: dW = dpre^T feat, db = colsum(dpre), dfeat = dpre W
    if ((rc = sgemm(a, a->dpre, 1, HID, t.feat, a->feat, 1, GA.h_w, a->feat, HID, a->feat, B, nullptr, -1))) return rc;
    if ((rc = colsum(a, a->dpre, HID, nullptr, B, HID, GA.h_b))) return rc;
    if ((rc = sgemm(a, a->dpre, HID, 1, PA.h_w, a->feat, 1, a->dfeat, a->feat, B, a->feat, HID, nullptr, -1))) return rc;
    k_side_bwd<<<B, 128, 0, a->st>>>(a->dfeat + a->flat, a->feat, t.o16, t.h1, PA.fb_w, PA.fa_w, SKY, SKY, 0, a->d16, a->dh1, nullptr, 0, a->side);
    LAUNCH_OK(a);
    if ((rc = side_wgrads(a, t.x1, t.h1, SKY, GA))) return rc;
    if (a->trunk && (rc = trunk_backward(a, t, PA, GA, a->mb_img, ACT_ELU))) return rc;
  }
  if ((rc = adam_step(a, 0))) return rc;
  if (dual && a->kld_mode == 0) {       // calib_sac.py:360-364: rho += admm_rho * g, g recomputed from the SAME actions
    k_kld_coef<<<1, 256, 0, a->st>>>(a->pi, a->mb_hint, B * A, a->cfg.hint_threshold, a->admm_rho, a->rho, a->hint_out, 1, 0);
    LAUNCH_OK(a);
  }
  if (dual && a->kld_mode == 1) {
    // demix_sac.py:660-670: a fresh no-grad sample from the UPDATED actor (training-mode BatchNorm: running statistics move again)
    if (!host_noise) {
      CKC(launch_fill_normal_p(a->eps_dual, B * A, a->cfg.rng_seed, &a->step_dev->noise[2], a->st));
      a->launches++;
    }
    TrunkWs& t = a->ws[3];
    if (a->trunk && (rc = trunk_forward(a, t, PA, 0, a->mb_img, ACT_ELU, 1, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky, SKY, nullptr, 0, PA.fa_w, PA.fa_b, PA.fb_w, PA.fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    if ((rc = sgemm(a, t.feat, a->feat, 1, PA.h_w, 1, a->feat, a->x21, HID, B, HID, a->feat, PA.h_b, ACT_ELU))) return rc;
    k_actor_head<<<B, 128, 0, a->st>>>(a->x21, PA.mu_w, PA.mu_b, PA.sg_w, PA.sg_b, a->eps_dual, A, a->a2, a->sig2, a->raw2, a->logp2, nullptr, nullptr, a->sigma_mode);
    LAUNCH_OK(a);
    k_kld_coef<<<1, 256, 0, a->st>>>(a->a2, a->mb_hint, B * A, a->cfg.hint_threshold, a->admm_rho, a->rho, a->hint_out + 4, 1, 1);
    LAUNCH_OK(a);
  }
  if ((rc = sc_calib_soft_update(a, (float)a->cfg.tau))) return rc;
  return SC_OK;
}


// every launch of one demixing TD3 learn() (demix_td3.py:459-603) on a->st.  `host_rng`: the PER uniforms and the target-noise scalar
// were uploaded by the caller (reference RNG streams); otherwise both come from the device Philox streams named in step_dev.
static int td3_body(sc_calib* a, bool host_rng, bool hint, int actor_step) {
  const int B = a->B, A = a->A, SKY = a->SKY;
  const size_t NI = (size_t)IMG * IMG;
  int rc;
  // ---- PER.sample_buffer (:80-120): stratified draw on the float64 tree -> tree indices, data indices, IS weights, sum(w) ----
  CKC(launch_per_sample(&a->step_dev->per_s, 1, B, a->st));
  a->launches++;
  if (!host_rng) {
    CKC(launch_fill_normal_p(a->tnoise, 1, a->cfg.rng_seed, &a->step_dev->noise[3], a->st));
    a->launches++;
  }
  k_gather<<<dim3(16, B), 256, 0, a->st>>>(a->r_img, a->bidx, NI, a->mb_img);  LAUNCH_OK(a);
  k_gather<<<dim3(16, B), 256, 0, a->st>>>(a->r_img2, a->bidx, NI, a->mb_img2);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 128, 0, a->st>>>(a->r_sky, a->bidx, SKY, a->mb_sky);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 128, 0, a->st>>>(a->r_sky2, a->bidx, SKY, a->mb_sky2);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 128, 0, a->st>>>(a->r_act, a->bidx, A, a->mb_act);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 128, 0, a->st>>>(a->r_hint, a->bidx, A, a->mb_hint);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 32, 0, a->st>>>(a->r_rew, a->bidx, 1, a->mb_rew);  LAUNCH_OK(a);
  k_gather<<<dim3(1, B), 32, 0, a->st>>>(a->r_done, a->bidx, 1, a->mb_done);  LAUNCH_OK(a);

  const NetP PA = net_ptrs(a->params + a->off_net[0], a->La, true), GA = net_ptrs(a->grads + a->off_net[0], a->La, true);
  const NetP PTA = net_ptrs(a->targets + 2 * a->Lc.size, a->La, true);
  NetP PC[2], GC[2], PT[2];
  for (int k = 0; k < 2; ++k) {
    PC[k] = net_ptrs(a->params + a->off_net[1 + k], a->Lc, false);
    GC[k] = net_ptrs(a->grads + a->off_net[1 + k], a->Lc, false);
    PT[k] = net_ptrs(a->targets + (size_t)k * a->Lc.size, a->Lc, false);
  }
  const float gamma = (float)a->cfg.gamma;

  // ---- 1. target actor on s' in eval mode (:484-491): a' = clamp(tanh(.) + n, -1, 1), ONE scalar n per step ----
  {
    TrunkWs& t = a->ws[0];
    if ((rc = trunk_forward(a, t, PTA, 5, a->mb_img2, ACT_ELU, 0, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky2, SKY, nullptr, 0, PTA.fa_w, PTA.fa_b, PTA.fb_w, PTA.fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    if ((rc = sgemm(a, t.feat, a->feat, 1, PTA.h_w, 1, a->feat, a->x21, HID, B, HID, a->feat, PTA.h_b, ACT_ELU))) return rc;
    k_td3_head<<<B, 128, 0, a->st>>>(a->x21, PTA.mu_w, PTA.mu_b, A, a->tnoise, host_rng ? 0 : 1, nullptr, a->a2);
    LAUNCH_OK(a);
  }
  // ---- 2. target critics on (s', a') in eval mode (:495-504) ----
  for (int k = 0; k < 2; ++k) {
    TrunkWs& t = a->ws[0];
    if ((rc = trunk_forward(a, t, PT[k], 3 + k, a->mb_img2, ACT_RELU, 0, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky2, SKY, a->a2, A, PT[k].fa_w, PT[k].fa_b, PT[k].fb_w, PT[k].fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat, 1);
    LAUNCH_OK(a);
    k_critic_head<<<B, 256, 0, a->st>>>(t.feat, PT[k].h_w, PT[k].h_b, a->qt[k], a->feat);
    LAUNCH_OK(a);
  }
  k_td3_target<<<(B + 127) / 128, 128, 0, a->st>>>(a->qt[0], a->qt[1], a->mb_rew, a->mb_done, gamma, B, a->y);
  LAUNCH_OK(a);
  // ---- 3. critics on (s, a): the priority pre-pass (:506-511) and the training pass (:516-517) are the same training-mode forward
  //         (same weights, same batch statistics), so it runs once and the running statistics take both momentum updates ----
  for (int k = 0; k < 2; ++k) {
    TrunkWs& t = a->ws[1 + k];
    if ((rc = trunk_forward(a, t, PC[k], 1 + k, a->mb_img, ACT_RELU, 1, B, 2))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky, SKY, a->mb_act, A, PC[k].fa_w, PC[k].fa_b, PC[k].fb_w, PC[k].fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat, 1);
    LAUNCH_OK(a);
    k_critic_head<<<B, 256, 0, a->st>>>(t.feat, PC[k].h_w, PC[k].h_b, a->q[k], a->feat);
    LAUNCH_OK(a);
  }
  k_td3_per_err<<<(B + 127) / 128, 128, 0, a->st>>>(a->q[0], a->q[1], a->y, B, a->per_err);
  LAUNCH_OK(a);
  CKC(launch_per_update(&a->step_dev->per_u, 1, B, a->st));          // PER.batch_update (:122-131): priorities -> tree, reference order
  a->launches++;
  for (int k = 0; k < 2; ++k) {
    TrunkWs& t = a->ws[1 + k];
    k_td3_critic_seed<<<1, 256, 0, a->st>>>(a->q[k], a->y, a->is_wsum, B, a->dq, a->stats + k);
    LAUNCH_OK(a);
    k_outer<<<grid_for((size_t)B * a->feat), 256, 0, a->st>>>(a->dq, PC[k].h_w, B, a->feat, a->dfeat);
    LAUNCH_OK(a);
    if ((rc = colsum(a, t.feat, a->feat, a->dq, B, a->feat, GC[k].h_w))) return rc;
    if ((rc = colsum(a, a->dq, 1, nullptr, B, 1, GC[k].h_b))) return rc;
    k_side_bwd<<<B, 128, 0, a->st>>>(a->dfeat + a->flat, a->feat, t.o16, t.h1, PC[k].fb_w, PC[k].fa_w, a->K1C, SKY, A, a->d16, a->dh1, nullptr, 0, a->side, 1);
    LAUNCH_OK(a);
    if ((rc = side_wgrads(a, t.x1, t.h1, a->K1C, GC[k]))) return rc;
    if ((rc = trunk_backward(a, t, PC[k], GC[k], a->mb_img, ACT_RELU))) return rc;
  }
  for (int k = 0; k < 2; ++k)
    if ((rc = adam_step(a, 1 + k))) return rc;
  if (!actor_step) return SC_OK;

  // ---- 4. delayed actor update (:529-600): one Adam step on -mean(q1(s, pi(s)) w), or the five-step adaptive-ADMM loop with hints.
  //         critic_1 stays in train() mode: every one of its forwards moves its BatchNorm running statistics; its weights do not
  //         change inside the loop, so the trunk runs once and the statistics take n_it momentum updates ----
  const int n_it = hint ? 5 : 1;
  if (hint) { CKC(launch_admm_update(a->admm_dev, 1, a->st)); a->launches++; }        // lambda <- 0, rho <- admm_rho (:543-546)
  if ((rc = trunk_forward(a, a->ws[0], PC[0], 1, a->mb_img, ACT_RELU, 1, B, n_it))) return rc;
  for (int it = 0; it < n_it; ++it) {
    TrunkWs& t = a->ws[3];
    if ((rc = trunk_forward(a, t, PA, 0, a->mb_img, ACT_ELU, 1, B))) return rc;
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky, SKY, nullptr, 0, PA.fa_w, PA.fa_b, PA.fb_w, PA.fb_b, t.x1, t.h1, t.o16, t.feat, a->side, a->feat, a->flat);
    LAUNCH_OK(a);
    if ((rc = sgemm(a, t.feat, a->feat, 1, PA.h_w, 1, a->feat, a->x21, HID, B, HID, a->feat, PA.h_b, ACT_ELU))) return rc;
    k_td3_head<<<B, 128, 0, a->st>>>(a->x21, PA.mu_w, PA.mu_b, A, nullptr, 0, nullptr, a->pi);
    LAUNCH_OK(a);
    TrunkWs& tc = a->ws[0];
    k_side_fwd<<<B, 128, 0, a->st>>>(a->mb_sky, SKY, a->pi, A, PC[0].fa_w, PC[0].fa_b, PC[0].fb_w, PC[0].fb_b, a->pside[0].x1, a->pside[0].h1, a->pside[0].o16,
                                     tc.feat, a->side, a->feat, a->flat, 1);
    LAUNCH_OK(a);
    k_critic_head<<<B, 256, 0, a->st>>>(tc.feat, PC[0].h_w, PC[0].h_b, a->qpi[0], a->feat);
    LAUNCH_OK(a);
    k_td3_actor_seed<<<1, 256, 0, a->st>>>(a->qpi[0], PC[0].h_w, a->is_wsum, B, a->ds[0], a->side, a->flat, a->stats + 2);
    LAUNCH_OK(a);
    k_side_bwd<<<B, 128, 0, a->st>>>(a->ds[0], a->side, a->pside[0].o16, a->pside[0].h1, PC[0].fb_w, PC[0].fa_w, a->K1C, SKY, A, a->d16, a->dh1, a->dact, 0, a->side, 1);
    LAUNCH_OK(a);
    if (hint) {
      k_td3_hint<<<1, 256, 0, a->st>>>(a->pi, a->mb_hint, a->lambda, a->admm_rho_dev, a->is_wsum, B, B * A, a->dact, a->stats + 2);
      LAUNCH_OK(a);
    }
    k_td3_head_bwd<<<B, 128, 0, a->st>>>(a->dact, a->pi, A, PA.mu_w, a->x21, a->dmu, a->dpre);
    LAUNCH_OK(a);
    if ((rc = sgemm(a, a->dmu, 1, A, a->x21, HID, 1, GA.mu_w, HID, A, HID, B, nullptr, -1))) return rc;
    if ((rc = colsum(a, a->dmu, A, nullptr, B, A, GA.mu_b))) return rc;
    if ((rc = sgemm(a, a->dpre, 1, HID, t.feat, a->feat, 1, GA.h_w, a->feat, HID, a->feat, B, nullptr, -1))) return rc;
    if ((rc = colsum(a, a->dpre, HID, nullptr, B, HID, GA.h_b))) return rc;
    if ((rc = sgemm(a, a->dpre, HID, 1, PA.h_w, a->feat, 1, a->dfeat, a->feat, B, a->feat, HID, nullptr, -1))) return rc;
    k_side_bwd<<<B, 128, 0, a->st>>>(a->dfeat + a->flat, a->feat, t.o16, t.h1, PA.fb_w, PA.fa_w, SKY, SKY, 0, a->d16, a->dh1, nullptr, 0, a->side);
    LAUNCH_OK(a);
    if ((rc = side_wgrads(a, t.x1, t.h1, SKY, GA))) return rc;
    if ((rc = trunk_backward(a, t, PA, GA, a->mb_img, ACT_ELU))) return rc;
    if ((rc = adam_step(a, 0, &a->step_dev->adam_x[it]))) return rc;
    if (hint) { CKC(launch_admm_update(a->admm_dev + 1 + it, 1, a->st)); a->launches++; }   // multipliers, adaptive penalty (:564-597)
  }
  return sc_calib_soft_update(a, (float)a->cfg.tau);                                       // :603, actor steps only
}

int sc_calib_learn(sc_calib* a, const sc_noise* noise, sc_learn_stats* stats) {
  if (!a) return failc(SC_ERR_ARG, "null handle");
  if (stats) memset(stats, 0, sizeof(*stats));
  if (a->mem_cntr < a->B) return SC_OK;                                     // calib_sac.py:301-302
  if (!a->td3 && !a->batch_ready) return failc(SC_ERR_STATE, "no minibatch selected: call sc_calib_set_batch_host first");
  const int B = a->B, A = a->A;
  int rc;
  const bool hint = a->cfg.use_hint != 0;
  // TD3: `dual` is "the delayed actor update runs in this call" (demix_td3.py:527-529) -- the second launch sequence / graph
  const int dual = a->td3 ? (((a->learn_counter + 1) % a->interval) == 0) : (hint && (a->learn_counter % 10 == 0));
  const bool host_noise = a->td3 ? (noise && noise->per_u01_host) : (noise && noise->eps_next_host && noise->eps_cur_host);
  if (a->td3 && host_noise && !noise->has_td3_target_noise) return failc(SC_ERR_ARG, "per_u01_host given but td3_target_noise missing");
  const int n_it = a->td3 && dual ? (hint ? 5 : 1) : 0;
  // ---- step scalars: Adam bias corrections (t counts steps per optimizer) and the Philox streams of this step's draws ----
  {
    const int k = a->step_k;
    a->step_k ^= 1;
    CKC(cudaEventSynchronize(a->ev_step[k]));
    sc_calib::StepScalars& h = a->step_host[k];
    a->t_critic++;
    if (!a->td3) a->t_actor++;
    h.adam[0] = adam_hyper_c(a->cfg.lr_actor, a->t_actor);
    h.adam[1] = h.adam[2] = adam_hyper_c(a->cfg.lr_critic, a->t_critic);
    for (int it = 0; it < 5; ++it) h.adam_x[it] = adam_hyper_c(a->cfg.lr_actor, a->t_actor + 1 + (it < n_it ? it : 0));
    a->t_actor += n_it;
    if (a->td3) {
      a->per_beta = a->per_beta + 1e-4 < 1.0 ? a->per_beta + 1e-4 : 1.0;          // PER.sample_buffer :95 (before the draws)
      PerSampleP& ps = h.per_s;
      ps.tree = a->tree; ps.C = a->C; ps.B = B;
      ps.u01 = host_noise ? a->u01_dev : nullptr;
      ps.seed = a->cfg.rng_seed; ps.ctr = 0xE000000000ull + (unsigned long long)a->noise_ctr;
      ps.beta = a->per_beta;
      ps.tree_idx = a->tree_idx; ps.data_idx = a->bidx; ps.is_w = a->is_w; ps.is_wsum = a->is_wsum;
      PerUpdateP& pu = h.per_u;
      pu.tree = a->tree; pu.C = a->C; pu.B = B; pu.idx = a->tree_idx; pu.err = a->per_err; pu.clip = 100.f;     // demix_td3.py:34
      pu.ps_out = a->per_ps; pu.ps_in = nullptr;
    }
    h.noise[0] = 2 * (unsigned long long)a->noise_ctr;
    h.noise[1] = 2 * (unsigned long long)a->noise_ctr + 1;
    h.noise[2] = 0xD000000000ull + (unsigned long long)a->noise_ctr;
    h.noise[3] = 0xF000000000ull + (unsigned long long)a->noise_ctr;          // TD3 target-noise scalar
    if (!host_noise) a->noise_ctr++;
    CKC(cudaMemcpyAsync(a->step_dev, &h, sizeof(h), cudaMemcpyHostToDevice, a->st));
    CKC(cudaEventRecord(a->ev_step[k], a->st));
  }
  if (host_noise && a->td3) {
    // the caller's numpy stream: B uniforms of PER.sample_buffer (:100) and the scalar of :493, clamped in float64 like the reference
    CKC(cudaStreamSynchronize(a->st));
    memcpy(a->h_u01, noise->per_u01_host, (size_t)B * sizeof(double));
    double n = noise->td3_target_noise;
    n = n < -0.5 ? -0.5 : (n > 0.5 ? 0.5 : n);
    float* nf = reinterpret_cast<float*>(a->h_u01 + B);
    nf[0] = (float)n;
    CKC(cudaMemcpyAsync(a->u01_dev, a->h_u01, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, a->st));
    CKC(cudaMemcpyAsync(a->tnoise, nf, sizeof(float), cudaMemcpyHostToDevice, a->st));
  }
  if (host_noise && !a->td3) {
    CKC(cudaStreamSynchronize(a->st));
    float* h = a->h_stage;
    memcpy(h, noise->eps_next_host, (size_t)B * A * 4);
    memcpy(h + (size_t)B * A, noise->eps_cur_host, (size_t)B * A * 4);
    CKC(cudaMemcpyAsync(a->eps_next, h, (size_t)B * A * 4, cudaMemcpyHostToDevice, a->st));
    CKC(cudaMemcpyAsync(a->eps_cur, h + (size_t)B * A, (size_t)B * A * 4, cudaMemcpyHostToDevice, a->st));
    if (dual && a->kld_mode == 1) {
      if (!noise->eps_dual_host) return failc(SC_ERR_ARG, "host noise given but eps_dual_host missing on a dual-update step");
      CKC(cudaStreamSynchronize(a->st));
      memcpy(h, noise->eps_dual_host, (size_t)B * A * 4);
      CKC(cudaMemcpyAsync(a->eps_dual, h, (size_t)B * A * 4, cudaMemcpyHostToDevice, a->st));
    }
  }
  // ---- the launches: eager the first time a shape is seen, captured the second time, replayed afterwards ----
  bool done = false;
  if (a->use_graph && !host_noise) {
    if (a->graph[dual] == nullptr && a->graph_seen[dual] >= 1) {
      cudaStream_t user = a->st;
      const int64_t launches0 = a->launches;
      a->st = a->st_cap;
      cudaGraph_t g = nullptr;
      cudaError_t e = cudaStreamBeginCapture(a->st_cap, cudaStreamCaptureModeRelaxed);
      int brc = SC_OK;
      if (e == cudaSuccess) {
        brc = a->td3 ? td3_body(a, host_noise, hint, dual) : calib_body(a, host_noise, hint, dual);
        cudaError_t e2 = cudaStreamEndCapture(a->st_cap, &g);
        if (e == cudaSuccess) e = e2;
      }
      a->st = user;
      a->launches = launches0;
      if (e == cudaSuccess && brc == SC_OK) e = cudaGraphInstantiate(&a->graph[dual], g, 0);
      if (g) cudaGraphDestroy(g);
      if (e != cudaSuccess || brc != SC_OK) { cudaGetLastError(); a->graph[dual] = nullptr; a->use_graph = 0; }
    }
    a->graph_seen[dual]++;
    if (a->graph[dual] != nullptr) {
      CKC(cudaGraphLaunch(a->graph[dual], a->st));
      a->launches += a->graph_launches[dual];
      done = true;
    }
  }
  if (!done) {
    const int64_t l0 = a->launches;
    if ((rc = a->td3 ? td3_body(a, host_noise, hint, dual) : calib_body(a, host_noise, hint, dual))) return rc;
    a->graph_launches[dual] = a->launches - l0;
  }
  a->learn_counter++;
  a->batch_ready = false;
  if (stats) {
    CKC(cudaMemcpyAsync(a->h_stats, a->stats, 8 * sizeof(float), cudaMemcpyDeviceToHost, a->st));
    CKC(cudaMemcpyAsync(a->h_stats + 8, a->hint_out, 4 * sizeof(float), cudaMemcpyDeviceToHost, a->st));
    CKC(cudaMemcpyAsync(a->h_stats + 12, a->td3 ? a->admm_rho_dev : a->rho, sizeof(float), cudaMemcpyDeviceToHost, a->st));
    CKC(cudaStreamSynchronize(a->st));
    stats->critic_1_loss = a->h_stats[0]; stats->critic_2_loss = a->h_stats[1]; stats->actor_loss = a->h_stats[2];
    stats->logp_mean = a->h_stats[3];
    stats->gfun = hint ? a->h_stats[9] : 0.f;
    stats->rho = a->h_stats[12];
    stats->learn_counter = (int32_t)a->learn_counter;
    stats->did_learn = 1;
    stats->actor_step = a->td3 ? dual : 1;
    if (a->td3) { stats->gfun = 0.f; stats->logp_mean = 0.f; if (!dual) stats->actor_loss = 0.f; }
  }
  return SC_OK;
}

}  // extern "C"
