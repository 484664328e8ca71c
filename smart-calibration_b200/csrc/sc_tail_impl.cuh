// Row-local "tail" of a batched SAC / TD3 phase as its own kernel: everything between the two wide tcgen05 forward layers
// (sc_dense_tc.cu) and the wide input-gradient GEMM (k_dgrad_tc) that only touches the narrow layers:
//   PROG_CRITIC: policy head on s' (actor fc3 -> LN -> ELU -> mu / log sigma -> tanh-Gaussian sample + log-prob), action branch
//                (fc21 -> LN -> ELU -> fc22 -> LN -> ELU) and fc3 -> Q of both target critics and both online critics, TD target,
//                TD error, PER priority, and the backward pass of the online critics down to dz12 / dz22 / dz21;
//   PROG_ACTOR:  policy on s (reparameterised), Q_k(s, pi(s)) with the updated critics, actor loss gradient through the critics'
//                action branch and the policy head down to dz3 / dz2 of the actor.
// Reference bodies: enet_sac.py:569-585,589-598 / enet_td3.py:247-288,303-308 (same arithmetic as the row-local stage of
// sc_rowphase_impl.cuh, which remains the choose_action path and the SMARTCAL_TAIL=row A/B switch).
//
// Geometry: one CTA carries TR batch rows (4, or 2 for a stand-alone agent: sc_tail.cu) of ONE agent (blockIdx.y) through the whole chain; activations stay in shared
// memory.  Every weight matrix of the chain is streamed from L2 exactly once per CTA: a producer warp walks the phase's flat
// chunk schedule (16 KB of whole weight rows per chunk) and keeps a 5-slot shared-memory ring full with 1-D bulk copies
// (cp.async.bulk, completion on mbarrier transaction counts), ACROSS stage boundaries -- while the consumers normalise one
// layer the next matrices are already landing -- so the chain is no longer a sequence of dependent L2 round trips (the row-phase
// tail fetched each weight row from L2 inside the chain: 33 us per phase at one row per CTA).
//   forward chunk  Z[r][n] = sum_k W[n][k] x[r][k]:  a warp takes whole features; lane = 4 (8) consecutive k with its x quads of
//                  all TR rows held in registers for the stage; W quads are read conflict-free along the row; butterfly reduce;
//   transposed chunk  dX[r][k] = sum_n d[r][n] W[n][k]:  lane = 4 (8) consecutive k, a warp takes rows n = warp, warp+8, ...;
//                  accumulators stay in registers across the chunks of a matrix and are combined across warps once.
// FP32 FFMA throughout (these layers are 4 % of the step's FLOPs; their GEMMs have 4 rows per CTA -- no tensor-core shape).
#include "sc_common.cuh"
#include "sc_params.h"

#if !defined(SC_TAIL_TR) || !defined(SC_TAIL_NS)
#error "include from sc_tail.cu / sc_tail_tr2.cu, which define SC_TAIL_TR (batch rows per CTA) and SC_TAIL_NS (variant namespace)"
#endif

namespace sc {
namespace SC_TAIL_NS {

constexpr int TR = SC_TAIL_TR;               // batch rows per CTA: 4 when agents share the GPU, 2 for a stand-alone agent (sc_tail.cu)
constexpr int TL_THREADS = 512;             // consumer threads: the chain is latency-bound, 16 warps hide twice what 8 did
constexpr int TL_ALL = TL_THREADS + 32;     // + the producer warp
constexpr int TL_WARPS = 16;
constexpr int RED_SLOTS = 8;                // cross-warp partial slots of the transposed product (warps w and w + 8 share slot w)
constexpr int CH_FLOATS = 4096;             // floats per weight chunk (16 KB): 16 rows of 256 or 32 rows of 128
constexpr int NSLOT = 3;                    // ring depth
constexpr int PSTAGE_FLOATS = 9600;         // staged small parameters (biases, LayerNorm affine, fc21, fc3, heads) of the actor + 4 critics
constexpr int MAXCH = 32;                   // chunks per phase (critic 20, actor 24)
constexpr int NETS4 = 4;

struct TailS {
  float *xh2, *zh2;              // [TR][H2]   actor layer-2 rows (+ normalised, actor phase)
  float *x3, *zh3;               // [TR][H3]
  float *head;                   // [TR][2*AMAX] mu | raw logsigma (later d mu | d logsigma)
  float *an;                     // [2][TR][AMAX] actions fed to the critics: [0] policy sample, [1] replay action
  float *smp;                    // [TR][AMAX][4] tanh(u), sigma, eps, clamp-pass mask
  float *lp;                     // [TR][AMAX]
  float *sc;                     // [TR][16] per-row scalars
  float *rsx;                    // [TR][8]: [2] rs3, [3] rs2, [4+k] rs12 of online critic k
  float *h12;                    // [NETS4][TR][H2]
  float *zh12;                   // [2][TR][H2]
  float *v21, *zh21;             // [NETS4][TR][HA1]
  float *v22, *zh22;             // [NETS4][TR][HA2]
  float *rsn;                    // [NETS4][TR][2]
  float *qv;                     // [NETS4][TR]
  float *ga, *gb;                // [2][TR][H2]
  float *gc, *gd;                // [2][TR][HA1]  (actor phase: [TR][H2])
  float *ge, *gf;                // [2][TR][HA2]
  float *red;                    // [RED_SLOTS][TR][H2] cross-warp partials
  float *wbuf;                   // [NSLOT][CH_FLOATS] weight chunk ring
  float *pstage;                 // [PSTAGE_FLOATS] small parameters
};
constexpr int TAIL_SMEM_FLOATS = 2 * TR * H2 + 2 * TR * H3 + TR * 2 * AMAX + 2 * TR * AMAX + TR * AMAX * 4 + TR * AMAX + TR * 16 + TR * 8 +
                                 NETS4 * TR * H2 + 2 * TR * H2 + 2 * NETS4 * TR * HA1 + 2 * NETS4 * TR * HA2 + NETS4 * TR * 2 + NETS4 * TR +
                                 4 * TR * H2 + 4 * TR * HA1 + 4 * TR * HA2 + RED_SLOTS * TR * H2 + NSLOT * CH_FLOATS + PSTAGE_FLOATS;
__device__ TailS tail_carve(float* sm) {
  TailS t;
  t.xh2 = sm; sm += TR * H2;
  t.zh2 = sm; sm += TR * H2;
  t.x3 = sm; sm += TR * H3;
  t.zh3 = sm; sm += TR * H3;
  t.head = sm; sm += TR * 2 * AMAX;
  t.an = sm; sm += 2 * TR * AMAX;
  t.smp = sm; sm += TR * AMAX * 4;
  t.lp = sm; sm += TR * AMAX;
  t.sc = sm; sm += TR * 16;
  t.rsx = sm; sm += TR * 8;
  t.h12 = sm; sm += NETS4 * TR * H2;
  t.zh12 = sm; sm += 2 * TR * H2;
  t.v21 = sm; sm += NETS4 * TR * HA1;
  t.zh21 = sm; sm += NETS4 * TR * HA1;
  t.v22 = sm; sm += NETS4 * TR * HA2;
  t.zh22 = sm; sm += NETS4 * TR * HA2;
  t.rsn = sm; sm += NETS4 * TR * 2;
  t.qv = sm; sm += NETS4 * TR;
  t.ga = sm; sm += 2 * TR * H2;
  t.gb = sm; sm += 2 * TR * H2;
  t.gc = sm; sm += 2 * TR * HA1;
  t.gd = sm; sm += 2 * TR * HA1;
  t.ge = sm; sm += 2 * TR * HA2;
  t.gf = sm; sm += 2 * TR * HA2;
  t.red = sm; sm += RED_SLOTS * TR * H2;
  t.wbuf = sm; sm += NSLOT * CH_FLOATS;
  t.pstage = sm; sm += PSTAGE_FLOATS;
  return t;
}

enum { TSC_LOGP = 0, TSC_DQ1 = 6, TSC_DQ2 = 7, TSC_GLP = 8, TSC_REWARD = 9, TSC_DONE = 10 };

// network parameters are read through generic loads: the small tensors are staged in shared memory at kernel start (below)
__device__ __forceinline__ float ldp(const float* p) { return *p; }
__device__ __forceinline__ float4 ldp4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ unsigned long long* g_tail_ts = nullptr;     // debug timeline target of t_mark_g (set per kernel by thread 0)
__device__ __forceinline__ void t_mark_g(int i) {
  if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 64 && g_tail_ts != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_tail_ts[(size_t)(2400 + 2 * blockIdx.x) * 8 + i] = t;
  }
}
// barrier over the consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void tsync() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }
static_assert(TL_THREADS == 512, "tsync() is written for 512 consumer threads");

// ---- weight chunk stream ---------------------------------------------------------------------------------------------------
// A phase's weight traffic is a flat list of chunks (whole rows of one matrix, <= 16 KB, contiguous in global memory because
// these matrices have pitch == K).  Producer and consumers walk it in the same order; `cur` counts chunks consumed so far.
struct Chunk { const float* src; int rows, K; };
struct PCopy { const float* src; float* dst; unsigned bytes; };
struct Stream {
  Chunk* sched;                  // [MAXCH] in shared memory
  int total;
  uint64_t *full, *empty;        // [NSLOT] mbarriers
  float* ring;
  uint64_t* pbar;                // completion of the small-parameter copies
};
__device__ __forceinline__ int sched_add(Chunk* sched, int n, const float* W, int N, int K) {
  const int per = CH_FLOATS / K;
  for (int n0 = 0; n0 < N; n0 += per) sched[n++] = Chunk{W + (size_t)n0 * K, min(per, N - n0), K};
  return n;
}
// consumer side: wait for chunk `c` / release it
__device__ __forceinline__ const float* chunk_acquire(const Stream& st, int c) {
  mbar_wait(st.full + (c % NSLOT), (unsigned)(c / NSLOT) & 1u);
  return st.ring + (size_t)(c % NSLOT) * CH_FLOATS;
}
__device__ __forceinline__ void chunk_release(const Stream& st, int c) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(st.empty + (c % NSLOT));
}

struct LinJob {
  int N;                         // rows of W (output features of the forward product)
  const float* bias;             // forward only (nullable)
  const float* x;                // forward: input rows [TR][K] in shared memory; transposed: d rows [TR][N]
  float* out;                    // forward: [TR][N] in shared memory
};

// forward: out[r][n] = bias[n] + sum_k W[n][k] x[r][k] for every job; the chunks of the jobs are the next ones of the stream
template <int K>
__device__ void linear_fwd(const LinJob* jobs, int njobs, const Stream& st, int& cur) {
  constexpr int KV = K / 128;                         // quads of a lane
  constexpr int PER = CH_FLOATS / K;                  // rows per chunk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = 0; j < njobs; ++j) {
    const LinJob jb = jobs[j];
    float4 xq[TR][KV];
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int v = 0; v < KV; ++v) xq[r][v] = *reinterpret_cast<const float4*>(jb.x + r * K + 4 * lane + 128 * v);
    for (int n0 = 0; n0 < jb.N; n0 += PER, ++cur) {
      const float* wb = chunk_acquire(st, cur);
      const int rows = min(PER, jb.N - n0);
      for (int f = warp; f < rows; f += TL_WARPS) {
        float acc[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) acc[r] = 0.f;
#pragma unroll
        for (int v = 0; v < KV; ++v) {
          const float4 wv = *reinterpret_cast<const float4*>(wb + f * K + 4 * lane + 128 * v);
#pragma unroll
          for (int r = 0; r < TR; ++r) acc[r] = dot4(wv, xq[r][v], acc[r]);
        }
        // transposing butterfly: the TR row sums over 32 lanes; every row is summed over the lanes in the same order
        // (xor 16, 8, 4, 2, 1) whatever TR is, so a row's result does not depend on how many rows share the CTA
        static_assert(TR == 4 || TR == 2, "reduction below is written for 2 or 4 rows");
        const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
        float v;
        if constexpr (TR == 4) {       // 6 shuffles; lanes 0 / 8 / 16 / 24 end with rows 0 / 1 / 2 / 3
          float k0 = hi16 ? acc[2] : acc[0], k1 = hi16 ? acc[3] : acc[1];
          k0 += __shfl_xor_sync(0xffffffffu, hi16 ? acc[0] : acc[2], 16);
          k1 += __shfl_xor_sync(0xffffffffu, hi16 ? acc[1] : acc[3], 16);
          v = hi8 ? k1 : k0;
          v += __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
        } else {                       // 5 shuffles; lanes 0 / 16 end with rows 0 / 1
          v = hi16 ? acc[1] : acc[0];
          v += __shfl_xor_sync(0xffffffffu, hi16 ? acc[0] : acc[1], 16);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
        }
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if ((lane & (32 / TR - 1)) == 0) {
          const int r = lane / (32 / TR);
          if (jb.bias != nullptr) v += ldp(jb.bias + n0 + f);
          jb.out[r * jb.N + n0 + f] = v;
        }
      }
      chunk_release(st, cur);
    }
  }
  tsync();
}

// transposed: out[r][k] = sum_n d[r][n] W[n][k] for ONE job (N rows, K columns); result left in out ([TR][K], shared memory)
template <int K>
__device__ void linear_bwd(const LinJob& jb, float* out, const Stream& st, int& cur, const TailS& s) {
  constexpr int KV = K / 128;                         // float4 quads per lane
  constexpr int PER = CH_FLOATS / K;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float4 acc[TR][KV];
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int v = 0; v < KV; ++v) acc[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int n0 = 0; n0 < jb.N; n0 += PER, ++cur) {
    const float* wb = chunk_acquire(st, cur);
    const int rows = min(PER, jb.N - n0);
    for (int n = warp; n < rows; n += TL_WARPS) {
      float4 wv[KV];
#pragma unroll
      for (int v = 0; v < KV; ++v) wv[v] = *reinterpret_cast<const float4*>(wb + n * K + 4 * lane + 128 * v);
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const float d = jb.x[r * jb.N + n0 + n];
#pragma unroll
        for (int v = 0; v < KV; ++v) {
          acc[r][v].x = fmaf(d, wv[v].x, acc[r][v].x);
          acc[r][v].y = fmaf(d, wv[v].y, acc[r][v].y);
          acc[r][v].z = fmaf(d, wv[v].z, acc[r][v].z);
          acc[r][v].w = fmaf(d, wv[v].w, acc[r][v].w);
        }
      }
    }
    chunk_release(st, cur);
  }
  // cross-warp combination in fixed order: slot w <- warp w, then += warp w + 8, then slots 0..7 are summed
  if (warp < RED_SLOTS) {
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int v = 0; v < KV; ++v) *reinterpret_cast<float4*>(s.red + (warp * TR + r) * K + 4 * lane + 128 * v) = acc[r][v];
  }
  tsync();
  if (warp >= RED_SLOTS) {
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int v = 0; v < KV; ++v) {
        float4* q = reinterpret_cast<float4*>(s.red + ((warp - RED_SLOTS) * TR + r) * K + 4 * lane + 128 * v);
        float4 t = *q;
        t.x += acc[r][v].x; t.y += acc[r][v].y; t.z += acc[r][v].z; t.w += acc[r][v].w;
        *q = t;
      }
  }
  tsync();
  for (int e = tid; e < TR * K; e += TL_THREADS) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < RED_SLOTS; ++w) v += s.red[w * TR * K + e];
    out[e] = v;
  }
  tsync();
}

// ---- row helpers (rows indexed rix = net*TR + t; warps stride over them) -----------------------------------------------------
__device__ void t_ln_elu(int nn, float* z, float* zh, int N, const float* const* gamma, const float* const* beta, float* rs, int rs_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rix = warp; rix < nn * TR; rix += TL_WARPS) {
    const int net = rix / TR;
    float* zr = z + rix * N;
    float sm = 0.f;
    for (int j = lane; j < N; j += 32) sm += zr[j];
    const float mean = warp_sum(sm) / (float)N;
    float m2 = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float d = zr[j] - mean;
      m2 = fmaf(d, d, m2);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(m2) / (float)N + SC_LN_EPS);
    const float* g = gamma[net];
    const float* b = beta[net];
    for (int j = lane; j < N; j += 32) {
      const float n = (zr[j] - mean) * rstd;
      zh[rix * N + j] = n;
      zr[j] = elu(fmaf(n, ldp(g + j), ldp(b + j)));
    }
    if (lane == 0) rs[rix * rs_stride] = rstd;
  }
}
// dh -> du (in place), dz
__device__ void t_ln_bwd(int nn, float* dh, const float* h, const float* zh, int N, const float* const* gamma, const float* rs, int rs_stride,
                         float* dz) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rix = warp; rix < nn * TR; rix += TL_WARPS) {
    const float* g = gamma[rix / TR];
    const int o = rix * N;
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float du = dh[o + j] * elu_grad_from_out(h[o + j]);
      dh[o + j] = du;
      const float t = du * ldp(g + j);
      s1 += t;
      s2 = fmaf(t, zh[o + j], s2);
    }
    const float c1 = warp_sum(s1) / (float)N, c2 = warp_sum(s2) / (float)N;
    const float r = rs[rix * rs_stride];
    for (int j = lane; j < N; j += 32) dz[o + j] = r * (dh[o + j] * ldp(g + j) - c1 - zh[o + j] * c2);
  }
}
struct TRows { int row[TR]; };
__device__ __forceinline__ void t_load_rows(float* dst, const float* src, int N, const TRows& rs) {   // dst [TR][N]
  for (int e = threadIdx.x; e < TR * (N >> 2); e += TL_THREADS) {
    const int t = e / (N >> 2), c4 = e - t * (N >> 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rs.row[t] >= 0) v = ld_act4(src + (size_t)rs.row[t] * N + 4 * c4);
    *reinterpret_cast<float4*>(dst + t * N + 4 * c4) = v;
  }
}
__device__ __forceinline__ void t_store_rows(float* dst, const float* src, int N, const TRows& rs) {   // src [TR][N]
  if (dst == nullptr) return;
  for (int e = threadIdx.x; e < TR * (N >> 2); e += TL_THREADS) {
    const int t = e / (N >> 2), c4 = e - t * (N >> 2);
    if (rs.row[t] >= 0)
      *reinterpret_cast<float4*>(dst + (size_t)rs.row[t] * N + 4 * c4) = *reinterpret_cast<const float4*>(src + t * N + 4 * c4);
  }
}
// the same rows as tf32 big / small parts (hl = [big | small], `plane` floats apart): operands of the tcgen05 dgrad / wgrad GEMMs
__device__ __forceinline__ void t_store_rows_hl(float* hl, size_t plane, const float* src, int N, const TRows& rs) {
  if (hl == nullptr) return;
  for (int e = threadIdx.x; e < TR * (N >> 2); e += TL_THREADS) {
    const int t = e / (N >> 2), c4 = e - t * (N >> 2);
    if (rs.row[t] >= 0) {
      float4 big, small;
      split_rn4(*reinterpret_cast<const float4*>(src + t * N + 4 * c4), big, small);
      float* d = hl + (size_t)rs.row[t] * N + 4 * c4;
      *reinterpret_cast<float4*>(d) = big;
      *reinterpret_cast<float4*>(d + plane) = small;
    }
  }
}

// action branches (fc21 -> LN -> ELU -> fc22 -> LN -> ELU) of `nn` critics at once; net i reads actions an_sel[i] ([TR][AMAX]).
// Leaves h21/zh21/h22/zh22 in s.v21.. and 1/std in s.rsn[(net*TR+t)*2 + {0,1}].
__device__ void t_action_branches(int nn, const CriticP* const* c, const float* const* an_sel, int A, int Ap, const TailS& s, const Stream& st,
                                  int& cur) {
  for (int e = threadIdx.x; e < nn * TR * HA1; e += TL_THREADS) {
    const int rix = e / HA1, n = e - rix * HA1, net = rix / TR, tt = rix - net * TR;
    const float* w = c[net]->w21 + n * Ap;
    const float* a = an_sel[net] + tt * AMAX;
    float acc = 0.f;
    for (int j = 0; j < A; ++j) acc = fmaf(ldp(w + j), a[j], acc);
    s.v21[e] = acc + ldp(c[net]->b21 + n);
  }
  tsync();
  if (nn == NETS4) t_mark_g(10);
  const float *g21[NETS4], *b21[NETS4], *g22[NETS4], *b22[NETS4];
  LinJob jobs[NETS4];
  for (int i = 0; i < nn; ++i) {
    g21[i] = c[i]->g21; b21[i] = c[i]->be21; g22[i] = c[i]->g22; b22[i] = c[i]->be22;
    jobs[i] = LinJob{HA2, c[i]->b22, s.v21 + i * TR * HA1, s.v22 + i * TR * HA2};
  }
  t_ln_elu(nn, s.v21, s.zh21, HA1, g21, b21, s.rsn, 2);
  tsync();
  if (nn == NETS4) t_mark_g(11);
  linear_fwd<HA1>(jobs, nn, st, cur);
  if (nn == NETS4) t_mark_g(12);
  t_ln_elu(nn, s.v22, s.zh22, HA2, g22, b22, s.rsn + 1, 2);
  tsync();
}
// q[net][t] = b3 + w3[0:H2].h12[net][t] + w3[H2:].h22[net][t]
__device__ void t_critic_q(int nn, const CriticP* const* c, const TailS& s) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rix = warp; rix < nn * TR; rix += TL_WARPS) {
    const CriticP* cc = c[rix / TR];
    float v = 0.f;
    for (int k = lane; k < H2; k += 32) v = fmaf(ldp(cc->w3 + k), s.h12[rix * H2 + k], v);
    for (int k = lane; k < HA2; k += 32) v = fmaf(ldp(cc->w3 + H2 + k), s.v22[rix * HA2 + k], v);
    v = warp_sum(v);
    if (lane == 0) s.qv[rix] = v + ldp(cc->b3);
  }
}
// actor layer 3 + heads on the rows staged in s.xh2: h3 / zh3 in s.x3 / s.zh3, raw head outputs in s.head, 1/std in s.rsx[t*8+2]
__device__ void t_actor_forward(const RowPhaseParams& p, const TailS& s, const Stream& st, int& cur) {
  const ActorP& a = p.actor;
  const float *g3[1] = {a.g3}, *be3[1] = {a.be3};
  LinJob j3{H3, a.b3, s.xh2, s.x3};
  linear_fwd<H2>(&j3, 1, st, cur);
  t_ln_elu(1, s.x3, s.zh3, H3, g3, be3, s.rsx + 2, 8);
  tsync();
  // heads: (mu | logsigma) rows of width A each -> s.head[t][0:A] and [AMAX:AMAX+A]; one warp per output, weights from L2 (<= 2A rows)
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = a.wls != nullptr ? 2 * p.A : p.A;
    for (int o = warp; o < nh; o += TL_WARPS) {
      const bool ls = o >= p.A;
      const int j = ls ? o - p.A : o;
      const float4 wv = ldp4((ls ? a.wls : a.wmu) + (size_t)j * H3 + 4 * lane);        // H3 == 128 == 32 lanes x float4
      float acc[TR];
#pragma unroll
      for (int tt = 0; tt < TR; ++tt) acc[tt] = warp_sum(dot4(wv, *reinterpret_cast<const float4*>(s.x3 + tt * H3 + 4 * lane), 0.f));
      if (lane == 0) {
        const float b = ldp((ls ? a.bls : a.bmu) + j);
#pragma unroll
        for (int tt = 0; tt < TR; ++tt) s.head[tt * 2 * AMAX + (ls ? AMAX : 0) + j] = acc[tt] + b;
      }
    }
  }
  tsync();
}
__device__ __forceinline__ float t_eps_value(const RowPhaseParams& p, int row, int j) {
  const size_t e = (size_t)row * p.A + j;
  if (!p.eps_inline) {
    const float v = ld_act(p.eps + e);
    if (p.eps_store != nullptr && p.eps_store != p.eps) p.eps_store[e] = v;
    return v;
  }
  const float2 z = philox_normal2(p.eps_seed, p.eps_stream, (uint64_t)(e >> 1));
  const float v = (e & 1) ? z.y : z.x;
  if (p.eps_store != nullptr) p.eps_store[e] = v;
  return v;
}
// tanh-squashed Gaussian sample + log-prob (SAC, ref ActorNetwork.sample_normal enet_sac.py:446-466) or the deterministic TD3 head
__device__ void t_policy_head(const RowPhaseParams& p, const TailS& s, const TRows& rs, bool td3_target) {
  const int tid = threadIdx.x;
  if (tid < TR * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    float* smp = s.smp + (r * AMAX + j) * 4;
    if (p.kind == 0) {
      const float mu = s.head[r * 2 * AMAX + j];
      const float raw = s.head[r * 2 * AMAX + AMAX + j];
      const float ls = fminf(fmaxf(raw, -20.f), 2.f);
      const float sg = expf(ls);
      const float e = rs.row[r] >= 0 ? t_eps_value(p, rs.row[r], j) : 0.f;
      const float u = __fadd_rn(mu, __fmul_rn(e, sg));                 // loc + eps*scale, two roundings like torch
      const float th = tanhf(u);
      const float dm = u - mu;
      // Normal.log_prob: -((v-loc)^2)/(2 var) - log(scale) - log(sqrt(2 pi)); then the tanh correction
      float lp = -(dm * dm) / (2.f * (sg * sg)) - logf(sg) - 0.91893853320467274178f;
      lp -= logf(p.maxa * (1.f - th * th) + 1e-6f);
      s.an[r * AMAX + j] = th * p.maxa;
      smp[0] = th; smp[1] = sg; smp[2] = e; smp[3] = (raw >= -20.f && raw <= 2.f) ? 1.f : 0.f;
      s.lp[r * AMAX + j] = lp;
    } else {
      float act = tanhf(s.head[r * 2 * AMAX + j]);
      smp[0] = act;
      if (td3_target) act = fminf(fmaxf(act + p.td3_noise_c, -1.f), 1.f);   // enet_td3.py:248-251
      s.an[r * AMAX + j] = act;
    }
  }
  tsync();
  if (p.kind == 0 && tid < TR) {
    float v = 0.f;
    for (int j = 0; j < p.A; ++j) v += s.lp[tid * AMAX + j];
    s.sc[tid * 16 + TSC_LOGP] = v;
  }
  tsync();
}

// optional stage timeline (SMARTCAL_STAMPS): thread 0 of CTA b stamps globaltimer into debug_ts rows 2400 + 2 b .. (16 marks)
__device__ __forceinline__ void t_mark(const RowPhaseParams& p, int i) {
  if (p.tstamp != nullptr && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.tstamp[(size_t)(2400 + 2 * blockIdx.x) * 8 + i] = t;
  }
}

// ======================================================================================================
// PROG_CRITIC
// ======================================================================================================
__device__ void t_critic(const RowPhaseParams& p, const TailS& s, const TRows& rs, const Stream& st) {
  int cur = 0;
  const int tid = threadIdx.x;
  t_mark(p, 0);
  t_load_rows(s.xh2, p.aw.t_h2, H2, rs);
  for (int k = 0; k < 2; ++k) {
    t_load_rows(s.h12 + k * TR * H2, p.tw[k].h12, H2, rs);
    t_load_rows(s.h12 + (2 + k) * TR * H2, p.cw[k].h12, H2, rs);
    t_load_rows(s.zh12 + k * TR * H2, p.cw[k].zh12, H2, rs);
  }
  if (tid < TR * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    s.an[TR * AMAX + r * AMAX + j] = rs.row[r] >= 0 ? ld_w(p.rp.action + (size_t)p.bidx[rs.row[r]] * p.Ap + j) : 0.f;
  }
  if (tid < 2 * TR) {
    const int k = tid / TR, r = tid - k * TR;
    s.rsx[r * 8 + 4 + k] = rs.row[r] >= 0 ? ld_act(p.cw[k].rs12 + rs.row[r]) : 0.f;
  }
  if (tid < TR && rs.row[tid] >= 0) {
    const int slot = p.bidx[rs.row[tid]];
    s.sc[tid * 16 + TSC_REWARD] = ld_w(p.rp.reward + slot);
    s.sc[tid * 16 + TSC_DONE] = p.rp.done[slot] != 0 ? 1.f : 0.f;
  }
  mbar_wait(st.pbar, 0);          // staged parameters have landed (the copies ran under the row loads)
  tsync();
  t_mark(p, 1);
  // ---- policy on s' (SAC: online actor, sampled; TD3: target actor + clipped scalar noise) ----
  t_actor_forward(p, s, st, cur);
  t_mark(p, 2);
  t_policy_head(p, s, rs, true);
  t_mark(p, 3);
  // ---- action branches of all four critics at once: targets on a', online critics on the replay action ----
  const CriticP* nets[NETS4] = {&p.targ[0], &p.targ[1], &p.crit[0], &p.crit[1]};
  const float* an_sel[NETS4] = {s.an, s.an, s.an + TR * AMAX, s.an + TR * AMAX};
  t_action_branches(NETS4, nets, an_sel, p.A, p.Ap, s, st, cur);
  t_mark(p, 4);
  t_critic_q(NETS4, nets, s);
  tsync();
  t_mark(p, 5);
  const float wsum = p.is_wsum ? ld_act(p.is_wsum) : (float)p.B;
  const float coef = 2.f * wsum / ((float)p.B * (float)p.B) * p.closs_scale;   // d/dq of mean(w)*mean((q-y)^2) [DDPG: sum]
  if (tid < TR) {
    float* sc = s.sc + tid * 16;
    const int row = rs.row[tid];
    float dq1 = 0.f, dq2 = 0.f;
    if (row >= 0) {
      const bool done = sc[TSC_DONE] != 0.f;
      const float qt1 = s.qv[0 * TR + tid], qt2 = s.qv[1 * TR + tid], q1 = s.qv[2 * TR + tid], q2 = s.qv[3 * TR + tid];
      float yv;
      if (p.kind == 0) {   // enet_sac.py:573-575
        float m = __fsub_rn(fminf(qt1, qt2), __fmul_rn(p.alpha, sc[TSC_LOGP]));
        if (done) m = 0.f;
        yv = __fadd_rn(__fmul_rn(p.scale, sc[TSC_REWARD]), __fmul_rn(p.gamma, m));
      } else {             // enet_td3.py:255-260
        const float m = done ? 0.f : fminf(qt1, qt2);
        yv = __fadd_rn(sc[TSC_REWARD], __fmul_rn(p.gamma, m));
      }
      p.y[row] = yv;
      dq1 = coef * (q1 - yv);
      dq2 = coef * (q2 - yv);
      p.cw[0].q[row] = q1; p.cw[1].q[row] = q2;
      p.cw[0].dq[row] = dq1; p.cw[1].dq[row] = dq2;
      const float er = 0.5f * (fabsf(q1 - yv) + fabsf(q2 - yv));      // enet_td3.py:266-269
      p.err[row] = er;
      // PER.batch_update priority (enet_sac.py:318-320): float32 (err + eps) -> min(., clip) -> power(., float32 0.6), in double, rounded once
      if (p.per_ps != nullptr) p.per_ps[row] = (float)pow((double)fminf(__fadd_rn(er, 0.01f), p.per_clip), (double)0.6f);
      for (int k = 0; k < 2; ++k) {
        p.cw[k].rs21[row] = s.rsn[((2 + k) * TR + tid) * 2 + 0];
        p.cw[k].rs22[row] = s.rsn[((2 + k) * TR + tid) * 2 + 1];
      }
    }
    sc[TSC_DQ1] = dq1;
    sc[TSC_DQ2] = dq2;
  }
  for (int k = 0; k < 2; ++k) {      // forward caches of the online critics (inputs of the weight-gradient kernels)
    t_store_rows(p.cw[k].h21, s.v21 + (2 + k) * TR * HA1, HA1, rs);
    t_store_rows(p.cw[k].zh21, s.zh21 + (2 + k) * TR * HA1, HA1, rs);
    t_store_rows(p.cw[k].h22, s.v22 + (2 + k) * TR * HA2, HA2, rs);
    t_store_rows(p.cw[k].zh22, s.zh22 + (2 + k) * TR * HA2, HA2, rs);
  }
  tsync();
  t_mark(p, 6);
  // ---- backward of both online critics: d cat = dq * w3 ----
  for (int e = tid; e < 2 * TR * HCAT; e += TL_THREADS) {
    const int rix = e / HCAT, j = e - rix * HCAT, k = rix / TR, r = rix - k * TR;
    const float gval = s.sc[r * 16 + TSC_DQ1 + k] * ldp(p.crit[k].w3 + j);
    if (j < H2) s.ga[rix * H2 + j] = gval; else s.ge[rix * HA2 + (j - H2)] = gval;
  }
  if (tid < 2 * TR) s.qv[tid] = s.rsx[(tid % TR) * 8 + 4 + tid / TR];      // rs12 of (k, r) as a dense [2*TR] array
  tsync();
  {
    const float *g12[2] = {p.crit[0].g12, p.crit[1].g12}, *g22[2] = {p.crit[0].g22, p.crit[1].g22};
    t_ln_bwd(2, s.ga, s.h12 + 2 * TR * H2, s.zh12, H2, g12, s.qv, 1, s.gb);
    t_ln_bwd(2, s.ge, s.v22 + 2 * TR * HA2, s.zh22 + 2 * TR * HA2, HA2, g22, s.rsn + 2 * TR * 2 + 1, 2, s.gf);
  }
  tsync();
  t_mark(p, 7);
  for (int k = 0; k < 2; ++k) {
    t_store_rows(p.cw[k].du12, s.ga + k * TR * H2, H2, rs);
    t_store_rows(p.cw[k].dz12, s.gb + k * TR * H2, H2, rs);
    t_store_rows_hl(p.cw[k].dz12_hl, (size_t)p.B * H2, s.gb + k * TR * H2, H2, rs);
    t_store_rows(p.cw[k].du22, s.ge + k * TR * HA2, HA2, rs);
    t_store_rows(p.cw[k].dz22, s.gf + k * TR * HA2, HA2, rs);
  }
  for (int k = 0; k < 2; ++k) {
    LinJob jb{HA2, nullptr, s.gf + k * TR * HA2, nullptr};
    linear_bwd<HA1>(jb, s.gc + k * TR * HA1, st, cur, s);
  }
  t_mark(p, 8);
  {
    const float* g21[2] = {p.crit[0].g21, p.crit[1].g21};
    t_ln_bwd(2, s.gc, s.v21 + 2 * TR * HA1, s.zh21 + 2 * TR * HA1, HA1, g21, s.rsn + 2 * TR * 2, 2, s.gd);
  }
  tsync();
  for (int k = 0; k < 2; ++k) {
    t_store_rows(p.cw[k].du21, s.gc + k * TR * HA1, HA1, rs);
    t_store_rows(p.cw[k].dz21, s.gd + k * TR * HA1, HA1, rs);
  }
  t_mark(p, 9);
}

// ======================================================================================================
// PROG_ACTOR
// ======================================================================================================
__device__ void t_actor(const RowPhaseParams& p, const TailS& s, const TRows& rs, const Stream& st) {
  int cur = 0;
  const int tid = threadIdx.x;
  const ActorP& a = p.actor;
  const int nq = p.n_q;
  t_load_rows(s.xh2, p.aw.h2, H2, rs);
  t_load_rows(s.zh2, p.aw.zh2, H2, rs);
  for (int k = 0; k < nq; ++k) t_load_rows(s.h12 + k * TR * H2, p.cw[k].p_h12, H2, rs);
  if (tid < TR) s.rsx[tid * 8 + 3] = rs.row[tid] >= 0 ? ld_act(p.aw.rs2 + rs.row[tid]) : 0.f;
  if (tid < TR * p.A && (p.hint_coef != nullptr || p.admm_lambda != nullptr)) {      // replay hints of the rows (second half of s.an)
    const int r = tid / p.A, j = tid - r * p.A;
    s.an[TR * AMAX + r * AMAX + j] = rs.row[r] >= 0 ? ld_w(p.rp.hint + (size_t)p.bidx[rs.row[r]] * p.Ap + j) : 0.f;
  }
  mbar_wait(st.pbar, 0);                              // staged parameters have landed (the copies ran under the row loads)
  tsync();
  t_actor_forward(p, s, st, cur);
  t_store_rows(p.aw.h3, s.x3, H3, rs);
  t_store_rows(p.aw.zh3, s.zh3, H3, rs);
  if (tid < TR && rs.row[tid] >= 0) p.aw.rs3[rs.row[tid]] = s.rsx[tid * 8 + 2];
  t_policy_head(p, s, rs, false);
  if (tid < TR * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    if (rs.row[r] >= 0) p.aw.pi[(size_t)rs.row[r] * p.A + j] = s.an[r * AMAX + j];
  }
  if (tid < TR && rs.row[tid] >= 0 && p.kind == 0) p.aw.logp[rs.row[tid]] = s.sc[tid * 16 + TSC_LOGP];
  // ---- Q_k(s, pi(s)) with the freshly updated critics ----
  const CriticP* nets[NETS4] = {&p.crit[0], &p.crit[1], nullptr, nullptr};
  const float* an_sel[NETS4] = {s.an, s.an, nullptr, nullptr};
  t_action_branches(nq, nets, an_sel, p.A, p.Ap, s, st, cur);
  t_critic_q(nq, nets, s);
  tsync();
  const float wsum = p.is_wsum ? ld_act(p.is_wsum) : (float)p.B;
  const float coef = wsum / ((float)p.B * (float)p.B);           // d/dx of mean(x*w) with the (B,B) broadcast
  if (tid < TR) {
    float* sc = s.sc + tid * 16;
    float g1 = 0.f, g2 = 0.f;
    if (rs.row[tid] >= 0) {
      const float q1 = s.qv[tid], q2 = nq == 2 ? s.qv[TR + tid] : 0.f;
      if (nq == 2) {   // torch.min backward: gradient to the smaller, halves on ties
        if (q1 < q2) g1 = -coef; else if (q2 < q1) g2 = -coef; else { g1 = -0.5f * coef; g2 = -0.5f * coef; }
        p.aw.qmin[rs.row[tid]] = fminf(q1, q2);       // critic_value (enet_sac.py:592)
      } else {
        g1 = -coef;
        p.aw.qmin[rs.row[tid]] = q1;
      }
    }
    sc[TSC_DQ1] = g1;
    sc[TSC_DQ2] = g2;
    sc[TSC_GLP] = rs.row[tid] >= 0 ? p.alpha * coef : 0.f;
  }
  tsync();
  // ---- back through the action branches of the critics to d loss / d action ----
  for (int e = tid; e < nq * TR * HA2; e += TL_THREADS) {
    const int rix = e / HA2, j = e - rix * HA2, k = rix / TR, r = rix - k * TR;
    s.ge[e] = s.sc[r * 16 + TSC_DQ1 + k] * ldp(p.crit[k].w3 + H2 + j);
  }
  tsync();
  {
    const float *g22[2] = {p.crit[0].g22, p.crit[1].g22}, *g21[2] = {p.crit[0].g21, p.crit[1].g21};
    t_ln_bwd(nq, s.ge, s.v22, s.zh22, HA2, g22, s.rsn + 1, 2, s.gf);
    tsync();
    for (int k = 0; k < nq; ++k) {
      LinJob jb{HA2, nullptr, s.gf + k * TR * HA2, nullptr};
      linear_bwd<HA1>(jb, s.gc + k * TR * HA1, st, cur, s);
    }
    t_ln_bwd(nq, s.gc, s.v21, s.zh21, HA1, g21, s.rsn, 2, s.gd);
    tsync();
  }
  {   // d action[r][j] = sum_k sum_n dz21_k[r][n] * w21_k[n][j]   (one warp per (row, j), critics in fixed order)
    const int warp = tid >> 5, lane = tid & 31;
    for (int o = warp; o < TR * p.A; o += TL_WARPS) {
      const int r = o / p.A, j = o - r * p.A;
      float tot = 0.f;
      for (int k = 0; k < nq; ++k) {
        float v = 0.f;
        for (int n = lane; n < HA1; n += 32) v = fmaf(s.gd[(k * TR + r) * HA1 + n], ldp(p.crit[k].w21 + n * p.Ap + j), v);
        tot += warp_sum(v);
      }
      if (lane == 0) s.lp[r * AMAX + j] = tot;
    }
  }
  tsync();
  // ---- back through the policy head ----
  if (tid < TR * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    float da = s.lp[r * AMAX + j];
    const float* smp = s.smp + (r * AMAX + j) * 4;
    const int row = rs.row[r];
    if (p.hint_coef != nullptr && row >= 0)
      da = fmaf(ld_act(p.hint_coef), s.an[r * AMAX + j] - s.an[TR * AMAX + r * AMAX + j], da);
    if (p.admm_lambda != nullptr && row >= 0) {
      const float numel = (float)p.B * (float)p.A, wmean = wsum / (float)p.B;
      const float diff = s.an[r * AMAX + j] - s.an[TR * AMAX + r * AMAX + j];
      da += wmean / numel * ld_act(p.admm_lambda + (size_t)row * p.A + j) + wmean * ld_act(p.admm_rho) / (numel * numel) * diff;
    }
    if (p.kind == 0) {
      const float th = smp[0], sg = smp[1], e = smp[2];
      const float omt = 1.f - th * th;
      const float glp = s.sc[r * 16 + TSC_GLP];
      const float du = da * p.maxa * omt + glp * (2.f * th * omt * p.maxa / (p.maxa * omt + 1e-6f));
      const float dl = (du * sg * e - glp) * smp[3];
      s.head[r * 2 * AMAX + j] = du;
      s.head[r * 2 * AMAX + AMAX + j] = dl;
      if (row >= 0) {
        p.aw.dhead[(size_t)row * 2 * p.A + j] = du;
        p.aw.dhead[(size_t)row * 2 * p.A + p.A + j] = dl;
      }
    } else {
      const float act = smp[0];
      const float dz = da * (1.f - act * act);
      s.head[r * 2 * AMAX + j] = dz;
      if (row >= 0) p.aw.dhead[(size_t)row * p.A + j] = dz;
    }
  }
  tsync();
  for (int e = tid; e < TR * H3; e += TL_THREADS) {
    const int r = e / H3, k = e - r * H3;
    float v = 0.f;
    for (int j = 0; j < p.A; ++j) {
      v = fmaf(s.head[r * 2 * AMAX + j], ldp(a.wmu + j * H3 + k), v);
      if (a.wls != nullptr) v = fmaf(s.head[r * 2 * AMAX + AMAX + j], ldp(a.wls + j * H3 + k), v);
    }
    s.ga[e] = v;
  }
  tsync();
  {
    const float *g3[1] = {a.g3}, *g2[1] = {a.g2};
    t_ln_bwd(1, s.ga, s.x3, s.zh3, H3, g3, s.rsx + 2, 8, s.gb);
    tsync();
    t_store_rows(p.aw.du3, s.ga, H3, rs);
    t_store_rows(p.aw.dz3, s.gb, H3, rs);
    t_store_rows_hl(p.aw.dz3_hl, (size_t)p.B * H3, s.gb, H3, rs);
    LinJob jb{H3, nullptr, s.gb, nullptr};
    linear_bwd<H2>(jb, s.gc, st, cur, s);                            // s.gc has room for [TR][H2] (2*TR*HA1 floats)
    t_ln_bwd(1, s.gc, s.xh2, s.zh2, H2, g2, s.rsx + 3, 8, s.gd);
    tsync();
    t_store_rows(p.aw.du2, s.gc, H2, rs);
    t_store_rows(p.aw.dz2, s.gd, H2, rs);
    t_store_rows_hl(p.aw.dz2_hl, (size_t)p.B * H2, s.gd, H2, rs);
  }
}

__global__ void __launch_bounds__(TL_ALL, 1) k_tail(const RowPhaseParams* __restrict__ pp) {
  __shared__ RowPhaseParams p_sm;
  __shared__ Chunk sched[MAXCH];
  __shared__ __align__(8) uint64_t mbars[2 * NSLOT + 1];
  __shared__ int n_chunks;
  __shared__ PCopy pcopy[16];
  __shared__ int n_pcopy;
  __shared__ unsigned pcopy_bytes;
  const RowPhaseParams& p = load_params(pp, &p_sm);           // blockIdx.y = agent of the ensemble
  extern __shared__ __align__(128) float tail_smem[];
  const TailS s = tail_carve(tail_smem);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(mbars + i, 1);                    // full: one arrive.expect_tx by the producer + the copy's bytes
      mbar_init(mbars + NSLOT + i, TL_WARPS);     // empty: one arrival per consumer warp
    }
    mbar_init(mbars + 2 * NSLOT, 1);              // small-parameter stage
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    // the phase's weight stream, in consumption order (all of these matrices have row pitch == K)
    int n = 0;
    n = sched_add(sched, n, p.actor.w3, H3, H2);
    if (p.program == PROG_CRITIC) {
      for (int k = 0; k < 2; ++k) n = sched_add(sched, n, p.targ[k].w22, HA2, HA1);
      for (int k = 0; k < 2; ++k) n = sched_add(sched, n, p.crit[k].w22, HA2, HA1);
      for (int k = 0; k < 2; ++k) n = sched_add(sched, n, p.crit[k].w22, HA2, HA1);      // transposed pass of the online critics
    } else {
      for (int k = 0; k < p.n_q; ++k) n = sched_add(sched, n, p.crit[k].w22, HA2, HA1);
      for (int k = 0; k < p.n_q; ++k) n = sched_add(sched, n, p.crit[k].w22, HA2, HA1);
      n = sched_add(sched, n, p.actor.w3, H3, H2);                                       // transposed pass of actor fc3
    }
    n_chunks = n;
    // small parameters (biases, LayerNorm affine, fc21, fc3 -> Q, policy heads): three contiguous ranges per critic, two per actor
    // (reference state_dict order, sc_api.cu critic_layout / actor_layout) are copied to shared memory once and the pointers of
    // this CTA's parameter block are redirected there, so no stage of the chain waits on a dependent L2 read
    int nc = 0;
    unsigned total = 0;
    float* cursor = s.pstage;
    auto stage = [&](const float* g0, const float* g1) -> float* {
      const int nf = (int)(g1 - g0);
      if (nf <= 0 || (cursor - s.pstage) + nf > PSTAGE_FLOATS || nc >= 16) return nullptr;
      pcopy[nc++] = PCopy{g0, cursor, (unsigned)nf * 4u};
      total += (unsigned)nf * 4u;
      float* d = cursor;
      cursor += nf;
      return d;
    };
    auto fix = [&](float*& q, const float* g0, float* d) { if (d != nullptr && q != nullptr) q = d + (q - g0); };
    auto stage_critic = [&](CriticP& c) {
      float* d1 = stage(c.w21, c.b21 + HA1);
      fix(c.b21, c.w21, d1); fix(c.w21, c.w21, d1);
      float* d2 = stage(c.b22, c.b3 + 4);
      fix(c.w3, c.b22, d2); fix(c.b3, c.b22, d2); fix(c.b22, c.b22, d2);
      float* d3 = stage(c.g12, c.be22 + HA2);
      fix(c.be12, c.g12, d3); fix(c.g21, c.g12, d3); fix(c.be21, c.g12, d3); fix(c.g22, c.g12, d3); fix(c.be22, c.g12, d3); fix(c.g12, c.g12, d3);
    };
    {
      ActorP& a = p_sm.actor;
      const float* end1 = (a.wls != nullptr ? a.bls : a.bmu) + ((p_sm.A + 3) & ~3);
      float* d1 = stage(a.b3, end1);
      fix(a.wmu, a.b3, d1); fix(a.bmu, a.b3, d1); fix(a.wls, a.b3, d1); fix(a.bls, a.b3, d1); fix(a.b3, a.b3, d1);
      float* d2 = stage(a.g2, a.be3 + H3);
      fix(a.be2, a.g2, d2); fix(a.g3, a.g2, d2); fix(a.be3, a.g2, d2); fix(a.g2, a.g2, d2);
    }
    if (p_sm.program == PROG_CRITIC) {
      stage_critic(p_sm.targ[0]); stage_critic(p_sm.targ[1]);
      stage_critic(p_sm.crit[0]); stage_critic(p_sm.crit[1]);
    } else {
      for (int k = 0; k < p_sm.n_q; ++k) stage_critic(p_sm.crit[k]);
    }
    n_pcopy = nc;
    pcopy_bytes = total;
  }
  if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_tail_ts = (p.program == PROG_CRITIC) ? p.tstamp : nullptr;
  __syncthreads();
  pdl_wait();                     // (set-up above overlapped the previous kernel; the weights it streams were written long before)
  const Stream st{sched, n_chunks, mbars, mbars + NSLOT, s.wbuf, mbars + 2 * NSLOT};
  if (tid >= TL_THREADS) {
    // ---------------- producer warp: small parameters first, then it keeps the ring full across stage boundaries ----------------
    if (tid == TL_THREADS) {
      mbar_expect_tx(st.pbar, pcopy_bytes);           // (also with nothing to stage: the consumers wait on this barrier once)
      for (int i = 0; i < n_pcopy; ++i) bulk_g2s(pcopy[i].dst, pcopy[i].src, pcopy[i].bytes, st.pbar);
    }
    __syncwarp();
    for (int c = 0; c < st.total; ++c) {
      if (tid == TL_THREADS) {
        const int slot = c % NSLOT;
        if (c >= NSLOT) mbar_wait(st.empty + slot, (unsigned)((c / NSLOT) - 1) & 1u);
        const Chunk ch = sched[c];
        const unsigned nbytes = (unsigned)(ch.rows * ch.K) * 4u;
        mbar_expect_tx(st.full + slot, nbytes);
        bulk_g2s(st.ring + (size_t)slot * CH_FLOATS, ch.src, nbytes, st.full + slot);
      }
      __syncwarp();
    }
  } else {
    TRows rs;
#pragma unroll
    for (int t = 0; t < TR; ++t) {
      const int r = (int)blockIdx.x * TR + t;
      rs.row[t] = r < p.B ? r : -1;
    }
    if (p.program == PROG_CRITIC) t_critic(p, s, rs, st);
    else t_actor(p, s, rs, st);
  }
  __syncthreads();
  pdl_launch();                   // late trigger: the dependent (the wide dgrad GEMM) reads everything this CTA stored
}

static size_t tail_smem_bytes() { return (size_t)TAIL_SMEM_FLOATS * sizeof(float); }

cudaError_t launch_tail_variant(const RowPhaseParams* pp, int E, int B, cudaStream_t st) {
  static bool configured[64];
  int d = 0;
  cudaGetDevice(&d);
  d &= 63;
  if (!configured[d]) {
    cudaError_t e = cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem_bytes());
    if (e != cudaSuccess) return e;
    configured[d] = true;
  }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3((B + TR - 1) / TR, E, 1);
  cfg.blockDim = dim3(TL_ALL, 1, 1);
  cfg.dynamicSmemBytes = tail_smem_bytes();
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_tail, pp);
}

}  // namespace SC_TAIL_NS
}  // namespace sc
