// Kernel parameter blocks shared by host (sc_api.cu) and device code.  Plain pointers into the agent arena.
#pragma once
#include <stdint.h>

namespace sc {

// hidden widths of the reference elastic-net MLPs (ref enet_sac.py:357-363,415-419)
constexpr int H1 = 512, H2 = 256, H3 = 128, HA1 = 128, HA2 = 64, HCAT = H2 + HA2;
constexpr int AMAX = 32;          // max n_actions supported by the row-local tail

struct CriticP {                   // one CriticNetwork: PyTorch-native row-major W[out][in], pitch = roundup4(in)
  float *w11, *b11, *w12, *b12, *w21, *b21, *w22, *b22, *w3, *b3;
  float *g11, *be11, *g12, *be12, *g21, *be21, *g22, *be22;
};
struct ActorP {                    // SAC: wmu/bmu + wls/bls heads; TD3: single head in wmu/bmu (fc4), wls = nullptr
  float *w1, *b1, *w2, *b2, *w3, *b3, *wmu, *bmu, *wls, *bls;
  float *g1, *be1, *g2, *be2, *g3, *be3;
};

struct CriticWs {                  // activation / gradient workspace of one ONLINE critic, all [B][width]
  float *h11, *zh11, *rs11, *du11, *dz11;    // width H1   (h = ELU(LN(z)), zh = normalised z, rs = 1/std per row)
  float *h12, *zh12, *rs12, *du12, *dz12;    // width H2
  float *h21, *zh21, *rs21, *du21, *dz21;    // width HA1
  float *h22, *zh22, *rs22, *du22, *dz22;    // width HA2
  float *q, *dq;                             // [B]
  float *p_h11, *p_h12;                      // state-branch forward with the UPDATED weights (actor phase)
  // tf32 big / small parts [2][B][width] of the arrays that are tensor-core operands (written by their producers)
  float *h11_hl, *p_h11_hl, *dz11_hl, *dz12_hl;
};
struct TargetWs { float *h11, *h12, *h11_hl; };
struct ActorWs {
  float *t_h1, *t_h2;                        // pass on s' (critic phase): SAC online actor / TD3 target actor
  float *h1, *zh1, *rs1, *du1, *dz1;         // pass on s (actor phase)
  float *h2, *zh2, *rs2, *du2, *dz2;
  float *h3, *zh3, *rs3, *du3, *dz3;
  float *dhead;                              // [B][2A] (SAC: dmu | dlogsigma) or [B][A] (TD3: d pre-tanh)
  float *pi, *logp, *qmin;                   // [B][A], [B], [B] (min_k Q_k(s,pi(s)) or Q_1 for TD3)
  float *t_h1_hl, *h1_hl, *h2_hl, *dz1_hl, *dz2_hl, *dz3_hl;  // tf32 big / small parts [2][B][width] (tensor-core operands)
};

struct ReplayP {
  float *state, *new_state;                  // [C][Sp]
  float *action, *hint;                      // [C][Ap]
  float *reward;                             // [C]
  uint8_t* done;                             // [C]
  double* tree;                              // [2C-1] float64 sum tree (PER) or nullptr
  int C, S, Sp, A, Ap;
};

enum RowProgram { PROG_CRITIC = 0, PROG_ACTOR = 1, PROG_ACT = 2 };

// CUtensorMap descriptors (device pointers into a host-encoded table) of every matrix the cluster-tiled stages stream
// through TMA.  "A" maps: activations [rows][ld] box {32 floats, 32 rows}; "W" maps: weights [N][ld] same box;
// "P" maps: weight panels for dgrad, box {32 floats, 128 rows}.
struct TensorMaps {
  const void *mb_s, *mb_s2;                  // gathered minibatch states / next states [B][Sp32]
  const void *t_h1, *th11[2], *ch11[2];      // critic phase, layer-1 outputs
  const void *h1, *ph11[2];                  // actor phase, layer-1 outputs
  const void *dz12[2], *dz2;                 // dgrad inputs
  const void *obs, *aux;                     // TD3 ADMM hint loop (enet_td3.py:310-361): d loss1 / d a = mean(w)/numel * lambda + mean(w)*rho/numel^2 * (a - hint)
  const float* admm_lambda;                  // [B][A] Lagrange multipliers, or nullptr
  const float* admm_rho;                     // device scalar: current (adaptive) penalty
  // PROG_ACT
  const void *actor_w1, *actor_w2;           // actor evaluated in this phase (online or target)
  const void *crit_w11[2], *crit_w12[2], *targ_w11[2], *targ_w12[2];
  const void *actor_w2p, *crit_w12p[2];      // dgrad panels
};

struct RowPhaseParams {
  int program, kind;                         // RowProgram, SC_KIND_*
  // device-RNG mode: the N(0,1) draws of the policy head are generated in the kernel (Philox counter = element index / 2, the
  // same values k_fill_normal would have written) and stored to eps_store for inspection
  int eps_inline; unsigned long long eps_seed, eps_stream; float* eps_store;
  int skip_dense;                            // 1: gather + both dense stages already ran as k_dense_tc launches (sc_dense_tc.cu)
  int skip_tail;                             // 1: the row-local tail already ran as a k_tail launch (sc_tail.cu): only the wide dgrad stage is left
  float closs_scale;                         // critic-loss gradient scale: 1 (mean of squares), B (DDPG's sum of squares)
  int B, S, Sp, A, Ap;
  ActorP actor;                              // actor evaluated in this phase
  CriticP crit[2];                           // online critics
  CriticP targ[2];                           // target critics
  CriticWs cw[2];
  TargetWs tw[2];
  ActorWs aw;
  ReplayP rp;
  TensorMaps tm;
  float *mb_s, *mb_s2;                       // [B][Sp32] minibatch gather buffers (zero padded to a multiple of 32 columns)
  int Sp32;
  const int* bidx;                           // [B] replay slots of the minibatch
  const float* eps;                          // [B][A] N(0,1) for this phase (nullptr: none)
  const float* is_wsum;                      // device scalar: sum of IS weights (PER) or nullptr
  float* y;                                  // [B] TD target
  float* err;                                // [B] 0.5(|q1-y|+|q2-y|) (PER errors)
  float* per_ps;                             // [B] float32 priorities min(err+eps,clip)^alpha, or nullptr (no PER)
  float per_clip;
  float* part;                               // LN partial statistics [slot][B][CS][2]
  float gamma, scale, alpha, maxa, td3_noise_c;
  int n_q;                                   // critics used by the actor phase (SAC 2, TD3 1)
  // hint penalty (SAC, enet_sac.py:601-602): coefficient already folded by the host-side pre-pass
  const float* hint_coef;                    // device scalar: (admm_rho*g + rho) * 2*max(0,m-thr) * 2/(B*A), or nullptr
  // TD3 ADMM hint loop (enet_td3.py:310-361): d loss1 / d a = mean(w)/numel * lambda + mean(w)*rho/numel^2 * (a - hint)
  const float* admm_lambda;                  // [B][A] Lagrange multipliers, or nullptr
  const float* admm_rho;                     // device scalar: current (adaptive) penalty
  // PROG_ACT
  const float* obs;                          // [Sp]
  float* act_out;                            // [A]
  float* act_aux;                            // scratch [H1 + H2] activations for the single row
  int dbg;                                   // profiling experiment ($SMARTCAL_DBG): 1 = skip the FFMA work of the GEMM stages
  unsigned long long* tstamp;                // optional [CTAs][8] globaltimer stamps at stage boundaries (debug/profiling)
};

// ---- wgrad + Adam (+ target soft update) work list ------------------------------------------------------
struct AdamHyper {                 // torch.optim.Adam scalars evaluated on the host in double like torch does
  float lr_neg_step;               // -(lr / (1 - beta1^t))
  float bc2_sqrt;                  // sqrt(1 - beta2^t)
  float one_minus_b1, b2, one_minus_b2, eps;
};
struct WgradOp {
  const float* D; int ldd;         // [B][N] upstream pre-activation gradient (this step)
  const float* X; int ldx;         // [B][K] layer input; gather != 0: X is a replay array indexed by bidx
  int gather;
  int N, K;
  float *W, *M, *V, *T;            // parameter, Adam moments, target (nullptr: none); row pitch ldw
  int ldw;
  int tiles_k;                     // number of 32-wide k tiles
  int tile_base, tile_count;       // position in the flat tile list
  int opt;                         // index into AdamHyper[3] (0 actor, 1 critic_1, 2 critic_2)
};
struct WgradTcOp {                 // wide-layer weight gradient on tcgen05 (sc_wgrad_tc.cu): 128 (n) x 64 (k) tiles
  const void* mapD;                // dZ [B][N] big part, boxes {32 floats, 32 rows}, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; the small part's map follows it
  const void* mapX;                // X  [B][K] big part, same box; small part's map follows
  int N, K;
  float *W, *M, *V, *T; int ldw;
  float *Whi, *Wlo, *Thi, *Tlo;    // tf32 big / small copies of W and T kept for the forward / dgrad GEMMs (same pitch)
  int tiles_k;                     // number of 64-wide k tiles
  int tile_base, tile_count;
  int opt;
};
struct VecOp {                     // g[n] = sum_b D[b][n] * (E ? E[b][n] : 1): biases and LayerNorm affine
  const float* D; const float* E; int ld;
  int N;
  float *P, *M, *V, *T;
  int tile_base, tile_count;       // 32 columns per tile
  int opt;
};
struct WgradParams {
  const WgradOp* ops; int nops;
  const VecOp* vops; int nvops;
  int total_tiles;
  const int* tile_map;             // [total_tiles] owner of each tile: op index, or 1000 + vector-op index
  int B;
  const int* bidx;
  AdamHyper hyper[3];
  float tau, one_minus_tau;        // soft update of T (enet_sac.py:531): tau*p + (1-tau)*t
  int do_soft;                     // 0: leave targets alone (TD3 off-steps)
  // loss/diagnostic reduction done by the CTA that draws tile index total_tiles
  int stats_mode;                  // 0 none, 1 critic phase, 2 actor phase
  const float *q1, *q2, *y, *logp, *pi_q;   // [B]
  const float* is_wsum;
  float* stats;                    // device float[16]
  float alpha;
};

// ---- dense forward stages on tcgen05 (sc_dense_tc.cu): H = ELU(LN(X W^T + b)) for up to 5 layers of equal width ----
struct DenseTcOp {
  // every operand arrives split into tf32 big / small parts by its producer: map = big part, the CUtensorMap after it = small part
  const void* mapX;                // X [B][K], boxes {32 floats, 128 rows}, CU_TENSOR_MAP_SWIZZLE_128B
  const void* mapW;                // W [N][K] (PyTorch layout), boxes {32 floats, 64 rows}
  const float *bias, *gamma, *beta;
  float *H, *ZH, *RS;              // forward: outputs [B][N] (ZH, RS nullable) as the tail reads them; dgrad: the PRODUCING layer's h / zh / 1/std (inputs)
  int K;                           // reduction length, multiple of 32 (zero padded / TMA zero filled)
  // dgrad mode (k_dgrad_tc): X = dZ [B][K] of the layer above, W = that layer's weight [K][N] as MN-major panels (boxes {32 floats, 32 rows},
  // CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), gamma = LayerNorm weight of the producing layer; outputs dU = dH * ELU'(h) and dZ (LayerNorm backward)
  float *DU, *DZ;
  float *OUThl;                    // nullable: tf32 big / small parts [2][B][N] of the output the next GEMM consumes (forward: H, dgrad: DZ)
};
struct DenseTcParams {
  DenseTcOp ops[5];
  int nops, N, B, nrb;             // layer width, batch rows, 128-row blocks
  int ks;                          // reduction split (1 or 2); cluster size = (N / 64) * ks
  int late_trigger;                // 1: release the dependent launch only after the outputs are stored (next kernel is a big cluster grid)
  unsigned long long* tl;          // optional debug timeline of CTA 0
  unsigned long long* ts;          // optional [CTA][8] globaltimer stamps: entry, dependency met, accumulator done, end
};

// ---- small kernels of the step, parameter blocks (sc_per.cu, k_gather_mb) ----------------------------------------
struct GatherP {                   // k_gather_mb
  float *mb_s, *mb_s2, *hl;
  const float *state, *new_state;
  const int* bidx;
  int* bidx_mirror;                // nullable: host-supplied indices are copied here (the arena's "batch_idx" buffer)
  int B, Sp, Sp32;
};
struct PerSampleP {                // k_per_sample
  const double* tree; int C, B;
  const double* u01;               // B host-drawn uniforms, or nullptr: Philox(seed, ctr)
  unsigned long long seed, ctr;
  double beta;
  int *tree_idx, *data_idx;
  float *is_w, *is_wsum;
};
struct PerUpdateP {                // k_per_update
  double* tree; int C, B;
  const int* idx; const float* err;
  float clip;
  float* ps_out; const float* ps_in;
};
struct UniformP { int n, B; unsigned long long seed, ctr; int* out; };   // k_sample_uniform
struct HintCoefP {                 // k_hint_coef
  const float *pi, *hint; int Ap; const int* bidx; int B, A;
  float thr, admm_rho;
  const float* rho; float* out; float* rho_update;
  float* stats_rho;                // nullable: *stats_rho = rho after the (optional) dual update
  float* stats_g;                  // nullable: stats_g[0] = g, stats_g[1] = penalty added to the actor loss
};
struct AdmmP {                     // k_admm_update
  int mode, it;
  const float *pi, *hint; int Ap; const int* bidx; int B, A;
  float *lambda, *y0, *a0, *rho;
  float rho0, corr_min;
  float* stats_rho;                // nullable: current penalty for the host-side diagnostics block
};
struct SplitP {                    // k_split: hl = [big | small] parts of up to 12 arrays (fallback / refresh paths)
  int n;
  const float* src[12]; float* hl[12]; unsigned count[12], plane[12];   // floats (multiples of 4); small part at hl + plane
};
struct PrefetchP {                 // k_l2_prefetch: ranges pulled into L2 at the start of a step, in order of first use
  int n;
  const void* ptr[8]; unsigned long long bytes[8];
};
struct StoreP {                    // k_replay_store
  ReplayP rp;
  const float* tr;
  int slot, priority_mode;
  double priority, clip;
};

struct WgradTcParams {
  const WgradTcOp* ops; int nops;
  int total_tiles;
  const int* tile_map;             // [total_tiles] -> op index
  int B;
  AdamHyper hyper[3];
  float tau, one_minus_tau;
  int do_soft;
  unsigned long long* tl;          // optional debug timeline of CTA 0 (SMARTCAL_STAMPS)
};

}  // namespace sc
