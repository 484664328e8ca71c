// C ABI of libsmartcal_b200 (see include/smartcal_b200.h): arena layout, step sequencing, host<->device staging.
#include <cuda.h>
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
#include "sc_params.h"

namespace sc {
size_t rowphase_smem_bytes(int variant);
int rowphase_max_rows(int variant);
int rowphase_panel_rows(int variant);
void rowphase_cluster_capacity(int* out4);
// every kernel of a learn() step takes a device array of E parameter blocks (blockIdx.y = agent of the ensemble)
cudaError_t launch_rowphase(const RowPhaseParams* pp, int E, int variant, int cluster_size, int nclusters, int tm, cudaStream_t st);
cudaError_t launch_tail(const RowPhaseParams* pp, int E, int B, cudaStream_t st);
size_t wgrad_smem_bytes();
cudaError_t launch_wgrad(const WgradParams* pp, int E, int grid, cudaStream_t st);
cudaError_t launch_wgrad_tc(const WgradTcParams* pp, int E, int grid, cudaStream_t st);
cudaError_t launch_dense_tc(const DenseTcParams* pp, int E, int nops, int nrb, int N, int ks, cudaStream_t st);
int dense_tc_max_clusters(int cs);
int dgrad_tc_max_clusters(int cs);
cudaError_t launch_dgrad_tc(const DenseTcParams* pp, int E, int nops, int nrb, int N, int ks, cudaStream_t st);
cudaError_t launch_gather_mb(const GatherP* pp, int E, int B, cudaStream_t st);
cudaError_t launch_split(const SplitP* pp, int E, int grid, cudaStream_t st);
cudaError_t launch_l2_prefetch(const PrefetchP* pp, int E, int grid, cudaStream_t st);
cudaError_t launch_soft_update(const float* src, float* dst, size_t n, float tau, cudaStream_t st);
cudaError_t launch_replay_store(const ReplayP& rp, const float* tr, int slot, int mode, double priority, double clip, cudaStream_t st);
cudaError_t launch_replay_store_t(const StoreP* pp, int E, cudaStream_t st);
cudaError_t launch_replay_fill(const ReplayP& rp, const float* tr, int n, cudaStream_t st);
cudaError_t launch_tree_build(double* tree, int C, const double* pr, int n, cudaStream_t st);
cudaError_t launch_per_sample(const PerSampleP* pp, int E, int B, cudaStream_t st);
cudaError_t launch_per_update(const PerUpdateP* pp, int E, int B, cudaStream_t st);
cudaError_t launch_sample_uniform(const UniformP* pp, int E, int n, cudaStream_t st);
cudaError_t launch_fill_normal(float* out, int n, uint64_t seed, uint64_t stream, cudaStream_t st);
cudaError_t launch_admm_update(const AdmmP* pp, int E, cudaStream_t st);
cudaError_t launch_hint_coef(const HintCoefP* pp, int E, cudaStream_t st);
}  // namespace sc

using namespace sc;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                 \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) return fail(SC_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

// every ABI entry selects the agent's device for its duration (and restores the caller's): handles are usable from any
// current-device context, e.g. one agent per GPU inside one process
struct DevGuard {
  int prev = -1, dev;
  explicit DevGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DevGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
#define DEV(a) DevGuard dev_guard_((a)->device)

static inline int round4i(int x) { return (x + 3) & ~3; }
static inline int round32i(int x) { return (x + 31) & ~31; }

// ---- network tensor tables (reference state_dict order) -------------------------------
struct TDesc {
  std::string name;
  int rows, cols, pitch;
  size_t off;   // floats from the start of the network block
};
struct NetLayout {
  std::vector<TDesc> t;
  size_t size = 0;
  void add(const char* name, int rows, int cols, int pitch = 0) {
    TDesc d{name, rows, cols, pitch ? pitch : round4i(cols), size};
    size += (size_t)rows * d.pitch;
    t.push_back(d);
  }
  const TDesc* find(const char* name) const {
    for (auto& d : t)
      if (d.name == name) return &d;
    return nullptr;
  }
  size_t off(const char* name) const { return find(name)->off; }
};
static NetLayout critic_layout(int S, int A) {
  NetLayout L;
  L.add("fc11.weight", H1, S, round32i(S)); L.add("fc11.bias", 1, H1);   // layer-1 rows padded to whole 32-float TMA k-blocks
  L.add("fc12.weight", H2, H1); L.add("fc12.bias", 1, H2);
  L.add("fc21.weight", HA1, A); L.add("fc21.bias", 1, HA1);
  L.add("fc22.weight", HA2, HA1); L.add("fc22.bias", 1, HA2);
  L.add("fc3.weight", 1, HCAT); L.add("fc3.bias", 1, 1);
  L.add("bn11.weight", 1, H1); L.add("bn11.bias", 1, H1);
  L.add("bn12.weight", 1, H2); L.add("bn12.bias", 1, H2);
  L.add("bn21.weight", 1, HA1); L.add("bn21.bias", 1, HA1);
  L.add("bn22.weight", 1, HA2); L.add("bn22.bias", 1, HA2);
  return L;
}
static NetLayout actor_layout(int S, int A, int kind) {
  NetLayout L;
  L.add("fc1.weight", H1, S, round32i(S)); L.add("fc1.bias", 1, H1);
  L.add("fc2.weight", H2, H1); L.add("fc2.bias", 1, H2);
  L.add("fc3.weight", H3, H2); L.add("fc3.bias", 1, H3);
  if (kind == SC_KIND_SAC) {
    L.add("fc4mu.weight", A, H3); L.add("fc4mu.bias", 1, A);
    L.add("fc4logsigma.weight", A, H3); L.add("fc4logsigma.bias", 1, A);
  } else {
    L.add("fc4.weight", A, H3); L.add("fc4.bias", 1, A);
  }
  L.add("bn1.weight", 1, H1); L.add("bn1.bias", 1, H1);
  L.add("bn2.weight", 1, H2); L.add("bn2.bias", 1, H2);
  L.add("bn3.weight", 1, H3); L.add("bn3.bias", 1, H3);
  return L;
}
static CriticP critic_ptrs(float* b, const NetLayout& L) {
  CriticP c;
  c.w11 = b + L.off("fc11.weight"); c.b11 = b + L.off("fc11.bias");
  c.w12 = b + L.off("fc12.weight"); c.b12 = b + L.off("fc12.bias");
  c.w21 = b + L.off("fc21.weight"); c.b21 = b + L.off("fc21.bias");
  c.w22 = b + L.off("fc22.weight"); c.b22 = b + L.off("fc22.bias");
  c.w3 = b + L.off("fc3.weight"); c.b3 = b + L.off("fc3.bias");
  c.g11 = b + L.off("bn11.weight"); c.be11 = b + L.off("bn11.bias");
  c.g12 = b + L.off("bn12.weight"); c.be12 = b + L.off("bn12.bias");
  c.g21 = b + L.off("bn21.weight"); c.be21 = b + L.off("bn21.bias");
  c.g22 = b + L.off("bn22.weight"); c.be22 = b + L.off("bn22.bias");
  return c;
}
static ActorP actor_ptrs(float* b, const NetLayout& L, int kind) {
  ActorP a;
  a.w1 = b + L.off("fc1.weight"); a.b1 = b + L.off("fc1.bias");
  a.w2 = b + L.off("fc2.weight"); a.b2 = b + L.off("fc2.bias");
  a.w3 = b + L.off("fc3.weight"); a.b3 = b + L.off("fc3.bias");
  if (kind == SC_KIND_SAC) {
    a.wmu = b + L.off("fc4mu.weight"); a.bmu = b + L.off("fc4mu.bias");
    a.wls = b + L.off("fc4logsigma.weight"); a.bls = b + L.off("fc4logsigma.bias");
  } else {
    a.wmu = b + L.off("fc4.weight"); a.bmu = b + L.off("fc4.bias");
    a.wls = nullptr; a.bls = nullptr;
  }
  a.g1 = b + L.off("bn1.weight"); a.be1 = b + L.off("bn1.bias");
  a.g2 = b + L.off("bn2.weight"); a.be2 = b + L.off("bn2.bias");
  a.g3 = b + L.off("bn3.weight"); a.be3 = b + L.off("bn3.bias");
  return a;
}

// ---- arena carving ---------------------------------------------------------------------------------------
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base((char*)b) {}
  void* take(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  float* f(size_t n) { return (float*)take(n * sizeof(float)); }
};

struct KernelTimer {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
  double ms = 0;
  int64_t launches = 0;
};

// ---- step tape + launch plan -----------------------------------------------------------------------------------------------
// One learn() step (of one agent, or of all E agents of an ensemble in lock-step) is planned on the host as a list of
// records -- kernel launches, stream forks / joins, host data blocks -- while every parameter block and every host-supplied
// array (noise, PER uniforms, batch indices, a transition to store) is written into ONE pinned "tape".  The step then costs
// one host->device copy of the tape plus the launches, each kernel reading its parameters from the device copy
// (blockIdx.y = agent).  Identical plans (same records, same tape offsets) are replayed as a CUDA graph.
enum { KID_STORE = 0, KID_PER_SAMPLE, KID_UNIFORM, KID_GATHER, KID_DENSE, KID_ROW, KID_PER_UPDATE, KID_WGRAD, KID_WGRAD_TC, KID_HINT,
       KID_ADMM, KID_TAIL, KID_DGRAD, KID_SPLIT, KID_PREFETCH };
enum { REC_KERNEL = 0, REC_DATA, REC_FORK, REC_JOIN_RECORD, REC_JOIN_WAIT, REC_TBEGIN, REC_TEND };
struct Rec {
  int type, kid, stream;           // stream: 0 main, 1 side (PER), 2 side2 (narrow weight gradients)
  int g[5];                        // kernel geometry (grid / cluster / variant ...), or fork/join pair, or timer id
  size_t off, bytes;               // tape offset of the E parameter blocks / data blocks, bytes per agent
};
struct GraphEntry { int seen = 0; cudaGraphExec_t exec = nullptr; };
struct sc_group {
  std::vector<sc_agent*> ag;
  int E = 0, device = 0;
  cudaStream_t st = nullptr;                  // agent 0's stream: eager launches, tape copies and graph launches go here
  cudaStream_t st_side = nullptr, st_side2 = nullptr, st_cap = nullptr;   // library-owned (st may be the legacy default stream, which cannot be captured)
  cudaEvent_t ev_fork[3] = {nullptr, nullptr, nullptr}, ev_join[3] = {nullptr, nullptr, nullptr};
  // the tape is double-buffered on both sides and uploaded on its own copy stream: the copy for step n+1 runs while the kernels
  // of step n execute (no copy bubble between steps); buffer k is reused once the step that read it two steps ago is done
  char* tape_dev[2] = {nullptr, nullptr};
  char* tape_host[2] = {nullptr, nullptr};
  cudaEvent_t ev_up[2] = {nullptr, nullptr};    // upload of buffer k complete
  cudaEvent_t ev_done[2] = {nullptr, nullptr};  // kernels of the step that read buffer k complete
  cudaStream_t st_copy = nullptr;
  int tk = 0;
  size_t tape_cap = 0, used = 0;
  float *stats_dev = nullptr, *h_stats = nullptr;          // [E][32] (h_stats: two pinned slots of E x 32 floats)
  // deferred read-back (sc_agent_set_stats_mode): the losses of step n are copied to a pinned slot without a synchronisation
  // and handed out by the call for step n+1, so the host prepares the next step while the device still runs this one
  int stats_deferred = 0, stats_slot = 0, stats_pending = 0;
  cudaEvent_t ev_stats[2] = {nullptr, nullptr};
  cudaEvent_t ev_sync = nullptr;                           // synchronous read-back: polled instead of a blocking stream synchronisation
  std::vector<int> pend_flags;                             // per agent: bit 0 did_learn, 1 actor_step, 2 hint_stats
  std::vector<int32_t> pend_counter;
  std::vector<Rec> recs;
  std::map<uint64_t, GraphEntry> graphs;
  int use_graph = 1;
  int plan_error = 0;
  int64_t launches = 0, graph_launches = 0;
};
struct sc_ensemble { sc_group g; };
extern "C" {
static int group_init(sc_group* g, sc_agent** agents, int n);
static void group_free(sc_group* g);
}
constexpr size_t IMM_SLOT = 80 * 1024;   // parameters + B float64 uniforms (B <= 8192)

struct sc_agent {
  sc_config cfg;
  int device;
  sc_group* own;                     // the agent's private group of one (stand-alone learn())
  sc_group* grp;                     // group the agent is being planned in (own, or an ensemble's)
  int gidx;                          // index within grp
  size_t cursor;                     // next plan record (agents after the first re-walk the records of agent 0)
  // immediate parameter slots for the stand-alone entry points (act, sample, update): pinned ring + device ring
  char *imm_host, *imm_dev;
  cudaEvent_t ev_imm[4];
  int imm_k;
  cudaStream_t st;
  int eps_inline[3]; unsigned long long eps_stream[3];
  const float* eps_dev[3];           // where this step's noise arrays live (tape data block, or the arena buffer for in-kernel draws)
  // pinned staging regions are guarded by one event each (recorded after the copy that reads the region, waited on before
  // the host rewrites it) instead of stream-wide synchronisation: 3 transition, 4 PER uniforms, 5 batch indices
  cudaEvent_t ev_stage[6];
  void* arena;
  bool owns_arena;
  size_t arena_bytes;
  int S, Sp, Sp32, A, Ap, B, C, T;    // T = packed transition length; Sp32 = S padded to whole 32-float k-blocks
  float *mb_s, *mb_s2;                // [B][Sp32] minibatch gather buffers
  float* mb_hl;                       // [4][B][Sp32]: tf32 big / small parts of mb_s, then of mb_s2 (X operands of k_dense_tc)
  const void *tmd_s_hi, *tmd_s_lo, *tmd_s2_hi, *tmd_s2_lo;
  unsigned char* tmaps;               // device table of CUtensorMap (128 B each)
  TensorMaps tm_critic, tm_actor, tm_act;   // per-program map pointer sets
  TensorMaps tmd_critic, tmd_actor;         // the same matrices with 128-row boxes for the tcgen05 dense stages (sc_dense_tc.cu)
  const void *tmg_dz12[2], *tmg_dz2;        // k_dgrad_tc: dZ of the layer above, 128-row boxes
  const void *tmg_w12p[2], *tmg_w2p;        //             W12 / W2 as MN-major panels (boxes {32 floats, 32 rows}, ATOM_32B swizzle)
  bool dgrad_tc;
  bool dense_tc;                            // batched passes run their two wide forward layers as k_dense_tc launches
  int dense_cap[3];                         // co-resident k_dense_tc clusters of 4 / 8 / 16 CTAs on this device
  NetLayout Lc, La;
  size_t online_floats;               // actor + c1 + c2
  float *params, *adam_m, *adam_v;    // online blocks
  float* targets;                     // SAC: t1,t2 ; TD3: ta,t1,t2
  // tf32 big / small copies of the parameter and target blocks ([2][floats]: what the tcgen05 forward / dgrad GEMMs read as W).
  // Kept current by the Adam epilogue of k_wgrad_tc; anything else that may have written parameters (a pointer handed out by
  // sc_agent_tensor, sc_agent_soft_update, the FP32 weight-gradient fallback) sets split_dirty and the next learn() refreshes them.
  float *params_hl, *targets_hl;
  size_t target_floats;
  bool split_dirty;
  bool tc_covers[2];                  // [critic, actor]: every wide layer of that update runs in k_wgrad_tc (whose Adam epilogue writes the copies)
  std::map<const float*, float*> hl_of;   // fp32 activation array -> its [big | small] copy
  size_t off_actor, off_c1, off_c2;   // within the online block
  size_t toff_actor, toff_c1, toff_c2;// within the target block
  ReplayP rp;
  CriticWs cw[2];
  TargetWs tw[2];
  ActorWs aw;
  std::map<std::string, sc_tensor_info> buffers;
  float *y, *err, *part, *is_w, *is_wsum, *eps[3], *stats, *hint_out, *rho_dev, *ps_out, *admm_lambda, *admm_y0, *admm_a0;
  int *bidx, *tree_idx;
  const int* bidx_step;               // minibatch slots of the step being planned: bidx, or a host-supplied list in the step tape
  // PER presampling (device RNG): while the rest of learn() runs, the side stream already draws the NEXT minibatch from the
  // updated tree into the alternate buffers; learn() swaps them in if nothing touched the replay state in between
  int *bidx_alt, *tree_idx_alt;
  float *is_w_alt, *is_wsum_alt;
  uint64_t sample_ctr;                // draws of the minibatch samplers (separate from the noise counter)
  bool presampled, quiet;             // quiet: no store / tree write / buffer access since the end of the previous learn()
  double presample_beta;
  unsigned long long* tstamp;
  double* u01;
  float *obs_dev, *act_eps, *act_out, *act_aux, *store_stage;
  WgradOp *critic_ops, *actor_ops;
  int *critic_tile_map, *actor_tile_map;
  // wide layers on tcgen05 (sc_wgrad_tc.cu): op tables, tile maps, tensor maps
  WgradTcOp *critic_tc_ops, *actor_tc_ops;
  int *critic_tc_map, *actor_tc_map;
  int n_critic_tc, n_actor_tc, critic_tc_tiles, actor_tc_tiles;
  unsigned char* tmaps_tc;
  bool use_tc;
  VecOp *critic_vops, *actor_vops;
  int n_critic_ops, n_critic_vops, n_actor_ops, n_actor_vops, critic_tiles, actor_tiles;
  // pinned host staging
  float* h_stage;                     // [3][B*A] noise slots | [T] transition | [Sp32] observation
  float* h_stats;
  int32_t* h_idx;
  double* h_u01;
  float* h_act;
  // state
  int row_variant, cluster_size, nclusters, tile_tm, wgrad_grid, sm_count;
  int64_t mem_cntr, learn_counter, data_pointer, data_length;
  int64_t t_actor, t_critic;          // Adam step counts
  double per_beta;
  bool batch_ready;
  float alpha, rho, admm_rho;
  uint64_t noise_ctr;
  int64_t launches;
  bool timing;
  KernelTimer timers[8];              // 0..5 phase groups (bench.py), 6/7 the tcgen05 dense launches of the critic / actor phase
};

// Anything that may change what the next minibatch draw would see (stores, tree writes, raw buffer access, an explicit
// sample) drops a pending presample and gives its draw counter back, so results do not depend on whether presampling ran.
static void touch_replay(sc_agent* a) {
  if (!a) return;
  if (a->presampled) { a->presampled = false; a->sample_ctr--; }
  a->quiet = false;
}
static void timer_begin(sc_agent* a, int which, cudaEvent_t* ev0, cudaEvent_t* ev1) {
  if (!a->timing) return;
  KernelTimer& t = a->timers[which];
  if (t.pool.empty()) {
    cudaEventCreate(ev0);
    cudaEventCreate(ev1);
  } else {
    *ev0 = t.pool.back().first;
    *ev1 = t.pool.back().second;
    t.pool.pop_back();
  }
  cudaEventRecord(*ev0, a->st);
}
static void timer_end(sc_agent* a, int which, cudaEvent_t ev0, cudaEvent_t ev1) {
  if (!a->timing) return;
  cudaEventRecord(ev1, a->st);
  a->timers[which].pending.push_back({ev0, ev1});
  a->timers[which].launches++;
}

// Builds the whole arena layout; with base == nullptr only measures.
static size_t carve(sc_agent* a, void* base) {
  Carver cv(base);
  const int S = a->S, Sp = a->Sp, A = a->A, Ap = a->Ap, B = a->B, C = a->C;
  const bool td3 = a->cfg.kind == SC_KIND_TD3;
  a->off_actor = 0;
  a->off_c1 = a->La.size;
  a->off_c2 = a->La.size + a->Lc.size;
  a->online_floats = a->La.size + 2 * a->Lc.size;
  a->params = cv.f(a->online_floats);
  a->adam_m = cv.f(a->online_floats);
  a->adam_v = cv.f(a->online_floats);
  a->toff_actor = 0;
  a->toff_c1 = td3 ? a->La.size : 0;
  a->toff_c2 = a->toff_c1 + a->Lc.size;
  a->target_floats = (td3 ? a->La.size : 0) + 2 * a->Lc.size;
  a->targets = cv.f(a->target_floats);
  a->params_hl = cv.f(2 * a->online_floats);
  a->targets_hl = cv.f(2 * a->target_floats);
  // replay
  ReplayP& rp = a->rp;
  rp.C = C; rp.S = S; rp.Sp = Sp; rp.A = A; rp.Ap = Ap;
  rp.state = cv.f((size_t)C * Sp);
  rp.new_state = cv.f((size_t)C * Sp);
  rp.action = cv.f((size_t)C * Ap);
  rp.hint = cv.f((size_t)C * Ap);
  rp.reward = cv.f(C);
  rp.done = (uint8_t*)cv.take(C);
  rp.tree = a->cfg.prioritized ? (double*)cv.take(sizeof(double) * (2 * (size_t)C)) : nullptr;
  // workspace
  auto reg = [&](const std::string& name, void* p, int64_t rows, int64_t cols, int64_t pitch, int eb) {
    sc_tensor_info ti;
    ti.dev_ptr = p; ti.rows = rows; ti.cols = cols; ti.pitch = pitch; ti.elem_bytes = eb; ti.is_int = 0;
    a->buffers[name] = ti;
  };
  auto mat = [&](const std::string& name, int width) {
    float* p = cv.f((size_t)B * width);
    reg(name, p, B, width, width, 4);
    return p;
  };
  a->hl_of.clear();
  auto hl = [&](float* base, int width) {              // [big | small] copy of a tensor-core operand
    float* p = cv.f((size_t)2 * B * width);
    a->hl_of[base] = p;
    return p;
  };
  for (int k = 0; k < 2; ++k) {
    const std::string c = k == 0 ? "c1." : "c2.";
    CriticWs& w = a->cw[k];
    w.h11 = mat(c + "h11", H1); w.zh11 = mat(c + "zh11", H1); w.rs11 = mat(c + "rs11", 1); w.du11 = mat(c + "du11", H1); w.dz11 = mat(c + "dz11", H1);
    w.h12 = mat(c + "h12", H2); w.zh12 = mat(c + "zh12", H2); w.rs12 = mat(c + "rs12", 1); w.du12 = mat(c + "du12", H2); w.dz12 = mat(c + "dz12", H2);
    w.h21 = mat(c + "h21", HA1); w.zh21 = mat(c + "zh21", HA1); w.rs21 = mat(c + "rs21", 1); w.du21 = mat(c + "du21", HA1); w.dz21 = mat(c + "dz21", HA1);
    w.h22 = mat(c + "h22", HA2); w.zh22 = mat(c + "zh22", HA2); w.rs22 = mat(c + "rs22", 1); w.du22 = mat(c + "du22", HA2); w.dz22 = mat(c + "dz22", HA2);
    w.q = mat(c + "q", 1); w.dq = mat(c + "dq", 1);
    w.p_h11 = mat(c + "p_h11", H1); w.p_h12 = mat(c + "p_h12", H2);
    const std::string t = k == 0 ? "t1." : "t2.";
    a->tw[k].h11 = mat(t + "h11", H1); a->tw[k].h12 = mat(t + "h12", H2);
    w.h11_hl = hl(w.h11, H1); w.p_h11_hl = hl(w.p_h11, H1); w.dz11_hl = hl(w.dz11, H1); w.dz12_hl = hl(w.dz12, H2);
    a->tw[k].h11_hl = hl(a->tw[k].h11, H1);
  }
  ActorWs& aw = a->aw;
  aw.t_h1 = mat("a.t_h1", H1); aw.t_h2 = mat("a.t_h2", H2);
  aw.h1 = mat("a.h1", H1); aw.zh1 = mat("a.zh1", H1); aw.rs1 = mat("a.rs1", 1); aw.du1 = mat("a.du1", H1); aw.dz1 = mat("a.dz1", H1);
  aw.h2 = mat("a.h2", H2); aw.zh2 = mat("a.zh2", H2); aw.rs2 = mat("a.rs2", 1); aw.du2 = mat("a.du2", H2); aw.dz2 = mat("a.dz2", H2);
  aw.h3 = mat("a.h3", H3); aw.zh3 = mat("a.zh3", H3); aw.rs3 = mat("a.rs3", 1); aw.du3 = mat("a.du3", H3); aw.dz3 = mat("a.dz3", H3);
  aw.t_h1_hl = hl(aw.t_h1, H1); aw.h1_hl = hl(aw.h1, H1); aw.dz1_hl = hl(aw.dz1, H1); aw.dz2_hl = hl(aw.dz2, H2);
  aw.h2_hl = hl(aw.h2, H2); aw.dz3_hl = hl(aw.dz3, H3);     // fc3 of the actor (128 x 256) is a tensor-core weight gradient too
  aw.dhead = mat("a.dhead", 2 * A); aw.pi = mat("a.pi", A); aw.logp = mat("a.logp", 1); aw.qmin = mat("a.qmin", 1);
  a->y = mat("y", 1);
  a->err = mat("err", 1);
  a->is_w = mat("is_weights", 1);
  a->ps_out = mat("per_ps", 1);
  a->part = cv.f((size_t)8 * B * 16 * 2);
  a->is_wsum = cv.f(4);
  for (int i = 0; i < 3; ++i) a->eps[i] = cv.f((size_t)B * A);
  reg("eps_next", a->eps[0], B, A, A, 4);
  reg("eps_cur", a->eps[1], B, A, A, 4);
  a->stats = cv.f(16);
  a->hint_out = cv.f(8);
  a->admm_lambda = cv.f((size_t)B * A);
  a->admm_y0 = cv.f((size_t)B * A);
  a->admm_a0 = cv.f((size_t)B * A);
  reg("admm_lambda", a->admm_lambda, B, A, A, 4);
  a->rho_dev = cv.f(4);
  reg("rho", a->rho_dev, 1, 1, 1, 4);
  a->bidx = (int*)cv.take(sizeof(int) * B);
  a->tree_idx = (int*)cv.take(sizeof(int) * B);
  a->bidx_alt = (int*)cv.take(sizeof(int) * B);
  a->tree_idx_alt = (int*)cv.take(sizeof(int) * B);
  a->is_w_alt = cv.f(B);
  a->is_wsum_alt = cv.f(4);
  reg("batch_idx", a->bidx, B, 1, 1, 4);
  reg("tree_idx", a->tree_idx, B, 1, 1, 4);
  a->buffers["batch_idx"].is_int = 1;   // int32
  a->buffers["tree_idx"].is_int = 1;
  a->u01 = (double*)cv.take(sizeof(double) * B);
  a->mb_s = cv.f((size_t)B * a->Sp32);
  a->mb_s2 = cv.f((size_t)B * a->Sp32);
  a->mb_hl = cv.f((size_t)4 * B * a->Sp32);
  a->hl_of[a->mb_s] = a->mb_hl;
  a->hl_of[a->mb_s2] = a->mb_hl + (size_t)2 * B * a->Sp32;
  reg("mb_state", a->mb_s, B, S, a->Sp32, 4);
  reg("mb_new_state", a->mb_s2, B, S, a->Sp32, 4);
  a->tmaps = (unsigned char*)cv.take(128 * 128);
  a->obs_dev = cv.f(a->Sp32);
  a->act_eps = cv.f(round4i(A));
  a->act_out = cv.f(round4i(A));
  a->act_aux = cv.f(H1 + H2);
  a->store_stage = cv.f((size_t)a->T * 1);
  a->tstamp = (unsigned long long*)cv.take(sizeof(unsigned long long) * 8 * 3072);
  reg("debug_ts", a->tstamp, 3072, 8, 8, 8);
  a->critic_tc_ops = (WgradTcOp*)cv.take(sizeof(WgradTcOp) * 8);
  a->actor_tc_ops = (WgradTcOp*)cv.take(sizeof(WgradTcOp) * 8);
  a->critic_tc_map = (int*)cv.take(sizeof(int) * 512);
  a->actor_tc_map = (int*)cv.take(sizeof(int) * 512);
  a->tmaps_tc = (unsigned char*)cv.take(128 * 48);
  a->critic_ops = (WgradOp*)cv.take(sizeof(WgradOp) * 16);
  a->actor_ops = (WgradOp*)cv.take(sizeof(WgradOp) * 8);
  a->critic_vops = (VecOp*)cv.take(sizeof(VecOp) * 32);
  a->critic_tile_map = (int*)cv.take(sizeof(int) * 4096);
  a->actor_tile_map = (int*)cv.take(sizeof(int) * 2048);
  a->actor_vops = (VecOp*)cv.take(sizeof(VecOp) * 16);
  reg("state_memory", rp.state, C, S, Sp, 4);
  reg("new_state_memory", rp.new_state, C, S, Sp, 4);
  reg("action_memory", rp.action, C, A, Ap, 4);
  reg("hint_memory", rp.hint, C, A, Ap, 4);
  reg("reward_memory", rp.reward, C, 1, 1, 4);
  reg("terminal_memory", rp.done, C, 1, 1, 1);
  if (rp.tree) reg("tree", rp.tree, 2 * (int64_t)C - 1, 1, 1, 8);
  return (cv.off + 255) & ~(size_t)255;
}

static int check_cfg(const sc_config* c) {
  if (!c) return fail(SC_ERR_ARG, "null config");
  if (c->abi_version != SC_ABI_VERSION) return fail(SC_ERR_ARG, "abi_version %d != %d", c->abi_version, SC_ABI_VERSION);
  if (c->kind != SC_KIND_SAC && c->kind != SC_KIND_TD3) return fail(SC_ERR_ARG, "unknown kind %d", c->kind);
  if (c->state_dim < 1 || c->n_actions < 1 || c->n_actions > AMAX) return fail(SC_ERR_ARG, "state_dim>=1 and 1<=n_actions<=%d required", AMAX);
  if (c->batch_size < 1 || c->batch_size > 8192) return fail(SC_ERR_ARG, "batch_size must be in [1,8192]");
  if (c->mem_size < 1) return fail(SC_ERR_ARG, "mem_size must be positive");
  if (c->prioritized && (c->mem_size & (c->mem_size - 1))) return fail(SC_ERR_ARG, "Capacity must be power of 2.");
  return SC_OK;
}
static void init_dims(sc_agent* a, const sc_config* cfg) {
  a->cfg = *cfg;
  a->S = cfg->state_dim; a->Sp = round4i(a->S); a->Sp32 = round32i(a->S);
  a->A = cfg->n_actions; a->Ap = round4i(a->A);
  a->B = cfg->batch_size; a->C = cfg->mem_size;
  a->T = 2 * a->S + 2 * a->A + 2;
  a->Lc = critic_layout(a->S, a->A);
  a->La = actor_layout(a->S, a->A, cfg->kind);
}

extern "C" {

const char* sc_last_error(void) { return g_err; }
void sc_set_last_error_(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }   // shared with sc_calib.cu (not part of the ABI)
int sc_abi_version(void) { return SC_ABI_VERSION; }

size_t sc_agent_arena_bytes(const sc_config* cfg) {
  if (check_cfg(cfg) != SC_OK) return 0;
  sc_agent tmp;
  init_dims(&tmp, cfg);
  return carve(&tmp, nullptr);
}

struct TcCand { const float* D; int ldd; const float* X; int ldx; int N, K; float *W, *M, *V, *T; int ldw, opt; };
static std::vector<TcCand>* g_tc_sink = nullptr;      // when set, wide ops are routed to the tcgen05 kernel instead of the FFMA tile list
static void add_w(std::vector<WgradOp>& v, int& tiles, const float* D, int ldd, const float* X, int ldx, int gather, int N, int K,
                  float* W, float* M, float* V, float* T, int ldw, int opt) {
  if (g_tc_sink != nullptr && !gather && N % 128 == 0 && K >= 64 && (ldd & 3) == 0 && (ldx & 3) == 0 && (ldw & 3) == 0) {
    g_tc_sink->push_back(TcCand{D, ldd, X, ldx, N, K, W, M, V, T, ldw, opt});
    return;
  }
  WgradOp o;
  o.D = D; o.ldd = ldd; o.X = X; o.ldx = ldx; o.gather = gather; o.N = N; o.K = K;
  o.W = W; o.M = M; o.V = V; o.T = T; o.ldw = ldw;
  o.tiles_k = (K + 31) / 32;
  o.tile_base = tiles;
  o.tile_count = ((N + 31) / 32) * o.tiles_k;
  o.opt = opt;
  tiles += o.tile_count;
  v.push_back(o);
}
static void add_v(std::vector<VecOp>& v, int& tiles, const float* D, const float* E, int ld, int N, float* P, float* M, float* V,
                  float* T, int opt) {
  VecOp o;
  o.D = D; o.E = E; o.ld = ld; o.N = N; o.P = P; o.M = M; o.V = V; o.T = T;
  o.tile_base = tiles;
  o.tile_count = (N + 31) / 32;
  o.opt = opt;
  tiles += o.tile_count;
  v.push_back(o);
}

// ---- TMA tensor maps ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// row-major float matrix [rows][ld] (ld a multiple of 32), box {32 floats = 128 B, box_rows}, 128-byte swizzle
static int encode_map(CUtensorMap* out, const void* base, int ld, int rows, int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(SC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SC_ERR_CUDA, "cuTensorMapEncodeTiled(ld=%d rows=%d box_rows=%d) failed: %d", ld, rows, box_rows, (int)r);
  return SC_OK;
}

extern "C++" {
namespace sc {
// shared with sc_calib.cu: row-major float matrix [rows][ld], box {32 floats, box_rows}, 128-byte swizzle, out-of-range elements read as 0
int encode_map_f32_2d(void* out, const void* base, int ld, int rows, int box_rows) {
  return encode_map(reinterpret_cast<CUtensorMap*>(out), base, ld, rows, box_rows);
}
}  // namespace sc
}

static int build_tensor_maps(sc_agent* a) {
  std::vector<CUtensorMap> host;
  host.reserve(128);
  int rc = SC_OK;
  auto add = [&](const void* base, int ld, int rows, int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) -> const void* {
    alignas(64) CUtensorMap m;
    if (rc == SC_OK) rc = encode_map(&m, base, ld, rows, box_rows, swz);
    host.push_back(m);
    return a->tmaps + (host.size() - 1) * sizeof(CUtensorMap);
  };
  const bool td3 = a->cfg.kind == SC_KIND_TD3;
  const int B = a->B, K1 = a->Sp32;
  TensorMaps t;
  memset(&t, 0, sizeof(t));
  // box heights follow the row-phase geometry: activations are fetched one GEMM tile (8 TM rows) at a time, weights one
  // CTA column slice (N / cluster size features) at a time, dgrad weight panels PROW reduction rows x 32 input columns
  const int RB = 8 * a->tile_tm, CSZ = a->cluster_size;
  const int W1 = ((H1 + CSZ - 1) / CSZ + 3) & ~3, W2 = ((H2 + CSZ - 1) / CSZ + 3) & ~3, PR = rowphase_panel_rows(a->row_variant);
  if ((W1 != 16 && W1 != 32 && W1 != 64) || (W2 != 16 && W2 != 32 && W2 != 64))
    return fail(SC_ERR_STATE, "row-phase column slices must be 16, 32 or 64 wide (got %d, %d)", W1, W2);
  t.mb_s = add(a->mb_s, K1, B, RB);
  t.mb_s2 = add(a->mb_s2, K1, B, RB);
  t.t_h1 = add(a->aw.t_h1, H1, B, RB);
  t.h1 = add(a->aw.h1, H1, B, RB);
  t.dz2 = add(a->aw.dz2, H2, B, RB);
  t.obs = add(a->obs_dev, K1, 1, RB);
  t.aux = add(a->act_aux, H1, 1, RB);
  CriticP C[2] = {critic_ptrs(a->params + a->off_c1, a->Lc), critic_ptrs(a->params + a->off_c2, a->Lc)};
  CriticP T[2] = {critic_ptrs(a->targets + a->toff_c1, a->Lc), critic_ptrs(a->targets + a->toff_c2, a->Lc)};
  for (int k = 0; k < 2; ++k) {
    t.th11[k] = add(a->tw[k].h11, H1, B, RB);
    t.ch11[k] = add(a->cw[k].h11, H1, B, RB);
    t.ph11[k] = add(a->cw[k].p_h11, H1, B, RB);
    t.dz12[k] = add(a->cw[k].dz12, H2, B, RB);
    t.crit_w11[k] = add(C[k].w11, K1, H1, W1);
    t.crit_w12[k] = add(C[k].w12, H1, H2, W2);
    t.targ_w11[k] = add(T[k].w11, K1, H1, W1);
    t.targ_w12[k] = add(T[k].w12, H1, H2, W2);
    t.crit_w12p[k] = add(C[k].w12, H1, H2, PR);
  }
  ActorP on = actor_ptrs(a->params + a->off_actor, a->La, a->cfg.kind);
  const void* on_w1 = add(on.w1, K1, H1, W1);
  const void* on_w2 = add(on.w2, H1, H2, W2);
  t.actor_w2p = add(on.w2, H1, H2, PR);
  const void *tg_w1 = on_w1, *tg_w2 = on_w2;
  if (td3) {
    ActorP tg = actor_ptrs(a->targets + a->toff_actor, a->La, a->cfg.kind);
    tg_w1 = add(tg.w1, K1, H1, W1);
    tg_w2 = add(tg.w2, H1, H2, W2);
  }
  // second set for the tcgen05 dense stages: 128-row boxes of the layer inputs and of the [N][K] weights
  TensorMaps d;
  memset(&d, 0, sizeof(d));
  const void *d_on_w1 = nullptr, *d_on_w2 = nullptr, *d_tg_w1 = nullptr, *d_tg_w2 = nullptr;
  {
    const int DB = 128, DW = 64;               // box rows: 128 batch rows, 64 output features (sc_dense_tc.cu tile)
    // every tensor-core operand is read as its tf32 big / small parts: two consecutive maps (big, then small) per matrix
    auto add2 = [&](const float* hlbase, size_t plane, int ld, int rows, int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) -> const void* {
      const void* first = add(hlbase, ld, rows, box_rows, swz);
      add(hlbase + plane, ld, rows, box_rows, swz);
      return first;
    };
    auto act2 = [&](const float* fp32, int width, int box_rows) { return add2(a->hl_of.at(fp32), (size_t)B * width, width, B, box_rows); };
    auto w_hl = [&](const float* W) -> const float* {            // big part of a weight matrix; the small part is one block further
      if (W >= a->params && W < a->params + a->online_floats) return a->params_hl + (W - a->params);
      return a->targets_hl + (W - a->targets);
    };
    auto w_plane = [&](const float* W) -> size_t { return (W >= a->params && W < a->params + a->online_floats) ? a->online_floats : a->target_floats; };
    auto wt2 = [&](const float* W, int ld, int rows, int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
      return add2(w_hl(W), w_plane(W), ld, rows, box_rows, swz);
    };
    d.mb_s = add(a->mb_s, K1, B, DB);
    d.mb_s2 = add(a->mb_s2, K1, B, DB);
    const size_t plane = (size_t)B * K1;
    a->tmd_s_hi = add(a->mb_hl, K1, B, DB);
    a->tmd_s_lo = add(a->mb_hl + plane, K1, B, DB);
    a->tmd_s2_hi = add(a->mb_hl + 2 * plane, K1, B, DB);
    a->tmd_s2_lo = add(a->mb_hl + 3 * plane, K1, B, DB);
    d.t_h1 = act2(a->aw.t_h1, H1, DB);
    d.h1 = act2(a->aw.h1, H1, DB);
    for (int k = 0; k < 2; ++k) {
      d.th11[k] = act2(a->tw[k].h11, H1, DB);
      d.ch11[k] = act2(a->cw[k].h11, H1, DB);
      d.ph11[k] = act2(a->cw[k].p_h11, H1, DB);
      d.crit_w11[k] = wt2(C[k].w11, K1, H1, DW);
      d.crit_w12[k] = wt2(C[k].w12, H1, H2, DW);
      d.targ_w11[k] = wt2(T[k].w11, K1, H1, DW);
      d.targ_w12[k] = wt2(T[k].w12, H1, H2, DW);
    }
    d_on_w1 = wt2(on.w1, K1, H1, DW);
    d_on_w2 = wt2(on.w2, H1, H2, DW);
    for (int k = 0; k < 2; ++k) {
      a->tmg_dz12[k] = act2(a->cw[k].dz12, H2, DB);
      a->tmg_w12p[k] = wt2(C[k].w12, H1, H2, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    }
    a->tmg_dz2 = act2(a->aw.dz2, H2, DB);
    a->tmg_w2p = wt2(on.w2, H1, H2, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    d_tg_w1 = d_on_w1; d_tg_w2 = d_on_w2;
    if (td3) {
      ActorP tg = actor_ptrs(a->targets + a->toff_actor, a->La, a->cfg.kind);
      d_tg_w1 = wt2(tg.w1, K1, H1, DW);
      d_tg_w2 = wt2(tg.w2, H1, H2, DW);
    }
  }
  if (rc != SC_OK) return rc;
  if (host.size() > 128) return fail(SC_ERR_STATE, "tensor map table overflow");
  a->tmd_critic = d; a->tmd_actor = d;
  a->tmd_critic.actor_w1 = td3 ? d_tg_w1 : d_on_w1;
  a->tmd_critic.actor_w2 = td3 ? d_tg_w2 : d_on_w2;
  a->tmd_actor.actor_w1 = d_on_w1; a->tmd_actor.actor_w2 = d_on_w2;
  a->tm_critic = t; a->tm_actor = t; a->tm_act = t;
  a->tm_critic.actor_w1 = td3 ? tg_w1 : on_w1;      // critic phase evaluates the TD3 target actor / the SAC online actor on s'
  a->tm_critic.actor_w2 = td3 ? tg_w2 : on_w2;
  a->tm_actor.actor_w1 = on_w1; a->tm_actor.actor_w2 = on_w2;
  a->tm_act.actor_w1 = on_w1; a->tm_act.actor_w2 = on_w2;
  CK(cudaMemcpyAsync(a->tmaps, host.data(), host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, a->st));
  CK(cudaStreamSynchronize(a->st));
  return SC_OK;
}

// tcgen05 weight-gradient tables: tensor maps (ATOM_32B swizzle, boxes of 32 batch rows x 32 floats), op list, tile map
static int upload_tc_tables(sc_agent* a, const std::vector<TcCand>& cand, WgradTcOp* dev_ops, int* dev_map, int& nops, int& ntiles,
                            std::vector<CUtensorMap>& maps) {
  std::vector<WgradTcOp> ops;
  std::vector<int> tmap;
  int tiles = 0;
  for (const TcCand& c : cand) {
    WgradTcOp o;
    // operands as tf32 big / small parts: two consecutive maps each (the kernel addresses the small part as map + 1)
    alignas(64) CUtensorMap md[2], mx[2];
    auto dh = a->hl_of.find(c.D), xh = a->hl_of.find(c.X);
    if (dh == a->hl_of.end() || xh == a->hl_of.end()) return fail(SC_ERR_STATE, "tcgen05 wgrad operand without a big / small copy");
    int rc = SC_OK;
    for (int part = 0; part < 2 && rc == SC_OK; ++part) {
      rc = encode_map(&md[part], dh->second + (size_t)part * a->B * c.ldd, c.ldd, a->B, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
      if (rc == SC_OK) rc = encode_map(&mx[part], xh->second + (size_t)part * a->B * c.ldx, c.ldx, a->B, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    }
    if (rc != SC_OK) return rc;
    maps.push_back(md[0]);
    o.mapD = a->tmaps_tc + (maps.size() - 1) * sizeof(CUtensorMap);
    maps.push_back(md[1]);
    maps.push_back(mx[0]);
    o.mapX = a->tmaps_tc + (maps.size() - 1) * sizeof(CUtensorMap);
    maps.push_back(mx[1]);
    o.N = c.N; o.K = c.K; o.W = c.W; o.M = c.M; o.V = c.V; o.T = c.T; o.ldw = c.ldw; o.opt = c.opt;
    o.Whi = a->params_hl + (c.W - a->params); o.Wlo = o.Whi + a->online_floats;
    o.Thi = c.T ? a->targets_hl + (c.T - a->targets) : nullptr; o.Tlo = c.T ? o.Thi + a->target_floats : nullptr;
    o.tiles_k = (c.K + 63) / 64;
    o.tile_base = tiles;
    o.tile_count = (c.N / 128) * o.tiles_k;
    for (int t = 0; t < o.tile_count; ++t) tmap.push_back((int)ops.size());
    tiles += o.tile_count;
    ops.push_back(o);
  }
  if (ops.size() > 8 || tiles > 512 || maps.size() > 48) return fail(SC_ERR_STATE, "tcgen05 wgrad table overflow");
  nops = (int)ops.size(); ntiles = tiles;
  if (!ops.empty()) {
    CK(cudaMemcpyAsync(dev_ops, ops.data(), ops.size() * sizeof(WgradTcOp), cudaMemcpyHostToDevice, a->st));
    CK(cudaMemcpyAsync(dev_map, tmap.data(), tmap.size() * sizeof(int), cudaMemcpyHostToDevice, a->st));
    CK(cudaStreamSynchronize(a->st));
  }
  return SC_OK;
}

static int build_op_tables(sc_agent* a) {
  const bool td3 = a->cfg.kind == SC_KIND_TD3;
  std::vector<WgradOp> cw;
  std::vector<VecOp> cv;
  int tiles = 0;
  // wide layers go to the tensor-core kernel unless SMARTCAL_WGRAD=ffma pins the FP32 tile kernel (A/B comparisons)
  const char* wg = getenv("SMARTCAL_WGRAD");
  a->use_tc = !(wg && !strcmp(wg, "ffma"));
  // wide forward layers of the batched passes: tcgen05 launches unless SMARTCAL_DENSE=mma keeps them in the row-phase kernel
  const char* dn = getenv("SMARTCAL_DENSE");
  a->dense_tc = !(dn && !strcmp(dn, "mma"));
  if (a->dense_tc) {
    a->dense_cap[0] = dense_tc_max_clusters(4); a->dense_cap[1] = dense_tc_max_clusters(8); a->dense_cap[2] = dense_tc_max_clusters(16);
    if (a->dense_cap[0] < 1 || a->dense_cap[1] < 1) a->dense_tc = false;     // cannot host the clusters: keep the wide layers in k_rowphase
  }
  // wide input gradients on tcgen05 as well (SMARTCAL_DGRAD=mma keeps the mma.sync stage of k_rowphase)
  const char* dg = getenv("SMARTCAL_DGRAD");
  a->dgrad_tc = a->dense_tc && !(dg && !strcmp(dg, "mma")) && dgrad_tc_max_clusters(8) >= 1;
  std::vector<TcCand> tc_critic, tc_actor;
  std::vector<CUtensorMap> tc_maps;
  g_tc_sink = a->use_tc ? &tc_critic : nullptr;
  for (int k = 0; k < 2; ++k) {
    const size_t off = k == 0 ? a->off_c1 : a->off_c2;
    const size_t toff = k == 0 ? a->toff_c1 : a->toff_c2;
    CriticP P = critic_ptrs(a->params + off, a->Lc), M = critic_ptrs(a->adam_m + off, a->Lc), V = critic_ptrs(a->adam_v + off, a->Lc),
            T = critic_ptrs(a->targets + toff, a->Lc);
    const CriticWs& w = a->cw[k];
    const int opt = 1 + k;
    add_w(cw, tiles, w.dz11, H1, a->mb_s, a->Sp32, 0, H1, a->S, P.w11, M.w11, V.w11, T.w11, a->Sp32, opt);
    add_w(cw, tiles, w.dz12, H2, w.h11, H1, 0, H2, H1, P.w12, M.w12, V.w12, T.w12, H1, opt);
    add_w(cw, tiles, w.dz21, HA1, a->rp.action, a->Ap, 1, HA1, a->A, P.w21, M.w21, V.w21, T.w21, a->Ap, opt);
    add_w(cw, tiles, w.dz22, HA2, w.h21, HA1, 0, HA2, HA1, P.w22, M.w22, V.w22, T.w22, HA1, opt);
    add_w(cw, tiles, w.dq, 1, w.h12, H2, 0, 1, H2, P.w3, M.w3, V.w3, T.w3, HCAT, opt);
    add_w(cw, tiles, w.dq, 1, w.h22, HA2, 0, 1, HA2, P.w3 + H2, M.w3 + H2, V.w3 + H2, T.w3 + H2, HCAT, opt);
    add_v(cv, tiles, w.dz11, nullptr, H1, H1, P.b11, M.b11, V.b11, T.b11, opt);
    add_v(cv, tiles, w.du11, w.zh11, H1, H1, P.g11, M.g11, V.g11, T.g11, opt);
    add_v(cv, tiles, w.du11, nullptr, H1, H1, P.be11, M.be11, V.be11, T.be11, opt);
    add_v(cv, tiles, w.dz12, nullptr, H2, H2, P.b12, M.b12, V.b12, T.b12, opt);
    add_v(cv, tiles, w.du12, w.zh12, H2, H2, P.g12, M.g12, V.g12, T.g12, opt);
    add_v(cv, tiles, w.du12, nullptr, H2, H2, P.be12, M.be12, V.be12, T.be12, opt);
    add_v(cv, tiles, w.dz21, nullptr, HA1, HA1, P.b21, M.b21, V.b21, T.b21, opt);
    add_v(cv, tiles, w.du21, w.zh21, HA1, HA1, P.g21, M.g21, V.g21, T.g21, opt);
    add_v(cv, tiles, w.du21, nullptr, HA1, HA1, P.be21, M.be21, V.be21, T.be21, opt);
    add_v(cv, tiles, w.dz22, nullptr, HA2, HA2, P.b22, M.b22, V.b22, T.b22, opt);
    add_v(cv, tiles, w.du22, w.zh22, HA2, HA2, P.g22, M.g22, V.g22, T.g22, opt);
    add_v(cv, tiles, w.du22, nullptr, HA2, HA2, P.be22, M.be22, V.be22, T.be22, opt);
    add_v(cv, tiles, w.dq, nullptr, 1, 1, P.b3, M.b3, V.b3, T.b3, opt);
  }
  g_tc_sink = a->use_tc ? &tc_actor : nullptr;
  a->n_critic_ops = (int)cw.size(); a->n_critic_vops = (int)cv.size(); a->critic_tiles = tiles;
  auto tile_table = [](const std::vector<WgradOp>& w, const std::vector<VecOp>& v, int total) {
    std::vector<int> t(total, -1);
    for (size_t i = 0; i < w.size(); ++i)
      for (int k = 0; k < w[i].tile_count; ++k) t[w[i].tile_base + k] = (int)i;
    for (size_t i = 0; i < v.size(); ++i)
      for (int k = 0; k < v[i].tile_count; ++k) t[v[i].tile_base + k] = 1000 + (int)i;
    return t;
  };
  if (tiles > 4096) return fail(SC_ERR_STATE, "critic tile table overflow (%d)", tiles);
  {
    std::vector<int> tt = tile_table(cw, cv, tiles);
    CK(cudaMemcpyAsync(a->critic_tile_map, tt.data(), tt.size() * sizeof(int), cudaMemcpyHostToDevice, a->st));
    CK(cudaStreamSynchronize(a->st));
  }
  CK(cudaMemcpyAsync(a->critic_ops, cw.data(), cw.size() * sizeof(WgradOp), cudaMemcpyHostToDevice, a->st));
  CK(cudaMemcpyAsync(a->critic_vops, cv.data(), cv.size() * sizeof(VecOp), cudaMemcpyHostToDevice, a->st));

  std::vector<WgradOp> awv;
  std::vector<VecOp> avv;
  tiles = 0;
  {
    ActorP P = actor_ptrs(a->params + a->off_actor, a->La, a->cfg.kind), M = actor_ptrs(a->adam_m + a->off_actor, a->La, a->cfg.kind),
           V = actor_ptrs(a->adam_v + a->off_actor, a->La, a->cfg.kind);
    ActorP T;
    memset(&T, 0, sizeof(T));
    if (td3) T = actor_ptrs(a->targets + a->toff_actor, a->La, a->cfg.kind);
    const ActorWs& w = a->aw;
    const int A = a->A, ldh = td3 ? A : 2 * A;
    add_w(awv, tiles, w.dz1, H1, a->mb_s, a->Sp32, 0, H1, a->S, P.w1, M.w1, V.w1, T.w1, a->Sp32, 0);
    add_w(awv, tiles, w.dz2, H2, w.h1, H1, 0, H2, H1, P.w2, M.w2, V.w2, T.w2, H1, 0);
    add_w(awv, tiles, w.dz3, H3, w.h2, H2, 0, H3, H2, P.w3, M.w3, V.w3, T.w3, H2, 0);
    add_w(awv, tiles, w.dhead, ldh, w.h3, H3, 0, A, H3, P.wmu, M.wmu, V.wmu, T.wmu, H3, 0);
    if (!td3) add_w(awv, tiles, w.dhead + A, ldh, w.h3, H3, 0, A, H3, P.wls, M.wls, V.wls, T.wls, H3, 0);
    add_v(avv, tiles, w.dz1, nullptr, H1, H1, P.b1, M.b1, V.b1, T.b1, 0);
    add_v(avv, tiles, w.du1, w.zh1, H1, H1, P.g1, M.g1, V.g1, T.g1, 0);
    add_v(avv, tiles, w.du1, nullptr, H1, H1, P.be1, M.be1, V.be1, T.be1, 0);
    add_v(avv, tiles, w.dz2, nullptr, H2, H2, P.b2, M.b2, V.b2, T.b2, 0);
    add_v(avv, tiles, w.du2, w.zh2, H2, H2, P.g2, M.g2, V.g2, T.g2, 0);
    add_v(avv, tiles, w.du2, nullptr, H2, H2, P.be2, M.be2, V.be2, T.be2, 0);
    add_v(avv, tiles, w.dz3, nullptr, H3, H3, P.b3, M.b3, V.b3, T.b3, 0);
    add_v(avv, tiles, w.du3, w.zh3, H3, H3, P.g3, M.g3, V.g3, T.g3, 0);
    add_v(avv, tiles, w.du3, nullptr, H3, H3, P.be3, M.be3, V.be3, T.be3, 0);
    add_v(avv, tiles, w.dhead, nullptr, ldh, A, P.bmu, M.bmu, V.bmu, T.bmu, 0);
    if (!td3) add_v(avv, tiles, w.dhead + A, nullptr, ldh, A, P.bls, M.bls, V.bls, T.bls, 0);
  }
  g_tc_sink = nullptr;
  {
    // are all the weights k_dense_tc reads (fc1 / fc2 of every network) updated by the tensor-core kernel?
    int nc = 0, na = 0;
    for (const TcCand& c : tc_critic) nc += (c.K == a->S || c.K == H1) ? 1 : 0;
    for (const TcCand& c : tc_actor) na += (c.N == H1 || c.N == H2) ? 1 : 0;
    a->tc_covers[0] = nc == 4;
    a->tc_covers[1] = na == 2;
  }
  a->n_actor_ops = (int)awv.size(); a->n_actor_vops = (int)avv.size(); a->actor_tiles = tiles;
  {
    int rc = upload_tc_tables(a, tc_critic, a->critic_tc_ops, a->critic_tc_map, a->n_critic_tc, a->critic_tc_tiles, tc_maps);
    if (rc == SC_OK) rc = upload_tc_tables(a, tc_actor, a->actor_tc_ops, a->actor_tc_map, a->n_actor_tc, a->actor_tc_tiles, tc_maps);
    if (rc != SC_OK) return rc;
    if (!tc_maps.empty()) {
      CK(cudaMemcpyAsync(a->tmaps_tc, tc_maps.data(), tc_maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, a->st));
      CK(cudaStreamSynchronize(a->st));
    }
  }
  if (tiles > 2048) return fail(SC_ERR_STATE, "actor tile table overflow (%d)", tiles);
  {
    std::vector<int> tt = tile_table(awv, avv, tiles);
    CK(cudaMemcpyAsync(a->actor_tile_map, tt.data(), tt.size() * sizeof(int), cudaMemcpyHostToDevice, a->st));
    CK(cudaStreamSynchronize(a->st));
  }
  CK(cudaMemcpyAsync(a->actor_ops, awv.data(), awv.size() * sizeof(WgradOp), cudaMemcpyHostToDevice, a->st));
  CK(cudaMemcpyAsync(a->actor_vops, avv.data(), avv.size() * sizeof(VecOp), cudaMemcpyHostToDevice, a->st));
  CK(cudaStreamSynchronize(a->st));
  return SC_OK;
}

int sc_agent_create(const sc_config* cfg, void* arena, size_t arena_bytes, void* cuda_stream, sc_agent** out) {
  if (!out) return fail(SC_ERR_ARG, "null out");
  *out = nullptr;
  int rc = check_cfg(cfg);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(SC_ERR_CUDA, "no CUDA device: libsmartcal_b200 has no CPU fallback");
  }
  cudaDeviceProp prop;
  int dev = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 9) return fail(SC_ERR_CUDA, "compute capability %d.%d lacks thread-block clusters (built for sm_100a)", prop.major, prop.minor);
  sc_agent* a = new sc_agent();
  init_dims(a, cfg);
  a->st = (cudaStream_t)cuda_stream;
  a->device = dev;
  a->own = nullptr; a->grp = nullptr; a->gidx = 0; a->cursor = 0;
  a->imm_host = a->imm_dev = nullptr; a->imm_k = 0;
  for (auto& e : a->ev_imm) e = nullptr;
  for (auto& e : a->ev_stage) e = nullptr;
  for (auto& e : a->ev_stage) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : a->ev_imm) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaHostAlloc((void**)&a->imm_host, 4 * IMM_SLOT, cudaHostAllocDefault));
  CK(cudaMalloc((void**)&a->imm_dev, 4 * IMM_SLOT));
  const size_t need = carve(a, nullptr);
  if (arena) {
    if (arena_bytes < need) { delete a; return fail(SC_ERR_ARG, "arena too small: %zu < %zu", arena_bytes, need); }
    if (((uintptr_t)arena) & 255) { delete a; return fail(SC_ERR_ARG, "arena must be 256-byte aligned"); }
    a->arena = arena; a->owns_arena = false;
  } else {
    cudaError_t e = cudaMalloc(&a->arena, need);
    if (e != cudaSuccess) { delete a; return fail(SC_ERR_CUDA, "cudaMalloc(%zu): %s", need, cudaGetErrorString(e)); }
    a->owns_arena = true;
  }
  a->arena_bytes = need;
  a->buffers.clear();
  carve(a, a->arena);
  CK(cudaMemsetAsync(a->arena, 0, need, a->st));
  const size_t stage_floats = (size_t)3 * a->B * a->A + a->T + a->Sp32 + 64;
  CK(cudaHostAlloc((void**)&a->h_stage, stage_floats * sizeof(float), cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&a->h_stats, 32 * sizeof(float), cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&a->h_idx, (size_t)a->B * sizeof(int32_t), cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&a->h_u01, (size_t)a->B * sizeof(double), cudaHostAllocDefault));
  CK(cudaHostAlloc((void**)&a->h_act, 64 * sizeof(float), cudaHostAllocDefault));
  {
    // Row-phase geometry.  Candidates are (variant, cluster size); each gets as many clusters as are co-resident (at most
    // one per 8-row group) and is scored by the padded GEMM work one SM has to issue:
    //     ctas_per_sm * ceil(rows_per_cluster / (8 TM)) * 8 TM * (512 / cluster_size) columns
    // times a penalty for small register tiles (shared-memory-pipe bound) and a bonus for two co-resident CTAs
    // (they overlap each other's latencies).
    int cap[4] = {0, 0, 0, 0};
    rowphase_cluster_capacity(cap);
    const char* force = getenv("SMARTCAL_GEOM");       // "variant,cs" e.g. "1,8" pins a geometry (profiling)
    int fv = 0, fc = 0;
    if (force) sscanf(force, "%d,%d", &fv, &fc);
    double best = 1e30;
    a->cluster_size = 0;
    // measured on B200 (profiles/r1_geometry.md): at B=256 the 2-CTA/SM variant with 32 clusters of 8 rows is the fastest
    if (!force && cap[3] >= (a->B + 7) / 8) {
      best = 0; a->row_variant = 2; a->cluster_size = 8; a->nclusters = (a->B + 7) / 8; a->tile_tm = 1;
    }
    for (int c = 0; c < 4; ++c) {
      const int variant = c < 2 ? 1 : 2, cs = (c & 1) ? 8 : 16;
      if (cap[c] < 1 || (fv && (fv != variant || fc != cs))) continue;
      const int tm_max = variant == 1 ? 4 : 2, tm_min = variant == 1 ? 2 : 1;
      int ncl = cap[c];
      if (ncl > (a->B + 7) / 8) ncl = (a->B + 7) / 8;
      while ((a->B + ncl - 1) / ncl > rowphase_max_rows(variant)) ncl *= 2;   // slab height limit of the LayerNorm staging buffers
      const int per = (a->B + ncl - 1) / ncl;
      for (int tm = tm_min; tm <= tm_max; ++tm) {
        const int padded = ((per + 8 * tm - 1) / (8 * tm)) * 8 * tm;
        const int waves = (ncl + cap[c] - 1) / cap[c];
        const double lds_pen = fmax(1.0, 2.0 * (tm + 8) / (tm * 8.0));      // LDS cycles ~ 2(TM+8) per TM*8 FFMA cycles
        const double cost = (double)variant * waves * padded * (512.0 / cs) * lds_pen * (variant == 2 ? 0.8 : 1.0);
        if (cost < best) { best = cost; a->row_variant = variant; a->cluster_size = cs; a->nclusters = ncl; a->tile_tm = tm; }
      }
    }
    if (getenv("SMARTCAL_DEBUG"))
      fprintf(stderr, "[smartcal] SMs=%d co-resident clusters occ1/16=%d occ1/8=%d occ2/16=%d occ2/8=%d -> variant %d, cluster %d x %d, TM=%d\n",
              prop.multiProcessorCount, cap[0], cap[1], cap[2], cap[3], a->row_variant, a->cluster_size, a->nclusters, a->tile_tm);
  }
  if (a->cluster_size == 0) { sc_agent_destroy(a); return fail(SC_ERR_CUDA, "device cannot host an 8-CTA cluster with %zu B of shared memory", rowphase_smem_bytes(2)); }
  a->wgrad_grid = prop.multiProcessorCount * 2;
  a->sm_count = prop.multiProcessorCount;
  a->mem_cntr = a->learn_counter = a->data_pointer = a->data_length = 0;
  a->t_actor = a->t_critic = 0;
  a->per_beta = 0.4;   // PER.beta, enet_sac.py:210
  a->batch_ready = false;
  a->presampled = false; a->quiet = false; a->presample_beta = 0.0; a->sample_ctr = 0;
  a->split_dirty = true;
  a->alpha = (float)cfg->alpha;
  a->rho = 0.f;
  a->admm_rho = cfg->admm_rho;
  a->noise_ctr = 0;
  a->launches = 0;
  a->timing = false;
  rc = build_op_tables(a);
  if (rc) { sc_agent_destroy(a); return rc; }
  rc = build_tensor_maps(a);
  if (rc) { sc_agent_destroy(a); return rc; }
  a->own = new sc_group();
  sc_agent* one[1] = {a};
  rc = group_init(a->own, one, 1);
  if (rc) { sc_agent_destroy(a); return rc; }
  *out = a;
  return SC_OK;
}

int sc_agent_destroy(sc_agent* a) {
  if (!a) return SC_OK;
  DEV(a);
  cudaStreamSynchronize(a->st);
  for (auto& t : a->timers) {
    for (auto& p : t.pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto& p : t.pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  }
  if (a->own) { group_free(a->own); delete a->own; a->own = nullptr; }
  for (auto& e : a->ev_stage) if (e) cudaEventDestroy(e);
  for (auto& e : a->ev_imm) if (e) cudaEventDestroy(e);
  if (a->imm_host) cudaFreeHost(a->imm_host);
  if (a->imm_dev) cudaFree(a->imm_dev);
  if (a->h_stage) cudaFreeHost(a->h_stage);
  if (a->h_stats) cudaFreeHost(a->h_stats);
  if (a->h_idx) cudaFreeHost(a->h_idx);
  if (a->h_u01) cudaFreeHost(a->h_u01);
  if (a->h_act) cudaFreeHost(a->h_act);
  if (a->owns_arena && a->arena) cudaFree(a->arena);
  delete a;
  return SC_OK;
}

int sc_agent_sync(sc_agent* a) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  DEV(a);
  CK(cudaStreamSynchronize(a->st));
  return SC_OK;
}

static float* net_base(sc_agent* a, int net, int slot, const NetLayout** L) {
  const bool td3 = a->cfg.kind == SC_KIND_TD3;
  float* online = slot == SC_SLOT_PARAM ? a->params : slot == SC_SLOT_ADAM_M ? a->adam_m : a->adam_v;
  switch (net) {
    case SC_NET_ACTOR: *L = &a->La; return online + a->off_actor;
    case SC_NET_CRITIC_1: *L = &a->Lc; return online + a->off_c1;
    case SC_NET_CRITIC_2: *L = &a->Lc; return online + a->off_c2;
    case SC_NET_TARGET_CRITIC_1: *L = &a->Lc; return slot == SC_SLOT_PARAM ? a->targets + a->toff_c1 : nullptr;
    case SC_NET_TARGET_CRITIC_2: *L = &a->Lc; return slot == SC_SLOT_PARAM ? a->targets + a->toff_c2 : nullptr;
    case SC_NET_TARGET_ACTOR: *L = &a->La; return (td3 && slot == SC_SLOT_PARAM) ? a->targets + a->toff_actor : nullptr;
  }
  return nullptr;
}

int sc_agent_tensor(sc_agent* a, int net, int slot, const char* name, sc_tensor_info* out) {
  if (!a || !name || !out) return fail(SC_ERR_ARG, "null argument");
  if (slot < 0 || slot > 2) return fail(SC_ERR_ARG, "bad slot %d", slot);
  const NetLayout* L = nullptr;
  float* base = net_base(a, net, slot, &L);
  if (!base) return fail(SC_ERR_ARG, "network %d / slot %d does not exist for this agent kind", net, slot);
  if (slot == SC_SLOT_PARAM) a->split_dirty = true;    // the caller may write through this pointer: re-derive the tf32 operand copies before the next step
  if (strcmp(name, "*") == 0) {
    out->dev_ptr = base; out->rows = 1; out->cols = (int64_t)L->size; out->pitch = (int64_t)L->size; out->elem_bytes = 4; out->is_int = 0;
    return SC_OK;
  }
  const TDesc* d = L->find(name);
  if (!d) return fail(SC_ERR_ARG, "unknown tensor '%s'", name);
  out->dev_ptr = base + d->off; out->rows = d->rows; out->cols = d->cols; out->pitch = d->pitch; out->elem_bytes = 4; out->is_int = 0;
  return SC_OK;
}

int sc_agent_buffer(sc_agent* a, const char* name, sc_tensor_info* out) {
  if (!a || !name || !out) return fail(SC_ERR_ARG, "null argument");
  if (strcmp(name, "rho") != 0) touch_replay(a);   // the caller may write the tree / replay arrays through the returned pointer
  auto it = a->buffers.find(name);
  if (it == a->buffers.end()) return fail(SC_ERR_ARG, "unknown buffer '%s'", name);
  *out = it->second;
  return SC_OK;
}

int sc_agent_get_counters(sc_agent* a, int64_t* mem_cntr, int64_t* learn_counter, int64_t* data_pointer, double* per_beta) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  if (mem_cntr) *mem_cntr = a->mem_cntr;
  if (learn_counter) *learn_counter = a->learn_counter;
  if (data_pointer) *data_pointer = a->data_pointer;
  if (per_beta) *per_beta = a->per_beta;
  return SC_OK;
}
int sc_agent_set_counters(sc_agent* a, int64_t mem_cntr, int64_t learn_counter, int64_t data_pointer, double per_beta) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  touch_replay(a);
  a->mem_cntr = mem_cntr; a->learn_counter = learn_counter; a->data_pointer = data_pointer; a->per_beta = per_beta;
  a->data_length = mem_cntr < a->C ? mem_cntr : a->C;
  return SC_OK;
}
int sc_agent_set_scalars(sc_agent* a, float alpha, float rho, float admm_rho) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  DEV(a);
  a->alpha = alpha; a->rho = rho; a->admm_rho = admm_rho;
  CK(cudaStreamSynchronize(a->st));
  CK(cudaMemcpy(a->rho_dev, &rho, sizeof(float), cudaMemcpyHostToDevice));   // rho evolves on the device
  return SC_OK;
}

// ---- immediate parameter slots (stand-alone entry points outside learn()) ------------------------------------------------
struct ImmSlot { char* host; char* dev; int k; };
static int imm_begin(sc_agent* a, ImmSlot& s) {
  s.k = a->imm_k;
  a->imm_k = (a->imm_k + 1) & 3;
  CK(cudaEventSynchronize(a->ev_imm[s.k]));                  // the copy that read this pinned slot last (normally long complete)
  s.host = a->imm_host + (size_t)s.k * IMM_SLOT;
  s.dev = a->imm_dev + (size_t)s.k * IMM_SLOT;
  return SC_OK;
}
static int imm_commit(sc_agent* a, const ImmSlot& s, size_t bytes) {
  CK(cudaMemcpyAsync(s.dev, s.host, bytes, cudaMemcpyHostToDevice, a->st));
  CK(cudaEventRecord(a->ev_imm[s.k], a->st));
  return SC_OK;
}

// ---- replay ----------------------------------------------------------------------------------------------
// slot bookkeeping of ReplayBuffer.store_transition (enet_sac.py:35-44) / PER.store_transition (:237-268)
static int next_slot(sc_agent* a, int& priority_mode) {
  int slot;
  if (a->cfg.prioritized) {
    if (priority_mode == 0) priority_mode = 1;
    slot = (int)a->data_pointer;
    a->data_pointer = (a->data_pointer + 1) % a->C;
    if (a->data_length < a->C) a->data_length++;
  } else {
    priority_mode = 0;
    slot = (int)(a->mem_cntr % a->C);
  }
  a->mem_cntr++;
  return slot;
}
static int store_common(sc_agent* a, const float* tr_dev, int priority_mode, double priority) {
  const int slot = next_slot(a, priority_mode);
  CK(launch_replay_store(a->rp, tr_dev, slot, priority_mode, priority, (double)a->cfg.per_clip, a->st));
  a->launches++;
  return SC_OK;
}
int sc_replay_store_host(sc_agent* a, const float* transition_host, int priority_mode, double priority) {
  if (!a || !transition_host) return fail(SC_ERR_ARG, "null argument");
  DEV(a);
  touch_replay(a);
  // the pinned slot is reused by the next call: wait for the copy that read it last (normally long complete)
  float* h = a->h_stage;
  CK(cudaEventSynchronize(a->ev_stage[3]));
  memcpy(h, transition_host, (size_t)a->T * sizeof(float));
  CK(cudaMemcpyAsync(a->store_stage, h, (size_t)a->T * sizeof(float), cudaMemcpyHostToDevice, a->st));
  CK(cudaEventRecord(a->ev_stage[3], a->st));
  return store_common(a, a->store_stage, priority_mode, priority);
}
int sc_replay_store_dev(sc_agent* a, const float* transitions_dev, int index, int priority_mode, double priority) {
  if (!a || !transitions_dev || index < 0) return fail(SC_ERR_ARG, "bad argument");
  DEV(a);
  touch_replay(a);
  return store_common(a, transitions_dev + (size_t)index * a->T, priority_mode, priority);
}
int sc_replay_fill_host(sc_agent* a, const float* transitions_host, int n, const double* priorities_host) {
  if (!a || !transitions_host || n < 0 || n > a->C) return fail(SC_ERR_ARG, "bad argument");
  DEV(a);
  touch_replay(a);
  float* tmp = nullptr;
  CK(cudaMalloc(&tmp, (size_t)n * a->T * sizeof(float)));
  CK(cudaMemcpyAsync(tmp, transitions_host, (size_t)n * a->T * sizeof(float), cudaMemcpyHostToDevice, a->st));
  CK(launch_replay_fill(a->rp, tmp, n, a->st));
  a->launches++;
  if (a->rp.tree) {
    double* pr = nullptr;
    CK(cudaMalloc(&pr, (size_t)a->C * sizeof(double)));
    std::vector<double> h(a->C, 0.0);
    for (int i = 0; i < n; ++i) h[i] = priorities_host ? priorities_host[i] : (double)a->cfg.per_clip;
    CK(cudaMemcpyAsync(pr, h.data(), (size_t)a->C * sizeof(double), cudaMemcpyHostToDevice, a->st));
    CK(launch_tree_build(a->rp.tree, a->C, pr, n, a->st));
    a->launches++;
    CK(cudaStreamSynchronize(a->st));
    cudaFree(pr);
  }
  CK(cudaStreamSynchronize(a->st));
  cudaFree(tmp);
  a->mem_cntr = n;
  a->data_pointer = n % a->C;
  a->data_length = n;
  return SC_OK;
}

int sc_replay_set_batch_host(sc_agent* a, const int32_t* idx_host, int n) {
  if (!a || !idx_host || n != a->B) return fail(SC_ERR_ARG, "expected %d indices", a ? a->B : 0);
  DEV(a);
  touch_replay(a);
  CK(cudaEventSynchronize(a->ev_stage[5]));   // h_idx may still be in flight from the previous step
  memcpy(a->h_idx, idx_host, (size_t)n * sizeof(int32_t));
  CK(cudaMemcpyAsync(a->bidx, a->h_idx, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, a->st));
  CK(cudaEventRecord(a->ev_stage[5], a->st));
  a->batch_ready = true;
  return SC_OK;
}
static UniformP uniform_params(sc_agent* a, int n) {
  UniformP u;
  u.n = n; u.B = a->B; u.seed = a->cfg.rng_seed; u.ctr = 0x5000000000ull + (a->sample_ctr++); u.out = a->bidx;
  return u;
}
int sc_replay_sample_uniform_dev(sc_agent* a) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  DEV(a);
  touch_replay(a);
  const int n = (int)(a->mem_cntr < a->C ? a->mem_cntr : a->C);
  if (n < a->B) return fail(SC_ERR_STATE, "only %d transitions stored, batch is %d", n, a->B);
  ImmSlot s;
  int rc = imm_begin(a, s);
  if (rc) return rc;
  *reinterpret_cast<UniformP*>(s.host) = uniform_params(a, n);
  if ((rc = imm_commit(a, s, sizeof(UniformP)))) return rc;
  CK(launch_sample_uniform(reinterpret_cast<const UniformP*>(s.dev), 1, n, a->st));
  a->launches++;
  a->batch_ready = true;
  return SC_OK;
}
static PerSampleP per_sample_params(sc_agent* a, const double* u01_dev, double beta, bool alt) {
  PerSampleP p;
  p.tree = a->rp.tree; p.C = a->C; p.B = a->B; p.u01 = u01_dev;
  p.seed = a->cfg.rng_seed; p.ctr = 0x6000000000ull + (a->sample_ctr++);
  p.beta = beta;
  p.tree_idx = alt ? a->tree_idx_alt : a->tree_idx;
  p.data_idx = alt ? a->bidx_alt : a->bidx;
  p.is_w = alt ? a->is_w_alt : a->is_w;
  p.is_wsum = alt ? a->is_wsum_alt : a->is_wsum;
  return p;
}
int sc_per_sample(sc_agent* a, const double* u01_host, int32_t* tree_idx_host_out, float* is_weights_host_out) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  if (!a->rp.tree) return fail(SC_ERR_STATE, "agent was created with prioritized=0");
  DEV(a);
  touch_replay(a);
  if (sizeof(PerSampleP) + 16 + (size_t)a->B * sizeof(double) > IMM_SLOT) return fail(SC_ERR_UNSUPPORTED, "batch too large for the immediate slot");
  ImmSlot s;
  int rc = imm_begin(a, s);
  if (rc) return rc;
  const size_t uoff = (sizeof(PerSampleP) + 15) & ~(size_t)15;
  size_t bytes = sizeof(PerSampleP);
  const double* u = nullptr;
  if (u01_host) {
    memcpy(s.host + uoff, u01_host, (size_t)a->B * sizeof(double));
    u = reinterpret_cast<const double*>(s.dev + uoff);
    bytes = uoff + (size_t)a->B * sizeof(double);
  }
  a->per_beta = fmin(1.0, a->per_beta + 1e-4);   // enet_sac.py:288
  *reinterpret_cast<PerSampleP*>(s.host) = per_sample_params(a, u, a->per_beta, false);
  if ((rc = imm_commit(a, s, bytes))) return rc;
  cudaEvent_t e0, e1;
  timer_begin(a, 4, &e0, &e1);
  CK(launch_per_sample(reinterpret_cast<const PerSampleP*>(s.dev), 1, a->B, a->st));
  timer_end(a, 4, e0, e1);
  a->launches++;
  a->batch_ready = true;
  if (tree_idx_host_out) CK(cudaMemcpyAsync(tree_idx_host_out, a->tree_idx, (size_t)a->B * sizeof(int32_t), cudaMemcpyDeviceToHost, a->st));
  if (is_weights_host_out) CK(cudaMemcpyAsync(is_weights_host_out, a->is_w, (size_t)a->B * sizeof(float), cudaMemcpyDeviceToHost, a->st));
  if (tree_idx_host_out || is_weights_host_out) CK(cudaStreamSynchronize(a->st));
  return SC_OK;
}
static PerUpdateP per_update_params(sc_agent* a, float* ps_out, const float* ps_in) {
  PerUpdateP p;
  p.tree = a->rp.tree; p.C = a->C; p.B = a->B; p.idx = a->tree_idx; p.err = a->err; p.clip = a->cfg.per_clip;
  p.ps_out = ps_out; p.ps_in = ps_in;
  return p;
}
int sc_per_update_host(sc_agent* a, const int32_t* tree_idx_host, const float* abs_err_host, int n, float* priorities_host_out) {
  if (!a || !tree_idx_host || !abs_err_host || n != a->B) return fail(SC_ERR_ARG, "bad argument");
  if (!a->rp.tree) return fail(SC_ERR_STATE, "agent was created with prioritized=0");
  DEV(a);
  touch_replay(a);
  CK(cudaMemcpyAsync(a->tree_idx, tree_idx_host, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, a->st));
  CK(cudaMemcpyAsync(a->err, abs_err_host, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, a->st));
  ImmSlot s;
  int rc = imm_begin(a, s);
  if (rc) return rc;
  *reinterpret_cast<PerUpdateP*>(s.host) = per_update_params(a, a->ps_out, nullptr);
  if ((rc = imm_commit(a, s, sizeof(PerUpdateP)))) return rc;
  CK(launch_per_update(reinterpret_cast<const PerUpdateP*>(s.dev), 1, a->B, a->st));
  a->launches++;
  if (priorities_host_out) CK(cudaMemcpyAsync(priorities_host_out, a->ps_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, a->st));
  CK(cudaStreamSynchronize(a->st));
  return SC_OK;
}

// ---- kernel parameter assembly -----------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static void fill_common(sc_agent* a, RowPhaseParams& p) {
  static const bool stamps = getenv("SMARTCAL_STAMPS") != nullptr;
  static const int dbg = env_int("SMARTCAL_DBG", 0);
  memset(&p, 0, sizeof(p));
  p.kind = a->cfg.kind;
  p.closs_scale = (a->cfg.flags & SC_FLAG_CRITIC_LOSS_SUM) ? (float)a->B : 1.f;
  p.B = a->B; p.S = a->S; p.Sp = a->Sp; p.A = a->A; p.Ap = a->Ap;
  p.crit[0] = critic_ptrs(a->params + a->off_c1, a->Lc);
  p.crit[1] = critic_ptrs(a->params + a->off_c2, a->Lc);
  p.targ[0] = critic_ptrs(a->targets + a->toff_c1, a->Lc);
  p.targ[1] = critic_ptrs(a->targets + a->toff_c2, a->Lc);
  p.cw[0] = a->cw[0]; p.cw[1] = a->cw[1];
  p.tw[0] = a->tw[0]; p.tw[1] = a->tw[1];
  p.aw = a->aw;
  p.rp = a->rp;
  p.bidx = a->bidx_step;
  p.is_wsum = a->cfg.prioritized ? a->is_wsum : nullptr;
  p.y = a->y; p.err = a->err; p.part = a->part;
  p.per_ps = a->cfg.prioritized ? a->ps_out : nullptr;
  p.per_clip = a->cfg.per_clip;
  p.gamma = (float)a->cfg.gamma;
  p.scale = (float)a->cfg.reward_scale;
  p.alpha = a->alpha;
  p.maxa = a->cfg.max_action;
  p.n_q = a->cfg.kind == SC_KIND_SAC ? 2 : 1;
  p.mb_s = a->mb_s; p.mb_s2 = a->mb_s2; p.Sp32 = a->Sp32;
  p.tstamp = stamps ? a->tstamp : nullptr;
  p.dbg = dbg;
}
static AdamHyper adam_hyper(double lr, int64_t t) {
  // torch/optim/adam.py::_single_tensor_adam scalar math, in double like Python floats
  const double b1 = 0.9, b2 = 0.999;
  const double bc1 = 1.0 - pow(b1, (double)t), bc2 = 1.0 - pow(b2, (double)t);
  AdamHyper h;
  h.lr_neg_step = (float)(-(lr / bc1));
  h.bc2_sqrt = (float)sqrt(bc2);
  h.one_minus_b1 = (float)(1.0 - b1);
  h.b2 = (float)b2;
  h.one_minus_b2 = (float)(1.0 - b2);
  h.eps = 1e-8f;
  return h;
}
static void fill_wgrad(sc_agent* a, WgradParams& w, bool actor, int do_soft, int stats_mode) {
  memset(&w, 0, sizeof(w));
  w.ops = actor ? a->actor_ops : a->critic_ops;
  w.nops = actor ? a->n_actor_ops : a->n_critic_ops;
  w.vops = actor ? a->actor_vops : a->critic_vops;
  w.nvops = actor ? a->n_actor_vops : a->n_critic_vops;
  w.total_tiles = actor ? a->actor_tiles : a->critic_tiles;
  w.tile_map = actor ? a->actor_tile_map : a->critic_tile_map;
  w.B = a->B;
  w.bidx = a->bidx_step;
  w.hyper[0] = adam_hyper(a->cfg.lr_actor, a->t_actor);
  w.hyper[1] = adam_hyper(a->cfg.lr_critic, a->t_critic);
  w.hyper[2] = w.hyper[1];
  w.tau = (float)a->cfg.tau;
  w.one_minus_tau = (float)(1.0 - a->cfg.tau);
  w.do_soft = do_soft;
  w.stats_mode = stats_mode;
  w.q1 = a->cw[0].q; w.q2 = a->cw[1].q; w.y = a->y;
  w.logp = a->cfg.kind == SC_KIND_SAC ? a->aw.logp : nullptr;
  w.pi_q = a->aw.qmin;
  w.is_wsum = a->cfg.prioritized ? a->is_wsum : nullptr;
  w.stats = a->grp->stats_dev + (size_t)a->gidx * 32;
  w.alpha = a->cfg.kind == SC_KIND_SAC ? a->alpha : 0.f;
}

// ---- plan records ------------------------------------------------------------------------------------------------
static int plan_fail(sc_group* g, const char* what) {
  g->plan_error = 1;
  return fail(SC_ERR_STATE, "ensemble agents diverged while planning a step (%s): all agents of an ensemble must be in lock-step "
                            "(same configuration, counters and call sequence)", what);
}
// appends (agent 0) or re-walks (agents 1..E-1) one record; returns its index or -1
static int plan_rec(sc_agent* a, int type, int kid, int stream, const int (&geo)[5], size_t bytes) {
  sc_group* g = a->grp;
  if (g->plan_error) return -1;
  if (a->gidx == 0) {
    Rec r;
    r.type = type; r.kid = kid; r.stream = stream;
    for (int i = 0; i < 5; ++i) r.g[i] = geo[i];
    r.bytes = bytes;
    g->used = (g->used + 255) & ~(size_t)255;
    r.off = g->used;
    g->used += bytes * (size_t)g->E;
    if (g->used > g->tape_cap) { g->plan_error = 1; fail(SC_ERR_STATE, "step tape overflow (%zu > %zu bytes)", g->used, g->tape_cap); return -1; }
    g->recs.push_back(r);
    if (type == REC_KERNEL) { a->launches++; g->launches++; }
  } else {
    if (a->cursor >= g->recs.size()) { plan_fail(g, "extra record"); return -1; }
    const Rec& r = g->recs[a->cursor];
    bool same = r.type == type && r.kid == kid && r.stream == stream && r.bytes == bytes;
    for (int i = 0; i < 5; ++i) same = same && r.g[i] == geo[i];
    if (!same) { plan_fail(g, "different record"); return -1; }
  }
  return (int)(a->cursor++);
}
extern "C++" {
template <class P>
static P* emit_kernel(sc_agent* a, int kid, int stream, int g0 = 0, int g1 = 0, int g2 = 0, int g3 = 0, int g4 = 0) {
  static_assert(sizeof(P) % 8 == 0, "parameter blocks are arrays in the tape");
  const int geo[5] = {g0, g1, g2, g3, g4};
  const int i = plan_rec(a, REC_KERNEL, kid, stream, geo, sizeof(P));
  if (i < 0) return nullptr;
  sc_group* g = a->grp;
  return reinterpret_cast<P*>(g->tape_host[g->tk] + g->recs[i].off + (size_t)a->gidx * sizeof(P));
}
}  // extern "C++"
// host-supplied array of this step (noise, uniforms, indices, a transition): lands in the tape, *dev is where kernels read it
static void* emit_data(sc_agent* a, size_t bytes, const void** dev) {
  const int geo[5] = {0, 0, 0, 0, 0};
  const size_t b = (bytes + 15) & ~(size_t)15;
  const int i = plan_rec(a, REC_DATA, 0, 0, geo, b);
  if (i < 0) return nullptr;
  sc_group* g = a->grp;
  const size_t off = g->recs[i].off + (size_t)a->gidx * b;
  *dev = g->tape_dev[g->tk] + off;
  return g->tape_host[g->tk] + off;
}
static int emit_ctl(sc_agent* a, int type, int which) {
  const int geo[5] = {which, 0, 0, 0, 0};
  return plan_rec(a, type, 0, 0, geo, 0) < 0 ? SC_ERR_STATE : SC_OK;
}
#define PL(ptr) do { if ((ptr) == nullptr) return SC_ERR_STATE; } while (0)

// tf32 big / small copies written by k_split: the paths whose producer kernel does not write them itself (fallback kernels
// selected by SMARTCAL_DENSE / _TAIL / _DGRAD / _WGRAD, parameters written from outside)
struct SplitItem { const float* src; float* hl; size_t count, plane; };
static int plan_split(sc_agent* a, const std::vector<SplitItem>& items) {
  if (items.empty()) return SC_OK;
  if (items.size() > 12) return fail(SC_ERR_STATE, "split list overflow");
  size_t total = 0;
  for (const SplitItem& it : items) total += it.count;
  int grid = (int)((total / 4 + 1023) / 1024);
  if (grid < 1) grid = 1;
  if (grid > 2 * a->sm_count) grid = 2 * a->sm_count;
  SplitP* sp = emit_kernel<SplitP>(a, KID_SPLIT, 0, grid);
  PL(sp);
  memset(sp, 0, sizeof(SplitP));
  sp->n = (int)items.size();
  for (size_t i = 0; i < items.size(); ++i) {
    sp->src[i] = items[i].src; sp->hl[i] = items[i].hl; sp->count[i] = (unsigned)items[i].count; sp->plane[i] = (unsigned)items[i].plane;
  }
  return SC_OK;
}
static int plan_split_weights(sc_agent* a) {
  return plan_split(a, {SplitItem{a->params, a->params_hl, a->online_floats, a->online_floats},
                        SplitItem{a->targets, a->targets_hl, a->target_floats, a->target_floats}});
}
static void split_act(const sc_agent* a, std::vector<SplitItem>& v, const float* fp32, int width) {
  v.push_back(SplitItem{fp32, a->hl_of.at(fp32), (size_t)a->B * width, (size_t)a->B * width});
}

// The two wide forward layers of a batched phase as tcgen05 launches (sc_dense_tc.cu); the op lists mirror the dense(...)
// tables of k_rowphase (sc_rowphase_impl.cuh, PROG_CRITIC / PROG_ACTOR).
static int plan_dense_tc(sc_agent* a, const RowPhaseParams& p) {
  static const bool stamps = getenv("SMARTCAL_STAMPS") != nullptr;
  const bool critic = p.program == PROG_CRITIC;
  const TensorMaps& m = critic ? a->tmd_critic : a->tmd_actor;
  const ActorP& ac = p.actor;
  DenseTcParams d1, d2;
  memset(&d1, 0, sizeof(d1));
  memset(&d2, 0, sizeof(d2));
  d1.N = H1; d2.N = H2;
  d1.B = d2.B = p.B;
  d1.nrb = d2.nrb = (p.B + 127) / 128;
  const int K1 = p.Sp32;
  int n = 0;
  // (the layer-1 input is the gathered minibatch, which k_gather_mb also wrote as tf32 big / small parts)
  auto op = [&](const void* mx1, const void* mw1, const float* b1, const float* g1, const float* be1, float* h1, float* zh1, float* rs1,
                const void* mx2, const void* mw2, const float* b2, const float* g2, const float* be2, float* h2, float* zh2, float* rs2) {
    const bool next = mx1 == m.mb_s2;
    // (tmd_s*_hi / _lo are consecutive table entries, like every big / small pair)
    d1.ops[n] = DenseTcOp{next ? a->tmd_s2_hi : a->tmd_s_hi, mw1, b1, g1, be1, h1, zh1, rs1, K1, nullptr, nullptr, a->hl_of.at(h1)};
    d2.ops[n] = DenseTcOp{mx2, mw2, b2, g2, be2, h2, zh2, rs2, H1, nullptr, nullptr, a->hl_of.count(h2) ? a->hl_of.at(h2) : nullptr};
    ++n;
  };
  if (critic) {
    GatherP* gp = emit_kernel<GatherP>(a, KID_GATHER, 0, a->B);
    PL(gp);
    gp->mb_s = a->mb_s; gp->mb_s2 = a->mb_s2; gp->hl = a->mb_hl; gp->state = p.rp.state; gp->new_state = p.rp.new_state;
    gp->bidx = p.bidx; gp->bidx_mirror = (p.bidx != a->bidx) ? a->bidx : nullptr;
    gp->B = p.B; gp->Sp = p.Sp; gp->Sp32 = p.Sp32;
    op(m.mb_s2, m.actor_w1, ac.b1, ac.g1, ac.be1, p.aw.t_h1, nullptr, nullptr,
       m.t_h1, m.actor_w2, ac.b2, ac.g2, ac.be2, p.aw.t_h2, nullptr, nullptr);
    for (int k = 0; k < 2; ++k)
      op(m.mb_s2, m.targ_w11[k], p.targ[k].b11, p.targ[k].g11, p.targ[k].be11, p.tw[k].h11, nullptr, nullptr,
         m.th11[k], m.targ_w12[k], p.targ[k].b12, p.targ[k].g12, p.targ[k].be12, p.tw[k].h12, nullptr, nullptr);
    for (int k = 0; k < 2; ++k)
      op(m.mb_s, m.crit_w11[k], p.crit[k].b11, p.crit[k].g11, p.crit[k].be11, p.cw[k].h11, p.cw[k].zh11, p.cw[k].rs11,
         m.ch11[k], m.crit_w12[k], p.crit[k].b12, p.crit[k].g12, p.crit[k].be12, p.cw[k].h12, p.cw[k].zh12, p.cw[k].rs12);
  } else {
    op(m.mb_s, m.actor_w1, ac.b1, ac.g1, ac.be1, p.aw.h1, p.aw.zh1, p.aw.rs1,
       m.h1, m.actor_w2, ac.b2, ac.g2, ac.be2, p.aw.h2, p.aw.zh2, p.aw.rs2);
    for (int k = 0; k < p.n_q; ++k)
      op(m.mb_s, m.crit_w11[k], p.crit[k].b11, p.crit[k].g11, p.crit[k].be11, p.cw[k].p_h11, nullptr, nullptr,
         m.ph11[k], m.crit_w12[k], p.crit[k].b12, p.crit[k].g12, p.crit[k].be12, p.cw[k].p_h12, nullptr, nullptr);
  }
  d1.nops = d2.nops = n;
  // reduction split: the K loop of one CTA is the latency of the launch, so K is halved across CTA pairs wherever the grid
  // of ONE agent still fits the machine in one wave (SMARTCAL_DENSE_KS=a,b pins the two stages' splits).  The split changes
  // the summation order, so it is decided per agent, never by the ensemble size: an agent's bits do not depend on E.
  d1.ks = 1; d2.ks = 2;
  d2.late_trigger = 1;
  if (n * d1.nrb * (H1 / 64) * 2 <= a->sm_count && (K1 / 32) % 2 == 0 && n * d1.nrb <= a->dense_cap[2]) d1.ks = 2;   // 16-CTA clusters
  if (n * d2.nrb * (H2 / 64) * 2 > a->sm_count) d2.ks = 1;
  if (const char* e = getenv("SMARTCAL_DENSE_KS")) {
    int k1 = 0, k2 = 0;
    if (sscanf(e, "%d,%d", &k1, &k2) == 2 && k1 >= 1 && k1 <= 2 && k2 >= 1 && k2 <= 2 && (K1 / 32) % k1 == 0 && (k1 == 1 || a->dense_cap[2] >= 1)) {
      d1.ks = k1; d2.ks = k2;
    }
  }
  if (stamps) {
    if (critic) d1.tl = a->tstamp + 2880 * 8;
    d1.ts = a->tstamp + (critic ? 1280 : 1792) * 8;
    d2.ts = a->tstamp + (critic ? 1408 : 1920) * 8;
  }
  DenseTcParams* t1 = emit_kernel<DenseTcParams>(a, KID_DENSE, 0, d1.nops, d1.nrb, d1.N, d1.ks);
  PL(t1);
  *t1 = d1;
  DenseTcParams* t2 = emit_kernel<DenseTcParams>(a, KID_DENSE, 0, d2.nops, d2.nrb, d2.N, d2.ks);
  PL(t2);
  *t2 = d2;
  return SC_OK;
}

static int plan_rowphase(sc_agent* a, const RowPhaseParams& p_in, int timer_id) {
  int rc;
  if (timer_id >= 0 && a->timing && (rc = emit_ctl(a, REC_TBEGIN, timer_id))) return rc;
  RowPhaseParams p = p_in;
  if (a->dense_tc) {
    const int dense_timer = timer_id == 0 ? 6 : timer_id == 2 ? 7 : -1;
    if (dense_timer >= 0 && a->timing && (rc = emit_ctl(a, REC_TBEGIN, dense_timer))) return rc;
    if ((rc = plan_dense_tc(a, p))) return rc;
    if (dense_timer >= 0 && a->timing && (rc = emit_ctl(a, REC_TEND, dense_timer))) return rc;
    p.skip_dense = 1;
    // the row-local tail as its own launch (sc_tail.cu; SMARTCAL_TAIL=row keeps it inside k_rowphase for A/B runs)
    static const bool tail_row = getenv("SMARTCAL_TAIL") != nullptr && !strcmp(getenv("SMARTCAL_TAIL"), "row");
    if (!tail_row) {
      RowPhaseParams* tt = emit_kernel<RowPhaseParams>(a, KID_TAIL, 0, a->B);
      PL(tt);
      *tt = p;
      p.skip_tail = 1;
      if (a->dgrad_tc) {
        // the wide input gradient (dz12 -> du11 / dz11 of both critics, or dz2 -> du1 / dz1 of the actor) on tcgen05
        DenseTcParams g;
        memset(&g, 0, sizeof(g));
        g.N = H1; g.B = p.B; g.nrb = (p.B + 127) / 128; g.ks = 1; g.late_trigger = 0;
        {
          // reduction split across CTA pairs (16-CTA clusters) where ONE agent's clusters are co-resident: the K loop is the latency
          // of this launch and only 16 / 32 CTAs would run it otherwise (decided per agent: bits do not depend on the ensemble size)
          const int nclusters = (p.program == PROG_CRITIC ? 2 : 1) * g.nrb;
          static const bool ks1 = getenv("SMARTCAL_DGRAD_KS") != nullptr && getenv("SMARTCAL_DGRAD_KS")[0] == '1';
          if (!ks1 && nclusters <= a->dense_cap[2] && nclusters * 16 <= a->sm_count) g.ks = 2;
        }
        if (p.program == PROG_CRITIC) {
          for (int k = 0; k < 2; ++k) {
            g.ops[k] = DenseTcOp{a->tmg_dz12[k], a->tmg_w12p[k], nullptr, p.crit[k].g11, nullptr, p.cw[k].h11, p.cw[k].zh11, p.cw[k].rs11, H2,
                                 p.cw[k].du11, p.cw[k].dz11, p.cw[k].dz11_hl};
          }
          g.nops = 2;
        } else {
          g.ops[0] = DenseTcOp{a->tmg_dz2, a->tmg_w2p, nullptr, p.actor.g1, nullptr, p.aw.h1, p.aw.zh1, p.aw.rs1, H2, p.aw.du1, p.aw.dz1, p.aw.dz1_hl};
          g.nops = 1;
        }
        DenseTcParams* gp = emit_kernel<DenseTcParams>(a, KID_DGRAD, 0, g.nops, g.nrb, g.N, g.ks);
        PL(gp);
        *gp = g;
        if (timer_id >= 0 && a->timing && (rc = emit_ctl(a, REC_TEND, timer_id))) return rc;
        return SC_OK;
      }
    }
  }
  RowPhaseParams* t = emit_kernel<RowPhaseParams>(a, KID_ROW, 0, a->row_variant, a->cluster_size, a->nclusters, a->tile_tm);
  PL(t);
  *t = p;
  if (a->use_tc && p.program != PROG_ACT) {
    // k_rowphase writes fp32 only: the tcgen05 weight-gradient kernel reads big / small parts
    std::vector<SplitItem> v;
    const bool critic = p.program == PROG_CRITIC;
    if (!p.skip_dense) {
      v.push_back(SplitItem{a->mb_s, a->mb_hl, (size_t)a->B * a->Sp32, (size_t)a->B * a->Sp32});
      if (critic) { split_act(a, v, a->cw[0].h11, H1); split_act(a, v, a->cw[1].h11, H1); }
      else { split_act(a, v, a->aw.h1, H1); split_act(a, v, a->aw.h2, H2); }
    }
    if (!p.skip_tail) {
      if (critic) { split_act(a, v, a->cw[0].dz12, H2); split_act(a, v, a->cw[1].dz12, H2); }
      else { split_act(a, v, a->aw.dz2, H2); split_act(a, v, a->aw.dz3, H3); }
    }
    if (critic) { split_act(a, v, a->cw[0].dz11, H1); split_act(a, v, a->cw[1].dz11, H1); }
    else split_act(a, v, a->aw.dz1, H1);
    if ((rc = plan_split(a, v))) return rc;
  }
  if (timer_id >= 0 && a->timing && (rc = emit_ctl(a, REC_TEND, timer_id))) return rc;
  return SC_OK;
}
// per-agent grid width of the grid-stride tile kernels: a stand-alone agent gets the whole machine, an ensemble shares it
static int shared_grid(const sc_agent* a, int full, int tiles) {
  const int E = a->grp->E;
  int gx = (full + E - 1) / E;
  if (gx < 8) gx = 8;
  if (gx > tiles) gx = tiles;
  return gx < 1 ? 1 : gx;
}
static int plan_wgrad(sc_agent* a, const WgradParams& w, int timer_id) {
  static const bool stamps = getenv("SMARTCAL_STAMPS") != nullptr;
  int rc;
  if (a->timing && (rc = emit_ctl(a, REC_TBEGIN, timer_id))) return rc;
  const bool actor = w.ops == a->actor_ops;
  const int tc_tiles = actor ? a->actor_tc_tiles : a->critic_tc_tiles;
  const int grid = shared_grid(a, a->wgrad_grid, w.total_tiles + 1);
  if (tc_tiles > 0) {
    WgradTcParams t;
    memset(&t, 0, sizeof(t));
    t.ops = actor ? a->actor_tc_ops : a->critic_tc_ops;
    t.nops = actor ? a->n_actor_tc : a->n_critic_tc;
    t.total_tiles = tc_tiles;
    t.tile_map = actor ? a->actor_tc_map : a->critic_tc_map;
    t.B = a->B;
    for (int i = 0; i < 3; ++i) t.hyper[i] = w.hyper[i];
    t.tau = w.tau; t.one_minus_tau = w.one_minus_tau; t.do_soft = w.do_soft;
    t.tl = (stamps && !actor) ? a->tstamp + 2304 * 8 : nullptr;
    // the wide layers (tensor-core kernel, <= 88 CTAs of one SM each) and the narrow layers + vector parameters + loss
    // readout (FP32 tile kernel) are independent: the second runs on the side stream next to the first
    const bool fork = !a->timing;
    if (fork && (rc = emit_ctl(a, REC_FORK, 2))) return rc;
    if (fork) {
      WgradParams* wp = emit_kernel<WgradParams>(a, KID_WGRAD, 2, grid);
      PL(wp);
      *wp = w;
      if ((rc = emit_ctl(a, REC_JOIN_RECORD, 2))) return rc;
    }
    WgradTcParams* tp = emit_kernel<WgradTcParams>(a, KID_WGRAD_TC, 0, shared_grid(a, a->sm_count, tc_tiles));
    PL(tp);
    *tp = t;
    if (fork) {
      if ((rc = emit_ctl(a, REC_JOIN_WAIT, 2))) return rc;
    } else {
      WgradParams* wp = emit_kernel<WgradParams>(a, KID_WGRAD, 0, grid);
      PL(wp);
      *wp = w;
    }
  } else {
    WgradParams* wp = emit_kernel<WgradParams>(a, KID_WGRAD, 0, grid);
    PL(wp);
    *wp = w;
  }
  // wide layers updated by the FP32 kernel (SMARTCAL_WGRAD=ffma, or a layer too small for the tensor-core tiles): refresh the operand copies
  if (a->dense_tc && !a->tc_covers[actor ? 1 : 0] && (rc = plan_split_weights(a))) return rc;
  if (a->timing && (rc = emit_ctl(a, REC_TEND, timer_id))) return rc;
  return SC_OK;
}

int sc_agent_act(sc_agent* a, const float* obs_host, const float* eps_host, float* action_host_out) {
  if (!a || !obs_host || !action_host_out) return fail(SC_ERR_ARG, "null argument");
  DEV(a);
  // parameters, observation and noise travel in ONE pinned block and one host->device copy
  const size_t obs_off = (sizeof(RowPhaseParams) + 255) & ~(size_t)255, eps_off = obs_off + (size_t)a->Sp32 * sizeof(float);
  const size_t bytes = eps_off + (size_t)round4i(a->A) * sizeof(float);
  if (bytes > IMM_SLOT) return fail(SC_ERR_UNSUPPORTED, "observation too large for the immediate slot");
  ImmSlot s;
  int rc = imm_begin(a, s);
  if (rc) return rc;
  float* h_obs = reinterpret_cast<float*>(s.host + obs_off);
  memset(h_obs, 0, (size_t)a->Sp32 * sizeof(float));
  memcpy(h_obs, obs_host, (size_t)a->S * sizeof(float));
  RowPhaseParams& p = *reinterpret_cast<RowPhaseParams*>(s.host);
  a->bidx_step = a->bidx;
  fill_common(a, p);
  p.program = PROG_ACT;
  p.tm = a->tm_act;
  p.B = 1;
  p.actor = actor_ptrs(a->params + a->off_actor, a->La, a->cfg.kind);
  p.obs = a->obs_dev;
  p.act_out = a->act_out;
  p.act_aux = a->act_aux;
  if (eps_host) {
    memcpy(s.host + eps_off, eps_host, (size_t)a->A * sizeof(float));
    p.eps = reinterpret_cast<const float*>(s.dev + eps_off);
  } else if (a->cfg.kind == SC_KIND_SAC) {
    CK(launch_fill_normal(a->act_eps, a->A, a->cfg.rng_seed, 0x7000000000ull + (a->noise_ctr++), a->st));
    a->launches++;
    p.eps = a->act_eps;
  }
  if ((rc = imm_commit(a, s, bytes))) return rc;
  CK(cudaMemcpyAsync(a->obs_dev, s.dev + obs_off, (size_t)a->Sp32 * sizeof(float), cudaMemcpyDeviceToDevice, a->st));
  CK(launch_rowphase(reinterpret_cast<const RowPhaseParams*>(s.dev), 1, a->row_variant, a->cluster_size, 1, a->tile_tm, a->st));
  a->launches++;
  CK(cudaMemcpyAsync(a->h_act + 32, a->act_out, (size_t)a->A * sizeof(float), cudaMemcpyDeviceToHost, a->st));
  CK(cudaStreamSynchronize(a->st));
  memcpy(action_host_out, a->h_act + 32, (size_t)a->A * sizeof(float));
  return SC_OK;
}

int sc_agent_soft_update(sc_agent* a, float tau) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  DEV(a);
  const bool td3 = a->cfg.kind == SC_KIND_TD3;
  CK(launch_soft_update(a->params + a->off_c1, a->targets + a->toff_c1, a->Lc.size, tau, a->st));
  CK(launch_soft_update(a->params + a->off_c2, a->targets + a->toff_c2, a->Lc.size, tau, a->st));
  a->launches += 2;
  if (td3) {
    CK(launch_soft_update(a->params + a->off_actor, a->targets + a->toff_actor, a->La.size, tau, a->st));
    a->launches++;
  }
  a->split_dirty = true;
  return SC_OK;
}
int sc_agent_params_changed(sc_agent* a) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  a->split_dirty = true;
  return SC_OK;
}

// noise of slot `slot` for this step: host array -> tape data block; NULL -> drawn in the kernel that consumes it (Philox)
static int plan_noise(sc_agent* a, const float* host, int slot, uint64_t stream_tag) {
  const size_t n = (size_t)a->B * a->A;
  if (host) {
    const void* dev = nullptr;
    void* h = emit_data(a, n * sizeof(float), &dev);
    PL(h);
    memcpy(h, host, n * sizeof(float));
    a->eps_dev[slot] = reinterpret_cast<const float*>(dev);
    a->eps_inline[slot] = 0;
  } else {
    a->eps_dev[slot] = a->eps[slot];
    a->eps_inline[slot] = 1;
    a->eps_stream[slot] = stream_tag + (a->noise_ctr++);
  }
  return SC_OK;
}
static void set_eps(sc_agent* a, RowPhaseParams& p, int slot) {
  p.eps = a->eps_dev[slot];
  p.eps_inline = a->eps_inline[slot];
  p.eps_seed = a->cfg.rng_seed;
  p.eps_stream = a->eps_stream[slot];
  p.eps_store = a->eps[slot];                 // the values used are left in the arena ("eps_next" / "eps_cur") for inspection
}
static HintCoefP hint_params(sc_agent* a, float* out, float* rho_update, float* stats_rho, float* stats_g) {
  HintCoefP h;
  h.pi = a->aw.pi; h.hint = a->rp.hint; h.Ap = a->Ap; h.bidx = a->bidx_step; h.B = a->B; h.A = a->A;
  h.thr = a->cfg.hint_threshold; h.admm_rho = a->admm_rho; h.rho = a->rho_dev; h.out = out; h.rho_update = rho_update;
  h.stats_rho = stats_rho; h.stats_g = stats_g;
  return h;
}
static AdmmP admm_params(sc_agent* a, int mode, int it, float* stats_rho) {
  AdmmP m;
  m.mode = mode; m.it = it; m.pi = a->aw.pi; m.hint = a->rp.hint; m.Ap = a->Ap; m.bidx = a->bidx_step; m.B = a->B; m.A = a->A;
  m.lambda = a->admm_lambda; m.y0 = a->admm_y0; m.a0 = a->admm_a0; m.rho = a->rho_dev; m.rho0 = a->admm_rho; m.corr_min = 0.5f;
  m.stats_rho = stats_rho;
  return m;
}

struct StepArgs {
  const sc_noise* noise = nullptr;
  const float* store_tr = nullptr;            // fused store: one packed transition (host) stored before the update
  int store_mode = 0;
  double store_priority = 0.0;
};
struct StepOut { bool did_learn = false, actor_step = false, hint_stats = false; };

// Plans one agent's step into its group's tape (Agent.store_transition + Agent.learn, enet_sac.py:544-628 / enet_td3.py:222-364)
static int plan_step(sc_agent* a, const StepArgs& sa, StepOut& so) {
  const sc_noise* noise = sa.noise;
  int rc;
  a->bidx_step = a->bidx;
  float* stats = a->grp->stats_dev + (size_t)a->gidx * 32;
  if (sa.store_tr) {
    touch_replay(a);
    int mode = sa.store_mode;
    const int slot = next_slot(a, mode);
    const void* dev = nullptr;
    void* h = emit_data(a, (size_t)a->T * sizeof(float), &dev);
    PL(h);
    memcpy(h, sa.store_tr, (size_t)a->T * sizeof(float));
    StoreP* sp = emit_kernel<StoreP>(a, KID_STORE, 0);
    PL(sp);
    sp->rp = a->rp; sp->tr = reinterpret_cast<const float*>(dev); sp->slot = slot; sp->priority_mode = mode;
    sp->priority = sa.store_priority; sp->clip = (double)a->cfg.per_clip;
  }
  if (a->mem_cntr < a->B) return SC_OK;   // enet_sac.py:556-557
  so.did_learn = true;
  if (a->split_dirty) {                   // parameters may have been written from outside since the last step (group_step levels the flag across an ensemble)
    if (a->dense_tc && (rc = plan_split_weights(a))) return rc;
    a->split_dirty = false;
  }
  // Experiment, off by default (SMARTCAL_PREFETCH=1): the step's parameter traffic requested from HBM up front on side stream 2
  // (joined before the first weight-gradient launch).  Measured on B200: 207 vs 202 us per step with the L2 flushed between
  // steps, 192 vs 185 us back to back -- the extra launch and fork / join cost more than the cold misses they were meant to hide.
  static const bool prefetch_on = getenv("SMARTCAL_PREFETCH") && getenv("SMARTCAL_PREFETCH")[0] == '1';
  const bool prefetched = prefetch_on && !a->timing && a->grp->E == 1 && a->dense_tc;
  if (prefetched) {
    if ((rc = emit_ctl(a, REC_FORK, 2))) return rc;
    PrefetchP* pf = emit_kernel<PrefetchP>(a, KID_PREFETCH, 2, 2 * a->sm_count);
    PL(pf);
    memset(pf, 0, sizeof(PrefetchP));
    const void* ptrs[6] = {a->params_hl, a->targets_hl, a->params, a->targets, a->adam_m, a->adam_v};
    const size_t fl[6] = {2 * a->online_floats, 2 * a->target_floats, a->online_floats, a->target_floats, a->online_floats, a->online_floats};
    pf->n = 6;
    for (int i = 0; i < 6; ++i) { pf->ptr[i] = ptrs[i]; pf->bytes[i] = (unsigned long long)fl[i] * sizeof(float); }
    if ((rc = emit_ctl(a, REC_JOIN_RECORD, 2))) return rc;
  }
  const bool sac = a->cfg.kind == SC_KIND_SAC;
  const bool per = a->cfg.prioritized != 0;
  const bool was_quiet = a->quiet;          // nothing touched the replay state since the previous learn() returned
  bool internal_sample = false;
  if (!a->batch_ready) {
    internal_sample = true;
    const double* u01_host = noise ? noise->per_u01_host : nullptr;
    const int32_t* idx_host = noise ? noise->batch_idx_host : nullptr;
    if (per && a->presampled && !u01_host) {
      // the minibatch was drawn on the side stream during the previous step (after its priority update): swap it in
      std::swap(a->bidx, a->bidx_alt); std::swap(a->tree_idx, a->tree_idx_alt);
      std::swap(a->is_w, a->is_w_alt); std::swap(a->is_wsum, a->is_wsum_alt);
      a->buffers["batch_idx"].dev_ptr = a->bidx; a->buffers["tree_idx"].dev_ptr = a->tree_idx; a->buffers["is_weights"].dev_ptr = a->is_w;
      a->bidx_step = a->bidx;
      a->per_beta = a->presample_beta;
      a->presampled = false;
    } else if (per) {
      touch_replay(a);
      const double* u = nullptr;
      if (u01_host) {
        const void* dev = nullptr;
        void* h = emit_data(a, (size_t)a->B * sizeof(double), &dev);
        PL(h);
        memcpy(h, u01_host, (size_t)a->B * sizeof(double));
        u = reinterpret_cast<const double*>(dev);
      }
      a->per_beta = fmin(1.0, a->per_beta + 1e-4);   // enet_sac.py:288
      if (a->timing && (rc = emit_ctl(a, REC_TBEGIN, 4))) return rc;
      PerSampleP* ps = emit_kernel<PerSampleP>(a, KID_PER_SAMPLE, 0, a->B);
      PL(ps);
      *ps = per_sample_params(a, u, a->per_beta, false);
      if (a->timing && (rc = emit_ctl(a, REC_TEND, 4))) return rc;
    } else if (idx_host && a->dense_tc) {
      touch_replay(a);
      const void* dev = nullptr;
      void* h = emit_data(a, (size_t)a->B * sizeof(int32_t), &dev);
      PL(h);
      memcpy(h, idx_host, (size_t)a->B * sizeof(int32_t));
      a->bidx_step = reinterpret_cast<const int*>(dev);      // (k_gather_mb mirrors it into the arena's batch_idx buffer)
    } else if (idx_host) {
      if ((rc = sc_replay_set_batch_host(a, idx_host, a->B))) return rc;
    } else {
      touch_replay(a);
      const int n = (int)(a->mem_cntr < a->C ? a->mem_cntr : a->C);
      UniformP* up = emit_kernel<UniformP>(a, KID_UNIFORM, 0, n);
      PL(up);
      *up = uniform_params(a, n);
    }
  }
  if (a->presampled) { a->presampled = false; a->sample_ctr--; }   // (a batch supplied from outside supersedes it)
  a->batch_ready = false;
  if (sac) {
    if ((rc = plan_noise(a, noise ? noise->eps_next_host : nullptr, 0, 0x1000000000ull))) return rc;
    if ((rc = plan_noise(a, noise ? noise->eps_cur_host : nullptr, 1, 0x2000000000ull))) return rc;
  }
  // ---- critic phase ----
  RowPhaseParams p;
  fill_common(a, p);
  p.program = PROG_CRITIC;
  p.tm = a->tm_critic;
  if (sac) {
    p.actor = actor_ptrs(a->params + a->off_actor, a->La, a->cfg.kind);
    set_eps(a, p, 0);
  } else {
    p.actor = actor_ptrs(a->targets + a->toff_actor, a->La, a->cfg.kind);
    double nz = 0.0;
    if (noise && noise->has_td3_target_noise) nz = noise->td3_target_noise;
    nz = nz < -0.5 ? -0.5 : (nz > 0.5 ? 0.5 : nz);          // enet_td3.py:249 clamp in float64, cast with the add
    p.td3_noise_c = (float)nz;
  }
  if ((rc = plan_rowphase(a, p, 0))) return rc;
  bool per_forked = false;
  if (per) {
    // PER.batch_update only depends on the priorities written by the critic row-phase and is only consumed by the NEXT
    // sample: it runs on a side stream under the weight-gradient / actor kernels and is joined at the end of the step
    if (a->timing) {
      if ((rc = emit_ctl(a, REC_TBEGIN, 5))) return rc;
      PerUpdateP* pu = emit_kernel<PerUpdateP>(a, KID_PER_UPDATE, 0, a->B);
      PL(pu);
      *pu = per_update_params(a, nullptr, a->ps_out);
      if ((rc = emit_ctl(a, REC_TEND, 5))) return rc;
    } else {
      const bool no_presample = getenv("SMARTCAL_NO_PRESAMPLE") != nullptr;    // (read per step: tests toggle it)
      if ((rc = emit_ctl(a, REC_FORK, 1))) return rc;
      PerUpdateP* pu = emit_kernel<PerUpdateP>(a, KID_PER_UPDATE, 1, a->B);
      PL(pu);
      *pu = per_update_params(a, nullptr, a->ps_out);
      if (internal_sample && was_quiet && !no_presample && !(noise && noise->per_u01_host)) {
        // learn() is being called back to back with the library's own sampling: draw the next minibatch now, from the tree
        // this update just produced (PER.sample_buffer of the next step, enet_sac.py:270-312), under the remaining kernels
        a->presample_beta = fmin(1.0, a->per_beta + 1e-4);
        PerSampleP* ps = emit_kernel<PerSampleP>(a, KID_PER_SAMPLE, 1, a->B);
        PL(ps);
        *ps = per_sample_params(a, nullptr, a->presample_beta, true);
        a->presampled = true;
      }
      if ((rc = emit_ctl(a, REC_JOIN_RECORD, 1))) return rc;
      per_forked = true;
    }
  }
  bool actor_step = true;
  if (!sac) actor_step = ((a->learn_counter + 1) % (a->cfg.update_actor_interval > 0 ? a->cfg.update_actor_interval : 1)) == 0;
  so.actor_step = actor_step;
  a->t_critic++;
  WgradParams w;
  fill_wgrad(a, w, false, actor_step ? 1 : 0, 1);
  if (prefetched && (rc = emit_ctl(a, REC_JOIN_WAIT, 2))) return rc;      // (long done: closes the fork before stream 2 is forked again)
  if ((rc = plan_wgrad(a, w, 1))) return rc;
  // ---- actor phase ----
  if (actor_step) {
    fill_common(a, p);
    p.program = PROG_ACTOR;
    p.tm = a->tm_actor;
    p.actor = actor_ptrs(a->params + a->off_actor, a->La, a->cfg.kind);
    if (sac) set_eps(a, p, 1); else p.eps = nullptr;
    if (sac && a->cfg.use_hint) {
      // pre-pass: forward only matters (pi), then the batch-wide penalty scalar (enet_sac.py:601-602)
      if ((rc = plan_rowphase(a, p, -1))) return rc;
      HintCoefP* hp = emit_kernel<HintCoefP>(a, KID_HINT, 0);
      PL(hp);
      *hp = hint_params(a, a->hint_out, nullptr, stats + 16, stats + 20);
      p.hint_coef = a->hint_out;
      so.hint_stats = true;
    }
    if (!sac && a->cfg.use_hint) {
      // adaptive-ADMM inner loop (enet_td3.py:310-361): Nadmm = 5 actor steps against the hint constraint; the penalty
      // lives on the device (rho_dev) and may be re-estimated after the 4th step; targets move once, after the loop
      AdmmP* m0 = emit_kernel<AdmmP>(a, KID_ADMM, 0);
      PL(m0);
      *m0 = admm_params(a, 0, 0, nullptr);
      p.admm_lambda = a->admm_lambda;
      p.admm_rho = a->rho_dev;
      for (int it = 0; it < 5; ++it) {
        if ((rc = plan_rowphase(a, p, 2))) return rc;
        a->t_actor++;
        fill_wgrad(a, w, true, it == 4 ? 1 : 0, 2);
        if ((rc = plan_wgrad(a, w, 3))) return rc;
        AdmmP* mi = emit_kernel<AdmmP>(a, KID_ADMM, 0);
        PL(mi);
        *mi = admm_params(a, 1, it, stats + 16);
      }
      so.hint_stats = true;
    } else {
      if ((rc = plan_rowphase(a, p, 2))) return rc;
      a->t_actor++;
      fill_wgrad(a, w, true, 1, 2);
      if ((rc = plan_wgrad(a, w, 3))) return rc;
    }
    if (sac && a->cfg.use_hint && a->learn_counter % 10 == 0) {
      // dual ascent on the hint multiplier with a fresh no-grad sample from the UPDATED actor (enet_sac.py:608-617)
      if ((rc = plan_noise(a, noise ? noise->eps_dual_host : nullptr, 2, 0x3000000000ull))) return rc;
      set_eps(a, p, 2);
      p.hint_coef = nullptr;
      if ((rc = plan_rowphase(a, p, -1))) return rc;
      HintCoefP* hp = emit_kernel<HintCoefP>(a, KID_HINT, 0);
      PL(hp);
      *hp = hint_params(a, a->hint_out + 4, a->rho_dev, stats + 16, nullptr);
    }
  }
  if (per_forked && (rc = emit_ctl(a, REC_JOIN_WAIT, 1))) return rc;     // join: later work on the stream sees the updated tree
  a->quiet = internal_sample;               // cleared by anything that may change what the next sample would see
  a->learn_counter++;
  return SC_OK;
}

// ---- group execution ---------------------------------------------------------------------------------------------
static void group_free(sc_group* g) {
  DevGuard dev_guard_(g->device);
  if (g->st) cudaStreamSynchronize(g->st);
  for (auto& kv : g->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  g->graphs.clear();
  for (cudaStream_t* s : {&g->st_side, &g->st_side2, &g->st_cap, &g->st_copy})
    if (*s) { cudaStreamSynchronize(*s); cudaStreamDestroy(*s); *s = nullptr; }
  for (auto& e : g->ev_fork) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto& e : g->ev_join) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto& e : g->ev_up) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto& e : g->ev_done) if (e) { cudaEventDestroy(e); e = nullptr; }
  for (auto& e : g->ev_stats) if (e) { cudaEventDestroy(e); e = nullptr; }
  if (g->ev_sync) { cudaEventDestroy(g->ev_sync); g->ev_sync = nullptr; }
  for (auto& h : g->tape_host) if (h) { cudaFreeHost(h); h = nullptr; }
  for (auto& d : g->tape_dev) if (d) { cudaFree(d); d = nullptr; }
  if (g->stats_dev) { cudaFree(g->stats_dev); g->stats_dev = nullptr; }
  if (g->h_stats) { cudaFreeHost(g->h_stats); g->h_stats = nullptr; }
}
static int group_init(sc_group* g, sc_agent** agents, int n) {
  g->ag.assign(agents, agents + n);
  g->E = n;
  g->device = agents[0]->device;
  g->st = agents[0]->st;
  DevGuard dev_guard_(g->device);
  CK(cudaStreamCreateWithFlags(&g->st_side, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g->st_side2, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g->st_cap, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g->st_copy, cudaStreamNonBlocking));
  for (auto& e : g->ev_fork) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : g->ev_join) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : g->ev_up) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : g->ev_done) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : g->ev_stats) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&g->ev_sync, cudaEventDisableTiming));
  g->pend_flags.assign(n, 0);
  g->pend_counter.assign(n, 0);
  // per agent and step: <= 2 + 7 row-phase blocks, 4 + 14 dense blocks, 12 weight-gradient blocks, the small kernels, and
  // the host arrays (3 noise tensors, B uniforms or indices, one transition)
  const sc_agent* a = agents[0];
  const size_t per_agent = 18 * sizeof(RowPhaseParams) + 18 * sizeof(DenseTcParams) + 12 * (sizeof(WgradParams) + sizeof(WgradTcParams)) + 8192 +
                           (size_t)3 * a->B * a->A * sizeof(float) + (size_t)a->B * sizeof(double) + (size_t)a->T * sizeof(float) + 64 * 256;
  g->tape_cap = ((per_agent * (size_t)n) + 4095) & ~(size_t)4095;
  for (auto& d : g->tape_dev) CK(cudaMalloc((void**)&d, g->tape_cap));
  for (auto& h : g->tape_host) CK(cudaHostAlloc((void**)&h, g->tape_cap, cudaHostAllocDefault));
  CK(cudaMalloc((void**)&g->stats_dev, (size_t)n * 32 * sizeof(float)));
  CK(cudaMemset(g->stats_dev, 0, (size_t)n * 32 * sizeof(float)));
  CK(cudaHostAlloc((void**)&g->h_stats, (size_t)2 * n * 32 * sizeof(float), cudaHostAllocDefault));
  const char* e = getenv("SMARTCAL_GRAPH");
  g->use_graph = (e && e[0] == '0') ? 0 : 1;
  return SC_OK;
}
static int group_begin(sc_group* g) {
  CK(cudaEventSynchronize(g->ev_up[g->tk]));     // the tape copy that read this pinned buffer two steps ago
  g->recs.clear();
  g->used = 0;
  g->plan_error = 0;
  for (int i = 0; i < g->E; ++i) { g->ag[i]->grp = g; g->ag[i]->gidx = i; g->ag[i]->cursor = 0; }
  return SC_OK;
}
static cudaError_t issue_recs(sc_group* g, cudaStream_t st0) {
  cudaStream_t sts[3] = {st0, g->st_side, g->st_side2};
  sc_agent* a0 = g->ag[0];
  cudaEvent_t tev[8][2];
  cudaError_t e = cudaSuccess;
  for (const Rec& r : g->recs) {
    cudaStream_t st = sts[r.stream];
    const char* pp = g->tape_dev[g->tk] + r.off;
    switch (r.type) {
      case REC_DATA: break;
      case REC_FORK:
        e = cudaEventRecord(g->ev_fork[r.g[0]], st0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(sts[r.g[0]], g->ev_fork[r.g[0]], 0);
        break;
      case REC_JOIN_RECORD: e = cudaEventRecord(g->ev_join[r.g[0]], sts[r.g[0]]); break;
      case REC_JOIN_WAIT: e = cudaStreamWaitEvent(st0, g->ev_join[r.g[0]], 0); break;
      case REC_TBEGIN: timer_begin(a0, r.g[0], &tev[r.g[0]][0], &tev[r.g[0]][1]); break;
      case REC_TEND: timer_end(a0, r.g[0], tev[r.g[0]][0], tev[r.g[0]][1]); break;
      case REC_KERNEL:
        switch (r.kid) {
          case KID_STORE: e = launch_replay_store_t(reinterpret_cast<const StoreP*>(pp), g->E, st); break;
          case KID_PER_SAMPLE: e = launch_per_sample(reinterpret_cast<const PerSampleP*>(pp), g->E, r.g[0], st); break;
          case KID_UNIFORM: e = launch_sample_uniform(reinterpret_cast<const UniformP*>(pp), g->E, r.g[0], st); break;
          case KID_GATHER: e = launch_gather_mb(reinterpret_cast<const GatherP*>(pp), g->E, r.g[0], st); break;
          case KID_DENSE: e = launch_dense_tc(reinterpret_cast<const DenseTcParams*>(pp), g->E, r.g[0], r.g[1], r.g[2], r.g[3], st); break;
          case KID_ROW: e = launch_rowphase(reinterpret_cast<const RowPhaseParams*>(pp), g->E, r.g[0], r.g[1], r.g[2], r.g[3], st); break;
          case KID_PER_UPDATE: e = launch_per_update(reinterpret_cast<const PerUpdateP*>(pp), g->E, r.g[0], st); break;
          case KID_WGRAD: e = launch_wgrad(reinterpret_cast<const WgradParams*>(pp), g->E, r.g[0], st); break;
          case KID_WGRAD_TC: e = launch_wgrad_tc(reinterpret_cast<const WgradTcParams*>(pp), g->E, r.g[0], st); break;
          case KID_HINT: e = launch_hint_coef(reinterpret_cast<const HintCoefP*>(pp), g->E, st); break;
          case KID_ADMM: e = launch_admm_update(reinterpret_cast<const AdmmP*>(pp), g->E, st); break;
          case KID_TAIL: e = launch_tail(reinterpret_cast<const RowPhaseParams*>(pp), g->E, r.g[0], st); break;
          case KID_DGRAD: e = launch_dgrad_tc(reinterpret_cast<const DenseTcParams*>(pp), g->E, r.g[0], r.g[1], r.g[2], r.g[3], st); break;
          case KID_SPLIT: e = launch_split(reinterpret_cast<const SplitP*>(pp), g->E, r.g[0], st); break;
          case KID_PREFETCH: e = launch_l2_prefetch(reinterpret_cast<const PrefetchP*>(pp), g->E, r.g[0], st); break;
          default: e = cudaErrorInvalidValue;
        }
        break;
    }
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}
static int group_issue(sc_group* g);
static uint64_t plan_key(const sc_group* g) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
  mix((uint64_t)g->E | ((uint64_t)g->tk << 32));               // (the kernels of a graph read one particular device tape buffer)
  for (const Rec& r : g->recs) {
    mix((uint64_t)r.type | ((uint64_t)r.kid << 8) | ((uint64_t)r.stream << 16));
    for (int i = 0; i < 5; ++i) mix((uint64_t)(uint32_t)r.g[i]);
    mix((uint64_t)r.off);
    mix((uint64_t)r.bytes);
  }
  return h;
}
// uploads the tape and runs the planned records: eagerly the first time a plan shape is seen, from its CUDA graph afterwards
static int group_run(sc_group* g) {
  if (g->plan_error) return SC_ERR_STATE;
  for (int i = 0; i < g->E; ++i)
    if (g->ag[i]->cursor != g->recs.size()) return plan_fail(g, "missing records");
  if (g->recs.empty()) return SC_OK;
  // upload on the copy stream as soon as the step that read this device buffer last (two steps ago) has finished; the
  // agents' stream only waits for the copy, which normally completed while the previous step was still running
  const int k = g->tk;
  CK(cudaStreamWaitEvent(g->st_copy, g->ev_done[k], 0));
  CK(cudaMemcpyAsync(g->tape_dev[k], g->tape_host[k], g->used, cudaMemcpyHostToDevice, g->st_copy));
  CK(cudaEventRecord(g->ev_up[k], g->st_copy));
  CK(cudaStreamWaitEvent(g->st, g->ev_up[k], 0));
  int rc = group_issue(g);
  CK(cudaEventRecord(g->ev_done[k], g->st));
  g->tk ^= 1;
  return rc;
}
static int group_issue(sc_group* g) {
  bool timing = false;
  for (const Rec& r : g->recs) timing = timing || r.type == REC_TBEGIN;
  if (g->use_graph && !timing) {
    GraphEntry& ge = g->graphs[plan_key(g)];
    if (ge.exec == nullptr && ge.seen >= 1) {
      // second occurrence of this plan shape: capture it (on a library-owned stream: the caller's may be the legacy default
      // stream) -- programmatic dependent launches become programmatic graph edges
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamBeginCapture(g->st_cap, cudaStreamCaptureModeRelaxed);
      if (e == cudaSuccess) {
        e = issue_recs(g, g->st_cap);
        cudaError_t e2 = cudaStreamEndCapture(g->st_cap, &graph);
        if (e == cudaSuccess) e = e2;
      }
      if (e == cudaSuccess) e = cudaGraphInstantiate(&ge.exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (e != cudaSuccess) {
        cudaGetLastError();
        ge.exec = nullptr;
        g->use_graph = 0;                                     // this driver cannot graph the step: stay eager
        if (getenv("SMARTCAL_DEBUG")) fprintf(stderr, "[smartcal] CUDA graph capture failed (%s): eager launches\n", cudaGetErrorString(e));
      }
    }
    ge.seen++;
    if (ge.exec != nullptr) {
      CK(cudaGraphLaunch(ge.exec, g->st));
      g->graph_launches++;
      return SC_OK;
    }
  }
  CK(issue_recs(g, g->st));
  return SC_OK;
}
static void fill_stats(sc_agent* a, const float* hs, const StepOut& so, sc_learn_stats* out) {
  const bool sac = a->cfg.kind == SC_KIND_SAC;
  out->critic_1_loss = hs[0];
  out->critic_2_loss = hs[1];
  out->actor_loss = hs[2];
  out->q1_mean = hs[3];
  out->q2_mean = hs[4];
  out->target_mean = hs[5];
  out->logp_mean = hs[6];
  if (so.hint_stats) a->rho = hs[16];
  out->rho = a->rho;
  out->gfun = 0.f;
  if (sac && a->cfg.use_hint && so.actor_step) {
    out->gfun = hs[20];
    out->actor_loss += hs[21];
  }
  out->learn_counter = (int32_t)a->learn_counter;
  out->did_learn = 1;
  out->actor_step = so.actor_step ? 1 : 0;
}
// one step of every agent of the group; args / stats arrays have E entries (stats nullable)
static int group_step(sc_group* g, const StepArgs* args, sc_learn_stats* stats_host) {
  DevGuard dev_guard_(g->device);
  int rc = group_begin(g);
  if (rc) return rc;
  std::vector<StepOut> so(g->E);
  {
    bool dirty = false;
    for (int i = 0; i < g->E; ++i) dirty = dirty || g->ag[i]->split_dirty;
    for (int i = 0; i < g->E; ++i) g->ag[i]->split_dirty = dirty;
  }
  for (int i = 0; i < g->E; ++i) {
    if (stats_host) memset(&stats_host[i], 0, sizeof(sc_learn_stats));
    if ((rc = plan_step(g->ag[i], args[i], so[i]))) break;
    if (so[i].did_learn != so[0].did_learn) { rc = plan_fail(g, "replay fill level"); break; }
  }
  for (int i = 0; i < g->E; ++i) g->ag[i]->grp = g->ag[i]->own;
  if (rc) return rc;
  if ((rc = group_run(g))) return rc;
  if (stats_host && g->stats_deferred) {
    // hand out the PREVIOUS step's block (its copy finished long ago), then queue this step's copy behind its kernels
    if (g->stats_pending) {
      const int k = g->stats_slot ^ 1;
      CK(cudaEventSynchronize(g->ev_stats[k]));
      for (int i = 0; i < g->E; ++i) {
        StepOut o;
        o.did_learn = (g->pend_flags[i] & 1) != 0; o.actor_step = (g->pend_flags[i] & 2) != 0; o.hint_stats = (g->pend_flags[i] & 4) != 0;
        if (o.did_learn) {
          fill_stats(g->ag[i], g->h_stats + ((size_t)k * g->E + i) * 32, o, &stats_host[i]);
          stats_host[i].learn_counter = g->pend_counter[i];
        }
      }
    }
    g->stats_pending = 0;
    if (so[0].did_learn) {
      const int k = g->stats_slot;
      CK(cudaMemcpyAsync(g->h_stats + (size_t)k * g->E * 32, g->stats_dev, (size_t)g->E * 32 * sizeof(float), cudaMemcpyDeviceToHost, g->st));
      CK(cudaEventRecord(g->ev_stats[k], g->st));
      for (int i = 0; i < g->E; ++i) {
        g->pend_flags[i] = (so[i].did_learn ? 1 : 0) | (so[i].actor_step ? 2 : 0) | (so[i].hint_stats ? 4 : 0);
        g->pend_counter[i] = (int32_t)g->ag[i]->learn_counter;
      }
      g->stats_slot ^= 1;
      g->stats_pending = 1;
    }
    return SC_OK;
  }
  if (stats_host) {
    if (so[0].did_learn) {
      CK(cudaMemcpyAsync(g->h_stats, g->stats_dev, (size_t)g->E * 32 * sizeof(float), cudaMemcpyDeviceToHost, g->st));
      // the one synchronisation of a step that reads its losses back: poll the copy's event instead of blocking in the driver (a
      // blocked thread is woken tens of microseconds after the copy lands, which is 10 % of a step)
      CK(cudaEventRecord(g->ev_sync, g->st));
      {
        cudaError_t q;
        while ((q = cudaEventQuery(g->ev_sync)) == cudaErrorNotReady) {}
        if (q != cudaSuccess) return fail(SC_ERR_CUDA, "%s:%d loss read-back: %s", __FILE__, __LINE__, cudaGetErrorString(q));
      }
      for (int i = 0; i < g->E; ++i) fill_stats(g->ag[i], g->h_stats + (size_t)i * 32, so[i], &stats_host[i]);
    } else {
      for (int i = 0; i < g->E; ++i) stats_host[i].learn_counter = (int32_t)g->ag[i]->learn_counter;
    }
  }
  return SC_OK;
}

int sc_agent_learn(sc_agent* a, const sc_noise* noise, sc_learn_stats* stats_host) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  StepArgs sa;
  sa.noise = noise;
  return group_step(a->own, &sa, stats_host);
}
int sc_agent_set_stats_mode(sc_agent* a, int deferred) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  a->own->stats_deferred = deferred != 0;
  a->own->stats_pending = 0;
  return SC_OK;
}
int sc_agent_store_learn(sc_agent* a, const float* transition_host, int priority_mode, double priority, const sc_noise* noise,
                         sc_learn_stats* stats_host) {
  if (!a || !transition_host) return fail(SC_ERR_ARG, "null argument");
  StepArgs sa;
  sa.noise = noise; sa.store_tr = transition_host; sa.store_mode = priority_mode; sa.store_priority = priority;
  return group_step(a->own, &sa, stats_host);
}

// ---- ensembles: E independent agents (the reference's seed-ensemble axis, elasticnet/do.sh:1-6) stepped by ONE launch sequence ----
int sc_ensemble_create(sc_agent** agents, int n, sc_ensemble** out) {
  if (!agents || !out || n < 1 || n > 1024) return fail(SC_ERR_ARG, "bad argument");
  *out = nullptr;
  const sc_agent* a0 = agents[0];
  for (int i = 0; i < n; ++i) {
    const sc_agent* a = agents[i];
    if (!a) return fail(SC_ERR_ARG, "null agent %d", i);
    sc_config c = a->cfg, c0 = a0->cfg;
    c.rng_seed = c0.rng_seed = 0;                             // agents may differ in their seed (and their weights / replay contents) only
    if (memcmp(&c, &c0, sizeof(c)) != 0 || a->device != a0->device || a->st != a0->st)
      return fail(SC_ERR_ARG, "agent %d differs from agent 0 in configuration, device or stream", i);
    if (a->row_variant != a0->row_variant || a->cluster_size != a0->cluster_size || a->nclusters != a0->nclusters || a->tile_tm != a0->tile_tm)
      return fail(SC_ERR_STATE, "agent %d has a different kernel geometry", i);
    for (int j = 0; j < i; ++j)
      if (agents[j] == a) return fail(SC_ERR_ARG, "agent %d listed twice", i);
  }
  sc_ensemble* e = new sc_ensemble();
  int rc = group_init(&e->g, agents, n);
  if (rc) { group_free(&e->g); delete e; return rc; }
  *out = e;
  return SC_OK;
}
int sc_ensemble_destroy(sc_ensemble* e) {
  if (!e) return SC_OK;
  group_free(&e->g);
  delete e;
  return SC_OK;
}
int sc_ensemble_learn(sc_ensemble* e, const sc_noise* noise_array, sc_learn_stats* stats_host_array) {
  if (!e) return fail(SC_ERR_ARG, "null ensemble");
  std::vector<StepArgs> sa(e->g.E);
  for (int i = 0; i < e->g.E; ++i) sa[i].noise = noise_array ? &noise_array[i] : nullptr;
  return group_step(&e->g, sa.data(), stats_host_array);
}
int sc_ensemble_store_learn(sc_ensemble* e, const float* transitions_host, int priority_mode, const double* priorities,
                            const sc_noise* noise_array, sc_learn_stats* stats_host_array) {
  if (!e || !transitions_host) return fail(SC_ERR_ARG, "null argument");
  std::vector<StepArgs> sa(e->g.E);
  for (int i = 0; i < e->g.E; ++i) {
    sa[i].noise = noise_array ? &noise_array[i] : nullptr;
    sa[i].store_tr = transitions_host + (size_t)i * e->g.ag[i]->T;
    sa[i].store_mode = priority_mode;
    sa[i].store_priority = priorities ? priorities[i] : 0.0;
  }
  return group_step(&e->g, sa.data(), stats_host_array);
}
int64_t sc_ensemble_launch_count(sc_ensemble* e) { return e ? e->g.launches : -1; }

int64_t sc_agent_launch_count(sc_agent* a) { return a ? a->launches : -1; }

int sc_agent_enable_kernel_timing(sc_agent* a, int on) {
  if (!a) return fail(SC_ERR_ARG, "null agent");
  a->timing = on != 0;
  return SC_OK;
}
int sc_agent_kernel_time(sc_agent* a, int which, int reset, double* ms_total, int64_t* launches) {
  if (!a || which < 0 || which >= 8) return fail(SC_ERR_ARG, "bad argument");
  DEV(a);
  CK(cudaStreamSynchronize(a->st));
  KernelTimer& t = a->timers[which];
  for (auto& pr : t.pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) t.ms += ms;
    t.pool.push_back(pr);
  }
  t.pending.clear();
  if (ms_total) *ms_total = t.ms;
  if (launches) *launches = t.launches;
  if (reset) { t.ms = 0; t.launches = 0; }
  return SC_OK;
}

}  // extern "C"
