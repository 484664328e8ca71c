/*
 * smartcal_b200.h -- C ABI of libsmartcal_b200.so (sm_100a).
 *
 * The reference (SarodYatawatta/smart-calibration) has NO native/FFI layer: callers use the Python
 * classes of elasticnet/enet_sac.py and enet_td3.py directly.  This header is therefore the boundary a
 * maintainer binds INSTEAD of the torch-eager bodies of those classes; each entry point cites the
 * reference method whose body it replaces.  The Python facade in
 * smart-calibration_b200/elasticnet/{enet_sac,enet_td3}.py is a ctypes binding of exactly these symbols
 * (INTEGRATION.md shows the stub).
 *
 * Conventions: every function returns 0 on success, non-zero on failure (never throws, never aborts);
 * sc_last_error() returns a thread-local message.  All work is enqueued on the CUDA stream passed to
 * sc_agent_create (cudaStream_t as void*; NULL = legacy default stream).  Pointers named *_host are
 * host memory (pinned or pageable), everything else is device memory.  Plain pointers and sizes only.
 * There is NO CPU fallback: creating an agent without a CUDA device fails with SC_ERR_CUDA.
 */
#ifndef SMARTCAL_B200_H
#define SMARTCAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_ABI_VERSION 3

enum { SC_OK = 0, SC_ERR_ARG = 1, SC_ERR_CUDA = 2, SC_ERR_STATE = 3, SC_ERR_UNSUPPORTED = 4 };

/* agent kind: which reference Agent the handle mirrors */
enum { SC_KIND_SAC = 0 /* enet_sac.py:478 */, SC_KIND_TD3 = 1 /* enet_td3.py:124 */ };

/* network ids (reference attribute names) */
enum {
  SC_NET_ACTOR = 0,           /* agent.actor            */
  SC_NET_CRITIC_1 = 1,        /* agent.critic_1         */
  SC_NET_CRITIC_2 = 2,        /* agent.critic_2         */
  SC_NET_TARGET_CRITIC_1 = 3, /* agent.target_critic_1  */
  SC_NET_TARGET_CRITIC_2 = 4, /* agent.target_critic_2  */
  SC_NET_TARGET_ACTOR = 5     /* agent.target_actor (TD3 only) */
};

/* which copy of a network's tensors */
enum { SC_SLOT_PARAM = 0, SC_SLOT_ADAM_M = 1 /* exp_avg */, SC_SLOT_ADAM_V = 2 /* exp_avg_sq */ };

/* DDPG (enet_ddpg.py:192-303) runs on the TD3 program: kind = SC_KIND_TD3, update_actor_interval = 1, target noise 0,
 * both critic slots loaded with the same network, and this flag for its sum-of-squares critic loss (enet_ddpg.py:281) */
enum { SC_FLAG_CRITIC_LOSS_SUM = 1 };

/* constructor kwargs of the reference Agent (enet_sac.py:479-480, enet_td3.py:125-126) */
typedef struct sc_config {
  int32_t abi_version;   /* = SC_ABI_VERSION */
  int32_t kind;          /* SC_KIND_*                                   */
  int32_t state_dim;     /* input_dims[0] (= N + N*M = 420 elastic net)  */
  int32_t n_actions;     /* n_actions                                   */
  int32_t batch_size;    /* batch_size                                  */
  int32_t mem_size;      /* max_mem_size (power of two when prioritized) */
  int32_t prioritized;   /* prioritized (SAC: derived TD3-style wiring) */
  int32_t use_hint;      /* use_hint                                    */
  int32_t update_actor_interval; /* TD3 only (enet_td3.py:137)          */
  int32_t flags;         /* SC_FLAG_* (0 for SAC / TD3)                  */
  /* Python floats of the reference constructor stay doubles so that every derived scalar
   * (lr/(1-beta1^t), 1-tau, ...) rounds to float32 exactly where torch rounds it */
  double gamma, tau, lr_actor, lr_critic;
  double reward_scale;   /* SAC `scale` (enet_sac.py:503); TD3 ignores   */
  double alpha;          /* SAC temperature (enet_sac.py:501)            */
  float max_action;      /* 1 in the reference (enet_sac.py:486)         */
  float per_clip;        /* PER.absolute_error_upper (100, enet_sac.py:212) */
  float hint_threshold;  /* 0.1  (enet_sac.py:514)                       */
  float admm_rho;        /* 0.01 (enet_sac.py:516)                       */
  uint64_t rng_seed;     /* device Philox stream (used when noise pointers are NULL) */
} sc_config;

typedef struct sc_agent sc_agent;

/* a 2-D view of a tensor inside the agent's HBM arena (row-major, `pitch` floats between rows) */
typedef struct sc_tensor_info {
  void* dev_ptr;
  int64_t rows, cols, pitch;
  int32_t elem_bytes; /* 4 = float32, 8 = float64 (sum tree), 1 = u8 */
  int32_t is_int;     /* 1: elements are int32 (index buffers) */
} sc_tensor_info;

/* per-step diagnostics, written to host by sc_agent_learn when `stats_host` != NULL (forces a sync) */
typedef struct sc_learn_stats {
  float critic_1_loss, critic_2_loss, actor_loss; /* F.mse_loss / actor_loss of enet_sac.py:580-602 */
  float q1_mean, q2_mean, target_mean, logp_mean; /* batch means, diagnostics only                   */
  float gfun, rho;                                /* hint penalty state (enet_sac.py:601,617)        */
  int32_t learn_counter;                          /* agent.learn_counter AFTER this call             */
  int32_t did_learn;                              /* 0 when mem_cntr < batch_size (enet_sac.py:556)  */
  int32_t actor_step;                             /* TD3: whether the delayed actor update ran       */
} sc_learn_stats;

/* randomness for one learn(); NULL pointers select the on-device Philox stream.
 * Host arrays reproduce the reference's own RNG streams (see oracle/sac_ref.py header). */
typedef struct sc_noise {
  const float* eps_next_host;    /* (B,A) N(0,1): Normal.sample  on s'   (enet_sac.py:570)      */
  const float* eps_cur_host;     /* (B,A) N(0,1): Normal.rsample on s    (enet_sac.py:589)      */
  const float* eps_dual_host;    /* (B,A) N(0,1): dual update, every 10th call (enet_sac.py:611) */
  double td3_target_noise;       /* np.random.normal(scale=0.2) scalar   (enet_td3.py:249)      */
  int32_t has_td3_target_noise;
  int32_t reserved;
  /* the minibatch draw of this learn(), when the caller owns the numpy stream and has not already sampled through
   * sc_per_sample / sc_replay_set_batch_host (they travel in the same host->device copy as the noise): */
  const double* per_u01_host;    /* (B) float64 uniforms of PER.sample_buffer (enet_sac.py:294), prioritized agents   */
  const int32_t* batch_idx_host; /* (B) np.random.choice(max_mem, B, replace=False) (enet_sac.py:48), uniform buffer */
} sc_noise;

const char* sc_last_error(void);
int sc_abi_version(void);

/* ---- lifetime: replaces Agent.__init__ (enet_sac.py:479-521 / enet_td3.py:125-164) ------------------ */
/* Bytes of HBM the agent needs (nets, Adam moments, targets, replay, sum tree, activation workspace). */
size_t sc_agent_arena_bytes(const sc_config* cfg);
/* `arena` may be caller-owned device memory of at least sc_agent_arena_bytes() (e.g. a torch tensor,
 * 256-byte aligned) or NULL to let the library cudaMalloc it.  Parameters start zeroed: the caller
 * uploads initial weights through sc_agent_tensor() views, then calls sc_agent_soft_update(tau=1). */
int sc_agent_create(const sc_config* cfg, void* arena, size_t arena_bytes, void* cuda_stream, sc_agent** out);
int sc_agent_destroy(sc_agent* a);
int sc_agent_sync(sc_agent* a); /* cudaStreamSynchronize on the agent's stream */

/* ---- state-dict interop: nn.Module.state_dict()/load_state_dict() of the reference networks ---------- */
/* `name` is a reference state_dict key ("fc11.weight", "bn3.bias", "fc4mu.weight", ...).  The special
 * name "*" returns the whole contiguous block of the network (1 x n_floats) -- used for one-shot NCCL
 * broadcast of actor weights (replaces Learner.get_actor_params, distributed_per_sac.py:84-90). */
int sc_agent_tensor(sc_agent* a, int net, int slot, const char* name, sc_tensor_info* out);
/* replay arrays ("state_memory", "new_state_memory", "action_memory", "reward_memory",
 * "terminal_memory", "hint_memory", "tree" = float64 sum tree) and, for tests, activation workspace
 * buffers ("c1.h11", "c1.dz12", "y", "q1", ...). */
int sc_agent_buffer(sc_agent* a, const char* name, sc_tensor_info* out);
int sc_agent_get_counters(sc_agent* a, int64_t* mem_cntr, int64_t* learn_counter, int64_t* data_pointer, double* per_beta);
int sc_agent_set_counters(sc_agent* a, int64_t mem_cntr, int64_t learn_counter, int64_t data_pointer, double per_beta);
int sc_agent_set_scalars(sc_agent* a, float alpha, float rho, float admm_rho);

/* ---- replay: replaces ReplayBuffer.store_transition (enet_sac.py:35-44) and
 *      PER.store_transition / store_transition_from_buffer (enet_sac.py:237-268) -------------------- */
/* One packed transition = [state(S) | new_state(S) | action(A) | hint(A) | reward | done] float32.
 * priority_mode: 0 = uniform buffer (no tree), 1 = PER with error==None (max leaf, or clip if empty),
 * 2 = PER with `priority` already evaluated by the caller (min((|err|+eps)^alpha, clip)),
 * 3 = like 2 but the propagated change is formed in float32 (reference behaviour when the reward
 *     handed to TD3.store_transition is a float32 torch scalar, enet_td3.py:205). */
int sc_replay_store_host(sc_agent* a, const float* transition_host, int priority_mode, double priority);
/* n packed transitions already in HBM (e.g. landed by an NCCL gather); stores transition `index`. */
int sc_replay_store_dev(sc_agent* a, const float* transitions_dev, int index, int priority_mode, double priority);
/* bulk fill of slots [0,n) from packed host transitions (synthetic benchmarks / checkpoint load) */
int sc_replay_fill_host(sc_agent* a, const float* transitions_host, int n, const double* priorities_host_or_null);

/* ---- sampling: replaces ReplayBuffer.sample_buffer (enet_sac.py:46-57) and PER.sample_buffer (:270-312) */
/* indices drawn by the caller's numpy stream (np.random.choice(max_mem,B,replace=False)) */
int sc_replay_set_batch_host(sc_agent* a, const int32_t* idx_host, int n);
/* on-device draw without replacement (Philox keys + selection), no host traffic */
int sc_replay_sample_uniform_dev(sc_agent* a);
/* stratified proportional sampling on the device sum tree.  u01_host: B float64 uniforms in [0,1) from the
 * caller's numpy stream, or NULL for device Philox.  Optional host outputs (force a sync when non-NULL). */
int sc_per_sample(sc_agent* a, const double* u01_host, int32_t* tree_idx_host_out, float* is_weights_host_out);
/* replaces PER.batch_update (enet_sac.py:314-323): abs errors (B) -> priorities -> tree.  Host arrays. */
int sc_per_update_host(sc_agent* a, const int32_t* tree_idx_host, const float* abs_err_host, int n, float* priorities_host_out);

/* ---- inference: replaces Agent.choose_action (enet_sac.py:547-553, enet_td3.py:207-220 post-warm-up) */
/* obs_host: (S) already concatenated (eig|A); eps_host: (A) N(0,1) (SAC) or additive noise (TD3), or NULL
 * (SAC: device Philox; TD3: no noise).  action_host_out: (A). */
int sc_agent_act(sc_agent* a, const float* obs_host, const float* eps_host, float* action_host_out);

/* ---- update: replaces Agent.learn (enet_sac.py:555-628, enet_td3.py:222-364) and
 *      Agent.update_network_parameters (enet_sac.py:523-542) ------------------------------------------ */
int sc_agent_learn(sc_agent* a, const sc_noise* noise_or_null, sc_learn_stats* stats_host_or_null);
int sc_agent_soft_update(sc_agent* a, float tau);
/* The tensor-core GEMMs read every wide weight matrix as a pair of tf32 "big / small" copies that the library keeps next
 * to the fp32 parameters (written by its own Adam kernels).  Handing out a parameter pointer (sc_agent_tensor with
 * SC_SLOT_PARAM) or sc_agent_soft_update marks the copies stale and the next learn() re-derives them (one small kernel), so
 * load_state_dict-style writes through FRESHLY obtained views need nothing more.  A caller that writes parameters through a
 * pointer it obtained EARLIER (e.g. an NCCL broadcast into a cached view, distributed_per_sac.py:84-90) calls this after
 * the write. */
int sc_agent_params_changed(sc_agent* a);
/* Agent.store_transition (enet_sac.py:544-545) followed by Agent.learn() as ONE call: the packed transition, the noise and
 * the sampling randomness travel in a single pinned block / host->device copy, the kernels replay from a CUDA graph.
 * Arguments as for sc_replay_store_host + sc_agent_learn. */
int sc_agent_store_learn(sc_agent* a, const float* transition_host, int priority_mode, double priority,
                         const sc_noise* noise_or_null, sc_learn_stats* stats_host_or_null);

/* Loss read-back mode of sc_agent_learn / sc_agent_store_learn when `stats_host` != NULL.  0 (default): the call waits for the
 * step and returns ITS losses.  1: deferred -- the call queues the copy of this step's losses behind its kernels and returns the
 * losses of the PREVIOUS step (did_learn = 0 on the first call); no synchronisation, so the host prepares step n+1 (random
 * draws, transition packing) while the device runs step n.  Every step's losses are still read exactly once. */
int sc_agent_set_stats_mode(sc_agent* a, int deferred);

/* ---- ensembles: E independent agents (the reference's seed-ensemble axis, elasticnet/do.sh:1-6: one python process per
 *      seed) stepped in lock-step by ONE launch sequence -- every kernel of learn() runs once with blockIdx.y = agent.
 *      Agents must share the configuration (except rng_seed), device and stream; weights, replay contents and randomness
 *      are per agent, and each agent's results are bit-identical to the same agent stepped alone. ------------------------ */
typedef struct sc_ensemble sc_ensemble;
int sc_ensemble_create(sc_agent** agents, int n, sc_ensemble** out);   /* agents stay owned by the caller and individually usable */
int sc_ensemble_destroy(sc_ensemble* e);
/* noise / stats: arrays of n entries (or NULL) */
int sc_ensemble_learn(sc_ensemble* e, const sc_noise* noise_array_or_null, sc_learn_stats* stats_host_array_or_null);
/* transitions_host: n packed transitions, one per agent; priorities_or_null: n doubles (priority_mode 2 / 3) */
int sc_ensemble_store_learn(sc_ensemble* e, const float* transitions_host, int priority_mode, const double* priorities_or_null,
                            const sc_noise* noise_array_or_null, sc_learn_stats* stats_host_array_or_null);
int64_t sc_ensemble_launch_count(sc_ensemble* e);
/* kernels launched by this library since the handle was created (bench.py "gpu_launches") */
int64_t sc_agent_launch_count(sc_agent* a);
/* CUDA-event time (ms) accumulated since the last call with reset!=0 (after sc_agent_enable_kernel_timing): which = 0 critic
 * row-phase group, 1 critic wgrad+Adam, 2 actor row-phase group, 3 actor wgrad+Adam, 4 PER sample, 5 PER update,
 * 6 / 7 the tensor-core dense launches inside groups 0 / 2 */
int sc_agent_kernel_time(sc_agent* a, int which, int reset, double* ms_total, int64_t* launches);
int sc_agent_enable_kernel_timing(sc_agent* a, int on);

/* =====================================================================================================================
 * Calibration conv-SAC agent: replaces calibration/calib_sac.py:234-423 (Agent with conv Actor/Critic networks,
 * BatchNorm2d in training mode inside learn(), KLD hint penalty).  Same conventions as the sc_agent_* group.
 * ================================================================================================================== */
/* constructor kwargs of calib_sac.Agent (:235-236); input_dims is fixed to [1, img, img] with img = 128 */
typedef struct sc_calib_config {
  int32_t abi_version;   /* = SC_ABI_VERSION */
  int32_t batch_size, mem_size, n_actions;
  int32_t M;             /* sky rows - 1: the sky tensor is (M+1) x 7 (calib_sac.py:30-36) */
  int32_t img;           /* 128 */
  int32_t use_hint;
  int32_t variant;       /* 0 = calibration/calib_sac.py; SC_CALIB_* flags select the demixing agents of the same network family */
  double gamma, tau, lr_actor, lr_critic, alpha;
  float hint_threshold, admm_rho;
  uint64_t rng_seed;
  int32_t update_actor_interval; /* SC_CALIB_TD3 only (demix_td3.py:357,368); ignored otherwise */
  int32_t reserved;
} sc_calib_config;
/* network-family variants (sc_calib_config.variant).  demixing_rl/demix_sac.py:372-527 = LOGSIGMA | DEMIX_HINT | FLAT_META;
 * demixing_fuzzy/demix_sac.py:96-276 with use_influence=False = the same + NO_TRUNK | SIDE32. */
enum {
  SC_CALIB_LOGSIGMA = 1,    /* second policy head is fc22logsigma: sigma = exp(clamp(., -20, 2)) (demix_sac.py:467,492-494) instead of clamp(fc22sigma, 1e-6, 1) */
  SC_CALIB_NO_TRUNK = 2,    /* no conv trunk / influence map (use_influence=False): features = metadata network only                                   */
  SC_CALIB_DEMIX_HINT = 4,  /* KLD hint without clamping and dual update from a fresh no-grad sample of the updated actor (demix_sac.py:636-676)        */
  SC_CALIB_FLAT_META = 8,   /* metadata is a flat vector of M values (demix) instead of the (M+1) x 7 sky table (calib)                                 */
  SC_CALIB_SIDE32 = 16,     /* INNER_DIM = 32 (demixing_fuzzy/demix_sac.py:13) instead of 16; MLP family only                                          */
  /* demixing_rl/demix_td3.py:355-603 (DemixingAgent of the TD3 file) = TD3 | FLAT_META: one tanh policy head `fc22` (:318,342), critic
   * side input cat((action, metadata)) (:279), a target actor (SC_NET_TARGET_ACTOR), eval-mode target passes (:484-486), the scalar
   * target noise (:492-493), prioritized replay with the float64 sum tree inside the arena ("tree" buffer; PER :23-167,
   * absolute_error_upper 100), PER.mse / actor loss with the (B,)x(B,1) broadcast (:137-140,:538), the delayed actor update
   * (update_actor_interval), with use_hint the five-step adaptive-ADMM actor loop (:543-600; admm_rho = the constructor's 0.1, :400),
   * and update_network_parameters over EVERY state_dict entry incl. BatchNorm buffers (:408-432), on actor steps only (:603).
   * sc_calib_learn draws the minibatch itself: noise->per_u01_host (B float64 uniforms) + noise->td3_target_noise reproduce the
   * reference's numpy stream, NULL selects device Philox; sc_calib_set_batch_host is not used.  sc_calib_act: eps_host = the additive
   * exploration noise (n_actions values, :451) or NULL for none; the warm-up branch (:440-441) is host-side. */
  SC_CALIB_TD3 = 32
};
typedef struct sc_calib sc_calib;

size_t sc_calib_arena_bytes(const sc_calib_config* cfg);
int sc_calib_create(const sc_calib_config* cfg, void* arena, size_t arena_bytes, void* cuda_stream, sc_calib** out);
int sc_calib_destroy(sc_calib* a);
int sc_calib_sync(sc_calib* a);
/* `net` = SC_NET_ACTOR / SC_NET_CRITIC_{1,2} / SC_NET_TARGET_CRITIC_{1,2} (/ SC_NET_TARGET_ACTOR, TD3); `name` = a calib_sac state_dict key
 * ("conv2.weight" as [Cout][Cin*25], "bn1.running_var", "bn3.num_batches_tracked" (int64), "fc21.weight", ...).
 * slot 3 = gradient of the last backward pass (tests). */
int sc_calib_tensor(sc_calib* a, int net, int slot, const char* name, sc_tensor_info* out);
/* replay arrays "state_memory_img", "state_memory_sky", "new_state_memory_img", "new_state_memory_sky",
 * "action_memory", "hint_memory", "reward_memory", "terminal_memory" (float 0/1) and a few workspace buffers */
int sc_calib_buffer(sc_calib* a, const char* name, sc_tensor_info* out);
int sc_calib_get_counters(sc_calib* a, int64_t* mem_cntr, int64_t* learn_counter);
int sc_calib_set_counters(sc_calib* a, int64_t mem_cntr, int64_t learn_counter);
int sc_calib_set_scalars(sc_calib* a, float rho, float admm_rho);
/* ReplayBuffer.store_transition (calib_sac.py:44-56): img (img*img), sky ((M+1)*7), action / hint (n_actions), host */
int sc_calib_store_host(sc_calib* a, const float* img, const float* sky, const float* img2, const float* sky2, const float* action,
                        const float* hint, float reward, int done);
/* indices of np.random.choice(max_mem, B, replace=False) (calib_sac.py:60) */
int sc_calib_set_batch_host(sc_calib* a, const int32_t* idx_host, int n);
/* Agent.choose_action (calib_sac.py:285-295): actor in eval() mode; eps (n_actions) N(0,1) or NULL for device Philox */
int sc_calib_act(sc_calib* a, const float* img_host, const float* sky_host, const float* eps_host, float* action_host_out);
/* Agent.learn (calib_sac.py:300-378); noise->eps_next_host / eps_cur_host (B, n_actions) or NULL for device Philox */
int sc_calib_learn(sc_calib* a, const sc_noise* noise_or_null, sc_learn_stats* stats_host_or_null);
int sc_calib_soft_update(sc_calib* a, float tau);   /* update_network_parameters (calib_sac.py:271-276) */
int64_t sc_calib_launch_count(sc_calib* a);
/* SC_CALIB_TD3: PER.beta (demix_td3.py:32,95) lives outside the arena; get / set for checkpoints (PER.load_checkpoint :147-160) */
int sc_calib_get_per_beta(sc_calib* a, double* beta_out);
int sc_calib_set_per_beta(sc_calib* a, double beta);

#ifdef __cplusplus
}
#endif
#endif /* SMARTCAL_B200_H */
