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
// (included by sc_rowphase.cu once per build configuration; see SC_CFG_* there)
namespace sc {
namespace SC_CFG_NS {

constexpr int KC = SC_CFG_KC;            // reduction elements per ring slot
constexpr int KBLK = KC / 32;            // 32-float (128-byte) k-blocks per slot
constexpr int NSTAGE = SC_CFG_NSTAGE;    // ring slots: the producer runs up to NSTAGE-1 slots ahead of the consumers
constexpr int RT = SC_CFG_RT;            // rows a CTA carries through the row-local tail at once
constexpr int REDROWS = SC_CFG_REDROWS;  // max rows of a GEMM tile (8*TM)
constexpr int GEMV_U = SC_CFG_MINBLOCKS == 2 ? 1 : 2;   // tail GEMV items in flight per warp (register budget)
constexpr int WCOLS = 64;                // max output features (dgrad: input columns) of a CTA's slice
constexpr int PROW = KC < 128 ? KC : 128;   // rows of a dgrad weight panel box
constexpr int SLOT_A = KBLK * REDROWS * 32;  // [k-block][tile rows][32]
constexpr int SLOT_W = KC * WCOLS;           // forward: [k-block][<=64 features][32]; dgrad: [KC/PROW][<=2][PROW][32]
constexpr int SLOT = SLOT_A + SLOT_W;

constexpr int MAXOPS = 8;
constexpr int NPMAX = 16;          // max LayerNorm partials per row (one per CTA of the cluster)
constexpr int MAXROWS = SC_CFG_MAXROWS;   // max slab rows per cluster

struct DenseOp {                          // Z = X W^T + b, then LayerNorm + ELU
  const void *mapA, *mapW;               // tensor maps of X [rows][K] (box: tile rows) and W [N][K] (box: the CTA's feature slice); K padded to 32
  const float *bias, *gamma, *beta;
  int K, N;
  float *H, *ZH, *RS; int ldo;
};
struct DgradOp {                          // dH_in = DZ W, then back through ELU / LayerNorm of the producing layer
  const void *mapA, *mapP;               // tensor maps of DZ [rows][Nl] and the W panel ([Nl][Kin], box 32 x PROW)
  int Nl, Kin;
  const float *Hp, *ZHp, *RSp, *gp;      // layer that produced the input (width Kin)
  float *DU, *DZo; int ldo;
};

// dynamic shared memory map (floats): operand ring | split-reduction partials | output tile | LN statistics
constexpr int SM_RING = NSTAGE * SLOT;
constexpr int SM_RED = 160 * REDROWS;                  // max over (KS, NC) in {(8,16),(4,32),(2,64)} of KS * REDROWS * (NC + 4)
constexpr int SM_Z = REDROWS * WCOLS;
constexpr int SM_STAT = MAXOPS * MAXROWS * 2;
constexpr int SM_ZK = SC_CFG_ZK;                       // raw pre-LayerNorm outputs of a stage, kept on chip between the two barriers
constexpr int SM_TOTAL_FLOATS = SM_RING + SM_RED + SM_Z + SM_STAT + SM_ZK;

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

// valid columns of LayerNorm partial `rk` (= the slice of CTA rk): clamp(N - rk*Nc, 0, Nc)
__device__ __forceinline__ int part_count(int N, int Nc, int rk) { return max(0, min(Nc, N - rk * Nc)); }

struct Pipe {
  int dbg;
  long long* prof;   // optional cycle accumulators of thread 0: [0] full-wait, [1] mma, [2] park+combine, [3] tile epilogue, [4] LN combine, [5] LN normalise
  uint64_t* full;     // [NSTAGE] armed by the producer (expect_tx), completed by the bulk copies
  uint64_t* empty;    // [NSTAGE] one arrival per consumer warp
  unsigned long long* tl;   // optional timeline of CTA 0 / thread 0 (globaltimer), tl[0] = next free index
};
__device__ __forceinline__ void tl_mark(const Pipe& pipe, int tag) {
  if (pipe.tl != nullptr && threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long i = pipe.tl[0];
    if (i < 500) { pipe.tl[1 + i] = (t << 8) | (unsigned long long)tag; pipe.tl[0] = i + 1; }
  }
}

// One output tile = all reduction chunks of (op, row sub-tile) for the CTA's whole column slice (MT feature tiles of 16).
// The 8 consumer warps are (feature tile mt = warp % MT, reduction part kp = warp / MT).  `g` counts chunks since the
// kernel began (ring slot = g % NSTAGE, mbarrier phase parity = (g / NSTAGE) & 1); producer and consumers walk the same sequence.
template <int NT8, int MT>
__device__ __forceinline__ void fwd_tile(const float* ring, const Pipe& pipe, unsigned& g, int K, float* red, float* sZ,
                                         const float* bias_slice, int ncv) {
  constexpr int KS = SC_WARPS / MT, NC = 16 * MT;
  Acc3 acc3[NT8];
#pragma unroll
  for (int i = 0; i < NT8; ++i) acc3_zero(acc3[i]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp % MT, kp = warp / MT;
  const bool prof = pipe.prof != nullptr && threadIdx.x == 0;
  tl_mark(pipe, 1);
  for (int k0 = 0; k0 < K; k0 += KC, ++g) {
    const unsigned slot = g % NSTAGE;
    long long t0 = prof ? clock64() : 0;
    mbar_wait(pipe.full + slot, (g / NSTAGE) & 1u);
    long long t1 = prof ? clock64() : 0;
    const float* sA = ring + slot * SLOT;
    const int nk8 = min(KC, K - k0) >> 3;
    if (!(pipe.dbg & 1)) mma3_nt<NT8>(sA, 8 * NT8, sA + SLOT_A, NC, mt, (nk8 * kp) / KS, (nk8 * (kp + 1)) / KS, acc3);
    __syncwarp();
    if (lane == 0) mbar_arrive(pipe.empty + slot);
    if (prof) { pipe.prof[0] += t1 - t0; pipe.prof[1] += clock64() - t1; }
  }
  long long t2 = prof ? clock64() : 0;
  tl_mark(pipe, 2);
  float acc[NT8][4];
#pragma unroll
  for (int i = 0; i < NT8; ++i) acc3_fold(acc3[i], acc[i]);
  park_partial<NT8>(red, kp, mt, NC, acc);
  consumer_sync();
  combine_partials<NT8>(red, KS, NC, sZ, bias_slice, ncv);
  consumer_sync();
  if (prof) pipe.prof[2] += clock64() - t2;
  tl_mark(pipe, 3);
}


template <int NT8, int MT>
__device__ __forceinline__ void bwd_tile(const float* ring, const Pipe& pipe, unsigned& g, int Nl, float* red, float* sZ) {
  constexpr int KS = SC_WARPS / MT, NC = 16 * MT, NCH = (NC + 31) / 32;
  Acc3 acc3[NT8];
#pragma unroll
  for (int i = 0; i < NT8; ++i) acc3_zero(acc3[i]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp % MT, kp = warp / MT;
  for (int n0 = 0; n0 < Nl; n0 += KC, ++g) {
    const unsigned slot = g % NSTAGE;
    mbar_wait(pipe.full + slot, (g / NSTAGE) & 1u);
    const float* sA = ring + slot * SLOT;
    const int nk8 = min(KC, Nl - n0) >> 3;
    mma3_nn<NT8, PROW>(sA, 8 * NT8, sA + SLOT_A, NCH, mt, (nk8 * kp) / KS, (nk8 * (kp + 1)) / KS, acc3);
    __syncwarp();
    if (lane == 0) mbar_arrive(pipe.empty + slot);
  }
  float acc[NT8][4];
#pragma unroll
  for (int i = 0; i < NT8; ++i) acc3_fold(acc3[i], acc[i]);
  park_partial<NT8>(red, kp, mt, NC, acc);
  consumer_sync();
  combine_partials<NT8>(red, KS, NC, sZ, nullptr, 0);
  consumer_sync();
}

// ======================================================================================================
// cluster-tiled forward stage
// ======================================================================================================
template <int CS, int TM>
__device__ void dense_stage(const DenseOp* ops, int nops, int B, int row0, int nrows, const int* __restrict__ bidx,
                            float* part, float* smem, const Pipe& pipe, unsigned& g, unsigned long long* ts = nullptr) {
  constexpr int ROWS = 8 * TM;
  float* ring = smem;
  float* red = ring + SM_RING;
  float* sZ = red + SM_RED;
  float* stat = sZ + SM_Z;
  float* zk = stat + SM_STAT;
  const int rank = (int)cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsub = (nrows + ROWS - 1) / ROWS;
  // raw z of every (op, row, own column) stays in shared memory when it fits, else it is parked in H (global)
  int zk_need = 0;
  for (int o = 0; o < nops; ++o) zk_need += nsub * ROWS * round4((ops[o].N + CS - 1) / CS);
  const bool use_zk = zk_need <= SM_ZK;
  auto stamp = [&](int i) {
    if (ts != nullptr && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      ts[i] = t;
    }
  };

  if (warp == SC_WARPS) {
    // ---------------- TMA producer warp: streams every reduction chunk of the stage through the ring ----------------
    if (nrows > 0) {
      for (int o = 0; o < nops; ++o) {
        const DenseOp op = ops[o];
        const int Nc = round4((op.N + CS - 1) / CS);      // 16, 32 or 64 (checked at create)
        const int n0 = rank * Nc;
        for (int sub = 0; sub < nsub; ++sub) {
          const int r0 = row0 + sub * ROWS;
          for (int k0 = 0; k0 < op.K; k0 += KC, ++g) {
            const unsigned slot = g % NSTAGE;
            const int nkb = (min(KC, op.K - k0) + 31) >> 5;               // 32-float k-blocks in this chunk
            mbar_wait(pipe.empty + slot, ((g / NSTAGE) & 1u) ^ 1u);
            float* sA = ring + slot * SLOT;
            if (lane == 0) mbar_expect_tx(pipe.full + slot, (unsigned)nkb * (unsigned)(ROWS + Nc) * 128u);
            __syncwarp();
            // lanes 0..KBLK-1: X boxes [ROWS][32], lanes KBLK..2KBLK-1: W boxes [Nc][32]; rows beyond the matrix are zero-filled
            if (lane < nkb) tma_load_2d(sA + lane * ROWS * 32, op.mapA, k0 + 32 * lane, r0, pipe.full + slot);
            else if (lane >= KBLK && lane - KBLK < nkb)
              tma_load_2d(sA + SLOT_A + (lane - KBLK) * Nc * 32, op.mapW, k0 + 32 * (lane - KBLK), n0, pipe.full + slot);
          }
        }
      }
    }
  } else if (nrows > 0) {
    // ---------------- consumer warps ----------------
    int zbase = 0;
    for (int o = 0; o < nops; ++o) {
      const DenseOp op = ops[o];
      const int Nc = round4((op.N + CS - 1) / CS);
      float* zko = zk + zbase;
      zbase += nsub * ROWS * Nc;
      const int n0 = rank * Nc;
      const int ncv = max(0, min(Nc, op.N - n0));
      const float* bias_slice = op.bias ? op.bias + n0 : nullptr;
      for (int sub = 0; sub < nsub; ++sub) {
        const int r0 = row0 + sub * ROWS;
        const int rv = min(ROWS, nrows - sub * ROWS);
        {
          switch (Nc >> 4) {
            case 4: fwd_tile<TM, 4>(ring, pipe, g, op.K, red, sZ, bias_slice, ncv); break;
            case 2: fwd_tile<TM, 2>(ring, pipe, g, op.K, red, sZ, bias_slice, ncv); break;
            default: fwd_tile<TM, 1>(ring, pipe, g, op.K, red, sZ, bias_slice, ncv); break;
          }
        }
        // raw z -> zk / H (own column slice) and per-row (mean, M2) over the valid columns of the slice
        const long long te = (pipe.prof && tid == 0) ? clock64() : 0;
        const int row = tid >> 3, sub8 = tid & 7;
        float s = 0.f;
        if (row < rv)
          for (int c = sub8; c < ncv; c += 8) s += sZ[row * Nc + c];
        s = group8_sum(s);
        const float mean = ncv > 0 ? s / (float)ncv : 0.f;
        float m2 = 0.f;
        if (row < rv)
          for (int c = sub8; c < ncv; c += 8) {
            const float d = sZ[row * Nc + c] - mean;
            m2 = fmaf(d, d, m2);
          }
        m2 = group8_sum(m2);
        if (row < rv) {
          if (use_zk) {
            float* zrow = zko + (sub * ROWS + row) * Nc;
            for (int c = sub8; c < ncv; c += 8) zrow[c] = sZ[row * Nc + c];
          } else {
            float* hrow = op.H + (size_t)(r0 + row) * op.ldo + n0;
            for (int c = sub8; c < ncv; c += 8) hrow[c] = sZ[row * Nc + c];
          }
          if (sub8 == 0) {
            float* pp = part + (((size_t)o * B + (r0 + row)) * NPMAX + rank) * 2;
            pp[0] = mean;
            pp[1] = m2;
          }
        }
        if (pipe.prof && tid == 0) pipe.prof[3] += clock64() - te;
        tl_mark(pipe, 4);
      }
    }
  }
  stamp(0);
  tl_mark(pipe, 5);
  cluster_sync();
  tl_mark(pipe, 6);
  stamp(1);
  const long long tf0 = (pipe.prof && tid == 0) ? clock64() : 0;
  if (warp < SC_WARPS) {
    // ---- combine the LayerNorm partials of every (op,row) in fixed order (Chan et al.) ----
    for (int i = tid; i < nops * nrows; i += SC_THREADS) {
      const int o = i / nrows, r = i - o * nrows;
      const DenseOp& op = ops[o];
      const int Nc = round4((op.N + CS - 1) / CS);
      const float* pp = part + (((size_t)o * B + (row0 + r)) * NPMAX) * 2;
      // all partials are fetched first (independent L2 loads in flight together), then combined in fixed order
      float2 pv[CS];
#pragma unroll
      for (int p = 0; p < CS; ++p) pv[p] = __ldcg(reinterpret_cast<const float2*>(pp) + p);
      float msum = 0.f;
#pragma unroll
      for (int p = 0; p < CS; ++p) msum = fmaf((float)part_count(op.N, Nc, p), pv[p].x, msum);
      const float mean = msum / (float)op.N;
      float m2 = 0.f;
#pragma unroll
      for (int p = 0; p < CS; ++p) {
        const int cnt = part_count(op.N, Nc, p);
        const float d = pv[p].x - mean;
        if (cnt > 0) m2 += pv[p].y + (float)cnt * d * d;
      }
      const float rstd = 1.0f / sqrtf(m2 / (float)op.N + SC_LN_EPS);
      stat[2 * i] = mean;
      stat[2 * i + 1] = rstd;
      if (rank == 0 && op.RS != nullptr) op.RS[row0 + r] = rstd;
    }
    consumer_sync();
    if (pipe.prof && tid == 0) pipe.prof[4] += clock64() - tf0;
    // normalise + affine + ELU on the own column slice (raw z read back from shared memory, or from H when it did not fit)
    int zbase = 0;
    for (int o = 0; o < nops; ++o) {
      const DenseOp& op = ops[o];
      const int Nc = round4((op.N + CS - 1) / CS);
      const int n0 = rank * Nc;
      const float* zko = zk + zbase;
      zbase += nsub * ROWS * Nc;
      for (int e = tid; e < nrows * Nc; e += SC_THREADS) {
        const int r = e / Nc, c = e - r * Nc, n = n0 + c;
        if (n < op.N) {
          const size_t off = (size_t)(row0 + r) * op.ldo + n;
          const float z = use_zk ? zko[r * Nc + c] : ld_act(op.H + off);     // (sub*ROWS + row) == r
          const float zh = (z - stat[2 * (o * nrows + r)]) * stat[2 * (o * nrows + r) + 1];
          const float u = fmaf(zh, ld_w(op.gamma + n), ld_w(op.beta + n));
          op.H[off] = elu(u);
          if (op.ZH != nullptr) op.ZH[off] = zh;
        }
      }
    }
  }
  if (pipe.prof && tid == 0) pipe.prof[5] += clock64() - tf0;
  stamp(2);
  tl_mark(pipe, 7);
  cluster_sync();
  tl_mark(pipe, 8);
  stamp(3);
}

// ======================================================================================================
// cluster-tiled dgrad stage: dH_in = DZ * W, then back through ELU and LayerNorm of the producing layer
// ======================================================================================================
template <int CS, int TM>
__device__ void dgrad_stage(const DgradOp* ops, int nops, int B, int row0, int nrows, float* part, float* smem,
                            const Pipe& pipe, unsigned& g) {
  constexpr int ROWS = 8 * TM;
  float* ring = smem;
  float* red = ring + SM_RING;
  float* sZ = red + SM_RED;
  float* stat = sZ + SM_Z;
  const int rank = (int)cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsub = (nrows + ROWS - 1) / ROWS;
  // per (op, sub-tile): [h | zh] slices of the producing layer, ROWS x WCOLS floats each, in the (idle) raw-z area
  float* stash = stat + SM_STAT;
  const bool use_stash = 2 * nops * nsub * ROWS * WCOLS <= SM_ZK;

  if (warp == SC_WARPS) {
    if (nrows > 0) {
      for (int o = 0; o < nops; ++o) {
        const DgradOp op = ops[o];
        const int Kc = round4((op.Kin + CS - 1) / CS);     // 16, 32 or 64 input columns per CTA
        const int nch = (Kc + 31) / 32;
        const int k0 = rank * Kc;
        for (int sub = 0; sub < nsub; ++sub) {
          const int r0 = row0 + sub * ROWS;
          for (int n0 = 0; n0 < op.Nl; n0 += KC, ++g) {
            const unsigned slot = g % NSTAGE;
            const int nn = min(KC, op.Nl - n0);
            const int nkb = (nn + 31) >> 5;
            const int npan = (nn + PROW - 1) / PROW;                          // weight panels (PROW reduction rows each)
            mbar_wait(pipe.empty + slot, ((g / NSTAGE) & 1u) ^ 1u);
            float* sA = ring + slot * SLOT;
            if (lane == 0) mbar_expect_tx(pipe.full + slot, (unsigned)nkb * ROWS * 128u + (unsigned)(npan * nch) * PROW * 128u);
            __syncwarp();
            if (lane < nkb) tma_load_2d(sA + lane * ROWS * 32, op.mapA, n0 + 32 * lane, r0, pipe.full + slot);
            else if (lane >= KBLK && lane - KBLK < npan * nch) {             // panel (pn, cc): [PROW n][32 input columns]
              const int q = lane - KBLK, pn = q / nch, cc = q - pn * nch;
              tma_load_2d(sA + SLOT_A + q * PROW * 32, op.mapP, k0 + 32 * cc, n0 + PROW * pn, pipe.full + slot);
            }
          }
        }
      }
    }
  } else if (nrows > 0) {
    for (int o = 0; o < nops; ++o) {
      const DgradOp op = ops[o];
      const int Kc = round4((op.Kin + CS - 1) / CS);
      const int k0 = rank * Kc;
      const int kcv = max(0, min(Kc, op.Kin - k0));
      for (int sub = 0; sub < nsub; ++sub) {
        const int r0 = row0 + sub * ROWS;
        const int rv = min(ROWS, nrows - sub * ROWS);
        // the producing layer's h and zh slices of this tile do not depend on the GEMM: fetch them into shared memory now
        // (cp.async, no registers held) so that their L2 latency runs under the k-loop; they stay there for the second pass
        float* st_h = stash + (size_t)((o * nsub + sub) * 2) * ROWS * WCOLS;
        float* st_z = st_h + ROWS * WCOLS;
        if (use_stash) {
          const int q4 = Kc >> 2;
          for (int e = tid; e < 2 * rv * q4; e += SC_THREADS) {
            const int arr = e / (rv * q4), f = e - arr * rv * q4, row = f / q4, c = (f - row * q4) * 4;
            if (c < kcv) {
              const float* src = (arr ? op.ZHp : op.Hp) + (size_t)(r0 + row) * op.ldo + k0 + c;
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32((arr ? st_z : st_h) + row * Kc + c)), "l"(src) : "memory");
            }
          }
          asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        switch (Kc >> 4) {
          case 4: bwd_tile<TM, 4>(ring, pipe, g, op.Nl, red, sZ); break;
          case 2: bwd_tile<TM, 2>(ring, pipe, g, op.Nl, red, sZ); break;
          default: bwd_tile<TM, 1>(ring, pipe, g, op.Nl, red, sZ); break;
        }
        if (use_stash) {
          asm volatile("cp.async.wait_all;\n" ::: "memory");
          consumer_sync();
        }
        const int row = tid >> 3, sub8 = tid & 7;
        float s1 = 0.f, s2 = 0.f;
        if (row < rv) {
          const size_t base = (size_t)(r0 + row) * op.ldo + k0;
          for (int c = sub8; c < kcv; c += 8) {
            const float dh = sZ[row * Kc + c];
            const float h = use_stash ? st_h[row * Kc + c] : ld_act(op.Hp + base + c);
            const float zh = use_stash ? st_z[row * Kc + c] : ld_act(op.ZHp + base + c);
            const float du = dh * elu_grad_from_out(h);
            const float t = du * ld_w(op.gp + k0 + c);
            s1 += t;
            s2 = fmaf(t, zh, s2);
            op.DU[base + c] = du;
            if (use_stash) st_h[row * Kc + c] = t;          // (h is not needed again: its slot carries t to the second pass)
            else op.DZo[base + c] = t;
          }
        }
        s1 = group8_sum(s1);
        s2 = group8_sum(s2);
        if (row < rv && sub8 == 0) {
          float* pp = part + (((size_t)o * B + (r0 + row)) * NPMAX + rank) * 2;
          pp[0] = s1;
          pp[1] = s2;
        }
      }
    }
  }
  cluster_sync();
  if (warp < SC_WARPS) {
    float* rsv = red;                                       // 1/std of the producing LayerNorm per (op, row); `red` is idle now
    for (int i = tid; i < nops * nrows; i += SC_THREADS) {
      const int o = i / nrows, r = i - o * nrows;
      const DgradOp& op = ops[o];
      const int Kc = round4((op.Kin + CS - 1) / CS);
      const float* pp = part + (((size_t)o * B + (row0 + r)) * NPMAX) * 2;
      float2 pv[CS];
#pragma unroll
      for (int p = 0; p < CS; ++p) pv[p] = __ldcg(reinterpret_cast<const float2*>(pp) + p);
      const float rs = ld_act(op.RSp + row0 + r);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int p = 0; p < CS; ++p) {
        if (part_count(op.Kin, Kc, p) > 0) {
          s1 += pv[p].x;
          s2 += pv[p].y;
        }
      }
      stat[2 * i] = s1 / (float)op.Kin;
      stat[2 * i + 1] = s2 / (float)op.Kin;
      rsv[i] = rs;
    }
    consumer_sync();
    for (int o = 0; o < nops; ++o) {
      const DgradOp& op = ops[o];
      const int Kc = round4((op.Kin + CS - 1) / CS);
      const int k0 = rank * Kc;
      for (int e = tid; e < nrows * Kc; e += SC_THREADS) {
        const int r = e / Kc, c = e - r * Kc, k = k0 + c;
        if (k < op.Kin) {
          const size_t off = (size_t)(row0 + r) * op.ldo + k;
          float t, zh;
          if (use_stash) {
            const int sub = r / ROWS, row = r - sub * ROWS;
            const float* st_h = stash + (size_t)((o * nsub + sub) * 2) * ROWS * WCOLS;
            t = st_h[row * Kc + c];
            zh = st_h[ROWS * WCOLS + row * Kc + c];
          } else {
            t = ld_act(op.DZo + off);
            zh = ld_act(op.ZHp + off);
          }
          op.DZo[off] = rsv[o * nrows + r] * (t - stat[2 * (o * nrows + r)] - zh * stat[2 * (o * nrows + r) + 1]);
        }
      }
    }
  }
  // no trailing cluster barrier: consumers are in the next kernel
}

// ======================================================================================================
// row-local helpers (RT rows at a time, vectors in shared memory)
// ======================================================================================================
struct RowSet {          // the RT rows a CTA is carrying through the tail
  int row[RT];           // global batch row, or -1
};

// shared-memory map of the tail (floats); lives in the same dynamic buffer as the GEMM ring (different stages)
constexpr int NETS = 4;   // critic phase: target critics 0,1 and online critics 2,3; actor phase: online critics 0,1
struct TailSm {
  float *xh2, *zh2r;              // [RT][H2]   actor layer-2 output rows (+ normalised, actor phase)
  float *x3, *zh3;                // [RT][H3]
  float *head;                    // [RT][2*AMAX] mu | raw logsigma (later: d mu | d logsigma)
  float *an;                      // [2][RT][AMAX] actions fed to the critics: [0] policy sample, [1] replay action
  float *smp;                     // [RT][AMAX][4]: tanh(u), sigma, eps, clamp-pass mask
  float *lp;                      // [RT][AMAX] per-action log-prob terms / d loss / d action accumulator
  float *sc;                      // [RT][16] per-row scalars
  float *rsx;                     // [RT][8]: [2] rs3, [3] rs2, [4+k] rs12 of online critic k
  float *h12r;                    // [NETS][RT][H2] state-branch outputs per net
  float *zh12r;                   // [2][RT][H2]
  float *v21, *zh21;              // [NETS][RT][HA1]
  float *v22, *zh22;              // [NETS][RT][HA2]
  float *rsn;                     // [NETS][RT][2] 1/std of layers 21, 22
  float *qv;                      // [NETS][RT]
  float *ga, *gb;                 // [2][RT][H2]  d u12 / d z12   (actor phase: [RT][H3|H2] scratch)
  float *gc, *gd;                 // [2][RT][HA1] d u21 / d z21
  float *ge, *gf;                 // [2][RT][HA2] d u22 / d z22
  float *scratch;                 // gemvT partials (2048*RT floats)
};
__device__ TailSm carve_tail(float* sm) {
  TailSm t;
  t.xh2 = sm; sm += RT * H2;
  t.zh2r = sm; sm += RT * H2;
  t.x3 = sm; sm += RT * H3;
  t.zh3 = sm; sm += RT * H3;
  t.head = sm; sm += RT * 2 * AMAX;
  t.an = sm; sm += 2 * RT * AMAX;
  t.smp = sm; sm += RT * AMAX * 4;
  t.lp = sm; sm += RT * AMAX;
  t.sc = sm; sm += RT * 16;
  t.rsx = sm; sm += RT * 8;
  t.h12r = sm; sm += NETS * RT * H2;
  t.zh12r = sm; sm += 2 * RT * H2;
  t.v21 = sm; sm += NETS * RT * HA1;
  t.zh21 = sm; sm += NETS * RT * HA1;
  t.v22 = sm; sm += NETS * RT * HA2;
  t.zh22 = sm; sm += NETS * RT * HA2;
  t.rsn = sm; sm += NETS * RT * 2;
  t.qv = sm; sm += NETS * RT;
  t.ga = sm; sm += 2 * RT * H2;
  t.gb = sm; sm += 2 * RT * H2;
  t.gc = sm; sm += 2 * RT * HA1;
  t.gd = sm; sm += 2 * RT * HA1;
  t.ge = sm; sm += 2 * RT * HA2;
  t.gf = sm; sm += 2 * RT * HA2;
  t.scratch = sm; sm += 2048 * RT;
  return t;
}

enum { SC_LOGP = 0, SC_Y = 3, SC_DQ1 = 6, SC_DQ2 = 7, SC_GLP = 8 };

// ---- multi-net row helpers: NR = nn*RT rows, row index rix = net*RT + t; warps stride over the rows ---------------
// in place z -> ELU(LN(z)), keeps the normalised value and 1/std
__device__ void ln_elu_multi(int nn, float* z, float* zh, int N, const float* const* gamma, const float* const* beta, float* rs,
                             int rs_stride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rix = warp; rix < nn * RT; rix += SC_WARPS) {
    const int net = rix / RT;
    float* zr = z + rix * N;
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s += zr[j];
    const float mean = warp_sum(s) / (float)N;
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
      zr[j] = elu(fmaf(n, ld_w(g + j), ld_w(b + j)));
    }
    if (lane == 0) rs[rix * rs_stride] = rstd;
  }
}
// dh -> du (in place), dz; per-row h, zh, rs given as arrays indexed like dh
__device__ void ln_bwd_multi(int nn, float* dh, const float* h, const float* zh, int N, const float* const* gamma, const float* rs,
                             int rs_stride, float* dz) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rix = warp; rix < nn * RT; rix += SC_WARPS) {
    const float* g = gamma[rix / RT];
    const int o = rix * N;
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float du = dh[o + j] * elu_grad_from_out(h[o + j]);
      dh[o + j] = du;
      const float t = du * ld_w(g + j);
      s1 += t;
      s2 = fmaf(t, zh[o + j], s2);
    }
    const float c1 = warp_sum(s1) / (float)N, c2 = warp_sum(s2) / (float)N;
    const float r = rs[rix * rs_stride];
    for (int j = lane; j < N; j += 32) dz[o + j] = r * (dh[o + j] * ld_w(g + j) - c1 - zh[o + j] * c2);
  }
}
// zout[net][t][n] = bias_net[n] + sum_k W_net[n][k] * xin[net][t][k].  A warp works on 4 output features at a time
// (lane / 8) with the reduction split over 8 lanes (128 contiguous bytes of a weight row per load instruction), and on
// U such items at once, so every lane has U*K/32 independent 16-byte weight loads in flight: the tail is bound by the
// latency of these L2 reads, not by arithmetic.
template <int K, int U>
__device__ void gemv_multi(int nn, const float* const* W, int ldw, const float* const* bias, int N, const float* xin, float* zout) {
  constexpr int KI = K / 32;                       // float4 loads per lane per output feature
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 7, q = lane >> 3;
  const int groups = (N + 3) >> 2, items = nn * groups;
  for (int base = warp; base < items; base += SC_WARPS * U) {
    float4 w[U][KI];
    int net[U], n[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int item = base + u * SC_WARPS;
      const bool ok = item < items;
      net[u] = ok ? item / groups : 0;
      n[u] = ok ? (item - net[u] * groups) * 4 + q : N;
      const float* wr = W[net[u]] + (size_t)min(n[u], N - 1) * ldw + 4 * sub;
#pragma unroll
      for (int i = 0; i < KI; ++i) w[u][i] = n[u] < N ? ld_w4(wr + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float* x0 = xin + net[u] * RT * K + 4 * sub;
      float acc[RT];
#pragma unroll
      for (int t = 0; t < RT; ++t) {
        acc[t] = 0.f;
#pragma unroll
        for (int i = 0; i < KI; ++i) acc[t] = dot4(w[u][i], *reinterpret_cast<const float4*>(x0 + t * K + 32 * i), acc[t]);
        acc[t] = group8_sum(acc[t]);
      }
      if (sub == 0 && n[u] < N) {
        const float* bn = bias[net[u]];
        const float b = bn ? ld_w(bn + n[u]) : 0.f;
#pragma unroll
        for (int t = 0; t < RT; ++t) zout[(net[u] * RT + t) * N + n[u]] = acc[t] + b;
      }
    }
  }
}
// out[net][t][k] = sum_n din[net][t][n] * W_net[n][k]; threads split (net, part of n, k4); fixed-order combine; 2 barriers
__device__ void gemvT_multi(int nn, const float* const* W, int ldw, int N, int K, const float* din, float* out, float* scratch) {
  const int K4 = K >> 2;
  const int tpn = SC_THREADS / nn;                 // threads per net
  const int parts = tpn / K4;
  const int tid = threadIdx.x;
  const int net = tid / tpn, tl = tid - net * tpn;
  const int part = tl / K4, k4 = tl - part * K4;
  if (net < nn && part < parts) {
    const float* Wn = W[net];
    const float* d0 = din + net * RT * N;
    float4 acc[RT];
#pragma unroll
    for (int t = 0; t < RT; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int n = part; n < N; n += parts) {
      const float4 w = ld_w4(Wn + (size_t)n * ldw + 4 * k4);
#pragma unroll
      for (int t = 0; t < RT; ++t) {
        const float d = d0[t * N + n];
        acc[t].x = fmaf(d, w.x, acc[t].x);
        acc[t].y = fmaf(d, w.y, acc[t].y);
        acc[t].z = fmaf(d, w.z, acc[t].z);
        acc[t].w = fmaf(d, w.w, acc[t].w);
      }
    }
#pragma unroll
    for (int t = 0; t < RT; ++t)
      *reinterpret_cast<float4*>(scratch + (((size_t)net * parts + part) * RT + t) * K + 4 * k4) = acc[t];
  }
  consumer_sync();
  for (int e = tid; e < nn * RT * K; e += SC_THREADS) {
    const int rix = e / K, k = e - rix * K;
    const int nt = rix / RT, t = rix - nt * RT;
    float s = 0.f;
    for (int pp = 0; pp < parts; ++pp) s += scratch[(((size_t)nt * parts + pp) * RT + t) * K + k];
    out[e] = s;
  }
  consumer_sync();
}

__device__ __forceinline__ void load_rows_to(float* dst, const float* src, int N, const RowSet& rs) {   // dst [RT][N]
  for (int e = threadIdx.x; e < RT * (N >> 2); e += SC_THREADS) {
    const int t = e / (N >> 2), c4 = e - t * (N >> 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rs.row[t] >= 0) v = ld_act4(src + (size_t)rs.row[t] * N + 4 * c4);
    *reinterpret_cast<float4*>(dst + t * N + 4 * c4) = v;
  }
}
__device__ __forceinline__ void store_rows_from(float* dst, const float* src, int N, const RowSet& rs) {   // src [RT][N]
  if (dst == nullptr) return;
  for (int e = threadIdx.x; e < RT * (N >> 2); e += SC_THREADS) {
    const int t = e / (N >> 2), c4 = e - t * (N >> 2);
    if (rs.row[t] >= 0)
      *reinterpret_cast<float4*>(dst + (size_t)rs.row[t] * N + 4 * c4) = *reinterpret_cast<const float4*>(src + t * N + 4 * c4);
  }
}

// action branches (fc21 -> LN -> ELU -> fc22 -> LN -> ELU) of `nn` critics at once; net i reads actions an_sel[i]
// ([RT][AMAX]).  Leaves h21/zh21/h22/zh22 in t.v21.. and 1/std in t.rsn[(net*RT+t)*2 + {0,1}].  4 barriers.
__device__ void action_branches(int nn, const CriticP* const* c, const float* const* an_sel, int A, int Ap, const TailSm& t) {
  for (int e = threadIdx.x; e < nn * RT * HA1; e += SC_THREADS) {
    const int rix = e / HA1, n = e - rix * HA1, net = rix / RT, tt = rix - net * RT;
    const float* w = c[net]->w21 + n * Ap;
    const float* a = an_sel[net] + tt * AMAX;
    float acc = 0.f;
    for (int j = 0; j < A; ++j) acc = fmaf(ld_w(w + j), a[j], acc);
    t.v21[e] = acc + ld_w(c[net]->b21 + n);
  }
  consumer_sync();
  const float *g21[NETS], *b21[NETS], *g22[NETS], *b22[NETS], *w22[NETS], *bb22[NETS];
  for (int i = 0; i < nn; ++i) {
    g21[i] = c[i]->g21; b21[i] = c[i]->be21; g22[i] = c[i]->g22; b22[i] = c[i]->be22; w22[i] = c[i]->w22; bb22[i] = c[i]->b22;
  }
  ln_elu_multi(nn, t.v21, t.zh21, HA1, g21, b21, t.rsn, 2);
  consumer_sync();
  gemv_multi<HA1, GEMV_U * 2>(nn, w22, HA1, bb22, HA2, t.v21, t.v22);
  consumer_sync();
  ln_elu_multi(nn, t.v22, t.zh22, HA2, g22, b22, t.rsn + 1, 2);
  consumer_sync();
}
// q[net][t] = b3 + w3[0:H2].h12r[net][t] + w3[H2:].h22[net][t]
__device__ void critic_q_multi(int nn, const CriticP* const* c, const TailSm& t) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rix = warp; rix < nn * RT; rix += SC_WARPS) {
    const CriticP* cc = c[rix / RT];
    float s = 0.f;
    for (int k = lane; k < H2; k += 32) s = fmaf(ld_w(cc->w3 + k), t.h12r[rix * H2 + k], s);
    for (int k = lane; k < HA2; k += 32) s = fmaf(ld_w(cc->w3 + H2 + k), t.v22[rix * HA2 + k], s);
    s = warp_sum(s);
    if (lane == 0) t.qv[rix] = s + ld_w(cc->b3);
  }
}

// actor layer 3 + heads on the rows already staged in t.xh2.  Leaves h3/zh3 in t.x3/t.zh3, raw head outputs in t.head,
// 1/std of layer 3 in t.rsx[t*8+2].  Ends synchronised (3 barriers).
__device__ void actor_tail_forward(const RowPhaseParams& p, const TailSm& t) {
  const ActorP& a = p.actor;
  const float *w3[1] = {a.w3}, *b3[1] = {a.b3}, *g3[1] = {a.g3}, *be3[1] = {a.be3};
  gemv_multi<H2, GEMV_U>(1, w3, H2, b3, H3, t.xh2, t.x3);
  consumer_sync();
  ln_elu_multi(1, t.x3, t.zh3, H3, g3, be3, t.rsx + 2, 8);
  consumer_sync();
  // heads: (mu | logsigma) rows of width A each, output to t.head[t][0:A] and [AMAX:AMAX+A]
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nh = a.wls != nullptr ? 2 * p.A : p.A;
    for (int o = warp; o < nh; o += SC_WARPS) {
      const bool ls = o >= p.A;
      const int j = ls ? o - p.A : o;
      const float* w = (ls ? a.wls : a.wmu) + (size_t)j * H3;
      const float4 wv = ld_w4(w + 4 * lane);           // H3 == 128 == 32 lanes x float4
      float acc[RT];
#pragma unroll
      for (int tt = 0; tt < RT; ++tt) acc[tt] = warp_sum(dot4(wv, *reinterpret_cast<const float4*>(t.x3 + tt * H3 + 4 * lane), 0.f));
      if (lane == 0) {
        const float b = ld_w((ls ? a.bls : a.bmu) + j);
#pragma unroll
        for (int tt = 0; tt < RT; ++tt) t.head[tt * 2 * AMAX + (ls ? AMAX : 0) + j] = acc[tt] + b;
      }
    }
  }
  consumer_sync();
}

__device__ __forceinline__ float eps_value(const RowPhaseParams& p, int row, int j) {
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

// tanh-squashed Gaussian sample + log-prob (SAC, ref ActorNetwork.sample_normal enet_sac.py:446-466) or the
// deterministic TD3 head.  eps_of(row, j) supplies N(0,1).  Results: t.an[0] (action), t.sc[SC_LOGP], t.smp.  2 barriers.
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
      t.lp[r * AMAX + j] = lp;
    } else {
      float act = tanhf(t.head[r * 2 * AMAX + j]);
      smp[0] = act;
      if (td3_target) act = fminf(fmaxf(act + p.td3_noise_c, -1.f), 1.f);   // enet_td3.py:248-251
      t.an[r * AMAX + j] = act;
    }
  }
  consumer_sync();
  if (p.kind == 0 && tid < RT) {
    float s = 0.f;
    for (int j = 0; j < p.A; ++j) s += t.lp[tid * AMAX + j];
    t.sc[tid * 16 + SC_LOGP] = s;
  }
  consumer_sync();
}

// ======================================================================================================
// PROG_CRITIC tail: policy on s', both target critics, TD target, both online critics forward + backward
// ======================================================================================================
__device__ void tail_critic(const RowPhaseParams& p, const TailSm& t, const RowSet& rs, long long* prof = nullptr) {
  const int tid = threadIdx.x;
  long long tprev = (prof && tid == 0) ? clock64() : 0;
  auto mark = [&](int i) { if (prof && tid == 0) { const long long c = clock64(); prof[i] += c - tprev; tprev = c; } };
  // ---- stage every global row this pass needs (one round of L2 latency) ----
  load_rows_to(t.xh2, p.aw.t_h2, H2, rs);
  for (int k = 0; k < 2; ++k) {
    load_rows_to(t.h12r + k * RT * H2, p.tw[k].h12, H2, rs);
    load_rows_to(t.h12r + (2 + k) * RT * H2, p.cw[k].h12, H2, rs);
    load_rows_to(t.zh12r + k * RT * H2, p.cw[k].zh12, H2, rs);
  }
  if (tid < RT * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    t.an[RT * AMAX + r * AMAX + j] = rs.row[r] >= 0 ? ld_w(p.rp.action + (size_t)p.bidx[rs.row[r]] * p.Ap + j) : 0.f;
  }
  if (tid < 2 * RT) {
    const int k = tid / RT, r = tid - k * RT;
    t.rsx[r * 8 + 4 + k] = rs.row[r] >= 0 ? ld_act(p.cw[k].rs12 + rs.row[r]) : 0.f;
  }
  consumer_sync();
  mark(8);
  // ---- policy on s' (SAC: online actor, sampled; TD3: target actor + clipped scalar noise) ----
  actor_tail_forward(p, t);
  mark(9);
  policy_head(p, t, rs, true, [&](int row, int j) { return eps_value(p, row, j); });
  mark(10);
  // ---- action branches of all four critics at once: targets on a', online critics on the replay action ----
  const CriticP* nets[NETS] = {&p.targ[0], &p.targ[1], &p.crit[0], &p.crit[1]};
  const float* an_sel[NETS] = {t.an, t.an, t.an + RT * AMAX, t.an + RT * AMAX};
  action_branches(NETS, nets, an_sel, p.A, p.Ap, t);
  mark(11);
  critic_q_multi(NETS, nets, t);
  consumer_sync();
  mark(12);
  const float wsum = p.is_wsum ? ld_act(p.is_wsum) : (float)p.B;
  const float coef = 2.f * wsum / ((float)p.B * (float)p.B) * p.closs_scale;   // d/dq of mean(w)*mean((q-y)^2) [DDPG: sum]
  if (tid < RT) {
    float* s = t.sc + tid * 16;
    const int row = rs.row[tid];
    float dq1 = 0.f, dq2 = 0.f;
    if (row >= 0) {
      const int slot = p.bidx[row];
      const bool done = p.rp.done[slot] != 0;
      const float qt1 = t.qv[0 * RT + tid], qt2 = t.qv[1 * RT + tid], q1 = t.qv[2 * RT + tid], q2 = t.qv[3 * RT + tid];
      float yv;
      if (p.kind == 0) {   // enet_sac.py:573-575
        float m = __fsub_rn(fminf(qt1, qt2), __fmul_rn(p.alpha, s[SC_LOGP]));
        if (done) m = 0.f;
        yv = __fadd_rn(__fmul_rn(p.scale, ld_w(p.rp.reward + slot)), __fmul_rn(p.gamma, m));
      } else {             // enet_td3.py:255-260
        const float m = done ? 0.f : fminf(qt1, qt2);
        yv = __fadd_rn(ld_w(p.rp.reward + slot), __fmul_rn(p.gamma, m));
      }
      p.y[row] = yv;
      dq1 = coef * (q1 - yv);
      dq2 = coef * (q2 - yv);
      p.cw[0].q[row] = q1; p.cw[1].q[row] = q2;
      p.cw[0].dq[row] = dq1; p.cw[1].dq[row] = dq2;
      const float er = 0.5f * (fabsf(q1 - yv) + fabsf(q2 - yv));      // enet_td3.py:266-269
      p.err[row] = er;
      // PER.batch_update priority (enet_sac.py:318-320): float32 (err + eps) -> min(., clip) -> power(., float32 0.6),
      // evaluated in double and rounded once; done here so that the fp64 pow runs on ~120 SMs instead of one
      if (p.per_ps != nullptr) p.per_ps[row] = (float)pow((double)fminf(__fadd_rn(er, 0.01f), p.per_clip), (double)0.6f);
      for (int k = 0; k < 2; ++k) {
        p.cw[k].rs21[row] = t.rsn[((2 + k) * RT + tid) * 2 + 0];
        p.cw[k].rs22[row] = t.rsn[((2 + k) * RT + tid) * 2 + 1];
      }
    }
    s[SC_DQ1] = dq1;
    s[SC_DQ2] = dq2;
  }
  for (int k = 0; k < 2; ++k) {      // forward caches of the online critics (inputs of the weight-gradient kernel)
    store_rows_from(p.cw[k].h21, t.v21 + (2 + k) * RT * HA1, HA1, rs);
    store_rows_from(p.cw[k].zh21, t.zh21 + (2 + k) * RT * HA1, HA1, rs);
    store_rows_from(p.cw[k].h22, t.v22 + (2 + k) * RT * HA2, HA2, rs);
    store_rows_from(p.cw[k].zh22, t.zh22 + (2 + k) * RT * HA2, HA2, rs);
  }
  consumer_sync();
  mark(13);
  // ---- backward of both online critics: d cat = dq * w3 ----
  for (int e = tid; e < 2 * RT * HCAT; e += SC_THREADS) {
    const int rix = e / HCAT, j = e - rix * HCAT, k = rix / RT, r = rix - k * RT;
    const float gval = t.sc[r * 16 + SC_DQ1 + k] * ld_w(p.crit[k].w3 + j);
    if (j < H2) t.ga[rix * H2 + j] = gval; else t.ge[rix * HA2 + (j - H2)] = gval;
  }
  consumer_sync();
  {
    const float *g12[2] = {p.crit[0].g12, p.crit[1].g12}, *g22[2] = {p.crit[0].g22, p.crit[1].g22};
    // rs12 of (k, r) sits at t.rsx[r*8 + 4 + k]: present it as a [2*RT] array with stride... copy to a dense temp
    if (tid < 2 * RT) t.qv[tid] = t.rsx[(tid % RT) * 8 + 4 + tid / RT];
    consumer_sync();
    ln_bwd_multi(2, t.ga, t.h12r + 2 * RT * H2, t.zh12r, H2, g12, t.qv, 1, t.gb);
    ln_bwd_multi(2, t.ge, t.v22 + 2 * RT * HA2, t.zh22 + 2 * RT * HA2, HA2, g22, t.rsn + 2 * RT * 2 + 1, 2, t.gf);
  }
  consumer_sync();
  mark(14);
  for (int k = 0; k < 2; ++k) {
    store_rows_from(p.cw[k].du12, t.ga + k * RT * H2, H2, rs);
    store_rows_from(p.cw[k].dz12, t.gb + k * RT * H2, H2, rs);
    store_rows_from(p.cw[k].du22, t.ge + k * RT * HA2, HA2, rs);
    store_rows_from(p.cw[k].dz22, t.gf + k * RT * HA2, HA2, rs);
  }
  {
    const float* w22[2] = {p.crit[0].w22, p.crit[1].w22};
    gemvT_multi(2, w22, HA1, HA2, HA1, t.gf, t.gc, t.scratch);
    const float* g21[2] = {p.crit[0].g21, p.crit[1].g21};
    ln_bwd_multi(2, t.gc, t.v21 + 2 * RT * HA1, t.zh21 + 2 * RT * HA1, HA1, g21, t.rsn + 2 * RT * 2, 2, t.gd);
  }
  consumer_sync();
  for (int k = 0; k < 2; ++k) {
    store_rows_from(p.cw[k].du21, t.gc + k * RT * HA1, HA1, rs);
    store_rows_from(p.cw[k].dz21, t.gd + k * RT * HA1, HA1, rs);
  }
  consumer_sync();
  mark(15);
}

// ======================================================================================================
// PROG_ACTOR tail: policy on s, Q_k(s, pi(s)) with the updated critics, backward to d z2 of the actor
// ======================================================================================================
__device__ void tail_actor(const RowPhaseParams& p, const TailSm& t, const RowSet& rs) {
  const int tid = threadIdx.x;
  const ActorP& a = p.actor;
  const int nq = p.n_q;
  load_rows_to(t.xh2, p.aw.h2, H2, rs);
  load_rows_to(t.zh2r, p.aw.zh2, H2, rs);
  for (int k = 0; k < nq; ++k) load_rows_to(t.h12r + k * RT * H2, p.cw[k].p_h12, H2, rs);
  if (tid < RT) t.rsx[tid * 8 + 3] = rs.row[tid] >= 0 ? ld_act(p.aw.rs2 + rs.row[tid]) : 0.f;
  consumer_sync();
  actor_tail_forward(p, t);
  store_rows_from(p.aw.h3, t.x3, H3, rs);
  store_rows_from(p.aw.zh3, t.zh3, H3, rs);
  if (tid < RT && rs.row[tid] >= 0) p.aw.rs3[rs.row[tid]] = t.rsx[tid * 8 + 2];
  policy_head(p, t, rs, false, [&](int row, int j) { return eps_value(p, row, j); });
  if (tid < RT * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    if (rs.row[r] >= 0) p.aw.pi[(size_t)rs.row[r] * p.A + j] = t.an[r * AMAX + j];
  }
  if (tid < RT && rs.row[tid] >= 0 && p.kind == 0) p.aw.logp[rs.row[tid]] = t.sc[tid * 16 + SC_LOGP];
  // ---- Q_k(s, pi(s)) with the freshly updated critics ----
  const CriticP* nets[NETS] = {&p.crit[0], &p.crit[1], nullptr, nullptr};
  const float* an_sel[NETS] = {t.an, t.an, nullptr, nullptr};
  action_branches(nq, nets, an_sel, p.A, p.Ap, t);
  critic_q_multi(nq, nets, t);
  consumer_sync();
  const float wsum = p.is_wsum ? ld_act(p.is_wsum) : (float)p.B;
  const float coef = wsum / ((float)p.B * (float)p.B);           // d/dx of mean(x*w) with the (B,B) broadcast
  if (tid < RT) {
    float* s = t.sc + tid * 16;
    float g1 = 0.f, g2 = 0.f;
    if (rs.row[tid] >= 0) {
      const float q1 = t.qv[tid], q2 = nq == 2 ? t.qv[RT + tid] : 0.f;
      if (nq == 2) {   // torch.min backward: gradient to the smaller, halves on ties
        if (q1 < q2) g1 = -coef; else if (q2 < q1) g2 = -coef; else { g1 = -0.5f * coef; g2 = -0.5f * coef; }
        p.aw.qmin[rs.row[tid]] = fminf(q1, q2);       // critic_value (enet_sac.py:592)
      } else {
        g1 = -coef;
        p.aw.qmin[rs.row[tid]] = q1;
      }
    }
    s[SC_DQ1] = g1;
    s[SC_DQ2] = g2;
    s[SC_GLP] = rs.row[tid] >= 0 ? p.alpha * coef : 0.f;
  }
  consumer_sync();
  // ---- back through the action branches of the critics to d loss / d action ----
  for (int e = tid; e < nq * RT * HA2; e += SC_THREADS) {
    const int rix = e / HA2, j = e - rix * HA2, k = rix / RT, r = rix - k * RT;
    t.ge[e] = t.sc[r * 16 + SC_DQ1 + k] * ld_w(p.crit[k].w3 + H2 + j);
  }
  consumer_sync();
  {
    const float *g22[2] = {p.crit[0].g22, p.crit[1].g22}, *g21[2] = {p.crit[0].g21, p.crit[1].g21};
    const float* w22[2] = {p.crit[0].w22, p.crit[1].w22};
    ln_bwd_multi(nq, t.ge, t.v22, t.zh22, HA2, g22, t.rsn + 1, 2, t.gf);
    consumer_sync();
    gemvT_multi(nq, w22, HA1, HA2, HA1, t.gf, t.gc, t.scratch);
    ln_bwd_multi(nq, t.gc, t.v21, t.zh21, HA1, g21, t.rsn, 2, t.gd);
    consumer_sync();
  }
  {   // d action[r][j] = sum_k sum_n dz21_k[r][n] * w21_k[n][j]   (one warp per (row, j), critics in fixed order)
    const int warp = tid >> 5, lane = tid & 31;
    for (int o = warp; o < RT * p.A; o += SC_WARPS) {
      const int r = o / p.A, j = o - r * p.A;
      float tot = 0.f;
      for (int k = 0; k < nq; ++k) {
        float s = 0.f;
        for (int n = lane; n < HA1; n += 32) s = fmaf(t.gd[(k * RT + r) * HA1 + n], ld_w(p.crit[k].w21 + n * p.Ap + j), s);
        tot += warp_sum(s);
      }
      if (lane == 0) t.lp[r * AMAX + j] = tot;
    }
  }
  consumer_sync();
  // ---- back through the policy head ----
  if (tid < RT * p.A) {
    const int r = tid / p.A, j = tid - r * p.A;
    float da = t.lp[r * AMAX + j];
    const float* smp = t.smp + (r * AMAX + j) * 4;
    const int row = rs.row[r];
    if (p.hint_coef != nullptr && row >= 0)
      da = fmaf(ld_act(p.hint_coef), t.an[r * AMAX + j] - ld_w(p.rp.hint + (size_t)p.bidx[row] * p.Ap + j), da);
    if (p.admm_lambda != nullptr && row >= 0) {
      const float numel = (float)p.B * (float)p.A, wmean = wsum / (float)p.B;
      const float diff = t.an[r * AMAX + j] - ld_w(p.rp.hint + (size_t)p.bidx[row] * p.Ap + j);
      da += wmean / numel * ld_act(p.admm_lambda + (size_t)row * p.A + j) + wmean * ld_act(p.admm_rho) / (numel * numel) * diff;
    }
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
  consumer_sync();
  for (int e = tid; e < RT * H3; e += SC_THREADS) {
    const int r = e / H3, k = e - r * H3;
    float s = 0.f;
    for (int j = 0; j < p.A; ++j) {
      s = fmaf(t.head[r * 2 * AMAX + j], ld_w(a.wmu + j * H3 + k), s);
      if (a.wls != nullptr) s = fmaf(t.head[r * 2 * AMAX + AMAX + j], ld_w(a.wls + j * H3 + k), s);
    }
    t.ga[e] = s;
  }
  consumer_sync();
  {
    const float *g3[1] = {a.g3}, *g2[1] = {a.g2}, *w3[1] = {a.w3};
    ln_bwd_multi(1, t.ga, t.x3, t.zh3, H3, g3, t.rsx + 2, 8, t.gb);
    consumer_sync();
    store_rows_from(p.aw.du3, t.ga, H3, rs);
    store_rows_from(p.aw.dz3, t.gb, H3, rs);
    gemvT_multi(1, w3, H2, H3, H2, t.gb, t.gc, t.scratch);     // t.gc has room for [RT][H2] (2*RT*HA1 floats)
    ln_bwd_multi(1, t.gc, t.xh2, t.zh2r, H2, g2, t.rsx + 3, 8, t.gd);
    consumer_sync();
    store_rows_from(p.aw.du2, t.gc, H2, rs);
    store_rows_from(p.aw.dz2, t.gd, H2, rs);
  }
  consumer_sync();
}

// gather the cluster's minibatch rows of `src` ([C][Sp] replay array) into the dense TMA-readable buffer dst [B][Sp32]
__device__ __forceinline__ void gather_rows(float* dst, const float* __restrict__ src, const int* __restrict__ bidx, int row0,
                                            int nrows, int S, int Sp, int Sp32, int rank, int cs) {
  const int n4 = Sp32 >> 2, v4 = Sp >> 2;
  for (int lr = rank; lr < nrows; lr += cs) {
    const int row = row0 + lr;
    const float* sp = src + (size_t)__ldg(bidx + row) * Sp;
    for (int c = threadIdx.x; c < n4; c += SC_ROW_THREADS) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < v4) v = ld_w4(sp + 4 * c);
      *reinterpret_cast<float4*>(dst + (size_t)row * Sp32 + 4 * c) = v;
    }
  }
}

template <int CS, int TM>
__global__ void __launch_bounds__(SC_ROW_THREADS, SC_CFG_MINBLOCKS) k_rowphase(const RowPhaseParams* __restrict__ pp) {
  __shared__ RowPhaseParams p_sm;
  const RowPhaseParams& p = load_params(pp, &p_sm);           // blockIdx.y = agent of the ensemble
  extern __shared__ __align__(16) float smem_raw[];
  // the TMA boxes use the 128-byte swizzle, whose pattern repeats every 1024 bytes: align the ring accordingly
  float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ DenseOp dops[MAXOPS];
  __shared__ DgradOp gops[2];
  __shared__ __align__(8) uint64_t mbars[2 * NSTAGE];
  const int tid = threadIdx.x;
  const bool consumer = tid < SC_THREADS;
  __shared__ long long prof_acc[16];
  if (tid < 16) prof_acc[tid] = 0;
  unsigned g = 0;                 // chunks streamed so far (identical sequence in producer and consumers)
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(mbars + i, 1);
      mbar_init(mbars + NSTAGE + i, SC_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  pdl_launch();
  pdl_wait();                     // (programmatic dependent launch: the barrier setup above overlapped the previous kernel)
  const Pipe pipe{p.dbg, p.tstamp ? prof_acc : nullptr, mbars, mbars + NSTAGE,
                  (p.tstamp && p.program == PROG_CRITIC) ? p.tstamp + 2048 * 8 : nullptr};
  if (pipe.tl != nullptr && tid == 0 && blockIdx.x == 0) pipe.tl[0] = 0;
  const int rank = (int)cluster_ctarank();
  const int ncl = (int)num_clusters();
  const int cid = (int)cluster_idx();
  const int per = (p.B + ncl - 1) / ncl;
  const int row0 = cid * per;
  const int nrows = max(0, min(per, p.B - row0));
  // (clusters with nrows == 0 still walk the barriers; all their loops are empty)

  auto dense = [&](int i, const void* mapA, const void* mapW, const float* b, const float* g_, const float* be, int K, int N,
                   float* H, float* ZH, float* RS) {
    if (tid == 0) dops[i] = DenseOp{mapA, mapW, b, g_, be, K, N, H, ZH, RS, N};
  };
  const TailSm ts = carve_tail(smem);
  auto stamp = [&](int i) {
    if (p.tstamp != nullptr && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.tstamp[(size_t)blockIdx.x * 8 + i] = t;
    }
  };
  const TensorMaps& tm = p.tm;
  const int K1 = p.Sp32;          // layer-1 reduction length, zero-padded to whole 32-float k-blocks

  if (p.program == PROG_CRITIC) {
    const ActorP& a = p.actor;
    stamp(0);
    if (!p.skip_dense) {
    // stage 0: fused minibatch gather (replay rows -> dense buffers the TMA unit can tile), one cluster barrier
    gather_rows(p.mb_s, p.rp.state, p.bidx, row0, nrows, p.S, p.Sp, p.Sp32, rank, CS);
    gather_rows(p.mb_s2, p.rp.new_state, p.bidx, row0, nrows, p.S, p.Sp, p.Sp32, rank, CS);
    // stage 1: first layers.  s' feeds the policy and both target critics, s feeds both online critics.
    dense(0, tm.mb_s2, tm.actor_w1, a.b1, a.g1, a.be1, K1, H1, p.aw.t_h1, nullptr, nullptr);
    dense(1, tm.mb_s2, tm.targ_w11[0], p.targ[0].b11, p.targ[0].g11, p.targ[0].be11, K1, H1, p.tw[0].h11, nullptr, nullptr);
    dense(2, tm.mb_s2, tm.targ_w11[1], p.targ[1].b11, p.targ[1].g11, p.targ[1].be11, K1, H1, p.tw[1].h11, nullptr, nullptr);
    dense(3, tm.mb_s, tm.crit_w11[0], p.crit[0].b11, p.crit[0].g11, p.crit[0].be11, K1, H1, p.cw[0].h11, p.cw[0].zh11, p.cw[0].rs11);
    dense(4, tm.mb_s, tm.crit_w11[1], p.crit[1].b11, p.crit[1].g11, p.crit[1].be11, K1, H1, p.cw[1].h11, p.cw[1].zh11, p.cw[1].rs11);
    cluster_sync();
    stamp(6);
    dense_stage<CS, TM>(dops, 5, p.B, row0, nrows, p.bidx, p.part, smem, pipe, g);
    stamp(1);
    // stage 2: second layers
    dense(0, tm.t_h1, tm.actor_w2, a.b2, a.g2, a.be2, H1, H2, p.aw.t_h2, nullptr, nullptr);
    dense(1, tm.th11[0], tm.targ_w12[0], p.targ[0].b12, p.targ[0].g12, p.targ[0].be12, H1, H2, p.tw[0].h12, nullptr, nullptr);
    dense(2, tm.th11[1], tm.targ_w12[1], p.targ[1].b12, p.targ[1].g12, p.targ[1].be12, H1, H2, p.tw[1].h12, nullptr, nullptr);
    dense(3, tm.ch11[0], tm.crit_w12[0], p.crit[0].b12, p.crit[0].g12, p.crit[0].be12, H1, H2, p.cw[0].h12, p.cw[0].zh12, p.cw[0].rs12);
    dense(4, tm.ch11[1], tm.crit_w12[1], p.crit[1].b12, p.crit[1].g12, p.crit[1].be12, H1, H2, p.cw[1].h12, p.cw[1].zh12, p.cw[1].rs12);
    __syncthreads();
    dense_stage<CS, TM>(dops, 5, p.B, row0, nrows, p.bidx, p.part, smem, pipe, g,
                        p.tstamp ? p.tstamp + 512 * 8 + (size_t)blockIdx.x * 8 : nullptr);
    }
    stamp(2);
    // stage 3: row-local tail
    if (consumer && !p.skip_tail)
      for (int gr = rank; gr < nrows; gr += CS * RT) {
        RowSet rs;
#pragma unroll
        for (int t = 0; t < RT; ++t) rs.row[t] = (gr + t * CS < nrows) ? row0 + gr + t * CS : -1;
        tail_critic(p, ts, rs, p.tstamp ? prof_acc : nullptr);
      }
    stamp(3);
    if (tid == 0) {
      for (int k = 0; k < 2; ++k)
        gops[k] = DgradOp{tm.dz12[k], tm.crit_w12p[k], H2, H1, p.cw[k].h11, p.cw[k].zh11, p.cw[k].rs11,
                          p.crit[k].g11, p.cw[k].du11, p.cw[k].dz11, H1};
    }
    cluster_sync();
    stamp(4);
    // stage 4: dz12 -> layer-11 gradients
    dgrad_stage<CS, TM>(gops, 2, p.B, row0, nrows, p.part, smem, pipe, g);
    stamp(5);
    if (p.tstamp != nullptr && tid == 0)
      for (int i = 0; i < 16; ++i) p.tstamp[(size_t)(1024 + (i >> 3) * 1536 + blockIdx.x) * 8 + (i & 7)] = (unsigned long long)prof_acc[i];
  } else if (p.program == PROG_ACTOR) {
    const ActorP& a = p.actor;
    // mb_s was gathered by the critic phase of this step (same minibatch)
    auto astamp = [&](int i) {
      if (p.tstamp != nullptr && tid == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.tstamp[(size_t)(1536 + blockIdx.x) * 8 + i] = t;
      }
    };
    astamp(0);
    if (!p.skip_dense) {
    dense(0, tm.mb_s, tm.actor_w1, a.b1, a.g1, a.be1, K1, H1, p.aw.h1, p.aw.zh1, p.aw.rs1);
    for (int k = 0; k < p.n_q; ++k)
      dense(1 + k, tm.mb_s, tm.crit_w11[k], p.crit[k].b11, p.crit[k].g11, p.crit[k].be11, K1, H1, p.cw[k].p_h11, nullptr, nullptr);
    __syncthreads();
    dense_stage<CS, TM>(dops, 1 + p.n_q, p.B, row0, nrows, p.bidx, p.part, smem, pipe, g);
    astamp(1);
    dense(0, tm.h1, tm.actor_w2, a.b2, a.g2, a.be2, H1, H2, p.aw.h2, p.aw.zh2, p.aw.rs2);
    for (int k = 0; k < p.n_q; ++k)
      dense(1 + k, tm.ph11[k], tm.crit_w12[k], p.crit[k].b12, p.crit[k].g12, p.crit[k].be12, H1, H2, p.cw[k].p_h12, nullptr, nullptr);
    __syncthreads();
    dense_stage<CS, TM>(dops, 1 + p.n_q, p.B, row0, nrows, p.bidx, p.part, smem, pipe, g);
    }
    astamp(2);
    if (consumer && !p.skip_tail)
      for (int gr = rank; gr < nrows; gr += CS * RT) {
        RowSet rs;
#pragma unroll
        for (int t = 0; t < RT; ++t) rs.row[t] = (gr + t * CS < nrows) ? row0 + gr + t * CS : -1;
        tail_actor(p, ts, rs);
      }
    astamp(3);
    if (tid == 0)
      gops[0] = DgradOp{tm.dz2, tm.actor_w2p, H2, H1, p.aw.h1, p.aw.zh1, p.aw.rs1, a.g1, p.aw.du1, p.aw.dz1, H1};
    cluster_sync();
    astamp(4);
    dgrad_stage<CS, TM>(gops, 1, p.B, row0, nrows, p.part, smem, pipe, g);
    astamp(5);
  } else {   // PROG_ACT: one observation through the actor (enet_sac.py:547-553)
    const ActorP& a = p.actor;
    dense(0, tm.obs, tm.actor_w1, a.b1, a.g1, a.be1, K1, H1, p.act_aux, nullptr, nullptr);
    __syncthreads();
    dense_stage<CS, TM>(dops, 1, 1, row0, nrows, nullptr, p.part, smem, pipe, g);
    dense(0, tm.aux, tm.actor_w2, a.b2, a.g2, a.be2, H1, H2, p.act_aux + H1, nullptr, nullptr);
    __syncthreads();
    dense_stage<CS, TM>(dops, 1, 1, row0, nrows, nullptr, p.part, smem, pipe, g);
    if (rank == 0 && nrows > 0 && consumer) {
      RowSet rs;
      rs.row[0] = 0;
#pragma unroll
      for (int t = 1; t < RT; ++t) rs.row[t] = -1;
      load_rows_to(ts.xh2, p.act_aux + H1, H2, rs);
      consumer_sync();
      actor_tail_forward(p, ts);
      policy_head(p, ts, rs, false, [&](int, int j) { return p.eps ? ld_act(p.eps + j) : 0.f; });
      if (tid < p.A) {
        float v = ts.an[tid];
        if (p.kind != 0 && p.eps != nullptr) v = fminf(fmaxf(v + ld_act(p.eps + tid), -1.f), 1.f);   // enet_td3.py:215-217
        p.act_out[tid] = v;
      }
    }
  }
}


constexpr size_t SMEM_BYTES = (size_t)SM_TOTAL_FLOATS * sizeof(float) + 1024;   // + slack for the 1024-byte alignment

}  // namespace SC_CFG_NS
}  // namespace sc
