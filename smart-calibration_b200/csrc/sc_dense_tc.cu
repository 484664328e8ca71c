// The two wide forward layers of every network pass (fc1/fc11: S -> 512, fc2/fc12: 512 -> 256; ref enet_sac.py:357-368,
// 415-424) on the 5th-generation tensor cores: Z = X W^T as a tcgen05.mma kind::tf32 GEMM with the batch rows on M (128 per
// CTA, one accumulator lane per row) and 64 output features on N, accumulator in tensor memory, followed by the fused
// bias + LayerNorm + ELU epilogue.  A row's LayerNorm statistics span the N / 64 CTAs of a thread-block cluster, which
// exchange per-row (mean, M2) partials through distributed shared memory and combine them in fixed order (Chan et al.),
// exactly like the mma.sync row-phase stage this replaces for batched passes (sc_rowphase_impl.cuh dense_stage).
//
// Both operands are K-major (X [rows][K], W [features][K]): TMA boxes of [128 rows][32 floats] with the 128-byte swizzle,
// described to the MMA as 8-row atoms of 1024 B (SBO), +32 B per 8-deep k-step.  kind::tf32 truncates its fp32 inputs to 10
// mantissa bits, so fp32 fidelity is restored with three products per k-step, X_lo W + X W_lo + X W.  Both operands arrive
// already split into tf32 big / small parts: whoever produces a tensor-core operand (k_gather_mb, the epilogues below, k_tail,
// the Adam epilogue of k_wgrad_tc, k_split) writes the two parts once, so the K loop here is TMA -> MMA with nothing in between.
//
// Warp roles (512 threads): warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue (one thread); epilogue on all 16
// warps: warp w reads TMEM lanes 32 (w & 3).. (its rows) x columns 16 (w >> 2).. (its features).
// A CTA's K loop is a chain of TMA round trips, which is why the tile is only 64 wide: more, smaller CTAs run more of those
// chains side by side.  For the same reason the reduction can be split over a CTA pair
// (DenseTcParams.ks = 2: cluster = feature blocks x 2, the second CTA's accumulators travel through distributed shared memory).
// Launched with programmatic dependent launch: everything up to pdl_wait() overlaps the previous kernel.
#include "sc_common.cuh"
#include "sc_params.h"

namespace sc {

constexpr int DT_M = 128, DT_N = 64;                         // batch rows / output features per CTA
constexpr int DT_STAGES = 3;
constexpr int DT_CW = DT_N / 4;                              // epilogue: features per thread
constexpr int DT_X_BYTES = DT_M * 128, DT_W_BYTES = DT_N * 128;   // one k-chunk of 32 floats
constexpr int DT_STAGE_BYTES = 2 * (DT_X_BYTES + DT_W_BYTES);     // big + small parts: 48 KB
constexpr int DT_THREADS = 512;
constexpr int DT_MAXCS = 8;
constexpr int DT_TP = DT_N + 4;                              // output tile pitch (floats)
static_assert(2 * DT_M * DT_TP * 4 <= DT_STAGE_BYTES * DT_STAGES, "output tiles must fit in the operand ring");
constexpr int DT_PART_BYTES = DT_M * DT_TP * 4;             // split-K: the partner CTA's partial accumulators land here

__device__ __forceinline__ uint64_t dt_desc(uint32_t saddr) {
  uint64_t d = 0;                                            // K-major, SWIZZLE_128B, 8-row atoms of 1024 B
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t dt_desc_mn(uint32_t saddr) {
  // MN-major, SWIZZLE_128B_BASE32B (boxes of [32 reduction rows][32 floats] written with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B):
  // LBO = 4096 B (next box of 32 MN elements), SBO = 512 B, version 1 -- the layout sc_wgrad_tc.cu established with tools/probe/tc_probe.cu
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((4096u >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ void dt_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void dt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MODE 0: forward layer  H = ELU(LN(X W^T + b))            (X [B][K] K-major, W [N][K] K-major)
// MODE 1: input gradient dU = (dZ W) * ELU'(h), dZ_in = LayerNorm backward of the producing layer
//         (dZ [B][K] K-major with K = width of the layer above, W [K][N] MN-major panels, N = width of the producing layer).
//         Replaces the mma.sync dgrad stage of k_rowphase (sc_rowphase_impl.cuh dgrad_stage) for batched passes.
template <int MODE>
__global__ void __launch_bounds__(DT_THREADS, 1) k_dense_tc_t(const DenseTcParams* __restrict__ pp) {
  __shared__ DenseTcParams p_sm;
  const DenseTcParams& p = load_params(pp, &p_sm);             // blockIdx.y = agent of the ensemble
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[DT_STAGES], empty[DT_STAGES], acc_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float psum[4][DT_M], pm2[4][DT_M];
  __shared__ __align__(16) float svec[3][DT_N];             // bias, gamma, beta of this CTA's features
  __shared__ __align__(8) float2 xstat[DT_MAXCS][DT_M];     // (mean, M2) of every cluster CTA's 128-feature slice, per row
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  const int tile = (int)cluster_idx();
  const int op_i = tile / p.nrb, rb = tile - op_i * p.nrb;
  const DenseTcOp& op = p.ops[op_i];
  // cluster = (N / 64 feature blocks) x (ks reduction halves): rank = kh * nfb + fb.  With ks = 2 the kh = 1 CTA reduces the
  // second half of K and hands its raw accumulators to CTA fb (kh = 0) through distributed shared memory.
  const int nfb = cs / p.ks, fb = rank % nfb, kh = rank / nfb;
  const int n0 = fb * DT_N, r0 = rb * DT_M;
  float* part = reinterpret_cast<float*>(smem + (size_t)DT_STAGES * DT_STAGE_BYTES);
  // debug timeline of CTA 0: each of three threads appends (globaltimer << 8 | tag) to its own shared-memory list (flushed at
  // the end, so that taking a mark costs no memory round trip)
  __shared__ unsigned long long tlbuf[3][62];
  __shared__ int tln[3];
  auto mark = [&](int slot, int tag) {
    if (p.tl != nullptr && blockIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      const int n = tln[slot];
      if (n < 62) { tlbuf[slot][n] = (t << 8) | (unsigned long long)tag; tln[slot] = n + 1; }
    }
  };
  auto stamp = [&](int i) {
    if (p.ts != nullptr && tid == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.ts[(size_t)blockIdx.x * 8 + i] = t;
    }
  };
  stamp(0);
  if (tid < 3) tln[tid] = 0;
  if (tid == 0) {
    for (int i = 0; i < DT_STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    mbar_init(&acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(DT_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // first cluster barrier phase: its wait (before the statistics exchange) proves every peer CTA is resident
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  pdl_wait();                                                 // everything above overlapped the previous kernel's tail
  stamp(1);
  const uint32_t tmem = tmem_slot;
  const int nk_all = (op.K + 31) >> 5;
  const int nk = nk_all / p.ks, kc0 = kh * nk;                // (launch_dense_tc checks divisibility)

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    for (int s = 0; s < nk; ++s) {
      const int slot = s % DT_STAGES;
      if (s >= DT_STAGES) mbar_wait(empty + slot, ((s / DT_STAGES) - 1) & 1u);
      uint8_t* st = smem + (size_t)slot * DT_STAGE_BYTES;
      if (lane == 0) {
        mark(0, 1);
        mbar_expect_tx(full + slot, DT_STAGE_BYTES);
      }
      __syncwarp();
      // stage layout: [X big | W big | X small | W small]; lanes 0 / 1 fetch the big / small parts (the small part's map follows the big part's)
      if (lane < 2) {
        const void* mx = reinterpret_cast<const uint8_t*>(op.mapX) + 128 * lane;
        const void* mw = reinterpret_cast<const uint8_t*>(op.mapW) + 128 * lane;
        uint8_t* sp = st + lane * (DT_X_BYTES + DT_W_BYTES);
        tma_load_2d(reinterpret_cast<float*>(sp), mx, 32 * (kc0 + s), r0, full + slot);
        if constexpr (MODE == 0) {
          tma_load_2d(reinterpret_cast<float*>(sp + DT_X_BYTES), mw, 32 * (kc0 + s), n0, full + slot);
        } else {          // two panels [32 reduction rows][32 output columns]
          tma_load_2d(reinterpret_cast<float*>(sp + DT_X_BYTES), mw, n0, 32 * (kc0 + s), full + slot);
          tma_load_2d(reinterpret_cast<float*>(sp + DT_X_BYTES + 4096), mw, n0 + 32, 32 * (kc0 + s), full + slot);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issue (one thread) ----------------
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (MODE == 1 ? (1u << 16) : 0u) | ((uint32_t)(DT_N >> 3) << 17) | ((uint32_t)(DT_M >> 4) << 24);
    // one elected lane issues; elect.sync (rather than a lane test) lets the compiler keep the descriptors in uniform
    // registers and emit the MMAs back to back -- a divergent branch costs ~70 cycles of uniform-datapath fix-up per MMA
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
    const uint64_t d_x0 = dt_desc(smem_u32(smem));
    const uint64_t d_w0 = MODE == 0 ? d_x0 + (uint64_t)(DT_X_BYTES >> 4) : dt_desc_mn(smem_u32(smem) + DT_X_BYTES);
    const uint64_t d_xl0 = d_x0 + (uint64_t)((DT_X_BYTES + DT_W_BYTES) >> 4), d_wl0 = d_w0 + (uint64_t)((DT_X_BYTES + DT_W_BYTES) >> 4);
    constexpr uint64_t WSTEP = MODE == 0 ? 2 : 64;             // B operand advance per 8-deep k-step: +32 B (K-major) / +1024 B (MN-major)
    for (int s = 0; s < nk; ++s) {
      const int slot = s % DT_STAGES;
      mbar_wait(full + slot, (s / DT_STAGES) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        mark(1, 2);
        const uint64_t so = (uint64_t)((slot * DT_STAGE_BYTES) >> 4);        // (descriptor addresses count 16-byte units)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t o = so + (uint64_t)(2 * j), ow = so + WSTEP * (uint64_t)j;   // A: + 32 B per 8-deep k-step
          dt_mma(tmem, d_xl0 + o, d_w0 + ow, idesc, (s | j) != 0 ? 1u : 0u);
          dt_mma(tmem, d_x0 + o, d_wl0 + ow, idesc, 1u);
          dt_mma(tmem, d_x0 + o, d_w0 + ow, idesc, 1u);
        }
        mark(1, 3);
        dt_commit(empty + slot);                              // frees the stage when these MMAs have read it
        if (s == nk - 1) dt_commit(&acc_done);
      }
      __syncwarp();
    }
  } else {
    // ---------------- the other warps: this CTA's bias / LayerNorm vectors ----------------
    const int wl = tid - 64;
    if (wl < 3 * DT_N) {
      const int a = wl / DT_N, j = wl - a * DT_N;
      const float* src = a == 0 ? op.bias : a == 1 ? op.gamma : op.beta;
      svec[a][j] = src ? ld_w(src + n0 + j) : 0.f;            // (made visible to the epilogue by its first __syncthreads)
    }
  }
  // ---------------- epilogue (all warps): bias, LayerNorm across the cluster, ELU, stores ----------------
  // dgrad mode: this thread's 16 columns of the producing layer's h / zh are requested now, ahead of the accumulator
  float4 hq[MODE == 1 ? DT_CW / 4 : 1], zq[MODE == 1 ? DT_CW / 4 : 1];
  if constexpr (MODE == 1) {
    const int row_ = (warp & 3) * 32 + lane, cb_ = (warp >> 2) * DT_CW;
    const bool ok = kh == 0 && r0 + row_ < p.B;
    const size_t off = (size_t)(ok ? r0 + row_ : 0) * p.N + n0 + cb_;
#pragma unroll
    for (int j = 0; j < DT_CW / 4; ++j) {
      hq[j] = ok ? ld_act4(op.H + off + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
      zq[j] = ok ? ld_act4(op.ZH + off + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  mbar_wait(&acc_done, 0);
  if (!p.late_trigger) pdl_launch();                          // the next kernel may set itself up under this epilogue
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 64) mark(2, 6);
  stamp(2);
  const int q = warp & 3, c = warp >> 2;
  const int row = q * 32 + lane;                              // TMEM lane = batch row of the tile
  const int cb = c * DT_CW;                                   // first of this thread's DT_CW features within the CTA slice
  float z[DT_CW];
  {
    uint32_t v[DT_CW];
    const uint32_t addr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cb;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < DT_CW; ++j) z[j] = __uint_as_float(v[j]);
  }
  __syncthreads();                                            // svec is complete
  if (tid == 64) mark(2, 8);
  if (p.ks > 1) {
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");      // phase 1: every peer is resident
    if (kh > 0) {
      const uint32_t local = smem_u32(part + row * DT_TP + cb);
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(fb));
#pragma unroll
      for (int j = 0; j < DT_CW; j += 4)
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote + 4u * j), "f"(z[j]), "f"(z[j + 1]), "f"(z[j + 2]), "f"(z[j + 3])
                     : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");    // phase 2: partials delivered
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (kh > 0) {
      // nothing else to compute here: keep the cluster barriers of the statistics exchange balanced, then leave
      asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
      if (p.late_trigger) pdl_launch();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      stamp(3);
      if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(DT_N));
      return;
    }
#pragma unroll
    for (int j = 0; j < DT_CW; j += 4) {
      const float4 q4 = *reinterpret_cast<const float4*>(part + row * DT_TP + cb + j);
      z[j] += q4.x; z[j + 1] += q4.y; z[j + 2] += q4.z; z[j + 3] += q4.w;
    }
  }
  if constexpr (MODE == 1) {
    // ---- dgrad epilogue: dU = dH * ELU'(h);  t = dU * gamma;  dZ = rs * (t - mean(t) - zh * mean(t * zh))  (row means over all N columns) ----
    float tv[DT_CW], zhv[DT_CW];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < DT_CW; ++j) {
      const float4 h4 = hq[j >> 2], z4 = zq[j >> 2];
      const float h = (j & 3) == 0 ? h4.x : (j & 3) == 1 ? h4.y : (j & 3) == 2 ? h4.z : h4.w;
      zhv[j] = (j & 3) == 0 ? z4.x : (j & 3) == 1 ? z4.y : (j & 3) == 2 ? z4.z : z4.w;
      const float du = z[j] * elu_grad_from_out(h);
      z[j] = du;
      tv[j] = du * svec[1][cb + j];
      s1 += tv[j];
      s2 = fmaf(tv[j], zhv[j], s2);
    }
    psum[c][row] = s1;
    pm2[c][row] = s2;
    __syncthreads();
    if (p.ks == 1) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (tid < DT_M * 2) {
      const int r_ = tid & (DT_M - 1);
      const float a1 = ((psum[0][r_] + psum[1][r_]) + psum[2][r_]) + psum[3][r_];
      const float a2 = ((pm2[0][r_] + pm2[1][r_]) + pm2[2][r_]) + pm2[3][r_];
      const uint32_t local = smem_u32(&xstat[fb][r_]);
      for (int r = tid >> 7; r < nfb; r += 2) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a1), "f"(a2) : "memory");
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    float c1 = 0.f, c2 = 0.f;
    for (int r = 0; r < nfb; ++r) {                           // slices in rank order
      const float2 pv = xstat[r][row];
      c1 += pv.x;
      c2 += pv.y;
    }
    const float fN = (float)p.N;
    c1 /= fN;
    c2 /= fN;
    const float rsr = (r0 + row < p.B) ? ld_act(op.RS + r0 + row) : 0.f;
    float* tile_u = reinterpret_cast<float*>(smem);
    float* tile_d = tile_u + DT_M * DT_TP;
#pragma unroll
    for (int j = 0; j < DT_CW; j += 4) {
      float4 u4 = make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]), d4;
      d4.x = rsr * (tv[j] - c1 - zhv[j] * c2);
      d4.y = rsr * (tv[j + 1] - c1 - zhv[j + 1] * c2);
      d4.z = rsr * (tv[j + 2] - c1 - zhv[j + 2] * c2);
      d4.w = rsr * (tv[j + 3] - c1 - zhv[j + 3] * c2);
      *reinterpret_cast<float4*>(tile_u + row * DT_TP + cb + j) = u4;
      *reinterpret_cast<float4*>(tile_d + row * DT_TP + cb + j) = d4;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < DT_M * DT_N / 4 / DT_THREADS; ++u) {
      const int f = tid + u * DT_THREADS, r = f / (DT_N / 4), c4 = f % (DT_N / 4);
      if (r0 + r < p.B) {
        const size_t off = (size_t)(r0 + r) * p.N + n0 + 4 * c4;
        *reinterpret_cast<float4*>(op.DU + off) = *reinterpret_cast<const float4*>(tile_u + r * DT_TP + 4 * c4);
        const float4 d4 = *reinterpret_cast<const float4*>(tile_d + r * DT_TP + 4 * c4);
        *reinterpret_cast<float4*>(op.DZ + off) = d4;
        if (op.OUThl != nullptr) {                            // dZ is the next operand of the layer-1 weight gradient
          float4 big, small;
          split_rn4(d4, big, small);
          *reinterpret_cast<float4*>(op.OUThl + off) = big;
          *reinterpret_cast<float4*>(op.OUThl + (size_t)p.B * p.N + off) = small;
        }
      }
    }
  } else {
#pragma unroll
  for (int j = 0; j < DT_CW; ++j) z[j] += svec[0][cb + j];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < DT_CW; ++j) s += z[j];
  psum[c][row] = s;
  __syncthreads();
  const float mean_c = (((psum[0][row] + psum[1][row]) + psum[2][row]) + psum[3][row]) * (1.0f / (float)DT_N);
  float m2 = 0.f;
#pragma unroll
  for (int j = 0; j < DT_CW; ++j) {
    const float d = z[j] - mean_c;
    m2 = fmaf(d, d, m2);
  }
  pm2[c][row] = m2;
  __syncthreads();
  if (p.ks == 1) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  if (tid == 64) mark(2, 9);
  if (tid < DT_M * 2) {
    // (row, peer-half) pairs: thread t sends row t & 127 to the peers of parity t >> 7
    const int r_ = tid & (DT_M - 1);
    const float mc = (((psum[0][r_] + psum[1][r_]) + psum[2][r_]) + psum[3][r_]) * (1.0f / (float)DT_N);
    const float m2c = ((pm2[0][r_] + pm2[1][r_]) + pm2[2][r_]) + pm2[3][r_];
    const uint32_t local = smem_u32(&xstat[fb][r_]);
    for (int r = tid >> 7; r < nfb; r += 2) {
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
      asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(mc), "f"(m2c) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  if (tid == 64) mark(2, 10);
  // combine the slices in rank order (same arithmetic as the mma.sync stage: count-weighted mean, Chan's M2 update)
  const float fN = (float)p.N;
  float msum = 0.f;
  for (int r = 0; r < nfb; ++r) msum = fmaf((float)DT_N, xstat[r][row].x, msum);
  const float mean = msum / fN;
  float m2t = 0.f;
  for (int r = 0; r < nfb; ++r) {
    const float2 pv = xstat[r][row];
    const float d = pv.x - mean;
    m2t += pv.y + (float)DT_N * d * d;
  }
  const float rstd = 1.0f / sqrtf(m2t / fN + SC_LN_EPS);
  // outputs go through a shared-memory tile (the operand ring is idle now) so that every 256-byte row segment is written
  // by 16 consecutive lanes
  float* tile_h = reinterpret_cast<float*>(smem);
  float* tile_z = tile_h + DT_M * DT_TP;
#pragma unroll
  for (int j = 0; j < DT_CW; j += 4) {
    const float4 g4 = *reinterpret_cast<const float4*>(&svec[1][cb + j]), b4 = *reinterpret_cast<const float4*>(&svec[2][cb + j]);
    float4 zh, h;
    zh.x = (z[j] - mean) * rstd;
    zh.y = (z[j + 1] - mean) * rstd;
    zh.z = (z[j + 2] - mean) * rstd;
    zh.w = (z[j + 3] - mean) * rstd;
    h.x = elu(fmaf(zh.x, g4.x, b4.x));
    h.y = elu(fmaf(zh.y, g4.y, b4.y));
    h.z = elu(fmaf(zh.z, g4.z, b4.z));
    h.w = elu(fmaf(zh.w, g4.w, b4.w));
    *reinterpret_cast<float4*>(tile_h + row * DT_TP + cb + j) = h;
    if (op.ZH != nullptr) *reinterpret_cast<float4*>(tile_z + row * DT_TP + cb + j) = zh;
  }
  if (fb == 0 && c == 0 && op.RS != nullptr && r0 + row < p.B) op.RS[r0 + row] = rstd;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < DT_M * DT_N / 4 / DT_THREADS; ++u) {
    const int f = tid + u * DT_THREADS, r = f / (DT_N / 4), c4 = f % (DT_N / 4);
    if (r0 + r < p.B) {
      const size_t off = (size_t)(r0 + r) * p.N + n0 + 4 * c4;
      const float4 h4 = *reinterpret_cast<const float4*>(tile_h + r * DT_TP + 4 * c4);
      *reinterpret_cast<float4*>(op.H + off) = h4;
      if (op.OUThl != nullptr) {                              // H is the X operand of the next layer and of its weight gradient
        float4 big, small;
        split_rn4(h4, big, small);
        *reinterpret_cast<float4*>(op.OUThl + off) = big;
        *reinterpret_cast<float4*>(op.OUThl + (size_t)p.B * p.N + off) = small;
      }
      if (op.ZH != nullptr) *reinterpret_cast<float4*>(op.ZH + off) = *reinterpret_cast<const float4*>(tile_z + r * DT_TP + 4 * c4);
    }
  }
  }
  if (tid == 64) mark(2, 7);
  if (p.late_trigger) pdl_launch();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  stamp(3);
  if (p.tl != nullptr && blockIdx.x == 0 && tid < 3) {
    p.tl[64 * tid] = (unsigned long long)tln[tid];
    for (int i = 0; i < tln[tid]; ++i) p.tl[64 * tid + 1 + i] = tlbuf[tid][i];
  }
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(DT_N));
}

// minibatch gather: replay rows [C][Sp] -> dense TMA-readable buffers [B][Sp32] (zero padded), states and next states:
// the fp32 rows (X operand of the layer-1 weight gradients) and their tf32 big / small parts (X operand of k_dense_tc,
// which then skips the split of the X boxes in every one of its CTAs)
__global__ void __launch_bounds__(128) k_gather_mb(const GatherP* __restrict__ pp) {
  const GatherP P = pp[blockIdx.y];
  float *mb_s = P.mb_s, *mb_s2 = P.mb_s2, *hl = P.hl;
  const float* __restrict__ state = P.state;
  const float* __restrict__ new_state = P.new_state;
  const int* __restrict__ bidx = P.bidx;
  const int B = P.B, Sp = P.Sp, Sp32 = P.Sp32;
  const int r = blockIdx.x >> 1, which = blockIdx.x & 1;
  pdl_launch();
  pdl_wait();
  if (P.bidx_mirror != nullptr && which == 0 && threadIdx.x == 0) P.bidx_mirror[r] = __ldg(bidx + r);
  const float* src = (which ? new_state : state) + (size_t)__ldg(bidx + r) * Sp;
  float* dst = (which ? mb_s2 : mb_s) + (size_t)r * Sp32;
  float* dhi = hl + ((size_t)(2 * which) * B + r) * Sp32;    // hl = [s_hi | s_lo | s2_hi | s2_lo], each [B][Sp32]
  float* dlo = dhi + (size_t)B * Sp32;
  const int n4 = Sp32 >> 2, v4 = Sp >> 2;
  for (int c4 = threadIdx.x; c4 < n4; c4 += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c4 < v4) v = ld_w4(src + 4 * c4);
    *reinterpret_cast<float4*>(dst + 4 * c4) = v;
    if (hl != nullptr) {
      float4 h, y;
      split_rn4(v, h, y);
      *reinterpret_cast<float4*>(dhi + 4 * c4) = h;
      *reinterpret_cast<float4*>(dlo + 4 * c4) = y;
    }
  }
}

// big / small parts of whole arrays: parameter arenas after they were written from outside (load_state_dict, soft update,
// the FP32 weight-gradient kernel) and the activations of the fallback kernels that do not write them themselves
__global__ void __launch_bounds__(256) k_split(const SplitP* __restrict__ pp) {
  const SplitP& P = pp[blockIdx.y];
  pdl_launch();
  pdl_wait();
  for (int e = 0; e < P.n; ++e) {
    const float4* __restrict__ src = reinterpret_cast<const float4*>(P.src[e]);
    float4* hi = reinterpret_cast<float4*>(P.hl[e]);
    float4* lo = reinterpret_cast<float4*>(P.hl[e] + P.plane[e]);
    const unsigned n4 = P.count[e] >> 2;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      float4 big, small;
      split_rn4(__ldcg(src + i), big, small);
      hi[i] = big;
      lo[i] = small;
    }
  }
}
cudaError_t launch_split(const SplitP* pp, int E, int grid, cudaStream_t st) {
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(grid, E, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_split, pp);
}

// Experiment behind SMARTCAL_PREFETCH=1 (off by default; sc_api.cu plan_step has the measurement): the step's parameter traffic
// (weights as big / small parts, fp32 parameters, targets, Adam moments: ~32 MB) requested from HBM at the start of learn().
__global__ void __launch_bounds__(256) k_l2_prefetch(const PrefetchP* __restrict__ pp) {
  const PrefetchP& P = pp[blockIdx.y];
  size_t base = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x, me = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int e = 0; e < P.n; ++e) {
    const size_t lines = (size_t)((P.bytes[e] + 127) >> 7);
    const char* p = reinterpret_cast<const char*>(P.ptr[e]);
    // (lines of all ranges are dealt round-robin over the grid in range order, so the first range is requested first)
    for (size_t i = (me + stride - base % stride) % stride; i < lines; i += stride)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (i << 7)));
    base += lines;
  }
}
cudaError_t launch_l2_prefetch(const PrefetchP* pp, int E, int grid, cudaStream_t st) {
  k_l2_prefetch<<<dim3(grid, E, 1), 256, 0, st>>>(pp);
  return cudaGetLastError();
}

size_t dense_tc_smem_bytes() { return (size_t)DT_STAGES * DT_STAGE_BYTES + DT_PART_BYTES + 1024; }
size_t dgrad_tc_smem_bytes() { return dense_tc_smem_bytes(); }

// co-resident clusters of `cs` CTAs this device can host for k_dense_tc (0: such clusters cannot be launched at all)
int dense_tc_max_clusters(int cs) {
  if (cudaFuncSetAttribute(k_dense_tc_t<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_tc_smem_bytes()) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cs > 8 && cudaFuncSetAttribute(k_dense_tc_t<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(cs * 8, 1, 1);
  cfg.blockDim = dim3(DT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = dense_tc_smem_bytes();
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_dense_tc_t<0>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static bool dt_first_use(bool (&done)[64]) {
  int d = 0;
  cudaGetDevice(&d);
  d &= 63;
  if (done[d]) return false;
  done[d] = true;
  return true;
}

// pp: device array of E parameter blocks (one per agent, identical geometry nops, nrb, N, ks; every op's K / 32 divisible by ks)
cudaError_t launch_dense_tc(const DenseTcParams* pp, int E, int nops, int nrb, int N, int ks, cudaStream_t st) {
  static bool configured[64], big[64];
  if (dt_first_use(configured)) {
    cudaError_t e = cudaFuncSetAttribute(k_dense_tc_t<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense_tc_smem_bytes());
    if (e != cudaSuccess) return e;
  }
  const int nfb = N / DT_N, cs = nfb * ks;
  if (nfb < 1 || nfb > DT_MAXCS || nfb * DT_N != N || ks < 1 || ks > 2 || cs > 16) return cudaErrorInvalidValue;
  if (cs > 8 && dt_first_use(big)) {
    cudaError_t e = cudaFuncSetAttribute(k_dense_tc_t<0>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[2];
  cfg.gridDim = dim3(nops * nrb * cs, E, 1);
  cfg.blockDim = dim3(DT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = dense_tc_smem_bytes();
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, k_dense_tc_t<0>, pp);
}

// the wide input-gradient GEMM (MODE 1): same geometry rules as launch_dense_tc
cudaError_t launch_dgrad_tc(const DenseTcParams* pp, int E, int nops, int nrb, int N, int ks, cudaStream_t st) {
  static bool configured[64], big[64];
  if (dt_first_use(configured)) {
    cudaError_t e = cudaFuncSetAttribute(k_dense_tc_t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dgrad_tc_smem_bytes());
    if (e != cudaSuccess) return e;
  }
  const int nfb = N / DT_N, cs = nfb * ks;
  if (nfb < 1 || nfb > DT_MAXCS || nfb * DT_N != N || ks < 1 || ks > 2 || cs > 16) return cudaErrorInvalidValue;
  if (cs > 8 && dt_first_use(big)) {
    cudaError_t e = cudaFuncSetAttribute(k_dense_tc_t<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[2];
  cfg.gridDim = dim3(nops * nrb * cs, E, 1);
  cfg.blockDim = dim3(DT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = dgrad_tc_smem_bytes();
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, k_dense_tc_t<1>, pp);
}
int dgrad_tc_max_clusters(int cs) {
  if (cudaFuncSetAttribute(k_dense_tc_t<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dgrad_tc_smem_bytes()) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(cs * 8, 1, 1);
  cfg.blockDim = dim3(DT_THREADS, 1, 1);
  cfg.dynamicSmemBytes = dgrad_tc_smem_bytes();
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, k_dense_tc_t<1>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

cudaError_t launch_gather_mb(const GatherP* pp, int E, int B, cudaStream_t st) {
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(2 * B, E, 1);
  cfg.blockDim = dim3(128, 1, 1);
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_gather_mb, pp);
}

}  // namespace sc
