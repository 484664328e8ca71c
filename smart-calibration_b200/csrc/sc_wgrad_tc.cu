// Weight gradients of the wide layers on the 5th-generation tensor cores: dW[n][k] = sum_b dZ[b][n] X[b][k] as a
// tcgen05.mma kind::tf32 GEMM with M = 128 output rows (n), N = 64 output columns (k) and the batch as the reduction,
// accumulator in tensor memory, followed by the same fused torch-exact Adam + target soft update as sc_wgrad.cu.
//
// Both operands are "MN-major" for this product (dZ is n-contiguous, X is k-contiguous), which for 32-bit elements means
// the SWIZZLE_128B_BASE32B canonical layout: TMA boxes of [32 batch rows][32 floats] written with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, described to the MMA with leading-byte-offset = one box (next 32 n / k), stride-byte-
// offset = 512 B (next 4 batch rows) and a start address that advances 1024 B per 8-row k-step (validated against a CPU GEMM
// with tools/probe/tc_probe.cu).  kind::tf32 truncates its fp32 inputs to 10 mantissa bits, so fp32 fidelity is restored with
// three products per k-step: A*B + A_lo*B + A*B_lo.  Both operands arrive already split into tf32 big / small parts by their
// producers (k_gather_mb, the k_dense_tc epilogues, k_tail), so the K loop is TMA -> MMA; the Adam epilogue in turn writes the
// big / small parts of the updated weights (and targets) that the forward and dgrad GEMMs of the next pass read.
//
// Warp roles (512 threads): warp 0 = TMA producer, warp 1 = TMEM allocation + MMA issue (one thread);
// epilogue: warps 2..5 move the accumulator TMEM -> registers -> shared tile, then ALL warps apply Adam (its IEEE divisions and
// square roots make the epilogue latency-bound on few threads).  Three 48 KB stages of 32 batch rows.
#include <stdlib.h>
#include "sc_common.cuh"
#include "sc_params.h"

namespace sc {

constexpr int TC_M = 128, TC_N = 64, TC_BR = 32;            // tile rows (n), tile columns (k), batch rows per stage
constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = (TC_M / 32) * TC_BR * 128;       // 16 KB: 4 boxes of [32 b][32 n]
constexpr int TC_B_BYTES = (TC_N / 32) * TC_BR * 128;       // 8 KB
constexpr int TC_STAGE_BYTES = 2 * (TC_A_BYTES + TC_B_BYTES);   // big + small parts
constexpr int TC_TILE_PITCH = TC_N + 4;
constexpr int TC_THREADS = 512;
constexpr int TC_SPLIT_THREADS = TC_THREADS - 64;          // warps 2..11

__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  // MN-major, SWIZZLE_128B_BASE32B: LBO = 4096 B (next box of 32 MN elements), SBO = 512 B (next 4 k-rows), version 1
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((4096u >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void adam_math_tc(float g, float& m, float& v, float& p, const AdamHyper& h) {
  m = fmaf(h.one_minus_b1, g - m, m);
  v = __fadd_rn(__fmul_rn(v, h.b2), __fmul_rn(__fmul_rn(h.one_minus_b2, g), g));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), h.bc2_sqrt), h.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(h.lr_neg_step, m), denom));
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_wgrad_tc(const WgradTcParams* __restrict__ pp) {
  __shared__ WgradTcParams p_sm;
  const WgradTcParams& p = load_params(pp, &p_sm);            // blockIdx.y = agent of the ensemble
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[TC_STAGES], empty[TC_STAGES], acc_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // debug timeline: thread `who` of CTA 0 appends (globaltimer << 8 | tag) at tl[1 + 64 * who_slot + n]
  auto mark = [&](int slot, int tag) {
    if (p.tl != nullptr && blockIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      unsigned long long* q = p.tl + 64 * slot;
      const unsigned long long n = q[0];
      if (n < 62) { q[1 + n] = (t << 8) | (unsigned long long)tag; q[0] = n + 1; }
    }
  };
  if (p.tl != nullptr && blockIdx.x == 0 && tid < 3) p.tl[64 * tid] = 0;
  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    mbar_init(&acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TC_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();                                                 // (programmatic dependent launch: the set-up above ran under the previous kernel)
  const uint32_t tmem = tmem_slot;
  const int nstages = (p.B + TC_BR - 1) / TC_BR;
  unsigned it = 0;                                            // stages streamed so far by this CTA (all roles count alike)
  unsigned tiles_done = 0;

  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tiles_done) {
    const WgradTcOp op = p.ops[__ldg(p.tile_map + tile)];
    const int lt = tile - op.tile_base;
    const int tn = lt / op.tiles_k, tk = lt - tn * op.tiles_k;
    const int n0 = tn * TC_M, k0 = tk * TC_N;
    // The Adam operands of this tile (parameter, both moments, target) do not depend on the GEMM: every thread requests its 4 x 4
    // float4 of them before the K loop (the producer / MMA warps after theirs), so the epilogue starts with them in registers
    const bool soft = op.T != nullptr && p.do_soft;
    constexpr int NF = TC_M * TC_N / 4;                     // float4 elements of the tile
    constexpr int EU = NF / TC_THREADS;                     // float4 elements per thread
    static_assert(EU * TC_THREADS == NF, "tile / thread mismatch");
    float4 pm[EU], pv[EU], pp[EU], pt[EU];
    bool ok[EU];
    size_t idx[EU];
    auto load_adam_operands = [&]() {
#pragma unroll
      for (int u = 0; u < EU; ++u) {
        const int f = tid + u * TC_THREADS, r = f >> 4, c4 = f & 15;
        ok[u] = (n0 + r < op.N) && (k0 + 4 * c4 < op.K);
        idx[u] = ok[u] ? (size_t)(n0 + r) * op.ldw + k0 + 4 * c4 : 0;
        pm[u] = pv[u] = pp[u] = pt[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok[u]) {
          pm[u] = __ldcs(reinterpret_cast<const float4*>(op.M + idx[u]));
          pv[u] = __ldcs(reinterpret_cast<const float4*>(op.V + idx[u]));
          pp[u] = *reinterpret_cast<const float4*>(op.W + idx[u]);
          if (soft) pt[u] = *reinterpret_cast<const float4*>(op.T + idx[u]);
        }
      }
    };
    if (warp >= 2) load_adam_operands();
    if (warp == 0) {
      // ---------------- TMA producer ----------------
      for (int s = 0; s < nstages; ++s) {
        const unsigned g = it + s, slot = g % TC_STAGES;
        if (g >= TC_STAGES) mbar_wait(empty + slot, ((g / TC_STAGES) - 1) & 1u);
        uint8_t* st = smem + (size_t)slot * TC_STAGE_BYTES;
        if (lane == 0) { mark(0, 1); mbar_expect_tx(full + slot, TC_STAGE_BYTES); }
        __syncwarp();
        // stage layout [dZ big | X big | dZ small | X small], 4 KB boxes; lanes 0..5 fetch the big parts, 6..11 the small parts
        constexpr int NBOX = TC_M / 32 + TC_N / 32;
        if (lane < 2 * NBOX) {
          const int part = lane / NBOX, b = lane - part * NBOX;
          uint8_t* sp = st + part * (TC_A_BYTES + TC_B_BYTES);
          if (b < TC_M / 32)
            tma_load_2d(reinterpret_cast<float*>(sp + b * 4096), reinterpret_cast<const uint8_t*>(op.mapD) + 128 * part, n0 + 32 * b, s * TC_BR, full + slot);
          else
            tma_load_2d(reinterpret_cast<float*>(sp + TC_A_BYTES + (b - TC_M / 32) * 4096), reinterpret_cast<const uint8_t*>(op.mapX) + 128 * part,
                        k0 + 32 * (b - TC_M / 32), s * TC_BR, full + slot);
        }
      }
    } else if (warp == 1) {
      // ---------------- MMA issue (one thread) ----------------
      // instruction descriptor: D = F32, A = B = TF32, both MN-major, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
      for (int s = 0; s < nstages; ++s) {
        const unsigned g = it + s, slot = g % TC_STAGES;
        mbar_wait(full + slot, (g / TC_STAGES) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          mark(1, 2);
          const uint32_t a_raw = smem_u32(smem + (size_t)slot * TC_STAGE_BYTES), b_raw = a_raw + TC_A_BYTES;
          const uint32_t a_lo = b_raw + TC_B_BYTES, b_lo = a_lo + TC_A_BYTES;
#pragma unroll
          for (int j = 0; j < TC_BR / 8; ++j) {
            const uint32_t o = (uint32_t)j * 1024u;
            tc_mma(tmem, tc_desc(a_lo + o), tc_desc(b_raw + o), idesc, (s | j) != 0 ? 1u : 0u);
            tc_mma(tmem, tc_desc(a_raw + o), tc_desc(b_lo + o), idesc, 1u);
            tc_mma(tmem, tc_desc(a_raw + o), tc_desc(b_raw + o), idesc, 1u);
          }
          mark(1, 3);
          tc_commit(empty + slot);                            // frees the stage when these MMAs have read it
          if (s == nstages - 1) tc_commit(&acc_done);
        }
        __syncwarp();
      }
    } else {
      const int wl = tid - 64;                                // 0..447
      // ---------------- epilogue part 1 (warps 2..5): TMEM -> shared tile ----------------
      if (warp < 6) {
        mbar_wait(&acc_done, tiles_done & 1u);
        if (wl == 0) mark(2, 6);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* tile_w = reinterpret_cast<float*>(smem);       // all MMAs of this tile are complete: stage 0 is free
        const int row = (warp & 3) * 32 + lane;               // TMEM lane this thread can read
#pragma unroll 1
        for (int c0 = 0; c0 < TC_N; c0 += 16) {               // (16 columns at a time: the Adam operands are live in registers too)
          uint32_t v[16];
          const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
              : "r"(addr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(tile_w + row * TC_TILE_PITCH + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
    }
    if (warp < 2) load_adam_operands();
    __syncthreads();                                          // tile is in shared memory; every warp joins the Adam pass
    if (tid == 64) mark(2, 7);
    {
      // ---------------- epilogue part 2 (all warps): Adam (+ soft update), coalesced ----------------
      const float* tile_s = reinterpret_cast<const float*>(smem);
      const AdamHyper h = p.hyper[op.opt];
      {
        const int f0 = tid;
#pragma unroll
        for (int u = 0; u < EU; ++u) {
          if (!ok[u]) continue;
          const int f = f0 + u * TC_THREADS, r = f >> 4, c4 = f & 15;
          const float4 g4 = *reinterpret_cast<const float4*>(tile_s + r * TC_TILE_PITCH + 4 * c4);
          float4 m = pm[u], v = pv[u], w = pp[u];
          adam_math_tc(g4.x, m.x, v.x, w.x, h);
          adam_math_tc(g4.y, m.y, v.y, w.y, h);
          adam_math_tc(g4.z, m.z, v.z, w.z, h);
          adam_math_tc(g4.w, m.w, v.w, w.w, h);
          __stcs(reinterpret_cast<float4*>(op.M + idx[u]), m);
          __stcs(reinterpret_cast<float4*>(op.V + idx[u]), v);
          *reinterpret_cast<float4*>(op.W + idx[u]) = w;
          if (op.Whi != nullptr) {                            // operand copies for the next forward / dgrad GEMMs
            float4 big, small;
            split_rn4(w, big, small);
            *reinterpret_cast<float4*>(op.Whi + idx[u]) = big;
            *reinterpret_cast<float4*>(op.Wlo + idx[u]) = small;
          }
          if (soft) {
            float4 t = pt[u];
            t.x = __fadd_rn(__fmul_rn(p.tau, w.x), __fmul_rn(p.one_minus_tau, t.x));
            t.y = __fadd_rn(__fmul_rn(p.tau, w.y), __fmul_rn(p.one_minus_tau, t.y));
            t.z = __fadd_rn(__fmul_rn(p.tau, w.z), __fmul_rn(p.one_minus_tau, t.z));
            t.w = __fadd_rn(__fmul_rn(p.tau, w.w), __fmul_rn(p.one_minus_tau, t.w));
            *reinterpret_cast<float4*>(op.T + idx[u]) = t;
            if (op.Thi != nullptr) {
              float4 big, small;
              split_rn4(t, big, small);
              *reinterpret_cast<float4*>(op.Thi + idx[u]) = big;
              *reinterpret_cast<float4*>(op.Tlo + idx[u]) = small;
            }
          }
        }
      }
    }
    if (tid == 64) mark(2, 8);
    it += (unsigned)nstages;
    __syncthreads();        // the shared tile (stage 0) and the accumulator are free again before the next tile starts
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_N));
}

size_t wgrad_tc_smem_bytes() { return (size_t)TC_STAGES * TC_STAGE_BYTES + 1024; }

cudaError_t launch_wgrad_tc(const WgradTcParams* pp, int E, int grid, cudaStream_t st) {
  static bool configured[64];
  int d = 0;
  cudaGetDevice(&d);
  d &= 63;
  if (!configured[d]) {
    cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wgrad_tc_smem_bytes());
    if (e != cudaSuccess) return e;
    configured[d] = true;
  }
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("SMARTCAL_PDL_WGRAD"); pdl = (e && e[0] == '0') ? 0 : 1; }
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(grid, E, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = wgrad_tc_smem_bytes();
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_wgrad_tc, pp);
}

}  // namespace sc
