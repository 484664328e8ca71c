// Device-side helpers shared by the sm_100a kernels of libsmartcal_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SC_THREADS 256          // compute ("consumer") threads per CTA in every kernel
#define SC_WARPS 8
#define SC_ROW_THREADS 288      // row-phase kernel: 8 consumer warps + 1 TMA producer warp
#define SC_MT 32                // batch rows per GEMM sub-tile
#define SC_MAX_CLUSTER 16
#define SC_LN_EPS 1e-5f         // nn.LayerNorm default eps (ref enet_sac.py:365-368)

namespace sc {

// ---- thread-block-cluster primitives (PTX; one cluster owns a slab of batch rows) ---------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_idx() {
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned num_clusters() {
  unsigned r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
// Release/acquire barrier over every thread of the cluster: global-memory writes made before it by any CTA
// of the cluster are visible to every CTA of the cluster after it.
__device__ __forceinline__ void cluster_sync() {
  // generic-proxy global writes made before the barrier are later read through the async proxy (TMA bulk copies)
  asm volatile("fence.proxy.async;\n" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become
// resident while its predecessor in the stream still runs; pdl_wait() blocks until that predecessor has completed and its
// writes are visible, pdl_launch() lets the successor start its own prologue early.  Both are no-ops otherwise.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// Every kernel of a learn() step takes its parameter block from a device-resident array with one entry per agent of the
// launch (blockIdx.y = agent; a stand-alone agent is an ensemble of one).  The block is written by the single host->device
// copy at the head of the step (sc_api.cu "step tape"), long before any kernel of the step starts, so fetching it here --
// ahead of griddepcontrol.wait -- is safe under programmatic dependent launch.  All threads of the CTA must call this.
template <typename P>
__device__ __forceinline__ const P& load_params(const P* __restrict__ pp, P* smem_copy) {
  static_assert(sizeof(P) % 4 == 0, "parameter blocks are copied word by word");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(pp + blockIdx.y);
  uint32_t* dst = reinterpret_cast<uint32_t*>(smem_copy);
  for (int i = threadIdx.x; i < (int)(sizeof(P) / 4); i += blockDim.x) dst[i] = __ldg(src + i);
  __syncthreads();
  return *smem_copy;
}

// barrier over the 256 consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

// ---- mbarrier + TMA bulk copy (1-D cp.async.bulk, global -> shared, completion counted in bytes) -----------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// the same wait for warps that are NOT on the critical path (epilogue warps waiting for a whole K loop, producers waiting for a
// free slot): between polls the warp sleeps, so that its polling does not take issue slots from the warps that do the work
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* b, unsigned parity, unsigned ns) {
  for (;;) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void bulk_g2s(float* smem_dst, const float* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D tiled TMA load through a CUtensorMap that lives in global memory: box {32 floats, rows} with 128-byte swizzle
__device__ __forceinline__ void tma_load_2d(float* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// ---- loads ------------------------------------------------------------------------------------------
// activations written earlier in the SAME launch by other CTAs: L2-coherent load (never the .nc path)
__device__ __forceinline__ float4 ld_act4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ld_act(const float* p) { return __ldcg(p); }
// weights / replay rows: constant for the duration of a launch -> read-only path
__device__ __forceinline__ float4 ld_w4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float ld_w(const float* p) { return __ldg(p); }

// ---- warp reductions (deterministic butterfly: every lane ends with the same value) --------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group8_sum(float v) {  // sum over aligned groups of 8 lanes
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// ---- activation -----------------------------------------------------------------------------------------
// F.elu, alpha = 1: x > 0 ? x : exp(x) - 1   (ATen computes exp(x) - 1, not expm1)
// x = big + small with big = x rounded to tf32 (so the tensor core's truncation leaves it unchanged) and small = the
// remainder rounded to tf32: the operand pair of the 3xTF32 products.  Every producer of a tensor-core operand (gather,
// dense / dgrad epilogues, tail, Adam epilogue, k_split) writes the two parts next to the fp32 array ("hl" = [big | small]).
__device__ __forceinline__ void split_rn(float x, float& big, float& small) {
  uint32_t b, s;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(x));
  big = __uint_as_float(b);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(s) : "f"(x - big));
  small = __uint_as_float(s);
}
__device__ __forceinline__ void split_rn4(const float4 x, float4& big, float4& small) {
  split_rn(x.x, big.x, small.x);
  split_rn(x.y, big.y, small.y);
  split_rn(x.z, big.z, small.z);
  split_rn(x.w, big.w, small.w);
}
__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : expf(x) - 1.f; }
// d elu / dx expressed through the OUTPUT h: 1 for h > 0, exp(x) = h + 1 otherwise
__device__ __forceinline__ float elu_grad_from_out(float h) { return h > 0.f ? 1.f : h + 1.f; }

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

// smallest pitch >= k (k multiple of 4) with pitch % 32 == 4: rows r and r+8i land on disjoint banks for
// 128-bit loads (see DESIGN.md "shared-memory tiles")
__host__ __device__ __forceinline__ int smem_pitch(int k) { return k + ((4 - (k & 31)) + 32) % 32; }

// ---- Philox4x32-10 counter RNG (device noise when the caller does not inject the reference streams) ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u32_to_unit_open(uint32_t x) {  // (0,1]
  return (float)(x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
}
// two independent N(0,1) from one Philox block (Box-Muller)
__device__ __forceinline__ float2 philox_normal2(uint64_t seed, uint64_t stream, uint64_t idx) {
  uint4 c = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream, (uint32_t)(stream >> 32));
  uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float u1 = u32_to_unit_open(r.x), u2 = u32_to_unit_open(r.y);
  float rad = sqrtf(-2.f * logf(u1));
  float s, co;
  sincospif(2.f * u2, &s, &co);
  return make_float2(rad * co, rad * s);
}
__device__ __forceinline__ double philox_uniform_double(uint64_t seed, uint64_t stream, uint64_t idx) {
  uint4 c = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)stream, (uint32_t)(stream >> 32));
  uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint64_t m = (((uint64_t)r.x << 32) | r.y) >> 11;  // 53 bits
  return (double)m * (1.0 / 9007199254740992.0);
}

}  // namespace sc
