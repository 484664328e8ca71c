// Stand-alone probe (not part of the product): issue-to-completion cost of back-to-back tcgen05.mma instructions from one
// CTA, for the operand placements and shapes the dense forward kernel could use.  Operands are zero-filled shared memory /
// tensor memory laid out as valid K-major SWIZZLE_128B tiles; only timing is reported (clock64 around issue + commit wait).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// mode 0: SS (A, B in shared memory), 1: TS (A in tensor memory).  kind 0: tf32 (K = 8), 1: bf16 (K = 16).
template <int mode, int kind>
__global__ void __launch_bounds__(128) probe(int M, int N, int nacc, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 192 * 1024 / 16; i += 128) reinterpret_cast<float4*>(base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  {
    // zero the whole tensor memory (accumulators and the TS-mode A operand)
    for (int c0 = 0; c0 < 512; c0 += 8) {
      const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(addr), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t sA = smem_u32(base), sB = sA + 64 * 1024;     // A: 4 chunks of [128 rows][128 B]; B: 4 chunks of [256 rows][128 B]
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 0) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
    if (leader) {
      const uint32_t fmt = kind == 0 ? 2u : 1u;                  // tf32 = 2, bf16 = 1
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      const uint64_t da0 = desc_k_sw128(sA), db0 = desc_k_sw128(sB);
      const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? (uint32_t)N : 0u), a0 = tmem + 256;
      t0 = clock64();
      for (int it = 0; it < iters; it += 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          // k-step u: chunk u >> 2 (16 KB of A, 32 KB of B apart), 32 B per k-step inside the 128-byte rows
          const uint64_t da = da0 + (uint64_t)(((u >> 2) * (128 * 128) + (u & 3) * 32) >> 4);
          const uint64_t db = db0 + (uint64_t)(((u >> 2) * (256 * 128) + (u & 3) * 32) >> 4);
          const uint32_t d_addr = (u & 1) ? d1 : d0;
          const uint32_t a_addr = a0 + (uint32_t)((u & 7) * 8);
          if (mode == 0) {
            if (kind == 0)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
          } else {
            if (kind == 0)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_addr), "r"(a_addr), "l"(db), "r"(idesc), "r"(1u) : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_addr), "r"(a_addr), "l"(db), "r"(idesc), "r"(1u) : "memory");
          }
        }
      }
      t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = 192 * 1024 + 1024;
  cudaFuncSetAttribute(probe<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Case { int mode, kind, M, N, nacc; };
  const Case cases[] = {
      {0, 0, 128, 64, 1}, {0, 0, 128, 128, 1}, {0, 0, 128, 256, 1}, {0, 0, 128, 64, 2}, {0, 0, 128, 64, 4}, {0, 0, 128, 128, 2},
      {0, 0, 64, 64, 1},  {0, 0, 64, 128, 1},  {0, 0, 64, 256, 1},  {0, 0, 128, 32, 1}, {0, 0, 128, 16, 1},
      {1, 0, 128, 64, 1}, {1, 0, 128, 128, 1}, {1, 0, 128, 256, 1}, {1, 0, 128, 64, 2}, {1, 0, 64, 64, 1},
      {0, 1, 128, 64, 1}, {0, 1, 128, 128, 1}, {0, 1, 128, 256, 1}, {1, 1, 128, 256, 1},
  };
  for (const Case& c : cases) {
    long long best[2] = {1ll << 60, 1ll << 60};
    const int iters = 192;
    for (int rep = 0; rep < 3; ++rep) {
      if (c.mode == 0 && c.kind == 0) probe<0, 0><<<1, 128, smem>>>(c.M, c.N, c.nacc, iters, d);
      if (c.mode == 1 && c.kind == 0) probe<1, 0><<<1, 128, smem>>>(c.M, c.N, c.nacc, iters, d);
      if (c.mode == 0 && c.kind == 1) probe<0, 1><<<1, 128, smem>>>(c.M, c.N, c.nacc, iters, d);
      if (c.mode == 1 && c.kind == 1) probe<1, 1><<<1, 128, smem>>>(c.M, c.N, c.nacc, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s (mode %d kind %d M %d N %d)\n", cudaGetErrorString(e), c.mode, c.kind, c.M, c.N); return 1; }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (h[1] < best[1]) { best[0] = h[0]; best[1] = h[1]; }
    }
    const int K = c.kind == 0 ? 8 : 16;
    const double clk = (double)best[1] / iters;
    printf("%s %s M=%3d N=%3d K=%2d accumulators=%d: issue %.1f clk/MMA, complete %.1f clk/MMA -> %.0f MAC/clk, operand bytes/clk %.1f\n",
           c.mode == 0 ? "SS" : "TS", c.kind == 0 ? "tf32" : "bf16", c.M, c.N, K, c.nacc, (double)best[0] / iters, clk,
           (double)c.M * c.N * K / clk, ((c.mode == 0 ? c.M : 0) + c.N) * 32.0 / clk);
  }
  return 0;
}
