// Stand-alone probe (not part of the product): tcgen05.mma kind::tf32 with the A operand in tensor memory (TS mode),
// M = 64 output features, N = 8 batch rows, K = 64: D[f][r] = sum_k W[f][k] * X[r][k].
// A (64 features x 64 k) is written to TMEM with tcgen05.st from registers (lane = feature in the M=64 "half subpartition"
// layout: feature 16 q + t -> lane 32 q + t), B = X [8 rows][64 k] sits in shared memory K-major with the 128-byte swizzle
// exactly as a TMA box {32 floats, 8 rows} x 2 k-blocks produces it.  Prints max |D - reference|.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int M = 64, N = 8, K = 64;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
  // K-major, SWIZZLE_128B: rows of 128 B, 8-row atoms of 1024 B (SBO), LBO unused (1)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ Wg, const float* __restrict__ Xg, float* __restrict__ Dg, int a_col_per_k) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* sX = (float*)base;                       // [K/32 blocks][8 rows][32 floats], swizzled
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N * K; i += 128) {
    const int r = i / K, k = i % K;
    const int kb = k >> 5, c = (k & 31) >> 2, w = k & 3;
    sX[(kb * N + r) * 32 + (((c ^ (r & 7)) << 2) | w)] = Xg[r * K + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t tA = tmem + 32;                  // A at columns [32, 32 + 64), D at columns [0, 8)
  // ---- A -> TMEM: warp q holds features 16 q + t (t < 16); lanes 16..31 write zeros into unused lanes ----
  {
    const int f = 16 * warp + (lane & 15);
    for (int c0 = 0; c0 < K; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = lane < 16 ? __float_as_uint(Wg[f * K + c0 + j]) : 0u;
      const uint32_t addr = tA + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c0 * a_col_per_k);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                   "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    // D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint32_t a_addr = tA + (uint32_t)(ks * 8 * a_col_per_k);
      const uint64_t db = desc_k_sw128(smem_u32(sX) + (ks >> 2) * (N * 128) + (ks & 3) * 32);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem), "r"(a_addr), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t v[8];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) Dg[(size_t)tid * 8 + j] = __uint_as_float(v[j]);      // [TMEM lane][row]
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
  std::vector<float> W(M * K), X(N * K), D(128 * 8), R(M * N);
  srand(2);
  for (auto& v : W) v = (float)(rand() % 7 - 3);
  for (auto& v : X) v = (float)(rand() % 5 - 2);
  for (int f = 0; f < M; ++f)
    for (int r = 0; r < N; ++r) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += W[f * K + k] * X[r * K + k];
      R[f * N + r] = s;
    }
  float *dW, *dX, *dD;
  cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = N * K * 4 + 2048;
  for (int cpk = 1; cpk <= 1; ++cpk) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, smem>>>(dW, dX, dD, cpk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0; int bad = 0;
    for (int f = 0; f < M; ++f)
      for (int r = 0; r < N; ++r) {
        const int lane = (f >> 4) * 32 + (f & 15);
        const double d = fabs((double)D[lane * 8 + r] - R[f * N + r]);
        if (d > mx) mx = d;
        if (d > 0.5) ++bad;
      }
    printf("TS mode, M=64 half-subpartition layout: max err %.3f, wrong %d / %d; D[f=0][0..3] = %.1f %.1f %.1f %.1f (ref %.1f %.1f %.1f %.1f); D[f=17][0] lane48 = %.1f (ref %.1f)\n",
           mx, bad, M * N, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3], D[(32 + 1) * 8], R[17 * N]);
  }
  return 0;
}
