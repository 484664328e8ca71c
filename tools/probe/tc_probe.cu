// Stand-alone probe (not part of the product): one CTA computes D[128 m][N] = sum_k A[k][m] * B[k][n] with tcgen05.mma
// kind::tf32, both operands MN-major in shared memory in the layout a TMA box {32 floats, rows} with SWIZZLE_128B produces
// ([m/32 blocks][k rows][32 floats], 16-byte chunk c of row r stored at chunk c ^ (r & 7)), accumulator in TMEM.
// Prints max |D - reference| for integer-valued inputs (exact in tf32).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <string.h>
#include <math.h>

constexpr int M = 128, N = 64, K = 32;           // K = rows of one box (4 MMA k-steps of 8)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t ltype) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);                 // start address
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;       // leading byte offset
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                                 // version = 1 (Blackwell)
  d |= (uint64_t)ltype << 61;                             // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ Ag, const float* __restrict__ Bg, float* __restrict__ Dg, int lbo_a,
                                             int sbo_a, int lbo_b, int sbo_b, int kadv, int majors, int ltype) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* sA = (float*)base;                       // [M/32][K][32]
  float* sB = sA + M * K;                         // [N/32][K][32]
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // stage operands in the swizzled box layout
  for (int i = tid; i < M * K; i += 128) {
    const int k = i / M, m = i % M;
    const int blk = m >> 5, c = (m & 31) >> 2, w = m & 3;
    const int ia = ltype == 2 ? (((c ^ (k & 7)) << 2) | w) : ((((c >> 1) ^ (k & 3)) << 3) | ((c & 1) << 2) | w);
    sA[(blk * K + k) * 32 + ia] = Ag[k * M + m];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int k = i / N, n = i % N;
    const int blk = n >> 5, c = (n & 31) >> 2, w = n & 3;
    const int ib = ltype == 2 ? (((c ^ (k & 7)) << 2) | w) : ((((c >> 1) ^ (k & 3)) << 3) | ((c & 1) << 2) | w);
    sB[(blk * K + k) * 32 + ib] = Bg[k * N + n];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  {
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v = __float_as_uint(777.f);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(addr + c0), "r"(v) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  
  if (tid == 0) {
    // instruction descriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), a_major=MN (1<<15), b_major=MN (1<<16), N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(majors & 1) << 15) | ((uint32_t)((majors >> 1) & 1) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t da = make_desc_mn(smem_u32(sA) + ks * kadv, lbo_a, sbo_a, ltype);
      const uint64_t db = make_desc_mn(smem_u32(sB) + ks * kadv, lbo_b, sbo_b, ltype);
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // D: lane = m, column = n.  Each warp reads its 32 lanes, 32 columns at a time.
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
        "%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) Dg[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main(int argc, char** argv) {
  // candidate descriptor parameters: (lbo_a, sbo_a, lbo_b, sbo_b, k advance bytes)
  const int box = K * 128;       // bytes of one [K rows][32 floats] box
  int cand[][7] = {{box, 512, box, 512, 1024, 3, 1}, {512, box, 512, box, 1024, 3, 1}, {box, 1024, box, 1024, 1024, 3, 1}, {box, 512, box, 512, 512, 3, 1}};
  std::vector<float> A(K * M), B(K * N), D(M * N), R(M * N);
  srand(1);
  for (auto& v : A) v = (float)(rand() % 7 - 3);
  for (auto& v : B) v = (float)(rand() % 5 - 2);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[k * M + m] * B[k * N + n];
      R[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(M + N) * K * 4 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int nc = sizeof(cand) / sizeof(cand[0]);
  for (int c = 0; c < nc; ++c) {
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, cand[c][0], cand[c][1], cand[c][2], cand[c][3], cand[c][4], cand[c][5], cand[c][6]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cand %d: CUDA error %s\n", c, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs((double)D[i] - R[i]); if (d > mx) mx = d; if (d > 0.5) ++bad; }
    printf("majors %d cand %d (lboA %d sboA %d lboB %d sboB %d kadv %d): max err %.3f, wrong %d / %d; D[0][0..3] = %.1f %.1f %.1f %.1f (ref %.1f %.1f %.1f %.1f)\n", cand[c][5], c,
           cand[c][0], cand[c][1], cand[c][2], cand[c][3], cand[c][4], mx, bad, M * N, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
  }
  // ---- does kind::tf32 truncate or round its fp32 inputs? ----
  auto trunc13 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; };
  auto rna13 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; };
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  probe<<<1, 128, smem>>>(dA, dB, dD, cand[0][0], cand[0][1], cand[0][2], cand[0][3], cand[0][4], cand[0][5], cand[0][6]);
  cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double et = 0, er = 0, ef = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double st = 0, sr = 0, sf = 0;
      for (int k = 0; k < K; ++k) {
        st += (double)trunc13(A[k * M + m]) * trunc13(B[k * N + n]);
        sr += (double)rna13(A[k * M + m]) * rna13(B[k * N + n]);
        sf += (double)A[k * M + m] * B[k * N + n];
      }
      et = fmax(et, fabs(D[m * N + n] - st)); er = fmax(er, fabs(D[m * N + n] - sr)); ef = fmax(ef, fabs(D[m * N + n] - sf));
    }
  printf("random inputs: max |D - ref| with truncated inputs %.3e, rounded inputs %.3e, full fp32 inputs %.3e\n", et, er, ef);
  return 0;
}
