// Stride-2 5x5 convolutions of the calibration / demixing trunks (calib_sac.py:101-105, demix_sac.py:372-390) as implicit GEMMs on
// the 5th-generation tensor cores.  Forward:
//     z[b][co][oy][ox] = bias[co] + sum_k  A[(b, oy, ox)][k] W[co][k],   k = (ci, ky, kx),  A[(b, oy, ox)][k] = x[b][ci][2 oy + ky][2 ox + kx]
// is a GEMM with M = B * HOUT^2 output pixels (128 per tile, one accumulator lane per pixel), N = 32 output channels and
// K = 25 CIN.  torch's weight layout [Cout][Cin][5][5] already is the K-major B operand; it arrives by TMA, split into tf32
// big / small parts by k_conv_w_split.  The A operand never exists in memory: 16 gather warps read the activations of the previous
// layer straight from L1 / L2 (lanes = 32 neighbouring output pixels, i.e. 8-byte strides along an image row), split each value
// into its tf32 big / small parts and write them into the 128-byte-swizzled K-major layout the MMA descriptors expect.
// kind::tf32 keeps 10 mantissa bits, so fp32 fidelity is restored with the big / small products of the scheme of
// sc_dense_tc.cu (A_lo W + A W_lo + A W).
//
// Warp roles (704 threads): warp 0 = TMA producer of the weight chunks, warp 1 = TMEM allocation + MMA issue (one thread),
// warps 2..17 = im2col gather (four groups of four warps, group s fills stage slot s: thread = one pixel row x the 32 k of a chunk),
// warps 18..21 = epilogue: they read the finished accumulator (TMEM lane = pixel, two
// accumulator sets so that the next tile accumulates meanwhile) and store it channel by channel, which keeps every 128-byte
// segment of z written by one warp.  A CTA is persistent over tiles; four 40 KB stages [A big | W big | A small | W small].
#pragma once
#include <cuda.h>
#include "sc_common.cuh"

namespace sc {
namespace calib {

constexpr int CT_M = 128, CT_N = 32, CT_STAGES = 4;
constexpr int CT_A_BYTES = CT_M * 128, CT_W_BYTES = CT_N * 128;     // one chunk of 32 k
constexpr int CT_HALF_BYTES = CT_A_BYTES + CT_W_BYTES;
constexpr int CT_STAGE_BYTES = 2 * CT_HALF_BYTES;                   // 40 KB
constexpr int CT_GWARPS = 16, CT_EWARPS = 4;
constexpr int CT_THREADS = 64 + 32 * (CT_GWARPS + CT_EWARPS);
constexpr int CT_TMEM_COLS = 4 * CT_N;                              // two accumulator sets of (main | correction) x 32 columns

// k -> offset of x[ci][ky][kx] from a pixel's window origin, per layer shape; the same for every tile, so the gather loop carries
// no index arithmetic.  Constant memory: the index is warp-uniform, and a uniform-address LDS.128 measured ~11 wavefronts of the
// shared-memory data pipe this kernel is bound by.  Padded k (>= K) point at the origin: the weight operand is zero there.
constexpr int CT_KOFF_MAX = 32 * 25;                                // conv3: K = 800
__constant__ int c_koff[2][CT_KOFF_MAX];                            // [0]: CIN = 16 (62 -> 29), [1]: CIN = 32 (29 -> 13)
inline cudaError_t conv_tc_upload_tables() {
  static int h[2][CT_KOFF_MAX];
  const int cin[2] = {16, 32}, hin[2] = {62, 29};
  for (int t = 0; t < 2; ++t)
    for (int k = 0; k < CT_KOFF_MAX; ++k) {
      const int ci = k / 25, rem = k - 25 * ci, ky = rem / 5, kx = rem - 5 * ky;
      h[t][k] = k < cin[t] * 25 ? ci * hin[t] * hin[t] + ky * hin[t] + kx : 0;
    }
  return cudaMemcpyToSymbol(c_koff, h, sizeof(h));
}

struct ConvTcP {
  const float* x; size_t x_img_stride;        // input activations [B][CIN][HIN][HIN] (image stride in floats)
  const float* bias;
  float* z;                                   // [B][32][HOUT][HOUT]
  int B;
  unsigned long long* ts;                     // optional debug timeline of CTA 0 ($SMARTCAL_STAMPS): [role][1 + 255] (globaltimer << 8 | tag)
  int dbg;                                    // profiling experiments ($SMARTCAL_CONV_DBG): 1 = no activation loads, 2 = no MMAs, 4 = no proxy fence
};

__device__ __forceinline__ uint64_t ct_desc(uint32_t saddr) {       // K-major, SWIZZLE_128B, 8-row atoms of 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void ct_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// explicit shared-state-space stores: with generic stores ptxas puts a MEMBAR.ALL.CTA in front of the proxy fence, which also
// waits for the prefetched activation loads of the next chunk
__device__ __forceinline__ void ct_sts4(uint32_t saddr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void ct_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// tf32 big / small parts of the conv weights of one network: hl = [big | small], `n` floats each
__global__ void __launch_bounds__(256) k_conv_w_split(const float* __restrict__ w2, float* __restrict__ hl2, int n2, const float* __restrict__ w3,
                                                       float* __restrict__ hl3, int n3) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n2 + n3) / 4; i += gridDim.x * blockDim.x) {
    const bool second = i >= n2 / 4;
    const int j = second ? i - n2 / 4 : i;
    const float4 v = __ldg(reinterpret_cast<const float4*>(second ? w3 : w2) + j);
    float4 big, small;
    split_rn4(v, big, small);
    float* hl = second ? hl3 : hl2;
    const int n = second ? n3 : n2;
    reinterpret_cast<float4*>(hl)[j] = big;
    reinterpret_cast<float4*>(hl + n)[j] = small;
  }
}

template <int CIN, int HIN, int HOUT>
__global__ void __launch_bounds__(CT_THREADS, 1) k_conv_fwd_tc(const ConvTcP p, const __grid_constant__ CUtensorMap map_w_hi,
                                                                const __grid_constant__ CUtensorMap map_w_lo) {
  constexpr int K = CIN * 25, NK = (K + 31) / 32, OHW = HOUT * HOUT, HW_IN = HIN * HIN;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[CT_STAGES], empty[CT_STAGES], acc_done[2], acc_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float sbias[CT_N];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Mtot = p.B * OHW, ntiles = (Mtot + CT_M - 1) / CT_M;
  static_assert(NK * 32 <= CT_KOFF_MAX && (CIN == 16 || CIN == 32), "offset tables are built for the two trunk shapes");
  if (tid == 0) {
    for (int i = 0; i < CT_STAGES; ++i) {
      mbar_init(full + i, CT_GWARPS / CT_STAGES + 1);     // one arrival per gather warp of the slot's group + the producer's expect_tx
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_done + i, 1);
      mbar_init(acc_free + i, CT_EWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (tid < CT_N) sbias[tid] = __ldg(p.bias + tid);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(CT_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  // debug timeline: role 0 = producer lane 0, 1 = MMA leader, 2 = gather warp 0 lane 0, 3 = epilogue warp 0 lane 0
  unsigned long long tcount = 0;
  auto mark = [&](int role, int tag) {
    if (p.ts != nullptr && blockIdx.x == 0 && tcount < 255) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.ts[256 * role + 1 + tcount] = (t << 8) | (unsigned long long)tag;
      p.ts[256 * role] = ++tcount;
    }
  };

  if (warp == 0) {
    // ---------------- weight chunks by TMA ----------------
    unsigned g = 0;
    for (int it = 0; it < my_tiles; ++it)
      for (int kc = 0; kc < NK; ++kc, ++g) {
        const unsigned slot = g % CT_STAGES;
        if (g >= CT_STAGES) mbar_wait_sleep(empty + slot, ((g / CT_STAGES) - 1) & 1u, 100);
        if (lane == 0) {
          mark(0, 1);
          uint8_t* st = smem + (size_t)slot * CT_STAGE_BYTES;
          mbar_expect_tx(full + slot, 2 * CT_W_BYTES);
          tma_load_2d(reinterpret_cast<float*>(st + CT_A_BYTES), &map_w_hi, 32 * kc, 0, full + slot);
          tma_load_2d(reinterpret_cast<float*>(st + CT_HALF_BYTES + CT_A_BYTES), &map_w_lo, 32 * kc, 0, full + slot);
        }
        __syncwarp();
      }
  } else if (warp == 1) {
    // ---------------- MMA issue (one thread) ----------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(CT_N >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(leader));
    const uint64_t d_a0 = ct_desc(smem_u32(smem));
    const uint64_t d_w0 = d_a0 + (uint64_t)(CT_A_BYTES >> 4);
    const uint64_t d_al0 = d_a0 + (uint64_t)(CT_HALF_BYTES >> 4), d_wl0 = d_w0 + (uint64_t)(CT_HALF_BYTES >> 4);
    unsigned g = 0;
    for (int it = 0; it < my_tiles; ++it) {
      // two accumulator sets (tile parity): the epilogue warps drain one while the next tile accumulates into the other
      const uint32_t acc = tmem + (uint32_t)(it & 1) * (2 * CT_N);
      if (it >= 2) mbar_wait(acc_free + (it & 1), (unsigned)((it >> 1) - 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kc = 0; kc < NK; ++kc, ++g) {
        const unsigned slot = g % CT_STAGES;
        mbar_wait(full + slot, (g / CT_STAGES) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) {
          mark(1, 2);
          const uint64_t so = (uint64_t)((slot * CT_STAGE_BYTES) >> 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (p.dbg & 2) break;
            const uint64_t o = so + (uint64_t)(2 * j);            // + 32 B per 8-deep k-step
            // the big x big products in columns 0..31, the two correction products in columns 32..63.  The tensor core's FP32
            // accumulation is not round-to-nearest; keeping the 2^-11-sized terms out of the main accumulator quarters the number
            // of roundings at full magnitude (the BatchNorm backward of layer 1 amplifies forward differences)
            ct_mma(acc + CT_N, d_al0 + o, d_w0 + o, idesc, (kc | j) != 0 ? 1u : 0u);
            ct_mma(acc + CT_N, d_a0 + o, d_wl0 + o, idesc, 1u);
            ct_mma(acc, d_a0 + o, d_w0 + o, idesc, (kc | j) != 0 ? 1u : 0u);
          }
          mark(1, 3);
          ct_commit(empty + slot);
          if (kc == NK - 1) ct_commit(acc_done + (it & 1));
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + CT_GWARPS) {
    // ---------------- im2col gather ----------------
    // Four groups of four warps; group s fills stage slot s, i.e. every fourth chunk, a whole 32-float row per thread.  (The
    // proxy fence that publishes a stage to the tensor core compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and therefore waits
    // for every load the thread has in flight: prefetching the next chunk across it gains nothing.  Instead each thread keeps
    // 32 loads in flight per fence and the four groups overlap each other's load latency.)
    const int gw = warp - 2;
    const int r = (gw & 3) * 32 + lane;                      // pixel row of the tile this thread fills
    const int group = gw >> 2;                               // stage slot this warp group owns
    const unsigned total = (unsigned)my_tiles * NK;
    const uint32_t st = smem_u32(smem) + (uint32_t)group * CT_STAGE_BYTES + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    const float* xb = p.x;
    auto row_base = [&](int it) {
      // rows past the last pixel read pixel 0's window: their accumulator lanes are never stored
      const int m = ((int)blockIdx.x + it * (int)gridDim.x) * CT_M + r;
      const bool valid = m < Mtot;
      const int b = valid ? m / OHW : 0;
      const int pp = valid ? m - b * OHW : 0;
      const int oy = pp / HOUT, ox = pp - oy * HOUT;
      xb = p.x + (size_t)b * p.x_img_stride + (size_t)(2 * oy) * HIN + 2 * ox;
    };
    int it = 0, kc = group, it_base = -1;
    unsigned n = 0;                                          // chunks this group has filled
    for (unsigned g = group; g < total; g += CT_STAGES, kc += CT_STAGES, ++n) {
      while (kc >= NK) { kc -= NK; ++it; }
      if (it != it_base) { row_base(it); it_base = it; }
      if (gw == 0 && lane == 0) mark(2, 4);
      float v[32];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int* kt = c_koff[CIN == 16 ? 0 : 1] + 32 * kc + 4 * u;
        v[4 * u] = __ldg(xb + kt[0]); v[4 * u + 1] = __ldg(xb + kt[1]); v[4 * u + 2] = __ldg(xb + kt[2]); v[4 * u + 3] = __ldg(xb + kt[3]);
      }
      if (n > 0) mbar_wait_sleep(empty + group, (n - 1) & 1u, 40);
      if (gw == 0 && lane == 0) mark(2, 6);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float4 big, small;
        split_rn4(make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]), big, small);
        const uint32_t a = st + (((uint32_t)u ^ sw) << 4);    // swizzled 16-byte unit
        ct_sts4(a, big);
        ct_sts4(a + CT_HALF_BYTES, small);
      }
      if (!(p.dbg & 4)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic stores -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(full + group);
      if (gw == 0 && lane == 0) mark(2, 7);
    }
  } else {
    // ---------------- epilogue warps: TMEM lane = pixel; warp w may read lanes 32 (w & 3) .. + 31 ----------------
    const int q = warp & 3;
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      mbar_wait_sleep(acc_done + (it & 1), (unsigned)(it >> 1) & 1u, 400);
      if (warp == 2 + CT_GWARPS && lane == 0) mark(3, 8);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t a[32], c[32];
      const uint32_t addr = tmem + (uint32_t)(it & 1) * (2 * CT_N) + ((uint32_t)(q * 32) << 16);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
          "%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]), "=r"(c[8]), "=r"(c[9]), "=r"(c[10]),
            "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15]), "=r"(c[16]), "=r"(c[17]), "=r"(c[18]), "=r"(c[19]), "=r"(c[20]),
            "=r"(c[21]), "=r"(c[22]), "=r"(c[23]), "=r"(c[24]), "=r"(c[25]), "=r"(c[26]), "=r"(c[27]), "=r"(c[28]), "=r"(c[29]), "=r"(c[30]),
            "=r"(c[31])
          : "r"(addr + (uint32_t)CT_N));
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
          "%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]), "=r"(a[10]),
            "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]), "=r"(a[20]),
            "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]), "=r"(a[30]),
            "=r"(a[31])
          : "r"(addr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free + (it & 1));        // the accumulator set may be overwritten (tile it + 2)
      const int me = tile * CT_M + q * 32 + lane;
      if (me < Mtot) {
        const int be = me / OHW, pe = me - be * OHW;
        float* zp = p.z + (size_t)be * CT_N * OHW + pe;
#pragma unroll
        for (int ch = 0; ch < CT_N; ++ch) zp[(size_t)ch * OHW] = (__uint_as_float(a[ch]) + __uint_as_float(c[ch])) + sbias[ch];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(CT_TMEM_COLS));
}

inline size_t conv_tc_smem_bytes() { return (size_t)CT_STAGES * CT_STAGE_BYTES + 1024; }

}  // namespace calib
}  // namespace sc
