// k_tail (sc_tail_impl.cuh) is built twice: 4 batch rows per CTA (this file) and 2 (sc_tail_tr2.cu).  A stand-alone agent's tail is a
// chain of latency-bound stages on 64 CTAs; with 2 rows per CTA it runs on 128 SMs and every stage has half the work (measured:
// 183 -> 180 us per step).  When E agents share the GPU the machine is full either way and the doubled weight streaming of the
// 2-row variant costs 14 % of an ensemble step, so ensembles keep 4 rows -- and so does a stand-alone agent whose batch already
// fills the machine with 4-row CTAs (TD3 at batch 1024: 256 CTAs; with 2 rows its 512 CTAs ran in four waves instead of two,
// 319 -> 376 us per step, profiles/r2f_bench_n1.json).  Every row is computed by the same instructions in the same order in both
// variants (tests/test_gpu_round2.py::test_ensemble_agents_bit_identical_to_standalone compares them bit for bit).
#define SC_TAIL_TR 4
#define SC_TAIL_NS tail4
#include "sc_tail_impl.cuh"

namespace sc {
namespace tail2 { cudaError_t launch_tail_variant(const RowPhaseParams* pp, int E, int B, cudaStream_t st); }

cudaError_t launch_tail(const RowPhaseParams* pp, int E, int B, cudaStream_t st) {
  // 2 rows per CTA only while that grid still fits one wave of the 148 SMs (one CTA per SM: 215 KB of shared memory each)
  const bool two_rows = E == 1 && (B + 1) / 2 <= 148;
  return two_rows ? tail2::launch_tail_variant(pp, E, B, st) : tail4::launch_tail_variant(pp, E, B, st);
}
}  // namespace sc
