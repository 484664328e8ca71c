// Host-side launch helpers shared by the two build configurations of the row-phase kernel (sc_rowphase_occ{1,2}.cu).
#pragma once
#include <stdlib.h>
#include "sc_common.cuh"
#include "sc_params.h"

namespace sc {

template <typename K>
static cudaError_t configure(K kern, size_t smem, int cs) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && cs > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return e;
}
static void cluster_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int nclusters, int cs, size_t smem, cudaStream_t st, int E = 1) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(nclusters * cs, E, 1);
  cfg.blockDim = dim3(SC_ROW_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch (SMARTCAL_PDL_ROW=0 turns it off).  The kernel in front must release it LATE (k_dense_tc
  // stage 2 does, after its stores): released early, the clusters of this 256-CTA grid became resident piecemeal next to
  // the dense CTAs and the grid ran in two waves (+40 us, measured)
  static int pdl_row = -1;
  if (pdl_row < 0) { const char* e = getenv("SMARTCAL_PDL_ROW"); pdl_row = (e && e[0] == '0') ? 0 : 1; }
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_row ? 2 : 1;
}
template <typename K>
static cudaError_t launch_k(K kern, size_t smem, const RowPhaseParams* pp, int E, int cs, int nclusters, cudaStream_t st) {
  cudaError_t e = configure(kern, smem, cs);          // idempotent and cheap
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  cluster_cfg(cfg, attr, nclusters, cs, smem, st, E);
  return cudaLaunchKernelEx(&cfg, kern, pp);
}
template <typename K>
static int max_clusters(K kern, size_t smem, int cs) {
  if (configure(kern, smem, cs) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  cluster_cfg(cfg, attr, 64, cs, smem, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

}  // namespace sc
