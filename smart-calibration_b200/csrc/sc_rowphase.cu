// Cluster "row-phase" kernel, host side: two build configurations of sc_rowphase_impl.cuh and the launcher.
//
//   occ1: one CTA per SM  (213 KB shared memory: 3 x 48 KB operand slots of 128 reduction elements, <=32-row tiles, tail RT=4)
//   occ2: two CTAs per SM (99 KB: 3 x 20 KB slots of 64 reduction elements, <=16-row tiles, tail RT=2, <=112 registers)
// occ2 doubles the number of resident clusters, so a 256-row batch is cut into 16-row slabs (no padded rows) and two
// independent instruction streams share every SM, which hides the latencies this kernel is bound by.
#include "sc_common.cuh"
#include "sc_params.h"
#include "sc_tile.cuh"

#define SC_CFG_NS occ1
#define SC_CFG_KC 128
#define SC_CFG_NSTAGE 3
#define SC_CFG_MAXROWS 128
#define SC_CFG_RT 4
#define SC_CFG_REDROWS 32
#define SC_CFG_ZK 8192
#define SC_CFG_MINBLOCKS 1
#include "sc_rowphase_impl.cuh"
#undef SC_CFG_NS
#undef SC_CFG_KC
#undef SC_CFG_NSTAGE
#undef SC_CFG_MAXROWS
#undef SC_CFG_RT
#undef SC_CFG_REDROWS
#undef SC_CFG_ZK
#undef SC_CFG_MINBLOCKS

#define SC_CFG_NS occ2
#define SC_CFG_KC 64
#define SC_CFG_NSTAGE 3
#define SC_CFG_MAXROWS 64
#define SC_CFG_RT 2
#define SC_CFG_REDROWS 16
#define SC_CFG_ZK 5120
#define SC_CFG_MINBLOCKS 2
#include "sc_rowphase_impl.cuh"
#undef SC_CFG_NS
#undef SC_CFG_KC
#undef SC_CFG_NSTAGE
#undef SC_CFG_MAXROWS
#undef SC_CFG_RT
#undef SC_CFG_REDROWS
#undef SC_CFG_ZK
#undef SC_CFG_MINBLOCKS

namespace sc {

template <typename K>
static cudaError_t configure(K kern, size_t smem, int cs) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess && cs > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  return e;
}
static void cluster_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int nclusters, int cs, size_t smem, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(nclusters * cs, 1, 1);
  cfg.blockDim = dim3(SC_ROW_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}
template <typename K>
static cudaError_t launch_k(K kern, size_t smem, const RowPhaseParams& p, int cs, int nclusters, cudaStream_t st) {
  cudaError_t e = configure(kern, smem, cs);          // idempotent and cheap
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_cfg(cfg, attr, nclusters, cs, smem, st);
  return cudaLaunchKernelEx(&cfg, kern, p);
}
template <typename K>
static int max_clusters(K kern, size_t smem, int cs) {
  if (configure(kern, smem, cs) != cudaSuccess) { cudaGetLastError(); return 0; }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_cfg(cfg, attr, 64, cs, smem, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// co-resident clusters per (variant, cluster size): out[0] occ1/16, out[1] occ1/8, out[2] occ2/16, out[3] occ2/8
void rowphase_cluster_capacity(int* out) {
  out[0] = max_clusters(occ1::k_rowphase<16, 4>, occ1::SMEM_BYTES, 16);
  out[1] = max_clusters(occ1::k_rowphase<8, 4>, occ1::SMEM_BYTES, 8);
  out[2] = max_clusters(occ2::k_rowphase<16, 2>, occ2::SMEM_BYTES, 16);
  out[3] = max_clusters(occ2::k_rowphase<8, 2>, occ2::SMEM_BYTES, 8);
}
size_t rowphase_smem_bytes(int variant) { return variant == 2 ? occ2::SMEM_BYTES : occ1::SMEM_BYTES; }
int rowphase_max_rows(int variant) { return variant == 2 ? occ2::MAXROWS : occ1::MAXROWS; }       // slab height limit
int rowphase_panel_rows(int variant) { return variant == 2 ? occ2::PROW : occ1::PROW; }           // dgrad weight panel box rows

cudaError_t launch_rowphase(const RowPhaseParams& p, int variant, int cs, int nclusters, int tm, cudaStream_t st) {
  if (variant == 2) {
    if (cs == 16) return tm == 1 ? launch_k(occ2::k_rowphase<16, 1>, occ2::SMEM_BYTES, p, cs, nclusters, st)
                                 : launch_k(occ2::k_rowphase<16, 2>, occ2::SMEM_BYTES, p, cs, nclusters, st);
    return tm == 1 ? launch_k(occ2::k_rowphase<8, 1>, occ2::SMEM_BYTES, p, cs, nclusters, st)
                   : launch_k(occ2::k_rowphase<8, 2>, occ2::SMEM_BYTES, p, cs, nclusters, st);
  }
  if (cs == 16)
    return tm == 2 ? launch_k(occ1::k_rowphase<16, 2>, occ1::SMEM_BYTES, p, cs, nclusters, st)
         : tm == 3 ? launch_k(occ1::k_rowphase<16, 3>, occ1::SMEM_BYTES, p, cs, nclusters, st)
                   : launch_k(occ1::k_rowphase<16, 4>, occ1::SMEM_BYTES, p, cs, nclusters, st);
  return tm == 2 ? launch_k(occ1::k_rowphase<8, 2>, occ1::SMEM_BYTES, p, cs, nclusters, st)
       : tm == 3 ? launch_k(occ1::k_rowphase<8, 3>, occ1::SMEM_BYTES, p, cs, nclusters, st)
                 : launch_k(occ1::k_rowphase<8, 4>, occ1::SMEM_BYTES, p, cs, nclusters, st);
}

}  // namespace sc
