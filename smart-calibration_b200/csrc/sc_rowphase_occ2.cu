// Cluster "row-phase" kernel, build configuration occ2 (two CTAs per SM) (see sc_rowphase.cu for the dispatcher and the description).
#include "sc_common.cuh"
#include "sc_params.h"
#include "sc_tile.cuh"

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
#include "sc_rowphase_host.cuh"

namespace sc {
int rowphase_capacity_occ2(int cs) {
  return cs == 16 ? max_clusters(occ2::k_rowphase<16, 2>, occ2::SMEM_BYTES, 16) : max_clusters(occ2::k_rowphase<8, 2>, occ2::SMEM_BYTES, 8);
}
size_t rowphase_smem_occ2() { return occ2::SMEM_BYTES; }
int rowphase_max_rows_occ2() { return occ2::MAXROWS; }
int rowphase_panel_rows_occ2() { return occ2::PROW; }
cudaError_t launch_rowphase_occ2(const RowPhaseParams* p, int E, int cs, int nclusters, int tm, cudaStream_t st) {
  if (cs == 16) return tm == 1 ? launch_k(occ2::k_rowphase<16, 1>, occ2::SMEM_BYTES, p, E, cs, nclusters, st)
                               : launch_k(occ2::k_rowphase<16, 2>, occ2::SMEM_BYTES, p, E, cs, nclusters, st);
  return tm == 1 ? launch_k(occ2::k_rowphase<8, 1>, occ2::SMEM_BYTES, p, E, cs, nclusters, st)
                 : launch_k(occ2::k_rowphase<8, 2>, occ2::SMEM_BYTES, p, E, cs, nclusters, st);
}
}  // namespace sc
