// Cluster "row-phase" kernel, build configuration occ1 (one CTA per SM) (see sc_rowphase.cu for the dispatcher and the description).
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
#include "sc_rowphase_host.cuh"

namespace sc {
int rowphase_capacity_occ1(int cs) {
  return cs == 16 ? max_clusters(occ1::k_rowphase<16, 4>, occ1::SMEM_BYTES, 16) : max_clusters(occ1::k_rowphase<8, 4>, occ1::SMEM_BYTES, 8);
}
size_t rowphase_smem_occ1() { return occ1::SMEM_BYTES; }
int rowphase_max_rows_occ1() { return occ1::MAXROWS; }
int rowphase_panel_rows_occ1() { return occ1::PROW; }
cudaError_t launch_rowphase_occ1(const RowPhaseParams* p, int E, int cs, int nclusters, int tm, cudaStream_t st) {
  if (cs == 16)
    return tm == 2 ? launch_k(occ1::k_rowphase<16, 2>, occ1::SMEM_BYTES, p, E, cs, nclusters, st)
         : tm == 3 ? launch_k(occ1::k_rowphase<16, 3>, occ1::SMEM_BYTES, p, E, cs, nclusters, st)
                   : launch_k(occ1::k_rowphase<16, 4>, occ1::SMEM_BYTES, p, E, cs, nclusters, st);
  return tm == 2 ? launch_k(occ1::k_rowphase<8, 2>, occ1::SMEM_BYTES, p, E, cs, nclusters, st)
       : tm == 3 ? launch_k(occ1::k_rowphase<8, 3>, occ1::SMEM_BYTES, p, E, cs, nclusters, st)
                 : launch_k(occ1::k_rowphase<8, 4>, occ1::SMEM_BYTES, p, E, cs, nclusters, st);
}
}  // namespace sc
