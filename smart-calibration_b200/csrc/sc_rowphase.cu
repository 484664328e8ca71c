// Cluster "row-phase" kernel, host side: dispatcher over the two build configurations of sc_rowphase_impl.cuh
// (compiled as separate translation units, sc_rowphase_occ1.cu / sc_rowphase_occ2.cu, so that they build in parallel).
//
//   occ1: one CTA per SM  (213 KB shared memory: 3 x 48 KB operand slots of 128 reduction elements, <=32-row tiles, tail RT=4)
//   occ2: two CTAs per SM (99 KB: 3 x 20 KB slots of 64 reduction elements, <=16-row tiles, tail RT=2, <=112 registers)
// occ2 doubles the number of resident clusters, so a 256-row batch is cut into 16-row slabs (no padded rows) and two
// independent instruction streams share every SM, which hides the latencies this kernel is bound by.
#include <cuda_runtime.h>

#include "sc_params.h"

namespace sc {
int rowphase_capacity_occ1(int cs);
int rowphase_capacity_occ2(int cs);
size_t rowphase_smem_occ1();
size_t rowphase_smem_occ2();
int rowphase_max_rows_occ1();
int rowphase_max_rows_occ2();
int rowphase_panel_rows_occ1();
int rowphase_panel_rows_occ2();
cudaError_t launch_rowphase_occ1(const RowPhaseParams* p, int E, int cs, int nclusters, int tm, cudaStream_t st);
cudaError_t launch_rowphase_occ2(const RowPhaseParams* p, int E, int cs, int nclusters, int tm, cudaStream_t st);

// co-resident clusters per (variant, cluster size): out[0] occ1/16, out[1] occ1/8, out[2] occ2/16, out[3] occ2/8
void rowphase_cluster_capacity(int* out) {
  out[0] = rowphase_capacity_occ1(16);
  out[1] = rowphase_capacity_occ1(8);
  out[2] = rowphase_capacity_occ2(16);
  out[3] = rowphase_capacity_occ2(8);
}
size_t rowphase_smem_bytes(int variant) { return variant == 2 ? rowphase_smem_occ2() : rowphase_smem_occ1(); }
int rowphase_max_rows(int variant) { return variant == 2 ? rowphase_max_rows_occ2() : rowphase_max_rows_occ1(); }       // slab height limit
int rowphase_panel_rows(int variant) { return variant == 2 ? rowphase_panel_rows_occ2() : rowphase_panel_rows_occ1(); } // dgrad weight panel box rows

cudaError_t launch_rowphase(const RowPhaseParams* p, int E, int variant, int cs, int nclusters, int tm, cudaStream_t st) {
  return variant == 2 ? launch_rowphase_occ2(p, E, cs, nclusters, tm, st) : launch_rowphase_occ1(p, E, cs, nclusters, tm, st);
}

}  // namespace sc
