// the 2-rows-per-CTA build of k_tail (see sc_tail.cu)
#define SC_TAIL_TR 2
#define SC_TAIL_NS tail2
#include "sc_tail_impl.cuh"
