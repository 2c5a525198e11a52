// bake kernels, exponent mode fast: all sequence-period instantiations.
#define LYAP_TU_MODE kFast
#define LYAP_TU_NAME fast
#include "tu_bake_impl.cuh"
