// bake kernels, exponent mode exact: all sequence-period instantiations.
#define LYAP_TU_MODE kExact
#define LYAP_TU_NAME exact
#include "tu_bake_impl.cuh"
