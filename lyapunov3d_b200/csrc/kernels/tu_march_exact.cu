// hybrid-mode march kernels with the exact parity evaluator: all sequence-period instantiations.
#define LYAP_TU_MODE kExact
#define LYAP_TU_NAME exact
#include "tu_march_impl.cuh"
