// hybrid-mode march kernels with the host parity evaluator: all sequence-period instantiations.
#define LYAP_TU_MODE kHost
#define LYAP_TU_NAME host
#include "tu_march_impl.cuh"
