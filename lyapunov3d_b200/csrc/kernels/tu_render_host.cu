// render kernels, exponent mode host: all sequence-period instantiations.
#define LYAP_TU_MODE kHost
#define LYAP_TU_NAME host
#include "tu_render_impl.cuh"
