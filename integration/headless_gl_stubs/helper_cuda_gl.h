/* Stand-in for the CUDA-samples helper_cuda_gl.h: on a headless box the "GL device" is the CUDA device. */
#ifndef LYAP_HEADLESS_HELPER_CUDA_GL_H
#define LYAP_HEADLESS_HELPER_CUDA_GL_H
#include <helper_cuda.h>
static inline int findCudaGLDevice(int argc, const char **argv) { return findCudaDevice(argc, argv); }
#endif
