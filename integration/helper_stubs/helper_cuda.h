/* Minimal stand-ins for the CUDA-samples helpers the reference's programs include
 * (checkCudaErrors, getLastCudaError, findCudaDevice: reference GNUmakefile:6 points at
 * $(CUDA)/samples/common/inc, which is not part of this image).  Written for the integration
 * build in tests/test_integration_patch.py; not reference code. */
#ifndef LYAP_HELPER_CUDA_STUB_H
#define LYAP_HELPER_CUDA_STUB_H
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define checkCudaErrors(call)                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(EXIT_FAILURE);                                                                \
        }                                                                                      \
    } while (0)

#define getLastCudaError(msg)                                                                  \
    do {                                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s: %s\n", msg, cudaGetErrorString(e_));                          \
            exit(EXIT_FAILURE);                                                                \
        }                                                                                      \
    } while (0)

static inline int findCudaDevice(int argc, const char **argv)
{
    int dev = 0;
    for (int i = 1; i < argc; ++i)
        if (!strncmp(argv[i], "-device=", 8)) dev = atoi(argv[i] + 8);
    checkCudaErrors(cudaSetDevice(dev));
    return dev;
}
#endif
