/* Headless stand-in for <cuda_gl_interop.h>: a registered "GL buffer" is a cudaMalloc of the size the
 * fake glBufferData recorded (GL/freeglut.h in this directory).  The map / unmap / unregister calls of
 * the CUDA runtime are redirected for that one resource type.  Not NVIDIA code. */
#ifndef LYAP_HEADLESS_CUDA_GL_INTEROP_H
#define LYAP_HEADLESS_CUDA_GL_INTEROP_H
#include <cuda_runtime.h>
#include <stdlib.h>

#include <GL/freeglut.h>

struct lyap_headless_resource {
    void *dptr;
    size_t bytes;
};

static inline cudaError_t lyap_headless_register(struct cudaGraphicsResource **res, GLuint buffer, unsigned int)
{
    struct lyap_headless_resource *r = (struct lyap_headless_resource *)calloc(1, sizeof *r);
    r->bytes = buffer < LYAP_HEADLESS_MAX_BUFFERS ? lyap_headless_buffer_bytes[buffer] : 0;
    *res = (struct cudaGraphicsResource *)r;
    return r->bytes ? cudaMalloc(&r->dptr, r->bytes) : cudaErrorInvalidValue;
}
static inline cudaError_t lyap_headless_get_pointer(void **dptr, size_t *bytes, struct cudaGraphicsResource *res)
{
    struct lyap_headless_resource *r = (struct lyap_headless_resource *)res;
    *dptr = r->dptr;
    if (bytes) *bytes = r->bytes;
    return cudaSuccess;
}
static inline cudaError_t lyap_headless_unregister(struct cudaGraphicsResource *res)
{
    struct lyap_headless_resource *r = (struct lyap_headless_resource *)res;
    cudaError_t e = r && r->dptr ? cudaFree(r->dptr) : cudaSuccess;
    if (r) r->dptr = 0;
    return e;
}

#define cudaGraphicsGLRegisterBuffer(res, buf, flags) lyap_headless_register(res, buf, flags)
#define cudaGraphicsMapResources(n, res, stream) (cudaSuccess)
#define cudaGraphicsUnmapResources(n, res, stream) (cudaSuccess)
#define cudaGraphicsResourceGetMappedPointer(p, n, res) lyap_headless_get_pointer(p, n, res)
#define cudaGraphicsUnregisterResource(res) lyap_headless_unregister(res)
#endif
