/* Stand-in for the CUDA-samples helper_gl.h (extension queries): the headless "GL" supports whatever is asked. */
#ifndef LYAP_HEADLESS_HELPER_GL_H
#define LYAP_HEADLESS_HELPER_GL_H
static inline int isGLVersionSupported(unsigned, unsigned) { return 1; }
static inline int areGLExtensionsSupported(const char *) { return 1; }
#endif
