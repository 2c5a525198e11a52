/* Headless stand-in for <GL/freeglut.h> + <GL/gl.h>, for building the reference's lyap_interactive.cu
 * on a box without GL/GLUT (integration/README.md).  No window, no context: the GL calls are no-ops, a
 * "pixel buffer object" is just a recorded size (cuda_gl_interop.h in this directory backs it with
 * cudaMalloc), and glutMainLoop() runs the display callback ONCE and returns -- which is all the
 * reference program ever does with it: its render() saves the frame and calls cleanup()
 * (lyap_interactive.cu:696-741).  Written for this repository; not reference or Khronos code. */
#ifndef LYAP_HEADLESS_FREEGLUT_H
#define LYAP_HEADLESS_FREEGLUT_H
#include <stddef.h>

typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef unsigned int GLbitfield;
typedef int GLint;
typedef int GLsizei;
typedef double GLdouble;
typedef ptrdiff_t GLsizeiptr;

enum {
    GL_MODELVIEW = 0x1700, GL_PROJECTION = 0x1701, GL_COLOR_BUFFER_BIT = 0x4000, GL_DEPTH_TEST = 0x0B71, GL_RGBA = 0x1908,
    GL_UNSIGNED_BYTE = 0x1401, GL_STREAM_DRAW = 0x88E0, GL_PIXEL_UNPACK_BUFFER_ARB = 0x88EC,
    GLUT_RGB = 0, GLUT_DOUBLE = 2, GLUT_UP = 1
};

enum { LYAP_HEADLESS_MAX_BUFFERS = 16 };
static size_t lyap_headless_buffer_bytes[LYAP_HEADLESS_MAX_BUFFERS];
static GLuint lyap_headless_next_buffer = 1, lyap_headless_bound_buffer = 0;
static void (*lyap_headless_display)(void) = 0;

static inline void glViewport(GLint, GLint, GLsizei, GLsizei) {}
static inline void glMatrixMode(GLenum) {}
static inline void glLoadIdentity(void) {}
static inline void glOrtho(GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble) {}
static inline void glClear(GLbitfield) {}
static inline void glDisable(GLenum) {}
static inline void glRasterPos2i(GLint, GLint) {}
static inline void glDrawPixels(GLsizei, GLsizei, GLenum, GLenum, const void *) {}
static inline void glGenBuffers(GLsizei n, GLuint *ids)
{
    for (GLsizei k = 0; k < n; ++k) ids[k] = lyap_headless_next_buffer < LYAP_HEADLESS_MAX_BUFFERS ? lyap_headless_next_buffer++ : 0;
}
static inline void glBindBuffer(GLenum, GLuint id) { lyap_headless_bound_buffer = id; }
static inline void glBufferData(GLenum, GLsizeiptr bytes, const void *, GLenum)
{
    if (lyap_headless_bound_buffer < LYAP_HEADLESS_MAX_BUFFERS) lyap_headless_buffer_bytes[lyap_headless_bound_buffer] = (size_t)bytes;
}
static inline void glDeleteBuffers(GLsizei, const GLuint *) {}

static inline void glutInit(int *, char **) {}
static inline void glutInitDisplayMode(unsigned int) {}
static inline void glutInitWindowSize(int, int) {}
static inline int glutCreateWindow(const char *) { return 1; }
static inline int glutGetWindow(void) { return 1; }
static inline void glutDestroyWindow(int) {}
static inline void glutDisplayFunc(void (*f)(void)) { lyap_headless_display = f; }
static inline void glutSpaceballRotateFunc(void (*)(int, int, int)) {}
static inline void glutSpaceballMotionFunc(void (*)(int, int, int)) {}
static inline void glutSpaceballButtonFunc(void (*)(int, int)) {}
static inline void glutKeyboardFunc(void (*)(unsigned char, int, int)) {}
static inline void glutReshapeFunc(void (*)(int, int)) {}
static inline void glutIdleFunc(void (*)(void)) {}
static inline void glutCloseFunc(void (*)(void)) {}
static inline void glutPostRedisplay(void) {}
static inline void glutSwapBuffers(void) {}
static inline void glutReportErrors(void) {}
static inline void glutMainLoop(void)
{
    if (lyap_headless_display) lyap_headless_display();
}
#endif
