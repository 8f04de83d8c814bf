// Headless stand-in for <FL/gl.h>: every GL entry point used by the non-GUI reference
// files is a variadic no-op. Test infrastructure only.
#pragma once
typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
typedef double GLdouble;
typedef unsigned char GLubyte;
typedef unsigned char GLboolean;
typedef void GLvoid;
#define STG_SHIM_NOOP(name) template <class... A> inline void name(A...) {}
STG_SHIM_NOOP(glColor4f) STG_SHIM_NOOP(glColor3f) STG_SHIM_NOOP(glColor4dv)
STG_SHIM_NOOP(glPushMatrix) STG_SHIM_NOOP(glPopMatrix) STG_SHIM_NOOP(glScalef)
STG_SHIM_NOOP(glTranslatef) STG_SHIM_NOOP(glRotatef) STG_SHIM_NOOP(glPolygonMode)
STG_SHIM_NOOP(glRecti) STG_SHIM_NOOP(glRectf) STG_SHIM_NOOP(glEnableClientState)
STG_SHIM_NOOP(glDisableClientState) STG_SHIM_NOOP(glVertexPointer) STG_SHIM_NOOP(glColorPointer)
STG_SHIM_NOOP(glDrawArrays) STG_SHIM_NOOP(glEnable) STG_SHIM_NOOP(glDisable)
STG_SHIM_NOOP(glBegin) STG_SHIM_NOOP(glEnd) STG_SHIM_NOOP(glVertex2f) STG_SHIM_NOOP(glVertex3f)
STG_SHIM_NOOP(glVertex3dv) STG_SHIM_NOOP(glCallList) STG_SHIM_NOOP(glDepthMask)
STG_SHIM_NOOP(glEdgeFlag) STG_SHIM_NOOP(glEndList) STG_SHIM_NOOP(glNewList)
STG_SHIM_NOOP(glPolygonOffset) STG_SHIM_NOOP(glDeleteLists) STG_SHIM_NOOP(glPointSize)
STG_SHIM_NOOP(glBlendFunc) STG_SHIM_NOOP(glLineStipple) STG_SHIM_NOOP(glLineWidth)
inline GLuint glGenLists(int) { return 1; }
enum {
  GL_FALSE = 0, GL_TRUE = 1,
  GL_COLOR_ARRAY = 0x100, GL_DEPTH_TEST, GL_FLOAT, GL_FRONT_AND_BACK, GL_LINE, GL_QUADS,
  GL_VERTEX_ARRAY, GL_POLYGON, GL_QUAD_STRIP, GL_COMPILE, GL_FILL, GL_POLYGON_OFFSET_FILL,
  GL_FRONT, GL_POINTS, GL_BLEND, GL_LINES, GL_ONE_MINUS_SRC_ALPHA, GL_SRC_ALPHA,
  GL_LINE_STIPPLE, GL_LINE_STRIP
};
