#pragma once
#include <FL/gl.h>
struct GLUquadric {};
struct GLUtesselator {};
enum { GLU_TESS_BEGIN, GLU_TESS_COMBINE, GLU_TESS_EDGE_FLAG, GLU_TESS_END, GLU_TESS_ERROR,
       GLU_TESS_VERTEX, GLU_FILL, GLU_LINE, GLU_SILHOUETTE };
inline GLUquadric *gluNewQuadric() { static GLUquadric q; return &q; }
inline GLUtesselator *gluNewTess() { static GLUtesselator t; return &t; }
inline const GLubyte *gluErrorString(GLenum) { return (const GLubyte *)""; }
STG_SHIM_NOOP(gluDeleteQuadric) STG_SHIM_NOOP(gluQuadricDrawStyle) STG_SHIM_NOOP(gluSphere)
STG_SHIM_NOOP(gluPartialDisk) STG_SHIM_NOOP(gluTessCallback) STG_SHIM_NOOP(gluTessBeginPolygon)
STG_SHIM_NOOP(gluTessBeginContour) STG_SHIM_NOOP(gluTessVertex) STG_SHIM_NOOP(gluTessEndContour)
STG_SHIM_NOOP(gluTessEndPolygon)
