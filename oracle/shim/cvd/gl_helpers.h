// TEST INFRASTRUCTURE — stand-in for OpenGL + cvd/gl_helpers.h: every drawing call is a no-op (the
// reference's drawing is out of scope and is never reached: TrackFrame is called with bDraw = false).
#pragma once
#include <TooN/TooN.h>
#include "image.h"
enum { GL_POINTS, GL_LINES, GL_LINE_STRIP, GL_LINE_SMOOTH, GL_POINT_SMOOTH, GL_BLEND, GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA };
template <class... A> inline void glColor3f(A...) {}
template <class... A> inline void glColor4f(A...) {}
template <class... A> inline void glPointSize(A...) {}
template <class... A> inline void glLineWidth(A...) {}
template <class... A> inline void glBegin(A...) {}
inline void glEnd() {}
template <class... A> inline void glEnable(A...) {}
template <class... A> inline void glDisable(A...) {}
template <class... A> inline void glBlendFunc(A...) {}
namespace CVD {
template <class... A> inline void glVertex(const A&...) {}
template <class... A> inline void glColor(const A&...) {}
template <class... A> inline void glDrawPixels(const A&...) {}
}
