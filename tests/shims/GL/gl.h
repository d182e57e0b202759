// TEST INFRASTRUCTURE: OpenGL scalar typedefs, enough for `g++ -fsyntax-only` of SDR++ modules (tests/test_module_compile.py);
// libGL headers are not installed in this image and nothing here is ever linked.
#pragma once
typedef unsigned int GLenum; typedef unsigned int GLuint; typedef int GLint; typedef int GLsizei; typedef float GLfloat;
typedef unsigned char GLubyte; typedef unsigned char GLboolean; typedef void GLvoid; typedef unsigned int GLbitfield; typedef double GLdouble;
