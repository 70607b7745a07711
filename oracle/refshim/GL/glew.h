// stub: the hot path uses no GL entry points (oracle/refshim/README)
#pragma once
typedef unsigned int GLuint; typedef int GLint; typedef unsigned int GLenum; typedef float GLfloat; typedef int GLsizei;
