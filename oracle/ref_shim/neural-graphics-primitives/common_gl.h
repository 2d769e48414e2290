// Shadows the reference's common_gl.h (OpenGL loader): TEST INFRASTRUCTURE ONLY (oracle/_ref), no GL on this path.
#pragma once
typedef unsigned int GLuint; typedef int GLint; typedef unsigned int GLenum; typedef float GLfloat;
