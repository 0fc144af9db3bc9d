"""Inert stand-in for PyOpenGL (the reference constructor allocates GL buffers, object.py:82-98)."""
import ctypes  # noqa: F401  (object.py uses ctypes.c_void_p through this star-import)
GL_ARRAY_BUFFER = GL_ELEMENT_ARRAY_BUFFER = GL_DYNAMIC_DRAW = GL_STATIC_DRAW = GL_FLOAT = GL_FALSE = GL_TRIANGLES = GL_UNSIGNED_INT = 0


def glGenVertexArrays(n): return 1
def glGenBuffers(n): return 1
def glBindVertexArray(*a): pass
def glBindBuffer(*a): pass
def glBufferData(*a): pass
def glVertexAttribPointer(*a): pass
def glEnableVertexAttribArray(*a): pass
def glDrawElements(*a): pass
