"""Inert stand-in for PyGLM (mouse-drag vector math in the dynamic classes; unused when nothing is selected)."""
import numpy as np


def vec3(*a):
    if len(a) == 1:
        return np.array(a[0], dtype=np.float32).reshape(3)
    return np.array(a if a else (0.0, 0.0, 0.0), dtype=np.float32)


def length(v):
    return float(np.linalg.norm(v))


def dot(a, b):
    return float(np.dot(a, b))
