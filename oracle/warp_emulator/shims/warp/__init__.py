"""Stand-in for NVIDIA Warp: executes @wp.kernel bodies as plain Python on numpy arrays (ORACLE infrastructure only;
see oracle/warp_emulator/README.md).  Only the API surface the reference's hot path touches is provided."""
import ast
import inspect
import textwrap
import types as _pytypes

import numpy as np
import torch

from . import types, utils  # noqa: F401

float32 = np.float32
float64 = np.float64
int32 = np.int32


class _TypeTag:
    def __init__(self, name, tail, scalar):
        self.name, self.tail, self.scalar = name, tail, scalar

    def __call__(self, *a):
        if self.tail == (3,):
            v = np.zeros(3, self.scalar) if not a else np.array(a[0] if len(a) == 1 else a, dtype=self.scalar).reshape(3)
            return v.view(vec)
        if self.tail == (3, 3):
            m = np.zeros((3, 3), self.scalar) if not a else np.array(a, dtype=self.scalar).reshape(3, 3)
            return m.view(mat)
        return self.scalar(a[0] if a else 0)


class vec(np.ndarray):
    pass


class mat(np.ndarray):
    def __matmul__(self, o):
        return np.matmul(np.asarray(self), np.asarray(o)).view(mat if np.ndim(o) == 2 else vec)


vec3 = vec3f = _TypeTag("vec3f", (3,), np.float32)
vec3i = _TypeTag("vec3i", (3,), np.int32)
vec2i = _TypeTag("vec2i", (2,), np.int32)
mat33f = mat33 = _TypeTag("mat33f", (3, 3), np.float32)


def _tail(dtype):
    return dtype.tail if isinstance(dtype, _TypeTag) else ()


def _scalar(dtype):
    if isinstance(dtype, _TypeTag):
        return dtype.scalar
    return {float: np.float32, int: np.int32, None: np.float32}.get(dtype, dtype)


class _Annot:
    def __init__(self, **kw):
        self.kw = kw


class array:
    """numpy-backed wp.array.  ``data`` has the element tail dims appended (vec3 -> (...,3), mat33 -> (...,3,3))."""

    def __new__(cls, data=None, dtype=None, shape=None, ptr=None, length=None, device=None, requires_grad=False, ndim=None,
                copy=True, **kw):
        if data is None and shape is None and ptr is None and length is None:
            return _Annot(dtype=dtype, ndim=ndim)
        return object.__new__(cls)

    def __init__(self, data=None, dtype=None, shape=None, ptr=None, length=None, device=None, requires_grad=False, ndim=None,
                 copy=True, **kw):
        self.dtype = dtype if dtype is not None else float32
        tail = _tail(self.dtype)
        if data is not None:
            d = np.array(data, dtype=_scalar(self.dtype))
            self.data = d if d.shape[d.ndim - len(tail):] == tail or not tail else d.reshape((-1,) + tail)
        else:
            if shape is None:
                shape = (int(length),)
            if isinstance(shape, (int, np.integer)):
                shape = (int(shape),)
            self.data = np.zeros(tuple(int(s) for s in shape) + tail, dtype=_scalar(self.dtype))
        self.device = "cpu-emulated"

    @classmethod
    def wrap(cls, data, dtype):
        a = object.__new__(cls)
        a.data, a.dtype, a.device = data, dtype, "cpu-emulated"
        return a

    @property
    def shape(self):
        t = len(_tail(self.dtype))
        return self.data.shape[:self.data.ndim - t] if t else self.data.shape

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def ptr(self):
        return self.data.ctypes.data

    def numpy(self):
        return self.data

    def zero_(self):
        self.data[...] = 0

    def _elem(self, v):
        t = _tail(self.dtype)
        if v.ndim == len(t):
            if t == (3, 3):
                return v.view(mat)
            if t:
                return v.view(vec)
            return v[()]
        return array.wrap(v, self.dtype)

    def __getitem__(self, idx):
        return self._elem(self.data[idx])

    def __setitem__(self, idx, value):
        self.data[idx] = np.asarray(value)

    def __len__(self):
        return self.shape[0]


def zeros(shape=None, dtype=float32, device=None, requires_grad=False, **kw):
    return array(shape=shape, dtype=dtype)


def ones(shape=None, dtype=float32, device=None, requires_grad=False, **kw):
    a = array(shape=shape, dtype=dtype)
    a.data[...] = 1
    return a


def empty(shape=None, dtype=float32, device=None, **kw):
    return array(shape=shape, dtype=dtype)


def zeros_like(a, device=None, **kw):
    return array.wrap(np.zeros_like(a.data), a.dtype)


def empty_like(a, device=None, **kw):
    return zeros_like(a)


def from_torch(t, dtype=None, **kw):
    d = t.detach().numpy()                     # shares memory with the (CPU) tensor, like wp.from_torch
    if dtype is None:
        dtype = {torch.float32: float32, torch.int32: int32, torch.float64: float64}[t.dtype]
    if d.ndim == len(_tail(dtype)):                # a single vec/mat (e.g. IM = torch.eye(3)) is an array of one element
        d = d[None]
    return array.wrap(d, dtype)


def to_torch(a):
    return torch.from_numpy(a.data)


def from_numpy(arr, dtype=None, device=None, **kw):
    tail = _tail(dtype)
    d = np.array(arr, dtype=_scalar(dtype))
    if tail and d.shape[d.ndim - len(tail):] != tail:
        d = d.reshape((-1,) + tail)
    return array.wrap(d, dtype)


def copy(dest, src, dest_offset=0, src_offset=0, count=0):
    if count == 0:
        count = min(dest.shape[0] - dest_offset, src.shape[0] - src_offset)
    dest.data[dest_offset:dest_offset + count] = src.data[src_offset:src_offset + count]


def init():
    pass


def set_device(d):
    pass


def synchronize():
    pass


def synchronize_stream(s=None):
    pass


def get_stream(device=None):
    return None


# ---------------------------------------------------------------------------------------------- kernels
_tid = 0


def tid():
    return _tid


def _mm(a, b):
    """Warp's `@`: matrix product, or a plain product when both operands are scalars (quai_hexagon.py:51)."""
    if np.ndim(a) == 0 and np.ndim(b) == 0:
        return np.float32(np.float32(a) * np.float32(b))
    r = np.matmul(np.asarray(a), np.asarray(b))
    return r.view(mat) if r.ndim == 2 else r.view(vec)


class _MatMulToCall(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.MatMult):
            return ast.Call(func=ast.Name(id="__wp_mm", ctx=ast.Load()), args=[node.left, node.right], keywords=[])
        return node


class _ValueAssign(ast.NodeTransformer):
    """Warp vectors / matrices are VALUE types: `r = b` copies.  Python names alias, so a later `r[i] = ...` would also
    change `b` (exp2's solve3x3 does `Ap = b; r = b; ...; Ap[i] = 0.0`, exp2/p_hat/quasi_simulation/cublas.py:59-77).  Plain
    name-to-name assignments are routed through a copy."""

    def visit_Assign(self, node):
        self.generic_visit(node)
        if len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) and isinstance(node.value, ast.Name):
            node.value = ast.Call(func=ast.Name(id="__wp_val", ctx=ast.Load()), args=[node.value], keywords=[])
        return node


def _val(v):
    return v.copy() if isinstance(v, np.ndarray) else v


def _recompile(f, fall_through_return=False):
    """Re-compile a kernel/func body with `@` routed through _mm; optionally close an open-ended @wp.func."""
    src = textwrap.dedent(inspect.getsource(f))
    tree = ast.parse(src)
    fd = tree.body[0]
    fd.decorator_list = []
    if fall_through_return:
        names = {r.value.id for r in ast.walk(fd) if isinstance(r, ast.Return) and isinstance(r.value, ast.Name)}
        if not isinstance(fd.body[-1], ast.Return) and len(names) == 1:
            fd.body.append(ast.Return(value=ast.Name(id=names.pop(), ctx=ast.Load())))
    tree = _MatMulToCall().visit(tree)
    tree = _ValueAssign().visit(tree)
    ast.fix_missing_locations(tree)
    f.__globals__["__wp_mm"] = _mm
    f.__globals__["__wp_val"] = _val
    ns = {}
    exec(compile(tree, inspect.getsourcefile(f) or "<wp>", "exec"), f.__globals__, ns)
    return ns[fd.name]


class Kernel:
    def __init__(self, f):
        self.func = _recompile(f)
        self.__name__ = f.__name__
        self.params = list(inspect.signature(f).parameters.values())


def kernel(f=None, **kw):
    if f is None:
        return lambda g: Kernel(g)
    return Kernel(f)


def func(f):
    """A @wp.func whose control flow can fall off the end (solve3x3, cublas.py:59-91) returns the variable its
    explicit `return` statements return."""
    return _recompile(f, fall_through_return=True)


def launch(kernel=None, dim=None, inputs=(), outputs=(), device=None, **kw):
    global _tid
    n = int(dim[0] if isinstance(dim, (tuple, list)) else dim)
    f = kernel.func
    args = list(inputs) + list(outputs)
    for i, p in enumerate(kernel.params):                  # scalar arguments take their annotated Warp type
        if i < len(args) and p.annotation in (float, float32) and not isinstance(args[i], array):
            args[i] = np.float32(args[i])
    for t in range(n):
        _tid = t
        f(*args)


# ---------------------------------------------------------------------------------------------- builtins
class _Tr(float):
    def __matmul__(self, o):
        return np.float32(np.float32(self) * np.float32(o))


def trace(m):
    m = np.asarray(m)
    return _Tr(np.float32(m[0, 0] + m[1, 1] + m[2, 2]))


def transpose(m):
    return np.asarray(m).T.copy().view(mat)


def determinant(m):
    m = np.asarray(m)
    return np.float32(m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
                      + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))


def inverse(m):
    m = np.asarray(m)
    d = determinant(m)
    b = np.empty((3, 3), np.float32)
    b[0, 0] = m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]; b[0, 1] = m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]; b[0, 2] = m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]
    b[1, 0] = m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]; b[1, 1] = m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]; b[1, 2] = m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]
    b[2, 0] = m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]; b[2, 1] = m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]; b[2, 2] = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    return (b * (np.float32(1.0) / d)).view(mat)


def mul(a, b):
    if np.ndim(a) == 2 and np.ndim(b) == 1:
        return np.matmul(np.asarray(a), np.asarray(b)).view(vec)
    if np.ndim(a) == 2 and np.ndim(b) == 2:
        return np.matmul(np.asarray(a), np.asarray(b)).view(mat)
    return a * b


def dot(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.float32(a[0] * b[0] + a[1] * b[1] + a[2] * b[2])


def cross(a, b):
    return np.cross(np.asarray(a), np.asarray(b)).astype(np.float32).view(vec)


def length(a):
    return np.float32(np.sqrt(dot(a, a)))


def normalize(a):
    return (np.asarray(a) / length(a)).view(vec)


def outer(a, b):
    return np.outer(np.asarray(a), np.asarray(b)).astype(np.float32).view(mat)


def cw_mul(a, b):
    return (np.asarray(a) * np.asarray(b)).view(vec)


def diag(d):
    return np.diag(np.asarray(d)).astype(np.float32).view(mat)


def abs(x):  # noqa: A001
    return np.abs(x)


def sqrt(x):
    return np.float32(np.sqrt(x))


def atomic_add(arr, idx, value):
    old = np.array(arr.data[idx], copy=True)
    arr.data[idx] = arr.data[idx] + np.asarray(value)
    return old


def atomic_sub(arr, idx, value):
    old = np.array(arr.data[idx], copy=True)
    arr.data[idx] = arr.data[idx] - np.asarray(value)
    return old


def atomic_max(arr, idx, value):
    old = arr.data[idx]
    arr.data[idx] = max(arr.data[idx], value)
    return old


def lower_bound(arr, lo, hi, value):
    return int(lo + np.searchsorted(arr.data[lo:hi], value, side="left"))


def eig3(A, V, D):
    """wp.eig3 restated (McAdams et al. 2011; see oracle/ref_kernels.py).  Writes V and D in place."""
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import ref_kernels as K
    v, d = K.eig3(np.asarray(A, dtype=np.float32)[None])
    V[...] = v[0]
    D[...] = d[0]
