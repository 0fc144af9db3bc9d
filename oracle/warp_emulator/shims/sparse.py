"""Stand-in for the reference's vendored ``sparse.py`` (patched warp.sparse; it needs Warp's native
``bsr_matrix_from_triplets`` / ``bsr_transpose``, which do not exist here).  Same API for the five entry points the hot
path uses -- bsr_zeros, bsr_set_from_triplets (with the reference's value/row/col offset + nnz windowing,
sparse.py:173-278), bsr_transposed, bsr_mv (with x/y offsets, sparse.py:1143-1261), bsr_mm (sparse.py:950-1140) -- and
the documented semantics: COO sorted by (row, col), duplicates summed, structural product pattern, values accumulated
per row in (x_block, y_block) order, all in float32.   ORACLE infrastructure only."""
import ctypes  # noqa: F401  (object.py reaches ctypes and warp through `from sparse import *`)
from typing import Any, Generic, Optional, Tuple, TypeVar, Union  # noqa: F401

import numpy as np
import warp  # noqa: F401
import warp as wp


class BsrMatrix:
    def __init__(self, nrow, ncol, block_type):
        self.nrow, self.ncol, self.nnz = int(nrow), int(ncol), 0
        self.block_type = block_type
        self.offsets = wp.zeros(nrow + 1, dtype=wp.int32)
        self.columns = wp.zeros(0, dtype=wp.int32)
        self.values = wp.zeros(0, dtype=block_type)

    @property
    def scalar_type(self):
        return wp.float32

    @property
    def block_shape(self):
        return (3, 3)


def bsr_zeros(rows_of_blocks, cols_of_blocks, block_type, device=None):
    return BsrMatrix(rows_of_blocks, cols_of_blocks, block_type)


def _from_coo(dest, rows, cols, vals):
    order = np.lexsort((cols, rows))
    r, c = rows[order], cols[order]
    v = vals[order] if vals is not None else None
    if len(r):
        new = np.ones(len(r), dtype=bool)
        new[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
        seg = np.cumsum(new) - 1
        if v is not None:
            out = np.zeros((seg[-1] + 1, 3, 3), np.float32)
            np.add.at(out, seg, v)
            v = out
        r, c = r[new], c[new]
    off = np.zeros(dest.nrow + 1, np.int64)
    np.add.at(off, r + 1, 1)
    dest.offsets = wp.array.wrap(np.cumsum(off).astype(np.int32), wp.int32)
    dest.columns = wp.array.wrap(c.astype(np.int32), wp.int32)
    dest.nnz = len(r)
    dest.values = wp.array.wrap(v if v is not None else np.zeros((len(r), 3, 3), np.float32), wp.mat33f)


def bsr_set_from_triplets(dest, rows, columns, values, value_offset=0, row_offset=0, col_offset=0, nnz=0):
    if nnz == 0:
        nnz = rows.shape[0]
    if nnz + row_offset > rows.shape[0]:
        raise ValueError("rows < nnz + row_offset")
    if nnz + col_offset > columns.shape[0]:
        raise ValueError("columns < nnz + col_offset")
    if nnz + value_offset > values.shape[0]:
        raise ValueError("values < nnz + value_offset")
    r = rows.data[row_offset:row_offset + nnz].astype(np.int64)
    c = columns.data[col_offset:col_offset + nnz].astype(np.int64)
    v = values.data[value_offset:value_offset + nnz].astype(np.float32)
    _from_coo(dest, r, c, v)


def bsr_transposed(src):
    dest = BsrMatrix(src.ncol, src.nrow, src.block_type)
    rows = np.repeat(np.arange(src.nrow, dtype=np.int64), np.diff(src.offsets.data))
    _from_coo(dest, src.columns.data.astype(np.int64), rows, np.swapaxes(src.values.data, 1, 2).copy())
    return dest


def bsr_mv(A, x, y=None, x_offset=0, y_offset=0, alpha=1.0, beta=0.0, work_buffer=None):
    if A.ncol + x_offset > x.shape[0]:
        raise ValueError("Number of columns of A + offset of x > number of rows of x")
    if A.nrow + y_offset > y.shape[0]:
        raise ValueError("Number of rows of A + offset of y > number of rows of y")
    xs = x.data.copy() if x is y else x.data
    v = np.zeros((A.nrow, 3), np.float32)
    if alpha != 0.0 and A.nnz:
        rows = np.repeat(np.arange(A.nrow), np.diff(A.offsets.data))
        contrib = np.einsum("nij,nj->ni", A.values.data[:A.nnz], xs[A.columns.data[:A.nnz].astype(np.int64) + x_offset]).astype(np.float32)
        np.add.at(v, rows, contrib)                       # per row, in block order (thread per row, sparse.py:1162-1166)
        v = v * np.float32(alpha)
    ys = y.data[y_offset:y_offset + A.nrow]
    if beta != 0.0:
        v = v + np.float32(beta) * ys
    ys[...] = v
    return y


def bsr_mm(x, y, z=None, alpha=1.0, beta=0.0, work_arrays=None):
    xr = np.repeat(np.arange(x.nrow, dtype=np.int64), np.diff(x.offsets.data))
    xc = x.columns.data[:x.nnz].astype(np.int64)
    yoff = y.offsets.data.astype(np.int64)
    ycnt = (yoff[1:] - yoff[:-1])[xc]
    xk = np.repeat(np.arange(x.nnz, dtype=np.int64), ycnt)
    within = np.arange(len(xk), dtype=np.int64) - np.repeat(np.cumsum(ycnt) - ycnt, ycnt)
    ym = yoff[xc][xk] + within
    vals = np.matmul(np.float32(alpha) * x.values.data[:x.nnz][xk], y.values.data[:y.nnz][ym]).astype(np.float32)
    out = BsrMatrix(x.nrow, y.ncol, x.block_type)
    _from_coo(out, xr[xk], y.columns.data[:y.nnz].astype(np.int64)[ym], vals)
    return out


# exp2's sparse.py adds a `*_with_grad` twin (values allocated with requires_grad=True; same triplet semantics,
# /root/reference/exp2/p_hat/quasi_simulation/sparse.py:293-372)
def bsr_set_from_triplets_with_grad(dest, rows, columns, values, value_offset=0, row_offset=0, col_offset=0, nnz=0):
    return bsr_set_from_triplets(dest, rows, columns, values, value_offset=value_offset, row_offset=row_offset,
                                 col_offset=col_offset, nnz=nnz)
