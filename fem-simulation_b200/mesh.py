"""Voxel hexahedral meshes in the reference's numbering.

The reference obtains ``(points, cells)`` from ``pv.voxelize`` and permutes the VTK node order
to local node ``a = 4*bx + 2*by + bz`` (/root/reference/exp1/quasi_simulation/object.py:28-45).
These helpers generate the same kind of mesh without pyvista: lattice points numbered
x-fastest and restricted to the used ones, hexahedra numbered x-fastest.
"""
from __future__ import annotations

import numpy as np


def mesh_from_mask(mask: np.ndarray, dx: float, origin=(0.0, 0.0, 0.0), index_offset=(0, 0, 0)):
    """``mask[ix,iy,iz]`` (bool) marks the kept voxels.  Returns (points f32[N,3], hexagons i32[Nh,8]).
    ``index_offset`` shifts the integer lattice coordinates before they are scaled (a slab of a larger lattice gets
    bit-identical coordinates to the full mesh).  Through the host routines of the C ABI (femb_setup_mask_mesh_count_host /
    _fill_host); `mesh_from_mask_numpy` is the numpy restatement the tests compare it with."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    m = np.ascontiguousarray(np.asarray(mask, dtype=bool)).view(np.uint8)
    if m.ndim != 3 or min(m.shape) < 1:
        raise ValueError(f"mask must be a non-empty [nx, ny, nz] array, got shape {m.shape}")
    nx, ny, nz = (int(v) for v in m.shape)
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    nv, nh = C.c_int(0), C.c_int(0)
    _lib.check(lib.femb_setup_mask_mesh_count_host(nx, ny, nz, hp(m), C.cast(C.byref(nv), C.c_void_p), C.cast(C.byref(nh), C.c_void_p)),
               "femb_setup_mask_mesh_count_host")
    points, hexagons = np.empty((nv.value, 3), np.float32), np.empty((nh.value, 8), np.int32)
    o = np.ascontiguousarray(origin, dtype=np.float32).reshape(3)
    off = np.ascontiguousarray(index_offset, dtype=np.int32).reshape(3)
    _lib.check(lib.femb_setup_mask_mesh_fill_host(nx, ny, nz, hp(m), C.c_float(dx), hp(o), hp(off), nv.value, nh.value, hp(points),
                                                  hp(hexagons)), "femb_setup_mask_mesh_fill_host")
    return points, hexagons


def mesh_from_mask_numpy(mask: np.ndarray, dx: float, origin=(0.0, 0.0, 0.0), index_offset=(0, 0, 0)):
    """numpy restatement of `mesh_from_mask`."""
    mask = np.asarray(mask, dtype=bool)
    nx, ny, nz = mask.shape
    iz, iy, ix = np.nonzero(np.transpose(mask, (2, 1, 0)))          # x fastest, then y, then z
    ix = ix.astype(np.int64); iy = iy.astype(np.int64); iz = iz.astype(np.int64)
    px, py = nx + 1, ny + 1
    corner = np.empty((len(ix), 8), dtype=np.int64)
    for a in range(8):
        bx, by, bz = (a >> 2) & 1, (a >> 1) & 1, a & 1
        corner[:, a] = (ix + bx) + px * ((iy + by) + py * (iz + bz))
    mark = np.zeros((nx + 1) * (ny + 1) * (nz + 1), dtype=bool)     # the used lattice points, ascending (= np.unique(corner))
    mark[corner] = True
    used = np.nonzero(mark)[0]
    remap = np.full((nx + 1) * (ny + 1) * (nz + 1), -1, dtype=np.int64)
    remap[used] = np.arange(len(used))
    hexagons = remap[corner].astype(np.int32)
    gx = used % px
    gy = (used // px) % py
    gz = used // (px * py)
    o = np.asarray(origin, dtype=np.float32)
    off = np.asarray(index_offset, dtype=np.int64)
    points = (np.stack([gx, gy, gz], axis=1) + off).astype(np.float32) * np.float32(dx) + o
    return points.astype(np.float32), hexagons


def cube_mesh(n: int, dx: float | None = None, origin=(0.0, 0.0, 0.0)):
    """n x n x n voxels; dx defaults to 1/n."""
    if dx is None:
        dx = 1.0 / n
    return mesh_from_mask(np.ones((n, n, n), dtype=bool), dx, origin)


def box_mesh(nx: int, ny: int, nz: int, dx: float, origin=(0.0, 0.0, 0.0), index_offset=(0, 0, 0)):
    return mesh_from_mask(np.ones((nx, ny, nz), dtype=bool), dx, origin, index_offset)


def jagged_mesh(nx: int, ny: int, nz: int, dx: float, seed: int = 0, keep: float = 0.8, origin=(0.0, 0.0, 0.0)):
    """A reproducible irregular voxel set (random carving that keeps the bounding box corners)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((nx, ny, nz)) < keep
    mask[0, 0, 0] = mask[-1, -1, -1] = mask[0, -1, 0] = mask[-1, 0, -1] = True
    return mesh_from_mask(mask, dx, origin)
