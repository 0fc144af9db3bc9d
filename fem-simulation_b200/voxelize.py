"""Surface OBJ -> voxel hexahedral mesh (stand-in for ``pv.read`` + ``pv.voxelize``,
/root/reference/exp1/quasi_simulation/object.py:28-34).

Per-axis lattice ``arange(min, max, dx)``; a lattice point is *enclosed* if it lies inside the closed
surface (parity of ray crossings along +z); every lattice cell with at least one enclosed corner is
kept with all 8 of its corners.  Numbering: points x-fastest restricted to the used ones, cells
x-fastest, local node 4bx+2by+bz.  pyvista is absent here, so its exact point numbering cannot be
pinned -- the construction follows its documented algorithm.
"""
from __future__ import annotations

import numpy as np

from .mesh import mesh_from_mask


def read_obj(path):
    """vertices [N,3] float64 and triangles [F,3] int64 of a Wavefront OBJ file (polygons as fans), through the host routines
    of the C ABI (femb_setup_obj_count_host / _read_host); `read_obj_python` is the restatement the tests compare it with."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    import os
    if not os.path.isfile(path):
        raise FileNotFoundError(path)                        # what open() / pv.read raise
    nv, nt = C.c_longlong(0), C.c_longlong(0)
    bpath = str(path).encode()
    _lib.check(lib.femb_setup_obj_count_host(bpath, C.cast(C.byref(nv), C.c_void_p), C.cast(C.byref(nt), C.c_void_p)),
               "femb_setup_obj_count_host")
    verts, faces = np.empty((nv.value, 3), dtype=np.float64), np.empty((nt.value, 3), dtype=np.int64)
    _lib.check(lib.femb_setup_obj_read_host(bpath, nv.value, nt.value, verts.ctypes.data_as(C.c_void_p), faces.ctypes.data_as(C.c_void_p)),
               "femb_setup_obj_read_host")
    return verts, faces


def read_obj_python(path):
    verts, faces = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                verts.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(t.split("/")[0]) for t in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(verts, dtype=np.float64), np.asarray(faces, dtype=np.int64)


def points_inside(verts, faces, gx, gy, gz):
    """inside[ix,iy,iz] for lattice points (gx[ix], gy[iy], gz[iz]): crossing parity of the +z ray, through the host routine
    of the C ABI (femb_setup_points_inside_host, csrc/femb_setup_hierarchy.cu: same float64 expressions as
    `points_inside_numpy`, the restatement tests compare it with)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    tri = np.ascontiguousarray(np.asarray(verts, dtype=np.float64)[np.asarray(faces, dtype=np.int64)].reshape(-1, 9))
    g = [np.ascontiguousarray(a, dtype=np.float64) for a in (gx, gy, gz)]
    inside = np.empty((len(g[0]), len(g[1]), len(g[2])), dtype=np.uint8)
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.femb_setup_points_inside_host(tri.shape[0], hp(tri), len(g[0]), hp(g[0]), len(g[1]), hp(g[1]), len(g[2]), hp(g[2]),
                                                 hp(inside)), "femb_setup_points_inside_host")
    return inside.view(bool)


def points_inside_numpy(verts, faces, gx, gy, gz):
    """inside[ix,iy,iz] for lattice points (gx[ix], gy[iy], gz[iz]): crossing parity of the +z ray (numpy restatement)."""
    inside = np.zeros((len(gx), len(gy), len(gz)), dtype=bool)
    tri = verts[faces]                                       # [F,3,3]
    for t in tri:
        x0, x1 = t[:, 0].min(), t[:, 0].max()
        y0, y1 = t[:, 1].min(), t[:, 1].max()
        ix = np.nonzero((gx >= x0) & (gx <= x1))[0]
        iy = np.nonzero((gy >= y0) & (gy <= y1))[0]
        if len(ix) == 0 or len(iy) == 0:
            continue
        px, py = np.meshgrid(gx[ix], gy[iy], indexing="ij")
        a, b, c = t[0], t[1], t[2]
        d = (b[1] - c[1]) * (a[0] - c[0]) + (c[0] - b[0]) * (a[1] - c[1])
        if abs(d) < 1e-300:
            continue
        l0 = ((b[1] - c[1]) * (px - c[0]) + (c[0] - b[0]) * (py - c[1])) / d
        l1 = ((c[1] - a[1]) * (px - c[0]) + (a[0] - c[0]) * (py - c[1])) / d
        l2 = 1.0 - l0 - l1
        # half-open edge rule so that a ray through a shared edge counts once
        hit = (l0 >= 0) & (l1 >= 0) & (l2 > 0) if d > 0 else (l0 > 0) & (l1 > 0) & (l2 >= 0)
        zc = l0 * a[2] + l1 * b[2] + l2 * c[2]
        for (i, j) in zip(*np.nonzero(hit)):
            inside[ix[i], iy[j], :] ^= gz > zc[i, j]
    return inside


def voxelize_obj(path, dx):
    verts, faces = read_obj(path)
    lo, hi = verts.min(axis=0), verts.max(axis=0)
    gx, gy, gz = (np.arange(lo[k], hi[k], dx) for k in range(3))
    ins = points_inside(verts, faces, gx, gy, gz)
    nx, ny, nz = len(gx) - 1, len(gy) - 1, len(gz) - 1
    if min(nx, ny, nz) < 1:
        raise ValueError("mesh is smaller than one voxel")
    mask = np.zeros((nx, ny, nz), dtype=bool)
    for bx in range(2):
        for by in range(2):
            for bz in range(2):
                mask |= ins[bx:bx + nx, by:by + ny, bz:bz + nz]
    return mesh_from_mask(mask, dx, origin=lo.astype(np.float32))
