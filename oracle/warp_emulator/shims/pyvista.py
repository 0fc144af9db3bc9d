"""Stand-in for pyvista: ``pv.read`` keeps the path, ``pv.voxelize`` returns the synthetic voxel mesh it names
('mask:<npz path>' or 'cube:<n>'), points in x-fastest lattice order, cells as VTK hexahedra
(000,100,110,010,001,101,111,011) -- the order object.py:34 permutes to local node 4bx+2by+bz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _Cell:
    def __init__(self, ids):
        self.point_ids = ids


class _Voxels:
    def __init__(self, points, hexagons):
        self.points = np.asarray(points, dtype=np.float32)
        h = np.asarray(hexagons)
        # local a = 4bx+2by+bz  ->  VTK order p0..p7
        vtk = h[:, [0, 4, 6, 2, 1, 5, 7, 3]]
        self.cell = [_Cell([int(v) for v in row]) for row in vtk]


def read(path):
    return path


def voxelize(mesh, density=None, check_surface=False):
    from fem_simulation_b200 import mesh as M
    spec = str(mesh)
    if spec.startswith("cube:"):
        pts, hx = M.cube_mesh(int(spec[5:]), density)
    elif spec.startswith("npz:"):
        d = np.load(spec[4:])
        pts, hx = d["points"], d["hexagons"]
    else:
        raise ValueError(spec)
    return _Voxels(pts, hx)
