"""CPU: the OBJ -> voxel-mesh front end (fem_simulation_b200/voxelize.py, stand-in for pv.read + pv.voxelize,
/root/reference/exp1/quasi_simulation/object.py:28-45).  pyvista is not installed, so its
exact numbering cannot be pinned; these tests hold the construction to the documented algorithm: lattice arange(min, max, dx) per
axis, a point is enclosed iff it is inside the closed surface, a cell is kept iff one of its corners is enclosed, numbering
x-fastest, local node 4bx+2by+bz."""
import os

import numpy as np
import pytest

from fem_simulation_b200 import hierarchy, voxelize

ASSETS = "/root/reference/assets/objs"


def _write_obj(path, verts, faces):
    with open(path, "w") as f:
        for v in verts:
            f.write("v %r %r %r\n" % tuple(float(c) for c in v))
        for t in faces:
            f.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))


def _tetra():
    """An irregular tetrahedron (no lattice point lands on its surface) with outward-oriented faces, and its inside test."""
    v = np.array([[0.013, 0.021, 0.037], [1.071, 0.113, 0.091], [0.233, 0.957, 0.171], [0.317, 0.289, 1.043]])
    f = np.array([(0, 2, 1), (0, 1, 3), (1, 2, 3), (0, 3, 2)])

    def inside(p):
        T = np.linalg.inv((v[1:] - v[0]).T)
        lam = (np.asarray(p) - v[0]) @ T.T
        return (lam > 0).all(axis=-1) & (lam.sum(axis=-1) < 1)
    return v, f, inside


def test_points_inside_matches_the_analytic_tetrahedron():
    v, f, inside = _tetra()
    g = [np.arange(v[:, k].min(), v[:, k].max(), 0.05) for k in range(3)]
    ins = voxelize.points_inside(v, f, *g)
    X, Y, Z = np.meshgrid(*g, indexing="ij")
    ana = inside(np.stack([X, Y, Z], axis=-1))
    assert ana.sum() > 100 and np.array_equal(ins, ana)


def test_voxelize_tetrahedron_mesh_structure(tmp_path):
    v, f, inside = _tetra()
    p = tmp_path / "tetra.obj"
    _write_obj(p, v, f)
    dx = 0.05
    pts, hx = voxelize.voxelize_obj(str(p), dx)
    assert pts.dtype == np.float32 and hx.dtype == np.int32 and hx.shape[1] == 8
    # every point is used; points sit on the lattice min + dx*(i,j,k) and are numbered x-fastest
    assert np.array_equal(np.unique(hx), np.arange(len(pts)))
    lo = v.min(axis=0)
    ijk = np.rint((pts.astype(np.float64) - lo) / dx).astype(np.int64)
    assert np.abs(pts - (lo + dx * ijk)).max() < 1e-5
    key = ijk[:, 0] + 1000 * (ijk[:, 1] + 1000 * ijk[:, 2])
    assert np.all(np.diff(key) > 0)
    # local node a = 4bx + 2by + bz of every cell, cells x-fastest
    base = ijk[hx[:, 0]]
    for a in range(8):
        assert np.array_equal(ijk[hx[:, a]], base + np.array([(a >> 2) & 1, (a >> 1) & 1, a & 1]))
    ckey = base[:, 0] + 1000 * (base[:, 1] + 1000 * base[:, 2])
    assert np.all(np.diff(ckey) > 0)
    # a cell is kept iff at least one of its corners is enclosed
    g = [np.arange(lo[k], v[:, k].max(), dx) for k in range(3)]
    X, Y, Z = np.meshgrid(*g, indexing="ij")
    ana = inside(np.stack([X, Y, Z], axis=-1))
    nx, ny, nz = (len(a) - 1 for a in g)
    keep = np.zeros((nx, ny, nz), bool)
    for bx in (0, 1):
        for by in (0, 1):
            for bz in (0, 1):
                keep |= ana[bx:bx + nx, by:by + ny, bz:bz + nz]
    want = np.argwhere(np.transpose(keep, (2, 1, 0)))[:, ::-1]          # x fastest
    assert np.array_equal(base, want)
    # the result feeds the hierarchy like any other voxel mesh
    H = hierarchy.build_hierarchy(pts, hx, dx)
    assert H.layer >= 2 and H.dims[0] == len(pts)


@pytest.mark.skipif(not os.path.exists(os.path.join(ASSETS, "cube.obj")), reason="reference assets not present")
def test_reference_cube_obj():
    """assets/objs/cube.obj at the drivers' dx = 0.05 (exp1/quasi_simulation/main.py:12 uses the same call on dragon.obj)."""
    pts, hx = voxelize.voxelize_obj(os.path.join(ASSETS, "cube.obj"), 0.05)
    ext = pts.max(axis=0) - pts.min(axis=0)
    n = np.rint(ext / 0.05).astype(int)
    assert len(hx) > 0 and np.array_equal(np.unique(hx), np.arange(len(pts)))
    assert len(hx) <= int(np.prod(n)) and len(hx) >= int(np.prod(n - 2))     # a cube: (almost) every lattice cell is kept
    H = hierarchy.build_hierarchy(pts, hx, 0.05)
    assert H.layer == hierarchy.level_count(pts.min(axis=0), pts.max(axis=0), 0.05)


@pytest.mark.skipif(not os.path.exists(os.path.join(ASSETS, "rocker.obj")), reason="reference assets not present")
def test_reference_rocker_obj_builds_a_hierarchy():
    pts, hx = voxelize.voxelize_obj(os.path.join(ASSETS, "rocker.obj"), 0.05)
    assert len(hx) > 50 and np.array_equal(np.unique(hx), np.arange(len(pts)))
    H = hierarchy.build_hierarchy(pts, hx, 0.05, layers=2)
    us_hat = H.Us_hat[0]
    ps = np.zeros(us_hat.nrow, np.float64); np.add.at(ps, us_hat.rows(), us_hat.weights)
    assert np.allclose(ps, 1.0)                                          # prolongation is a partition of unity
    cv = np.asarray(H.color_vertex_num[0]); p = H.pattern[0]
    assert np.all(np.searchsorted(cv, p.U_row, side="right") != np.searchsorted(cv, p.U_col, side="right"))


def _box_triangles(lo, hi):
    """12 triangles of an axis-aligned box: lattice points land exactly on faces, edges and corners (the tie rules)."""
    (x0, y0, z0), (x1, y1, z1) = lo, hi
    v = np.array([[x0, y0, z0], [x1, y0, z0], [x0, y1, z0], [x1, y1, z0], [x0, y0, z1], [x1, y0, z1], [x0, y1, z1], [x1, y1, z1]], float)
    f = np.array([(0, 2, 1), (1, 2, 3), (4, 5, 6), (5, 7, 6), (0, 1, 4), (1, 5, 4), (2, 6, 3), (3, 6, 7), (0, 4, 2), (2, 4, 6), (1, 3, 5), (3, 7, 5)])
    return v, f


@pytest.mark.parametrize("case", ["tetra", "hull", "box_on_lattice", "empty"])
def test_native_inside_test_matches_numpy_restatement(case):
    """femb_setup_points_inside_host (C++/OpenMP) against voxelize.points_inside_numpy: the same booleans, point for point."""
    from scipy.spatial import ConvexHull
    if case == "tetra":
        v, f, _ = _tetra()
        grids = [[np.arange(v[:, k].min(), v[:, k].max(), dx) for k in range(3)] for dx in (0.05, 0.031, 0.2)]
    elif case == "hull":
        rng = np.random.default_rng(3)
        v = rng.normal(size=(300, 3)) * np.array([1.0, 0.6, 1.7])
        f = ConvexHull(v).simplices
        grids = [[np.arange(v[:, k].min(), v[:, k].max(), dx) for k in range(3)] for dx in (0.11, 0.07)]
    elif case == "box_on_lattice":
        v, f = _box_triangles((0.0, 0.0, 0.0), (1.0, 0.5, 0.75))
        grids = [[np.arange(-0.25, 1.5, 0.125)] * 3, [np.arange(0.0, 1.0, 0.25), np.arange(0.0, 0.5, 0.25), np.arange(0.0, 0.75, 0.25)]]
    else:
        v, f = np.zeros((3, 3)), np.zeros((0, 3), np.int64)
        grids = [[np.arange(0.0, 1.0, 0.25)] * 3]
    for g in grids:
        got = voxelize.points_inside(v, f, *g)
        want = voxelize.points_inside_numpy(v, f, *g)
        assert got.dtype == np.bool_ and got.shape == want.shape and np.array_equal(got, want)
        if case == "hull":
            assert want.sum() > 1000


@pytest.mark.parametrize("name", ["cube.obj", "rocker.obj", "horse.obj"])
def test_native_inside_test_on_reference_assets(name):
    path = os.path.join(ASSETS, name)
    if not os.path.exists(path):
        pytest.skip("reference assets not present")
    v, f = voxelize.read_obj(path)
    g = [np.arange(v[:, k].min(), v[:, k].max(), 0.05) for k in range(3)]
    assert np.array_equal(voxelize.points_inside(v, f, *g), voxelize.points_inside_numpy(v, f, *g))


def test_native_obj_reader_matches_python_parser(tmp_path):
    """femb_setup_obj_count_host / _read_host against the line-by-line Python parser: quads and pentagons as fans, a/b/c index
    forms, negative (relative) indices, comments, other records, exponents; and the reference's own files when present."""
    p = tmp_path / "mixed.obj"
    p.write_text("# comment\nmtllib x.mtl\nv 0 0 0\nv 1.5e0 0 0\nv 1 1 -2.5E-1\nvn 0 0 1\nvt 0.5 0.5\nv 0 1 0.125\n"
                 "f 1 2 3\nf 1/1/1 2/2/1 3//1 4\ng grp\nv 0.3333333333333333 7 1e-3\nf -1 -2 -3 -4 -5\nf 5//2 1//2 3//2\n")
    v, f = voxelize.read_obj(str(p))
    pv_, pf = voxelize.read_obj_python(str(p))
    assert v.dtype == np.float64 and f.dtype == np.int64 and v.shape == (5, 3) and f.shape == (7, 3)
    assert np.array_equal(v, pv_) and np.array_equal(f, pf)
    for name in ["cube.obj", "rocker.obj", "horse.obj"]:
        path = os.path.join(ASSETS, name)
        if os.path.exists(path):
            v, f = voxelize.read_obj(path)
            pv_, pf = voxelize.read_obj_python(path)
            assert np.array_equal(v, pv_) and np.array_equal(f, pf), name
    with pytest.raises(FileNotFoundError):
        voxelize.read_obj(str(tmp_path / "missing.obj"))
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0\n")
    from fem_simulation_b200 import _lib
    with pytest.raises(_lib.FembError):
        voxelize.read_obj(str(bad))


def test_native_mask_mesh_matches_numpy_restatement():
    """femb_setup_mask_mesh_count_host / _fill_host against mesh.mesh_from_mask_numpy: points bit for bit, hexahedra equal --
    boxes, ragged sets, offsets and origins of a slab, an empty set."""
    from fem_simulation_b200 import mesh
    rng = np.random.default_rng(9)
    cases = [(np.ones((4, 3, 5), bool), 0.05, (0.0, 0.0, 0.0), (0, 0, 0)),
             (rng.random((9, 11, 7)) < 0.5, 0.1, (-0.3, 1.7, 0.21), (0, 0, 0)),
             (rng.random((6, 5, 8)) < 0.2, 1.0 / 7, (0.0, 0.0, 0.0), (3, 0, 14)),
             (rng.random((13, 4, 6)) < 0.9, 0.0125, (5.5, -2.25, 1e-3), (0, 7, 0)),
             (np.zeros((3, 3, 3), bool), 0.05, (0.0, 0.0, 0.0), (0, 0, 0))]
    for mask, dx, origin, off in cases:
        p, h = mesh.mesh_from_mask(mask, dx, origin, off)
        q, g = mesh.mesh_from_mask_numpy(mask, dx, origin, off)
        assert p.dtype == np.float32 and h.dtype == np.int32 and p.shape == q.shape and h.shape == g.shape
        assert np.array_equal(p, q) and np.array_equal(h, g)
    with pytest.raises(ValueError):
        mesh.mesh_from_mask(np.ones((3, 3), bool), 0.1)
