"""CPU: the product's vectorised setup against the literal restatement of the reference loops
(oracle/ref_setup.py <- /root/reference/exp1/quasi_simulation/object.py:116-865): every integer map
bit-exact, every weight bit-exact."""
import numpy as np
import pytest

from fem_simulation_b200 import hierarchy, mesh
from oracle import ref_setup

CASES = {
    "cube5": lambda: (mesh.cube_mesh(5, 0.05), 0.05),                                  # 1 level
    "cube9": lambda: (mesh.cube_mesh(9, 0.05), 0.05),                                  # 2 levels
    "jag2": lambda: (mesh.jagged_mesh(10, 9, 11, 0.05, seed=1), 0.05),                 # 2 levels, ragged
    "jag3": lambda: (mesh.jagged_mesh(18, 17, 19, 0.05, seed=2, keep=0.7, origin=(-0.3, 0.1, 0.2)), 0.05),  # 3 levels
    "box_dx01": lambda: (mesh.box_mesh(9, 12, 10, 0.1, origin=(0.5, -1.0, 2.0)), 0.1),
}


def _eq(a, b, name):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{name}: shape {a.shape} != {b.shape}"
    assert np.array_equal(a, b), f"{name}: values differ"


def test_corrected_spacing_bit_exact_and_really_coarsens():
    """with `corrected_spacing` level l is coarsened with spacing dx 2^l instead of the reference's 2 dx for
    every l >= 1 (object.py:175-179).  The literal loops with the same one-line fix and the vectorised setup agree bit for
    bit on a sparse ragged mesh with a 4-level bounding box, and the corrected level 3 is smaller than level 2."""
    points, hexs = mesh.jagged_mesh(33, 32, 34, 0.05, seed=4, keep=0.1)
    Hb = hierarchy.build_hierarchy(points, hexs, 0.05)
    assert Hb.layer == 4 and Hb.dims[3] == Hb.dims[2]                 # the reference's degenerate top level
    _check_setup(points, hexs, 0.05, corrected=True)
    Hc = hierarchy.build_hierarchy(points, hexs, 0.05, corrected_spacing=True)
    assert Hc.dims[:3] == Hb.dims[:3] and Hc.dims[3] < Hc.dims[2]
    us_hat = Hc.Us_hat[2]
    ps = np.zeros(us_hat.nrow, np.float64); np.add.at(ps, us_hat.rows(), us_hat.weights)
    assert np.allclose(ps, 1.0)                                      # still a partition of unity


@pytest.mark.parametrize("case", list(CASES))
def test_setup_bit_exact(case):
    (points, hexs), dx = CASES[case]()
    _check_setup(points, hexs, dx)


def _check_setup(points, hexs, dx, corrected=False):
    S = ref_setup.build_setup(points, hexs, dx, [], corrected_spacing=corrected)
    H = hierarchy.build_hierarchy(points, hexs, dx, corrected_spacing=corrected)
    assert S["layer"] == H.layer and S["dims"] == H.dims
    for l in range(H.layer):
        for nm in ["hexs", "x", "vertex2index", "index2vertex", "color_vertex_num"]:
            _eq(S[nm][l], getattr(H, nm)[l], f"{nm}[{l}]")
        p, q = S["pattern"][l], H.pattern[l]
        for nm in ["nnz", "L_nnz", "off_l", "off_d", "off_u", "L_row", "L_col", "U_row", "U_col", "GS_L_row", "GS_L_col",
                   "GS_U_row", "GS_U_col", "GS_L_Ptr", "GS_U_Ptr", "block_offset", "hex_update_offset", "row", "col"]:
            _eq(p[nm], getattr(q, nm), f"pattern[{l}].{nm}")
    for l in range(H.layer - 1):
        _eq(S["fine_coarse"][l], H.fine_coarse[l], f"fine_coarse[{l}]")
        for nm in ["rowInd", "colInd", "rowInd_noOrder", "colInd_noOrder", "value", "norm"]:
            _eq(S["U_triplets"][l][nm], H.U_triplets[l][nm], f"U_triplets[{l}].{nm}")
        for nm in ["Ut", "Ut_noOrder", "Ut_hat", "Ut_noOrder_hat", "Us", "Us_noOrder", "Us_hat", "Us_noOrder_hat"]:
            a, b = S[nm][l], getattr(H, nm)[l]
            _eq(a["offsets"], b.offsets, nm + ".offsets"); _eq(a["columns"], b.columns, nm + ".columns")
            _eq(a["values"], b.weights, nm + ".values")
    _eq(S["fix_idx"], H.fix_idx, "fix_idx"); _eq(S["fix_values"], H.fix_values, "fix_values")
    _eq(S["shapeFuncGrad"], H.shapeFuncGrad, "shapeFuncGrad")
    _eq(S["pin_list"], hierarchy.default_pins(points, dx), "default pins")


def test_hierarchy_properties_large():
    """Size-independent properties on a mesh the literal oracle cannot finish quickly (n=40: 64 k hexes)."""
    points, hexs = mesh.cube_mesh(40)
    H = hierarchy.build_hierarchy(points, hexs, 1.0 / 40)
    assert H.layer == 4 and H.dims[0] == 41 ** 3
    for l in range(H.layer):
        p = H.pattern[l]
        n = H.dims[l]
        assert np.array_equal(np.sort(H.vertex2index[l]), np.arange(n))
        assert np.array_equal(H.index2vertex[l][H.vertex2index[l]], np.arange(n))
        assert p.nnz == 2 * p.L_nnz + n
        # pattern symmetric: (r,c) in U  <->  (c,r) in L
        assert np.array_equal(np.sort(p.U_row.astype(np.int64) * n + p.U_col), np.sort(p.L_col.astype(np.int64) * n + p.L_row))
        assert p.hex_update_offset.min() == 0 and p.hex_update_offset.max() == p.nnz - 1
        # no two vertices of one colour share a hexahedron (the GS sweeps rely on it)
        cv = np.asarray(H.color_vertex_num[l])
        assert np.all(np.searchsorted(cv, p.U_row, side="right") != np.searchsorted(cv, p.U_col, side="right"))
    for l in range(H.layer - 1):
        ut, us_hat = H.Ut[l], H.Us_hat[l]
        rs = np.zeros(ut.nrow, np.float64); np.add.at(rs, ut.rows(), ut.weights)
        assert np.allclose(rs, 1.0, atol=1e-6)                      # Ut rows are normalised
        ps = np.zeros(us_hat.nrow, np.float64); np.add.at(ps, us_hat.rows(), us_hat.weights)
        assert np.allclose(ps, 1.0)                                 # prolongation is a partition of unity
    # levels >= 3 do not coarsen (reference bug object.py:175-179, reproduced on purpose)
    assert H.dims[3] == H.dims[2]


def test_reference_cpu_functions_match():
    """cpu_function.py is importable here: pin our formulas against the reference's own functions."""
    import os, sys
    ref = "/root/reference/exp1/quasi_simulation"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    import torch
    sys.path.insert(0, ref)
    try:
        import cpu_function as cf
    finally:
        sys.path.remove(ref)
    sfg = torch.zeros((8, 8, 3), dtype=torch.float32)
    help_ = torch.tensor([[-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1], [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], dtype=torch.float32)
    quad = help_ * 0.57735
    cf.calShapeFuncGrad(sfg, help_, quad)
    assert np.array_equal(sfg.numpy(), hierarchy.shape_func_grad())
    assert np.array_equal(sfg.numpy(), ref_setup.cal_shape_func_grad())
    rng = np.random.default_rng(0)
    pts = (rng.integers(0, 40, size=(200, 3)) * 0.05).astype(np.float32) + np.float32(0.3)
    org = np.array([0.3, 0.3, 0.3], np.float32)
    for sp, sh in ((0.05, 0.5), (0.1, 0.1), (0.2, 0.1)):
        mine = hierarchy._ijk(pts, org, sp, sh)
        for k in range(len(pts)):
            if sh == 0.1:
                assert cf.ijk_index(torch.tensor(pts[k]), torch.tensor(org), sp) == mine[k].tolist()
            else:
                c = mine[k] % 2
                assert int(cf.color_ind(torch.tensor(pts[k]), torch.tensor(org), sp)) == int(c[0] + 2 * c[1] + 4 * c[2])


def test_galerkin_plan_emulation():
    """The slot plans of hierarchy.build_galerkin_plan drive the two CUDA stages; emulate both stages in numpy
    and compare with the oracle's bsr_mm(Ut_hat, bsr_mm(A, Us_hat)) (sparse.py:822-1140)."""
    from oracle import ref_solvers
    from oracle.ref_setup import bsr_from_triplets
    points, hexs = mesh.jagged_mesh(10, 9, 11, 0.05, seed=1)
    H = hierarchy.build_hierarchy(points, hexs, 0.05)
    rng = np.random.default_rng(0)
    for l in range(H.layer - 1):
        g = hierarchy.build_galerkin_plan(H, l)
        p, pc, us, ut = H.pattern[l], H.pattern[l + 1], H.Us_hat[l], H.Ut_hat[l]
        vals = rng.normal(size=(p.nnz, 3, 3))
        T = np.zeros((len(g["t_col"]), 3, 3))
        for s in range(p.nnz):
            i, j = p.row[s], p.col[s]
            for pp in range(us.offsets[j + 1] - us.offsets[j]):
                b8 = int(g["t_slot"][s, pp])
                assert g["dyadic"] and 2.0 ** -(b8 >> 5) == us.weights[us.offsets[j] + pp]
                T[g["t_ptr"][i] + (b8 & 31)] += 2.0 ** -(b8 >> 5) * vals[s]
        out = np.zeros((pc.nnz, 3, 3))
        for I in range(H.dims[l + 1]):
            for e in range(ut.offsets[I], ut.offsets[I + 1]):
                i, w, q = ut.columns[e], ut.weights[e], g["ut_q"][e]
                for t in range(g["t_ptr"][i], g["t_ptr"][i + 1]):
                    out[g["c_slot"][g["c_ptr"][I] + g["c_slot8"][t, q]]] += w * T[t]
        A = bsr_from_triplets(H.dims[l], H.dims[l], p.row, p.col, vals.astype(np.float32))
        A["values"] = A["values"].astype(np.float64)
        Us = dict(nrow=us.nrow, ncol=us.ncol, nnz=us.nnz, offsets=us.offsets, columns=us.columns, values=us.weights.astype(np.float64))
        Ut = dict(nrow=ut.nrow, ncol=ut.ncol, nnz=ut.nnz, offsets=ut.offsets, columns=ut.columns, values=ut.weights.astype(np.float64))
        Cm = ref_solvers.bsr_mm(Ut, ref_solvers.bsr_mm(A, Us))
        ref = Cm["values"][pc.block_offset]
        assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()


def test_exp3_hooks_edge_index_and_prolongation_chain():
    """exp3/simulation/object.py:258-276 (edge_index, dictionary order) and :294,429,438-441 (A_noOrder chain behind torch_A),
    vectorised vs the literal loops."""
    points, hexs = mesh.jagged_mesh(18, 17, 19, 0.05, seed=2, keep=0.7, origin=(-0.3, 0.1, 0.2))
    H = hierarchy.build_hierarchy(points, hexs, 0.05)
    hex_edge = [[0, 1], [0, 2], [1, 3], [2, 3], [4, 5], [4, 6], [5, 7], [6, 7], [0, 4], [1, 5], [2, 6], [3, 7]]
    for l in range(H.layer):
        seen = {}
        for h in H.hexs[l]:
            for e0, e1 in hex_edge:
                a, b = int(h[e0]), int(h[e1])
                if a > b:
                    a, b = b, a
                seen.setdefault((a, b), 1)
        lit = np.array(list(seen.keys()), dtype=np.int64).T
        _eq(hierarchy.edge_index(H.hexs[l]), lit, f"edge_index[{l}]")
    chain = hierarchy.prolongation_chain(H)
    assert len(chain) == H.layer - 1
    dense = None
    for l, m in enumerate(chain):
        u = H.Us_noOrder_hat[l]
        d = np.zeros((u.nrow, u.ncol), np.float64)
        d[u.rows(), u.columns] = u.weights
        dense = d if dense is None else dense @ d
        assert m.shape == (H.dims[0], H.dims[l + 1]) and m.has_sorted_indices
        assert np.array_equal(m.toarray().astype(np.float64), dense), f"A_noOrder[{l}]"      # dyadic weights: exact


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 3, 3), (9, 1, 9)])
def test_mesh_below_the_level_formula_is_an_error(shape):
    """object.py:143 gives layer < 1 when the thinnest extent is under ~4 dx; the reference constructor (and its literal
    restatement) then dies with IndexError.  The product raises the same class of error, with a message, instead of
    returning a hierarchy without levels."""
    pts, hx = mesh.box_mesh(*shape, 0.05)
    with pytest.raises(IndexError):
        ref_setup.build_setup(pts, hx, 0.05, [])
    with pytest.raises(IndexError, match="level count"):
        hierarchy.build_hierarchy(pts, hx, 0.05)
    with pytest.raises(ValueError):
        hierarchy.build_hierarchy(pts, hx, 0.05)
    assert hierarchy.build_hierarchy(pts, hx, 0.05, layers=1).layer == 1       # explicit override still works


def test_malformed_meshes_are_rejected():
    """Host-side validation before anything reaches the device: empty, mis-shaped, out-of-range and non-finite input."""
    pts, hx = mesh.cube_mesh(5, 0.05)
    bad = [(np.zeros((0, 3), np.float32), np.zeros((0, 8), np.int32)), (pts, hx[:0]), (pts[:, :2], hx), (pts, hx[:, :4])]
    h2 = hx.copy(); h2[3, 5] = len(pts)
    h3 = hx.copy(); h3[0, 0] = -1
    p2 = pts.copy(); p2[7, 1] = np.nan
    bad += [(pts, h2), (pts, h3), (p2, hx)]
    for p, h in bad:
        with pytest.raises(ValueError):
            hierarchy.build_hierarchy(p, h, 0.05)


def test_host_setup_routines_match_numpy_restatements():
    """csrc/femb_setup.cu (femb_setup_incidence/pattern/galerkin_*_host: per-row sorted unions, OpenMP) vs the numpy / scipy
    restatements (one global sort, scipy products), array by array, on ragged 3-level meshes and a box."""
    for pts, hx, dx in (mesh.jagged_mesh(18, 17, 19, 0.05, seed=2, keep=0.7) + (0.05,), mesh.box_mesh(7, 12, 9, 0.1) + (0.1,),
                        mesh.jagged_mesh(10, 9, 11, 0.05, seed=1, origin=(-0.3, 0.1, 0.2)) + (0.05,)):
        H = hierarchy.build_hierarchy(pts, hx, dx)
        for l in range(H.layer):
            a = H.pattern[l]
            b = hierarchy.build_pattern_numpy(H.hexs[l], H.vertex2index[l], H.color_vertex_num[l], H.dims[l])
            assert (a.n, a.nnz, a.L_nnz, a.off_l, a.off_d, a.off_u) == (b.n, b.nnz, b.L_nnz, b.off_l, b.off_d, b.off_u)
            for f in ("L_row", "L_col", "U_row", "U_col", "L_ptr", "U_ptr", "GS_L_row", "GS_L_col", "GS_U_row", "GS_U_col",
                      "block_offset", "hex_update_offset"):
                assert np.array_equal(getattr(a, f), getattr(b, f)), (l, f)
            assert a.GS_L_Ptr == b.GS_L_Ptr and a.GS_U_Ptr == b.GS_U_Ptr
            ve = hierarchy.vertex_incidence(H.hexs[l], H.dims[l])
            flat = H.hexs[l].reshape(-1)
            for v in range(0, H.dims[l], max(1, H.dims[l] // 50)):
                assert ve[v][ve[v] >= 0].tolist() == np.nonzero(flat == v)[0].tolist()
        for l in range(H.layer - 1):
            g1, g2 = hierarchy.build_galerkin_plan(H, l), hierarchy.build_galerkin_plan_numpy(H, l)
            assert set(g1) == set(g2)
            for k in g2:
                assert np.array_equal(np.asarray(g1[k]), np.asarray(g2[k])), (l, k)


def test_coo_to_csr_pair_matches_numpy_sort():
    """femb_setup_coo_to_csr_pair_host (counting sort by row + per-row insertion sort, two weight arrays on one pattern) vs the
    numpy restatement (one global stable sort): offsets, columns and both weight arrays; the two results share their pattern
    arrays; the transposed pair equals the transposes taken one at a time."""
    rng = np.random.default_rng(3)
    nrow, ncol = 57, 301
    cells = rng.choice(nrow * ncol, size=2000, replace=False)                # unique (row, col) pairs in random order
    rows, cols = (cells // ncol).astype(np.int32), (cells % ncol).astype(np.int32)
    wa, wb = rng.normal(size=2000).astype(np.float32), rng.normal(size=2000).astype(np.float32)
    a, b = hierarchy.coo_to_csr_pair(nrow, ncol, rows, cols, wa, wb)
    ra, rb = hierarchy.coo_to_csr_numpy(nrow, ncol, rows, cols, wa), hierarchy.coo_to_csr_numpy(nrow, ncol, rows, cols, wb)
    for got, want in ((a, ra), (b, rb)):
        assert (got.nrow, got.ncol) == (want.nrow, want.ncol)
        assert np.array_equal(got.offsets, want.offsets) and np.array_equal(got.columns, want.columns)
        assert np.array_equal(got.weights, want.weights)
    assert a.offsets is b.offsets and a.columns is b.columns
    one = hierarchy.coo_to_csr(nrow, ncol, rows, cols, wa)
    assert np.array_equal(one.offsets, ra.offsets) and np.array_equal(one.columns, ra.columns) and np.array_equal(one.weights, ra.weights)
    ta, tb = hierarchy.transposed_pair(a, b)
    for got, want in ((ta, a.transposed()), (tb, b.transposed())):
        assert np.array_equal(got.offsets, want.offsets) and np.array_equal(got.columns, want.columns)
        assert np.array_equal(got.weights, want.weights)
    empty_a, empty_b = hierarchy.coo_to_csr_pair(4, 5, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros(0, np.float32))
    assert empty_a.nnz == 0 and empty_b.offsets.tolist() == [0, 0, 0, 0, 0]


def _same_hierarchy(Ha, Hb):
    assert Ha.layer == Hb.layer and Ha.dims == Hb.dims and Ha.spacing == Hb.spacing
    for l in range(Ha.layer):
        for nm in ["hexs", "x", "vertex2index", "index2vertex", "color_vertex_num"]:
            _eq(getattr(Ha, nm)[l], getattr(Hb, nm)[l], f"{nm}[{l}]")
    for l in range(Ha.layer - 1):
        _eq(Ha.fine_coarse[l], Hb.fine_coarse[l], f"fine_coarse[{l}]")
        _eq(Ha.fine_hex_parent[l], Hb.fine_hex_parent[l], f"fine_hex_parent[{l}]")
        for nm in ["rowInd", "colInd", "rowInd_noOrder", "colInd_noOrder", "value", "norm"]:
            _eq(Ha.U_triplets[l][nm], Hb.U_triplets[l][nm], f"U_triplets[{l}].{nm}")
        for nm in ["Ut", "Ut_hat", "Ut_noOrder", "Us", "Us_hat", "Us_noOrder_hat"]:
            a, b = getattr(Ha, nm)[l], getattr(Hb, nm)[l]
            _eq(a.offsets, b.offsets, f"{nm}[{l}].offsets"); _eq(a.columns, b.columns, f"{nm}[{l}].columns")
            _eq(a.weights, b.weights, f"{nm}[{l}].weights")
    _eq(Ha.fix_idx, Hb.fix_idx, "fix_idx"); _eq(Ha.fix_values, Hb.fix_values, "fix_values")


@pytest.mark.parametrize("case", ["cube40", "jag_sparse", "offset_origin", "corrected"])
def test_native_hierarchy_matches_numpy_restatement(case):
    """csrc/femb_setup_hierarchy.cu (colour order, coarsening, interpolation entries: ordered scans over lattice tables)
    against the numpy restatement (np.unique + first occurrence), which the tests above hold to the literal loops."""
    kw = {}
    if case == "cube40":
        (points, hexs), dx = mesh.cube_mesh(40), 1.0 / 40                        # 5 levels, 64 000 hexahedra
    elif case == "jag_sparse":
        (points, hexs), dx = mesh.jagged_mesh(37, 41, 35, 0.05, seed=7, keep=0.3, origin=(-0.7, 0.3, 1.1)), 0.05
    elif case == "offset_origin":                                                # a slab that counts from the global origin
        (points, hexs), dx = mesh.box_mesh(12, 10, 9, 0.1, origin=(0.0, 0.0, 0.0), index_offset=(0, 0, 6)), 0.1
        kw = dict(origin=(0.0, 0.0, 0.0))
    else:
        (points, hexs), dx = mesh.jagged_mesh(33, 32, 34, 0.05, seed=4, keep=0.1), 0.05
        kw = dict(corrected_spacing=True)
    _same_hierarchy(hierarchy.build_hierarchy(points, hexs, dx, **kw), hierarchy.build_hierarchy_numpy(points, hexs, dx, **kw))


def test_native_hierarchy_hashed_lattice_and_stencil_table():
    """Two small blocks far apart: the coarse lattice of the bounding box (300^3 cells) is far larger than the mesh, the
    lookup tables switch from dense to hashed; numbering unchanged.  The stencil table equals the Python one."""
    import ctypes as C
    from fem_simulation_b200 import _lib
    lib = _lib.load()
    cnt, corners, w = np.empty(27, np.int32), np.empty((27, 8), np.int32), np.empty(27, np.float32)
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.femb_setup_stencil_host(hp(cnt), hp(corners), hp(w)) == 0
    _eq(cnt, hierarchy._ST_CNT, "cnt"); _eq(corners, hierarchy._ST_CORNERS, "corners"); _eq(w, hierarchy._ST_W, "weight")
    pa, ha = mesh.box_mesh(5, 6, 4, 0.05)
    pb, hb = mesh.box_mesh(4, 5, 6, 0.05, index_offset=(596, 594, 595))
    points, hexs = np.concatenate([pa, pb]), np.concatenate([ha, hb + len(pa)])
    _same_hierarchy(hierarchy.build_hierarchy(points, hexs, 0.05, layers=3), hierarchy.build_hierarchy_numpy(points, hexs, 0.05, layers=3))


def test_native_hierarchy_rejects_bad_input():
    import ctypes as C
    from fem_simulation_b200 import _lib
    lib = _lib.load()
    points, hexs = mesh.cube_mesh(4, 0.05)
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    parent, key = np.empty(len(hexs), np.int32), np.empty((len(hexs), 3), np.int32)
    n_ch, n_cv = C.c_int(0), C.c_int(0)
    ref = lambda v: C.cast(C.byref(v), C.c_void_p)
    above = np.array([1.0, 0.0, 0.0], np.float32)                                # origin above the mesh: negative cells
    assert lib.femb_setup_coarsen_count_host(len(points), len(hexs), hp(points), hp(hexs), hp(above), C.c_float(0.1), hp(parent),
                                             hp(key), ref(n_ch), ref(n_cv)) == -1
    assert b"origin" in lib.femb_last_error()
    bad = hexs.copy(); bad[3, 0] = len(points)
    zero = np.zeros(3, np.float32)
    assert lib.femb_setup_coarsen_count_host(len(points), len(hexs), hp(points), hp(bad), hp(zero), C.c_float(0.1), hp(parent),
                                             hp(key), ref(n_ch), ref(n_cv)) == -1
    fc = np.full((2, 27), -1, np.int32); fc[1, 3] = 99
    first, n_ent = np.empty(10, np.int32), C.c_int(0)
    assert lib.femb_setup_interp_count_host(10, 2, hp(fc), hp(first), ref(n_ent)) == -1


@pytest.mark.parametrize("case", ["jag3", "cube12_corrected", "slab_origin"])
def test_one_call_hierarchy_matches_python_build(case):
    """femb_setup_hierarchy_host (the whole of Object.__init__'s integer setup behind one C call, library-owned arrays)
    returns the arrays hierarchy.build_hierarchy assembles from the individual setup calls."""
    import ctypes as C
    from fem_simulation_b200 import _lib
    lib = _lib.load()
    corrected, origin = False, None
    if case == "jag3":
        (points, hexs), dx = CASES["jag3"]()
    elif case == "cube12_corrected":
        (points, hexs), dx, corrected = mesh.cube_mesh(12, 0.05), 0.05, True
    else:
        (points, hexs), dx = mesh.box_mesh(12, 10, 9, 0.1, index_offset=(0, 0, 6)), 0.1
        origin = np.zeros(3, np.float32)
    H = hierarchy.build_hierarchy(points, hexs, dx, corrected_spacing=corrected, origin=origin)
    hp = lambda a: a.ctypes.data_as(C.c_void_p)
    out = C.POINTER(_lib.FembHierarchy)()
    rc = lib.femb_setup_hierarchy_host(len(points), len(hexs), hp(points), hp(hexs), C.c_float(dx), H.layer, int(corrected),
                                       None if origin is None else hp(origin), C.cast(C.byref(out), C.c_void_p))
    assert rc == 0, lib.femb_last_error()
    try:
        h = out.contents
        assert h.n_levels == H.layer
        _eq(np.array(h.origin[:], np.float32), H.box_min, "origin")
        arr = lambda ptr, n, shape=None: np.ctypeslib.as_array(ptr, shape=(int(n),)).copy().reshape(shape or (int(n),))
        for l in range(H.layer):
            L, p = h.level[l], H.pattern[l]
            assert (L.n_vert, L.n_hex, L.L_nnz) == (H.dims[l], len(H.hexs[l]), p.L_nnz)
            assert np.float32(L.spacing) == np.float32(H.spacing[l])
            _eq(arr(L.hex, L.n_hex * 8, (-1, 8)), H.hexs[l], f"hex[{l}]")
            _eq(arr(L.x, L.n_vert * 3, (-1, 3)), H.x[l], f"x[{l}]")
            _eq(arr(L.v2i, L.n_vert), H.vertex2index[l], f"v2i[{l}]")
            _eq(arr(L.i2v, L.n_vert), H.index2vertex[l], f"i2v[{l}]")
            _eq(list(L.color_vertex_num), H.color_vertex_num[l], f"cvn[{l}]")
            for nm in ["L_ptr", "U_ptr"]:
                _eq(arr(getattr(L, nm), L.n_vert + 1), getattr(p, nm), f"{nm}[{l}]")
            for nm in ["L_col", "U_col"]:
                _eq(arr(getattr(L, nm), L.L_nnz), getattr(p, nm), f"{nm}[{l}]")
            _eq(arr(L.elem2nnz, L.n_hex * 64), p.hex_update_offset, f"elem2nnz[{l}]")
            _eq(arr(L.block_offset, p.nnz), p.block_offset, f"block_offset[{l}]")
            if l == H.layer - 1:
                assert not L.hex_parent and not L.fine_coarse and L.n_entries == 0 and not L.ut_off
                continue
            t, nc = H.U_triplets[l], H.dims[l + 1]
            _eq(arr(L.hex_parent, L.n_hex), H.fine_hex_parent[l], f"hex_parent[{l}]")
            _eq(arr(L.fine_coarse, len(H.hexs[l + 1]) * 27, (-1, 27)), H.fine_coarse[l], f"fine_coarse[{l}]")
            assert L.n_entries == len(t["value"])
            _eq(arr(L.coarse_id, L.n_entries), t["rowInd_noOrder"], "coarse_id"); _eq(arr(L.fine_id, L.n_entries), t["colInd_noOrder"], "fine_id")
            _eq(arr(L.weight, L.n_entries), t["value"], "weight"); _eq(arr(L.norm, nc), t["norm"], "norm")
            for pre, m, mh, nrow in [("ut", H.Ut[l], H.Ut_hat[l], nc), ("us", H.Us[l], H.Us_hat[l], L.n_vert)]:
                _eq(arr(getattr(L, pre + "_off"), nrow + 1), m.offsets, pre + "_off")
                _eq(arr(getattr(L, pre + "_col"), L.n_entries), m.columns, pre + "_col")
                _eq(arr(getattr(L, pre + "_w"), L.n_entries), m.weights, pre + "_w")
                _eq(arr(getattr(L, pre + "_hat_w"), L.n_entries), mh.weights, pre + "_hat_w")
            if l == 0:
                fix_idx = np.full((L.n_vert, 8), -1, np.int32)
                fix_idx[arr(L.fine_id, L.n_entries), arr(L.entry_k, L.n_entries)] = arr(L.coarse_id, L.n_entries)
                _eq(fix_idx, H.fix_idx, "fix_idx from entry_k")
    finally:
        assert lib.femb_setup_hierarchy_free(out) == 0
    bad = C.POINTER(_lib.FembHierarchy)()
    assert lib.femb_setup_hierarchy_host(len(points), len(hexs), hp(points), hp(hexs), C.c_float(dx), 0, 0, None,
                                         C.cast(C.byref(bad), C.c_void_p)) == -1 and not bad


def test_native_hierarchy_random_voxel_sets():
    """25 random voxel sets (sizes, spacings, fill ratios, origins, both spacing rules): the C++ build and the numpy restatement
    agree on every array (362 such cases were run when the C++ build was written)."""
    rng = np.random.default_rng(5)
    done = 0
    while done < 25:
        nx, ny, nz = rng.integers(4, 22, 3)
        dx = float(rng.choice([0.05, 0.1, 0.03, 0.0125, 1.0 / 7, 0.3]))
        mask = rng.random((nx, ny, nz)) < float(rng.choice([0.15, 0.4, 0.8, 1.0]))
        mask[rng.integers(0, nx), rng.integers(0, ny), rng.integers(0, nz)] = True
        points, hexs = mesh.mesh_from_mask(mask, dx, tuple(float(v) for v in rng.uniform(-3, 3, 3)))
        corrected = bool(rng.integers(0, 2))
        try:
            Hn = hierarchy.build_hierarchy_numpy(points, hexs, dx, corrected_spacing=corrected)
        except hierarchy.MeshTooSmallError:
            continue
        _same_hierarchy(hierarchy.build_hierarchy(points, hexs, dx, corrected_spacing=corrected), Hn)
        done += 1
