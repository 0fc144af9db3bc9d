"""GPU parity: every operator of the C ABI and every solver entry point of ``Object`` against the CPU
oracle (oracle/ref_*.py, the line-by-line restatement of the reference) on identical seeded inputs.

Tolerances (fp32 path, BASELINE.json north_star): integer maps bit-exact (tests/test_hierarchy.py);
floating-point arrays within REL = 1e-5 of the oracle, measured as max|a-b| / max|b| (the oracle's own
fp32-vs-fp64 gap is of the same order, see test_oracle.py).
"""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from fem_simulation_b200 import Object, _lib, mesh  # noqa: E402
from oracle import ref_kernels as K  # noqa: E402
from oracle import ref_solvers  # noqa: E402

REL = 1e-5


def close(a, b, rel=REL, what=""):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = max(float(np.abs(b).max()), 1e-30)
    err = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
    assert err <= rel * scale, f"{what}: max err {err:.3e} > {rel:g} * {scale:.3e}"


def close64(a, b32, b64, rel=REL, what=""):
    """|a - exact| <= max(rel * scale, 2 * |oracle_fp32 - exact|): where fp32 summation noise of the path itself
    exceeds `rel`, the CUDA result must be as close to the float64 twin as the float32 oracle is."""
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    b64 = np.asarray(b64, dtype=np.float64)
    scale = max(float(np.abs(b64).max()), 1e-30)
    err = float(np.abs(a.astype(np.float64) - b64).max())
    noise = float(np.abs(np.asarray(b32, dtype=np.float64) - b64).max())
    assert err <= max(rel * scale, 2.0 * noise), f"{what}: err {err:.3e}, fp32-oracle noise {noise:.3e}, scale {scale:.3e}"


MESHES = {
    "cube9": lambda: mesh.cube_mesh(9, 0.05),
    "jag2": lambda: mesh.jagged_mesh(10, 9, 11, 0.05, seed=1),
    "jag3": lambda: mesh.jagged_mesh(18, 17, 19, 0.05, seed=2, keep=0.7, origin=(-0.3, 0.1, 0.2)),
}


@pytest.fixture(scope="module", params=list(MESHES))
def pair(request):
    pts, hx = MESHES[request.param]()
    obj = Object((pts, hx), 0.05, [], use_graphs=False)
    ref = ref_solvers.RefObject(pts, hx, 0.05, [])
    ref.ref64 = ref_solvers.RefObject(pts, hx, 0.05, [], dtype=np.float64, setup=ref.S)
    rng = np.random.default_rng(7)
    xp = (ref.x_gpu_layer[0] + rng.normal(size=ref.x_gpu_layer[0].shape).astype(np.float32) * 0.002).astype(np.float32)
    return obj, ref, xp


def test_rest_data(pair):
    obj, ref, _ = pair
    for l in range(obj.layer):
        close(obj.x_gpu_layer[l] if l else obj.x_cpu, ref.x_gpu_layer[l] if l else ref.x_cpu, what=f"x_gpu_layer[{l}]")
        close(obj.det_pX_peps_gpu[l], ref.det[l], what=f"det[{l}]")
        close(obj.inverse_pX_peps_gpu[l], ref.inv[l], what=f"inv[{l}]")
        close(obj.m_gpu[l], ref.m[l], what=f"mass[{l}]")
    assert obj.N_pin == ref.N_pin and np.array_equal(obj.pin_gpu.cpu().numpy(), ref.pin)


def test_force_energy_hessian(pair):
    obj, ref, xp = pair
    x = torch.from_numpy(xp).to(obj.device)
    out = torch.zeros_like(x)
    obj._force(x, out)
    r64 = ref.ref64
    x64 = xp.astype(np.float64)
    g = np.zeros_like(xp); ref.force(xp, g)
    g64 = np.zeros_like(x64); r64.force(x64, g64)
    close64(out, g, g64, what="force (index order)")
    obj._force(x, out, ordered=False, gravity=False, pins=False)
    g = np.zeros_like(xp)
    K.elastic_force(xp, ref.hexs[0], None, ref.sfg, ref.det[0], ref.inv[0], ref.mu, ref.la, g)
    g64 = np.zeros_like(x64)
    K.elastic_force(x64, ref.hexs[0], None, r64.sfg, r64.det[0], r64.inv[0], ref.mu, ref.la, g64, np.float64)
    close64(out, g, g64, what="elastic force (vertex order)")
    e = obj.total_energy(x)
    e_ref = float(ref.total_energy(xp))
    assert abs(e - e_ref) <= 2e-5 * abs(e_ref)
    vals = obj.new_values(0)
    obj._hessian(0, x, vals, obj.diag_fixed)
    v = ref.MF_value_fixed.copy(); ref.hessian(0, xp, v)
    v64 = r64.MF_value_fixed.copy(); r64.hessian(0, x64, v64)
    close64(vals, v, v64, what="hessian [L;D;U] values")
    f2 = torch.zeros_like(x); vals2 = obj.new_values(0)
    obj._force_hessian(x, f2, vals2, obj.diag_fixed)
    g = np.zeros_like(xp); ref.force(xp, g)
    g64 = np.zeros_like(x64); r64.force(x64, g64)
    close64(f2, g, g64, what="fused force"); close64(vals2, v, v64, what="fused hessian")
    # symmetry H_ab = H_ba^T through the pattern
    if obj.layer > 1:
        x1 = torch.from_numpy(ref.x_gpu_layer[1] * np.float32(1.01)).to(obj.device)
        v1 = obj.new_values(1)
        obj._hessian(1, x1, v1)
        r1 = np.zeros_like(ref.UtAUs_values[0]); ref.hessian(1, ref.x_gpu_layer[1] * np.float32(1.01), r1)
        close(v1, r1, what="coarse re-discretised hessian")


@pytest.mark.parametrize("name", ["cube9", "jag3", "box", "odd"])
def test_lattice_fold_matches_fold_plan_path(name):
    """femb_hessian / femb_force_hessian on the lattice fold (plane-major element matrices, one thread per row component,
    compile-time offsets) vs the general fold-plan path on the same level: same values up to the summation order (incident
    hexahedra by local node instead of by element number); ragged meshes, odd hexahedron counts, every level."""
    pts, hx = MESHES[name]() if name in MESHES else (mesh.box_mesh(15, 6, 23, 0.05) if name == "box" else mesh.box_mesh(5, 5, 5, 0.05))
    obj = Object((pts, hx), 0.05, [], use_graphs=False)
    assert obj.levels[0].lat_row_rec and obj.levels[0].lat_inc8, "lattice fold missing on a voxel mesh"
    if name == "odd":
        assert hx.shape[0] % 2 == 1
    rng = np.random.default_rng(11)
    x = obj.x_gpu_layer[0] + torch.from_numpy(rng.normal(size=pts.shape).astype(np.float32) * 0.003).to(obj.device)
    for trial in range(2):                                   # twice: the kernels must leave nothing behind
        va, vb = obj.new_values(0), obj.new_values(0)
        fa, fb = torch.zeros_like(x), torch.zeros_like(x)
        va.fill_(7.0); fa.fill_(7.0)                         # every slot must be written, not accumulated
        obj._force_hessian(x, fa, va, obj.diag_fixed)
        obj._force_hessian(x, fb, vb, obj.diag_fixed, level=obj.level_without_lattice(0))
        close(va, vb.cpu().numpy(), rel=2e-6, what=f"{name}: lattice fold vs fold plan values")
        assert torch.equal(fa, fb), f"{name}: forces differ between the two layouts"
        vc = obj.new_values(0); vc.fill_(-3.0)
        obj._hessian(0, x, vc)
        vd = obj.new_values(0)
        obj._hessian(0, x, vd, level=obj.level_without_lattice(0))
        close(vc, vd.cpu().numpy(), rel=2e-6, what=f"{name}: lattice fold vs fold plan values (Hessian only, no diagonal add)")
    for l in range(1, obj.layer):
        assert obj.levels[l].lat_row_rec, f"lattice fold missing on level {l}"
        xl = obj.x_gpu_layer[l] * 1.01
        v1, v2 = obj.new_values(l), obj.new_values(l)
        obj._hessian(l, xl, v1)
        obj._hessian(l, xl, v2, level=obj.level_without_lattice(l))
        close(v1, v2.cpu().numpy(), rel=2e-6, what=f"{name}: level {l} lattice fold vs fold plan")


@pytest.mark.parametrize("name", ["cube9", "jag3", "box", "odd"])
def test_assembly_async_modes_are_bit_identical(name):
    """femb_set_assembly_async: the element kernel's TMA tensor store (swizzled shared-memory tile -> blocked plane-major
    scratch), the bulk-copy staging of the rest data and the L2 prefetches leave the same numbers as the store loop /
    per-lane loads: values and forces bit-identical, on every level, with and without forces, on box and ragged meshes."""
    pts, hx = MESHES[name]() if name in MESHES else (mesh.box_mesh(15, 6, 23, 0.05) if name == "box" else mesh.box_mesh(5, 5, 5, 0.05))
    obj = Object((pts, hx), 0.05, [], use_graphs=False)
    rng = np.random.default_rng(5)
    x = obj.x_gpu_layer[0] + torch.from_numpy(rng.normal(size=pts.shape).astype(np.float32) * 0.003).to(obj.device)
    out = {}
    try:
        for variant in (0, 1, 2, 3):
            obj.lib.femb_set_assembly_async(variant)
            for trial in range(2):
                v = obj.new_values(0); v.fill_(5.0)
                f = torch.zeros_like(x); f.fill_(5.0)
                obj._force_hessian(x, f, v, obj.diag_fixed)
                vh = obj.new_values(0); vh.fill_(-2.0)
                obj._hessian(0, x, vh)
                ff = torch.zeros_like(x); ff.fill_(4.0)
                obj._force(x, ff)
                coarse = [ff]
                for l in range(1, obj.layer):
                    vl = obj.new_values(l); vl.fill_(9.0)
                    obj._hessian(l, obj.x_gpu_layer[l] * 1.01, vl)
                    coarse.append(vl)
                torch.cuda.synchronize()
                out[variant] = (v, f, vh, coarse)
    finally:
        obj.lib.femb_set_assembly_async(3)
    base = out[0]
    assert torch.isfinite(base[0]).all()
    for variant in (1, 2, 3):
        got = out[variant]
        assert torch.equal(got[0], base[0]), f"{name}: values differ in assembly async mode {variant}"
        assert torch.equal(got[1], base[1]), f"{name}: forces differ in assembly async mode {variant}"
        assert torch.equal(got[2], base[2]), f"{name}: Hessian-only values differ in assembly async mode {variant}"
        for a, b in zip(got[3], base[3]):
            assert torch.equal(a, b), f"{name}: coarse values differ in assembly async mode {variant}"


def _load_state(obj, ref, xp):
    """Same matrix + rhs on both sides (built by the oracle, uploaded to the device)."""
    ref.MF_value[...] = ref.MF_value_fixed; ref.hessian(0, xp, ref.MF_value)
    ref.grad[...] = 0; ref.force(xp, ref.grad)
    ref.set_LUD(0)
    obj.MF_value_gpu.copy_(torch.from_numpy(ref.MF_value))
    ref.dev_B_fixed[0][...] = ref.grad
    obj.dev_B_fixed[0].copy_(torch.from_numpy(ref.grad))


def test_whole_step_exports_match_operator_sequences():
    """femb_newton_mg_step / femb_fas_step(1..3) / femb_vcycle / femb_dynamic_step / femb_cloth_step (csrc/femb_steps.cu), driven
    through ctypes from a context struct built HERE, vs the same step issued operator by operator by the Python drivers:
    identical kernels in identical order => identical bits."""
    from fem_simulation_b200 import Cloth, DynamicObject
    pts, hx = MESHES["jag3"]()
    a = Object((pts, hx), 0.05, [], use_graphs=False, use_c_steps=False)
    b = Object((pts, hx), 0.05, [], use_graphs=False, use_c_steps=False)
    assert a.layer == 3
    b._ensure_plans()
    st = lambda o: C.c_void_p(torch.cuda.current_stream(o.device).cuda_stream)
    for _ in range(2):
        a._newton_mg_body()
        _lib.check(b.lib.femb_newton_mg_step(C.byref(b.ctx), st(b)), "femb_newton_mg_step")
    assert torch.equal(a.x_gpu_layer[0], b.x_gpu_layer[0]) and torch.equal(a.UtAUs_values[1], b.UtAUs_values[1])
    for variant, args in ((2, (False, True, True, False)), (3, (False, True, True, True))):
        a._fas_v123_body(a.x_gpu_layer[0], *args)
        _lib.check(b.lib.femb_fas_step(C.byref(b.ctx), variant, None, st(b)), "femb_fas_step")
        assert torch.equal(a.x_gpu_layer[0], b.x_gpu_layer[0]), f"FASv{variant}"
    a.dev_x_solved[0].copy_(a.x_gpu_layer[0]); b.dev_x_solved[0].copy_(b.x_gpu_layer[0])
    a._fas_v123_body(a.dev_x_solved[0], True, False, False, False)
    _lib.check(b.lib.femb_fas_step(C.byref(b.ctx), 1, b.dev_x_solved[0].data_ptr(), st(b)), "femb_fas_step")
    assert torch.equal(a.dev_x_solved[0], b.dev_x_solved[0]), "FASv1"
    for o in (a, b):
        o.dev_B_fixed[0].normal_(generator=torch.Generator(device=o.device).manual_seed(5))
    a.VCycle(0)
    _lib.check(b.lib.femb_vcycle(C.byref(b.ctx), 0, st(b)), "femb_vcycle")
    assert torch.equal(a.dev_delta_x[0], b.dev_delta_x[0]) and torch.equal(a.dev_x_solved[0], b.dev_x_solved[0])
    da = DynamicObject((pts, hx), 0.05, [], use_graphs=False, use_c_steps=False)
    db = DynamicObject((pts, hx), 0.05, [], use_graphs=True, use_c_steps=True)
    ca, cb = Cloth(use_graphs=False, use_c_steps=False), Cloth(use_graphs=True, use_c_steps=True)
    for _ in range(3):
        da.step(); db.step(); ca.step(); cb.step()
    assert torch.equal(da.dev_x, db.dev_x) and torch.equal(da.dev_v, db.dev_v)
    assert torch.equal(ca.dev_x, cb.dev_x) and torch.equal(ca.dev_v, cb.dev_v)


def test_bsr_views_match_oracle_bsrs(pair):
    """The drop-in A / L / U / D attributes (object.py:291-311: BsrMatrix objects rebuilt by bsr_set_from_triplets every
    step) as views of the fixed [L;D;U] layout: .nrow/.ncol/.nnz/.offsets/.columns bit-exact and .values identical to the
    oracle's BSRs built from the same value array, every level; the transfer operators likewise."""
    from oracle.ref_setup import bsr_from_triplets
    obj, ref, xp = pair
    _load_state(obj, ref, xp)
    for l in range(obj.layer):
        if l > 0:
            vals = np.random.default_rng(l).normal(size=ref.values(l).shape).astype(np.float32)
            ref.values(l)[...] = vals
            obj._values(l).copy_(torch.from_numpy(vals))
        ref.set_LUD(l)
        p = ref.pat[l]
        n = ref.dims[l]
        full = bsr_from_triplets(n, n, np.concatenate([p["L_row"], p["D_row"], p["U_row"]]),
                                 np.concatenate([p["L_col"], p["D_col"], p["U_col"]]), ref.values(l))       # object.py:1254
        for mine, theirs, nm in ((obj.L[l], ref.Lm[l], "L"), (obj.U[l], ref.Um[l], "U"), (obj.D[l], ref.Dm[l], "D"), (obj.A[l], full, "A")):
            assert (mine.nrow, mine.ncol, mine.nnz) == (theirs["nrow"], theirs["ncol"], theirs["nnz"]), f"{nm}[{l}] shape"
            assert np.array_equal(mine.offsets.cpu().numpy(), theirs["offsets"]), f"{nm}[{l}].offsets"
            assert np.array_equal(mine.columns.cpu().numpy(), theirs["columns"]), f"{nm}[{l}].columns"
            assert np.array_equal(mine.values.cpu().numpy().reshape(-1, 3, 3), theirs["values"].reshape(-1, 3, 3)), f"{nm}[{l}].values"
    for name in ("Ut", "Us", "Ut_noOrder", "Us_noOrder", "Ut_hat", "Us_hat", "Ut_noOrder_hat", "Us_noOrder_hat"):
        for l in range(obj.layer - 1):
            mine, theirs = getattr(obj, name)[l], getattr(ref, name)[l]
            assert (mine.nrow, mine.ncol, mine.nnz) == (theirs["nrow"], theirs["ncol"], theirs["nnz"]), f"{name}[{l}] shape"
            assert np.array_equal(mine.offsets.cpu().numpy(), theirs["offsets"]) and np.array_equal(mine.columns.cpu().numpy(), theirs["columns"])
            tv = theirs["values"] if theirs["values"].ndim == 3 else theirs["values"][:, None, None] * np.eye(3, dtype=np.float32)
            assert np.array_equal(mine.values.cpu().numpy(), tv), f"{name}[{l}].values"


def test_spmv_parts(pair):
    obj, ref, xp = pair
    _load_state(obj, ref, xp)
    rng = np.random.default_rng(3)
    v = rng.normal(size=xp.shape).astype(np.float32)
    y0 = rng.normal(size=xp.shape).astype(np.float32)
    xv = torch.from_numpy(v).to(obj.device)
    for parts, mats in ((_lib.PART_L, [ref.Lm]), (_lib.PART_D, [ref.Dm]), (_lib.PART_U, [ref.Um]),
                        (_lib.PART_ALL, [ref.Lm, ref.Um, ref.Dm])):
        y = torch.from_numpy(y0.copy()).to(obj.device)
        obj._spmv(0, parts, -0.5, xv, 2.0, y)
        acc = np.zeros_like(v)
        for m in mats:
            ref_solvers.bsr_mv(m[0], v, acc, alpha=1.0, beta=1.0)
        close(y, -0.5 * acc + 2.0 * y0, what=f"spmv parts={parts}")
    r = torch.zeros_like(xv)
    obj._residual(0, obj.dev_B_fixed[0], xv, r)
    acc = np.zeros_like(v)
    for m in (ref.Lm, ref.Um, ref.Dm):
        ref_solvers.bsr_mv(m[0], v, acc, alpha=1.0, beta=1.0)
    close(r, ref.grad - acc, what="residual")


@pytest.mark.parametrize("iters", [1, 3])
def test_gauss_seidel(pair, iters):
    obj, ref, xp = pair
    _load_state(obj, ref, xp)
    obj.PerformGaussSeidel(0, iters)
    ref.PerformGaussSeidel(0, iters)
    close(obj.dev_delta_x[0], ref.dev_delta_x[0], what=f"GS x{iters}")


def test_jacobi_and_cg(pair):
    obj, ref, xp = pair
    _load_state(obj, ref, xp)
    obj.PerformJacobi(0, 1); ref.PerformJacobi(0, 1)
    close(obj.dev_delta_x[0], ref.dev_delta_x[0], what="Jacobi x1")
    obj.PerformJacobi(0, 3); ref.PerformJacobi(0, 3)
    close(obj.dev_delta_x[0], ref.dev_delta_x[0], what="Jacobi x3")
    obj.PerformConjugateGradient(0, 5); ref.PerformConjugateGradient(0, 5)
    close(obj.dev_delta_x[0], ref.dev_delta_x[0], rel=5e-5, what="CG(5)")
    obj.PerformConjugateGradient_WithInitX(0, 4); ref.PerformConjugateGradient_WithInitX(0, 4)
    close(obj.dev_delta_x[0], ref.dev_delta_x[0], rel=5e-5, what="CG with init x")
    # zero right-hand side: r1 < EPSILON exit leaves dx = 0 (object.py:944-945)
    obj.dev_B_fixed[0].zero_()
    obj.PerformConjugateGradient(0, 5)
    assert float(obj.dev_delta_x[0].abs().max()) == 0.0


def test_transfers(pair):
    obj, ref, xp = pair
    if obj.layer < 2:
        pytest.skip("single level")
    rng = np.random.default_rng(5)
    for name in ("Ut", "Ut_hat", "Ut_noOrder", "Ut_noOrder_hat", "Us", "Us_hat", "Us_noOrder", "Us_noOrder_hat"):
        T = getattr(obj, name)[0]
        R = getattr(ref, name)[0]
        v = rng.normal(size=(T.ncol, 3)).astype(np.float32)
        y0 = rng.normal(size=(T.nrow, 3)).astype(np.float32)
        y = torch.from_numpy(y0.copy()).to(obj.device)
        obj._transfer(T, 1.0, torch.from_numpy(v).to(obj.device), 1.0, y)
        yr = y0.copy(); ref_solvers.bsr_mv(R, v, yr, alpha=1.0, beta=1.0)
        close(y, yr, what=name)
        assert T.nnz == R["nnz"] and np.array_equal(T.columns.cpu().numpy(), R["columns"])
        close(T.values, R["values"][:, None, None] * np.eye(3, dtype=np.float32), rel=0, what=name + ".values")


def test_galerkin_and_spd(pair):
    obj, ref, xp = pair
    if obj.layer < 2:
        pytest.skip("single level")
    _load_state(obj, ref, xp)
    ref.galerkin()
    obj.galerkin()
    for l in range(1, obj.layer):
        close(obj.UtAUs_values[l - 1], ref.UtAUs_values[l - 1], rel=2e-5, what=f"Galerkin level {l}")
    blocks = ref.MF_value[:4096].copy()
    b = torch.from_numpy(blocks.copy()).to(obj.device)
    obj._spd(b)
    close(b, K.spd_matrix33f(blocks, 1e-3), rel=2e-5, what="spd_matrix33f")


def test_vcycle(pair):
    obj, ref, xp = pair
    _load_state(obj, ref, xp)
    ref.galerkin(); obj.galerkin()
    for l in range(1, obj.layer):          # identical coarse operators on both sides
        obj.UtAUs_values[l - 1].copy_(torch.from_numpy(ref.UtAUs_values[l - 1]))
    ref.VCycle(0); ref.finish()
    obj.VCycle(0); obj.finish()
    close(obj.dev_x_solved[0], ref.dev_x_solved[0], rel=5e-5, what="V-cycle correction")


@pytest.mark.parametrize("method,steps", [("NewtonMultigrid", 3), ("Newton", 3), ("FASv3", 4), ("FASv2", 3), ("FASv1", 2), ("FASv0", 2)])
@pytest.mark.parametrize("graphs", [False, True])
def test_nonlinear_steps(method, steps, graphs):
    pts, hx = mesh.cube_mesh(9, 0.05)
    obj = Object((pts, hx), 0.05, [], use_graphs=graphs)
    ref = ref_solvers.RefObject(pts, hx, 0.05, [])
    getattr(obj, method)(steps)
    getattr(ref, method)(steps)
    xa = obj.dev_x_solved[0] if method in ("FASv0", "FASv1") else obj.x_gpu_layer[0]
    xb = ref.dev_x_solved[0] if method in ("FASv0", "FASv1") else ref.x_gpu_layer[0]
    close(xa, xb, what=f"{method} positions after {steps} steps")
    # displacement (what the solver actually produced) to 1e-3 of its own scale
    da = xa.cpu().numpy() - ref.S["x0"]; db = xb - ref.S["x0"]
    close(da, db, rel=2e-3, what=f"{method} displacement")


def test_pinlist_and_three_levels():
    pts, hx = mesh.jagged_mesh(18, 17, 19, 0.05, seed=2, keep=0.7, origin=(-0.3, 0.1, 0.2))
    obj = Object((pts, hx), 0.05, [100], use_graphs=True)
    ref = ref_solvers.RefObject(pts, hx, 0.05, [100])
    assert obj.layer == 3 and obj.N_pin == 1
    obj.NewtonMultigrid(3); ref.NewtonMultigrid(3)
    close(obj.x_gpu_layer[0], ref.x_gpu_layer[0], what="3-level Newton-MG, displaced pin")
    close(obj.x_gpu_layer[0].cpu().numpy() - ref.S["x0"], ref.x_gpu_layer[0] - ref.S["x0"], rel=2e-3, what="displacement")


def test_convergence_curve():
    """compare(100): the sampled energy / |f|inf curves of both loops (object.py:1514-1598) against the golden
    oracle curves (tests/golden/compare_cube9.npz, float32 and float64 twins of the oracle).
    Energies and final positions: 1e-5 relative.  Force norms: the iteration is chaotic at round-off level (the
    1e-10 cut-offs of solve3x3 switch rows on and off), so the float32 oracle itself drifts from the float64 one
    (up to O(1) late in the FAS loop); the band is max(1e-3, 10 x running max of that oracle gap)."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "compare_cube9.npz"))
    pts, hx = mesh.cube_mesh(9, 0.05)
    obj = Object((pts, hx), 0.05, [], use_graphs=True)
    obj.compare(100)
    assert obj.plot_x == gold["plot_x_f64"].tolist() and len(obj.plot_x) == 100
    for mine, key in ((obj.plot_energy_newton, "energy_newton"), (obj.plot_energy_newtonMultigrid, "energy_fas")):
        close(np.array(mine), gold[key + "_f64"], rel=1e-5, what=key)
    for mine, key in ((obj.plot_InfNorm_newton, "lognorm_newton"), (obj.plot_InfNorm_newtonMultigrid, "lognorm_fas")):
        f, f32, f64 = 10.0 ** np.array(mine), 10.0 ** gold[key + "_f32"], 10.0 ** gold[key + "_f64"]
        band = np.maximum(1e-3, 10.0 * np.maximum.accumulate(np.abs(f32 - f64) / f64))
        dev = np.abs(f - f64) / f64
        assert np.all(dev <= band), f"{key}: worst dev/band {np.max(dev / band):.2f} at sample {int(np.argmax(dev / band))}"
        assert np.all(np.abs(f[:10] - f64[:10]) <= 1e-3 * f64[:10]), key + " (first 10 samples)"
    close(obj.x_gpu_layer[0], gold["x_final_f64"], rel=1e-5, what="final positions")


def test_empty_and_errors():
    lib = _lib.load()
    lv = _lib.FembLevel()
    assert lib.femb_spmv(C.byref(lv), None, 7, 1.0, None, 0.0, None, None) == -1
    assert b"invalid argument" in lib.femb_last_error()
    z = torch.zeros(3, device="cuda")
    assert lib.femb_axpby(0, 1.0, z.data_ptr(), 1.0, z.data_ptr(), None) == 0
    assert lib.femb_inf_norm(0, z.data_ptr(), z.data_ptr(), None) == 0
    assert lib.femb_launch_count() > 0


def test_pdl_gauss_seidel_is_bit_identical():
    """The colour launches of femb_gauss_seidel are programmatic dependent launches (matrix staging overlaps the
    previous colour's tail).  Same arithmetic => identical bits with femb_set_pdl(0), eagerly and in a CUDA graph,
    on a mesh with more CTAs per colour than fit on the device at once."""
    n = 56
    obj = Object(f"cube:{n}", 1.0 / n, [], use_graphs=False, layers=2)
    x = obj.x_gpu_layer[0]
    gen = torch.Generator(device="cpu").manual_seed(3)
    x.add_((torch.randn(x.shape, generator=gen) * 2e-4).to(x.device))
    obj._force_hessian(x, obj.dev_B_fixed[0], obj.MF_value_gpu, obj.diag_fixed)
    old = obj.lib.femb_set_pdl(0)
    try:
        obj.PerformGaussSeidel(0, 2)
        ref = obj.dev_delta_x[0].clone()
    finally:
        obj.lib.femb_set_pdl(old)
    for rep in range(3):
        obj.dev_delta_x[0].fill_(float("nan"))
        obj.PerformGaussSeidel(0, 2)
        torch.cuda.synchronize()
        assert torch.equal(ref, obj.dev_delta_x[0]), f"eager PDL sweep differs (rep {rep})"
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        obj.PerformGaussSeidel(0, 2)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            obj.PerformGaussSeidel(0, 2)
    for rep in range(3):
        obj.dev_delta_x[0].fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(ref, obj.dev_delta_x[0]), f"graph-replayed PDL sweep differs (rep {rep})"


def _fan_mesh(dx=0.05):
    """8 unit cubes around the origin that share ONLY vertex 0 (every other corner is a private copy): not a lattice
    mesh -- row 0 has 57 blocks, and its diagonal block 8 contributors, every other row 8 blocks."""
    pts, hexs = [(0.0, 0.0, 0.0)], []
    for ox in (-1, 0):
        for oy in (-1, 0):
            for oz in (-1, 0):
                h = []
                for bx in (0, 1):
                    for by in (0, 1):
                        for bz in (0, 1):
                            p = ((ox + bx) * dx, (oy + by) * dx, (oz + bz) * dx)
                            if p == (0.0, 0.0, 0.0):
                                h.append(0)
                            else:
                                pts.append(p); h.append(len(pts) - 1)
                hexs.append(h)
    return np.array(pts, np.float32), np.array(hexs, np.int32)


def test_rows_with_more_than_27_blocks():
    """Generality beyond hexahedral lattices (the fast paths assume <= 27 blocks per row): the fold's slow path, the row
    pass with 56 off-diagonal blocks in one row, and Gauss-Seidel / SpMV on that matrix, against a direct numpy scatter of the
    oracle's element matrices through the element->nnz map."""
    pts, hx = _fan_mesh()
    obj = Object((pts, hx), 0.05, [3], layers=1, use_graphs=False, pin_displacement=0.0)
    assert not obj.levels[0].lat_row_rec, "femb_build_lattice_fold must reject a mesh that is not a lattice subset"
    p = obj.hierarchy.pattern[0]
    assert int(np.diff(p.L_ptr).max() + np.diff(p.U_ptr).max()) >= 56
    rng = np.random.default_rng(5)
    xp = (pts + rng.normal(size=pts.shape).astype(np.float32) * 0.002).astype(np.float32)
    sfg = obj.hierarchy.shapeFuncGrad
    det, inv = K.prepare_kernal(pts, hx, sfg)
    ref = np.zeros((p.nnz, 3, 3), np.float32)
    K.elastic_hessian(xp, hx, sfg, det, inv, obj.mu, obj.la, ref, p.hex_update_offset)
    ref64 = np.zeros((p.nnz, 3, 3), np.float64)
    det64, inv64 = K.prepare_kernal(pts.astype(np.float64), hx, sfg, dtype=np.float64)
    K.elastic_hessian(xp.astype(np.float64), hx, sfg, det64, inv64, obj.mu, obj.la, ref64, p.hex_update_offset, dtype=np.float64)
    x = torch.from_numpy(xp).to(obj.device)
    vals = obj.new_values(0)
    obj._hessian(0, x, vals, None)
    close64(vals, ref, ref64, what="values of the fan mesh (slow fold path)")
    # y = A x with the 57-block row, and one symmetric Gauss-Seidel iteration, against dense algebra
    n = obj.N_verts
    A = np.zeros((n * 3, n * 3))
    v = vals.cpu().numpy().astype(np.float64)
    rows, cols = p.row, p.col
    for s in range(p.nnz):
        A[3 * rows[s]:3 * rows[s] + 3, 3 * cols[s]:3 * cols[s] + 3] += v[s]
    xv = rng.normal(size=(n, 3)).astype(np.float32)
    y = torch.zeros(n, 3, device=obj.device)
    obj.MF_value_gpu.copy_(vals)
    obj._spmv(0, _lib.PART_ALL, 1.0, torch.from_numpy(xv).to(obj.device), 0.0, y)
    close(y, (A @ xv.reshape(-1).astype(np.float64)).reshape(n, 3), rel=2e-5, what="SpMV with a 57-block row")


@pytest.mark.parametrize("rows,subs", [(8, 4), (8, 8), (16, 4), (16, 8), (32, 4), (32, 8)])
def test_row_pass_configurations(pair, rows, subs):
    """Every rows-per-CTA x threads-per-row instantiation of the row-pass kernels (Gauss-Seidel colour launches as
    programmatic dependent launches; residual) against the oracle."""
    obj, ref, xp = pair
    _load_state(obj, ref, xp)
    ref.PerformGaussSeidel(0, 2)
    old_r, old_s = obj.lib.femb_set_gs_rows_per_cta(rows), obj.lib.femb_set_gs_threads_per_row(subs)
    try:
        if True:
            obj.dev_delta_x[0].fill_(float("nan"))
            obj.PerformGaussSeidel(0, 2)
            close(obj.dev_delta_x[0], ref.dev_delta_x[0], what=f"GS x2 rows={rows} subs={subs}")
        r = torch.zeros_like(obj.dev_delta_x[0])
        obj._residual(0, obj.dev_B_fixed[0], obj.dev_delta_x[0], r)
        acc = np.zeros_like(ref.dev_delta_x[0])
        for m in (ref.Lm, ref.Um, ref.Dm):
            ref_solvers.bsr_mv(m[0], ref.dev_delta_x[0], acc, alpha=1.0, beta=1.0)
        close(r, ref.grad - acc, what=f"residual rows={rows} subs={subs}")
    finally:
        obj.lib.femb_set_gs_rows_per_cta(old_r)
        obj.lib.femb_set_gs_threads_per_row(old_s)


def test_corrected_hierarchy_newton_multigrid():
    """`Object(..., corrected_hierarchy=True)` (level l coarsened with spacing dx 2^l, so the 4th level really is
    coarser) against the oracle with the same one-line fix, two Newton-multigrid steps on a sparse ragged 4-level mesh."""
    pts, hx = mesh.jagged_mesh(33, 32, 34, 0.05, seed=4, keep=0.1)
    obj = Object((pts, hx), 0.05, [], use_graphs=False, corrected_hierarchy=True)
    ref = ref_solvers.RefObject(pts, hx, 0.05, [], corrected_spacing=True)
    assert obj.layer == 4 and obj.dims == ref.dims and obj.dims[3] < obj.dims[2]
    obj.NewtonMultigrid(2)
    ref.NewtonMultigrid(2)
    close(obj.x_gpu_layer[0], ref.x_gpu_layer[0], what="positions after 2 Newton-MG steps on the corrected hierarchy")
    close(obj.UtAUs_values[2], ref.UtAUs_values[2], rel=2e-5, what="Galerkin operator of the corrected level 3")


def test_fas_multilevel():
    """SURVEY 8f-3, second half: FAS with the tau correction on more than two levels (Object.FAS_multilevel, the
    generalisation of the reference's FASv0) -- with levels=2 it is FASv0 itself, bit for bit; on a 4-level corrected hierarchy
    it matches its oracle twin, and the deeper cycle reduces the force more per iteration than the two-level one."""
    pts, hx = mesh.cube_mesh(9, 0.05)
    a = Object((pts, hx), 0.05, [], use_graphs=True)
    b = Object((pts, hx), 0.05, [], use_graphs=True)
    a.FASv0(3); b.FAS_multilevel(3, levels=2)
    assert torch.equal(a.dev_x_solved[0], b.dev_x_solved[0])
    pts, hx = mesh.jagged_mesh(33, 32, 34, 0.05, seed=4, keep=0.1)
    obj = Object((pts, hx), 0.05, [], use_graphs=True, corrected_hierarchy=True)
    ref = ref_solvers.RefObject(pts, hx, 0.05, [], corrected_spacing=True)
    assert obj.layer == 4
    rng = np.random.default_rng(8)
    start = (ref.S["x0"] + rng.normal(size=pts.shape).astype(np.float32) * np.float32(0.002)).astype(np.float32)   # FASv0 has no load: strain the start state
    for levels, smoothing in ((3, 1), (4, 2)):
        obj.x_gpu_layer[0].copy_(torch.from_numpy(start)); ref.x_gpu_layer[0][...] = start
        obj.FAS_multilevel(2, levels=levels, smoothing=smoothing)
        ref.FAS_multilevel(2, levels=levels, smoothing=smoothing)
        close(obj.dev_x_solved[0], ref.dev_x_solved[0], what=f"FAS_multilevel({levels} levels) positions after 2 cycles")
        da = obj.dev_x_solved[0].cpu().numpy() - start; db = ref.dev_x_solved[0] - start
        close(da, db, rel=2e-3, what=f"FAS_multilevel({levels} levels) displacement")
        close(obj.UtAUs_values[levels - 2], ref.UtAUs_values[levels - 2], rel=2e-5, what="re-discretised Hessian of the coarsest level")


def test_exp3_hooks_on_device(pair):
    """`edge_index` and `torch_A` (what the exp3 GNN drivers read, exp3/simulation/object.py:258-276, 438-441) live on the
    device; torch_A[0] is the un-normalised prolongation in vertex order, so it must agree with femb_transfer_mv."""
    obj, ref, _ = pair
    ei = obj.edge_index
    assert len(ei) == obj.layer and all(e.device.type == "cuda" and e.shape[0] == 2 and e.dtype == torch.int64 for e in ei)
    assert bool((ei[0][0] < ei[0][1]).all())
    tA = obj.torch_A
    assert len(tA) == obj.layer - 1 and tuple(tA[0].shape) == (obj.dims[0], obj.dims[1])
    rng = np.random.default_rng(2)
    xc = torch.from_numpy(rng.normal(size=(obj.dims[1], 3)).astype(np.float32)).to(obj.device)
    y = torch.zeros(obj.dims[0], 3, device=obj.device)
    obj._transfer(obj.Us_noOrder_hat[0], 1.0, xc, 0.0, y)
    close(tA[0] @ xc, y.cpu().numpy(), what="torch_A[0] @ x vs femb_transfer_mv(Us_noOrder_hat[0])")


def test_compare_without_gravity_and_drag():
    """The remaining diagnostic drivers of the reference API (object.py:1630-1741 compare_withoutGravity, :1770-1852 drag;
    plots dropped) against the oracle twin, with one displaced pinned vertex as in exp1/quasi_simulation/main.py."""
    pts, hx = mesh.cube_mesh(9, 0.05)
    obj = Object((pts, hx), 0.05, [444], use_graphs=True)
    ref = ref_solvers.RefObject(pts, hx, 0.05, [444])
    d, r = obj.drag(5), ref.drag(5)
    for key in ("GS", "CG", "NewtonMG"):
        a, b = np.asarray(d[key]), np.asarray(r[key], dtype=np.float64)
        live = b > -4.0                                        # below that the residual is at fp32 round-off of the rhs
        assert live[0] and np.abs(a[live] - b[live]).max() < 2e-3, (key, a, b)
        assert np.all(a[~live] < -3.5)
    obj.x_gpu_layer[0].copy_(obj.x_cpu)
    ref.x_gpu_layer[0][...] = ref.x_cpu
    obj.compare_withoutGravity(20)
    ref.compare_withoutGravity(20)
    assert obj.plot_x == ref.plot_x == list(range(2, 21, 2))
    for name in ("plot_InfNorm_newton", "plot_InfNorm_newtonMultigrid"):
        assert np.abs(np.asarray(getattr(obj, name)) - np.asarray(getattr(ref, name), dtype=np.float64)).max() < 2e-3, name
    for name in ("plot_energy_newton", "plot_energy_newtonMultigrid"):
        close(np.asarray(getattr(obj, name), dtype=np.float64), np.asarray(getattr(ref, name), dtype=np.float64), rel=2e-5, what=name)
    close(obj.x_gpu_layer[0], ref.x_gpu_layer[0], what="positions after compare_withoutGravity")


def test_adam_update_operator_and_driver(pair):
    """femb_adam_update (updateM + updateV + scal_ x2 + updateX, cublas.py:172-194) against the numpy formula on random
    data (the driver itself is chaotic where the force vanishes: m_hat / sqrt(v_hat) is a sign), and `Object.Adam` filling the
    sampling lists like the reference (object.py:1067-1094)."""
    obj, ref, xp = pair
    rng = np.random.default_rng(9)
    n = 3000
    g, m, v, x = (rng.normal(size=n).astype(np.float32) for _ in range(4))
    v = np.abs(v)
    b1, b2, lr, eps, step = 0.9, 0.999, 1e-3, 1e-8, 3
    s1, s2 = np.float32(1 / (1 - b1 ** step)), np.float32(1 / (1 - b2 ** step))
    m2 = (np.float32(1 - b1) * g + np.float32(b1) * m).astype(np.float32)
    v2 = (np.float32(1 - b2) * (g * g) + np.float32(b2) * v).astype(np.float32)
    x2 = x - np.float32(lr) * ((m2 * s1) / (np.sqrt(v2 * s2) + np.float32(eps)))
    t = lambda a: torch.from_numpy(a.copy()).to(obj.device)
    tg, tm, tv, tx = t(g), t(m), t(v), t(x)
    tmh, tvh = torch.zeros_like(tm), torch.zeros_like(tv)
    _lib.check(obj.lib.femb_adam_update(n, tg.data_ptr(), tm.data_ptr(), tv.data_ptr(), tmh.data_ptr(), tvh.data_ptr(), tx.data_ptr(),
                                        b1, b2, float(s1), float(s2), lr, eps, None), "femb_adam_update")
    close(tm, m2, what="Adam m"); close(tv, v2, what="Adam v"); close(tmh, m2 * s1, what="Adam m_hat"); close(tx, x2, what="Adam x")
    x0 = obj.x_gpu_layer[0].clone()
    obj.plot_x.clear(); obj.plot_InfNorm.clear(); obj.plot_energy.clear()
    try:
        obj.Adam(200, lr=1e-5)
        assert obj.plot_x == list(range(2, 201, 2)) and len(obj.plot_energy) == 100
        moved = float((obj.x_gpu_layer[0] - x0).abs().max())
        assert 0.0 < moved <= 200 * 1e-5 * 10                 # an Adam step is a few lr at most
    finally:
        obj.x_gpu_layer[0].copy_(x0)
