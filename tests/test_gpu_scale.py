"""GPU parity at benchmark scale and on the BASELINE configurations ("what's weak" 1):

  * the n=64 cube (262 144 hexahedra, the kernels' large-grid paths; 5 levels by the reference formula): FASv3 and
    Newton-multigrid steps, CUDA vs the C oracle (oracle/femb_oracle.c, float32 and its float64 twin);
  * C1: `cube.obj` at dx = 0.05 (== the 19^3 lattice, tests/test_voxelize.py): `compare(1000)` -- 2000 steps --
    sampled energy / |f|inf curves vs the C oracle's curves;
  * C2 / C4: 300 frames of the dynamic hexahedral step and of the cloth step vs the numpy oracles.

Tolerances: positions / energies 1e-5 relative (north_star); block values 2e-5; where the float32 oracle itself is
further than that from its float64 twin (long chaotic runs: the 1e-10 cut-offs of solve3x3 switch rows on and off) the
bound is a multiple of that measured oracle gap, stated per assertion.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from fem_simulation_b200 import Object, hierarchy, mesh  # noqa: E402
from oracle import c_oracle  # noqa: E402


def rel(a, b):
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    return float(np.abs(a.astype(np.float64) - np.asarray(b, np.float64)).max() / max(float(np.abs(b).max()), 1e-30))


@pytest.fixture(scope="module")
def cube64():
    n = 64
    pts, hx = mesh.cube_mesh(n)
    obj = Object((pts, hx), 1.0 / n, [], use_graphs=True)
    H = obj.hierarchy
    assert H.layer == 5 and hx.shape[0] == 262144
    plans = [hierarchy.build_galerkin_plan(H, l) for l in range(H.layer - 1)]
    return n, pts, obj, H, plans


def test_cube64_fasv3_steps_match_c_oracle(cube64):
    n, pts, obj, H, plans = cube64
    co = c_oracle.from_hierarchy(H, plans, pts, 1.0 / n)
    co64 = c_oracle.from_hierarchy(H, plans, pts, 1.0 / n, dtype=np.float64)
    for c in (co, co64):
        c.set_threads(len(os.sched_getaffinity(0)))
    obj.x_gpu_layer[0].copy_(obj.x_cpu)
    for _ in range(3):
        obj.step(); co.fasv3_step(); co64.fasv3_step()
    x0 = obj.x_cpu.numpy()
    noise = np.abs(co.x.astype(np.float64) - co64.x).max()
    err = np.abs(obj.x_gpu_layer[0].cpu().numpy().astype(np.float64) - co64.x).max()
    assert rel(obj.x_gpu_layer[0], co.x) <= 1e-5
    assert err <= max(1e-5 * np.abs(co64.x - x0).max(), 4.0 * noise), (err, noise)      # displacements, not just positions
    # the matrices of the last step: fine Hessian + pin diagonal, re-discretised coarse Hessian + coarse pin penalty
    assert rel(obj.MF_value_gpu, co.values[0]) <= 2e-5
    assert rel(obj.UtAUs_values[0], co.values[1]) <= 2e-5


def test_cube64_newton_multigrid_steps_match_c_oracle(cube64):
    n, pts, obj, H, plans = cube64
    co = c_oracle.from_hierarchy(H, plans, pts, 1.0 / n)
    co64 = c_oracle.from_hierarchy(H, plans, pts, 1.0 / n, dtype=np.float64)
    for c in (co, co64):
        c.set_threads(len(os.sched_getaffinity(0)))
    obj.x_gpu_layer[0].copy_(obj.x_cpu)
    obj.NewtonMultigrid(2)
    for _ in range(2):
        co.newton_mg_step(); co64.newton_mg_step()
    x0 = obj.x_cpu.numpy()
    noise = np.abs(co.x.astype(np.float64) - co64.x).max()
    err = np.abs(obj.x_gpu_layer[0].cpu().numpy().astype(np.float64) - co64.x).max()
    assert rel(obj.x_gpu_layer[0], co.x) <= 1e-5
    assert err <= max(1e-5 * np.abs(co64.x - x0).max(), 4.0 * noise), (err, noise)
    assert rel(obj.MF_value_gpu, co.values[0]) <= 2e-5
    for l in range(1, H.layer):                                    # Galerkin chain 0 -> 1 -> 2 (-> the repeated levels), eig3 unpinned
        assert rel(obj.UtAUs_values[l - 1], co.values[l]) <= 2e-5, f"UtAUs_values[{l - 1}]"


def test_cube64_lattice_fold(cube64):
    """The lattice fold (csrc/femb_element.cu) at a size with 8 k element CTAs, against the general fold-plan path."""
    n, pts, obj, H, plans = cube64
    o2 = Object((pts, mesh.cube_mesh(n)[1]), 1.0 / n, [], use_graphs=False, layers=2)
    assert o2.levels[0].lat_row_rec
    x = o2.x_gpu_layer[0] + torch.randn_like(o2.x_gpu_layer[0]) * 1e-4
    va, vb = o2.new_values(0), o2.new_values(0)
    fa, fb = torch.zeros_like(x), torch.zeros_like(x)
    o2._force_hessian(x, fa, va, o2.diag_fixed)
    o2._force_hessian(x, fb, vb, o2.diag_fixed, level=o2.level_without_lattice(0))
    assert rel(va, vb.cpu().numpy()) <= 2e-6 and torch.equal(fa, fb)


def test_c1_cube_obj_compare_1000():
    """BASELINE config 1: cube.obj @ dx = 0.05, compare(1000): 100 samples of each loop."""
    asset = "/root/reference/assets/objs/cube.obj"
    if os.path.exists(asset):
        from fem_simulation_b200 import voxelize
        pts, hx = voxelize.voxelize_obj(asset, 0.05)
    else:                                                            # the GPU box has no /root/reference: the same lattice
        pts, hx = mesh.cube_mesh(19, 0.05)
    assert hx.shape[0] == 6859
    obj = Object((pts, hx), 0.05, [], use_graphs=True)
    assert obj.layer == 3
    obj.compare(1000)
    H = obj.hierarchy
    c32 = c_oracle.from_hierarchy(H, [], pts, 0.05)
    c64 = c_oracle.from_hierarchy(H, [], pts, 0.05, dtype=np.float64)
    g32, g64 = c32.compare(1000), c64.compare(1000)
    assert obj.plot_x == g64["plot_x"] and len(obj.plot_x) == 100
    for mine, key in ((obj.plot_energy_newton, "energy_newton"), (obj.plot_energy_newtonMultigrid, "energy_fas")):
        assert rel(np.array(mine), np.array(g64[key])) <= 1e-5, key
    for mine, key in ((obj.plot_InfNorm_newton, "lognorm_newton"), (obj.plot_InfNorm_newtonMultigrid, "lognorm_fas")):
        f, f32, f64 = 10.0 ** np.array(mine), 10.0 ** np.array(g32[key]), 10.0 ** np.array(g64[key])
        gap = np.abs(f32 - f64) / f64
        band = np.maximum(1e-3, 10.0 * np.maximum.accumulate(gap))
        dev = np.abs(f - f64) / f64
        # Once the iteration has converged to float32 round-off (the FAS loop after ~200 steps: |f|inf ~ 4e-5 and flat) the
        # norm only fluctuates: the float32 and float64 oracles part by > 1 % from sample `floor` on, and no float32
        # implementation can follow the curve sample by sample from (about) there.  Before: the tight band.  From two
        # samples before: the level must agree within half a decade.
        floor = int(np.argmax(gap > 1e-2)) if np.any(gap > 1e-2) else len(gap)
        tight = max(floor - 2, 10)
        assert np.all(dev[:tight] <= band[:tight]), (f"{key}: worst dev/band {np.max(dev[:tight] / band[:tight]):.2f} at sample "
                                                     f"{int(np.argmax(dev[:tight] / band[:tight]))} (noise floor from sample {floor})")
        assert np.all(np.abs(np.log10(f[tight:] / f64[tight:])) <= 0.5), key + " (level at the round-off floor)"
        assert np.all(np.abs(f[:10] - f64[:10]) <= 1e-3 * f64[:10]), key + " (first 10 samples)"
    noise = np.abs(c32.x.astype(np.float64) - c64.x).max()
    err = np.abs(obj.x_gpu_layer[0].cpu().numpy().astype(np.float64) - c64.x).max()
    assert err <= max(1e-5 * np.abs(c64.x).max(), 4.0 * noise), (err, noise)


def test_c4_cloth_300_frames():
    from fem_simulation_b200 import Cloth
    from oracle.ref_cloth import RefCloth
    c = Cloth(use_graphs=True)
    r, r64 = RefCloth(), RefCloth(dtype=np.float64)
    for _ in range(300):
        c.step(); r.render(); r64.render()
    x = c.dev_x.cpu().numpy().astype(np.float64)
    noise = np.abs(r.dev_x.astype(np.float64) - r64.dev_x).max()
    disp = np.abs(r64.dev_x - r64.x_cpu).max()
    err = np.abs(x - r64.dev_x).max()
    assert disp > 0.01                                               # the cloth really moved
    assert err <= max(1e-5 * np.abs(r64.dev_x).max(), 4.0 * noise), (err, noise, disp)


def test_c2_dynamic_300_frames():
    from fem_simulation_b200 import DynamicObject
    from oracle.ref_dynamic import RefDynamicObject
    pts, hx = mesh.jagged_mesh(10, 9, 11, 0.05, seed=1)
    d = DynamicObject((pts, hx), 0.05, [])
    r, r64 = RefDynamicObject(pts, hx, 0.05, []), RefDynamicObject(pts, hx, 0.05, [], dtype=np.float64)
    for _ in range(300):
        d.step(); r.render(); r64.render()
    x = d.dev_x.cpu().numpy().astype(np.float64)
    noise = np.abs(r.dev_x.astype(np.float64) - r64.dev_x).max()
    disp = np.abs(r64.dev_x - r64.S["x0"]).max()
    err = np.abs(x - r64.dev_x).max()
    assert disp > 1e-3
    assert err <= max(1e-5 * np.abs(r64.dev_x).max(), 4.0 * noise), (err, noise, disp)


def test_c0_bundled_horse_mesh_steps_match_c_oracle():
    """BASELINE config 0: Newton-multigrid and FAS on the reference's bundled `horse.obj` voxelised at dx = 0.05 (a ragged,
    non-box voxel set of 4 033 hexahedra, 3 levels; fixture tests/golden/horse_0.05.npz made by make_horse_fixture.py), CUDA vs
    the C oracle in float32 and float64: positions, displacements, fine and coarse operators, and the sampled compare() curve."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "horse_0.05.npz"))
    pts, hx, dx = g["points"], g["hexagons"], float(g["dx"])
    obj = Object((pts, hx), dx, [], use_graphs=True)
    H = obj.hierarchy
    assert H.layer == 3 and hx.shape[0] == 4033
    assert all(obj.levels[l].lat_row_rec for l in range(H.layer)), "a voxelised OBJ is a lattice subset on every level"
    plans = [hierarchy.build_galerkin_plan(H, l) for l in range(H.layer - 1)]
    x0 = obj.x_cpu.numpy()
    for method in ("nmg", "fasv3"):
        co = c_oracle.from_hierarchy(H, plans, pts, dx)
        co64 = c_oracle.from_hierarchy(H, plans, pts, dx, dtype=np.float64)
        obj.x_gpu_layer[0].copy_(obj.x_cpu)
        for _ in range(3):
            if method == "nmg":
                obj.NewtonMultigrid(1); co.newton_mg_step(); co64.newton_mg_step()
            else:
                obj.step(); co.fasv3_step(); co64.fasv3_step()
        noise = np.abs(co.x.astype(np.float64) - co64.x).max()
        err = np.abs(obj.x_gpu_layer[0].cpu().numpy().astype(np.float64) - co64.x).max()
        assert rel(obj.x_gpu_layer[0], co.x) <= 1e-5, method
        assert err <= max(1e-5 * np.abs(co64.x - x0).max(), 4.0 * noise), (method, err, noise)
        assert rel(obj.MF_value_gpu, co.values[0]) <= 2e-5, method
        assert rel(obj.UtAUs_values[0], co.values[1]) <= 2e-5, method
    obj.x_gpu_layer[0].copy_(obj.x_cpu)
    obj.compare(100)
    c64 = c_oracle.from_hierarchy(H, [], pts, dx, dtype=np.float64)
    g64 = c64.compare(100)
    assert obj.plot_x == g64["plot_x"]
    for mine, key in ((obj.plot_energy_newton, "energy_newton"), (obj.plot_energy_newtonMultigrid, "energy_fas")):
        assert rel(np.array(mine), np.array(g64[key])) <= 1e-5, key
