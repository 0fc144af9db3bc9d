"""Golden vectors produced by the reference's own Python source (unmodified ``object.py`` / ``quai_hexagon.py`` /
``cublas.py``) executed under the Warp stand-in of oracle/warp_emulator (tests/golden/README.md).

CPU: pins the oracle (literal setup, numpy kernels/solvers) and the product's vectorised setup against them.
GPU (``-m gpu``): the CUDA path through ``Object`` against the same files."""
import os

import numpy as np
import pytest

from fem_simulation_b200 import hierarchy
from oracle import ref_solvers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [c for c in ("jag9", "jag9_pin", "jag17", "jag17_pin") if os.path.exists(os.path.join(GOLD, f"ref_{c}.npz"))]
INT_KEYS = ["MF_L_row", "MF_L_col", "MF_U_row", "MF_U_col", "MF_GS_U_Ptr", "MF_GS_L_Ptr", "MF_GS_U_row", "MF_GS_U_col",
            "MF_GS_L_row", "MF_GS_L_col", "fix_idx", "pin", "pin_list"]


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module", params=CASES)
def gold(request):
    g = dict(np.load(os.path.join(GOLD, f"ref_{request.param}.npz")))
    g["pins"] = [int(v) for v in g["pinList"]]
    return g


def test_setup_matches_reference_class(gold):
    g = gold
    ref = ref_solvers.RefObject(g["points"], g["hexagons"], float(g["dx"]), g["pins"])
    H = hierarchy.build_hierarchy(g["points"], g["hexagons"], float(g["dx"]))
    S = ref.S
    L = int(g["layer"])
    assert S["layer"] == L == H.layer and list(g["dims"]) == S["dims"] == H.dims
    for l in range(L):
        for key, mine, prod in ((f"hexs{l}", S["hexs"][l], H.hexs[l]), (f"v2i{l}", S["vertex2index"][l], H.vertex2index[l]),
                                (f"i2v{l}", S["index2vertex"][l], H.index2vertex[l]),
                                (f"cvn{l}", S["color_vertex_num"][l], H.color_vertex_num[l]),
                                (f"hex_update_offset{l}", S["pattern"][l]["hex_update_offset"], H.pattern[l].hex_update_offset)):
            assert np.array_equal(g[key], np.asarray(mine)), key
            assert np.array_equal(g[key], np.asarray(prod)), key + " (product)"
        if l == 0:      # self.x[0] aliases the tensor the pin displacement is applied to (object.py:167,775)
            assert np.array_equal(g["x0"], S["x0"])
        else:
            assert np.array_equal(g[f"x{l}"], S["x"][l]) and np.array_equal(g[f"x{l}"], H.x[l])
        assert rel(ref.det[l], g[f"det{l}"]) <= 1e-6 and rel(ref.inv[l], g[f"inv{l}"]) <= 1e-6 and rel(ref.m[l], g[f"m{l}"]) <= 1e-6
    for l in range(L - 1):
        assert np.array_equal(g[f"fine_coarse{l}"], S["fine_coarse"][l]) and np.array_equal(g[f"fine_coarse{l}"], H.fine_coarse[l])
        for nm in ("Ut", "Us", "Ut_noOrder", "Us_noOrder", "Ut_hat", "Us_hat", "Ut_noOrder_hat", "Us_noOrder_hat"):
            m, pm = S[nm][l], getattr(H, nm)[l]
            assert np.array_equal(g[f"{nm}{l}_offsets"], m["offsets"]) and np.array_equal(g[f"{nm}{l}_columns"], m["columns"]), nm
            assert np.array_equal(g[f"{nm}{l}_offsets"], pm.offsets) and np.array_equal(g[f"{nm}{l}_columns"], pm.columns), nm
            blocks = g[f"{nm}{l}_values"]                                   # reference stores w * I3 blocks
            assert np.array_equal(blocks, m["values"][:, None, None] * np.eye(3, dtype=np.float32)), nm + " values"
            assert np.array_equal(blocks[:, 0, 0], pm.weights), nm + " weights (product)"
        p = S["pattern"][l + 1]
        q = H.pattern[l + 1]
        for key, mine, prod in ((f"UtAUs_block_offset{l}", p["block_offset"], q.block_offset),
                                (f"UtAUs_GS_U_Ptr{l}", p["GS_U_Ptr"], q.GS_U_Ptr), (f"UtAUs_GS_L_Ptr{l}", p["GS_L_Ptr"], q.GS_L_Ptr),
                                (f"UtAUs_L_row{l}", p["L_row"], q.L_row), (f"UtAUs_L_col{l}", p["L_col"], q.L_col),
                                (f"UtAUs_U_row{l}", p["U_row"], q.U_row), (f"UtAUs_U_col{l}", p["U_col"], q.U_col)):
            assert np.array_equal(g[key], np.asarray(mine)), key
            assert np.array_equal(g[key], np.asarray(prod)), key + " (product)"
    p0, q0 = S["pattern"][0], H.pattern[0]
    mine = dict(MF_L_row=p0["L_row"], MF_L_col=p0["L_col"], MF_U_row=p0["U_row"], MF_U_col=p0["U_col"], MF_GS_U_Ptr=p0["GS_U_Ptr"],
                MF_GS_L_Ptr=p0["GS_L_Ptr"], MF_GS_U_row=p0["GS_U_row"], MF_GS_U_col=p0["GS_U_col"], MF_GS_L_row=p0["GS_L_row"],
                MF_GS_L_col=p0["GS_L_col"], fix_idx=S["fix_idx"], pin=S["pin"], pin_list=S["pin_list"])
    prod = dict(MF_L_row=q0.L_row, MF_L_col=q0.L_col, MF_U_row=q0.U_row, MF_U_col=q0.U_col, MF_GS_U_Ptr=q0.GS_U_Ptr,
                MF_GS_L_Ptr=q0.GS_L_Ptr, MF_GS_U_row=q0.GS_U_row, MF_GS_U_col=q0.GS_U_col, MF_GS_L_row=q0.GS_L_row,
                MF_GS_L_col=q0.GS_L_col, fix_idx=H.fix_idx)
    for k in INT_KEYS:
        assert np.array_equal(g[k], np.asarray(mine[k])), k
        if k in prod:
            assert np.array_equal(g[k], np.asarray(prod[k])), k + " (product)"
    assert list(g["offs"]) == [p0["nnz"], p0["off_l"], p0["off_d"], p0["off_u"]]
    assert np.array_equal(g["fix_values"], S["fix_values"]) and np.array_equal(g["fix_values"], H.fix_values)
    assert np.array_equal(g["shapeFuncGrad"], S["shapeFuncGrad"]) and np.array_equal(g["shapeFuncGrad"], H.shapeFuncGrad)
    assert np.array_equal(g["MF_value_fixed"], S["MF_value_fixed"])
    assert np.array_equal(g["pin_pos"], S["pin_pos"])
    for l in range(L):
        assert rel(ref.x_gpu_layer[l] if l else S["x0"], g[f"x_gpu_layer{l}"]) <= 1e-6


METHODS = ["NewtonMultigrid", "Newton", "FASv3", "FASv2", "FASv1", "FASv0"]


@pytest.mark.parametrize("method", METHODS)
def test_oracle_steps_match_reference_class(gold, method):
    g = gold
    if f"{method}_x" not in g:
        pytest.skip("method not in this fixture")
    ref = ref_solvers.RefObject(g["points"], g["hexagons"], float(g["dx"]), g["pins"])
    getattr(ref, method)(int(g[f"{method}_steps"]))
    x = ref.dev_x_solved[0] if method in ("FASv0", "FASv1") else ref.x_gpu_layer[0]
    assert rel(x, g[f"{method}_x"]) <= 1e-5, "positions"
    disp_ref = g[f"{method}_x"] - ref.S["x0"]
    assert np.abs((x - ref.S["x0"]) - disp_ref).max() <= 2e-3 * np.abs(disp_ref).max(), "displacement"
    assert rel(ref.MF_value, g[f"{method}_MF_value"]) <= 1e-5, "MF_value"
    assert rel(ref.dev_delta_x[0], g[f"{method}_delta_x0"]) <= 2e-3, "delta_x"
    for l in range(len(ref.UtAUs_values)):
        if f"{method}_UtAUs_values{l}" in g:
            assert rel(ref.UtAUs_values[l], g[f"{method}_UtAUs_values{l}"]) <= 2e-5, f"coarse operator {l}"


@pytest.mark.gpu
@pytest.mark.parametrize("method", METHODS)
def test_cuda_steps_match_reference_class(gold, method):
    torch = pytest.importorskip("torch")
    from fem_simulation_b200 import Object
    g = gold
    if f"{method}_x" not in g:
        pytest.skip("method not in this fixture")
    obj = Object((g["points"], g["hexagons"]), float(g["dx"]), g["pins"])
    getattr(obj, method)(int(g[f"{method}_steps"]))
    x = (obj.dev_x_solved[0] if method in ("FASv0", "FASv1") else obj.x_gpu_layer[0]).cpu().numpy()
    assert rel(x, g[f"{method}_x"]) <= 1e-5, "positions"
    x0 = obj.x_cpu.numpy()
    disp_ref = g[f"{method}_x"] - x0
    assert np.abs((x - x0) - disp_ref).max() <= 2e-3 * np.abs(disp_ref).max(), "displacement"
    assert rel(obj.MF_value_gpu.cpu().numpy(), g[f"{method}_MF_value"]) <= 1e-5, "MF_value"
    for l in range(obj.layer - 1):
        if f"{method}_UtAUs_values{l}" in g:
            assert rel(obj.UtAUs_values[l].cpu().numpy(), g[f"{method}_UtAUs_values{l}"]) <= 2e-5, f"coarse operator {l}"


# ---- dynamic (exp1/simulation) and cloth (exp1/cloth_simulation) reference classes -------------------------------
def _load(name):
    p = os.path.join(GOLD, f"ref_{name}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{name} fixture not generated")
    return dict(np.load(p))


def test_cloth_oracle_matches_reference_class():
    from oracle.ref_cloth import RefCloth
    g = _load("cloth")
    r = RefCloth()
    for k, v in (("edge", r.edge), ("x_cpu", r.x_cpu), ("MF_row", r.MF_row), ("MF_col", r.MF_col), ("offset", r.offset),
                 ("d_offset", r.d_offset)):
        assert np.array_equal(g[k], v), k
    assert rel(r.rest_length, g["rest_length"]) <= 1e-6
    for f in range(int(g["frames"])):
        r.render()
        assert rel(r.dev_x, g[f"x_frame{f}"]) <= 1e-6, f"x frame {f}"
        assert rel(r.dev_v, g[f"v_frame{f}"]) <= 1e-3, f"v frame {f}"
    assert rel(r.MF_value, g["MF_value"]) <= 1e-5


@pytest.mark.gpu
def test_cloth_cuda_matches_reference_class():
    pytest.importorskip("torch")
    from fem_simulation_b200 import Cloth
    g = _load("cloth")
    c = Cloth()
    assert np.array_equal(c.offset.numpy(), g["offset"]) and np.array_equal(c.d_offset.numpy(), g["d_offset"])
    assert np.array_equal(c.MF_row.numpy(), g["MF_row"]) and np.array_equal(c.MF_col.numpy(), g["MF_col"])
    for f in range(int(g["frames"])):
        c.step()
        assert rel(c.dev_x.cpu().numpy(), g[f"x_frame{f}"]) <= 1e-6, f"x frame {f}"
        assert rel(c.dev_v.cpu().numpy(), g[f"v_frame{f}"]) <= 1e-3, f"v frame {f}"
    assert rel(c.MF_value_gpu.cpu().numpy(), g["MF_value"]) <= 1e-5


def test_dynamic_oracle_matches_reference_class():
    from oracle.ref_dynamic import RefDynamicObject
    g = _load("dyn5")
    r = RefDynamicObject(g["points"], g["hexagons"], float(g["dx"]), [int(v) for v in g["pinList"]])
    assert np.allclose(g["consts"], [r.control_mag, 0.033, 30.0, 0.9995, -9.8], rtol=1e-6)
    assert rel(r.m[0], g["m0"]) <= 1e-6 and rel(r.MF_value_fixed, g["MF_value_fixed"]) <= 1e-6
    for f in range(int(g["frames"])):
        r.render()
        assert rel(r.dev_x, g[f"x_frame{f}"]) <= 1e-6, f"x frame {f}"
        assert rel(r.dev_v, g[f"v_frame{f}"]) <= 1e-3, f"v frame {f}"
    assert rel(r.MF_value, g["MF_value"]) <= 1e-5


@pytest.mark.gpu
def test_dynamic_cuda_matches_reference_class():
    pytest.importorskip("torch")
    from fem_simulation_b200 import DynamicObject
    g = _load("dyn5")
    d = DynamicObject((g["points"], g["hexagons"]), float(g["dx"]), [int(v) for v in g["pinList"]])
    for f in range(int(g["frames"])):
        d.step()
        assert rel(d.dev_x.cpu().numpy(), g[f"x_frame{f}"]) <= 1e-6, f"x frame {f}"
        assert rel(d.dev_v.cpu().numpy(), g[f"v_frame{f}"]) <= 1e-3, f"v frame {f}"
    assert rel(d.MF_value_gpu.cpu().numpy(), g["MF_value"]) <= 1e-5


# ---- the diagnostic drivers (compare_withoutGravity, drag, gradientDescent, Adam), reference run under the emulator ----------
DIAG = {m: os.path.join(GOLD, f"ref_diag_{c}.npz") for m, c in (("compare_withoutGravity", "cmpng"), ("drag", "drag"),
                                                                ("gradientDescent", "gd"), ("Adam", "adam"))}
DIAG_HAVE = [m for m, f in DIAG.items() if os.path.exists(f)]
LISTS = ("plot_x", "plot_InfNorm", "plot_energy", "plot_InfNorm_newton", "plot_energy_newton", "plot_InfNorm_newtonMultigrid",
         "plot_energy_newtonMultigrid")


def _check_lists(obj, g, method, tol=2e-5, log_tol=2e-3):
    for nm in LISTS:
        want = g[f"{method}_{nm}"]
        got = np.asarray(getattr(obj, nm), dtype=np.float64)
        assert got.shape == want.shape, f"{method}: {nm} has {got.shape} samples, the reference {want.shape}"
        if want.size == 0:
            continue
        if "InfNorm_newton" in nm:                              # log10 of a norm
            assert np.abs(got - want).max() <= log_tol, f"{method}: {nm}"
        else:
            assert rel(got, want) <= tol, f"{method}: {nm} rel {rel(got, want):.2e}"


def _run_diag(make, method, to_numpy):
    g = dict(np.load(DIAG[method]))
    obj = make(g["points"], g["hexagons"], float(g["dx"]), [int(v) for v in g["pinList"]])
    if method == "drag":
        curves = obj.drag()
        for key in ("GS", "CG", "NewtonMG"):
            want, got = g["drag_" + key], np.asarray(curves[key], dtype=np.float64)
            assert got.shape == want.shape
            live = want > -4.0                                  # below: fp32 round-off of the right-hand side
            # CG in fp32 is not monotone here and sensitive to the order of its dot-product reductions from ~6 iterations on
            tol = 5e-3 if key == "CG" else 2e-3
            assert live[0] and np.abs(got[live] - want[live]).max() <= tol, (key, got, want)
        return
    getattr(obj, method)(int(g[f"{method}_steps"]))
    if method == "Adam":
        # m_hat / (sqrt(v_hat) + eps) is a sign wherever the force is at round-off level, so the trajectory is only
        # reproducible while those components have not moved anything: hold the first samples, loosely
        # (measured: the float32 oracle and the emulated reference agree to 5e-7 on the first sample and drift apart by
        # 0.1-0.6 % over the next four; the energy RISES in both -- the reference subtracts lr * m_hat with the force as m)
        for nm in ("plot_energy", "plot_InfNorm"):
            want, got = g[f"Adam_{nm}"], np.asarray(getattr(obj, nm), dtype=np.float64)
            assert rel(got[:1], want[:1]) <= 1e-5, f"Adam: first {nm} sample rel {rel(got[:1], want[:1]):.2e}"
            assert rel(got[:5], want[:5]) <= 2e-2, f"Adam: {nm} rel {rel(got[:5], want[:5]):.2e}"
        assert len(obj.plot_x) == len(g["Adam_plot_x"])
        return
    _check_lists(obj, g, method)
    assert rel(to_numpy(obj.x_gpu_layer[0]), g[f"{method}_x"]) <= 1e-5


@pytest.mark.parametrize("method", DIAG_HAVE)
def test_oracle_diagnostic_drivers_match_reference_class(method):
    _run_diag(lambda p, h, dx, pins: ref_solvers.RefObject(p, h, dx, pins), method, lambda x: x)


@pytest.mark.gpu
@pytest.mark.parametrize("method", DIAG_HAVE)
def test_cuda_diagnostic_drivers_match_reference_class(method):
    from fem_simulation_b200 import Object
    _run_diag(lambda p, h, dx, pins: Object((p, h), dx, pins), method, lambda x: x.cpu().numpy())


# ---- exp2 compare(): the unmodified exp2/p_hat/quasi_simulation/object.py under the emulator (ref_exp2.npz) -------------------
EXP2 = os.path.join(GOLD, "ref_exp2.npz")


def _exp2_gold():
    g = dict(np.load(EXP2))
    g["pins"] = [int(v) for v in g["pinList"]]
    return g


@pytest.mark.skipif(not os.path.exists(EXP2), reason="fixture not generated")
def test_exp2_oracle_matches_reference_class():
    """oracle/ref_exp2.py vs the reference class: default weights (p_hat_init), both |R|inf curves of compare() -- classic and
    "trained" weights --, the final state and the last coarse Hessian.  The reference's unseeded perturbation of vertex 732
    and the weights it loads from assets/p are recorded in the fixture (oracle/warp_emulator/generate_golden.py)."""
    from oracle.ref_exp2 import RefExp2Object
    g = _exp2_gold()
    r = RefExp2Object(g["points"], g["hexagons"], float(g["dx"]), g["pins"])
    assert np.array_equal(r.w, g["Ut_noOrder_value_init"]) and np.array_equal(r.dev_x_init, g["x_init"])
    r.x_gpu_layer[0][732] += g["perturbation"]
    assert np.array_equal(r.x_gpu_layer[0], g["x_start"])
    x_start = r.x_gpu_layer[0].copy()
    n = int(g["steps"])
    r.compare_loop(n, r.w)
    assert rel(np.array(r.plot_InfNorm), g["plot_InfNorm"]) <= 1e-5
    r.x_gpu_layer[0][...] = x_start
    r.plot_InfNorm = []
    r.compare_loop(n, g["Ut_noOrder_value_trained"])
    assert rel(np.array(r.plot_InfNorm), g["plot_InfNorm_newton"]) <= 1e-5
    assert rel(r.x_gpu_layer[0], g["x_final"]) <= 1e-6
    assert np.abs((r.x_gpu_layer[0] - x_start) - (g["x_final"] - g["x_start"])).max() <= 1e-3 * np.abs(g["x_final"] - g["x_start"]).max()
    assert rel(r.UtAUs_values[0], g["UtAUs_values0"]) <= 1e-5
    assert np.abs(g["plot_InfNorm"][1:] - g["plot_InfNorm_newton"][1:]).max() > 0     # the trained weights change the iteration


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(EXP2), reason="fixture not generated")
def test_exp2_cuda_matches_reference_class():
    import torch
    from fem_simulation_b200.object_exp2 import Exp2Object
    g = _exp2_gold()
    o = Exp2Object((g["points"], g["hexagons"]), float(g["dx"]), g["pins"])
    assert np.array_equal(o.Ut_noOrder_value.cpu().numpy(), g["Ut_noOrder_value_init"])
    n = int(g["steps"])
    o.compare(n, perturb_vertex=732, perturbation=tuple(float(v) for v in g["perturbation"]),
              trained=torch.from_numpy(g["Ut_noOrder_value_trained"]))
    # |R|inf at these states is what is left of O(1) stress terms that cancel (strains of a few per cent): the fp32 evaluation
    # itself is only good to about 1e-5 there, measured as the gap between the oracle's float32 and float64 runs of the same loop
    from oracle.ref_exp2 import RefExp2Object
    curves = {}
    for dt in (np.float32, np.float64):
        r = RefExp2Object(g["points"], g["hexagons"], float(g["dx"]), g["pins"], dtype=dt)
        r.x_gpu_layer[0][732] += g["perturbation"].astype(dt)
        r.compare_loop(n, r.w)
        curves[dt] = np.array(r.plot_InfNorm, dtype=np.float64)
    noise = rel(curves[np.float32], curves[np.float64])
    tol = max(1e-5, 2.0 * noise)
    assert tol <= 1e-4, noise
    assert rel(np.array(o.plot_InfNorm), g["plot_InfNorm"]) <= tol
    assert rel(np.array(o.plot_InfNorm_newton), g["plot_InfNorm_newton"]) <= tol
    x = o.x_gpu_layer[0].cpu().numpy()
    assert rel(x, g["x_final"]) <= 1e-6
    assert np.abs((x - g["x_start"]) - (g["x_final"] - g["x_start"])).max() <= 1e-3 * np.abs(g["x_final"] - g["x_start"]).max()
    assert rel(o.UtAUs_values[0].cpu().numpy(), g["UtAUs_values0"]) <= 2e-5
