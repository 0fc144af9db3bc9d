"""Runs the UNMODIFIED reference classes (``/root/reference/exp1/{quasi_simulation,simulation,cloth_simulation}/object.py``)
under the shims of this directory and writes golden fixtures to tests/golden/ref_<case>.npz.   Build-container only
(needs /root/reference).

    python oracle/warp_emulator/generate_golden.py [case ...]
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_DIRS = {"quasi": "/root/reference/exp1/quasi_simulation", "dynamic": "/root/reference/exp1/simulation",
            "cloth": "/root/reference/exp1/cloth_simulation", "exp2": "/root/reference/exp2/p_hat/quasi_simulation"}
sys.path.insert(0, ROOT)

_orig_to = torch.Tensor.to


def _to(self, *a, **k):                                  # `.to('cuda:0')` -> stay on the CPU
    if a and isinstance(a[0], str) and a[0].startswith("cuda"):
        return self
    return _orig_to(self, *a, **k)


torch.Tensor.to = _to

# torch has already imported the real (namespace) `cuda` package: plant the inert cudart stand-in inside it
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("cuda.cudart", os.path.join(HERE, "shims", "cuda", "cudart.py"))
_cudart = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_cudart)
sys.modules["cuda.cudart"] = _cudart
if "cuda" in sys.modules:
    sys.modules["cuda"].cudart = _cudart

from fem_simulation_b200 import mesh  # noqa: E402

CASES = {
    # name: (reference dir, mesh builder, dx, pinList, {method: steps})
    "jag9": ("quasi", lambda: mesh.jagged_mesh(9, 9, 10, 0.05, seed=3, keep=0.75), 0.05, [],
             {"NewtonMultigrid": 2, "FASv3": 2, "FASv0": 1, "Newton": 2}),
    # 3 levels (what the reference's hierarchy effectively is): Galerkin chain 0 -> 1 -> 2 and the 3-level V-cycle
    "jag17": ("quasi", lambda: mesh.jagged_mesh(16, 17, 16, 0.05, seed=13, keep=0.12), 0.05, [], {"NewtonMultigrid": 1}),
    "jag17_pin": ("quasi", lambda: mesh.jagged_mesh(17, 16, 16, 0.05, seed=17, keep=0.12), 0.05, [25],
                  {"NewtonMultigrid": 2, "FASv3": 1}),
    "jag9_pin": ("quasi", lambda: mesh.jagged_mesh(9, 10, 9, 0.05, seed=5, keep=0.75), 0.05, [40], {"FASv2": 2, "FASv1": 1}),
    # the diagnostic drivers: sampled lists / plotted curves instead of a final state; one case (= one process, one file)
    # per driver because the emulator needs tens of minutes for each
    "diag_cmpng": ("quasi", lambda: mesh.jagged_mesh(8, 9, 8, 0.05, seed=11, keep=0.7), 0.05, [40], {"compare_withoutGravity": 10}),
    "diag_drag": ("quasi", lambda: mesh.jagged_mesh(8, 9, 8, 0.05, seed=11, keep=0.7), 0.05, [40], {"drag": None}),
    "diag_gd": ("quasi", lambda: mesh.jagged_mesh(8, 9, 8, 0.05, seed=11, keep=0.7), 0.05, [40], {"gradientDescent": 100}),
    "diag_adam": ("quasi", lambda: mesh.jagged_mesh(8, 9, 8, 0.05, seed=11, keep=0.7), 0.05, [40], {"Adam": 100}),
    "dyn5": ("dynamic", lambda: mesh.jagged_mesh(5, 6, 5, 0.05, seed=7, keep=0.8), 0.05, [], {"render": 3}),
    "cloth": ("cloth", None, None, None, {"render": 3}),
    # exp2 (the canonical p_hat copy): compare() = the two-level forward step with classic and with "trained" position-
    # restriction weights.  The mesh must have more than 732 vertices (compare perturbs x[732], object.py:1294 there).
    "exp2": ("exp2", lambda: mesh.jagged_mesh(10, 9, 11, 0.05, seed=1), 0.05, [300], {"compare": 3}),
}


def use_reference(kind):
    """Put ONE reference experiment directory (they all name their modules object / cublas / cpu_function) and the shims
    on sys.path.  One case per process."""
    sys.path.insert(0, REF_DIRS[kind])
    sys.path.insert(0, os.path.join(HERE, "shims"))      # shadows warp / sparse / pyvista / OpenGL / cuda / matplotlib / glm


def snapshot_setup(obj):
    out = {"layer": int(obj.layer), "dims": np.array(obj.dims), "N_pin": obj.N_pin}
    for l in range(int(obj.layer)):
        out[f"hexs{l}"] = obj.hexs[l].numpy()
        out[f"x{l}"] = obj.x[l].numpy()
        out[f"v2i{l}"] = obj.dev_vertex2index[l].numpy()
        out[f"i2v{l}"] = obj.dev_index2vertex[l].numpy()
        out[f"cvn{l}"] = np.array(obj.color_vertex_num[l])
        out[f"hex_update_offset{l}"] = obj.hex_update_offset_gpu[l].numpy()
        out[f"x_gpu_layer{l}"] = obj.x_gpu_layer[l].numpy().copy()
        out[f"det{l}"] = obj.det_pX_peps_gpu[l].numpy().copy()
        out[f"inv{l}"] = obj.inverse_pX_peps_gpu[l].numpy().copy()
        out[f"m{l}"] = obj.m_gpu[l].numpy().copy()
    for l in range(int(obj.layer) - 1):
        out[f"fine_coarse{l}"] = obj.fine_coarse[l].numpy()
        for nm in ("Ut", "Us", "Ut_noOrder", "Us_noOrder", "Ut_hat", "Us_hat", "Ut_noOrder_hat", "Us_noOrder_hat"):
            m = getattr(obj, nm)[l]
            out[f"{nm}{l}_offsets"] = m.offsets.numpy()[:m.nrow + 1]
            out[f"{nm}{l}_columns"] = m.columns.numpy()[:m.nnz]
            out[f"{nm}{l}_values"] = m.values.numpy()[:m.nnz]
        out[f"UtAUs_block_offset{l}"] = obj.UtAUs_block_offset[l].numpy()
        out[f"UtAUs_GS_U_Ptr{l}"] = np.array(obj.UtAUs_GS_U_Ptr[l]); out[f"UtAUs_GS_L_Ptr{l}"] = np.array(obj.UtAUs_GS_L_Ptr[l])
        out[f"UtAUs_L_row{l}"] = obj.UtAUs_L_row[l].numpy(); out[f"UtAUs_L_col{l}"] = obj.UtAUs_L_col[l].numpy()
        out[f"UtAUs_U_row{l}"] = obj.UtAUs_U_row[l].numpy(); out[f"UtAUs_U_col{l}"] = obj.UtAUs_U_col[l].numpy()
    out["MF_L_row"] = obj.MF_L_row.numpy(); out["MF_L_col"] = obj.MF_L_col.numpy()
    out["MF_U_row"] = obj.MF_U_row.numpy(); out["MF_U_col"] = obj.MF_U_col.numpy()
    out["MF_GS_U_Ptr"] = np.array(obj.MF_GS_U_Ptr); out["MF_GS_L_Ptr"] = np.array(obj.MF_GS_L_Ptr)
    out["MF_GS_U_row"] = obj.MF_GS_U_row_gpu.numpy(); out["MF_GS_U_col"] = obj.MF_GS_U_col_gpu.numpy()
    out["MF_GS_L_row"] = obj.MF_GS_L_row_gpu.numpy(); out["MF_GS_L_col"] = obj.MF_GS_L_col_gpu.numpy()
    out["offs"] = np.array([obj.MF_nnz, obj.off_l, obj.off_d, obj.off_u])
    out["fix_idx"] = obj.fix_idx_gpu.numpy(); out["fix_values"] = obj.fix_values_gpu.numpy()
    out["pin"] = obj.pin_gpu.numpy(); out["pin_list"] = obj.pin_list_gpu.numpy(); out["pin_pos"] = obj.pin_pos_gpu.numpy()
    out["MF_value_fixed"] = obj.MF_value_fixed_gpu.numpy()
    out["shapeFuncGrad"] = obj.shapeFuncGrad_gpu.numpy()
    return out


def run_dynamic(name):
    kind, build, dx, pins, methods = CASES[name]
    use_reference(kind)
    import object as ref_object                           # exp1/simulation/object.py, unmodified
    pts, hx = build()
    path = os.path.join(ROOT, "tests", "golden", f"_mesh_{name}.npz")
    np.savez(path, points=pts, hexagons=hx)
    out = {"points": pts, "hexagons": hx, "dx": dx, "pinList": np.array(pins, dtype=np.int64)}
    with contextlib.redirect_stdout(io.StringIO()):
        obj = ref_object.Object("npz:" + path, dx, list(pins))
        out["m0"] = obj.m_gpu[0].numpy().copy()
        out["MF_value_fixed"] = obj.MF_value_fixed_gpu.numpy().copy()
        out["consts"] = np.array([obj.control_mag, obj.dt, obj.inv_dt, obj.damping, obj.g_gpu.numpy()[0]], dtype=np.float64)
        for f in range(methods["render"]):
            obj.render(False)
            out[f"x_frame{f}"] = obj.dev_x.numpy().copy()
            out[f"v_frame{f}"] = obj.dev_v.numpy().copy()
        out["MF_value"] = obj.MF_value_gpu.numpy().copy()
        out["frames"] = methods["render"]
    os.remove(path)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"), **out)
    print("written", name, flush=True)


def run_cloth(name):
    kind, _, _, _, methods = CASES[name]
    use_reference(kind)
    import object as ref_object                           # exp1/cloth_simulation/object.py, unmodified
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        obj = ref_object.Cloth()
        out["edge"] = obj.edge.numpy(); out["x_cpu"] = obj.x_cpu.numpy().copy()
        out["MF_row"] = obj.MF_row.numpy(); out["MF_col"] = obj.MF_col.numpy()
        out["offset"] = obj.offset_gpu.numpy(); out["d_offset"] = obj.d_offset.numpy()
        out["rest_length"] = obj.rest_length.numpy().copy()
        for f in range(methods["render"]):
            obj.render(False)
            out[f"x_frame{f}"] = obj.dev_x.numpy().copy()
            out[f"v_frame{f}"] = obj.dev_v.numpy().copy()
        out["MF_value"] = obj.MF_value_gpu.numpy().copy()
        out["frames"] = methods["render"]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"), **out)
    print("written", name, flush=True)


def run_exp2(name):
    """exp2/p_hat/quasi_simulation/object.py, unmodified.  Three things its compare() takes from outside are pinned here and
    recorded in the fixture: the unseeded torch.rand perturbation (torch.manual_seed), and the "trained" weights it loads
    from assets/p/Ut_noOrder_0.pth (dragon-sized: 26539 entries) -- torch.load is replaced by a stand-in that returns a
    deterministic perturbation of the classic weights of THIS mesh."""
    kind, build, dx, pins, methods = CASES[name]
    use_reference(kind)
    import object as ref_object
    pts, hx = build()
    path = os.path.join(ROOT, "tests", "golden", f"_mesh_{name}.npz")
    np.savez(path, points=pts, hexagons=hx)
    out = {"points": pts, "hexagons": hx, "dx": dx, "pinList": np.array(pins, dtype=np.int64)}
    with contextlib.redirect_stdout(io.StringIO()):
        obj = ref_object.Object("npz:" + path, dx, list(pins))
        w0 = obj.Ut_noOrder_value.numpy().copy()
        out["Ut_noOrder_value_init"] = w0
        out["x_init"] = obj.dev_x_init.numpy().copy()
        rng = np.random.default_rng(11)
        w_tr = (w0 * (1.0 + 0.05 * rng.standard_normal(len(w0)))).astype(np.float32)
        out["Ut_noOrder_value_trained"] = w_tr
        torch.load = lambda *a, **k: torch.from_numpy(w_tr.copy())
        torch.manual_seed(20251018)
        x_before = obj.x[0].numpy().copy() if hasattr(obj.x[0], "numpy") else np.array(obj.x[0])
        obj.compare(methods["compare"])
        x_after = obj.x[0].numpy().copy() if hasattr(obj.x[0], "numpy") else np.array(obj.x[0])
        out["perturbation"] = (x_after[732].astype(np.float64) - x_before[732].astype(np.float64)).astype(np.float32)
        out["x_start"] = obj.x_cpu.numpy().copy()
        out["plot_InfNorm"] = np.array(obj.plot_InfNorm, dtype=np.float64)
        out["plot_InfNorm_newton"] = np.array(obj.plot_InfNorm_newton, dtype=np.float64)
        out["x_final"] = obj.x_gpu_layer[0].numpy().copy()
        out["UtAUs_values0"] = obj.UtAUs_values[0].numpy().copy()
        out["steps"] = methods["compare"]
    os.remove(path)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"), **out)
    print("written", name, flush=True)


def run_case(name):
    if CASES[name][0] == "exp2":
        return run_exp2(name)
    if CASES[name][0] == "dynamic":
        return run_dynamic(name)
    if CASES[name][0] == "cloth":
        return run_cloth(name)
    use_reference(CASES[name][0])
    import object as ref_object                           # exp1/quasi_simulation/object.py, unmodified
    _, build, dx, pins, methods = CASES[name]
    pts, hx = build()
    path = os.path.join(ROOT, "tests", "golden", f"_mesh_{name}.npz")
    np.savez(path, points=pts, hexagons=hx)
    out = {"points": pts, "hexagons": hx, "dx": dx, "pinList": np.array(pins, dtype=np.int64)}
    for k, (method, steps) in enumerate(methods.items()):
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            obj = ref_object.Object("npz:" + path, dx, list(pins))
            if k == 0 and not name.startswith("diag_"):
                out.update(snapshot_setup(obj))
            if method == "drag":
                import matplotlib.pyplot as plt                   # the recording stand-in
                plt.PLOTS.clear()
                obj.drag()
                for a, kw in plt.PLOTS:
                    out["drag_" + kw["label"]] = np.array(a[1], dtype=np.float64)
            else:
                getattr(obj, method)(steps)
            for nm in ("plot_x", "plot_InfNorm", "plot_energy", "plot_InfNorm_newton", "plot_energy_newton",
                       "plot_InfNorm_newtonMultigrid", "plot_energy_newtonMultigrid"):
                out[f"{method}_{nm}"] = np.array(getattr(obj, nm), dtype=np.float64)
        xs = obj.dev_x_solved[0] if method in ("FASv0", "FASv1") else obj.x_gpu_layer[0]
        out[f"{method}_steps"] = -1 if steps is None else steps
        out[f"{method}_x"] = xs.numpy().copy()
        out[f"{method}_MF_value"] = obj.MF_value_gpu.numpy().copy()
        out[f"{method}_delta_x0"] = obj.dev_delta_x[0].numpy().copy()
        for l in range(int(obj.layer) - 1):
            out[f"{method}_UtAUs_values{l}"] = obj.UtAUs_values[l].numpy().copy()
        print(f"{name}: {method} x{steps} done in {time.time() - t0:.0f} s", flush=True)
    os.remove(path)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"), **out)
    print("written", name, flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 2:
        run_case(sys.argv[1])
    else:                                                 # one process per case: the experiments share module names
        import subprocess
        for c in (sys.argv[1:] or list(CASES)):
            subprocess.run([sys.executable, os.path.abspath(__file__), c], check=True)
