"""Generates tests/golden/compare_cube9.npz: the sampled energy / log10|f|inf curves of compare(100)
(object.py:1514-1598) on the 9^3-voxel cube from the CPU oracle in float32 AND float64.  The gap between the
two is the round-off yardstick the GPU convergence test uses.   python tests/golden/make_oracle_curves.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fem_simulation_b200 import mesh
from oracle import ref_solvers

pts, hx = mesh.cube_mesh(9, 0.05)
out = {}
for name, dt in (("f32", np.float32), ("f64", np.float64)):
    ref = ref_solvers.RefObject(pts, hx, 0.05, [], dtype=dt)
    ref.compare(100)
    out[f"plot_x_{name}"] = np.array(ref.plot_x)
    out[f"energy_newton_{name}"] = np.array(ref.plot_energy_newton, dtype=np.float64)
    out[f"energy_fas_{name}"] = np.array(ref.plot_energy_newtonMultigrid, dtype=np.float64)
    out[f"lognorm_newton_{name}"] = np.array(ref.plot_InfNorm_newton, dtype=np.float64)
    out[f"lognorm_fas_{name}"] = np.array(ref.plot_InfNorm_newtonMultigrid, dtype=np.float64)
    out[f"x_final_{name}"] = ref.x_gpu_layer[0].astype(np.float64)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "compare_cube9.npz"), **out)
print("written")
