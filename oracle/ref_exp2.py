"""ORACLE (test infrastructure only) -- forward path of exp2 ``compare``,
/root/reference/exp2/p_hat/quasi_simulation/object.py:1289-1357 (temp_kernel.py:2-48 materialises F/E; same physics)."""
from __future__ import annotations

import numpy as np

from . import ref_kernels as K
from .ref_setup import build_setup
from .ref_solvers import RefObject, bsr_mv


class RefExp2Object(RefObject):
    def __init__(self, points, hexagons, dx, pinList, dtype=np.float32):
        S = build_setup(points, hexagons, dx, pinList)
        if len(pinList):                                         # exp2 object.py:795: displacement 0.0
            S["x0"][pinList[0]][1] -= np.float32(0.1)
            S["pin_pos"] = S["x0"][S["pin_list"]].copy()
        super().__init__(points, hexagons, dx, pinList, dtype=dtype, setup=S)
        self.dev_x_init = S["x_rest"].astype(dtype)
        self.X = self.x_gpu_layer[0].copy()
        self.w = self.Ut_noOrder[0]["values"].copy()
        self.plot_InfNorm = []

    def compare_loop(self, iterations, w):
        x = self.x_gpu_layer[0]
        Ut = dict(self.Ut_noOrder[0], values=np.asarray(w, self.dtype))
        for _ in range(iterations):
            self.dev_R[0][...] = 0
            self.force(x, self.dev_R[0], gravity=False, pins=False)
            self.plot_InfNorm.append(K.inf_norm(self.dev_R[0]))
            bsr_mv(self.Ut_hat[0], self.dev_R[0], self.dev_B_fixed[1], alpha=1.0, beta=0.0)
            bsr_mv(Ut, self.dev_x_init, self.dev_x_solved[1], alpha=1.0, beta=0.0)
            self.det[1], self.inv[1] = K.prepare_kernal(self.dev_x_solved[1], self.hexs[1], self.sfg, self.dtype)
            x1 = np.zeros_like(self.dev_x_solved[1])
            bsr_mv(Ut, self.X, x1, alpha=1.0, beta=0.0)
            self.UtAUs_values[0][...] = 0
            self.hessian(1, x1, self.UtAUs_values[0])
            p = self.pat[1]
            D = self.UtAUs_values[0][p["off_d"]:p["off_d"] + self.dims[1]]
            self.dev_delta_x[1][...] = 0
            xs, _ = K.solve3x3(D, self.dev_B_fixed[1], self.dev_delta_x[1])
            self.dev_delta_x[1][...] = xs
            bsr_mv(self.Us_hat[0], self.dev_delta_x[1], self.dev_delta_x[0], alpha=1.0, beta=0.0)
            self.update_deltaX(x, self.dev_delta_x[0])
