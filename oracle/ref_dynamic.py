"""ORACLE (test infrastructure only) -- CPU restatement of the dynamic (implicit-Euler) class,
/root/reference/exp1/simulation/object.py (constants :24-28,705,732; setup :752; render :1286-1325) and its extra
kernels (hexagon.py:174-203, cublas.py:234-299 there)."""
from __future__ import annotations

import numpy as np

from .ref_setup import build_setup
from .ref_solvers import RefObject


class RefDynamicObject(RefObject):
    def __init__(self, points, hexagons, dx, pinList, dtype=np.float32):
        S = build_setup(points, hexagons, dx, pinList, control_mag=10.0)
        if len(pinList):                                           # :732 displaces by 0.05 (build_setup used 0.1)
            S["x0"][pinList[0]][1] += np.float32(0.05) - np.float32(0.1)
            S["pin_pos"] = S["x0"][S["pin_list"]].copy()
        super().__init__(points, hexagons, dx, pinList, dtype=dtype, control_mag=10.0, g=-9.8, setup=S)
        t = dtype
        self.dt, self.inv_dt, self.damping = t(0.033), t(30.0), t(0.9995)
        n = self.N_verts
        self.more_fixed = np.zeros(n, np.int32)
        self.dev_x = self.x_gpu_layer[0].copy()
        self.dev_old_x = np.zeros_like(self.dev_x); self.dev_inertia_x = np.zeros_like(self.dev_x); self.dev_v = np.zeros_like(self.dev_x)
        off_d = self.pat[0]["off_d"]
        for v in range(n):                                         # Hessian_Mass_Kernel at setup (:752)
            self.MF_value_fixed[off_d + self.v2i[0][v]] += self.m[0][v] * np.eye(3, dtype=dtype) * self.inv_dt * self.inv_dt

    def render(self, direction=(0.0, 0.0, 0.0)):
        t = self.dtype
        x = self.dev_x
        d = np.asarray(direction, dtype=t)
        sel = (self.more_fixed == 1) & (self.pin == 0)
        self.all_pin_pos[sel] = x[sel] + d                          # Fixed_Update_Kernel
        self.dev_old_x[...] = x
        self.dev_v[...] = self.dev_v * self.damping                # Basic_Update_Kernel
        x[...] = x + self.dev_v * self.dt
        self.dev_inertia_x[...] = x
        off_d = self.pat[0]["off_d"]
        self.MF_value[...] = 0
        fixed = (self.more_fixed == 1) | (self.pin == 1)           # Hessian_Diag_Kernel
        for v in np.nonzero(fixed)[0]:
            self.MF_value[off_d + self.v2i[0][v]] = t(self.control_mag) * np.eye(3, dtype=t)
        for v in range(self.N_verts):                              # Hessian_Mass_Kernel
            self.MF_value[off_d + self.v2i[0][v]] += self.m[0][v] * np.eye(3, dtype=t) * self.inv_dt * self.inv_dt
        self.hessian(0, x, self.MF_value)
        self.grad[...] = 0
        self.force(x, self.grad)
        mf = np.nonzero(self.more_fixed == 1)[0]                   # compute_partial_more_fixed_energy_X
        self.grad[self.v2i[0][mf]] += t(self.control_mag) * (self.all_pin_pos[mf] - x[mf])
        c = (self.m[0] * self.inv_dt * self.inv_dt)[:, None]       # compute_Inertia_Gradient_Kernel (arguments swapped at :1312)
        self.grad[self.v2i[0]] -= c * (x - self.dev_inertia_x)
        self.set_LUD(0)
        self.dev_B_fixed[0][...] = self.grad
        self.PerformConjugateGradient(layer=0, iterations=5)
        self.update_deltaX(x, self.dev_delta_x[0])
        self.dev_v[...] = (x - self.dev_old_x) * self.inv_dt       # updateVelocity
