"""``DynamicObject`` -- the implicit-Euler (dt = 0.033) variant of ``Object``.

Mirrors ``class Object`` of /root/reference/exp1/simulation/object.py: same hierarchy and solvers as the
quasi-static class plus inertia, mass on the diagonal and drag constraints; ``render(pause)`` performs ONE Newton
iteration per frame with 5 CG iterations (object.py:1286-1325 there).  The GL draw, ray picking (``select``) and the
window callbacks are UI and are not reproduced; ``drag(vertex, target)`` / ``clear()`` give the same constraint
mechanism (Control_Kernel / Fixed_Update_Kernel) without a mouse.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .object import Object


class DynamicObject(Object):
    def __init__(self, mesh_path, dx, pinList, **kw):
        kw.setdefault("control_mag", 10.0)          # exp1/simulation/object.py:24
        kw.setdefault("g", -9.8)                    # :705
        kw.setdefault("pin_displacement", 0.05)     # :732
        kw.setdefault("spd_level0", True)           # :1059, :1125
        super().__init__(mesh_path, dx, pinList, **kw)
        self.dt = 0.033                             # :26-28 (inv_dt is 30.0, not 1/0.033 -- kept)
        self.inv_dt = 30.0
        self.damping = 0.9995
        dev = self.device
        n = self.N_verts
        self.more_fixed_gpu = torch.zeros(n, dtype=torch.int32, device=dev)
        # MF_value_fixed also carries m/dt^2 (Hessian_Mass_Kernel at setup, :752) for Newton / NewtonMultigrid
        m_idx = self.m_gpu[0][self.dev_index2vertex[0].long()]
        self.diag_fixed = self.diag_fixed + m_idx * self.inv_dt * self.inv_dt
        self.diag_dyn = torch.zeros(n, device=dev)
        self.dev_x = self.x_gpu_layer[0].clone()
        self.dev_old_x = torch.zeros_like(self.dev_x)
        self.dev_inertia_x = torch.zeros_like(self.dev_x)
        self.dev_v = torch.zeros_like(self.dev_x)
        self.select_vertex = -1
        self.target = np.zeros(3, dtype=np.float32)
        self._dir = (C.c_float * 3)(0.0, 0.0, 0.0)

    # ---- drag constraint (select / moveSelect / clear of the reference, without the ray picking) ----------------
    def drag(self, vertex: int, target):
        """moveSelect (:1275-1279): constrain the vertices within sqrt(0.002) of ``vertex`` and pull them to ``target``."""
        self.select_vertex = int(vertex)
        self.target = np.asarray(target, dtype=np.float32)
        _lib.check(self.lib.femb_control(self.N_verts, self.select_vertex, self.dev_x.data_ptr(), self.more_fixed_gpu.data_ptr(),
                                         self._stream()), "femb_control")

    def clear(self):
        self.select_vertex = -1
        self.more_fixed_gpu.zero_()

    def _drag_dir(self):
        """render() head (:1288-1295): dir = (target - x_select) * min(|.|, 0.5), from the last host copy of x."""
        if self.select_vertex == -1:
            return (0.0, 0.0, 0.0)
        xs = self.dev_x[self.select_vertex].cpu().numpy()
        d = self.target - xs
        ln = float(np.linalg.norm(d))
        d = d * min(ln, 0.5)
        return tuple(float(v) for v in d)

    def _dyn_struct(self):
        d = _lib.FembDynamic()
        d.x, d.v, d.old_x, d.inertia_x = (t.data_ptr() for t in (self.dev_x, self.dev_v, self.dev_old_x, self.dev_inertia_x))
        d.more_fixed, d.diag_dyn = self.more_fixed_gpu.data_ptr(), self.diag_dyn.data_ptr()
        d.cg_p, d.cg_ap, d.cg_state = self.dev_P[0].data_ptr(), self.dev_AP[0].data_ptr(), self._cg_state.data_ptr()
        d.dir_host = C.cast(self._dir, C.c_void_p)
        d.dt, d.inv_dt, d.damping = self.dt, self.inv_dt, self.damping
        return d

    def _frame_body(self):
        if self.use_c_steps:
            if getattr(self, "_dyn", None) is None:
                self._dyn = self._dyn_struct()
            _lib.check(self.lib.femb_dynamic_step(C.byref(self.ctx), C.byref(self._dyn), self._stream()), "femb_dynamic_step")
            return
        lv0 = C.byref(self.levels[0])
        st = self._stream
        x = self.dev_x
        _lib.check(self.lib.femb_dyn_begin(self.N_verts, x.data_ptr(), self.dev_v.data_ptr(), self.dev_old_x.data_ptr(),
                                           self.dev_inertia_x.data_ptr(), self.all_pin_pos_gpu.data_ptr(),
                                           self.more_fixed_gpu.data_ptr(), self.pin_gpu.data_ptr(), self._dir, self.damping,
                                           self.dt, st()), "femb_dyn_begin")                               # :1296-1299
        _lib.check(self.lib.femb_dyn_diag(lv0, self.m_gpu[0].data_ptr(), self.more_fixed_gpu.data_ptr(), self.pin_gpu.data_ptr(),
                                          self.control_mag, self.inv_dt, self.diag_dyn.data_ptr(), st()), "femb_dyn_diag")
        self._force_hessian(x, self.grad_gpu, self.MF_value_gpu, self.diag_dyn)                            # :1302-1310
        _lib.check(self.lib.femb_dyn_force(lv0, x.data_ptr(), self.dev_inertia_x.data_ptr(), self.more_fixed_gpu.data_ptr(),
                                           self.all_pin_pos_gpu.data_ptr(), self.m_gpu[0].data_ptr(), self.control_mag,
                                           self.inv_dt, self.grad_gpu.data_ptr(), st()), "femb_dyn_force")  # :1311-1312
        self._copy(self.dev_B_fixed[0], self.grad_gpu)                                                     # :1318
        self.PerformConjugateGradient(layer=0, iterations=5)                                               # :1321
        self.update_deltaX(x, self.dev_delta_x[0])                                                         # :1324
        _lib.check(self.lib.femb_dyn_velocity(self.N_verts, x.data_ptr(), self.dev_old_x.data_ptr(), self.inv_dt,
                                              self.dev_v.data_ptr(), st()), "femb_dyn_velocity")           # :1325

    def step(self):
        """One frame of ``render(pause=False)`` without the GL draw."""
        d = tuple(C.c_float(v).value for v in self._drag_dir())          # as the kernel will see it (float32)
        if d != tuple(self._dir):
            self._dir[0], self._dir[1], self._dir[2] = d
            self._graphs.pop("dyn", None)            # the drag direction is a kernel argument
        if any(d):
            self._frame_body()                       # while a vertex is dragged the direction changes every frame: no re-capture
        else:
            self._run_step("dyn", self._frame_body)

    def render(self, pause=False):
        if not pause:
            self.step()
