"""``Cloth`` -- the 65x65 mass-spring cloth of /root/reference/exp1/cloth_simulation/object.py.

Same constructor (no arguments), constants (dt = 0.0033, k = 1000, damping 0.9995, g = -9.8, unit masses), grid /
edge construction (object.py:31-73 there), ``offset`` / ``d_offset`` maps (:139-176) and per-frame physics of
``render(pause)`` (:657-677): implicit-Euler predictor, spring Hessian + mass on a fixed pattern, force, 5 CG
iterations, update with vertices 0 and 65 frozen, velocity.  The multigrid methods of that class are dead code
 and are not reproduced; neither are the GL parts.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class Cloth:
    def __init__(self, *, res_x=64, res_y=64, device="cuda:0", use_graphs=True, use_c_steps=True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.FembError("femb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        _lib.check(self.lib.femb_check_device(self.device.index or 0), "femb_check_device")
        torch.cuda.set_device(self.device)
        self.use_graphs = use_graphs
        self.use_c_steps = use_c_steps        # a frame as ONE C-ABI call (femb_cloth_step)
        self._graph = None
        self.layer = 1
        self.control_mag, self.spd_value = 10.0, 1e-3
        self.dt = 0.0033
        self.inv_dt = 1.0 / self.dt
        self.damping, self.k, self.g = 0.9995, 1000.0, -9.8
        self.res_x, self.res_y = res_x, res_y
        x_cpu, face, edge = cloth_grid(res_x, res_y)
        self.x_cpu, self.face, self.edge = torch.from_numpy(x_cpu), torch.from_numpy(face), torch.from_numpy(edge)
        self.N_verts, self.N_face, self.N_edge = len(x_cpu), len(face), len(edge)
        pat = cloth_pattern(self.N_verts, edge)
        self.MF_nnz = pat["nnz"]
        self.MF_row, self.MF_col = torch.from_numpy(pat["MF_row"]), torch.from_numpy(pat["MF_col"])
        self.offset, self.d_offset = torch.from_numpy(pat["offset"]), torch.from_numpy(pat["d_offset"])
        dev = self.device
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.edge_gpu = d(edge)
        self.x_gpu = d(x_cpu)
        self.dev_x = self.x_gpu.clone()
        z = lambda: torch.zeros_like(self.dev_x)
        self.dev_v, self.dev_old_x, self.dev_inertia_x = z(), z(), z()
        self.dev_B, self.dev_P, self.dev_AP, self.dev_delta_x, self.dev_x_solved = z(), z(), z(), z(), z()
        self.grad_gpu, self.dev_B_fixed = z(), z()
        self.m_gpu = torch.ones(self.N_verts, device=dev)
        self.rest_length = torch.zeros(self.N_edge, device=dev)
        # internal [L;D;U] value layout of the same pattern (the row-pass kernels stream it); sorted -> internal slot
        n = self.N_verts
        self._slot_of_sorted = d(pat["ldu_slot"].astype(np.int64))
        pad = np.zeros(4, np.int32)
        self._arrs = dict(L_ptr=d(pat["L_ptr"]), L_col=d(np.concatenate([pat["L_col"], pad]))[:len(pat["L_col"])],
                          U_ptr=d(pat["U_ptr"]), U_col=d(np.concatenate([pat["U_col"], pad]))[:len(pat["U_col"])],
                          ident=d(np.arange(n, dtype=np.int32)), inc_ptr=d(pat["inc_ptr"]), inc=d(pat["inc"]),
                          off01=d(pat["ldu_slot"][pat["offset"][1::4]].astype(np.int32)),
                          off10=d(pat["ldu_slot"][pat["offset"][2::4]].astype(np.int32)),
                          dslot=d(pat["ldu_slot"][pat["d_offset"]].astype(np.int32)))
        lv = _lib.FembLevel()
        lv.n_vert, lv.n_hex, lv.nnz = n, 0, self.MF_nnz
        lv.off_l, lv.off_d, lv.off_u = 0, pat["L_nnz"], pat["L_nnz"] + n
        lv.color_start[0] = 0
        for c in range(1, 9):
            lv.color_start[c] = n
        lv.v2i = lv.i2v = self._arrs["ident"].data_ptr()
        for kname in ("L_ptr", "L_col", "U_ptr", "U_col"):
            setattr(lv, kname, self._arrs[kname].data_ptr())
        self.level = lv
        buf = torch.zeros(self.MF_nnz * 9 + 8, device=dev)
        self._values = buf[:self.MF_nnz * 9].view(self.MF_nnz, 3, 3)
        self._edge_h = torch.zeros(self.N_edge * 9, device=dev)
        self._edge_g = torch.zeros(self.N_edge * 3, device=dev)
        self._cg_state = torch.zeros(16 + 1024, device=dev)
        self._zeros_i = torch.zeros(n, dtype=torch.int32, device=dev)
        self._dir = (C.c_float * 3)(0.0, 0.0, 0.0)
        self.offset_gpu, self.d_offset_gpu = d(pat["offset"]), d(pat["d_offset"])
        self.MF_row_gpu, self.MF_col_gpu = d(pat["MF_row"]), d(pat["MF_col"])
        _lib.check(self.lib.femb_cloth_rest_length(self.N_edge, self.edge_gpu.data_ptr(), self.x_gpu.data_ptr(),
                                                   self.rest_length.data_ptr(), self._stream()), "femb_cloth_rest_length")
        torch.cuda.synchronize(dev)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def MF_value_gpu(self):
        """Block values in the reference's (row,col)-sorted order (object.py:129-158 there)."""
        return self._values[self._slot_of_sorted]

    def _cloth_struct(self):
        a = self._arrs
        k = _lib.FembCloth()
        k.level = C.cast(C.pointer(self.level), C.c_void_p)
        k.n_edge = self.N_edge
        k.edge, k.rest, k.mass = self.edge_gpu.data_ptr(), self.rest_length.data_ptr(), self.m_gpu.data_ptr()
        k.inc_ptr, k.inc = a["inc_ptr"].data_ptr(), a["inc"].data_ptr()
        k.off_p0p1, k.off_p1p0, k.d_slot = a["off01"].data_ptr(), a["off10"].data_ptr(), a["dslot"].data_ptr()
        k.zeros_i = self._zeros_i.data_ptr()
        k.x, k.v, k.old_x, k.inertia_x = (t.data_ptr() for t in (self.dev_x, self.dev_v, self.dev_old_x, self.dev_inertia_x))
        k.values, k.grad = self._values.data_ptr(), self.grad_gpu.data_ptr()
        k.B, k.B_fixed, k.delta_x = self.dev_B.data_ptr(), self.dev_B_fixed.data_ptr(), self.dev_delta_x.data_ptr()
        k.cg_p, k.cg_ap, k.cg_state = self.dev_P.data_ptr(), self.dev_AP.data_ptr(), self._cg_state.data_ptr()
        k.edge_h, k.edge_g = self._edge_h.data_ptr(), self._edge_g.data_ptr()
        k.k, k.g, k.dt, k.inv_dt, k.damping = self.k, self.g, self.dt, self.inv_dt, self.damping
        return k

    def _frame_body(self):
        if self.use_c_steps:
            if getattr(self, "_cloth", None) is None:
                self._cloth = self._cloth_struct()
            _lib.check(self.lib.femb_cloth_step(C.byref(self._cloth), self._stream()), "femb_cloth_step")
            return
        st, a = self._stream, self._arrs
        n = self.N_verts
        x = self.dev_x
        _lib.check(self.lib.femb_dyn_begin(n, x.data_ptr(), self.dev_v.data_ptr(), self.dev_old_x.data_ptr(),
                                           self.dev_inertia_x.data_ptr(), self.dev_x_solved.data_ptr(), self._zeros_i.data_ptr(),
                                           self._zeros_i.data_ptr(), self._dir, self.damping, self.dt, st()), "femb_dyn_begin")
        _lib.check(self.lib.femb_cloth_assemble(n, self.N_edge, self.edge_gpu.data_ptr(), self.rest_length.data_ptr(), self.k,
                                                x.data_ptr(), self.m_gpu.data_ptr(), self.g, self.dev_inertia_x.data_ptr(),
                                                self.inv_dt, a["inc_ptr"].data_ptr(), a["inc"].data_ptr(), a["off01"].data_ptr(),
                                                a["off10"].data_ptr(), a["dslot"].data_ptr(), self._edge_h.data_ptr(),
                                                self._edge_g.data_ptr(), self._values.data_ptr(), self.grad_gpu.data_ptr(), st()),
                   "femb_cloth_assemble")
        _lib.check(self.lib.femb_copy(3 * n, self.grad_gpu.data_ptr(), self.dev_B_fixed.data_ptr(), st()), "femb_copy")
        _lib.check(self.lib.femb_cg(C.byref(self.level), self._values.data_ptr(), self.dev_B_fixed.data_ptr(),
                                    self.dev_delta_x.data_ptr(), self.dev_B.data_ptr(), self.dev_P.data_ptr(),
                                    self.dev_AP.data_ptr(), self._cg_state.data_ptr(), 5, 1e-5, 0, st()), "femb_cg")
        _lib.check(self.lib.femb_cloth_update(n, self.dev_delta_x.data_ptr(), x.data_ptr(), 0, 65, st()), "femb_cloth_update")
        _lib.check(self.lib.femb_dyn_velocity(n, x.data_ptr(), self.dev_old_x.data_ptr(), self.inv_dt, self.dev_v.data_ptr(), st()),
                   "femb_dyn_velocity")

    def step(self):
        if not self.use_graphs:
            self._frame_body()
            return
        if self._graph is None:
            self._frame_body()
            torch.cuda.synchronize(self.device)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._frame_body()
            return
        self._graph.replay()

    def render(self, pause=False):
        if not pause:
            self.step()


def cloth_grid(res_x=64, res_y=64):
    """object.py:31-73 there: vertices, faces, edges (structural x, structural y, one diagonal)."""
    nvx = res_x + 1
    i = np.arange(nvx * (res_y + 1))
    x = np.zeros((len(i), 3), np.float32)
    x[:, 0] = (i % nvx).astype(np.float32) * np.float32(1.0 / res_x)
    x[:, 1] = (i // nvx).astype(np.float32) * np.float32(1.0 / res_y)
    face = np.zeros((res_x * res_y * 2, 3), np.int32)
    ii, jj = np.meshgrid(np.arange(res_y), np.arange(res_x), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    f0 = 2 * (ii * res_x + jj)
    face[f0] = np.stack([ii * nvx + jj, ii * nvx + jj + 1, (ii + 1) * nvx + jj], 1)
    face[f0 + 1] = np.stack([ii * nvx + jj + 1, (ii + 1) * nvx + jj + 1, (ii + 1) * nvx + jj], 1)
    n_e = res_x * (res_y + 1) + res_y * (res_x + 1) + res_x * res_y
    edge = np.zeros((n_e, 2), np.int32)
    a, b = np.meshgrid(np.arange(res_y + 1), np.arange(res_x), indexing="ij")
    a, b = a.reshape(-1), b.reshape(-1)
    edge[a * res_x + b] = np.stack([a * nvx + b, a * nvx + b + 1], 1)
    a, b = np.meshgrid(np.arange(res_x + 1), np.arange(res_y), indexing="ij")
    a, b = a.reshape(-1), b.reshape(-1)
    edge[(res_y + 1) * res_x + a * res_y + b] = np.stack([b * nvx + a, (b + 1) * nvx + a], 1)
    base = (res_y + 1) * res_x + (res_x + 1) * res_y
    edge[base + ii * res_x + jj] = np.stack([ii * nvx + jj, (ii + 1) * nvx + jj + 1], 1)
    return x, face, edge


def cloth_pattern(n, edge):
    """object.py:129-176 there: (row,col)-sorted pattern, ``offset[4e..4e+3]`` and ``d_offset[v]`` -- plus the [L;D;U]
    view of the same pattern and the vertex->edge incidence used by the CUDA kernels."""
    e = edge.astype(np.int64)
    rows = np.concatenate([e[:, 0], e[:, 1], np.arange(n)])
    cols = np.concatenate([e[:, 1], e[:, 0], np.arange(n)])
    key = rows * n + cols
    order = np.argsort(key, kind="stable")
    rank = np.empty(len(key), np.int64); rank[order] = np.arange(len(key))
    MF_row, MF_col = rows[order].astype(np.int32), cols[order].astype(np.int32)
    ne = len(e)
    offset = np.zeros(4 * ne, np.int32)
    d_rank = rank[2 * ne:]
    offset[0::4] = d_rank[e[:, 0]]; offset[1::4] = rank[:ne]; offset[2::4] = rank[ne:2 * ne]; offset[3::4] = d_rank[e[:, 1]]
    # [L;D;U] slots of the sorted entries
    isL, isU = MF_row > MF_col, MF_row < MF_col
    L_nnz = int(isL.sum())
    slot = np.empty(len(key), np.int64)
    slot[isL] = np.arange(L_nnz)
    slot[MF_row == MF_col] = L_nnz + MF_row[MF_row == MF_col]
    slot[isU] = L_nnz + n + np.arange(int(isU.sum()))
    L_row, U_row = MF_row[isL], MF_row[isU]
    inc_v = np.concatenate([e[:, 0], e[:, 1]])
    inc_e = np.concatenate([np.arange(ne) * 2, np.arange(ne) * 2 + 1])
    o2 = np.argsort(inc_v, kind="stable")
    inc_ptr = np.zeros(n + 1, np.int64); np.add.at(inc_ptr, inc_v + 1, 1)
    return dict(nnz=len(key), MF_row=MF_row, MF_col=MF_col, offset=offset, d_offset=d_rank.astype(np.int32), ldu_slot=slot,
                L_nnz=L_nnz, L_ptr=np.searchsorted(L_row, np.arange(n + 1)).astype(np.int32), L_col=MF_col[isL],
                U_ptr=np.searchsorted(U_row, np.arange(n + 1)).astype(np.int32), U_col=MF_col[isU],
                inc_ptr=np.cumsum(inc_ptr).astype(np.int32), inc=inc_e[o2].astype(np.int32))
