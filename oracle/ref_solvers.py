"""ORACLE (test infrastructure only) -- CPU restatement of the reference's solver methods.

``RefObject`` mirrors ``class Object`` of ``exp1/quasi_simulation/object.py`` (paths relative
to /root/reference): same attribute names, same method names, same order of operations,
numpy arrays instead of Warp arrays.  Every method cites the lines it follows.  The
``bsr_*`` helpers restate ``sparse.py`` (vendored warp.sparse): ``bsr_mv`` (:1143-1261),
``bsr_mm`` (:822-1140, same row-major accumulation order as ``_bsr_mm_compute_values``).

Parity status: see ``oracle/ref_setup.py``.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import numpy as np

from . import ref_kernels as K
from .ref_setup import build_setup, bsr_from_triplets

EPSILON = 1e-7


def _rows_of(m):
    return np.repeat(np.arange(m["nrow"]), np.diff(m["offsets"]))


def bsr_mv(A, x, y, alpha=1.0, beta=0.0, x_offset=0, y_offset=0):
    """sparse.py:1143-1171: y[r+yo] = alpha * sum_blocks A*x[c+xo] + beta*y[r+yo]."""
    t = y.dtype.type
    v = np.zeros((A["nrow"], 3), y.dtype)
    if A["nnz"]:
        rows = _rows_of(A)
        xs = x[A["columns"].astype(np.int64) + x_offset]
        vals = A["values"]
        contrib = vals[:, None] * xs if vals.ndim == 1 else np.einsum("nij,nj->ni", vals, xs)
        np.add.at(v, rows, contrib.astype(y.dtype))
    v = v * t(alpha)
    ys = y[y_offset:y_offset + A["nrow"]]
    if beta != 0.0:
        v = v + t(beta) * ys
    ys[...] = v
    return y


def _as_blocks(m):
    if m["values"].ndim == 3:
        return m["values"]
    return m["values"][:, None, None] * np.eye(3, dtype=m["values"].dtype)


def bsr_mm(X, Y):
    """sparse.py:950-1140: z = x*y, structural pattern (no pruning: values pointer is NULL at :1078-1089),
    values accumulated row by row in (x_block, y_block) order (:873-901)."""
    xr = _rows_of(X)
    ycnt = np.diff(Y["offsets"])[X["columns"]]
    xk = np.repeat(np.arange(X["nnz"]), ycnt)
    start = np.repeat(Y["offsets"][X["columns"]].astype(np.int64), ycnt)
    within = np.arange(len(xk)) - np.repeat(np.cumsum(ycnt) - ycnt, ycnt)
    ym = start + within
    vals = _as_blocks(X)[xk] @ _as_blocks(Y)[ym]
    return bsr_from_triplets(X["nrow"], Y["ncol"], xr[xk], Y["columns"][ym], vals)


class RefObject:
    """CPU twin of the reference ``Object`` (physics only; GL/plots dropped)."""

    def __init__(self, points, hexagons, dx, pinList, dtype=np.float32, control_mag=100.0,
                 mu=250.0, la=0.0, g=-9.8 / 2.0, spd_value=1e-3, setup=None, corrected_spacing=False):
        self.dtype = dtype
        S = setup if setup is not None else build_setup(points, hexagons, dx, pinList, control_mag, corrected_spacing)
        self.S = S
        self.layer = S["layer"]
        self.dims = S["dims"]
        self.control_mag = control_mag
        self.spd_value = spd_value
        self.mu, self.la, self.g = mu, la, g
        self.N_verts, self.N_hexagons, self.N_pin = S["N_verts"], S["N_hexagons"], S["N_pin"]
        self.hexs = S["hexs"]
        self.color_vertex_num = S["color_vertex_num"]
        self.v2i = S["vertex2index"]
        self.i2v = S["index2vertex"]
        self.pat = S["pattern"]
        self.sfg = S["shapeFuncGrad"].astype(dtype)
        L = self.layer
        cast = lambda m: dict(m, values=m["values"].astype(dtype))
        for name in ("Ut", "Us", "Ut_noOrder", "Us_noOrder", "Ut_hat", "Us_hat", "Ut_noOrder_hat", "Us_noOrder_hat"):
            setattr(self, name, [cast(m) for m in S[name]])
        # object.py:727-731 rest positions of every level
        self.x_gpu_layer = [S["x_rest"].astype(dtype)]
        for i in range(1, L):
            xl = np.zeros((self.dims[i], 3), dtype)
            bsr_mv(self.Ut_noOrder[i - 1], self.x_gpu_layer[i - 1], xl, alpha=1.0, beta=0.0)
            self.x_gpu_layer.append(xl)
        # object.py:756-759 rest data
        self.det, self.inv, self.vol, self.m = [], [], [], []
        for i in range(L):
            d, iv = K.prepare_kernal(self.x_gpu_layer[i], self.hexs[i], self.sfg, dtype)
            self.det.append(d); self.inv.append(iv)
            self.vol.append(K.sum_8(d))
            self.m.append(K.prepare_mass(self.vol[i], self.hexs[i], self.dims[i]))
        # object.py:761-795 pins
        self.x_gpu_layer[0] = S["x0"].astype(dtype)
        self.x_cpu = S["x0"].astype(dtype)                 # shares memory with x in the reference (:727)
        self.all_pin_pos = S["x0"].astype(dtype)
        self.pin, self.pin_list, self.pin_pos = S["pin"], S["pin_list"], S["pin_pos"].astype(dtype)
        self.MF_value_fixed = S["MF_value_fixed"].astype(dtype)
        self.MF_value = np.zeros_like(self.MF_value_fixed)
        self.UtAUs_values = [np.zeros((self.pat[i + 1]["nnz"], 3, 3), dtype) for i in range(L - 1)]
        z = lambda i: np.zeros((self.dims[i], 3), dtype)
        self.dev_E = [z(i) for i in range(L)]; self.dev_R = [z(i) for i in range(L)]
        self.dev_P = [z(i) for i in range(L)]; self.dev_AP = [z(i) for i in range(L)]
        self.dev_B = [z(i) for i in range(L)]; self.dev_B_fixed = [z(i) for i in range(L)]
        self.dev_delta_x = [z(i) for i in range(L)]; self.dev_x_solved = [z(i) for i in range(L)]
        self.grad = z(0)
        self.A = [None] * L
        self.Lm = [None] * L; self.Um = [None] * L; self.Dm = [None] * L
        self.plot_x, self.plot_InfNorm, self.plot_energy = [], [], []
        self.plot_InfNorm_newton, self.plot_energy_newton = [], []
        self.plot_InfNorm_newtonMultigrid, self.plot_energy_newtonMultigrid = [], []
        self.error_log = []
        self.not_spd = 0

    # ------------------------------------------------------------------ helpers
    def values(self, layer):
        return self.MF_value if layer == 0 else self.UtAUs_values[layer - 1]

    def set_LUD(self, layer):
        """The three bsr_set_from_triplets(L/U/D ...) calls, e.g. object.py:1255-1257 / 1262-1264."""
        p = self.pat[layer]; v = self.values(layer); n = self.dims[layer]
        self.Lm[layer] = bsr_from_triplets(n, n, p["L_row"], p["L_col"], v[p["off_l"]:p["off_l"] + p["L_nnz"]])
        self.Um[layer] = bsr_from_triplets(n, n, p["U_row"], p["U_col"], v[p["off_u"]:p["off_u"] + p["U_nnz"]])
        self.Dm[layer] = bsr_from_triplets(n, n, p["D_row"], p["D_col"], v[p["off_d"]:p["off_d"] + p["D_nnz"]])

    def force(self, x, out, gravity=True, pins=True):
        """The 3-launch force evaluation, e.g. object.py:1234-1236 (index order)."""
        K.elastic_force(x, self.hexs[0], self.v2i[0], self.sfg, self.det[0], self.inv[0], self.mu, self.la, out, self.dtype)
        if gravity:
            K.gravity_force(self.m[0], self.g, out, self.i2v[0])
        if pins:
            K.fixed_force(x, self.v2i[0], self.pin_list, out, self.pin_pos, self.control_mag)
        return out

    def hessian(self, layer, x, values):
        K.elastic_hessian(x, self.hexs[layer], self.sfg, self.det[layer], self.inv[layer], self.mu, self.la,
                          values, self.pat[layer]["hex_update_offset"], self.dtype)

    def total_energy(self, x):
        """object.py:1241-1243."""
        e = K.compute_elastic_energy(x, self.hexs[0], self.sfg, self.det[0], self.inv[0], self.mu, self.la, self.dtype)
        e = e + K.compute_gravity_energy(x, self.m[0], self.g, self.pin, self.all_pin_pos, self.control_mag, self.dtype)
        return self.dtype(e)

    def update_deltaX(self, x, dx):
        """cublas.py:20-24: x[i2v[idx]] += deltaX[idx]."""
        x[self.i2v[0]] = x[self.i2v[0]] + dx

    def _solve_color(self, layer, base, n):
        p = self.pat[layer]; v = self.values(layer)
        D = v[p["off_d"] + base:p["off_d"] + base + n]
        xs, ns = K.solve3x3(D, self.dev_B[layer][base:base + n], self.dev_delta_x[layer][base:base + n])
        self.not_spd += ns
        self.dev_delta_x[layer][base:base + n] = xs

    # ------------------------------------------------------------------ smoothers
    def PerformJacobi(self, layer=0, iterations=2):
        """object.py:869-884."""
        self.dev_delta_x[layer][...] = 0
        n = self.dims[layer]
        for _ in range(iterations):
            self.dev_B[layer][...] = self.dev_B_fixed[layer]
            bsr_mv(self.Lm[layer], self.dev_delta_x[layer], self.dev_B[layer], alpha=-1.0, beta=1.0)
            bsr_mv(self.Um[layer], self.dev_delta_x[layer], self.dev_B[layer], alpha=-1.0, beta=1.0)
            self.dev_delta_x[layer][...] = 0
            D = self.MF_value[self.pat[0]["off_d"]:self.pat[0]["off_d"] + n] if layer == 0 else self.Dm[layer]["values"]
            xs, ns = K.solve3x3(D, self.dev_B[layer], self.dev_delta_x[layer])
            self.not_spd += ns
            self.dev_delta_x[layer][...] = xs

    def PerformGaussSeidel(self, layer=0, iterations=10):
        """object.py:886-929 with the per-colour U/L slices (MF_GS_* / UtAUs_GS_*)."""
        p = self.pat[layer]; v = self.values(layer); cvn = self.color_vertex_num[layer]
        dx = self.dev_delta_x[layer]; B = self.dev_B[layer]
        dx[...] = 0
        for _ in range(iterations):
            B[...] = self.dev_B_fixed[layer]
            bsr_mv(self.Lm[layer], dx, B, alpha=-1.0, beta=1.0)
            dx[...] = 0
            for color in range(7, -1, -1):
                base = cvn[color]
                self._solve_color(layer, base, cvn[color + 1] - base)
                if color:
                    s0, s1 = p["GS_U_Ptr"][color - 1], p["GS_U_Ptr"][color]
                    nrow = cvn[color] - cvn[color - 1]
                    ncol = cvn[8] - cvn[color]
                    Uc = bsr_from_triplets(nrow, ncol, p["GS_U_row"][s0:s1], p["GS_U_col"][s0:s1],
                                           v[p["off_u"] + s0:p["off_u"] + s1])
                    bsr_mv(Uc, dx, B, x_offset=base, y_offset=cvn[color - 1], alpha=-1.0, beta=1.0)
            B[...] = self.dev_B_fixed[layer]
            bsr_mv(self.Um[layer], dx, B, alpha=-1.0, beta=1.0)
            dx[...] = 0
            for color in range(8):
                base = cvn[color]
                self._solve_color(layer, base, cvn[color + 1] - base)
                if color < 7:
                    s0, s1 = p["GS_L_Ptr"][color], p["GS_L_Ptr"][color + 1]
                    nrow = cvn[color + 2] - cvn[color + 1]
                    ncol = cvn[color + 1]
                    Lc = bsr_from_triplets(nrow, ncol, p["GS_L_row"][s0:s1], p["GS_L_col"][s0:s1],
                                           v[p["off_l"] + s0:p["off_l"] + s1])
                    bsr_mv(Lc, dx, B, x_offset=0, y_offset=cvn[color + 1], alpha=-1.0, beta=1.0)

    def _A_mv(self, layer, x, y, alpha, beta):
        bsr_mv(self.Lm[layer], x, y, alpha=alpha, beta=beta)
        bsr_mv(self.Um[layer], x, y, alpha=alpha, beta=1.0)
        bsr_mv(self.Dm[layer], x, y, alpha=alpha, beta=1.0)

    def _cg_loop(self, layer, iterations, tol):
        t = self.dtype
        B, Pv, AP, dx = self.dev_B[layer], self.dev_P[layer], self.dev_AP[layer], self.dev_delta_x[layer]
        r1 = t((B * B).sum(dtype=t))
        if r1 < EPSILON:
            return 0
        r = r1
        r0 = t(0)
        k = 1
        while r1 > tol * r and k <= iterations:
            if k > 1:
                beta = r1 / r0
                Pv[...] = Pv * t(beta)
                Pv[...] = Pv + t(1.0) * B
            else:
                Pv[...] = B
            AP[...] = 0
            bsr_mv(self.Lm[layer], Pv, AP, alpha=1.0, beta=1.0)
            bsr_mv(self.Um[layer], Pv, AP, alpha=1.0, beta=1.0)
            bsr_mv(self.Dm[layer], Pv, AP, alpha=1.0, beta=1.0)
            dot = t((Pv * AP).sum(dtype=t))
            if dot < 1e-10:
                break
            alpha = t(r1 / dot)
            dx[...] = dx + alpha * Pv
            B[...] = B + (-alpha) * AP
            r0 = r1
            r1 = t((B * B).sum(dtype=t))
            k += 1
        return k - 1

    def PerformConjugateGradient(self, layer=0, iterations=10, tol=1e-5):
        """object.py:931-973."""
        self.dev_delta_x[layer][...] = 0
        self.dev_B[layer][...] = self.dev_B_fixed[layer]
        return self._cg_loop(layer, iterations, tol)

    def PerformConjugateGradient_WithInitX(self, layer=0, iterations=10, tol=1e-5):
        """object.py:975-1019."""
        self.dev_B[layer][...] = self.dev_B_fixed[layer]
        bsr_mv(self.Lm[layer], self.dev_delta_x[layer], self.dev_B[layer], alpha=-1.0, beta=1.0)
        bsr_mv(self.Um[layer], self.dev_delta_x[layer], self.dev_B[layer], alpha=-1.0, beta=1.0)
        bsr_mv(self.Dm[layer], self.dev_delta_x[layer], self.dev_B[layer], alpha=-1.0, beta=1.0)
        return self._cg_loop(layer, iterations, tol)

    # ------------------------------------------------------------------ grid transfer
    def downSample(self, layer=0):
        """object.py:1021-1027."""
        self.dev_R[layer][...] = self.dev_B_fixed[layer]
        self.dev_x_solved[layer][...] = self.dev_delta_x[layer]
        bsr_mv(self.Lm[layer], self.dev_delta_x[layer], self.dev_R[layer], alpha=-1.0, beta=1.0)
        bsr_mv(self.Um[layer], self.dev_delta_x[layer], self.dev_R[layer], alpha=-1.0, beta=1.0)
        bsr_mv(self.Dm[layer], self.dev_delta_x[layer], self.dev_R[layer], alpha=-1.0, beta=1.0)
        bsr_mv(self.Ut_hat[layer], self.dev_R[layer], self.dev_B_fixed[layer + 1], alpha=1.0, beta=0.0)

    def upSample(self, layer=1):
        """object.py:1029-1035."""
        self.dev_x_solved[layer][...] = self.dev_x_solved[layer] + self.dev_delta_x[layer]
        bsr_mv(self.Us_hat[layer - 1], self.dev_x_solved[layer], self.dev_x_solved[layer - 1], alpha=1.0, beta=1.0)
        bsr_mv(self.Lm[layer - 1], self.dev_x_solved[layer - 1], self.dev_B_fixed[layer - 1], alpha=-1.0, beta=1.0)
        bsr_mv(self.Dm[layer - 1], self.dev_x_solved[layer - 1], self.dev_B_fixed[layer - 1], alpha=-1.0, beta=1.0)
        bsr_mv(self.Um[layer - 1], self.dev_x_solved[layer - 1], self.dev_B_fixed[layer - 1], alpha=-1.0, beta=1.0)

    def finish(self):
        """object.py:1037-1040."""
        self.dev_x_solved[0][...] = self.dev_x_solved[0] + self.dev_delta_x[0]
        for i in range(1, self.layer):
            self.dev_x_solved[i][...] = 0

    def showError(self, layer):
        """object.py:1048-1065 (diagnostic; norms are logged instead of printed)."""
        self.dev_B[layer][...] = self.dev_B_fixed[layer]
        before = K.inf_norm(self.dev_B[layer])
        self._A_mv_sub(layer)
        after = K.inf_norm(self.dev_B[layer])
        self.error_log.append((layer, float(before), float(after)))

    def _A_mv_sub(self, layer):
        bsr_mv(self.Lm[layer], self.dev_delta_x[layer], self.dev_B[layer], alpha=-1.0, beta=1.0)
        bsr_mv(self.Um[layer], self.dev_delta_x[layer], self.dev_B[layer], alpha=-1.0, beta=1.0)
        bsr_mv(self.Dm[layer], self.dev_delta_x[layer], self.dev_B[layer], alpha=-1.0, beta=1.0)

    def VCycle(self, layer):
        """object.py:1216-1226."""
        self.PerformGaussSeidel(layer, iterations=3)
        self.showError(layer)
        if layer < self.layer - 1:
            self.downSample(layer)
            self.VCycle(layer + 1)
            self.upSample(layer + 1)
        if layer == self.layer - 1:
            return 0
        self.PerformGaussSeidel(layer, iterations=3)
        self.showError(layer)

    # ------------------------------------------------------------------ non-linear drivers
    def _sample(self, step, iterations, grad, tag="", every=100):
        """The 'step % (iterations/100) == 0' block, e.g. object.py:1238-1248."""
        if step % (iterations / every) == 0:
            e = self.total_energy(self.x_gpu_layer[0])
            n = K.inf_norm(grad)
            if tag == "":
                self.plot_x.append(step); self.plot_InfNorm.append(n); self.plot_energy.append(e)
            elif tag == "newton":
                self.plot_x.append(step); self.plot_InfNorm_newton.append(np.log10(n)); self.plot_energy_newton.append(e)
            else:
                self.plot_InfNorm_newtonMultigrid.append(np.log10(n)); self.plot_energy_newtonMultigrid.append(e)

    def galerkin(self):
        """object.py:1252-1264: A_0 from triplets, A_i = Ut_hat (A_{i-1} Us_hat), spd_matrix33f, re-layout."""
        for i in range(self.layer):
            if i == 0:
                p = self.pat[0]
                self.A[0] = bsr_from_triplets(self.dims[0], self.dims[0], p["row"], p["col"], self.MF_value)
                self.set_LUD(0)
            else:
                self.A[i] = bsr_mm(self.Ut_hat[i - 1], bsr_mm(self.A[i - 1], self.Us_hat[i - 1]))
                nn = self.pat[i]["nnz"]
                assert self.A[i]["nnz"] == nn, "Galerkin pattern != coarse hex adjacency (block_values_reorder relies on it)"
                self.A[i]["values"][:nn] = K.spd_matrix33f(self.A[i]["values"][:nn], self.spd_value)
                self.UtAUs_values[i - 1][...] = self.A[i]["values"][self.pat[i]["block_offset"]]
                self.set_LUD(i)

    def NewtonMultigrid(self, iterations=100):
        """object.py:1230-1296."""
        x = self.x_gpu_layer[0]
        for step in range(1, iterations + 1):
            self.grad[...] = 0
            self.MF_value[...] = self.MF_value_fixed
            self.force(x, self.grad)
            self._sample(step, iterations, self.grad)
            self.hessian(0, x, self.MF_value)
            self.galerkin()
            self.dev_B_fixed[0][...] = self.grad
            self.VCycle(0)
            self.finish()
            self.update_deltaX(x, self.dev_x_solved[0])

    def Newton(self, iterations=1000):
        """object.py:1127-1160."""
        x = self.x_gpu_layer[0]
        for step in range(1, iterations + 1):
            self.grad[...] = 0
            self.MF_value[...] = self.MF_value_fixed
            self.force(x, self.grad)
            self.hessian(0, x, self.MF_value)
            self._sample(step, iterations, self.grad)
            self.set_LUD(0)
            self.dev_B_fixed[0][...] = self.grad
            self.PerformConjugateGradient(layer=0, iterations=5)
            self.update_deltaX(x, self.dev_delta_x[0])

    def _fas_fine(self, x, spd, gravity_pins, fixed_diag):
        """Common head of FASv0-v3 / compare: object.py:1311-1328 (v0/v1) and :1468-1487 (v2/v3)."""
        self.dev_B_fixed[0][...] = 0
        self.dev_B_fixed[1][...] = 0
        if fixed_diag:
            self.MF_value[...] = self.MF_value_fixed
        else:
            self.MF_value[...] = 0
        self.UtAUs_values[0][...] = 0
        self.hessian(0, x, self.MF_value)
        if spd:
            self.MF_value[...] = K.spd_matrix33f(self.MF_value, self.spd_value)
        g, pn = gravity_pins if isinstance(gravity_pins, tuple) else (gravity_pins, gravity_pins)
        self.force(x, self.dev_B_fixed[0], gravity=g, pins=pn)

    def _fas_coarse_setup(self, x, spd, gravity_pins, coarse_pin_penalty):
        """object.py:1489-1504: restrict state and re-evaluated residual, re-discretised coarse Hessian."""
        bsr_mv(self.Ut_noOrder[0], x, self.dev_x_solved[1], alpha=1.0, beta=0.0)
        self.dev_R[0][...] = 0
        g, pn = gravity_pins if isinstance(gravity_pins, tuple) else (gravity_pins, gravity_pins)
        self.force(x, self.dev_R[0], gravity=g, pins=pn)
        bsr_mv(self.Ut_hat[0], self.dev_R[0], self.dev_B_fixed[1], alpha=1.0, beta=0.0)
        self.hessian(1, self.dev_x_solved[1], self.UtAUs_values[0])
        if coarse_pin_penalty:
            K.compute_fix_hessian(self.v2i[1], self.pin, self.control_mag, self.pat[1]["off_d"], self.UtAUs_values[0],
                                  self.S["fix_idx"], self.S["fix_values"])
        if spd:
            self.UtAUs_values[0][...] = K.spd_matrix33f(self.UtAUs_values[0], self.spd_value)
        self.set_LUD(1)

    def _fas_v123_step(self, x, spd, gravity_pins, fixed_diag, coarse_pin_penalty, show):
        self._fas_fine(x, spd, gravity_pins, fixed_diag)
        self.set_LUD(0)
        self.PerformGaussSeidel(layer=0, iterations=1)
        if show:
            self.showError(0)
        self.update_deltaX(x, self.dev_delta_x[0])
        self._fas_coarse_setup(x, spd, gravity_pins, coarse_pin_penalty)
        self.PerformJacobi(layer=1, iterations=1)
        bsr_mv(self.Us_hat[0], self.dev_delta_x[1], self.dev_delta_x[0], alpha=1.0, beta=0.0)
        self.update_deltaX(x, self.dev_delta_x[0])

    def FASv0(self, iterations=100):
        """object.py:1308-1370 (the only variant with the FAS right-hand side R r + A_c x_c)."""
        self.dev_x_solved[0][...] = self.x_gpu_layer[0]
        x = self.dev_x_solved[0]
        t = self.dtype
        for _ in range(iterations):
            self._fas_fine(x, spd=True, gravity_pins=False, fixed_diag=False)
            self.set_LUD(0)
            self.PerformGaussSeidel(layer=0, iterations=1)
            self.showError(0)
            self.update_deltaX(x, self.dev_delta_x[0])
            self._fas_coarse_setup(x, spd=True, gravity_pins=False, coarse_pin_penalty=False)
            self.dev_delta_x[1][self.v2i[1]] = self.dev_x_solved[1]                      # V2I, :1350
            bsr_mv(self.Lm[1], self.dev_delta_x[1], self.dev_B_fixed[1], alpha=1.0, beta=1.0)
            bsr_mv(self.Um[1], self.dev_delta_x[1], self.dev_B_fixed[1], alpha=1.0, beta=1.0)
            bsr_mv(self.Dm[1], self.dev_delta_x[1], self.dev_B_fixed[1], alpha=1.0, beta=1.0)
            self.PerformConjugateGradient_WithInitX(layer=1, iterations=10)
            self.dev_B[1][self.i2v[1]] = self.dev_delta_x[1]                             # I2V, :1357
            self.dev_E[1][...] = t(-1.0) * self.dev_x_solved[1] + t(1.0) * self.dev_B[1]  # z_axpby, :1361
            bsr_mv(self.Us_noOrder_hat[0], self.dev_E[1], self.dev_E[0], alpha=1.0, beta=0.0)
            x[...] = t(1.0) * x + t(1.0) * self.dev_E[0]                                  # axpby, :1364
            self.showError(1)

    # ---- FASv0 generalised to more than two levels (no counterpart in the reference beyond levels=2) ----
    def _ml_smooth(self, l, nu):
        """y_l += GS_nu(A_l, b_l - A_l y_l) with the reference's PerformGaussSeidel (which starts from zero)."""
        b, y = self.dev_B_fixed[l].copy(), self.dev_delta_x[l].copy()
        r = b.copy()
        self._A_mv(l, y, r, alpha=-1.0, beta=1.0)
        self.dev_B_fixed[l][...] = r
        self.PerformGaussSeidel(layer=l, iterations=nu)
        self.dev_B_fixed[l][...] = b
        self.dev_delta_x[l][...] = y + self.dev_delta_x[l]

    def _ml_level(self, l, last, nu):
        t = self.dtype
        if l == last:
            self.PerformConjugateGradient_WithInitX(layer=l, iterations=10)
            return
        self._ml_smooth(l, nu)
        self.dev_B[l][self.i2v[l]] = self.dev_delta_x[l]
        bsr_mv(self.Ut_noOrder[l], self.dev_B[l], self.dev_x_solved[l + 1], alpha=1.0, beta=0.0)
        self.dev_R[l][...] = self.dev_B_fixed[l]
        self._A_mv(l, self.dev_delta_x[l], self.dev_R[l], alpha=-1.0, beta=1.0)
        bsr_mv(self.Ut_hat[l], self.dev_R[l], self.dev_B_fixed[l + 1], alpha=1.0, beta=0.0)
        self.UtAUs_values[l][...] = 0
        self.hessian(l + 1, self.dev_x_solved[l + 1], self.UtAUs_values[l])
        self.UtAUs_values[l][...] = K.spd_matrix33f(self.UtAUs_values[l], self.spd_value)
        self.set_LUD(l + 1)
        self.dev_delta_x[l + 1][self.v2i[l + 1]] = self.dev_x_solved[l + 1]
        self._A_mv(l + 1, self.dev_delta_x[l + 1], self.dev_B_fixed[l + 1], alpha=1.0, beta=1.0)
        self._ml_level(l + 1, last, nu)
        self.dev_B[l + 1][self.i2v[l + 1]] = self.dev_delta_x[l + 1]
        self.dev_E[l + 1][...] = t(-1.0) * self.dev_x_solved[l + 1] + t(1.0) * self.dev_B[l + 1]
        bsr_mv(self.Us_noOrder_hat[l], self.dev_E[l + 1], self.dev_E[l], alpha=1.0, beta=0.0)
        d = np.zeros_like(self.dev_delta_x[l])
        d[self.v2i[l]] = self.dev_E[l]
        self.dev_delta_x[l][...] = self.dev_delta_x[l] + d
        self._ml_smooth(l, nu)

    def FAS_multilevel(self, iterations=100, levels=None, smoothing=1):
        """Twin of Object.FAS_multilevel (fem-simulation_b200/object.py); levels=2 is FASv0 statement for statement."""
        L = self.layer if levels is None else int(levels)
        self.dev_x_solved[0][...] = self.x_gpu_layer[0]
        x = self.dev_x_solved[0]
        t = self.dtype
        for _ in range(iterations):
            self._fas_fine(x, spd=True, gravity_pins=False, fixed_diag=False)
            self.set_LUD(0)
            self.PerformGaussSeidel(layer=0, iterations=1)
            self.update_deltaX(x, self.dev_delta_x[0])
            self._fas_coarse_setup(x, spd=True, gravity_pins=False, coarse_pin_penalty=False)
            self.dev_delta_x[1][self.v2i[1]] = self.dev_x_solved[1]
            self._A_mv(1, self.dev_delta_x[1], self.dev_B_fixed[1], alpha=1.0, beta=1.0)
            self._ml_level(1, L - 1, int(smoothing))
            self.dev_B[1][self.i2v[1]] = self.dev_delta_x[1]
            self.dev_E[1][...] = t(-1.0) * self.dev_x_solved[1] + t(1.0) * self.dev_B[1]
            bsr_mv(self.Us_noOrder_hat[0], self.dev_E[1], self.dev_E[0], alpha=1.0, beta=0.0)
            x[...] = t(1.0) * x + t(1.0) * self.dev_E[0]

    def FASv1(self, iterations=100):
        """object.py:1372-1415."""
        self.dev_x_solved[0][...] = self.x_gpu_layer[0]
        for _ in range(iterations):
            self._fas_v123_step(self.dev_x_solved[0], spd=True, gravity_pins=False, fixed_diag=False,
                                coarse_pin_penalty=False, show=True)

    def FASv2(self, iterations=100):
        """object.py:1418-1463."""
        for _ in range(iterations):
            self._fas_v123_step(self.x_gpu_layer[0], spd=False, gravity_pins=True, fixed_diag=True,
                                coarse_pin_penalty=False, show=True)

    def FASv3(self, iterations=100):
        """object.py:1466-1511 (== the per-frame body of render(), :1167-1210)."""
        for _ in range(iterations):
            self._fas_v123_step(self.x_gpu_layer[0], spd=False, gravity_pins=True, fixed_diag=True,
                                coarse_pin_penalty=True, show=False)

    def compare(self, iterations=100):
        """object.py:1514-1598 (the two loops; plotting dropped)."""
        x = self.x_gpu_layer[0]
        for step in range(1, iterations + 1):
            self._fas_fine(x, spd=False, gravity_pins=True, fixed_diag=True)
            self._sample(step, iterations, self.dev_B_fixed[0], "newton")
            self.set_LUD(0)
            self.PerformGaussSeidel(layer=0, iterations=1)
            self.update_deltaX(x, self.dev_delta_x[0])
        x[...] = self.x_cpu                                                               # :1546
        for step in range(1, iterations + 1):
            self._fas_fine(x, spd=False, gravity_pins=True, fixed_diag=True)
            self._sample(step, iterations, self.dev_B_fixed[0], "fas")
            self.set_LUD(0)
            self.PerformGaussSeidel(layer=0, iterations=1)
            self.update_deltaX(x, self.dev_delta_x[0])
            self._fas_coarse_setup(x, spd=False, gravity_pins=True, coarse_pin_penalty=False)
            self.PerformJacobi(layer=1, iterations=1)
            bsr_mv(self.Us_hat[0], self.dev_delta_x[1], self.dev_delta_x[0], alpha=1.0, beta=0.0)
            self.update_deltaX(x, self.dev_delta_x[0])

    def compare_withoutGravity(self, iterations=100):
        """object.py:1630-1741: the two-level FAS body without gravity (pin springs only), first with the coarse pin
        penalty ("add attachment"), then from the rest state again without it; sampled every iterations/10 steps."""
        x = self.x_gpu_layer[0]
        for tag, penalty in (("newton", True), ("without", False)):
            if tag == "without":
                x[...] = self.x_cpu                                                       # :1686
            for step in range(1, iterations + 1):
                self._fas_fine(x, spd=False, gravity_pins=(False, True), fixed_diag=True)
                self._sample(step, iterations, self.dev_B_fixed[0], tag, every=10)
                self.set_LUD(0)
                self.PerformGaussSeidel(layer=0, iterations=1)
                self.update_deltaX(x, self.dev_delta_x[0])
                self._fas_coarse_setup(x, spd=False, gravity_pins=(False, True), coarse_pin_penalty=penalty)
                self.PerformJacobi(layer=1, iterations=1)
                bsr_mv(self.Us_hat[0], self.dev_delta_x[1], self.dev_delta_x[0], alpha=1.0, beta=0.0)
                self.update_deltaX(x, self.dev_delta_x[0])

    def _drag_linearise(self):
        x = self.x_gpu_layer[0]
        self.MF_value[...] = self.MF_value_fixed
        self.grad[...] = 0
        self.force(x, self.grad, gravity=False, pins=True)
        self.hessian(0, x, self.MF_value)
        self.MF_value[...] = K.spd_matrix33f(self.MF_value, self.spd_value)
        self.set_LUD(0)
        self.dev_B_fixed[0][...] = self.grad

    def _residual_norm(self, layer):
        self.dev_B[layer][...] = self.dev_B_fixed[layer]
        self._A_mv_sub(layer)
        return float(K.inf_norm(self.dev_B[layer]))

    def drag(self, n=10):
        """object.py:1770-1852 without the plots: log10 of the residual infinity norm after i = 0..n-1 iterations of
        Gauss-Seidel, of CG, and after i V-cycles (Newton-MG), linearised at the rest state with the pin springs only."""
        out = {"GS": [], "CG": [], "NewtonMG": []}
        for i in range(n):
            self._drag_linearise()
            self.PerformGaussSeidel(layer=0, iterations=i)
            out["GS"].append(np.log10(self._residual_norm(0)))
        for i in range(n):
            self.x_gpu_layer[0][...] = self.x_cpu
            self._drag_linearise()
            self.PerformConjugateGradient(layer=0, iterations=i)
            out["CG"].append(np.log10(self._residual_norm(0)))
        for it in range(n):
            self.x_gpu_layer[0][...] = self.x_cpu
            self._drag_linearise()
            self.galerkin()
            norm = float(K.inf_norm(self.grad))
            for _ in range(it):
                self.VCycle(0)
                norm = self._residual_norm(0)
            self.finish()
            out["NewtonMG"].append(np.log10(norm))
        return out

    def Adam(self, iterations=1000, lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
        """object.py:1067-1094 with updateM / updateV / scal_ / updateX of cublas.py:172-194 (float32 like the kernels)."""
        t = self.dtype
        x = self.x_gpu_layer[0]
        m = np.zeros_like(x); v = np.zeros_like(x)
        for step in range(1, iterations + 1):
            g = np.zeros_like(x)
            self._force_vertex_order(x, g)
            self._sample(step, iterations, g, "")
            m = (t(1 - beta1) * g + t(beta1) * m).astype(t)
            v = (t(1 - beta2) * (g * g) + t(beta2) * v).astype(t)
            m_hat = (m * t(1 / (1 - beta1 ** step))).astype(t)
            v_hat = (v * t(1 / (1 - beta2 ** step))).astype(t)
            x -= (t(lr) * (m_hat / (np.sqrt(v_hat) + t(epsilon)))).astype(t)

    def _force_vertex_order(self, x, out):
        """The *_noOrder force kernels (quai_hexagon.py:135-158, 169-186): the index-order force scattered back."""
        gi = np.zeros_like(out)
        self.force(x, gi)
        out[self.i2v[0]] = gi

    def gradientDescent(self, iterations=100, lr=1e-3):
        """object.py:1097-1125: steepest descent on the vertex-order force with energy backtracking (x0.2, <= 20 tries); the
        energy it compares with is the one of the last SAMPLE (self.energy is only refreshed inside the sampling block)."""
        t = self.dtype
        x = self.x_gpu_layer[0]
        energy = getattr(self, "_gd_energy", t(0.0))
        for step in range(1, iterations + 1):
            g = np.zeros_like(x)
            self._force_vertex_order(x, g)
            if step % (iterations / 100) == 0:
                energy = self.total_energy(x)
                self.plot_x.append(step); self.plot_InfNorm.append(K.inf_norm(g)); self.plot_energy.append(energy)
            alpha = t(lr)
            for _ in range(20):
                trial = (x + alpha * g).astype(t)                                          # add_grad, cublas.py:143-147
                if self.total_energy(trial) < energy:
                    x[...] = trial
                    break
                alpha = t(alpha * t(0.2))
        self._gd_energy = energy
