"""ORACLE (test / CPU-baseline infrastructure only): ctypes driver of oracle/femb_oracle.c.

Builds an ``OrcCtx`` from the numpy arrays of a hierarchy (the same integer maps the product uses -- they
are validated bit-exactly against oracle/ref_setup.py) and rest data computed by the numpy oracle, and runs
whole Newton-multigrid / FASv3 steps on the host cores.  Only tests/, __graft_entry__ and bench.py's
CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
P = C.c_void_p


class OrcLevel(C.Structure):
    _fields_ = [("n_vert", C.c_int), ("n_hex", C.c_int), ("nnz", C.c_int), ("off_l", C.c_int), ("off_d", C.c_int),
                ("off_u", C.c_int), ("color_start", C.c_int * 9), ("hex", P), ("v2i", P), ("i2v", P), ("L_ptr", P),
                ("L_col", P), ("U_ptr", P), ("U_col", P), ("elem2nnz", P), ("invJ", P), ("detJ", P), ("mass", P),
                ("values", P), ("B", P), ("B_fixed", P), ("delta_x", P), ("x_solved", P), ("R", P),
                ("vert_elem", P), ("elem_k", P), ("elem_f", P)]


class OrcCsr(C.Structure):
    _fields_ = [("nrow", C.c_int), ("ncol", C.c_int), ("nnz", C.c_int), ("offsets", P), ("columns", P), ("weights", P)]


class OrcPlan(C.Structure):
    _fields_ = [("t_nnz", C.c_int), ("t_ptr", P), ("t_col", P), ("c_ptr", P), ("c_col", P), ("c_slot", P), ("t_val", P),
                ("c_val", P)]


def _ctx_type(real):
    class OrcCtx(C.Structure):
        _fields_ = [("layers", C.c_int), ("lv", C.POINTER(OrcLevel)), ("Ut_hat", C.POINTER(OrcCsr)), ("Us_hat", C.POINTER(OrcCsr)),
                    ("Ut_noOrder", C.POINTER(OrcCsr)), ("Ut_noOrder_hat", C.POINTER(OrcCsr)), ("plan", C.POINTER(OrcPlan)),
                    ("x", P), ("grad", P), ("pin", P), ("pin_pos", P), ("diag_fixed", P), ("mu", real), ("la", real),
                    ("g", real), ("control_mag", real), ("spd_value", real), ("sfg", real * 192)]
    return OrcCtx


OrcCtx = _ctx_type(C.c_float)
OrcCtx64 = _ctx_type(C.c_double)

_libs = {}


def load(native: bool = True, double: bool = False):
    """Compile (host-tuned when ``native``) and load the C oracle; ``double`` = the float64 twin (-DORC_DOUBLE)."""
    key = (native, double)
    if key in _libs:
        return _libs[key]
    tag = "portable"
    flags = "-O3 -fopenmp -fPIC -std=c11 -Wno-unknown-pragmas" + (" -DORC_DOUBLE" if double else "")
    if native:
        try:
            cpu = [l for l in open("/proc/cpuinfo") if l.startswith("flags")][0]
        except Exception:
            cpu = "unknown"
        tag = hashlib.md5(cpu.encode()).hexdigest()[:10]
        flags += " -march=native"
    out = os.path.join(_DIR, "_build", f"libfemb_oracle_{tag}{'_f64' if double else ''}.so")
    src = os.path.join(_DIR, "femb_oracle.c")
    stale = lambda: not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src)
    if stale():
        import fcntl
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out + ".lock", "w") as lock:                  # several ranks may get here at once: one compiles, atomically
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if stale():
                    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
                    tmp = f"{out}.{os.getpid()}.tmp"
                    subprocess.run([cc] + flags.split() + ["-shared", "-o", tmp, src, "-lm"], check=True)
                    os.replace(tmp, out)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    lib = C.CDLL(out)
    lib.orc_num_threads.restype = C.c_int
    lib.orc_energy.restype = C.c_double if double else C.c_float
    lib.orc_inf_norm.restype = C.c_double if double else C.c_float
    _libs[key] = lib
    return lib


def _p(a):
    return a.ctypes.data_as(P)


def vertex_incidence(hexagons, n_vert):
    """[n_vert, 8] int32: e*8+a of the hexahedra incident to every vertex in ascending order, -1 padded."""
    ve = np.full((n_vert, 8), -1, dtype=np.int32)
    flat = np.asarray(hexagons).reshape(-1).astype(np.int64)
    o2 = np.argsort(flat, kind="stable")
    cnt = np.bincount(flat, minlength=n_vert)
    start = np.cumsum(cnt) - cnt
    ve[flat[o2], np.arange(len(o2)) - start[flat[o2]]] = o2.astype(np.int32)
    return ve


class COracle:
    """Owns the numpy buffers behind an OrcCtx.  ``H`` is a fem_simulation_b200.hierarchy.Hierarchy (integer maps
    only), ``rest`` a oracle.ref_solvers.RefObject-like provider of det/inv/mass/x0/pins (numpy oracle)."""

    def __init__(self, H, plans, x0, det, inv, mass, pin, pin_pos, diag_fixed, mu=250.0, la=0.0, g=-4.9, control_mag=100.0,
                 spd_value=1e-3, native=True, dtype=np.float32):
        self.double = np.dtype(dtype) == np.float64
        self.lib = load(native, self.double)
        self.keep = []
        rt = np.float64 if self.double else np.float32
        k = lambda a, dt: self._keep(np.ascontiguousarray(a, dtype=(rt if dt == np.float32 else dt)))
        L = H.layer
        self.L = L
        self.levels = (OrcLevel * L)()
        self.values, self.vec = [], []
        for l in range(L):
            p = H.pattern[l]
            lv = self.levels[l]
            lv.n_vert, lv.n_hex, lv.nnz = H.dims[l], H.hexs[l].shape[0], p.nnz
            lv.off_l, lv.off_d, lv.off_u = p.off_l, p.off_d, p.off_u
            for c in range(9):
                lv.color_start[c] = H.color_vertex_num[l][c]
            lv.hex = _p(k(H.hexs[l], np.int32)); lv.v2i = _p(k(H.vertex2index[l], np.int32)); lv.i2v = _p(k(H.index2vertex[l], np.int32))
            lv.L_ptr = _p(k(p.L_ptr, np.int32)); lv.L_col = _p(k(p.L_col, np.int32))
            lv.U_ptr = _p(k(p.U_ptr, np.int32)); lv.U_col = _p(k(p.U_col, np.int32))
            lv.elem2nnz = _p(k(p.hex_update_offset, np.int32))
            lv.invJ = _p(k(inv[l], np.float32)); lv.detJ = _p(k(det[l], np.float32)); lv.mass = _p(k(mass[l], np.float32))
            lv.vert_elem = _p(k(vertex_incidence(H.hexs[l], H.dims[l]), np.int32))
            lv.elem_k = _p(self._keep(np.zeros(H.hexs[l].shape[0] * 576, rt)))
            lv.elem_f = _p(self._keep(np.zeros(H.hexs[l].shape[0] * 24, rt)))
            vals = self._keep(np.zeros((p.nnz, 3, 3), rt)); self.values.append(vals); lv.values = _p(vals)
            vs = {}
            for nm in ("B", "B_fixed", "delta_x", "x_solved", "R"):
                vs[nm] = self._keep(np.zeros((H.dims[l], 3), rt)); setattr(lv, nm, _p(vs[nm]))
            self.vec.append(vs)

        def csr_array(lst):
            arr = (OrcCsr * max(len(lst), 1))()
            for i, m in enumerate(lst):
                arr[i].nrow, arr[i].ncol, arr[i].nnz = m.nrow, m.ncol, m.nnz
                arr[i].offsets = _p(k(m.offsets, np.int32)); arr[i].columns = _p(k(m.columns, np.int32)); arr[i].weights = _p(k(m.weights, np.float32))
            return arr
        self.Ut_hat, self.Us_hat = csr_array(H.Ut_hat), csr_array(H.Us_hat)
        self.Ut_noOrder, self.Ut_noOrder_hat = csr_array(H.Ut_noOrder), csr_array(H.Ut_noOrder_hat)
        self.plans = (OrcPlan * max(L - 1, 1))()
        self.has_plans = len(plans) >= L - 1                   # the Galerkin plans are needed by newton_mg_step only
        for l in range(L - 1 if self.has_plans else 0):
            g_ = plans[l]
            pl = self.plans[l]
            pl.t_nnz = len(g_["t_col"])
            pl.t_ptr = _p(k(g_["t_ptr"], np.int32)); pl.t_col = _p(k(g_["t_col"], np.int32))
            pl.c_ptr = _p(k(g_["c_ptr"], np.int32)); pl.c_col = _p(k(g_["c_col"], np.int32)); pl.c_slot = _p(k(g_["c_slot"], np.int32))
            pl.t_val = _p(self._keep(np.zeros(pl.t_nnz * 9, rt)))
            pl.c_val = _p(self._keep(np.zeros(len(g_["c_col"]) * 9, rt)))
        self.x = self._keep(np.ascontiguousarray(x0, rt).copy())
        self.x0 = self.x.copy()
        self.grad = self._keep(np.zeros_like(self.x))
        ctx = (OrcCtx64 if self.double else OrcCtx)()
        ctx.layers = L
        ctx.lv = C.cast(self.levels, C.POINTER(OrcLevel))
        ctx.Ut_hat = C.cast(self.Ut_hat, C.POINTER(OrcCsr)); ctx.Us_hat = C.cast(self.Us_hat, C.POINTER(OrcCsr))
        ctx.Ut_noOrder = C.cast(self.Ut_noOrder, C.POINTER(OrcCsr)); ctx.Ut_noOrder_hat = C.cast(self.Ut_noOrder_hat, C.POINTER(OrcCsr))
        ctx.plan = C.cast(self.plans, C.POINTER(OrcPlan))
        ctx.x, ctx.grad = _p(self.x), _p(self.grad)
        ctx.pin = _p(k(pin, np.int32)); ctx.pin_pos = _p(k(pin_pos, np.float32)); ctx.diag_fixed = _p(k(diag_fixed, np.float32))
        ctx.mu, ctx.la, ctx.g, ctx.control_mag, ctx.spd_value = mu, la, g, control_mag, spd_value
        sfg = np.ascontiguousarray(H.shapeFuncGrad, np.float32).reshape(-1)      # the table is float32 in the reference (object.py:701-722)
        for i in range(192):
            ctx.sfg[i] = float(sfg[i])
        self.ctx = ctx

    def _keep(self, a):
        self.keep.append(a)
        return a

    @property
    def threads(self):
        return int(self.lib.orc_num_threads())

    def newton_mg_step(self):
        assert self.has_plans, "COracle built without Galerkin plans"
        self.lib.orc_newton_mg_step(C.byref(self.ctx))

    def fasv3_step(self):
        self.lib.orc_fasv3_step(C.byref(self.ctx))

    def force(self, x, out, with_gravity_pins=True):
        self.lib.orc_force(C.byref(self.ctx), 0, _p(x), _p(out), 1 if with_gravity_pins else 0)

    def hessian(self, level, x, values, diag_add=None):
        self.lib.orc_hessian(C.byref(self.ctx), level, _p(x), _p(values), _p(diag_add) if diag_add is not None else None)

    def hessian_literal(self, level, x, values, diag_add=None):
        self.lib.orc_hessian_literal(C.byref(self.ctx), level, _p(x), _p(values), _p(diag_add) if diag_add is not None else None)

    def fasv2_step(self):
        self.lib.orc_fasv2_step(C.byref(self.ctx))

    def compare_newton_step(self):
        self.lib.orc_compare_newton_step(C.byref(self.ctx))

    def inf_norm(self, v):
        return float(self.lib.orc_inf_norm(_p(v), int(v.size)))

    def compare(self, iterations=100):
        """Object.compare (object.py:1514-1598) on the C oracle: returns dict(plot_x, energy_newton, lognorm_newton,
        energy_fas, lognorm_fas); sampling `step % (iterations/100) == 0`, energy before the step, |f|inf of the force
        evaluated at the start of the step."""
        out = dict(plot_x=[], energy_newton=[], lognorm_newton=[], energy_fas=[], lognorm_fas=[])
        for tag, fn in (("newton", self.compare_newton_step), ("fas", self.fasv2_step)):
            self.x[...] = self.x0                                 # object.py:1546
            for step in range(1, iterations + 1):
                sampled = step % (iterations / 100) == 0
                e = self.energy(self.x) if sampled else None
                fn()
                if sampled:
                    if tag == "newton":
                        out["plot_x"].append(step)
                    out["energy_" + tag].append(e)
                    out["lognorm_" + tag].append(float(np.log10(self.inf_norm(self.vec[0]["B_fixed"]))))
        return out

    def set_threads(self, n):
        """Use n OpenMP threads from now on (torchrun exports OMP_NUM_THREADS=1)."""
        self.lib.orc_set_threads(int(n))

    def energy(self, x):
        return float(self.lib.orc_energy(C.byref(self.ctx), _p(x)))


def from_ref(ref, H, plans, native=True):
    """COracle whose rest data / pins come from a numpy-oracle RefObject built on the same mesh."""
    diag = np.zeros(H.dims[0], np.float32)
    diag[H.vertex2index[0][ref.pin_list]] = ref.control_mag
    return COracle(H, plans, ref.S["x0"], ref.det, ref.inv, ref.m, ref.pin, ref.all_pin_pos, diag, ref.mu, ref.la, ref.g,
                   ref.control_mag, ref.spd_value, native)


def from_hierarchy(H, plans, points, dx, pin_list=None, mu=250.0, la=0.0, g=-4.9, control_mag=100.0, spd_value=1e-3,
                   native=True, dtype=np.float32):
    """COracle for meshes too large for the literal setup of oracle/ref_setup.py: integer maps come from the
    vectorised hierarchy (bit-exact with the literal one, tests/test_hierarchy.py), rest data from the numpy
    restatement of prepare_kernal / Sum_8 / prepare_mass (oracle/ref_kernels.py)."""
    from . import ref_kernels as K
    x0 = np.ascontiguousarray(points, np.float32).astype(dtype)
    xs = [x0]
    for l in range(1, H.layer):                               # object.py:730-731
        m = H.Ut_noOrder[l - 1]
        y = np.zeros((m.nrow, 3), dtype)
        np.add.at(y, m.rows(), m.weights[:, None] * xs[l - 1][m.columns])
        xs.append(y)
    det, inv, mass = [], [], []
    sfg = H.shapeFuncGrad
    for l in range(H.layer):                                  # object.py:756-759
        d, iv = K.prepare_kernal(xs[l], H.hexs[l], sfg, dtype=dtype)
        det.append(d); inv.append(iv)
        mass.append(K.prepare_mass(K.sum_8(d), H.hexs[l], H.dims[l]))
    if pin_list is None:
        y = x0[:, 1]
        pin_list = np.nonzero(y >= np.float32(np.float32(y.max() - np.float32(dx)) - np.float32(1e-5)))[0]
    pin = np.zeros(H.dims[0], np.int32); pin[pin_list] = 1
    diag = np.zeros(H.dims[0], np.float32); diag[H.vertex2index[0][pin_list]] = control_mag
    return COracle(H, plans, x0, det, inv, mass, pin, x0, diag, mu, la, g, control_mag, spd_value, native, dtype)
