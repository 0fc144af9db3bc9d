"""``Object`` -- host-side mirror of the reference simulation API on top of libfemb200.

Same constructor signature, method names, attribute names and order of operations as
``class Object`` in /root/reference/exp1/quasi_simulation/object.py (cited per method as
``object.py:LINE``); Warp arrays become ``torch.cuda`` tensors, every ``wp.launch`` / ``bsr_*``
call becomes one call through the C ABI of ``include/femb200.h`` on torch's current stream, the
GL window, pyvista and matplotlib parts are dropped.  The per-step
COO->BSR rebuilds of the reference disappear: all patterns are fixed at setup.

No CPU fallback: without libfemb200.so or without a CUDA device the constructor raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, hierarchy, mesh

EPSILON = 1e-7


class BsrView:
    """What drivers read from the reference's BsrMatrix: nrow, ncol, nnz, offsets, columns, values."""

    def __init__(self, nrow, ncol, offsets, columns, values_fn):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.offsets, self.columns = offsets, columns
        self._values_fn = values_fn

    @property
    def nnz(self):
        return int(self.columns.shape[0])

    @property
    def values(self):
        return self._values_fn()


class TransferMatrix(BsrView):
    """Interpolation operator with scalar weights on the device (+ the struct the C ABI takes)."""

    def __init__(self, sw: hierarchy.SparseW, device):
        self.weights = torch.from_numpy(sw.weights).to(device)
        offsets = torch.from_numpy(sw.offsets).to(device)
        columns = torch.from_numpy(sw.columns).to(device)
        eye = torch.eye(3, device=device)
        super().__init__(sw.nrow, sw.ncol, offsets, columns, lambda: self.weights[:, None, None] * eye)
        self.c = _lib.FembTransfer(sw.nrow, sw.ncol, sw.nnz, offsets.data_ptr(), columns.data_ptr(), self.weights.data_ptr())

    def set_weights(self, w):
        self.weights.copy_(w)


def _dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


class Object:
    def __init__(self, mesh_path, dx, pinList, *, device="cuda:0", layers=None, use_graphs=True, verbose=False,
                 control_mag=100.0, mu=250.0, la=0.0, g=-9.8 / 2.0, pin_displacement=0.1, spd_level0=False,
                 corrected_hierarchy=False, lattice_fold=True,
                 lattice_origin=None, use_c_steps=True, pins_exact=False):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.FembError("femb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        _lib.check(self.lib.femb_check_device(self.device.index or 0), "femb_check_device")
        torch.cuda.set_device(self.device)
        self.verbose = verbose
        self.use_graphs = use_graphs
        self.use_c_steps = use_c_steps        # step bodies as ONE C-ABI call (femb_newton_mg_step / femb_fas_step) instead of ~25
        self._graphs = {}
        # object.py:23-25, 743-748
        self.control_mag = float(control_mag)
        self.spd_value = 1e-3
        self.mu, self.la, self.g = float(mu), float(la), float(g)
        self.dx = dx
        self.spd_level0 = spd_level0           # exp1/simulation applies spd_matrix33f on level 0 too (object.py:1059,1125 there)

        points, hexagons = self._load_mesh(mesh_path, dx)
        self.N_verts, self.N_hexagons = int(points.shape[0]), int(hexagons.shape[0])
        H = hierarchy.build_hierarchy(points, hexagons, dx, layers, corrected_spacing=corrected_hierarchy, origin=lattice_origin)
        self.hierarchy = H
        self.layer = H.layer
        self.dims = H.dims
        self.hexs = [torch.from_numpy(h) for h in H.hexs]
        self.x = [torch.from_numpy(x.copy()) for x in H.x]
        self.fine_coarse = [torch.from_numpy(f) for f in H.fine_coarse]
        self.color_num = [8] * H.layer
        self.color_vertex_num = H.color_vertex_num
        self.box_min, self.box_max = torch.from_numpy(H.box_min), torch.from_numpy(H.box_max)
        dev = self.device
        L = self.layer

        # ---- per-level device data + C structs ------------------------------------------------
        self.dev_vertex2index = [_dev(a, dev) for a in H.vertex2index]
        self.dev_index2vertex = [_dev(a, dev) for a in H.index2vertex]
        self.hexagons_gpu = [_dev(a, dev) for a in H.hexs]
        self.hex_update_offset_gpu = [_dev(p.hex_update_offset, dev) for p in H.pattern]
        self._pat = []
        self.levels = []
        self.inverse_pX_peps_gpu, self.det_pX_peps_gpu, self.vol, self.m_gpu = [], [], [], []
        for l in range(L):
            p = H.pattern[l]
            n, nh = H.dims[l], H.hexs[l].shape[0]
            ve = hierarchy.vertex_incidence(H.hexs[l], n)                # e*8+a of the incident hexahedra, -1 padded
            pad = np.zeros(4, np.int32)                                  # bulk copies are rounded out to 16 bytes
            d = dict(L_ptr=_dev(p.L_ptr, dev), L_col=_dev(np.concatenate([p.L_col, pad]), dev)[:p.L_nnz],
                     U_ptr=_dev(p.U_ptr, dev), U_col=_dev(np.concatenate([p.U_col, pad]), dev)[:p.L_nnz],
                     vert_elem=_dev(ve, dev),
                     invJ=torch.zeros((nh, 8, 3, 3), device=dev), detJ=torch.zeros((nh, 8), device=dev))
            self._pat.append(d)
            lv = _lib.FembLevel()
            lv.n_vert, lv.n_hex, lv.nnz = n, nh, p.nnz
            lv.off_l, lv.off_d, lv.off_u = p.off_l, p.off_d, p.off_u
            for c in range(9):
                lv.color_start[c] = H.color_vertex_num[l][c]
            lv.hex = self.hexagons_gpu[l].data_ptr()
            lv.v2i = self.dev_vertex2index[l].data_ptr()
            lv.i2v = self.dev_index2vertex[l].data_ptr()
            lv.elem2nnz = self.hex_update_offset_gpu[l].data_ptr()
            for k in ("L_ptr", "L_col", "U_ptr", "U_col", "vert_elem", "invJ", "detJ"):
                setattr(lv, k, d[k].data_ptr())
            d["fold_plan"] = torch.empty(n * 128 + 16, dtype=torch.uint8, device=dev)      # FEMB_FOLD_PLAN_BYTES per vertex
            _lib.check(self.lib.femb_build_fold_plan(C.byref(lv), d["fold_plan"].data_ptr(), self._stream()), "femb_build_fold_plan")
            lv.fold_plan = d["fold_plan"].data_ptr()
            if lattice_fold:
                self._attach_lattice_fold(lv, d)
            self.levels.append(lv)
            self.inverse_pX_peps_gpu.append(d["invJ"])
            self.det_pX_peps_gpu.append(d["detJ"])
            self.vol.append(torch.zeros(nh, device=dev))
            self.m_gpu.append(torch.zeros(n, device=dev))

        # reference attribute names of the level-0 pattern (object.py:499-599)
        p0 = H.pattern[0]
        self.MF_nnz, self.MF_D_nnz, self.MF_L_nnz, self.MF_U_nnz = p0.nnz, p0.n, p0.L_nnz, p0.L_nnz
        self.off_l, self.off_d, self.off_u = p0.off_l, p0.off_d, p0.off_u
        self.MF_GS_U_Ptr, self.MF_GS_L_Ptr = p0.GS_U_Ptr, p0.GS_L_Ptr
        self.UtAUs_nnz = [H.pattern[l].nnz for l in range(1, L)]
        self.UtAUs_off_l = [H.pattern[l].off_l for l in range(1, L)]
        self.UtAUs_off_d = [H.pattern[l].off_d for l in range(1, L)]
        self.UtAUs_off_u = [H.pattern[l].off_u for l in range(1, L)]
        self.UtAUs_block_offset = [torch.from_numpy(H.pattern[l].block_offset) for l in range(1, L)]

        # ---- interpolation operators (object.py:247-443) -----------------------------------------
        for name in ("Ut", "Us", "Ut_noOrder", "Us_noOrder", "Ut_hat", "Us_hat", "Ut_noOrder_hat", "Us_noOrder_hat"):
            setattr(self, name, [TransferMatrix(m, dev) for m in getattr(H, name)])
        self.fix_idx_gpu = _dev(H.fix_idx, dev)
        self.fix_values_gpu = _dev(H.fix_values, dev)

        # ---- rest state of every level (object.py:727-731, 756-759) ------------------------------
        self.x_cpu = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32).copy())
        self.x_gpu_layer = [torch.zeros((H.dims[i], 3), device=dev) for i in range(L)]
        self.x_gpu_layer[0].copy_(self.x_cpu)
        for i in range(1, L):
            self._transfer(self.Ut_noOrder[i - 1], 1.0, self.x_gpu_layer[i - 1], 0.0, self.x_gpu_layer[i])
        for i in range(L):
            self._prepare_rest(i)

        # ---- pins (object.py:761-795) ------------------------------------------------------------
        pinList = list(pinList)
        if pins_exact:
            pass            # the list as given, possibly empty, nothing displaced (a slab of a partitioned mesh: distributed.py)
        elif len(pinList) == 0:
            pinList = hierarchy.default_pins(points, dx).tolist()
        else:
            self.x_cpu[pinList[0]][1] += pin_displacement
            self.x_gpu_layer[0].copy_(self.x_cpu)
        self.pinList = pinList
        self.N_pin = len(pinList)
        pin_cpu = torch.zeros(self.N_verts, dtype=torch.int32)
        pin_cpu[torch.tensor(pinList, dtype=torch.long)] = 1
        self.pin_gpu = pin_cpu.to(dev)
        self.all_pin_pos_gpu = self.x_cpu.to(dev)
        self.pin_list_gpu = torch.tensor(pinList, dtype=torch.int32, device=dev)
        self.pin_pos_gpu = self.all_pin_pos_gpu[self.pin_list_gpu.long()].clone()
        # the k*I pin blocks of MF_value_fixed (object.py:791-793) as a per-row diagonal add, index order
        self.diag_fixed = torch.zeros(self.N_verts, device=dev)
        self.diag_fixed[self.dev_vertex2index[0][self.pin_list_gpu.long()].long()] = self.control_mag

        # ---- matrices and work vectors (object.py:797-855) ---------------------------------------
        self.MF_value_gpu = self.new_values(0)
        self.UtAUs_values = [self.new_values(l) for l in range(1, L)]
        vec = lambda i: torch.zeros((H.dims[i], 3), device=dev)
        self.dev_E = [vec(i) for i in range(L)]; self.dev_R = [vec(i) for i in range(L)]
        self.dev_P = [vec(i) for i in range(L)]; self.dev_AP = [vec(i) for i in range(L)]
        self.dev_B = [vec(i) for i in range(L)]; self.dev_B_fixed = [vec(i) for i in range(L)]
        self.dev_delta_x = [vec(i) for i in range(L)]; self.dev_x_solved = [vec(i) for i in range(L)]
        self.X = vec(0); self.dev_temp_X = vec(0)
        self.x_lineSearch = self.x_gpu_layer[0].clone()
        self.grad_gpu = vec(0)
        self.energy = torch.zeros(1, device=dev); self.energy_lineSearch = torch.zeros(1, device=dev)
        self.squared_sum = torch.zeros(1, device=dev); self.dot_sum = torch.zeros(1, device=dev)
        self.norm_max = torch.zeros(1, device=dev)
        self.not_spd = torch.zeros(1, dtype=torch.int32, device=dev)
        self._scratch = torch.zeros(4096, device=dev)
        self._cg_state = torch.zeros(16 + 1024, device=dev)
        nh_max = max(h.shape[0] for h in H.hexs)
        self._elem_f = torch.empty(nh_max * 24, device=dev)
        self._elem_k = torch.empty((nh_max + 31) // 32 * 32 * _lib.ELEM_K_FLOATS, device=dev)
        self.L = [self._part_view(l, "L") for l in range(L)]
        self.U = [self._part_view(l, "U") for l in range(L)]
        self.D = [self._part_view(l, "D") for l in range(L)]
        self.A = [self._part_view(l, "A") for l in range(L)]
        self._plans = [None] * (L - 1)             # Galerkin plans are built on first use (FAS-only runs never need them)

        self._build_context()
        self.plot_x, self.plot_y, self.plot_InfNorm, self.plot_energy = [], [], [], []
        self.plot_InfNorm_newton, self.plot_energy_newton = [], []
        self.plot_InfNorm_newtonMultigrid, self.plot_energy_newtonMultigrid = [], []
        self.error_log = []
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------------------------------
    # setup helpers
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _load_mesh(mesh_path, dx):
        """object.py:28-45.  Accepts (points, hexagons), 'cube:n' / 'box:nx,ny,nz' (synthetic voxel meshes,
        SURVEY.md 8d C5) or a surface OBJ to voxelise."""
        if isinstance(mesh_path, (tuple, list)):
            return np.asarray(mesh_path[0], np.float32), np.asarray(mesh_path[1], np.int32)
        if isinstance(mesh_path, str) and mesh_path.startswith("cube:"):
            return mesh.cube_mesh(int(mesh_path[5:]), dx)
        if isinstance(mesh_path, str) and mesh_path.startswith("box:"):
            nx, ny, nz = (int(v) for v in mesh_path[4:].split(","))
            return mesh.box_mesh(nx, ny, nz, dx)
        from . import voxelize
        return voxelize.voxelize_obj(mesh_path, dx)

    def _attach_lattice_fold(self, lv, d):
        """FembLevel.lat_*: per vertex the hexahedron in which it is local node a, and the row record lattice offset ->
        sorted position (femb_build_lattice_fold).  With it the assembly keeps the element matrices plane-major and folds
        them with compile-time offsets; levels that are not lattice subsets keep the fold-plan path."""
        dev = self.device
        n = lv.n_vert
        d["lat_inc8"] = torch.empty(max(n, 1) * 8, dtype=torch.int32, device=dev)
        d["lat_row_rec"] = torch.empty(max(n, 1) * _lib.ROW_REC_BYTES, dtype=torch.uint8, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(self.lib.femb_build_lattice_fold(C.byref(lv), d["lat_inc8"].data_ptr(), d["lat_row_rec"].data_ptr(), bad.data_ptr(),
                                                    self._stream()), "femb_build_lattice_fold")
        if int(bad.item()) != 0:                                         # not a lattice subset (e.g. a row with more than 27 blocks)
            del d["lat_inc8"], d["lat_row_rec"]
            return
        lv.lat_inc8 = d["lat_inc8"].data_ptr()
        lv.lat_row_rec = d["lat_row_rec"].data_ptr()

    def _build_context(self):
        """FembContext (include/femb200.h): the arrays of this object as one host struct for the whole-step entry points."""
        L = self.layer
        P = C.c_void_p
        k = self._ctx_keep = {}
        k["levels"] = (_lib.FembLevel * L)(*self.levels)
        for name in ("Ut_hat", "Us_hat", "Ut_noOrder", "Ut_noOrder_hat"):
            k[name] = (_lib.FembTransfer * max(L - 1, 1))(*[m.c for m in getattr(self, name)])
        k["plans"] = (_lib.FembGalerkinPlan * max(L - 1, 1))()
        k["t_values"] = (P * max(L - 1, 1))()
        ptrs = lambda ts: (P * L)(*[t.data_ptr() for t in ts])
        k["values"] = ptrs([self._values(l) for l in range(L)])
        for name, ts in (("B", self.dev_B), ("B_fixed", self.dev_B_fixed), ("delta_x", self.dev_delta_x),
                         ("x_solved", self.dev_x_solved), ("R", self.dev_R)):
            k[name] = ptrs(ts)
        c = _lib.FembContext()
        c.layers = L
        addr = lambda a: C.cast(a, P)
        c.levels = addr(k["levels"])
        c.Ut_hat, c.Us_hat = addr(k["Ut_hat"]), addr(k["Us_hat"])
        c.Ut_noOrder, c.Ut_noOrder_hat = addr(k["Ut_noOrder"]), addr(k["Ut_noOrder_hat"])
        c.plans, c.t_values = addr(k["plans"]), addr(k["t_values"])
        for name in ("values", "B", "B_fixed", "delta_x", "x_solved", "R"):
            setattr(c, name, addr(k[name]))
        c.x, c.grad = self.x_gpu_layer[0].data_ptr(), self.grad_gpu.data_ptr()
        c.diag_fixed, c.mass0 = self.diag_fixed.data_ptr(), self.m_gpu[0].data_ptr()
        c.pin, c.pin_pos = self.pin_gpu.data_ptr(), self.all_pin_pos_gpu.data_ptr()
        c.elem_f, c.elem_k, c.not_spd = self._elem_f.data_ptr(), self._elem_k.data_ptr(), self.not_spd.data_ptr()
        c.mu, c.la, c.g, c.control_mag, c.spd_value = self.mu, self.la, self.g, self.control_mag, self.spd_value
        c.spd_level0 = 1 if self.spd_level0 else 0
        self.ctx = c

    def _ensure_plans(self):
        for i in range(1, self.layer):
            if self._plans[i - 1] is None:
                self._plans[i - 1] = t = self._build_galerkin_plan(i - 1)
                self._ctx_keep["plans"][i - 1] = t["plan"]
                self._ctx_keep["t_values"][i - 1] = t["t_values"].data_ptr()

    def level_without_lattice(self, l):
        """A copy of FembLevel l without the lattice fold: the general fold-plan assembly path (tests, measurements)."""
        lv = _lib.FembLevel()
        C.memmove(C.byref(lv), C.byref(self.levels[l]), C.sizeof(_lib.FembLevel))
        lv.lat_inc8 = None
        lv.lat_row_rec = None
        return lv

    # hooks of the exp3 drivers (exp3/simulation/object.py:258-276, 294-295, 438-441), built on first use
    @property
    def edge_index(self):
        """Per level: the mesh edges as a [2, E] int64 tensor on the device, in the reference's first-seen order."""
        if getattr(self, "_edge_index", None) is None:
            self._edge_index = [torch.from_numpy(hierarchy.edge_index(h)).to(self.device) for h in self.hierarchy.hexs]
        return self._edge_index

    @property
    def torch_A(self):
        """Per coarse level l+1: torch sparse CSR [dims[0], dims[l+1]] of the chained un-normalised prolongation."""
        if getattr(self, "_torch_A", None) is None:
            self._torch_A = [torch.sparse_csr_tensor(torch.from_numpy(m.indptr.astype(np.int32)).to(self.device),
                                                     torch.from_numpy(m.indices.astype(np.int32)).to(self.device),
                                                     torch.from_numpy(m.data.astype(np.float32)).to(self.device),
                                                     size=m.shape, dtype=torch.float32)
                             for m in hierarchy.prolongation_chain(self.hierarchy)]
        return self._torch_A

    def new_values(self, layer):
        """Zeroed block-value array [nnz,3,3] of a level, padded as the bulk-copy kernels require (femb200.h)."""
        nnz = self.hierarchy.pattern[layer].nnz
        buf = torch.zeros(nnz * 9 + 8, device=self.device)
        return buf[:nnz * 9].view(nnz, 3, 3)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _prepare_rest(self, i):
        lv = self.levels[i]
        _lib.check(self.lib.femb_prepare_rest(lv.n_hex, lv.hex, self.x_gpu_layer[i].data_ptr(), lv.detJ, lv.invJ,
                                              self._stream()), "femb_prepare_rest")
        _lib.check(self.lib.femb_lumped_mass(C.byref(lv), self.vol[i].data_ptr(), self.m_gpu[i].data_ptr(), self._stream()),
                   "femb_lumped_mass")

    def _values(self, layer):
        return self.MF_value_gpu if layer == 0 else self.UtAUs_values[layer - 1]

    def _part_view(self, l, part):
        p = self.hierarchy.pattern[l]
        d = self._pat[l]
        n = p.n
        if part == "L":
            return BsrView(n, n, d["L_ptr"], d["L_col"], lambda: self._values(l)[p.off_l:p.off_d])
        if part == "U":
            return BsrView(n, n, d["U_ptr"], d["U_col"], lambda: self._values(l)[p.off_u:p.off_u + p.L_nnz])
        if part == "D":
            ar = torch.arange(n + 1, dtype=torch.int32, device=self.device)
            return BsrView(n, n, ar, ar[:n], lambda: self._values(l)[p.off_d:p.off_u])
        # full matrix in (row,col)-sorted order, built on demand (bsr_set_from_triplets(self.A[i], ...), object.py:1254)
        inv = np.empty(p.nnz, dtype=np.int64)
        inv[p.block_offset] = np.arange(p.nnz)
        rows = p.row[inv]
        offs = np.zeros(n + 1, dtype=np.int64)
        np.add.at(offs, rows.astype(np.int64) + 1, 1)
        offsets = torch.from_numpy(np.cumsum(offs).astype(np.int32))
        columns = torch.from_numpy(p.col[inv].astype(np.int32))
        inv_t = torch.from_numpy(inv)
        return BsrView(n, n, offsets, columns, lambda: self._values(l)[inv_t.to(self.device)])

    def _build_galerkin_plan(self, l):
        """Upload the fixed-pattern plan of A_{l+1} = Ut_hat (A_l Us_hat) (hierarchy.build_galerkin_plan)."""
        g = hierarchy.build_galerkin_plan(self.hierarchy, l)
        dev = self.device
        padb = np.zeros(16, np.uint8)
        t = dict(t_ptr=_dev(g["t_ptr"], dev), t_col=_dev(g["t_col"], dev), c_ptr=_dev(g["c_ptr"], dev),
                 c_col=_dev(g["c_col"], dev), c_slot=_dev(g["c_slot"], dev),
                 t_values=torch.empty(len(g["t_col"]) * 9, device=dev),
                 t_slot=_dev(np.concatenate([g["t_slot"].reshape(-1), padb]), dev),
                 c_slot8=_dev(g["c_slot8"].reshape(-1), dev), ut_q=_dev(g["ut_q"], dev))
        t["plan"] = _lib.FembGalerkinPlan(g["n_fine"], g["n_coarse"], len(g["t_col"]), t["t_ptr"].data_ptr(),
                                          t["t_col"].data_ptr(), t["c_ptr"].data_ptr(), t["c_col"].data_ptr(),
                                          t["c_slot"].data_ptr(), t["t_slot"].data_ptr(), t["c_slot8"].data_ptr(),
                                          t["ut_q"].data_ptr(), int(g["dyadic"]))
        return t

    # ------------------------------------------------------------------------------------------
    # thin operator wrappers (one C-ABI call each)
    # ------------------------------------------------------------------------------------------
    def _transfer(self, T, alpha, x, beta, y):
        _lib.check(self.lib.femb_transfer_mv(C.byref(T.c), alpha, x.data_ptr(), beta, y.data_ptr(), self._stream()),
                   "femb_transfer_mv")

    def _spmv(self, layer, parts, alpha, x, beta, y):
        _lib.check(self.lib.femb_spmv(C.byref(self.levels[layer]), self._values(layer).data_ptr(), parts, alpha, x.data_ptr(),
                                      beta, y.data_ptr(), self._stream()), "femb_spmv")

    def _residual(self, layer, b, x, r):
        _lib.check(self.lib.femb_residual(C.byref(self.levels[layer]), self._values(layer).data_ptr(), b.data_ptr(),
                                          x.data_ptr(), r.data_ptr(), self._stream()), "femb_residual")

    def _axpby(self, a, x, b, y):
        _lib.check(self.lib.femb_axpby(x.numel(), a, x.data_ptr(), b, y.data_ptr(), self._stream()), "femb_axpby")

    def _copy(self, dst, src):
        _lib.check(self.lib.femb_copy(src.numel(), src.data_ptr(), dst.data_ptr(), self._stream()), "femb_copy")

    def _force(self, x, out, layer=0, gravity=True, pins=True, ordered=True):
        """The force launches, e.g. object.py:1234-1236."""
        lv = self.levels[layer]
        _lib.check(self.lib.femb_force(C.byref(lv), x.data_ptr(), self.mu, self.la,
                                       self.m_gpu[layer].data_ptr() if gravity else None, self.g,
                                       self.pin_gpu.data_ptr() if pins else None,
                                       self.all_pin_pos_gpu.data_ptr() if pins else None, self.control_mag,
                                       1 if ordered else 0, self._elem_f.data_ptr(), out.data_ptr(), self._stream()),
                   "femb_force")

    def _hessian(self, layer, x, values, diag_add=None, level=None):
        """compute_elastic_hessian onto (a copy of MF_value_fixed | zeros), e.g. object.py:1233,1249.
        `level`: a FembLevel to use instead of self.levels[layer] (level_without_lattice)."""
        _lib.check(self.lib.femb_hessian(C.byref(level if level is not None else self.levels[layer]), x.data_ptr(), self.mu, self.la,
                                         diag_add.data_ptr() if diag_add is not None else None,
                                         self._elem_k.data_ptr(), values.data_ptr(), self._stream()), "femb_hessian")

    def _force_hessian(self, x, force_out, values, diag_add, gravity=True, pins=True, level=None):
        lv = level if level is not None else self.levels[0]
        _lib.check(self.lib.femb_force_hessian(C.byref(lv), x.data_ptr(), self.mu, self.la,
                                               self.m_gpu[0].data_ptr() if gravity else None, self.g,
                                               self.pin_gpu.data_ptr() if pins else None,
                                               self.all_pin_pos_gpu.data_ptr() if pins else None, self.control_mag,
                                               diag_add.data_ptr() if diag_add is not None else None,
                                               self._elem_f.data_ptr(), self._elem_k.data_ptr(), force_out.data_ptr(),
                                               values.data_ptr(), self._stream()), "femb_force_hessian")

    def _spd(self, values):
        _lib.check(self.lib.femb_spd_blocks(values.data_ptr(), values.shape[0], self.spd_value, self._stream()),
                   "femb_spd_blocks")

    def total_energy(self, x=None):
        """compute_elastic_energy + compute_gravity_energy (object.py:1241-1243); returns a python float (host sync)."""
        x = self.x_gpu_layer[0] if x is None else x
        _lib.check(self.lib.femb_energy(C.byref(self.levels[0]), x.data_ptr(), self.mu, self.la, self.m_gpu[0].data_ptr(),
                                        self.g, self.pin_gpu.data_ptr(), self.all_pin_pos_gpu.data_ptr(), self.control_mag,
                                        self._scratch.data_ptr(), self.energy.data_ptr(), self._stream()), "femb_energy")
        return float(self.energy.item())

    def inf_norm(self, x):
        _lib.check(self.lib.femb_inf_norm(x.numel(), x.data_ptr(), self.norm_max.data_ptr(), self._stream()), "femb_inf_norm")
        return float(self.norm_max.item())

    def update_deltaX(self, x, dx):
        _lib.check(self.lib.femb_update_x(C.byref(self.levels[0]), dx.data_ptr(), x.data_ptr(), self._stream()), "femb_update_x")

    # ------------------------------------------------------------------------------------------
    # smoothers / Krylov (object.py:869-1019)
    # ------------------------------------------------------------------------------------------
    def PerformJacobi(self, layer=0, iterations=2):
        _lib.check(self.lib.femb_jacobi(C.byref(self.levels[layer]), self._values(layer).data_ptr(),
                                        self.dev_B_fixed[layer].data_ptr(), self.dev_delta_x[layer].data_ptr(),
                                        self.dev_B[layer].data_ptr(), iterations, self.not_spd.data_ptr(), self._stream()),
                   "femb_jacobi")

    def PerformGaussSeidel(self, layer=0, iterations=10):
        _lib.check(self.lib.femb_gauss_seidel(C.byref(self.levels[layer]), self._values(layer).data_ptr(),
                                              self.dev_B_fixed[layer].data_ptr(), self.dev_delta_x[layer].data_ptr(),
                                              iterations, self.not_spd.data_ptr(), self._stream()), "femb_gauss_seidel")

    def PerformConjugateGradient(self, layer=0, iterations=10, tol=1e-5, _with_init=0):
        _lib.check(self.lib.femb_cg(C.byref(self.levels[layer]), self._values(layer).data_ptr(),
                                    self.dev_B_fixed[layer].data_ptr(), self.dev_delta_x[layer].data_ptr(),
                                    self.dev_B[layer].data_ptr(), self.dev_P[layer].data_ptr(), self.dev_AP[layer].data_ptr(),
                                    self._cg_state.data_ptr(), iterations, tol, _with_init, self._stream()), "femb_cg")

    def PerformConjugateGradient_WithInitX(self, layer=0, iterations=10, tol=1e-5):
        self.PerformConjugateGradient(layer, iterations, tol, _with_init=1)

    # ------------------------------------------------------------------------------------------
    # grid transfer and V-cycle (object.py:1021-1065, 1216-1226)
    # ------------------------------------------------------------------------------------------
    def downSample(self, layer=0):
        self._copy(self.dev_x_solved[layer], self.dev_delta_x[layer])
        self._residual(layer, self.dev_B_fixed[layer], self.dev_delta_x[layer], self.dev_R[layer])
        self._transfer(self.Ut_hat[layer], 1.0, self.dev_R[layer], 0.0, self.dev_B_fixed[layer + 1])

    def upSample(self, layer=1):
        self._axpby(1.0, self.dev_delta_x[layer], 1.0, self.dev_x_solved[layer])
        self._transfer(self.Us_hat[layer - 1], 1.0, self.dev_x_solved[layer], 1.0, self.dev_x_solved[layer - 1])
        self._spmv(layer - 1, _lib.PART_ALL, -1.0, self.dev_x_solved[layer - 1], 1.0, self.dev_B_fixed[layer - 1])

    def finish(self):
        self._axpby(1.0, self.dev_delta_x[0], 1.0, self.dev_x_solved[0])
        for i in range(1, self.layer):
            self.dev_x_solved[i].zero_()

    def showErrorInfNorm(self, layer, x):
        n = self.inf_norm(x)
        if self.verbose:
            print("Inf_norm : ", n)
        return n

    def showError(self, layer):
        """object.py:1048-1065; forces two host read-backs, so VCycle only calls it when verbose."""
        before = self.inf_norm(self.dev_B_fixed[layer])
        self._residual(layer, self.dev_B_fixed[layer], self.dev_delta_x[layer], self.dev_B[layer])
        after = self.inf_norm(self.dev_B[layer])
        self.error_log.append((layer, before, after))
        if self.verbose:
            print("layer : ", layer)
            print("before solve Inf_norm : ", before)
            print("after solve Inf_norm : ", after)

    def VCycle(self, layer):
        self.PerformGaussSeidel(layer, iterations=3)
        if self.verbose:
            self.showError(layer)
        if layer < self.layer - 1:
            self.downSample(layer)
            self.VCycle(layer + 1)
            self.upSample(layer + 1)
        if layer == self.layer - 1:
            return 0
        self.PerformGaussSeidel(layer, iterations=3)
        if self.verbose:
            self.showError(layer)

    # ------------------------------------------------------------------------------------------
    # non-linear drivers
    # ------------------------------------------------------------------------------------------
    def _sampled(self, step, iterations):
        return step % (iterations / 100) == 0                    # object.py:1238 (float division, sic)

    def _run_step(self, key, body):
        """Run one step body; with use_graphs the body is captured once into a CUDA graph and replayed
        (no host sync inside any body)."""
        if not self.use_graphs or self.verbose:
            body()
            return
        g = self._graphs.get(key)
        if g is None:
            body()                                               # warm-up (sets kernel attributes) -- a real step
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            self._graphs[key] = g
            # the capture itself does not execute: the warm-up call above was this step
            return
        g.replay()

    def galerkin(self):
        """object.py:1252-1264 minus the rebuilds: A_i = Ut_hat (A_{i-1} Us_hat), spd_matrix33f, [L;D;U] layout."""
        self._ensure_plans()
        for i in range(1, self.layer):
            t = self._plans[i - 1]
            _lib.check(self.lib.femb_galerkin(C.byref(self.levels[i - 1]), C.byref(self.levels[i]), self._values(i - 1).data_ptr(),
                                              C.byref(self.Ut_hat[i - 1].c), C.byref(self.Us_hat[i - 1].c), C.byref(t["plan"]),
                                              t["t_values"].data_ptr(), 1, self.spd_value, self.UtAUs_values[i - 1].data_ptr(),
                                              self._stream()), "femb_galerkin")

    def _newton_mg_body(self):
        if self.use_c_steps and not self.verbose:
            self._ensure_plans()
            self.ctx.diag_fixed = self.diag_fixed.data_ptr()                             # (DynamicObject replaces the tensor)
            _lib.check(self.lib.femb_newton_mg_step(C.byref(self.ctx), self._stream()), "femb_newton_mg_step")
            return
        x = self.x_gpu_layer[0]
        self._force_hessian(x, self.grad_gpu, self.MF_value_gpu, self.diag_fixed)       # :1232-1236, 1249
        if self.spd_level0:
            self._spd(self.MF_value_gpu)
        self.galerkin()                                                                  # :1252-1264
        self._copy(self.dev_B_fixed[0], self.grad_gpu)                                   # :1266
        self.VCycle(0)                                                                   # :1290
        self.finish()                                                                    # :1293
        self.update_deltaX(x, self.dev_x_solved[0])                                      # :1296

    def NewtonMultigrid(self, iterations=100):
        """object.py:1230-1296."""
        for step in range(1, iterations + 1):
            sampled = self._sampled(step, iterations)
            e = self.total_energy() if sampled else None
            self._run_step("nmg", self._newton_mg_body)
            if sampled:
                self._record(step, e, self.inf_norm(self.grad_gpu))

    def _record(self, step, energy, norm, tag=""):
        if self.verbose:
            print("Step : ", step); print("Energy : ", energy); print("Inf_norm : ", norm)
        if tag == "":
            self.plot_x.append(step); self.plot_InfNorm.append(norm); self.plot_energy.append(energy)
        elif tag == "newton":
            self.plot_x.append(step); self.plot_InfNorm_newton.append(np.log10(norm)); self.plot_energy_newton.append(energy)
        else:
            self.plot_InfNorm_newtonMultigrid.append(np.log10(norm)); self.plot_energy_newtonMultigrid.append(energy)

    def _newton_body(self):
        x = self.x_gpu_layer[0]
        self._force_hessian(x, self.grad_gpu, self.MF_value_gpu, self.diag_fixed)
        if self.spd_level0:
            self._spd(self.MF_value_gpu)
        self._copy(self.dev_B_fixed[0], self.grad_gpu)
        self.PerformConjugateGradient(layer=0, iterations=5)
        self.update_deltaX(x, self.dev_delta_x[0])

    def Newton(self, iterations=1000):
        """object.py:1127-1160."""
        for step in range(1, iterations + 1):
            sampled = self._sampled(step, iterations)
            e = self.total_energy() if sampled else None
            self._run_step("newton", self._newton_body)
            if sampled:
                self._record(step, e, self.inf_norm(self.grad_gpu))

    # -- FAS family (object.py:1308-1511) ---------------------------------------------------------
    def _fas_fine(self, x, spd, gravity_pins, fixed_diag):
        self.dev_B_fixed[1].zero_()
        g, pn = gravity_pins if isinstance(gravity_pins, tuple) else (gravity_pins, gravity_pins)
        self._force_hessian(x, self.dev_B_fixed[0], self.MF_value_gpu, self.diag_fixed if fixed_diag else None,
                            gravity=g, pins=pn)
        if spd:
            self._spd(self.MF_value_gpu)

    def _fas_coarse_setup(self, x, spd, gravity_pins, coarse_pin_penalty):
        self._transfer(self.Ut_noOrder[0], 1.0, x, 0.0, self.dev_x_solved[1])            # :1489
        g, pn = gravity_pins if isinstance(gravity_pins, tuple) else (gravity_pins, gravity_pins)
        self._force(x, self.dev_R[0], gravity=g, pins=pn)                                # :1491-1495
        self._transfer(self.Ut_hat[0], 1.0, self.dev_R[0], 0.0, self.dev_B_fixed[1])     # :1497
        self._hessian(1, self.dev_x_solved[1], self.UtAUs_values[0])                     # :1499
        if coarse_pin_penalty:                                                           # :1501
            _lib.check(self.lib.femb_coarse_pin_penalty(C.byref(self.levels[1]), C.byref(self.Ut_noOrder_hat[0].c),
                                                        self.pin_gpu.data_ptr(), self.control_mag,
                                                        self.UtAUs_values[0].data_ptr(), self._stream()),
                       "femb_coarse_pin_penalty")
        if spd:
            self._spd(self.UtAUs_values[0])

    def _fas_v123_body(self, x, spd, gravity_pins, fixed_diag, coarse_pin_penalty):
        variant = {(True, False, False, False): 1, (False, True, True, False): 2, (False, True, True, True): 3}.get(
            (spd, gravity_pins, fixed_diag, coarse_pin_penalty)) if not isinstance(gravity_pins, tuple) else None
        if self.use_c_steps and variant is not None:
            self.ctx.diag_fixed = self.diag_fixed.data_ptr()
            _lib.check(self.lib.femb_fas_step(C.byref(self.ctx), variant, x.data_ptr(), self._stream()), "femb_fas_step")
            return
        self._fas_fine(x, spd, gravity_pins, fixed_diag)
        self.PerformGaussSeidel(layer=0, iterations=1)                                   # :1483
        self.update_deltaX(x, self.dev_delta_x[0])                                       # :1487
        self._fas_coarse_setup(x, spd, gravity_pins, coarse_pin_penalty)
        self.PerformJacobi(layer=1, iterations=1)                                        # :1506
        self._transfer(self.Us_hat[0], 1.0, self.dev_delta_x[1], 0.0, self.dev_delta_x[0])   # :1508
        self.update_deltaX(x, self.dev_delta_x[0])                                       # :1509

    def _fas_v0_body(self):
        x = self.dev_x_solved[0]
        self._fas_fine(x, True, False, False)
        self.PerformGaussSeidel(layer=0, iterations=1)
        self.update_deltaX(x, self.dev_delta_x[0])
        self._fas_coarse_setup(x, True, False, False)
        lv1 = C.byref(self.levels[1])
        _lib.check(self.lib.femb_v2i(lv1, self.dev_x_solved[1].data_ptr(), self.dev_delta_x[1].data_ptr(), self._stream()), "femb_v2i")
        self._spmv(1, _lib.PART_ALL, 1.0, self.dev_delta_x[1], 1.0, self.dev_B_fixed[1])            # :1351-1353
        self.PerformConjugateGradient_WithInitX(layer=1, iterations=10)                  # :1356
        _lib.check(self.lib.femb_i2v(lv1, self.dev_delta_x[1].data_ptr(), self.dev_B[1].data_ptr(), self._stream()), "femb_i2v")
        self._copy(self.dev_E[1], self.dev_B[1])                                         # z_axpby, :1361
        self._axpby(-1.0, self.dev_x_solved[1], 1.0, self.dev_E[1])
        self._transfer(self.Us_noOrder_hat[0], 1.0, self.dev_E[1], 0.0, self.dev_E[0])   # :1363
        self._axpby(1.0, self.dev_E[0], 1.0, x)                                          # :1364

    def FASv0(self, iterations=100):
        self.dev_x_solved[0].copy_(self.x_gpu_layer[0])
        for _ in range(iterations):
            self._run_step("fasv0", self._fas_v0_body)
        self.x[0] = self.dev_x_solved[0].cpu()

    # ---- FAS with the tau correction on more than two levels: FASv0 (object.py:1308-1370) generalised ----
    def _ml_scratch(self, l):
        if getattr(self, "_ml_vec", None) is None:
            self._ml_vec = {}
        if l not in self._ml_vec:
            n = self.dims[l]
            self._ml_vec[l] = (torch.zeros((n, 3), device=self.device), torch.zeros((n, 3), device=self.device))
        return self._ml_vec[l]

    def _v2i(self, l, src, dst):
        _lib.check(self.lib.femb_v2i(C.byref(self.levels[l]), src.data_ptr(), dst.data_ptr(), self._stream()), "femb_v2i")

    def _i2v(self, l, src, dst):
        _lib.check(self.lib.femb_i2v(C.byref(self.levels[l]), src.data_ptr(), dst.data_ptr(), self._stream()), "femb_i2v")

    def _ml_smooth(self, l, nu):
        """y_l += GS_nu(A_l, b_l - A_l y_l): the reference's symmetric Gauss-Seidel starts from zero (object.py:886-929), so an
        iterate is smoothed in correction form."""
        r, d = self._ml_scratch(l)
        self._residual(l, self.dev_B_fixed[l], self.dev_delta_x[l], r)
        _lib.check(self.lib.femb_gauss_seidel(C.byref(self.levels[l]), self._values(l).data_ptr(), r.data_ptr(), d.data_ptr(),
                                              nu, self.not_spd.data_ptr(), self._stream()), "femb_gauss_seidel")
        self._axpby(1.0, d, 1.0, self.dev_delta_x[l])

    def _ml_level(self, l, last, nu):
        """Level l of the FAS cycle: A_l = Hessian re-discretised at the restricted state x_l (dev_x_solved[l]), right-hand side
        b_l = R r_{l-1} + A_l x_l in dev_B_fixed[l], iterate y_l in dev_delta_x[l] (index order, starts at x_l)."""
        if l == last:
            self.PerformConjugateGradient_WithInitX(layer=l, iterations=10)              # coarsest: as FASv0, :1356
            return
        self._ml_smooth(l, nu)
        self._i2v(l, self.dev_delta_x[l], self.dev_B[l])                                 # the smoothed state of level l
        self._transfer(self.Ut_noOrder[l], 1.0, self.dev_B[l], 0.0, self.dev_x_solved[l + 1])
        self._residual(l, self.dev_B_fixed[l], self.dev_delta_x[l], self.dev_R[l])
        self._transfer(self.Ut_hat[l], 1.0, self.dev_R[l], 0.0, self.dev_B_fixed[l + 1])
        self._hessian(l + 1, self.dev_x_solved[l + 1], self.UtAUs_values[l])
        self._spd(self.UtAUs_values[l])
        self._v2i(l + 1, self.dev_x_solved[l + 1], self.dev_delta_x[l + 1])
        self._spmv(l + 1, _lib.PART_ALL, 1.0, self.dev_delta_x[l + 1], 1.0, self.dev_B_fixed[l + 1])   # tau: + A_{l+1} x_{l+1}
        self._ml_level(l + 1, last, nu)
        self._i2v(l + 1, self.dev_delta_x[l + 1], self.dev_B[l + 1])
        self._copy(self.dev_E[l + 1], self.dev_B[l + 1])
        self._axpby(-1.0, self.dev_x_solved[l + 1], 1.0, self.dev_E[l + 1])               # coarse approximation of the error
        self._transfer(self.Us_noOrder_hat[l], 1.0, self.dev_E[l + 1], 0.0, self.dev_E[l])
        _, d = self._ml_scratch(l)
        self._v2i(l, self.dev_E[l], d)
        self._axpby(1.0, d, 1.0, self.dev_delta_x[l])
        self._ml_smooth(l, nu)

    def _fas_ml_body(self, last, nu):
        x = self.dev_x_solved[0]
        self._fas_fine(x, True, False, False)
        self.PerformGaussSeidel(layer=0, iterations=1)
        self.update_deltaX(x, self.dev_delta_x[0])
        self._fas_coarse_setup(x, True, False, False)
        self._v2i(1, self.dev_x_solved[1], self.dev_delta_x[1])
        self._spmv(1, _lib.PART_ALL, 1.0, self.dev_delta_x[1], 1.0, self.dev_B_fixed[1])
        self._ml_level(1, last, nu)
        self._i2v(1, self.dev_delta_x[1], self.dev_B[1])
        self._copy(self.dev_E[1], self.dev_B[1])
        self._axpby(-1.0, self.dev_x_solved[1], 1.0, self.dev_E[1])
        self._transfer(self.Us_noOrder_hat[0], 1.0, self.dev_E[1], 0.0, self.dev_E[0])
        self._axpby(1.0, self.dev_E[0], 1.0, x)

    def FAS_multilevel(self, iterations=100, levels=None, smoothing=1):
        """FAS V-cycle with the tau correction over `levels` levels (default: all; meant for `corrected_hierarchy=True`, whose
        levels really coarsen).  Every coarse level l solves A_l(x_l) y = R r_{l-1} + A_l(x_l) x_l for the full approximation y
        (A_l = Hessian re-discretised on the level-l mesh at the restricted state x_l = Ut_noOrder y_{l-1}, eigen-clamped),
        smooths it with `smoothing` symmetric Gauss-Seidel iterations before and after the coarser correction, and hands
        y - x_l back through Us_noOrder_hat.  With levels=2 this IS the reference's FASv0 (object.py:1308-1370), operation for
        operation."""
        L = self.layer if levels is None else int(levels)
        if not 2 <= L <= self.layer:
            raise ValueError(f"levels must be in [2, {self.layer}]")
        self.dev_x_solved[0].copy_(self.x_gpu_layer[0])
        for _ in range(iterations):
            self._run_step(f"fasml{L}_{smoothing}", lambda: self._fas_ml_body(L - 1, int(smoothing)))
        self.x[0] = self.dev_x_solved[0].cpu()

    def FASv1(self, iterations=100):
        self.dev_x_solved[0].copy_(self.x_gpu_layer[0])
        for _ in range(iterations):
            self._run_step("fasv1", lambda: self._fas_v123_body(self.dev_x_solved[0], True, False, False, False))
        self.x[0] = self.dev_x_solved[0].cpu()

    def FASv2(self, iterations=100):
        for _ in range(iterations):
            self._run_step("fasv2", lambda: self._fas_v123_body(self.x_gpu_layer[0], False, True, True, False))
        self.x[0] = self.x_gpu_layer[0].cpu()

    def FASv3(self, iterations=100):
        for _ in range(iterations):
            self._run_step("fasv3", lambda: self._fas_v123_body(self.x_gpu_layer[0], False, True, True, True))
        self.x[0] = self.x_gpu_layer[0].cpu()

    FAS = FASv3

    def step(self):
        """GL-free ``render(pause=False)``: one FASv3 body per frame (object.py:1167-1210)."""
        self._run_step("fasv3", lambda: self._fas_v123_body(self.x_gpu_layer[0], False, True, True, True))

    def render(self, pause=False):
        if not pause:
            self.step()

    def _compare_newton_body(self):
        x = self.x_gpu_layer[0]
        self._fas_fine(x, False, True, True)
        self.PerformGaussSeidel(layer=0, iterations=1)
        self.update_deltaX(x, self.dev_delta_x[0])

    def compare(self, iterations=100):
        """object.py:1514-1598 (the plots are left to the caller: plot_* lists are filled identically)."""
        for step in range(1, iterations + 1):
            sampled = self._sampled(step, iterations)
            e = self.total_energy() if sampled else None
            self._run_step("cmp_newton", self._compare_newton_body)
            if sampled:
                self._record(step, e, self.inf_norm(self.dev_B_fixed[0]), "newton")
        self.x_gpu_layer[0].copy_(self.x_cpu)                                            # :1546
        for step in range(1, iterations + 1):
            sampled = self._sampled(step, iterations)
            e = self.total_energy() if sampled else None
            self._run_step("fasv2", lambda: self._fas_v123_body(self.x_gpu_layer[0], False, True, True, False))
            if sampled:
                # dev_B_fixed[0] still holds the force evaluated at the start of the step (:1565)
                self._record(step, e, self.inf_norm(self.dev_B_fixed[0]), "fas")

    def compare_withoutGravity(self, iterations=100):
        """object.py:1630-1741 (plots left to the caller): the two-level FAS body without gravity (pin springs only), first
        with the coarse pin penalty ("add attachment" -> plot_*_newton), then from the rest state again without it
        (-> plot_*_newtonMultigrid); sampled every iterations/10 steps."""
        x = self.x_gpu_layer[0]
        for tag, penalty in (("newton", True), ("without", False)):
            if tag == "without":
                x.copy_(self.x_cpu)                                                      # :1686
            for step in range(1, iterations + 1):
                sampled = step % (iterations / 10) == 0
                e = self.total_energy() if sampled else None
                self._run_step("cmp_nograv_" + tag, lambda: self._fas_v123_body(x, False, (False, True), True, penalty))
                if sampled:
                    # dev_B_fixed[0] still holds the force evaluated at the start of the step (:1640-1652)
                    self._record(step, e, self.inf_norm(self.dev_B_fixed[0]), tag)

    def _drag_linearise(self):
        x = self.x_gpu_layer[0]
        self._force_hessian(x, self.grad_gpu, self.MF_value_gpu, self.diag_fixed, gravity=False, pins=True)
        self._spd(self.MF_value_gpu)
        self._copy(self.dev_B_fixed[0], self.grad_gpu)

    def _residual_norm(self, layer):
        self._residual(layer, self.dev_B_fixed[layer], self.dev_delta_x[layer], self.dev_B[layer])
        return self.inf_norm(self.dev_B[layer])

    def drag(self, n=10):
        """object.py:1770-1852 without the plots: log10 of the residual infinity norm after i = 0..n-1 iterations of
        Gauss-Seidel, of CG, and after i V-cycles (Newton-MG), linearised at the rest state with the pin springs only.
        Returns {"GS": [...], "CG": [...], "NewtonMG": [...]} (the three curves the reference plots)."""
        out = {"GS": [], "CG": [], "NewtonMG": []}
        for i in range(n):
            self._drag_linearise()
            self.PerformGaussSeidel(layer=0, iterations=i)
            out["GS"].append(float(np.log10(self._residual_norm(0))))
        for i in range(n):
            self.x_gpu_layer[0].copy_(self.x_cpu)
            self._drag_linearise()
            self.PerformConjugateGradient(layer=0, iterations=i)
            out["CG"].append(float(np.log10(self._residual_norm(0))))
        for it in range(n):
            self.x_gpu_layer[0].copy_(self.x_cpu)
            self._drag_linearise()
            self.galerkin()
            norm = self.inf_norm(self.grad_gpu)
            for _ in range(it):
                self.VCycle(0)
                norm = self._residual_norm(0)
            self.finish()
            out["NewtonMG"].append(float(np.log10(norm)))
        self.drag_curves = out
        return out

    def Adam(self, iterations=1000, lr=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8):
        """object.py:1067-1094: first-order baseline on the vertex-order force (sign as in the reference: x -= lr m_hat /
        (sqrt(v_hat) + eps) with the FORCE as `gradient`, cublas.py:185-194)."""
        x = self.x_gpu_layer[0]
        if getattr(self, "m", None) is None:
            self.m, self.v, self.m_hat, self.v_hat = (torch.zeros_like(x) for _ in range(4))       # object.py:816-819
        self.m.zero_(); self.v.zero_()
        for step in range(1, iterations + 1):
            self._force(x, self.grad_gpu, ordered=False)
            if self._sampled(step, iterations):
                self._record(step, self.total_energy(), self.inf_norm(self.grad_gpu))
            _lib.check(self.lib.femb_adam_update(x.numel(), self.grad_gpu.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                                 self.m_hat.data_ptr(), self.v_hat.data_ptr(), x.data_ptr(), beta1, beta2,
                                                 1 / (1 - beta1 ** step), 1 / (1 - beta2 ** step), lr, epsilon, self._stream()),
                       "femb_adam_update")

    def gradientDescent(self, iterations=100, lr=1e-3):
        """object.py:1097-1125 (vertex-order force, energy backtracking x0.2, <= 20 tries; host sync per try)."""
        x = self.x_gpu_layer[0]
        for step in range(1, iterations + 1):
            self._force(x, self.grad_gpu, ordered=False)
            if self._sampled(step, iterations):
                self._energy_value = self.total_energy()
                self._record(step, self._energy_value, self.inf_norm(self.grad_gpu))
            alpha = lr
            for _ in range(20):
                self._copy(self.x_lineSearch, x)
                self._axpby(alpha, self.grad_gpu, 1.0, self.x_lineSearch)
                e_try = self.total_energy(self.x_lineSearch)
                if e_try < getattr(self, "_energy_value", 0.0):
                    self._copy(x, self.x_lineSearch)
                    break
                alpha *= 0.2
