"""ctypes binding of libfemb200.so (C ABI in include/femb200.h) and its in-tree build.

The library is the product: if it is missing or cannot be loaded there is NO fallback --
``load()`` raises and so does every ``Object`` constructor.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "libfemb200.so")
SOURCES = ["femb_api.cu", "femb_element.cu", "femb_comm.cu", "femb_steps.cu", "femb_setup.cu", "femb_setup_hierarchy.cu", "femb_sparse.cu", "femb_dynamic.cu", "femb_cloth.cu", "femb_exp2.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-shared", "-ldl", "-lgomp"]

PART_L, PART_D, PART_U, PART_ALL = 1, 2, 4, 7

c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)


class FembLevel(C.Structure):
    _fields_ = [("n_vert", C.c_int), ("n_hex", C.c_int), ("nnz", C.c_int),
                ("off_l", C.c_int), ("off_d", C.c_int), ("off_u", C.c_int),
                ("color_start", C.c_int * 9),
                ("hex", C.c_void_p), ("v2i", C.c_void_p), ("i2v", C.c_void_p),
                ("L_ptr", C.c_void_p), ("L_col", C.c_void_p), ("U_ptr", C.c_void_p), ("U_col", C.c_void_p),
                ("elem2nnz", C.c_void_p),
                ("vert_elem", C.c_void_p), ("invJ", C.c_void_p), ("detJ", C.c_void_p),
                ("fold_plan", C.c_void_p),
                ("lat_inc8", C.c_void_p), ("lat_row_rec", C.c_void_p)]

ASM_CHUNK_HEXES = 32           # FEMB_ASM_CHUNK_HEXES
ELEM_K_FLOATS = 384            # FEMB_ELEM_K_FLOATS
ROW_REC_BYTES = 48             # FEMB_ROW_REC_BYTES


class FembTransfer(C.Structure):
    _fields_ = [("nrow", C.c_int), ("ncol", C.c_int), ("nnz", C.c_int),
                ("offsets", C.c_void_p), ("columns", C.c_void_p), ("weights", C.c_void_p)]


class FembGalerkinPlan(C.Structure):
    _fields_ = [("n_fine", C.c_int), ("n_coarse", C.c_int), ("t_nnz", C.c_int),
                ("t_ptr", C.c_void_p), ("t_col", C.c_void_p), ("c_ptr", C.c_void_p),
                ("c_col", C.c_void_p), ("c_slot", C.c_void_p), ("t_slot", C.c_void_p), ("c_slot8", C.c_void_p),
                ("ut_q", C.c_void_p), ("dyadic", C.c_int)]


MAX_NEIGHBOURS = 8             # FEMB_MAX_NEIGHBOURS
COMM_ID_BYTES = 128            # FEMB_COMM_ID_BYTES


_IP, _FP, _BP = C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_ubyte)


class FembHierarchyLevel(C.Structure):
    _fields_ = [("n_vert", C.c_int), ("n_hex", C.c_int), ("spacing", C.c_float), ("hex", _IP), ("x", _FP), ("v2i", _IP), ("i2v", _IP),
                ("color_vertex_num", C.c_int * 9), ("L_nnz", C.c_int), ("L_ptr", _IP), ("U_ptr", _IP), ("L_col", _IP), ("U_col", _IP),
                ("elem2nnz", _IP), ("block_offset", _IP), ("hex_parent", _IP), ("fine_coarse", _IP), ("n_entries", C.c_int),
                ("coarse_id", _IP), ("fine_id", _IP), ("entry_k", _BP), ("weight", _FP), ("weight_norm", _FP), ("norm", _FP),
                ("ut_off", _IP), ("ut_col", _IP), ("ut_w", _FP), ("ut_hat_w", _FP), ("us_off", _IP), ("us_col", _IP), ("us_w", _FP),
                ("us_hat_w", _FP)]


class FembHierarchy(C.Structure):
    _fields_ = [("n_levels", C.c_int), ("origin", C.c_float * 3), ("level", C.POINTER(FembHierarchyLevel))]


class FembHalo(C.Structure):
    _fields_ = [("n_nb", C.c_int), ("nb", C.c_int * MAX_NEIGHBOURS), ("n_send", C.c_int * MAX_NEIGHBOURS),
                ("n_recv", C.c_int * MAX_NEIGHBOURS), ("send_rows", C.c_void_p), ("recv_rows", C.c_void_p),
                ("send_buf", C.c_void_p), ("recv_buf", C.c_void_p)]


class FembContext(C.Structure):
    _fields_ = [("layers", C.c_int), ("levels", C.c_void_p), ("Ut_hat", C.c_void_p), ("Us_hat", C.c_void_p),
                ("Ut_noOrder", C.c_void_p), ("Ut_noOrder_hat", C.c_void_p), ("plans", C.c_void_p),
                ("values", C.c_void_p), ("t_values", C.c_void_p), ("B", C.c_void_p), ("B_fixed", C.c_void_p),
                ("delta_x", C.c_void_p), ("x_solved", C.c_void_p), ("R", C.c_void_p),
                ("x", C.c_void_p), ("grad", C.c_void_p), ("diag_fixed", C.c_void_p), ("mass0", C.c_void_p),
                ("pin", C.c_void_p), ("pin_pos", C.c_void_p), ("elem_f", C.c_void_p), ("elem_k", C.c_void_p),
                ("not_spd", C.c_void_p),
                ("mu", C.c_float), ("la", C.c_float), ("g", C.c_float), ("control_mag", C.c_float), ("spd_value", C.c_float),
                ("spd_level0", C.c_int)]


class FembDynamic(C.Structure):
    _fields_ = [("x", C.c_void_p), ("v", C.c_void_p), ("old_x", C.c_void_p), ("inertia_x", C.c_void_p),
                ("more_fixed", C.c_void_p), ("diag_dyn", C.c_void_p), ("cg_p", C.c_void_p), ("cg_ap", C.c_void_p),
                ("cg_state", C.c_void_p), ("dir_host", C.c_void_p),
                ("dt", C.c_float), ("inv_dt", C.c_float), ("damping", C.c_float)]


class FembCloth(C.Structure):
    _fields_ = [("level", C.c_void_p), ("n_edge", C.c_int), ("edge", C.c_void_p), ("rest", C.c_void_p), ("mass", C.c_void_p),
                ("inc_ptr", C.c_void_p), ("inc", C.c_void_p), ("off_p0p1", C.c_void_p), ("off_p1p0", C.c_void_p),
                ("d_slot", C.c_void_p), ("zeros_i", C.c_void_p),
                ("x", C.c_void_p), ("v", C.c_void_p), ("old_x", C.c_void_p), ("inertia_x", C.c_void_p),
                ("values", C.c_void_p), ("grad", C.c_void_p), ("B", C.c_void_p), ("B_fixed", C.c_void_p),
                ("delta_x", C.c_void_p), ("cg_p", C.c_void_p), ("cg_ap", C.c_void_p), ("cg_state", C.c_void_p),
                ("edge_h", C.c_void_p), ("edge_g", C.c_void_p),
                ("k", C.c_float), ("g", C.c_float), ("dt", C.c_float), ("inv_dt", C.c_float), ("damping", C.c_float)]


P = C.c_void_p
HL = C.POINTER(FembHalo)
CX = C.POINTER(FembContext)
LV = C.POINTER(FembLevel)
TR = C.POINTER(FembTransfer)
GP = C.POINTER(FembGalerkinPlan)
I, F = C.c_int, C.c_float

# name -> argtypes; every function returns int unless listed in _RESTYPE (mirrors include/femb200.h)
PROTOTYPES = {
    "femb_abi_version": [],
    "femb_last_error": [],
    "femb_launch_count": [],
    "femb_check_device": [I],
    "femb_setup_incidence_host": [I, I, P, P],
    "femb_setup_coo_to_csr_host": [I, I, P, P, P, P, P, P],
    "femb_setup_coo_to_csr_pair_host": [I, I, P, P, P, P, P, P, P, P],
    "femb_setup_pattern_count_host": [I, P, P, P, P, P, P],
    "femb_setup_pattern_fill_host": [I, I, P, P, P, P, P, P, P, P, P, P],
    "femb_setup_galerkin_count_host": [I, P, P, P, P, P, P, P],
    "femb_setup_galerkin_fill_host": [I, I] + [P] * 22,
    "femb_setup_stencil_host": [P, P, P],
    "femb_setup_color_order_host": [I, P, P, F, P, P, P],
    "femb_setup_coarsen_count_host": [I, I, P, P, P, F, P, P, P, P],
    "femb_setup_coarsen_fill_host": [I, I, P, P, P, F, F, P, P, I, I, P, P, P, P, P, P],
    "femb_setup_interp_count_host": [I, I, P, P, P],
    "femb_setup_interp_fill_host": [I, I, I, P, P, P, I, P, P, P, P, P, P],
    "femb_setup_mask_mesh_count_host": [I, I, I, P, P, P],
    "femb_setup_mask_mesh_fill_host": [I, I, I, P, F, P, P, I, I, P, P],
    "femb_setup_obj_count_host": [C.c_char_p, P, P],
    "femb_setup_obj_read_host": [C.c_char_p, C.c_longlong, C.c_longlong, P, P],
    "femb_setup_points_inside_host": [I, P, I, P, I, P, I, P, P],
    "femb_setup_hierarchy_host": [I, I, P, P, F, I, I, P, P],
    "femb_setup_hierarchy_free": [P],
    "femb_prepare_rest": [I, P, P, P, P, P],
    "femb_build_fold_plan": [LV, P, P],
    "femb_build_lattice_fold": [LV, P, P, P, P],
    "femb_lumped_mass": [LV, P, P, P],
    "femb_force": [LV, P, F, F, P, F, P, P, F, I, P, P, P],
    "femb_hessian": [LV, P, F, F, P, P, P, P],
    "femb_set_fold_config": [I, I],
    "femb_set_assembly_async": [I],
    "femb_force_hessian": [LV, P, F, F, P, F, P, P, F, P, P, P, P, P, P],
    "femb_energy": [LV, P, F, F, P, F, P, P, F, P, P, P],
    "femb_coarse_pin_penalty": [LV, TR, P, F, P, P],
    "femb_spmv": [LV, P, I, F, P, F, P, P],
    "femb_residual": [LV, P, P, P, P, P],
    "femb_gauss_seidel": [LV, P, P, P, I, P, P],
    "femb_set_pdl": [I],
    "femb_set_gs_threads_per_row": [I],
    "femb_set_gs_rows_per_cta": [I],
    "femb_gs_colors": [LV, P, P, P, I, I, I, P, P],
    "femb_jacobi": [LV, P, P, P, P, I, P, P],
    "femb_cg": [LV, P, P, P, P, P, P, P, I, F, I, P],
    "femb_spd_blocks": [P, I, F, P],
    "femb_transfer_mv": [TR, F, P, F, P, P],
    "femb_galerkin": [LV, LV, P, TR, TR, GP, P, I, F, P, P],
    "femb_dyn_begin": [I, P, P, P, P, P, P, P, C.POINTER(C.c_float), F, F, P],
    "femb_dyn_diag": [LV, P, P, P, F, F, P, P],
    "femb_dyn_force": [LV, P, P, P, P, P, F, F, P, P],
    "femb_dyn_velocity": [I, P, P, F, P, P],
    "femb_control": [I, I, P, P, P],
    "femb_cloth_rest_length": [I, P, P, P, P],
    "femb_cloth_assemble": [I, I, P, P, F, P, P, F, P, F, P, P, P, P, P, P, P, P, P, P],
    "femb_cloth_update": [I, P, P, I, I, P],
    "femb_vcycle": [CX, I, P],
    "femb_galerkin_chain": [CX, P],
    "femb_newton_mg_step": [CX, P],
    "femb_fas_step": [CX, I, P, P],
    "femb_dynamic_step": [CX, C.POINTER(FembDynamic), P],
    "femb_cloth_step": [C.POINTER(FembCloth), P],
    "femb_comm_unique_id": [P],
    "femb_comm_init": [C.POINTER(C.c_void_p), P, I, I],
    "femb_comm_destroy": [P],
    "femb_halo_exchange": [P, HL, P, I, P],
    "femb_comm_p2p_enable": [P, P, I, C.c_size_t, P],
    "femb_comm_p2p_error": [P, P],
    "femb_comm_allreduce": [P, P, I, I, P],
    "femb_axpby": [I, F, P, F, P, P],
    "femb_copy": [I, P, P, P],
    "femb_adam_update": [I, P, P, P, P, P, P, F, F, F, F, F, F, P],
    "femb_fill": [I, F, P, P],
    "femb_update_x": [LV, P, P, P],
    "femb_v2i": [LV, P, P, P],
    "femb_i2v": [LV, P, P, P],
    "femb_absmax_onehot": [I, P, P, P, P],
    "femb_solve3x3_vjp": [I, P, P, P, P, P],
    "femb_coarse_diag_vjp": [LV, P, P, F, F, P, P, P, P, P],
    "femb_exp2_weight_grad": [TR, P, P, P, P, P, F, P, P, P],
    "femb_exp2_weight_update": [I, P, P, F, P],
    "femb_inf_norm": [I, P, P, P],
    "femb_dot": [I, P, P, P, P, P],
}
_RESTYPE = {"femb_last_error": C.c_char_p, "femb_launch_count": C.c_longlong}


class FembError(RuntimeError):
    pass


def build(verbose: bool = False, force: bool = True) -> str:
    """Compile libfemb200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import fcntl
    src = [os.path.join(_PKG, "csrc", s) for s in SOURCES]
    # ranks of one job may all find the library stale at once: one builds (file lock), into a temporary file that replaces the
    # library atomically, the others wait and then find it fresh
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _needs_build():
                return LIB_PATH
            tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
            cmd = ["nvcc"] + NVCC_FLAGS + ["-I", os.path.join(_ROOT, "include"), "-o", tmp] + src
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise FembError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def _needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_PKG, "csrc", f) for f in os.listdir(os.path.join(_PKG, "csrc"))]
    deps.append(os.path.join(_ROOT, "include", "femb200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


_lib = None


def load(rebuild: bool = False):
    """Load (building first if sources are newer) and type the C ABI.  Raises if that is impossible."""
    global _lib
    if _lib is not None and not rebuild:
        return _lib
    if rebuild or _needs_build():
        build(force=rebuild)
    lib = C.CDLL(LIB_PATH)
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)            # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    if lib.femb_abi_version() != 5:
        raise FembError("libfemb200 ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = _lib.femb_last_error().decode() if _lib is not None else ""
        raise FembError(f"{what} failed with code {rc}: {msg}")
