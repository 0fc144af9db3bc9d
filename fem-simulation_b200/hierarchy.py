"""Host-side setup of the multigrid hierarchy, vectorised (numpy), numbering identical to the reference.

Mirrors ``Object.__init__`` of /root/reference/exp1/quasi_simulation/object.py:116-865, which builds
everything with per-element Python loops and dictionaries (minutes at 1e4 hexes).  Here every
dictionary "first seen" numbering is reproduced with stable sorts / ``np.unique`` so that the
integer outputs (colour/index order, coarse meshes, ``fine_coarse``, interpolation triplets, the
[L;D;U] slot maps and the element->nnz maps) are bit-identical while scaling to 1e7 hexes.
SURVEY.md Appendix B is the specification; each function cites the reference lines it replaces.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import ctypes as C

import numpy as np
import torch

F32 = np.float32


@dataclass
class SparseW:
    """CSR matrix with one scalar weight per entry; stands for a BSR matrix with blocks w*I3
    (the reference stores the 36-byte blocks: object.py:422-436)."""
    nrow: int
    ncol: int
    offsets: np.ndarray
    columns: np.ndarray
    weights: np.ndarray

    @property
    def nnz(self):
        return int(self.columns.shape[0])

    def rows(self):
        return np.repeat(np.arange(self.nrow, dtype=np.int32), np.diff(self.offsets))

    def transposed(self):
        return coo_to_csr(self.ncol, self.nrow, self.columns, self.rows(), self.weights)

    def block_values(self):
        return self.weights[:, None, None] * np.eye(3, dtype=F32)


def coo_to_csr_numpy(nrow, ncol, rows, cols, w) -> SparseW:
    """sort by (row, col); entries are unique here (sparse.py:173-278 would sum duplicates).  numpy restatement (tests)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    order = np.argsort(rows * np.int64(max(ncol, 1)) + cols, kind="stable")
    r = rows[order]
    offsets = np.zeros(nrow + 1, dtype=np.int64)
    np.add.at(offsets, r + 1, 1)
    return SparseW(int(nrow), int(ncol), np.cumsum(offsets).astype(np.int32), cols[order].astype(np.int32),
                   np.asarray(w, dtype=F32)[order])


def coo_to_csr(nrow, ncol, rows, cols, w) -> SparseW:
    """sort by (row, col) through the host routine of the C ABI (csrc/femb_setup.cu: counting sort by row + per-row sort)."""
    from . import _lib
    lib = _lib.load()
    r = np.ascontiguousarray(rows, dtype=np.int32)
    c = np.ascontiguousarray(cols, dtype=np.int32)
    ww = np.ascontiguousarray(w, dtype=F32)
    nnz = int(r.shape[0])
    offsets = np.empty(int(nrow) + 1, np.int32)
    columns, weights = np.empty(nnz, np.int32), np.empty(nnz, F32)
    hp = lambda a: a.ctypes.data_as(_lib.C.c_void_p)
    _lib.check(lib.femb_setup_coo_to_csr_host(int(nrow), nnz, hp(r), hp(c), hp(ww), hp(offsets), hp(columns), hp(weights)),
               "femb_setup_coo_to_csr_host")
    return SparseW(int(nrow), int(ncol), offsets, columns, weights)


def coo_to_csr_pair(nrow, ncol, rows, cols, w_a, w_b):
    """two matrices on one pattern (same rows / columns, two weight arrays) from one sort; they share the offset and column
    arrays (nobody writes to them)."""
    from . import _lib
    lib = _lib.load()
    r = np.ascontiguousarray(rows, dtype=np.int32)
    c = np.ascontiguousarray(cols, dtype=np.int32)
    wa, wb = np.ascontiguousarray(w_a, dtype=F32), np.ascontiguousarray(w_b, dtype=F32)
    nnz = int(r.shape[0])
    offsets = np.empty(int(nrow) + 1, np.int32)
    columns, wa_out, wb_out = np.empty(nnz, np.int32), np.empty(nnz, F32), np.empty(nnz, F32)
    hp = lambda a: a.ctypes.data_as(_lib.C.c_void_p)
    _lib.check(lib.femb_setup_coo_to_csr_pair_host(int(nrow), nnz, hp(r), hp(c), hp(wa), hp(wb), hp(offsets), hp(columns), hp(wa_out),
                                                   hp(wb_out)), "femb_setup_coo_to_csr_pair_host")
    return SparseW(int(nrow), int(ncol), offsets, columns, wa_out), SparseW(int(nrow), int(ncol), offsets, columns, wb_out)


def transposed_pair(a: SparseW, b: SparseW):
    """a^T and b^T for two matrices that share their pattern"""
    return coo_to_csr_pair(a.ncol, a.nrow, a.columns, a.rows(), a.weights, b.weights)


class LevelPattern:
    """Fixed block pattern of one level in the reference's [L;D;U] value layout (object.py:485-697).  The per-entry row arrays
    of the reference (`*_L_row`, `*_U_row`, `*_GS_*_row/col`: what its bsr_set_from_triplets calls consume) are derived from
    the row pointers on first use -- the product's kernels only read the pointers and the columns."""

    def __init__(self, n, nnz, L_nnz, off_l, off_d, off_u, L_col, U_col, L_ptr, U_ptr, GS_L_Ptr, GS_U_Ptr, block_offset,
                 hex_update_offset, cvn=None, L_row=None, U_row=None, GS_L_row=None, GS_L_col=None, GS_U_row=None, GS_U_col=None):
        self.n, self.nnz, self.L_nnz = n, nnz, L_nnz
        self.off_l, self.off_d, self.off_u = off_l, off_d, off_u
        self.L_col, self.U_col = L_col, U_col
        self.L_ptr, self.U_ptr = L_ptr, U_ptr          # [n+1] row pointers into the L / U sections (rows are contiguous there)
        self.GS_L_Ptr, self.GS_U_Ptr = GS_L_Ptr, GS_U_Ptr
        self.block_offset = block_offset               # UtAUs_block_offset: [L;D;U] slot -> rank in (row,col)-sorted order
        self.hex_update_offset = hex_update_offset
        self._cv = None if cvn is None else np.asarray(cvn, dtype=np.int64)
        self._lazy = dict(L_row=L_row, U_row=U_row, GS_L_row=GS_L_row, GS_L_col=GS_L_col, GS_U_row=GS_U_row, GS_U_col=GS_U_col)

    def _get(self, name):
        v = self._lazy[name]
        if v is None:
            rows = np.arange(self.n, dtype=np.int32)
            if name == "L_row":
                v = np.repeat(rows, np.diff(self.L_ptr))
            elif name == "U_row":
                v = np.repeat(rows, np.diff(self.U_ptr))
            elif name == "GS_L_col":
                v = self.L_col.copy()
            else:
                cv = self._cv
                row_colour = (np.searchsorted(cv, rows, side="right") - 1).astype(np.int64)
                if name == "GS_L_row":
                    v = (self.L_row - cv[np.repeat(row_colour, np.diff(self.L_ptr))]).astype(np.int32)
                else:
                    colU = np.repeat(row_colour, np.diff(self.U_ptr))
                    v = ((self.U_row - cv[colU]) if name == "GS_U_row" else (self.U_col - cv[colU + 1])).astype(np.int32)
            self._lazy[name] = v
        return v

    L_row = property(lambda self: self._get("L_row"))
    U_row = property(lambda self: self._get("U_row"))
    GS_L_row = property(lambda self: self._get("GS_L_row"))
    GS_L_col = property(lambda self: self._get("GS_L_col"))
    GS_U_row = property(lambda self: self._get("GS_U_row"))
    GS_U_col = property(lambda self: self._get("GS_U_col"))

    @property
    def D_row(self):
        return np.arange(self.n, dtype=np.int32)

    @property
    def row(self):
        return np.concatenate([self.L_row, self.D_row, self.U_row])

    @property
    def col(self):
        return np.concatenate([self.L_col, self.D_row, self.U_col])


@dataclass
class Hierarchy:
    dx: float
    layer: int
    box_min: np.ndarray
    box_max: np.ndarray
    dims: list
    hexs: list
    x: list                     # lattice positions per level (object.py:230; colouring/plots only)
    fine_coarse: list
    color_vertex_num: list
    vertex2index: list
    index2vertex: list
    Ut: list = field(default_factory=list)
    Ut_noOrder: list = field(default_factory=list)
    Ut_hat: list = field(default_factory=list)
    Ut_noOrder_hat: list = field(default_factory=list)
    Us: list = field(default_factory=list)
    Us_noOrder: list = field(default_factory=list)
    Us_hat: list = field(default_factory=list)
    Us_noOrder_hat: list = field(default_factory=list)
    U_triplets: list = field(default_factory=list)
    fix_idx: np.ndarray = None
    fix_values: np.ndarray = None
    pattern: list = field(default_factory=list)
    fine_hex_parent: list = field(default_factory=list)   # [layer-1] int32 [N_h(l)]: coarse hexahedron that contains fine hexahedron e
    spacing: list = field(default_factory=list)           # lattice spacing of every level (dx, 2 dx, 4 dx, then 4 dx again: object.py:175-179)
    shapeFuncGrad: np.ndarray = None


def shape_func_grad() -> np.ndarray:
    """shapeFuncGrad[a][q][c] (cpu_function.py:4-9, tables object.py:701-722), float32 products in
    the reference's left-to-right order, Gauss abscissa literal 0.57735."""
    h = np.array([[-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1],
                  [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], dtype=F32)
    quad = h * F32(0.57735)
    one, e8 = F32(1), F32(0.125)
    out = np.zeros((8, 8, 3), dtype=F32)
    for a in range(8):
        for q in range(8):
            out[a, q, 0] = h[a, 0] * (one + h[a, 1] * quad[q, 1]) * (one + h[a, 2] * quad[q, 2]) * e8
            out[a, q, 1] = h[a, 1] * (one + h[a, 0] * quad[q, 0]) * (one + h[a, 2] * quad[q, 2]) * e8
            out[a, q, 2] = h[a, 2] * (one + h[a, 1] * quad[q, 1]) * (one + h[a, 0] * quad[q, 0]) * e8
    return out


def _ijk(points: np.ndarray, origin: np.ndarray, spacing: float, shift: float) -> np.ndarray:
    """cpu_function.py:12-20: ((p - origin)/spacing + shift).int() in float32."""
    return ((points.astype(F32) - origin.astype(F32)) / F32(spacing) + F32(shift)).astype(np.int32)


def _first_seen_ids(keys: np.ndarray):
    """Dictionary numbering: id of each entry = rank of the first appearance of its key."""
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[order] = np.arange(len(uniq))
    return rank[inv.reshape(-1)], uniq[order]


def _color_order(color: np.ndarray):
    """object.py:151-158."""
    i2v = np.argsort(color, kind="stable").astype(np.int32)
    v2i = np.empty_like(i2v)
    v2i[i2v] = np.arange(len(i2v), dtype=np.int32)
    cvn = [0] + np.cumsum(np.bincount(color, minlength=8)).astype(np.int64).tolist()
    return v2i, i2v, [int(c) for c in cvn]


# slot (i,j,k) of a coarse hex -> ordered coarse corners and weight (object.py:283-412)
def _stencil_table():
    cnt = np.zeros(27, dtype=np.int64)
    corners = np.zeros((27, 8), dtype=np.int64)
    weight = np.zeros(27, dtype=F32)
    for i in range(3):
        for j in range(3):
            for k in range(3):
                s = 9 * i + 3 * j + k
                par = i % 2 + j % 2 + k % 2
                if par == 0:
                    c = [4 * (i // 2) + 2 * (j // 2) + (k // 2)]; w = 1.0
                elif par == 3:
                    c = list(range(8)); w = 0.125
                elif par == 1:
                    if i == 1:
                        c = [2 * (j // 2) + (k // 2), 4 + 2 * (j // 2) + (k // 2)]
                    elif j == 1:
                        c = [4 * (i // 2) + (k // 2), 4 * (i // 2) + 2 + (k // 2)]
                    else:
                        c = [4 * (i // 2) + 2 * (j // 2), 4 * (i // 2) + 2 * (j // 2) + 1]
                    w = 0.5
                else:
                    if i % 2 == 0:
                        c = [4 * (i // 2) + 2 * ii + jj for ii in range(2) for jj in range(2)]
                    elif j % 2 == 0:
                        c = [4 * ii + 2 * (j // 2) + jj for ii in range(2) for jj in range(2)]
                    else:
                        c = [4 * ii + 2 * jj + (k // 2) for ii in range(2) for jj in range(2)]
                    w = 0.25
                cnt[s] = len(c); corners[s, :len(c)] = c; weight[s] = w
    return cnt, corners, weight


_ST_CNT, _ST_CORNERS, _ST_W = _stencil_table()


def level_count(box_min, box_max, dx) -> int:
    """object.py:140-143, float32 torch arithmetic."""
    delta = torch.tensor(box_max, dtype=torch.float32) - torch.tensor(box_min, dtype=torch.float32)
    mb = torch.min(delta)
    mb = torch.log(mb / dx) / torch.log(torch.tensor(2.0))
    return int(mb.int() - 1)


def _hp(a):
    return a.ctypes.data_as(C.c_void_p)


def vertex_incidence(hexagons: np.ndarray, n: int) -> np.ndarray:
    """[n, 8] int32: e*8+a of the hexahedra incident to every vertex in ascending order, -1 padded (femb_setup_incidence_host)."""
    from . import _lib
    lib = _lib.load()
    hx = np.ascontiguousarray(hexagons, dtype=np.int32)
    ve = np.empty((n, 8), dtype=np.int32)
    _lib.check(lib.femb_setup_incidence_host(n, hx.shape[0], _hp(hx), _hp(ve)), "femb_setup_incidence_host")
    return ve


def build_pattern(hexagons: np.ndarray, v2i: np.ndarray, cvn: list, n: int, vert_elem: np.ndarray | None = None) -> LevelPattern:
    """object.py:485-697 for one level, through the host routines of the C ABI (csrc/femb_setup.cu: per-row sorted unions
    instead of one global sort); `build_pattern_numpy` is the numpy restatement the tests compare it with."""
    from . import _lib
    lib = _lib.load()
    hx = np.ascontiguousarray(hexagons, dtype=np.int32)
    v2i = np.ascontiguousarray(v2i, dtype=np.int32)
    i2v = np.empty_like(v2i)
    i2v[v2i] = np.arange(n, dtype=np.int32)
    ve = vertex_incidence(hx, n) if vert_elem is None else vert_elem
    L_ptr, U_ptr = np.empty(n + 1, np.int32), np.empty(n + 1, np.int32)
    _lib.check(lib.femb_setup_pattern_count_host(n, _hp(hx), _hp(v2i), _hp(i2v), _hp(ve), _hp(L_ptr), _hp(U_ptr)),
               "femb_setup_pattern_count_host")
    L_nnz = int(L_ptr[n])
    assert int(U_ptr[n]) == L_nnz
    nnz = 2 * L_nnz + n
    L_col, U_col = np.empty(L_nnz, np.int32), np.empty(L_nnz, np.int32)
    e2n = np.empty(hx.shape[0] * 64, np.int32)
    block_offset = np.empty(nnz, np.int32)
    _lib.check(lib.femb_setup_pattern_fill_host(n, hx.shape[0], _hp(hx), _hp(v2i), _hp(i2v), _hp(ve), _hp(L_ptr), _hp(U_ptr), _hp(L_col),
                                                _hp(U_col), _hp(e2n), _hp(block_offset)), "femb_setup_pattern_fill_host")
    cv = np.asarray(cvn, dtype=np.int64)
    GS_U_Ptr = [int(U_ptr[min(cv[p], n)]) for p in range(8)] + [0]
    GS_L_Ptr = [int(L_ptr[min(cv[p + 1], n)]) for p in range(8)] + [0]
    return LevelPattern(n=n, nnz=nnz, L_nnz=L_nnz, off_l=0, off_d=L_nnz, off_u=L_nnz + n, L_col=L_col, U_col=U_col, L_ptr=L_ptr,
                        U_ptr=U_ptr, GS_L_Ptr=GS_L_Ptr, GS_U_Ptr=GS_U_Ptr, block_offset=block_offset, hex_update_offset=e2n, cvn=cvn)


def build_pattern_numpy(hexagons: np.ndarray, v2i: np.ndarray, cvn: list, n: int) -> LevelPattern:
    """object.py:485-697 for one level (numpy restatement: one global sort of the 64 n_hex (row, col) keys)."""
    ih = v2i[hexagons].astype(np.int64)                                   # [Nh,8] index-order ids
    keys = (ih[:, :, None] * np.int64(n) + ih[:, None, :]).reshape(-1)    # (r,c) for e,j,k
    uniq, inv = np.unique(keys, return_inverse=True)
    r = (uniq // n).astype(np.int64)
    c = (uniq % n).astype(np.int64)
    nnz = len(uniq)
    L_nnz = (nnz - n) >> 1
    off_l, off_d, off_u = 0, L_nnz, L_nnz + n
    isL, isU = r > c, r < c
    slot = np.empty(nnz, dtype=np.int64)
    slot[r == c] = off_d + r[r == c]
    slot[isU] = off_u + np.arange(int(isU.sum()))
    slot[isL] = off_l + np.arange(int(isL.sum()))
    assert int(isU.sum()) == L_nnz and int(isL.sum()) == L_nnz
    block_offset = np.empty(nnz, dtype=np.int32)
    block_offset[slot] = np.arange(nnz, dtype=np.int32)
    L_row, L_col = r[isL].astype(np.int32), c[isL].astype(np.int32)
    U_row, U_col = r[isU].astype(np.int32), c[isU].astype(np.int32)
    cv = np.asarray(cvn, dtype=np.int64)
    colU = np.searchsorted(cv, U_row, side="right") - 1
    colL = np.searchsorted(cv, L_row, side="right") - 1
    GS_U_row = (U_row - cv[colU]).astype(np.int32)
    GS_U_col = (U_col - cv[colU + 1]).astype(np.int32)
    GS_L_row = (L_row - cv[colL]).astype(np.int32)
    GS_L_col = L_col.copy()
    GS_U_Ptr = [int(np.searchsorted(U_row, cv[p], side="left")) for p in range(8)] + [0]
    GS_L_Ptr = [int(np.searchsorted(L_row, cv[p + 1], side="left")) for p in range(8)] + [0]
    L_ptr = np.searchsorted(L_row, np.arange(n + 1), side="left").astype(np.int32)
    U_ptr = np.searchsorted(U_row, np.arange(n + 1), side="left").astype(np.int32)
    return LevelPattern(n=n, nnz=nnz, L_nnz=L_nnz, off_l=off_l, off_d=off_d, off_u=off_u,
                        L_row=L_row, L_col=L_col, U_row=U_row, U_col=U_col, L_ptr=L_ptr, U_ptr=U_ptr,
                        GS_L_row=GS_L_row, GS_L_col=GS_L_col, GS_U_row=GS_U_row, GS_U_col=GS_U_col,
                        GS_L_Ptr=GS_L_Ptr, GS_U_Ptr=GS_U_Ptr, block_offset=block_offset,
                        hex_update_offset=slot[inv.reshape(-1)].astype(np.int32))


class MeshTooSmallError(ValueError, IndexError):
    """The reference's level formula gives no level at all (object.py:143: layer = int(log2(min extent / dx)) - 1 < 1);
    its constructor then dies with ``IndexError: list index out of range`` -- same class of error here, with a message."""


def _color_order_native(x: np.ndarray, origin: np.ndarray, spacing: float):
    """object.py:147-158 through femb_setup_color_order_host."""
    from . import _lib
    lib = _lib.load()
    n = x.shape[0]
    o = np.ascontiguousarray(origin, dtype=F32)
    v2i, i2v, cvn = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(9, np.int32)
    _lib.check(lib.femb_setup_color_order_host(n, _hp(x), _hp(o), C.c_float(spacing), _hp(v2i), _hp(i2v), _hp(cvn)),
               "femb_setup_color_order_host")
    return v2i, i2v, [int(c) for c in cvn]


def _coarsen_numpy(xl, hx, box_min, fine_size, coarse_size):
    """object.py:171-241 for one level (numpy restatement: dictionary numbering through np.unique + first occurrence)."""
    BIG = np.int64(1) << 20
    key3 = _ijk(xl[hx[:, 0]], box_min, coarse_size, 0.1).astype(np.int64)        # [Nh,3]
    hex_id, ukeys = _first_seen_ids((key3[:, 0] * BIG + key3[:, 1]) * BIG + key3[:, 2])
    n_ch = len(ukeys)
    ck = np.stack([ukeys // (BIG * BIG), (ukeys // BIG) % BIG, ukeys % BIG], axis=1)  # coarse hex keys, id order
    jkl = np.array([[(a >> 2) & 1, (a >> 1) & 1, a & 1] for a in range(8)], dtype=np.int64)
    vk3 = ck[:, None, :] + jkl[None, :, :]                                        # [n_ch,8,3]
    vkeys = ((vk3[..., 0] * BIG + vk3[..., 1]) * BIG + vk3[..., 2]).reshape(-1)
    vid, uvk = _first_seen_ids(vkeys)
    coarse_hex = vid.reshape(n_ch, 8).astype(np.int32)
    vk = np.stack([uvk // (BIG * BIG), (uvk // BIG) % BIG, uvk % BIG], axis=1)
    # slot of node b of hexahedron e inside its coarse cell: 9 di + 3 dj + dk with (di, dj, dk) = ijk(vertex) - 2 key3(e);
    # the lattice index is a function of the vertex, so it is evaluated once per vertex, folded to 9 i + 3 j + k and gathered
    ijk_v = _ijk(xl, box_min, fine_size, 0.1).astype(np.int64)
    code_v = 9 * ijk_v[:, 0] + 3 * ijk_v[:, 1] + ijk_v[:, 2]
    slot = code_v[hx] - 2 * (9 * key3[:, 0] + 3 * key3[:, 1] + key3[:, 2])[:, None]
    fc = np.full((n_ch, 27), -1, dtype=np.int32)
    fc[np.repeat(hex_id, 8), slot.reshape(-1)] = hx.reshape(-1)
    coarse_vert = (vk.astype(F32) * F32(coarse_size) + box_min).astype(F32)
    ccolor = 4 * (vk[:, 2] % 2) + 2 * (vk[:, 1] % 2) + (vk[:, 0] % 2)
    cv2i, ci2v, ccvn = _color_order(ccolor)
    return hex_id.astype(np.int32), coarse_hex, coarse_vert, fc, cv2i, ci2v, ccvn


def _coarsen_native(xl, hx, box_min, fine_size, coarse_size):
    """object.py:171-241 for one level through femb_setup_coarsen_count_host / _fill_host (csrc/femb_setup_hierarchy.cu)."""
    from . import _lib
    lib = _lib.load()
    nv, nh = xl.shape[0], hx.shape[0]
    o = np.ascontiguousarray(box_min, dtype=F32)
    hex_id, key = np.empty(nh, np.int32), np.empty((nh, 3), np.int32)
    n_ch, n_cv = C.c_int(0), C.c_int(0)
    ref = lambda v: C.cast(C.byref(v), C.c_void_p)
    _lib.check(lib.femb_setup_coarsen_count_host(nv, nh, _hp(xl), _hp(hx), _hp(o), C.c_float(coarse_size), _hp(hex_id), _hp(key),
                                                 ref(n_ch), ref(n_cv)), "femb_setup_coarsen_count_host")
    n_ch, n_cv = n_ch.value, n_cv.value
    coarse_hex, coarse_vert = np.empty((n_ch, 8), np.int32), np.empty((n_cv, 3), F32)
    fc = np.empty((n_ch, 27), np.int32)
    cv2i, ci2v, ccvn = np.empty(n_cv, np.int32), np.empty(n_cv, np.int32), np.empty(9, np.int32)
    _lib.check(lib.femb_setup_coarsen_fill_host(nv, nh, _hp(xl), _hp(hx), _hp(o), C.c_float(fine_size), C.c_float(coarse_size),
                                                _hp(hex_id), _hp(key), n_ch, n_cv, _hp(coarse_hex), _hp(coarse_vert), _hp(fc),
                                                _hp(cv2i), _hp(ci2v), _hp(ccvn)), "femb_setup_coarsen_fill_host")
    return hex_id, coarse_hex, coarse_vert, fc, cv2i, ci2v, [int(c) for c in ccvn]


def _interp_numpy(fc, chex, n_fine, n_coarse):
    """object.py:265-420 for one level pair (numpy restatement): the interpolation entries (coarse vertex, fine vertex,
    position k of the coarse vertex among the fine vertex's parents, weight) in the reference's order, the column sums and
    the normalised weights."""
    chex = chex.astype(np.int64)
    flat = fc.reshape(-1).astype(np.int64)
    pos_valid = np.nonzero(flat != -1)[0]
    first_pos = np.full(n_fine, -1, dtype=np.int64)                    # first (HEX, slot) that names each fine vertex:
    first_pos[flat[pos_valid][::-1]] = pos_valid[::-1]                 # assigned back to front, the first one stays
    pos = np.sort(first_pos[first_pos >= 0])                           # (HEX,slot) that handles each fine vertex
    HEX, s = pos // 27, pos % 27
    cnt = _ST_CNT[s]
    ent_owner = np.repeat(np.arange(len(pos)), cnt)
    ent_k = np.arange(len(ent_owner)) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    coarse_id = chex[HEX[ent_owner], _ST_CORNERS[s[ent_owner], ent_k]]
    fine_id = flat[pos][ent_owner]
    w = _ST_W[s[ent_owner]]
    norm = np.zeros(n_coarse, dtype=F32)
    np.add.at(norm, coarse_id, w)                                      # dyadic weights: exact in any order
    w_norm = (w / norm[coarse_id]).astype(F32)
    return coarse_id.astype(np.int32), fine_id.astype(np.int32), ent_k, w, norm, w_norm


def _interp_native(fc, chex, n_fine, n_coarse):
    """object.py:265-420 for one level pair through femb_setup_interp_count_host / _fill_host."""
    from . import _lib
    lib = _lib.load()
    fc = np.ascontiguousarray(fc, dtype=np.int32)
    chex = np.ascontiguousarray(chex, dtype=np.int32)
    n_ch = fc.shape[0]
    first_pos = np.empty(n_fine, np.int32)
    n_ent = C.c_int(0)
    _lib.check(lib.femb_setup_interp_count_host(n_fine, n_ch, _hp(fc), _hp(first_pos), C.cast(C.byref(n_ent), C.c_void_p)),
               "femb_setup_interp_count_host")
    n_ent = n_ent.value
    coarse_id, fine_id, ent_k = np.empty(n_ent, np.int32), np.empty(n_ent, np.int32), np.empty(n_ent, np.uint8)
    w, norm, w_norm = np.empty(n_ent, F32), np.empty(n_coarse, F32), np.empty(n_ent, F32)
    _lib.check(lib.femb_setup_interp_fill_host(n_fine, n_coarse, n_ch, _hp(fc), _hp(chex), _hp(first_pos), n_ent, _hp(coarse_id),
                                               _hp(fine_id), _hp(ent_k), _hp(w), _hp(norm), _hp(w_norm)), "femb_setup_interp_fill_host")
    return coarse_id, fine_id, ent_k, w, norm, w_norm


def build_hierarchy_numpy(points, hexagons, dx, layers=None, corrected_spacing=False, origin=None) -> Hierarchy:
    """build_hierarchy with the colour order, the coarsening and the interpolation entries evaluated by the numpy
    restatements instead of csrc/femb_setup_hierarchy.cu (tests compare the two bit for bit)."""
    return build_hierarchy(points, hexagons, dx, layers, corrected_spacing, origin, native=False)


def build_hierarchy(points: np.ndarray, hexagons: np.ndarray, dx: float, layers: int | None = None,
                    corrected_spacing: bool = False, origin=None, native: bool = True) -> Hierarchy:
    """object.py:116-697.  ``layers`` overrides the reference's level-count formula (benchmarks only).
    ``corrected_spacing`` (off by default = the reference's behaviour) fixes object.py:175-179, which coarsens every
    level >= 1 with spacing 2 dx so that levels >= 3 are copies of level 2: with it level l has spacing dx 2^l and the
    hierarchy really gets coarser.
    ``origin`` (default: the mesh's own bounding-box minimum, like the reference) is the lattice origin the colours and
    the coarse cells are counted from; a slab of a partitioned mesh passes the GLOBAL minimum so that its colours and
    coarse cells coincide with the global ones whatever plane it starts on (distributed.py)."""
    x0 = np.ascontiguousarray(points, dtype=F32)
    hex0 = np.ascontiguousarray(hexagons, dtype=np.int32)
    if x0.ndim != 2 or x0.shape[1] != 3 or hex0.ndim != 2 or hex0.shape[1] != 8:
        raise ValueError(f"points must be [N,3] and hexagons [M,8], got {x0.shape} and {hex0.shape}")
    if x0.shape[0] == 0 or hex0.shape[0] == 0:
        raise ValueError("empty mesh: no vertices or no hexahedra")
    if hex0.min() < 0 or hex0.max() >= x0.shape[0]:
        raise ValueError(f"hexagon vertex ids must lie in [0, {x0.shape[0]}), got [{hex0.min()}, {hex0.max()}]")
    if not np.isfinite(x0).all():
        raise ValueError("non-finite vertex coordinates")
    box_min = np.minimum(x0.min(axis=0), F32(1000.0)).astype(F32)
    box_max = np.maximum(x0.max(axis=0), F32(-1000.0)).astype(F32)
    layer = level_count(box_min, box_max, dx) if layers is None else int(layers)
    if origin is not None:
        box_min = np.asarray(origin, dtype=F32).reshape(3)
    if layer < 1:
        raise MeshTooSmallError(f"level count {layer} < 1: the smallest bounding-box extent must be at least (about) 4 dx "
                                f"(extent {float((box_max - box_min).min()):.6g}, dx {dx:.6g}; object.py:143)")

    if native:                                                             # object.py:147-158
        v2i, i2v, cvn = _color_order_native(x0, box_min, dx)
    else:
        t = _ijk(x0, box_min, dx, 0.5) % 2
        color = (t[:, 0] + 2 * t[:, 1] + 4 * t[:, 2]).astype(np.int64)
        v2i, i2v, cvn = _color_order(color)
    coarsen = _coarsen_native if native else _coarsen_numpy
    interp = _interp_native if native else _interp_numpy
    H = Hierarchy(dx=dx, layer=layer, box_min=box_min, box_max=box_max, dims=[x0.shape[0]], hexs=[hex0], x=[x0],
                  fine_coarse=[], color_vertex_num=[cvn], vertex2index=[v2i], index2vertex=[i2v], spacing=[float(dx)])

    for lay in range(layer - 1):                                           # object.py:171-241
        fine_size = dx if lay == 0 else dx * 2                             # sic (object.py:175-179)
        if corrected_spacing:
            fine_size = dx * (2 ** lay)
        coarse_size = fine_size * 2
        hex_id, coarse_hex, coarse_vert, fc, cv2i, ci2v, ccvn = coarsen(H.x[lay], H.hexs[lay], box_min, fine_size, coarse_size)
        n_cv = coarse_vert.shape[0]
        H.dims.append(n_cv); H.hexs.append(coarse_hex); H.x.append(coarse_vert); H.fine_coarse.append(fc)
        H.fine_hex_parent.append(hex_id)
        H.spacing.append(float(coarse_size))
        H.color_vertex_num.append(ccvn); H.vertex2index.append(cv2i); H.index2vertex.append(ci2v)

    n0 = H.dims[0]
    H.fix_idx = np.full((n0, 8), -1, dtype=np.int32)
    H.fix_values = np.zeros((n0, 8), dtype=F32)
    for lay in range(layer - 1):                                           # object.py:265-443
        coarse_id, fine_id, ent_k, w, norm, w_norm = interp(H.fine_coarse[lay], H.hexs[lay + 1], H.dims[lay], H.dims[lay + 1])
        v2i_c, v2i_f = H.vertex2index[lay + 1], H.vertex2index[lay]
        rowInd, colInd = v2i_c[coarse_id], v2i_f[fine_id]
        H.U_triplets.append(dict(rowInd=rowInd.astype(np.int32), colInd=colInd.astype(np.int32),
                                 rowInd_noOrder=coarse_id, colInd_noOrder=fine_id,
                                 value=w, norm=norm))
        nc, nf = H.dims[lay + 1], H.dims[lay]
        ut, ut_hat = coo_to_csr_pair(nc, nf, rowInd, colInd, w_norm, w)                 # same pattern, two weight sets
        utn, utn_hat = coo_to_csr_pair(nc, nf, coarse_id, fine_id, w_norm, w)
        H.Ut.append(ut); H.Ut_noOrder.append(utn); H.Ut_hat.append(ut_hat); H.Ut_noOrder_hat.append(utn_hat)
        if lay == 0:
            H.fix_idx[fine_id, ent_k] = coarse_id
            H.fix_values[fine_id, ent_k] = w
    for lay in range(layer - 1):
        us, us_hat = transposed_pair(H.Ut[lay], H.Ut_hat[lay])
        usn, usn_hat = transposed_pair(H.Ut_noOrder[lay], H.Ut_noOrder_hat[lay])
        H.Us.append(us); H.Us_noOrder.append(usn); H.Us_hat.append(us_hat); H.Us_noOrder_hat.append(usn_hat)

    for lay in range(layer):                                               # object.py:485-697
        H.pattern.append(build_pattern(H.hexs[lay], H.vertex2index[lay], H.color_vertex_num[lay], H.dims[lay]))
    H.shapeFuncGrad = shape_func_grad()
    return H


HEX_EDGES = np.array([[0, 1], [0, 2], [1, 3], [2, 3], [4, 5], [4, 6], [5, 7], [6, 7], [0, 4], [1, 5], [2, 6], [3, 7]], dtype=np.int64)


def edge_index(hexagons: np.ndarray) -> np.ndarray:
    """exp3/simulation/object.py:258-276: the undirected edges (id0 < id1) of a hexahedral mesh, [2, E] int64, numbered in
    the order the reference's dictionary first meets them (hexahedra ascending, the 12 local edges in `hex_edge` order)."""
    h = np.asarray(hexagons, dtype=np.int64)
    a, b = h[:, HEX_EDGES[:, 0]].reshape(-1), h[:, HEX_EDGES[:, 1]].reshape(-1)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    n = int(h.max(initial=-1)) + 1
    _, uniq = _first_seen_ids(lo * n + hi)
    return np.stack([uniq // max(n, 1), uniq % max(n, 1)]).astype(np.int64)


def prolongation_chain(H: "Hierarchy"):
    """exp3/simulation/object.py:294, 429, 438-441: A_noOrder[l] = the un-normalised interpolation weights as a scalar CSR
    matrix dims[l] x dims[l+1] (vertex order), chained so that entry l maps level l+1 straight to level 0
    (A_noOrder[l+1] = A_noOrder[l] A_noOrder[l+1]).  Returns scipy CSR matrices with sorted columns (bsr_mm's order)."""
    import scipy.sparse as sp
    out = []
    for l in range(H.layer - 1):
        u = H.Us_noOrder_hat[l]
        m = sp.csr_matrix((u.weights.astype(F32), u.columns, u.offsets), shape=(u.nrow, u.ncol))
        if l > 0:
            m = (out[l - 1] @ m).tocsr()
        m.sort_indices()
        out.append(m.astype(F32))
    return out


def default_pins(points: np.ndarray, dx: float) -> np.ndarray:
    """object.py:762-773: every vertex with y >= y_max - dx - 1e-5 (float32 arithmetic)."""
    y = np.asarray(points, dtype=F32)[:, 1]
    thr = F32(F32(y.max() - F32(dx)) - F32(1e-5))
    return np.nonzero(y >= thr)[0].astype(np.int32)


def build_galerkin_plan(H: Hierarchy, l: int) -> dict:
    """Fixed patterns and slot plans of the Galerkin product A_{l+1} = Ut_hat[l] (A_l Us_hat[l]) through the host routines of
    the C ABI (csrc/femb_setup.cu); same arrays as `build_galerkin_plan_numpy` (tests/test_hierarchy.py)."""
    from . import _lib
    lib = _lib.load()
    p, pc = H.pattern[l], H.pattern[l + 1]
    nf, nc = H.dims[l], H.dims[l + 1]
    us, ut = H.Us_hat[l], H.Ut_hat[l]
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    fL_ptr, fL_col, fU_ptr, fU_col = i32(p.L_ptr), i32(p.L_col), i32(p.U_ptr), i32(p.U_col)
    cL_ptr, cL_col, cU_ptr, cU_col = i32(pc.L_ptr), i32(pc.L_col), i32(pc.U_ptr), i32(pc.U_col)
    us_off, us_col, us_w = i32(us.offsets), i32(us.columns), np.ascontiguousarray(us.weights, dtype=np.float32)
    ut_off, ut_col = i32(ut.offsets), i32(ut.columns)
    t_ptr = np.empty(nf + 1, np.int32)
    _lib.check(lib.femb_setup_galerkin_count_host(nf, _hp(fL_ptr), _hp(fL_col), _hp(fU_ptr), _hp(fU_col), _hp(us_off), _hp(us_col),
                                                  _hp(t_ptr)), "femb_setup_galerkin_count_host")
    t_nnz = int(t_ptr[nf])
    t_col = np.empty(t_nnz, np.int32)
    c_ptr = np.empty(nc + 1, np.int32)
    c_col, c_slot = np.empty(pc.nnz, np.int32), np.empty(pc.nnz, np.int32)
    t_slot = np.empty((p.nnz, 8), np.uint8)
    c_slot8 = np.empty((t_nnz, 8), np.uint8)
    ut_q = np.empty(ut.nnz, np.uint8)
    dyadic = C.c_int(0)
    _lib.check(lib.femb_setup_galerkin_fill_host(nf, nc, _hp(fL_ptr), _hp(fL_col), _hp(fU_ptr), _hp(fU_col), _hp(cL_ptr), _hp(cL_col),
                                                 _hp(cU_ptr), _hp(cU_col), _hp(us_off), _hp(us_col), _hp(us_w), _hp(ut_off), _hp(ut_col),
                                                 _hp(t_ptr), _hp(t_col), _hp(c_ptr), _hp(c_col), _hp(c_slot), _hp(t_slot), _hp(c_slot8),
                                                 _hp(ut_q), C.cast(C.byref(dyadic), C.c_void_p)), "femb_setup_galerkin_fill_host")
    return dict(n_fine=nf, n_coarse=nc, t_ptr=t_ptr, t_col=t_col, c_ptr=c_ptr, c_col=c_col, c_slot=c_slot, t_slot=t_slot,
                c_slot8=c_slot8, ut_q=ut_q, dyadic=bool(dyadic.value))


def build_galerkin_plan_numpy(H: Hierarchy, l: int) -> dict:
    """Fixed patterns and slot plans of the Galerkin product A_{l+1} = Ut_hat[l] (A_l Us_hat[l]) (numpy / scipy restatement).

    Replaces the symbolic half of ``bsr_mm`` (sparse.py:1018-1091), which the reference repeats every step
    (object.py:1259).  Returns numpy arrays:
      t_ptr/t_col   CSR pattern of T = A_l Us_hat (columns sorted)
      c_ptr/c_col   (row,col)-sorted pattern of A_{l+1}; c_slot = its [L;D;U] slot (inverse of UtAUs_block_offset)
      t_slot[s,p]   fine block s (in [L;D;U] slot order, column j), p-th parent J of j -> position of J in T row
      c_slot8[t,q]  T block t = (i,J), q-th parent I of i -> position of J in the sorted row I of A_{l+1}
      ut_q[e]       Ut_hat entry e = (I,i) -> q with parent_q(i) == I
    """
    import scipy.sparse as sp
    p, pc = H.pattern[l], H.pattern[l + 1]
    nf, nc = H.dims[l], H.dims[l + 1]
    us, ut = H.Us_hat[l], H.Ut_hat[l]
    A = sp.csr_matrix((np.ones(p.nnz, np.int32), (p.row, p.col)), shape=(nf, nf))
    Pm = sp.csr_matrix((np.ones(us.nnz, np.int32), us.columns, us.offsets), shape=(nf, nc))
    T = (A @ Pm).tocsr()
    T.sort_indices()
    R = sp.csr_matrix((np.ones(ut.nnz, np.int32), ut.columns, ut.offsets), shape=(nc, nf))
    Cc = (R @ T).tocsr()
    Cc.sort_indices()
    # the product pattern must equal the coarse hex adjacency (block_values_reorder relies on it, object.py:1261)
    inv = np.empty(pc.nnz, dtype=np.int64)
    inv[pc.block_offset] = np.arange(pc.nnz)                       # sorted rank -> [L;D;U] slot
    if Cc.nnz != pc.nnz or not np.array_equal(Cc.indices, pc.col[inv]) or not np.array_equal(
            np.repeat(np.arange(nc), np.diff(Cc.indptr)), pc.row[inv]):
        raise ValueError(f"Galerkin pattern of level {l + 1} differs from its hex adjacency pattern")
    t_ptr = T.indptr.astype(np.int64)
    t_col = T.indices.astype(np.int64)
    c_ptr = Cc.indptr.astype(np.int64)
    us_off = us.offsets.astype(np.int64)
    us_col = us.columns.astype(np.int64)
    cntp = np.diff(us_off)
    if cntp.max(initial=0) > 8 or np.diff(t_ptr).max(initial=0) > 27 or np.diff(c_ptr).max(initial=0) > 27:
        raise ValueError("Galerkin plan: more than 8 parents or more than 27 blocks in a row")
    Tkeys = np.repeat(np.arange(nf, dtype=np.int64), np.diff(t_ptr)) * nc + t_col
    Ckeys = np.repeat(np.arange(nc, dtype=np.int64), np.diff(c_ptr)) * nc + Cc.indices.astype(np.int64)

    def expand(rows_of, cols_for_parents, keys_sorted, row_ptr, key_row_is_parent):
        n = len(cols_for_parents)
        out = np.full((n, 8), 255, dtype=np.uint8)
        CH = 1 << 22
        for s0 in range(0, n, CH):
            src = cols_for_parents[s0:s0 + CH]
            np_s = cntp[src]
            owner = np.repeat(np.arange(len(src), dtype=np.int64), np_s)
            pidx = np.arange(len(owner), dtype=np.int64) - np.repeat(np.cumsum(np_s) - np_s, np_s)
            par = us_col[us_off[src[owner]] + pidx]
            other = rows_of[s0:s0 + CH][owner]
            if key_row_is_parent:
                key, row = par * nc + other, par
            else:
                key, row = other * nc + par, other
            pos = np.searchsorted(keys_sorted, key)
            if not np.array_equal(keys_sorted[np.minimum(pos, len(keys_sorted) - 1)], key):
                raise ValueError("Galerkin plan: product pattern is missing an entry")
            out[s0 + owner, pidx] = (pos - row_ptr[row]).astype(np.uint8)
        return out

    t_slot = expand(p.row.astype(np.int64), p.col.astype(np.int64), Tkeys, t_ptr, False)
    # dyadic interpolation weights (always the case for the reference's Us_hat) ride in bits 5-6 of the plan byte
    expo = -np.log2(us.weights.astype(np.float64))
    dyadic = bool(us.nnz == 0 or (np.all(expo == np.round(expo)) and expo.min() >= 0 and expo.max() <= 3))
    if dyadic:
        ebits = np.zeros((nf, 8), dtype=np.uint8)
        rows_u = np.repeat(np.arange(nf, dtype=np.int64), cntp)
        ebits[rows_u, np.arange(us.nnz, dtype=np.int64) - us_off[rows_u]] = expo.astype(np.uint8) << 5
        valid = t_slot != 255
        t_slot = np.where(valid, t_slot | ebits[p.col.astype(np.int64)], 255).astype(np.uint8)
    t_rows = np.repeat(np.arange(nf, dtype=np.int64), np.diff(t_ptr))
    c_slot8 = expand(t_col, t_rows, Ckeys, c_ptr, True)
    ut_rows = np.repeat(np.arange(nc, dtype=np.int64), np.diff(ut.offsets))
    ukeys = np.repeat(np.arange(nf, dtype=np.int64), cntp) * nc + us_col
    ut_cols = ut.columns.astype(np.int64)
    ut_q = (np.searchsorted(ukeys, ut_cols * nc + ut_rows) - us_off[ut_cols]).astype(np.uint8)
    return dict(n_fine=nf, n_coarse=nc, t_ptr=T.indptr.astype(np.int32), t_col=T.indices.astype(np.int32),
                c_ptr=Cc.indptr.astype(np.int32), c_col=Cc.indices.astype(np.int32), c_slot=inv.astype(np.int32),
                t_slot=t_slot, c_slot8=c_slot8, ut_q=ut_q, dyadic=dyadic)
