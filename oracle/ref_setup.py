"""ORACLE (test infrastructure, never shipped, never timed as the product).

Literal CPU restatement of the *setup* half of the reference hot path:
``Object.__init__`` of ``exp1/quasi_simulation/object.py:21-865`` (reference paths are
relative to /root/reference).  It keeps the reference's Python loops and dictionaries
on purpose -- the integer outputs (colour/index order, coarse meshes, fine_coarse,
interpolation triplets, [L;D;U] slot maps, element->nnz maps) must be bit-exact and the
numbering is defined by dictionary insertion order.  Only small meshes are practical
(minutes at ~5 k hexes); the product's vectorised setup is tested against this.

Parity status: **pinned only through the reference's own Python source executed under
the Warp/pyvista/OpenGL stand-ins of ``oracle/warp_emulator``** (see
``tests/golden/README.md``).  The reference ships no tests or golden vectors
(SURVEY.md section 4/8c); the third-party arithmetic (warp-lang, pyvista) is restated.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.
"""
from __future__ import annotations

import operator

import numpy as np
import torch

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# cpu_function.py:4-20
# ----------------------------------------------------------------------------------------------
def cal_shape_func_grad():
    """shapeFuncGrad[a][q][c]; cpu_function.py:4-9 with the tables of object.py:701-722.

    Evaluated with 0-dim float32 torch tensors exactly like the reference (so the
    rounding of every product is the reference's)."""
    sfg = torch.zeros((8, 8, 3), dtype=torch.float32)
    help_ = torch.tensor([[-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1],
                          [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], dtype=torch.float32)
    g = 0.57735
    quadrature = torch.tensor([[-g, -g, -g], [-g, -g, g], [-g, g, -g], [-g, g, g],
                               [g, -g, -g], [g, -g, g], [g, g, -g], [g, g, g]])
    for a in range(8):
        for q in range(8):
            sfg[a][q][0] = help_[a][0] * (1 + help_[a][1] * quadrature[q][1]) * (1 + help_[a][2] * quadrature[q][2]) * 0.125
            sfg[a][q][1] = help_[a][1] * (1 + help_[a][0] * quadrature[q][0]) * (1 + help_[a][2] * quadrature[q][2]) * 0.125
            sfg[a][q][2] = help_[a][2] * (1 + help_[a][1] * quadrature[q][1]) * (1 + help_[a][0] * quadrature[q][0]) * 0.125
    return sfg.numpy().copy()


def ijk_index(point, origin, spacing):
    """cpu_function.py:12-13 (float32 torch arithmetic, trunc)."""
    return ((point - origin) / spacing + torch.tensor([0.1, 0.1, 0.1])).int().tolist()


def color_ind(point, origin, spacing):
    """cpu_function.py:15-20."""
    temp = ((point - origin) / spacing + torch.tensor([0.5, 0.5, 0.5])).int()
    return int(temp[0] % 2 + 2 * (temp[1] % 2) + 4 * (temp[2] % 2))


# ----------------------------------------------------------------------------------------------
# sparse.py:173-278 (+ the native bsr_matrix_from_triplets it calls) -- structural part only
# ----------------------------------------------------------------------------------------------
def bsr_from_triplets(nrow, ncol, rows, cols, vals):
    """COO -> CSR-of-blocks: sort by (row, col), sum duplicates.  ``vals`` may be scalar
    weights [n] or blocks [n,3,3].  Restates warp's native bsr_matrix_from_triplets
    (source absent; semantics from sparse.py:184-194 and its callers)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals)
    if vals.dtype != np.float64:                      # float64 only for the oracle's float64 twin
        vals = vals.astype(F32)
    order = np.lexsort((cols, rows))
    r, c, v = rows[order], cols[order], vals[order]
    if len(r):
        new = np.ones(len(r), dtype=bool)
        new[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
        seg = np.cumsum(new) - 1
        out_v = np.zeros((seg[-1] + 1,) + v.shape[1:], dtype=v.dtype)
        np.add.at(out_v, seg, v)
        r, c = r[new], c[new]
    else:
        out_v = v
    offsets = np.zeros(nrow + 1, dtype=np.int32)
    np.add.at(offsets, r + 1, 1)
    offsets = np.cumsum(offsets).astype(np.int32)
    return dict(nrow=int(nrow), ncol=int(ncol), nnz=int(len(r)), offsets=offsets,
                columns=c.astype(np.int32), values=out_v)


def bsr_transposed(m):
    """sparse.py:301-372 (bsr_transposed); blocks here are w*I3 so the block transpose is a no-op."""
    rows = np.repeat(np.arange(m["nrow"]), np.diff(m["offsets"]))
    return bsr_from_triplets(m["ncol"], m["nrow"], m["columns"], rows, m["values"])


# ----------------------------------------------------------------------------------------------
# object.py:116-697  hierarchy, interpolation, patterns
# ----------------------------------------------------------------------------------------------
def build_setup(points, hexagons, dx, pinList, control_mag=100.0, corrected_spacing=False):
    """Restates object.py:116-865 for (points f32[N,3], hexagons i32[Nh,8] with local node
    4bx+2by+bz, dx, pinList).  Returns a dict of numpy arrays / python lists named like the
    reference attributes."""
    x = torch.tensor(np.asarray(points), dtype=torch.float32)
    hexagons = torch.tensor(np.asarray(hexagons), dtype=torch.int32)
    N_verts = x.shape[0]
    N_hexagons = hexagons.shape[0]
    pinList = list(pinList)
    S = dict(dx=dx, N_verts=N_verts, N_hexagons=N_hexagons, control_mag=control_mag)

    vertex2index = [torch.zeros((N_verts), dtype=torch.int32)]
    index2vertex = [torch.zeros((N_verts), dtype=torch.int32)]
    color_num = [8]
    color_vertex_num = [[0 for _ in range(9)]]
    color = torch.zeros((N_verts), dtype=torch.int32)

    # object.py:130-143 bounding box and level count
    box_min = torch.tensor([1000.0, 1000.0, 1000.0], dtype=torch.float32)
    box_max = torch.tensor([-1000.0, -1000.0, -1000.0], dtype=torch.float32)
    for j in range(3):
        box_min[j] = torch.min(torch.cat([x[:, j], box_min[j:j + 1]]))
        box_max[j] = torch.max(torch.cat([x[:, j], box_max[j:j + 1]]))
    delta_box = box_max - box_min
    min_box = torch.min(delta_box)
    min_box = torch.log(min_box / dx) / torch.log(torch.tensor(2.0))
    layer = int(min_box.int() - 1)
    S["layer"] = layer
    S["box_min"] = box_min.numpy().copy()
    S["box_max"] = box_max.numpy().copy()

    # object.py:147-158 colouring and index order, level 0
    for i in range(N_verts):
        color[i] = color_ind(x[i], box_min, dx)
    index_idx = 0
    for j in range(8):
        for i in range(N_verts):
            if color[i] == j:
                vertex2index[0][i] = index_idx
                index2vertex[0][index_idx] = i
                index_idx += 1
        color_vertex_num[0][j + 1] = index_idx

    dims = [N_verts]
    hexs = [hexagons]
    xs = [x]
    fine_coarse_all = []

    # object.py:171-241 coarsening
    for lay in range(layer - 1):
        coarse_hex_num = 0
        coarse_vert_num = 0
        fine_size = dx if lay == 0 else dx * 2          # object.py:175-179 (sic: 2dx for every lay>=1)
        if corrected_spacing:                            # the spacing of level `lay` is dx * 2^lay
            fine_size = dx * (2 ** lay)
        coarse_size = fine_size * 2
        hex_hash = {}
        for i in range(hexs[lay].shape[0]):
            ijk = ijk_index(xs[lay][hexs[lay][i][0]], box_min, coarse_size)
            hex_key = (ijk[0], ijk[1], ijk[2])
            if hex_key not in hex_hash:
                hex_hash[hex_key] = coarse_hex_num
                coarse_hex_num += 1
        coarse_hex = torch.zeros((coarse_hex_num, 8), dtype=torch.int32)
        vert_hash = {}
        for hex_key in hex_hash.keys():
            hex_id = hex_hash[hex_key]
            for j in range(2):
                for k in range(2):
                    for l in range(2):
                        vert_key = (hex_key[0] + j, hex_key[1] + k, hex_key[2] + l)
                        if vert_key not in vert_hash:
                            vert_hash[vert_key] = coarse_vert_num
                            coarse_vert_num += 1
                        coarse_hex[hex_id][4 * j + 2 * k + l] = vert_hash[vert_key]
        dims.append(coarse_vert_num)
        hexs.append(coarse_hex)
        fine_coarse = torch.full((coarse_hex_num, 27), -1, dtype=torch.int32)
        for i in range(hexs[lay].shape[0]):
            ijk_coarse = ijk_index(xs[lay][hexs[lay][i][0]], box_min, coarse_size)
            hex_id = hex_hash[(ijk_coarse[0], ijk_coarse[1], ijk_coarse[2])]
            for j in range(8):
                ijk_fine = ijk_index(xs[lay][hexs[lay][i][j]], box_min, fine_size)
                rel = [ijk_fine[0] - 2 * ijk_coarse[0], ijk_fine[1] - 2 * ijk_coarse[1], ijk_fine[2] - 2 * ijk_coarse[2]]
                fine_coarse[hex_id][9 * rel[0] + 3 * rel[1] + rel[2]] = hexs[lay][i][j]
        fine_coarse_all.append(fine_coarse)
        color_num.append(8)
        color_vertex_num.append([0 for _ in range(9)])
        vertex2index.append(torch.zeros((coarse_vert_num), dtype=torch.int32))
        index2vertex.append(torch.zeros((coarse_vert_num), dtype=torch.int32))
        color = torch.zeros((coarse_vert_num), dtype=torch.int32)
        coarse_vert = torch.zeros((coarse_vert_num, 3), dtype=torch.float32)
        for vert_key in vert_hash.keys():
            vert_id = vert_hash[vert_key]
            coarse_vert[vert_id] = torch.tensor(vert_key, dtype=torch.float32) * coarse_size + box_min
            color[vert_id] = 4 * (vert_key[2] % 2) + 2 * (vert_key[1] % 2) + (vert_key[0] % 2)
        xs.append(coarse_vert)
        index_idx = 0
        for j in range(8):
            for i in range(coarse_vert_num):
                if color[i] == j:
                    vertex2index[lay + 1][i] = index_idx
                    index2vertex[lay + 1][index_idx] = i
                    index_idx += 1
            color_vertex_num[lay + 1][j + 1] = index_idx

    S["dims"] = dims
    S["hexs"] = [h.numpy().copy() for h in hexs]
    S["x"] = [t.numpy().copy() for t in xs]
    S["fine_coarse"] = [t.numpy().copy() for t in fine_coarse_all]
    S["color_num"] = color_num
    S["color_vertex_num"] = color_vertex_num
    S["vertex2index"] = [t.numpy().copy() for t in vertex2index]
    S["index2vertex"] = [t.numpy().copy() for t in index2vertex]

    # object.py:247-443 interpolation operators
    fix_idx = np.full((N_verts, 8), -1, dtype=np.int32)
    fix_num = np.zeros(N_verts, dtype=np.int32)
    fix_values = np.zeros((N_verts, 8), dtype=F32)
    for name in ("Ut", "Us", "Ut_noOrder", "Us_noOrder", "Ut_hat", "Us_hat", "Ut_noOrder_hat", "Us_noOrder_hat"):
        S[name] = []
    S["U_triplets"] = []
    for lay in range(layer - 1):
        rowInd, colInd, rowInd_noOrder, colInd_noOrder, value = [], [], [], [], []
        fine_hash = {}
        fc = S["fine_coarse"][lay]
        chex = S["hexs"][lay + 1]
        v2i_c = S["vertex2index"][lay + 1]
        v2i_f = S["vertex2index"][lay]

        def emit(coarse_id, fine_id, w):
            rowInd.append(int(v2i_c[coarse_id]))
            colInd.append(int(v2i_f[fine_id]))
            rowInd_noOrder.append(int(coarse_id))
            colInd_noOrder.append(int(fine_id))
            if lay == 0:
                fix_idx[fine_id][fix_num[fine_id]] = coarse_id
                fix_values[fine_id][fix_num[fine_id]] = w
                fix_num[fine_id] += 1
            value.append(w)

        for HEX in range(fc.shape[0]):
            for i in range(3):
                for j in range(3):
                    for k in range(3):
                        fine_id = int(fc[HEX][9 * i + 3 * j + k])
                        if fine_id == -1 or fine_id in fine_hash:
                            continue
                        fine_hash[fine_id] = 0
                        par = i % 2 + j % 2 + k % 2
                        if par == 0:                                        # object.py:283-293
                            emit(chex[HEX][4 * (i // 2) + 2 * (j // 2) + (k // 2)], fine_id, 1.0)
                        elif par == 3:                                      # object.py:295-306
                            for l in range(8):
                                emit(chex[HEX][l], fine_id, 0.125)
                        elif par == 1:                                      # object.py:308-371
                            if i == 1:
                                emit(chex[HEX][4 * 0 + 2 * (j // 2) + (k // 2)], fine_id, 0.5)
                                emit(chex[HEX][4 * 1 + 2 * (j // 2) + (k // 2)], fine_id, 0.5)
                            elif j == 1:
                                emit(chex[HEX][4 * (i // 2) + 2 * 0 + (k // 2)], fine_id, 0.5)
                                emit(chex[HEX][4 * (i // 2) + 2 * 1 + (k // 2)], fine_id, 0.5)
                            elif k == 1:
                                emit(chex[HEX][4 * (i // 2) + 2 * (j // 2) + 0], fine_id, 0.5)
                                emit(chex[HEX][4 * (i // 2) + 2 * (j // 2) + 1], fine_id, 0.5)
                        elif par == 2:                                      # object.py:373-412
                            if i % 2 == 0:
                                for ii in range(2):
                                    for jj in range(2):
                                        emit(chex[HEX][4 * (i // 2) + 2 * ii + jj], fine_id, 0.25)
                            elif j % 2 == 0:
                                for ii in range(2):
                                    for jj in range(2):
                                        emit(chex[HEX][4 * ii + 2 * (j // 2) + jj], fine_id, 0.25)
                            elif k % 2 == 0:
                                for ii in range(2):
                                    for jj in range(2):
                                        emit(chex[HEX][4 * ii + 2 * jj + (k // 2)], fine_id, 0.25)
        # object.py:417-429 row normalisation (float32 accumulation in append order)
        norm = torch.zeros((dims[lay + 1]), dtype=torch.float32)
        for i in range(len(value)):
            norm[rowInd_noOrder[i]] += value[i]
        norm = norm.numpy()
        w = np.asarray(value, dtype=F32)
        w_norm = (w / norm[np.asarray(rowInd_noOrder, dtype=np.int64)]).astype(F32) if len(w) else w
        S["U_triplets"].append(dict(rowInd=np.asarray(rowInd, np.int32), colInd=np.asarray(colInd, np.int32),
                                    rowInd_noOrder=np.asarray(rowInd_noOrder, np.int32),
                                    colInd_noOrder=np.asarray(colInd_noOrder, np.int32), value=w, norm=norm.copy()))
        nc, nf = dims[lay + 1], dims[lay]
        S["Ut"].append(bsr_from_triplets(nc, nf, rowInd, colInd, w_norm))                           # object.py:433
        S["Ut_noOrder"].append(bsr_from_triplets(nc, nf, rowInd_noOrder, colInd_noOrder, w_norm))   # :434
        S["Ut_hat"].append(bsr_from_triplets(nc, nf, rowInd, colInd, w))                            # :435
        S["Ut_noOrder_hat"].append(bsr_from_triplets(nc, nf, rowInd_noOrder, colInd_noOrder, w))    # :436
    for lay in range(layer - 1):                                                                    # :439-443
        S["Us"].append(bsr_transposed(S["Ut"][lay]))
        S["Us_noOrder"].append(bsr_transposed(S["Ut_noOrder"][lay]))
        S["Us_hat"].append(bsr_transposed(S["Ut_hat"][lay]))
        S["Us_noOrder_hat"].append(bsr_transposed(S["Ut_noOrder_hat"][lay]))
    S["fix_idx"] = fix_idx
    S["fix_values"] = fix_values

    # object.py:449-697 fixed [L;D;U] pattern per level and element->nnz maps
    P = []
    for lay in range(layer):
        v2i = S["vertex2index"][lay]
        hx = S["hexs"][lay]
        cvn = color_vertex_num[lay]
        my_dict = {}
        for i in range(hx.shape[0]):
            for j in range(8):
                for k in range(8):
                    key = (int(v2i[hx[i][j]]), int(v2i[hx[i][k]]))
                    if key not in my_dict:
                        my_dict[key] = 0
        sorted_dict = dict(sorted(my_dict.items(), key=operator.itemgetter(0)))
        nnz = len(my_dict)
        D_nnz = dims[lay]
        L_nnz = (nnz - dims[lay]) >> 1
        U_nnz = L_nnz
        off_l, off_d, off_u = 0, L_nnz, L_nnz + D_nnz
        L_row = np.zeros(L_nnz, np.int32); L_col = np.zeros(L_nnz, np.int32)
        D_row = np.zeros(D_nnz, np.int32); D_col = np.zeros(D_nnz, np.int32)
        U_row = np.zeros(U_nnz, np.int32); U_col = np.zeros(U_nnz, np.int32)
        GS_L_col = np.zeros(L_nnz, np.int32); GS_L_row = np.zeros(L_nnz, np.int32)
        GS_U_col = np.zeros(U_nnz, np.int32); GS_U_row = np.zeros(U_nnz, np.int32)
        GS_U_Ptr = [0 for _ in range(9)]
        GS_L_Ptr = [0 for _ in range(9)]
        block_offset = np.zeros(nnz, np.int32)
        diag_offset = np.zeros(D_nnz, np.int32)
        i_l = i_d = i_u = i_f = 0
        p_base_l = -1
        p_base_u = -1
        now_base_l = 0
        next_base_l = cvn[1]
        r_base_u = -1
        c_base_u = 0
        for key in sorted_dict.keys():
            r, c = key
            if c == r:
                D_row[i_d] = r; D_col[i_d] = c
                sorted_dict[key] = r + off_d
                diag_offset[i_d] = r + off_d
                block_offset[i_d + off_d] = i_f
                i_d += 1; i_f += 1
            if r < c:
                while r >= c_base_u:
                    p_base_u += 1
                    GS_U_Ptr[p_base_u] = i_u
                    r_base_u = cvn[p_base_u]
                    c_base_u = cvn[p_base_u + 1]
                U_row[i_u] = r; U_col[i_u] = c
                GS_U_row[i_u] = r - r_base_u
                GS_U_col[i_u] = c - c_base_u
                block_offset[i_u + off_u] = i_f
                sorted_dict[key] = i_u + off_u
                i_u += 1; i_f += 1
            if r > c:
                while r >= next_base_l:
                    p_base_l += 1
                    GS_L_Ptr[p_base_l] = i_l
                    now_base_l = cvn[p_base_l + 1]
                    next_base_l = cvn[p_base_l + 2]
                L_row[i_l] = r; L_col[i_l] = c
                GS_L_row[i_l] = r - now_base_l
                GS_L_col[i_l] = c
                block_offset[i_l + off_l] = i_f
                sorted_dict[key] = i_l + off_l
                i_l += 1; i_f += 1
        hex_update_offset = np.zeros(hx.shape[0] * 64, np.int32)
        for i in range(hx.shape[0]):
            for j in range(8):
                for k in range(8):
                    hex_update_offset[i * 64 + j * 8 + k] = sorted_dict[(int(v2i[hx[i][j]]), int(v2i[hx[i][k]]))]
        while p_base_l < 7:
            p_base_l += 1
            GS_L_Ptr[p_base_l] = i_l
        while p_base_u < 7:
            p_base_u += 1
            GS_U_Ptr[p_base_u] = i_u
        P.append(dict(nnz=nnz, L_nnz=L_nnz, D_nnz=D_nnz, U_nnz=U_nnz, off_l=off_l, off_d=off_d, off_u=off_u,
                      L_row=L_row, L_col=L_col, D_row=D_row, D_col=D_col, U_row=U_row, U_col=U_col,
                      GS_L_row=GS_L_row, GS_L_col=GS_L_col, GS_U_row=GS_U_row, GS_U_col=GS_U_col,
                      GS_L_Ptr=GS_L_Ptr, GS_U_Ptr=GS_U_Ptr, block_offset=block_offset,
                      diag_offset=diag_offset, hex_update_offset=hex_update_offset,
                      row=np.concatenate([L_row, D_row, U_row]), col=np.concatenate([L_col, D_col, U_col])))
    S["pattern"] = P
    S["shapeFuncGrad"] = cal_shape_func_grad()

    # object.py:761-795 pins (the rest state is taken before the displacement at :775)
    S["x_rest"] = S["x"][0].copy()
    x0 = torch.tensor(S["x"][0].copy())
    if len(pinList) == 0:
        max_value = torch.max(x0[:, 1])
        for i in range(N_verts):
            if x0[i][1] >= max_value - dx - 1e-5:
                pinList.append(i)
    else:
        x0[pinList[0]][1] += 0.1
    S["x0"] = x0.numpy().copy()                      # x_gpu_layer[0] / all_pin_pos after pins
    pin = np.zeros(N_verts, np.int32)
    pin[np.asarray(pinList, dtype=np.int64)] = 1
    S["pin"] = pin
    S["pin_list"] = np.asarray(pinList, np.int32)
    S["pin_pos"] = S["x0"][S["pin_list"]].copy()
    S["N_pin"] = len(pinList)
    fixed = np.zeros((P[0]["nnz"], 3, 3), F32)
    for p in pinList:
        fixed[S["vertex2index"][0][p] + P[0]["off_d"]] = np.eye(3, dtype=F32) * F32(control_mag)
    S["MF_value_fixed"] = fixed
    return S
