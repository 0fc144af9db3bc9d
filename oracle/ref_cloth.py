"""ORACLE (test infrastructure only) -- CPU restatement of /root/reference/exp1/cloth_simulation: ``Cloth.__init__``
(object.py:21-196 there, literal loops), ``render`` (:657-677), ``PerformConjugateGradient`` (:261-302) and the kernels of
triangle.py:5-79 / cublas.py (update_deltaX_kernel freezes vertices 0 and 65, :22-26)."""
from __future__ import annotations

import operator

import numpy as np

EPSILON = 1e-7


class RefCloth:
    def __init__(self, res_x=64, res_y=64, dtype=np.float32):
        t = self.dtype = dtype
        self.dt = 0.0033
        self.inv_dt = t(1.0 / self.dt)
        self.damping, self.k, self.g = t(0.9995), t(1000.0), t(-9.8)
        nvx = res_x + 1
        n = nvx * (res_y + 1)
        x = np.zeros((n, 3), np.float32)
        for i in range(n):                                            # :36-41
            x[i][0] = np.float32(i % nvx) * np.float32(1.0 / res_x)
            x[i][1] = np.float32(i // nvx) * np.float32(1.0 / res_y)
        edge = np.zeros((res_x * (res_y + 1) + res_y * (res_x + 1) + res_x * res_y, 2), np.int32)
        for i in range(res_y + 1):                                    # :58-73
            for j in range(res_x):
                edge[i * res_x + j] = (i * nvx + j, i * nvx + j + 1)
        for i in range(res_x + 1):
            for j in range(res_y):
                edge[(res_y + 1) * res_x + i * res_y + j] = (j * nvx + i, (j + 1) * nvx + i)
        for i in range(res_y):
            for j in range(res_x):
                edge[(res_y + 1) * res_x + (res_x + 1) * res_y + i * res_x + j] = (i * nvx + j, (i + 1) * nvx + j + 1)
        self.N_verts, self.N_edge, self.edge = n, len(edge), edge
        self.x_cpu = x
        edge_hash = {}                                                # :139-176
        for i in range(self.N_edge):
            id0, id1 = int(edge[i][0]), int(edge[i][1])
            edge_hash[(id0, id1)] = i
            edge_hash[(id1, id0)] = i + self.N_edge
        for i in range(n):
            edge_hash[(i, i)] = self.N_edge * 2 + i
        sorted_dict = dict(sorted(edge_hash.items(), key=operator.itemgetter(0)))
        self.MF_nnz = n + 2 * self.N_edge
        self.MF_row = np.zeros(self.MF_nnz, np.int32); self.MF_col = np.zeros(self.MF_nnz, np.int32)
        self.d_offset = np.zeros(n, np.int32)
        for index, key in enumerate(sorted_dict.keys()):
            self.MF_row[index], self.MF_col[index] = key
            sorted_dict[key] = index
            if key[0] == key[1]:
                self.d_offset[key[0]] = index
        self.offset = np.zeros(4 * self.N_edge, np.int32)
        for i in range(self.N_edge):
            id0, id1 = int(edge[i][0]), int(edge[i][1])
            self.offset[4 * i:4 * i + 4] = (sorted_dict[(id0, id0)], sorted_dict[(id0, id1)], sorted_dict[(id1, id0)], sorted_dict[(id1, id1)])
        self.dev_x = x.astype(t)
        d = self.dev_x[edge[:, 0]] - self.dev_x[edge[:, 1]]
        self.rest_length = np.sqrt((d * d).sum(-1)).astype(t)          # compute_rest_length
        self.m = np.ones(n, t)
        z = lambda: np.zeros((n, 3), t)
        self.dev_v, self.dev_old_x, self.dev_inertia_x = z(), z(), z()
        self.dev_B, self.dev_P, self.dev_AP, self.dev_delta_x, self.dev_B_fixed, self.grad = z(), z(), z(), z(), z(), z()
        self.MF_value = np.zeros((self.MF_nnz, 3, 3), t)

    def _A_mv(self, v):
        out = np.zeros_like(v)
        np.add.at(out, self.MF_row, np.einsum("nij,nj->ni", self.MF_value, v[self.MF_col]))
        return out

    def PerformConjugateGradient(self, iterations=10, tol=1e-5):
        t = self.dtype
        self.dev_delta_x[...] = 0
        self.dev_B[...] = self.dev_B_fixed
        r1 = t((self.dev_B * self.dev_B).sum(dtype=t))
        if r1 < EPSILON:
            return
        r, r0, k = r1, t(0), 1
        while r1 > tol * r and k <= iterations:
            if k > 1:
                self.dev_P[...] = self.dev_P * t(r1 / r0) + self.dev_B
            else:
                self.dev_P[...] = self.dev_B_fixed
            self.dev_AP[...] = self._A_mv(self.dev_P)
            dot = t((self.dev_P * self.dev_AP).sum(dtype=t))
            if dot < 1e-10:
                break
            alpha = t(r1 / dot)
            self.dev_delta_x[...] = self.dev_delta_x + alpha * self.dev_P
            self.dev_B[...] = self.dev_B - alpha * self.dev_AP
            r0 = r1
            r1 = t((self.dev_B * self.dev_B).sum(dtype=t))
            k += 1

    def render(self):
        t = self.dtype
        x, e = self.dev_x, self.edge
        self.dev_old_x[...] = x
        self.dev_v[...] = self.dev_v * self.damping
        x[...] = x + self.dev_v * t(self.dt)
        self.dev_inertia_x[...] = x
        self.MF_value[...] = 0
        self.MF_value[self.d_offset] += (self.m * self.inv_dt * self.inv_dt)[:, None, None] * np.eye(3, dtype=t)
        d = x[e[:, 0]] - x[e[:, 1]]
        ln = np.sqrt((d * d).sum(-1))
        a = self.k * self.rest_length / ln
        b = a / ln / ln
        H = (self.k - a)[:, None, None] * np.eye(3, dtype=t) + b[:, None, None] * np.einsum("ni,nj->nij", d, d)
        np.add.at(self.MF_value, self.offset[0::4], H); np.add.at(self.MF_value, self.offset[1::4], -H)
        np.add.at(self.MF_value, self.offset[2::4], -H); np.add.at(self.MF_value, self.offset[3::4], H)
        self.grad[...] = 0
        gl = (self.k * (self.rest_length / ln - t(1.0)))[:, None] * d
        np.add.at(self.grad, e[:, 0], gl); np.add.at(self.grad, e[:, 1], -gl)
        self.grad[:, 1] += self.m * self.g
        self.grad -= (self.m * self.inv_dt * self.inv_dt)[:, None] * (x - self.dev_inertia_x)
        self.dev_B_fixed[...] = self.grad
        self.PerformConjugateGradient(iterations=5)
        upd = np.ones(self.N_verts, bool); upd[[0, 65]] = False
        x[upd] = x[upd] + self.dev_delta_x[upd]
        self.dev_v[...] = (x - self.dev_old_x) * self.inv_dt
