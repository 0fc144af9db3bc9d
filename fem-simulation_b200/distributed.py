"""Mesh-partitioned (domain-decomposed) FAS step over the GPUs of one box: one process per GPU, NCCL halo exchange.

The reference is single-GPU; this is the B200 build's addition (BASELINE.json
configs[4]: "FAS multigrid quasi-static step, mesh-partitioned at 1/2/4/8 B200").

Decomposition.  The voxel lattice is cut into slabs of planes along z; rank p owns the vertex planes [z_p, z_{p+1})
(cuts on even planes so the 2:1 coarse level nests; the last rank also owns the top plane).  Its local mesh is the
cells [z_p - 2, z_{p+1}): every owned fine row AND every owned coarse row (the re-discretised coarse Hessian of the
FAS step, object.py:1499) has all its incident elements locally -- assembly needs no reduction, the overlap elements
are computed redundantly.  Ghost vertex planes: z_p-2, z_p-1 below, z_{p+1} above (coarse: z_p-2 and z_{p+1}).

Each rank runs the ordinary single-GPU ``Object`` on its local mesh (global lattice parity is preserved, so the
8-colouring and the coarse cells coincide with the global ones) and exchanges ghost values with its two neighbours:
  * inside the symmetric Gauss-Seidel sweep after each colour GROUP: a z-plane has a single z-parity, so the four
    colours of the boundary plane are {4 cz + 0..3}; rows of one group never touch ghost rows of the same group, so
    one exchange after colours 7..4 and one after 3..0 (and the mirror in the forward sweep) reproduce the
    single-GPU sweep exactly -- 4 exchanges per iteration instead of 16;
  * the restricted state x1 = Ut_noOrder x (coarse ghosts) and the coarse correction dx1 before prolongation.
Six small neighbour messages per FASv3 step, no all-reduce on the data path.  Every local row then performs the same
floating-point operations as on one GPU; the only difference is the order of the <= 8 terms of the prolongation sum
(the local coarse numbering differs), i.e. results agree to 1 ulp-level round-off (tests/test_multigpu.py: 6e-8).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, mesh
from .object import Object


def slab_planes(nz_total: int, world: int, rank: int):
    """Owned vertex planes [z0, z1) of ``rank`` (the last rank also owns plane nz_total) and its local cell range."""
    if nz_total % world or (nz_total // world) % 2:
        raise ValueError("cells along z must split into an even number of planes per rank")
    nzl = nz_total // world
    z0, z1 = rank * nzl, (rank + 1) * nzl
    clo, chi = max(z0 - 2, 0), min(z1, nz_total)
    return z0, z1, clo, chi


class HaloPlan:
    """Per-neighbour send/recv index lists (into a level's vector in a given ordering), matched by lattice order."""

    def __init__(self):
        self.send = {}     # neighbour rank -> LongTensor of local rows
        self.recv = {}

    def neighbours(self):
        return sorted(set(self.send) | set(self.recv))


def exchange(vec: torch.Tensor, plan: HaloPlan, group=None):
    """vec[recv rows] <- neighbour's vec[send rows] (vec is [n,3] or [n]); all neighbours in one batched P2P call.
    One gather into a reused send buffer, one batched NCCL group, one scatter from a reused receive buffer."""
    nbs = plan.neighbours()
    if not nbs:
        return
    key = (vec.dtype, tuple(vec.shape[1:]), vec.device)
    cache = plan.__dict__.setdefault("_cache", {})
    c = cache.get(key)
    if c is None:
        s_nb = [nb for nb in nbs if nb in plan.send and plan.send[nb].numel()]
        r_nb = [nb for nb in nbs if nb in plan.recv and plan.recv[nb].numel()]
        s_idx = torch.cat([plan.send[nb] for nb in s_nb]) if s_nb else None
        r_idx = torch.cat([plan.recv[nb] for nb in r_nb]) if r_nb else None
        tail = tuple(vec.shape[1:])
        c = dict(s_nb=s_nb, r_nb=r_nb, s_idx=s_idx, r_idx=r_idx,
                 s_len=[plan.send[nb].numel() for nb in s_nb], r_len=[plan.recv[nb].numel() for nb in r_nb],
                 s_buf=torch.empty((s_idx.numel(),) + tail, dtype=vec.dtype, device=vec.device) if s_nb else None,
                 r_buf=torch.empty((r_idx.numel(),) + tail, dtype=vec.dtype, device=vec.device) if r_nb else None)
        cache[key] = c
    ops = []
    if c["s_nb"]:
        torch.index_select(vec, 0, c["s_idx"], out=c["s_buf"])
        o = 0
        for nb, n in zip(c["s_nb"], c["s_len"]):
            ops.append(dist.P2POp(dist.isend, c["s_buf"][o:o + n], nb, group))
            o += n
    if c["r_nb"]:
        o = 0
        for nb, n in zip(c["r_nb"], c["r_len"]):
            ops.append(dist.P2POp(dist.irecv, c["r_buf"][o:o + n], nb, group))
            o += n
    if not ops:
        return
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    if c["r_nb"]:
        vec.index_copy_(0, c["r_idx"], c["r_buf"])


def build_halo_plans(nx, ny, z0, z1, clo, chi, nz_total, rank, world, coarse_xyz, coarse_v2i, fine_v2i, dx, device):
    """Fine plans (vertex order and index order) and coarse plans (vertex order, index order)."""
    ps = (nx + 1) * (ny + 1)

    def fine_plane(gz):
        return torch.arange((gz - clo) * ps, (gz - clo + 1) * ps, dtype=torch.long)

    # coarse vertices: lattice key from their lattice positions (object.py:230), sorted (z,y,x) inside a plane
    key = np.rint(coarse_xyz / np.float32(2 * dx)).astype(np.int64)          # positions are key*2dx + box_min (x,y min = 0)
    kz_local = np.rint((coarse_xyz[:, 2] - np.float32(clo * dx)) / np.float32(2 * dx)).astype(np.int64)
    gz_c = kz_local * 2 + clo

    def coarse_plane(gz):
        ids = np.nonzero(gz_c == gz)[0]
        order = np.lexsort((key[ids, 0], key[ids, 1]))
        return torch.from_numpy(ids[order]).long()

    fv, fi, cv, ci = HaloPlan(), HaloPlan(), HaloPlan(), HaloPlan()
    if rank > 0:
        fv.send[rank - 1] = fine_plane(z0)
        fv.recv[rank - 1] = torch.cat([fine_plane(z0 - 2), fine_plane(z0 - 1)])
        cv.send[rank - 1] = coarse_plane(z0)
        cv.recv[rank - 1] = coarse_plane(z0 - 2)
    if rank < world - 1:
        fv.send[rank + 1] = torch.cat([fine_plane(z1 - 2), fine_plane(z1 - 1)])
        fv.recv[rank + 1] = fine_plane(z1)
        cv.send[rank + 1] = coarse_plane(z1 - 2)
        cv.recv[rank + 1] = coarse_plane(z1)
    f_v2i = torch.from_numpy(np.asarray(fine_v2i)).long()
    c_v2i = torch.from_numpy(np.asarray(coarse_v2i)).long()
    for src, dst, m in ((fv, fi, f_v2i), (cv, ci, c_v2i)):
        for nb in src.send:
            dst.send[nb] = m[src.send[nb]]
        for nb in src.recv:
            dst.recv[nb] = m[src.recv[nb]]
    for pl in (fv, fi, cv, ci):
        pl.send = {k: v.to(device) for k, v in pl.send.items()}
        pl.recv = {k: v.to(device) for k, v in pl.recv.items()}
    return fv, fi, cv, ci


class SlabObject:
    """One rank of the mesh-partitioned FASv3 step on an nx x ny x nz_total voxel box (dx spacing)."""

    def __init__(self, nx, ny, nz_total, dx, *, rank=None, world=None, device=None, group=None, graph_segments=True,
                 **object_kwargs):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.group = group
        self.nx, self.ny, self.nz_total, self.dx = nx, ny, nz_total, dx
        z0, z1, clo, chi = slab_planes(nz_total, self.world, self.rank)
        self.z0, self.z1, self.clo, self.chi = z0, z1, clo, chi
        pts, hx = mesh.box_mesh(nx, ny, chi - clo, dx, index_offset=(0, 0, clo))
        object_kwargs.setdefault("use_graphs", False)
        self.obj = Object((pts, hx), dx, [], device=device, layers=2, **object_kwargs)
        o = self.obj
        H = o.hierarchy
        ps = (nx + 1) * (ny + 1)
        gz = np.arange(o.N_verts) // ps + clo
        top = nz_total if self.rank == self.world - 1 else -1
        self.owned = torch.from_numpy((gz >= z0) & ((gz < z1) | (gz == top))).to(o.device)
        self.fine_v, self.fine_i, self.coarse_v, self.coarse_i = build_halo_plans(
            nx, ny, z0, z1, clo, chi, nz_total, self.rank, self.world, H.x[1], H.vertex2index[1], H.vertex2index[0], dx, o.device)
        # coarse rest state: ghost coarse vertices average over children this rank does not have (object.py:730-731)
        exchange(o.x_gpu_layer[1], self.coarse_v, group)
        o._prepare_rest(1)
        self.global_offset = ps * clo                  # local vertex id + offset = vertex id in the full box mesh
        self.graph_segments, self._graphs = graph_segments, {}
        torch.cuda.synchronize(o.device) if o.device.type == "cuda" else None

    def _seg(self, key, fn):
        """Run one exchange-free segment of the step.  With `graph_segments` the segment is captured into a CUDA graph the
        first time it runs (after executing it eagerly once, like Object._run_step) and replayed afterwards: a step is
        ~300 kernel launches, and with several ranks per host the eager launches -- not the GPUs -- set the pace.  The NCCL
        exchanges between the segments stay eager."""
        if not self.graph_segments or self.obj.device.type != "cuda":
            fn()
            return
        g = self._graphs.get(key)
        if g is None:
            fn()
            torch.cuda.synchronize(self.obj.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self._graphs[key] = g
            return
        g.replay()

    # the FASv3 loop body of object.py:1467-1511 with the halo exchanges in place
    def fasv3_step(self):
        o = self.obj
        x = o.x_gpu_layer[0]
        lv0 = C.byref(o.levels[0])
        st = o._stream
        dx0 = o.dev_delta_x[0]
        vals, b, nspd = o.MF_value_gpu.data_ptr(), o.dev_B_fixed[0].data_ptr(), o.not_spd.data_ptr()
        U, LU = _lib.PART_U, _lib.PART_L | _lib.PART_U

        def colours(first, last, parts):
            _lib.check(o.lib.femb_gs_colors(lv0, vals, b, dx0.data_ptr(), first, last, parts, nspd, st()), "femb_gs_colors")

        def seg_fine():
            o._fas_fine(x, False, True, True)                                               # :1468-1481
            dx0.zero_()
            colours(7, 4, U)                                                                # :1483, sweep in colour groups

        def seg_restrict():
            o.update_deltaX(x, dx0)                                                         # :1487
            o._transfer(o.Ut_noOrder[0], 1.0, x, 0.0, o.dev_x_solved[1])                    # :1489

        def seg_coarse():
            o._force(x, o.dev_R[0])                                                         # :1491-1495
            o._transfer(o.Ut_hat[0], 1.0, o.dev_R[0], 0.0, o.dev_B_fixed[1])                # :1497
            o._hessian(1, o.dev_x_solved[1], o.UtAUs_values[0])                             # :1499
            _lib.check(o.lib.femb_coarse_pin_penalty(C.byref(o.levels[1]), C.byref(o.Ut_noOrder_hat[0].c),
                                                     o.pin_gpu.data_ptr(), o.control_mag, o.UtAUs_values[0].data_ptr(), st()),
                       "femb_coarse_pin_penalty")
            o.PerformJacobi(layer=1, iterations=1)                                          # :1506

        def seg_prolong():
            o._transfer(o.Us_hat[0], 1.0, o.dev_delta_x[1], 0.0, dx0)                       # :1508
            o.update_deltaX(x, dx0)                                                         # :1509

        self._seg("fine", seg_fine)
        exchange(dx0, self.fine_i, self.group)
        self._seg("gs1", lambda: colours(3, 0, U))
        exchange(dx0, self.fine_i, self.group)
        self._seg("gs2", lambda: colours(0, 3, LU))
        exchange(dx0, self.fine_i, self.group)
        self._seg("gs3", lambda: colours(4, 7, LU))
        exchange(dx0, self.fine_i, self.group)
        self._seg("restrict", seg_restrict)
        exchange(o.dev_x_solved[1], self.coarse_v, self.group)
        self._seg("coarse", seg_coarse)
        exchange(o.dev_delta_x[1], self.coarse_i, self.group)
        self._seg("prolong", seg_prolong)

    def FASv3(self, iterations=100):
        for _ in range(iterations):
            self.fasv3_step()

    def owned_positions(self):
        """(global vertex ids, positions) of the owned vertices, for gathering / checking."""
        ids = torch.nonzero(self.owned).squeeze(1)
        return ids + self.global_offset, self.obj.x_gpu_layer[0][ids]

    def force_inf_norm(self):
        """max |f| over the owned rows of all ranks (the sampled quantity of object.py:1245)."""
        o = self.obj
        o._force(o.x_gpu_layer[0], o.grad_gpu, ordered=False)
        m = (o.grad_gpu.abs().amax(dim=1) * self.owned).max().reshape(1)
        if self.world > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
        return float(m.item())
