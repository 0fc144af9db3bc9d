"""Mesh-partitioned (domain-decomposed) FAS and Newton-multigrid steps over the GPUs of one box: one process per GPU,
halo exchange with the two neighbouring slabs.

The reference is single-GPU; this is the B200 build's addition (BASELINE.json
configs[4] and north_star: "a 10M-tet Newton-multigrid quasi-static step on 8xB200").

Decomposition.  The voxel lattice is cut into slabs of vertex planes along z; rank p owns the planes [z_p, z_{p+1})
(the last rank also the top plane).  Cuts sit on multiples of `align` = the plane stride of the coarsest really
coarsening level (2 for the two-level FAS step, 4 for the reference's hierarchy, whose levels >= 3 repeat level 2:
object.py:175-179), so every level nests in a slab; slabs may differ by `align` planes (216 planes over 8 ranks: six
slabs of 28 and two of 24).  The local mesh of a rank is the cells [z_p - overlap, z_{p+1}) with overlap = align:

  * every owned fine row has all its hexahedra locally -> assembly needs no reduction (overlap cells are recomputed);
  * the fine rows of the `overlap - 1` planes below z_p are complete as well, so the Galerkin products
    A_{l+1} = Ut_hat A_l Us_hat of every OWNED coarse row (children reach one plane of level l below it) and of the one
    ghost plane below it that the next product reads are exact from local data: no matrix entry is ever exchanged.

Each rank runs the ordinary single-GPU ``Object`` on its local mesh (global lattice parity is preserved, so colours and
coarse cells coincide with the global ones) and exchanges VECTOR ghosts only: per level the plane z_p - stride_l below
and the plane z_{p+1} above (level 0: all overlap planes, which also keeps the ghost copy of the state bit-identical
to its owner's):

  * inside a symmetric Gauss-Seidel sweep after each colour GROUP: a z-plane has one z-parity, so its rows are the
    colours {4 cz + 0..3}; rows of a group never read ghost rows of the same group -> 4 exchanges per iteration instead
    of 16 reproduce the single-GPU sweep exactly;
  * the residual before a restriction, the accumulated correction after a prolongation, the increment before the state
    update (so that owner and ghost apply identical increments).

No all-reduce on the data path (sampled norms / energies all-reduce one scalar).  Every owned row performs the same
floating-point operations as on one GPU except for the order of the <= 8 terms of a prolongation sum (local coarse
numbering), i.e. results agree to round-off (tests/test_multigpu.py).

`exchange` is the portable (torch.distributed, eager) exchange used by the CPU tests; on CUDA the same plans are handed to
the C ABI (femb_comm_*, csrc/femb_comm.cu), on the step's stream and capturable into the step's CUDA graph: by default through
peer memory (`comm="auto"` / `"p2p"`: a push kernel stores the boundary rows straight into the neighbour's arena over NVLink
and raises an epoch flag, a pull kernel waits for the flags and scatters), else NCCL (`comm="nccl"`: pack kernel -> send/recv
-> unpack kernel).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, mesh
from .object import Object


def slab_cuts(nz_total: int, world: int, align: int = 2):
    """Plane indices [z_0 = 0, z_1, ..., z_world = nz_total]: multiples of `align`, slab sizes differ by at most `align`."""
    if align < 1 or nz_total % align:
        raise ValueError(f"cells along z ({nz_total}) must be a multiple of the cut alignment {align}")
    groups = nz_total // align
    if groups < world:
        raise ValueError(f"{nz_total} cell layers cannot be cut into {world} slabs aligned to {align} planes")
    base, extra = divmod(groups, world)
    cuts = [0]
    for r in range(world):
        cuts.append(cuts[-1] + (base + (1 if r < extra else 0)) * align)
    return cuts


def slab_planes(nz_total: int, world: int, rank: int, align: int = 2, overlap: int | None = None):
    """Owned vertex planes [z0, z1) of ``rank`` (the last rank also owns plane nz_total) and its local cell range."""
    cuts = slab_cuts(nz_total, world, align)
    overlap = align if overlap is None else overlap
    z0, z1 = cuts[rank], cuts[rank + 1]
    if rank > 0 and cuts[rank] - cuts[rank - 1] < overlap:
        raise ValueError("slabs must be at least `overlap` planes thick")
    return z0, z1, max(z0 - overlap, 0), min(z1, nz_total)


def hierarchy_align(layers: int, corrected: bool = False) -> int:
    """Plane stride of the coarsest really coarsening level: the reference's spacing saturates at 4 dx (object.py:175-179)."""
    real = layers if corrected else min(layers, 3)
    return 2 ** max(real - 1, 1)


class HaloPlan:
    """Per-neighbour send/recv row lists (into a level's vector in a given ordering), matched by lattice order."""

    def __init__(self):
        self.send = {}     # neighbour rank -> LongTensor of local rows
        self.recv = {}

    def neighbours(self):
        return sorted(set(self.send) | set(self.recv))

    def to(self, device):
        self.send = {k: v.to(device) for k, v in self.send.items()}
        self.recv = {k: v.to(device) for k, v in self.recv.items()}
        return self


def exchange(vec: torch.Tensor, plan: HaloPlan, group=None):
    """vec[recv rows] <- neighbour's vec[send rows] (vec is [n,3] or [n]); all neighbours in one batched P2P call.
    One gather into a reused send buffer, one batched group, one scatter from a reused receive buffer."""
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


def level_planes(x_level: np.ndarray, dx: float):
    """Integer lattice coordinates (ix, iy, gz) of a level's vertices in units of the FINE spacing (global: the local
    meshes are slices of the global mesh, coordinates included)."""
    return np.rint(np.asarray(x_level, dtype=np.float64) / float(dx)).astype(np.int64)


def build_level_plans(H, dx, z0, z1, cuts, rank, world, overlap):
    """Per level l: (vertex-order plan, index-order plan, owned mask).  Ghost planes of rank p at level l: below, the level's
    planes in [z_p - span_l, z_p) with span_0 = overlap (all overlap planes: the state must stay current there) and
    span_l = stride_l for l >= 1 (the one plane owned rows read); above, plane z_{p+1}.  The sender derives its lists from
    the same rule applied to the neighbour's cut, both sides ordered by (gz, iy, ix)."""
    nz_total = cuts[-1]
    out = []
    for l in range(H.layer):
        ijk = level_planes(H.x[l], dx)
        stride = int(round(H.spacing[l] / dx))
        span = overlap if l == 0 else stride
        gz = ijk[:, 2]

        def rows(lo, hi):
            ids = np.nonzero((gz >= lo) & (gz < hi))[0]
            order = np.lexsort((ijk[ids, 0], ijk[ids, 1], gz[ids]))
            return torch.from_numpy(ids[order]).long()

        pv = HaloPlan()
        if rank > 0:
            pv.recv[rank - 1] = rows(z0 - span, z0)
            pv.send[rank - 1] = rows(z0, z0 + 1)                         # the lower neighbour's ghost plane above
        if rank < world - 1:
            pv.recv[rank + 1] = rows(z1, z1 + 1)
            pv.send[rank + 1] = rows(z1 - span, z1)                      # the upper neighbour's ghost planes below
        v2i = torch.from_numpy(np.asarray(H.vertex2index[l])).long()
        pi = HaloPlan()
        pi.send = {k: v2i[v] for k, v in pv.send.items()}
        pi.recv = {k: v2i[v] for k, v in pv.recv.items()}
        top = nz_total if rank == world - 1 else -1
        owned = (gz >= z0) & ((gz < z1) | (gz == top))
        out.append((pv, pi, owned))
    return out


_KEY_BASE = np.int64(1) << 21


def lattice_keys(ijk: np.ndarray) -> np.ndarray:
    """One int64 per lattice point, ascending in (gz, iy, ix): the order both sides of an exchange list their rows in."""
    return (ijk[:, 2] * _KEY_BASE + ijk[:, 1]) * _KEY_BASE + ijk[:, 0]


def level_recv_keys(H, dx, z0, z1, rank, world, overlap):
    """Per level: {neighbour rank: ascending lattice keys of the ghost rows this rank reads from it} -- the same planes as
    `build_level_plans`, restricted to the vertices the local mesh really has (a voxel SET, not a box: `SlabObject(mask=)`)."""
    out = []
    for l in range(H.layer):
        ijk = level_planes(H.x[l], dx)
        stride = int(round(H.spacing[l] / dx))
        span = overlap if l == 0 else stride
        gz, key = ijk[:, 2], lattice_keys(ijk)
        d = {}
        if rank > 0:
            d[rank - 1] = np.sort(key[(gz >= z0 - span) & (gz < z0)])
        if rank < world - 1:
            d[rank + 1] = np.sort(key[gz == z1])
        out.append(d)
    return out


def build_level_plans_matched(H, dx, z0, z1, cuts, rank, world, recv_keys_all):
    """`build_level_plans` for a voxel set: `recv_keys_all[r]` = `level_recv_keys` of rank r (gathered once at setup); a
    rank sends exactly the rows its neighbour asks for, in the neighbour's (ascending key) order.  On a box the result is
    `build_level_plans`' (tests/test_multigpu.py)."""
    nz_total = cuts[-1]
    out = []
    for l in range(H.layer):
        ijk = level_planes(H.x[l], dx)
        gz, key = ijk[:, 2], lattice_keys(ijk)
        order = np.argsort(key, kind="stable")
        skey = key[order]
        top = nz_total if rank == world - 1 else -1
        owned = (gz >= z0) & ((gz < z1) | (gz == top))

        def rows_of(keys, what):
            pos = np.minimum(np.searchsorted(skey, keys), len(skey) - 1)
            if len(keys) and not np.array_equal(skey[pos], keys):
                raise ValueError(f"level {l}: rank {rank} does not hold every lattice point of {what}")
            return order[pos]

        pv = HaloPlan()
        for nb in [r for r in (rank - 1, rank + 1) if 0 <= r < world]:
            recv = rows_of(recv_keys_all[rank][l][nb], f"its own ghost list for rank {nb}")
            send = rows_of(recv_keys_all[nb][l][rank], f"the ghost list of rank {nb}")
            if owned[recv].any() or not owned[send].all():
                raise ValueError(f"level {l}: ghost lists of ranks {rank} and {nb} disagree with the ownership rule")
            pv.recv[nb], pv.send[nb] = torch.from_numpy(recv).long(), torch.from_numpy(send).long()
        v2i = torch.from_numpy(np.asarray(H.vertex2index[l])).long()
        pi = HaloPlan()
        pi.send = {k: v2i[v] for k, v in pv.send.items()}
        pi.recv = {k: v2i[v] for k, v in pv.recv.items()}
        out.append((pv, pi, owned))
    return out


def mask_slab(mask: np.ndarray, dx: float, clo: int, chi: int):
    """Local mesh of the cells [clo, chi) of a global voxel set, the local ids of the vertices the reference would pin on
    the GLOBAL mesh (object.py:762-773: y >= y_max - dx - 1e-5, float32) and the global vertex id of every local vertex."""
    mask = np.asarray(mask, dtype=bool)
    nx, ny, nz = mask.shape
    pts, hx = mesh.mesh_from_mask(mask[:, :, clo:chi], dx, index_offset=(0, 0, clo))
    iy_top = int(np.nonzero(mask.any(axis=(0, 2)))[0].max()) + 1               # highest used vertex row
    y_max = np.float32(iy_top) * np.float32(dx)
    thr = np.float32(np.float32(y_max - np.float32(dx)) - np.float32(1e-5))
    pins = np.nonzero(pts[:, 1] >= thr)[0].astype(np.int32)
    px, py = nx + 1, ny + 1
    mark = np.zeros(px * py * (nz + 1), dtype=bool)                            # used lattice points of the global mesh
    cx, cy, cz = np.nonzero(mask)
    for a in range(8):
        mark[(cx + ((a >> 2) & 1)) + px * ((cy + ((a >> 1) & 1)) + py * (cz + (a & 1)))] = True
    gid_of_lattice = np.cumsum(mark) - 1
    ijk = level_planes(pts, dx)
    gid = gid_of_lattice[ijk[:, 0] + px * (ijk[:, 1] + py * ijk[:, 2])]
    return pts, hx, pins, gid.astype(np.int64)


class SlabObject:
    """One rank of the mesh-partitioned FASv3 / Newton-multigrid step on an nx x ny x nz_total voxel box (dx spacing), or on
    the voxel set `mask` [nx, ny, nz_total] (bool) inside that box (every slab must keep at least one voxel)."""

    def __init__(self, nx, ny, nz_total, dx, *, rank=None, world=None, device=None, group=None, graph_segments=True,
                 layers=2, comm="auto", mask=None, **object_kwargs):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.group = group
        self.nx, self.ny, self.nz_total, self.dx = nx, ny, nz_total, dx
        self.layers = int(layers)
        self.align = hierarchy_align(self.layers, bool(object_kwargs.get("corrected_hierarchy", False)))
        self.overlap = self.align
        self.cuts = slab_cuts(nz_total, self.world, self.align)
        z0, z1, clo, chi = slab_planes(nz_total, self.world, self.rank, self.align, self.overlap)
        self.z0, self.z1, self.clo, self.chi = z0, z1, clo, chi
        object_kwargs.setdefault("use_graphs", False)
        object_kwargs.setdefault("lattice_origin", (0.0, 0.0, 0.0))     # colours / coarse cells counted from the GLOBAL corner
        if mask is None:
            pts, hx = mesh.box_mesh(nx, ny, chi - clo, dx, index_offset=(0, 0, clo))
            self.global_ids = None
            self.obj = Object((pts, hx), dx, [], device=device, layers=self.layers, **object_kwargs)
        else:
            if tuple(np.shape(mask)) != (nx, ny, nz_total):
                raise ValueError(f"mask must have shape {(nx, ny, nz_total)}, got {tuple(np.shape(mask))}")
            if not np.asarray(mask, dtype=bool)[:, :, clo:chi].any():
                raise ValueError(f"rank {self.rank}: no voxel in the cell layers [{clo}, {chi})")
            pts, hx, pins, gid = mask_slab(mask, dx, clo, chi)
            self.global_ids = gid
            self.obj = Object((pts, hx), dx, pins.tolist(), device=device, layers=self.layers, pins_exact=True, **object_kwargs)
        o = self.obj
        H = o.hierarchy
        if mask is None:
            plans = build_level_plans(H, dx, z0, z1, self.cuts, self.rank, self.world, self.overlap)
        else:
            mine = level_recv_keys(H, dx, z0, z1, self.rank, self.world, self.overlap)
            everyone = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(everyone, mine, group=group)
            else:
                everyone[0] = mine
            plans = build_level_plans_matched(H, dx, z0, z1, self.cuts, self.rank, self.world, everyone)
        self.plan_v = [p[0].to(o.device) for p in plans]
        self.plan_i = [p[1].to(o.device) for p in plans]
        self.owned_level = [torch.from_numpy(p[2]).to(o.device) for p in plans]
        self.owned = self.owned_level[0]
        self.n_exchanges = 0
        self.comm = None
        if comm != "torch" and o.device.type == "cuda" and self.world > 1:
            from . import comm as comm_mod
            self.comm = comm_mod.HaloComm(o, self.rank, self.world, group)
            self.hx_v = [self.comm.plan(p) for p in self.plan_v]
            self.hx_i = [self.comm.plan(p) for p in self.plan_i]
            if comm != "nccl":
                # peer-memory transport: every neighbour of any plan, slots sized for the largest segment (collective call)
                nbs = sorted({nb for p in self.plan_v + self.plan_i for nb in p.neighbours()})
                seg = max([int(t.numel()) for p in self.plan_v + self.plan_i for t in list(p.send.values()) + list(p.recv.values())] + [1])
                seg_t = torch.tensor([seg], device=o.device)
                dist.all_reduce(seg_t, op=dist.ReduceOp.MAX, group=group)
                ok = self.comm.enable_p2p(nbs, int(seg_t.item()) * 3 * 4)
                if comm == "p2p" and not ok:
                    raise _lib.FembError("peer-memory transport requested but not available: " + o.lib.femb_last_error().decode())
        # coarse rest state of the re-discretising FAS step: ghost coarse vertices average over children this rank does
        # not have (object.py:730-731)
        if self.layers >= 2:
            self._exchange(o.x_gpu_layer[1], 1, "v")
            o._prepare_rest(1)
        ps = (nx + 1) * (ny + 1)
        self.global_offset = ps * clo                  # local vertex id + offset = vertex id in the full box mesh
        self.graph_segments, self._graphs = graph_segments, {}
        torch.cuda.synchronize(o.device) if o.device.type == "cuda" else None

    # ---- plumbing ------------------------------------------------------------------------------------------------
    def _exchange(self, vec, level, order):
        """Ghost rows of `vec` (a level-`level` vector in vertex ("v") or index ("i") order) <- their owners' values."""
        self.n_exchanges += 1
        if self.comm is not None:
            self.comm.exchange((self.hx_v if order == "v" else self.hx_i)[level], vec)
        else:
            exchange(vec, (self.plan_v if order == "v" else self.plan_i)[level], self.group)

    def _seg(self, key, fn):
        """Run one exchange-free segment of the step.  With `graph_segments` the segment is captured into a CUDA graph the
        first time it runs (after executing it eagerly once, like Object._run_step) and replayed afterwards."""
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

    def _run(self, key, body):
        """A whole step.  With the C-ABI communicator the exchanges are stream-ordered NCCL calls and the step is ONE CUDA
        graph per rank (`graph_segments`); with the torch.distributed exchange only the exchange-free segments are."""
        if self.comm is None or not self.graph_segments:
            body()
            return
        g = self._graphs.get(key)
        if g is None:
            self._whole = True
            body()                                      # eager (a real step: it is the one the caller asked for)
            torch.cuda.synchronize(self.obj.device)
            g = self.comm.capture(body)
            self._graphs[key] = g
            self._whole = False
            return
        g.replay()

    def _colours(self, l, first, last, parts):
        o = self.obj
        _lib.check(o.lib.femb_gs_colors(C.byref(o.levels[l]), o._values(l).data_ptr(), o.dev_B_fixed[l].data_ptr(),
                                        o.dev_delta_x[l].data_ptr(), first, last, parts, o.not_spd.data_ptr(), o._stream()),
                   "femb_gs_colors")

    def _seg_or_inline(self, key, fn):
        if self.comm is not None and self.graph_segments:
            fn()                                        # inside the whole-step capture (or its eager first run)
        else:
            self._seg(key, fn)

    def gauss_seidel(self, l, iterations, tag=""):
        """PerformGaussSeidel(l, iterations) (object.py:886-929) in colour groups with a ghost exchange after each."""
        o = self.obj
        U, LU = _lib.PART_U, _lib.PART_L | _lib.PART_U
        dxl = o.dev_delta_x[l]
        for it in range(iterations):
            back = U if it == 0 else LU                 # first half-sweep: dx == 0, only higher colours contribute
            k = f"gs{tag}_{l}_{min(it, 1)}"

            def first_group(it=it, back=back):
                if it == 0:
                    dxl.zero_()
                self._colours(l, 7, 4, back)
            self._seg_or_inline(k + "a", first_group)
            self._exchange(dxl, l, "i")
            self._seg_or_inline(k + "b", lambda back=back: self._colours(l, 3, 0, back))
            self._exchange(dxl, l, "i")
            self._seg_or_inline(f"gs_{l}_c", lambda: self._colours(l, 0, 3, LU))
            self._exchange(dxl, l, "i")
            self._seg_or_inline(f"gs_{l}_d", lambda: self._colours(l, 4, 7, LU))
            self._exchange(dxl, l, "i")

    # ---- the FASv3 loop body of object.py:1467-1511 with the halo exchanges in place -------------------------------
    def _fasv3_body(self):
        o = self.obj
        x = o.x_gpu_layer[0]
        dx0 = o.dev_delta_x[0]

        def seg_restrict():
            o.update_deltaX(x, dx0)                                                         # :1487
            o._transfer(o.Ut_noOrder[0], 1.0, x, 0.0, o.dev_x_solved[1])                    # :1489

        def seg_coarse():
            o._force(x, o.dev_R[0])                                                         # :1491-1495
            o._transfer(o.Ut_hat[0], 1.0, o.dev_R[0], 0.0, o.dev_B_fixed[1])                # :1497
            o._hessian(1, o.dev_x_solved[1], o.UtAUs_values[0])                             # :1499
            _lib.check(o.lib.femb_coarse_pin_penalty(C.byref(o.levels[1]), C.byref(o.Ut_noOrder_hat[0].c),
                                                     o.pin_gpu.data_ptr(), o.control_mag, o.UtAUs_values[0].data_ptr(), o._stream()),
                       "femb_coarse_pin_penalty")
            o.PerformJacobi(layer=1, iterations=1)                                          # :1506

        self._seg_or_inline("fine", lambda: o._fas_fine(x, False, True, True))              # :1468-1481
        self.gauss_seidel(0, 1, tag="f")                                                    # :1483
        self._seg_or_inline("restrict", seg_restrict)
        self._exchange(o.dev_x_solved[1], 1, "v")
        self._seg_or_inline("coarse", seg_coarse)
        self._exchange(o.dev_delta_x[1], 1, "i")
        self._seg_or_inline("prolong", lambda: o._transfer(o.Us_hat[0], 1.0, o.dev_delta_x[1], 0.0, dx0))   # :1508
        self._exchange(dx0, 0, "i")                     # owner and ghost apply bit-identical increments
        self._seg_or_inline("update", lambda: o.update_deltaX(x, dx0))                      # :1509

    def fasv3_step(self):
        self._run("fasv3", self._fasv3_body)

    def FASv3(self, iterations=100):
        for _ in range(iterations):
            self.fasv3_step()

    # ---- the Newton-multigrid loop body of object.py:1231-1296 --------------------------------------------------------
    def _vcycle(self, l):
        """VCycle(l) (object.py:1216-1226) with downSample / upSample (:1021-1035) and their exchanges."""
        o = self.obj
        self.gauss_seidel(l, 3)
        if l < o.layer - 1:
            def down():
                o._copy(o.dev_x_solved[l], o.dev_delta_x[l])
                o._residual(l, o.dev_B_fixed[l], o.dev_delta_x[l], o.dev_R[l])
            self._seg_or_inline(f"down_{l}", down)
            self._exchange(o.dev_R[l], l, "i")          # owned coarse rows restrict the residual of one ghost plane
            self._seg_or_inline(f"restr_{l}", lambda: o._transfer(o.Ut_hat[l], 1.0, o.dev_R[l], 0.0, o.dev_B_fixed[l + 1]))
            self._vcycle(l + 1)

            def up():
                o._axpby(1.0, o.dev_delta_x[l + 1], 1.0, o.dev_x_solved[l + 1])
                o._transfer(o.Us_hat[l], 1.0, o.dev_x_solved[l + 1], 1.0, o.dev_x_solved[l])
            self._seg_or_inline(f"up_{l}", up)
            self._exchange(o.dev_x_solved[l], l, "i")   # ghost rows were prolonged from parents this rank does not own
            self._seg_or_inline(f"rhs_{l}", lambda: o._spmv(l, _lib.PART_ALL, -1.0, o.dev_x_solved[l], 1.0, o.dev_B_fixed[l]))
        if l == o.layer - 1:
            return
        self.gauss_seidel(l, 3)

    def _nmg_body(self):
        o = self.obj
        x = o.x_gpu_layer[0]

        def assemble():
            o._force_hessian(x, o.grad_gpu, o.MF_value_gpu, o.diag_fixed)                   # :1232-1236, 1249
            if o.spd_level0:
                o._spd(o.MF_value_gpu)
            o.galerkin()                                                                    # :1252-1264
            o._copy(o.dev_B_fixed[0], o.grad_gpu)                                           # :1266
        self._seg_or_inline("assemble", assemble)
        self._vcycle(0)                                                                     # :1290
        self._seg_or_inline("finish", o.finish)                                             # :1293
        self._exchange(o.dev_x_solved[0], 0, "i")
        self._seg_or_inline("newton_update", lambda: o.update_deltaX(x, o.dev_x_solved[0]))  # :1296

    def nmg_step(self):
        self._run("nmg", self._nmg_body)

    def NewtonMultigrid(self, iterations=100):
        for _ in range(iterations):
            self.nmg_step()

    # ---- results ----------------------------------------------------------------------------------------------------
    def owned_positions(self):
        """(global vertex ids, positions) of the owned vertices, for gathering / checking."""
        ids = torch.nonzero(self.owned).squeeze(1)
        if self.global_ids is not None:
            return torch.from_numpy(self.global_ids).to(ids.device)[ids], self.obj.x_gpu_layer[0][ids]
        return ids + self.global_offset, self.obj.x_gpu_layer[0][ids]

    def force_inf_norm(self):
        """max |f| over the owned rows of all ranks (the sampled quantity of object.py:1245)."""
        o = self.obj
        o._force(o.x_gpu_layer[0], o.grad_gpu, ordered=False)
        m = (o.grad_gpu.abs().amax(dim=1) * self.owned).max().reshape(1)
        if self.world > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=self.group)
        return float(m.item())

    def checksum(self):
        """(sum |x| over owned vertices, all ranks) -- a size-independent fingerprint of the state for bench.py."""
        s = (self.obj.x_gpu_layer[0].abs().sum(dim=1).double() * self.owned).sum().reshape(1)
        if self.world > 1:
            dist.all_reduce(s, op=dist.ReduceOp.SUM, group=self.group)
        return float(s.item())
