"""Mesh-partitioned path (fem_simulation_b200.distributed).

CPU (gloo, world_size 2): slab bookkeeping and halo plans -- every global vertex has exactly one owner, the send /
recv lists of neighbouring ranks match entry by entry (fine and coarse level, vertex and index ordering), the local
colouring / coarse cells coincide with the global ones.
GPU (``-m gpu``, needs >= 2 GPUs, NCCL): the distributed FASv3 and Newton-multigrid steps reproduce the single-GPU steps
up to summation-order round-off (the local coarse numbering orders the <= 8 parents of a fine vertex differently in the
prolongation sum)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fem_simulation_b200 import distributed as D
from fem_simulation_b200 import hierarchy, mesh

NX, NY, NZ, DX = 6, 5, 8, 0.05


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _init(rank, world, port, backend):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)


def _voxel_set(nz, seed):
    """a reproducible ragged voxel set in the NX x NY x nz box: every cell layer keeps voxels, the box corners are kept"""
    rng = np.random.default_rng(seed)
    mask = rng.random((NX, NY, nz)) < 0.7
    mask[0, 0, :] = True
    mask[-1, -1, -1] = mask[0, -1, 0] = mask[-1, 0, 0] = True
    return mask


def _cpu_worker(rank, world, port, out, nz, layers, mask_seed=None):
    _init(rank, world, port, "gloo")
    try:
        align = D.hierarchy_align(layers)
        cuts = D.slab_cuts(nz, world, align)
        z0, z1, clo, chi = D.slab_planes(nz, world, rank, align)
        col = lambda Hh, l: np.searchsorted(np.asarray(Hh.color_vertex_num[l]), Hh.vertex2index[l], side="right") - 1
        if mask_seed is None:
            pts, hx = mesh.box_mesh(NX, NY, chi - clo, DX, index_offset=(0, 0, clo))
            gpts, ghx = mesh.box_mesh(NX, NY, nz, DX)
            gid0 = (NX + 1) * (NY + 1) * clo + np.arange(len(pts))
        else:
            mask = _voxel_set(nz, mask_seed)
            pts, hx, pins, gid0 = D.mask_slab(mask, DX, clo, chi)
            gpts, ghx = mesh.mesh_from_mask(mask, DX)
            # the slab pins what the reference pins on the whole mesh (object.py:762-773), whatever the slab's own top is
            gpins = hierarchy.default_pins(gpts, DX)
            assert np.array_equal(np.sort(gid0[pins]), np.intersect1d(gpins, gid0))
        H = hierarchy.build_hierarchy(pts, hx, DX, layers=layers, origin=(0.0, 0.0, 0.0))
        G = hierarchy.build_hierarchy(gpts, ghx, DX, layers=layers, origin=(0.0, 0.0, 0.0))
        # local slab == part of the global mesh: coordinates bit-identical, colours identical
        assert np.array_equal(pts, gpts[gid0])
        assert np.array_equal(col(H, 0), col(G, 0)[gid0])
        mine = D.level_recv_keys(H, DX, z0, z1, rank, world, align)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        plans = D.build_level_plans_matched(H, DX, z0, z1, cuts, rank, world, everyone)
        if mask_seed is None:                                                # on a box: the plane rule's lists, entry by entry
            for (pv, pi, owned), (qv, qi, qowned) in zip(plans, D.build_level_plans(H, DX, z0, z1, cuts, rank, world, align)):
                assert np.array_equal(owned, qowned)
                for a, b in [(pv, qv), (pi, qi)]:
                    assert sorted(a.send) == sorted(b.send) and sorted(a.recv) == sorted(b.recv)
                    assert all(torch.equal(a.send[k], b.send[k]) for k in a.send) and all(torch.equal(a.recv[k], b.recv[k]) for k in a.recv)
        n_owned_levels = []
        for l in range(H.layer):
            pv, pi, owned = plans[l]
            ijk = D.level_planes(H.x[l], DX)
            stride = int(round(H.spacing[l] / DX))
            # the local level is the matching slice of the global level: same lattice points, same colours
            gijk = D.level_planes(G.x[l], DX)
            gkey = {tuple(k): i for i, k in enumerate(gijk.tolist())}
            gid_np = np.array([gkey[tuple(k)] for k in ijk.tolist()])
            assert np.array_equal(col(H, l), col(G, l)[gid_np]), f"level {l} colours"
            gid = torch.from_numpy(gid_np.astype(np.float32))
            # vertex order: payload = global id of the lattice point at this level; every ghost row a local owned row reads
            # (one level plane below the slab -- all overlap planes on level 0 -- and the plane above) must end up current
            v = torch.stack([gid, gid * 2, gid * 3], dim=1)
            v[torch.from_numpy(~owned)] = -1.0
            D.exchange(v, pv)
            span = align if l == 0 else stride
            need = ((ijk[:, 2] >= z0 - span) & (ijk[:, 2] <= z1)) if rank < world - 1 else (ijk[:, 2] >= z0 - span)
            need_t = torch.from_numpy(need)
            assert torch.equal(v[need_t, 0], gid[need_t]), f"level {l} ghosts (vertex order)"
            v2i = torch.from_numpy(H.vertex2index[l]).long()
            w = torch.full((len(gid), 3), -1.0)
            own_t = torch.from_numpy(owned)
            w[v2i[own_t]] = torch.stack([gid, gid, gid], dim=1)[own_t]
            D.exchange(w, pi)
            assert torch.equal(w[v2i[need_t], 0], gid[need_t]), f"level {l} ghosts (index order)"
            # every owned row has its complete set of blocks locally (assembly and the Galerkin rows need no reduction): the
            # local pattern restricted to the owned rows is the global pattern's, lattice point by lattice point
            lp, gp = H.pattern[l], G.pattern[l]
            li2v, gi2v = np.asarray(H.index2vertex[l]), np.asarray(G.index2vertex[l])
            lrow, lcol = gid_np[li2v[lp.row]], gid_np[li2v[lp.col]]
            keep = owned[li2v[lp.row]]
            local_pairs = np.unique(lrow[keep] * len(gijk) + lcol[keep])
            grow, gcol = gi2v[gp.row].astype(np.int64), gi2v[gp.col].astype(np.int64)
            mine_g = np.zeros(len(gijk), bool); mine_g[gid_np[owned]] = True
            global_pairs = np.unique(grow[mine_g[grow]] * len(gijk) + gcol[mine_g[grow]])
            assert np.array_equal(local_pairs, global_pairs), f"level {l}: owned rows' blocks"
            # ownership is a partition of the global level
            cnt = torch.zeros(len(gijk))
            cnt[torch.from_numpy(gid_np[owned])] = 1
            dist.all_reduce(cnt)
            assert torch.all(cnt == 1), f"level {l} ownership"
            n_owned_levels.append(int(owned.sum()))
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nz,layers,mask_seed", [(2, NZ, 2, None), (4, NZ, 2, None), (8, 2 * NZ, 2, None), (2, 16, 3, None),
                                                       (3, 28, 4, None), (2, NZ, 2, 11), (4, 16, 2, 12), (3, 28, 4, 13)])
def test_halo_plans_gloo(world, nz, layers, mask_seed):
    """world 4 and 8 (2 cell layers per rank) add interior ranks, which talk to both neighbours; 8 is the size of the
    round-end scaling run; layers 3 / 4 are the Newton-multigrid decomposition (cuts on multiples of 4 planes, overlap 4,
    per-level ghost planes; 28 planes over 3 ranks: uneven slabs 12 / 8 / 8).  With a mask seed the mesh is a ragged voxel
    set instead of the box (`SlabObject(mask=)`): the ghost lists are matched through the gathered lattice keys."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, world, port, out, nz, layers, mask_seed)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    res = sorted(out.get(timeout=5) for _ in range(world))
    assert res == [(r, "ok") for r in range(world)], res


def test_slab_planes_validation():
    assert D.slab_planes(8, 2, 0) == (0, 4, 0, 4) and D.slab_planes(8, 2, 1) == (4, 8, 2, 8)
    with pytest.raises(ValueError):
        D.slab_planes(9, 2, 0)             # not a multiple of the cut alignment
    assert D.slab_cuts(12, 4, 2) == [0, 4, 8, 10, 12]                # uneven slabs are fine
    # the metric's 10 M cube on 8 ranks: cuts on multiples of 4 planes, six slabs of 28 and two of 24
    cuts = D.slab_cuts(216, 8, D.hierarchy_align(6))
    assert cuts[0] == 0 and cuts[-1] == 216 and all(c % 4 == 0 for c in cuts)
    assert sorted(set(np.diff(cuts).tolist())) == [24, 28]
    assert D.slab_planes(216, 8, 3, 4) == (84, 112, 80, 112)
    with pytest.raises(ValueError):
        D.slab_cuts(8, 4, 4)               # fewer aligned groups than ranks


def _gpu_voxel_set():
    rng = np.random.default_rng(21)
    mask = rng.random((12, 10, 32)) < 0.75
    mask[0, 0, :] = mask[-1, -1, :] = True                           # keeps the bounding box and the top row in every slab
    return mask


def _gpu_worker(rank, world, port, steps, out, method, layers, comm, graphs, masked=False):
    torch.cuda.set_device(rank)
    _init(rank, world, port, "nccl")
    try:
        so = D.SlabObject(12, 10, 32, DX, device=f"cuda:{rank}", layers=layers, comm=comm, graph_segments=graphs,
                          mask=_gpu_voxel_set() if masked else None)
        for _ in range(steps):
            so.fasv3_step() if method == "fasv3" else so.nmg_step()
        ids, xs = so.owned_positions()
        if so.comm is not None:
            so.comm.check()                                          # no pull kernel timed out
            if comm == "p2p":
                assert so.comm.transport.startswith("peer memory")
        out.put((rank, ids.cpu().numpy(), xs.cpu().numpy(), so.force_inf_norm()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("method,layers,comm,graphs,masked", [
    ("fasv3", 2, "p2p", True, False), ("nmg", 3, "p2p", True, False), ("nmg", 3, "p2p", False, False), ("nmg", 3, "nccl", True, False),
    ("nmg", 4, "torch", False, False), ("fasv3", 2, "torch", True, False),
    # a ragged voxel set instead of the box: written after the round's GPU budget was spent, never run on GPUs yet -- the
    # outcome is recorded (XPASS / xfail), not asserted, until it has been seen green once
    pytest.param("fasv3", 2, "nccl", True, True, marks=pytest.mark.xfail(strict=False, reason="voxel-set slabs: not yet run on GPUs")),
    pytest.param("nmg", 3, "nccl", True, True, marks=pytest.mark.xfail(strict=False, reason="voxel-set slabs: not yet run on GPUs"))])
def test_distributed_step_matches_single_gpu(method, layers, comm, graphs, masked):
    """The mesh-partitioned FASv3 / Newton-multigrid step (C-ABI exchanges inside one CUDA graph per rank -- peer-memory
    push / pull kernels or NCCL send / recv --, or the torch.distributed exchange between segment graphs) vs the single-GPU
    step on the same global mesh."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from fem_simulation_b200 import Object
    steps = 4
    pts, hx = mesh.mesh_from_mask(_gpu_voxel_set(), DX) if masked else mesh.box_mesh(12, 10, 32, DX)
    ref = Object((pts, hx), DX, [], layers=layers, use_graphs=False, lattice_origin=(0.0, 0.0, 0.0) if masked else None)
    ref.FASv3(steps) if method == "fasv3" else ref.NewtonMultigrid(steps)
    xref = ref.x_gpu_layer[0].cpu().numpy()
    ref._force(ref.x_gpu_layer[0], ref.grad_gpu)
    fref = float(ref.grad_gpu.abs().max())
    world = 4 if torch.cuda.device_count() >= 4 else 2               # 4 ranks exercise interior slabs (two neighbours)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, steps, out, method, layers, comm, graphs, masked)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time
    res, t_end = [], time.time() + 240
    while len(res) < world:                                          # a crashed rank must not stall the test for minutes
        try:
            res.append(out.get(timeout=2))
        except queue.Empty:
            if time.time() > t_end or any(p.exitcode not in (None, 0) for p in procs):
                for p in procs:
                    p.kill()
                pytest.fail(f"ranks exited {[p.exitcode for p in procs]} with {len(res)} of {world} results")
    for p in procs:
        p.join(60)
    x = np.full_like(xref, np.nan)
    for rank, ids, xs, fn in res:
        x[ids] = xs
        assert abs(fn - fref) <= 1e-5 * fref
    assert not np.isnan(x).any()
    assert np.abs(x - xref).max() <= 1e-6 * np.abs(xref).max(), f"max diff {np.abs(x - xref).max():.3e}"
    disp = xref - pts
    assert np.abs((x - pts) - disp).max() <= 1e-4 * np.abs(disp).max()


class _HostObject:
    """Stand-in for `Object` on a machine without a GPU: the hierarchy and the level-1 positions `SlabObject.__init__` touches
    (the CUDA state is not needed to check the constructor's bookkeeping)."""

    def __init__(self, mesh_path, dx, pinList, *, device=None, layers=None, lattice_origin=None, pins_exact=False, **_):
        pts, hx = mesh_path
        self.hierarchy = hierarchy.build_hierarchy(pts, hx, dx, layers, origin=lattice_origin)
        self.device = torch.device("cpu")
        self.pinList, self.pins_exact = list(pinList), pins_exact
        self.x_gpu_layer = [torch.from_numpy(x.copy()) for x in self.hierarchy.x]
        self.rest_prepared = []

    def _prepare_rest(self, level):
        self.rest_prepared.append(level)


def _ctor_worker(rank, world, port, out, masked):
    _init(rank, world, port, "gloo")
    try:
        D.Object = _HostObject                                               # this process only
        nz = 16
        mask = _voxel_set(nz, 31) if masked else None
        so = D.SlabObject(NX, NY, nz, DX, rank=rank, world=world, layers=3, comm="torch", mask=mask)
        assert so.comm is None and so.obj.rest_prepared == [1] and so.n_exchanges == 1
        gpts, _ = mesh.mesh_from_mask(mask, DX) if masked else mesh.box_mesh(NX, NY, nz, DX)
        # the constructor's exchange made the ghost rows of the level-1 positions current (they already were: lattice points)
        assert torch.equal(so.obj.x_gpu_layer[1], torch.from_numpy(so.obj.hierarchy.x[1]))
        ids, xs = so.owned_positions()
        assert np.array_equal(xs.numpy(), gpts[ids.numpy()])                 # local -> global vertex ids
        cnt = torch.zeros(len(gpts)); cnt[ids] = 1
        dist.all_reduce(cnt)
        assert torch.all(cnt == 1)                                           # every global vertex owned exactly once
        if masked:
            assert so.obj.pins_exact and len(so.obj.pinList) > 0
            got = torch.zeros(len(gpts)); got[torch.from_numpy(so.global_ids[np.asarray(so.obj.pinList)])] = 1
            dist.all_reduce(got)
            want = torch.zeros(len(gpts)); want[torch.from_numpy(hierarchy.default_pins(gpts, DX)).long()] = 1
            assert torch.equal(got > 0, want > 0)                            # union of the slabs' pins = the reference's pins
        else:
            assert so.obj.pinList == [] and not so.obj.pins_exact and so.global_ids is None
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("masked", [False, True])
def test_slab_object_constructor_bookkeeping_gloo(masked):
    """`SlabObject.__init__` itself (cuts, local mesh, plans -- gathered lattice keys for a voxel set --, the first exchange,
    global vertex ids, pins) at world size 2 over gloo, with a host-only stand-in for the CUDA `Object`."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ctor_worker, args=(r, world, port, out, masked)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    res = sorted(out.get(timeout=5) for _ in range(world))
    assert res == [(r, "ok") for r in range(world)], res
