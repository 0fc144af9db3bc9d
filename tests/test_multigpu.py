"""Mesh-partitioned path (fem_simulation_b200.distributed).

CPU (gloo, world_size 2): slab bookkeeping and halo plans -- every global vertex has exactly one owner, the send /
recv lists of neighbouring ranks match entry by entry (fine and coarse level, vertex and index ordering), the local
colouring / coarse cells coincide with the global ones.
GPU (``-m gpu``, needs >= 2 GPUs, NCCL): the distributed FASv3 step reproduces the single-GPU step up to summation-order
round-off (the local coarse numbering orders the <= 8 parents of a fine vertex differently in the prolongation sum)."""
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


def _cpu_worker(rank, world, port, out, nz):
    _init(rank, world, port, "gloo")
    try:
        z0, z1, clo, chi = D.slab_planes(nz, world, rank)
        pts, hx = mesh.box_mesh(NX, NY, chi - clo, DX, index_offset=(0, 0, clo))
        H = hierarchy.build_hierarchy(pts, hx, DX, layers=2)
        gpts, ghx = mesh.box_mesh(NX, NY, nz, DX)
        G = hierarchy.build_hierarchy(gpts, ghx, DX, layers=2)
        ps = (NX + 1) * (NY + 1)
        off = ps * clo
        # local slab == slice of the global mesh: coordinates bit-identical, colours identical
        assert np.array_equal(pts, gpts[off:off + len(pts)])
        col = lambda Hh: np.searchsorted(np.asarray(Hh.color_vertex_num[0]), Hh.vertex2index[0], side="right") - 1
        assert np.array_equal(col(H), col(G)[off:off + len(pts)])
        fv, fi, cv, ci = D.build_halo_plans(NX, NY, z0, z1, clo, chi, nz, rank, world, H.x[1], H.vertex2index[1],
                                            H.vertex2index[0], DX, torch.device("cpu"))
        # fine level, vertex order: payload = global vertex id
        gid = torch.arange(len(pts), dtype=torch.float32) + off
        gz = np.arange(len(pts)) // ps + clo
        owned = (gz >= z0) & ((gz < z1) | ((gz == nz) & (rank == world - 1)))
        v = torch.stack([gid, gid * 2, gid * 3], dim=1)
        v[torch.from_numpy(~owned)] = -1.0
        D.exchange(v, fv)
        assert torch.equal(v[:, 0], gid), "fine ghosts (vertex order)"
        # fine level, index order
        v2i = torch.from_numpy(H.vertex2index[0]).long()
        w = torch.full((len(pts), 3), -1.0)
        w[v2i[torch.from_numpy(owned)]] = torch.stack([gid, gid, gid], dim=1)[torch.from_numpy(owned)]
        D.exchange(w, fi)
        assert torch.equal(w[v2i, 0], gid), "fine ghosts (index order)"
        # coarse level: payload = global coarse lattice key
        cx = H.x[1]
        key = np.rint(cx / np.float32(2 * DX)).astype(np.int64)
        gkey = torch.from_numpy((key[:, 0] + 100 * key[:, 1] + 10000 * key[:, 2]).astype(np.float32))
        cgz = np.rint(cx[:, 2] / np.float32(DX)).astype(np.int64)
        cown = (cgz >= z0) & ((cgz < z1) | ((cgz == nz) & (rank == world - 1)))
        c = torch.stack([gkey, gkey, gkey], dim=1).clone()
        c[torch.from_numpy(~cown)] = -1.0
        D.exchange(c, cv)
        assert torch.equal(c[:, 0], gkey), "coarse ghosts (vertex order)"
        cv2i = torch.from_numpy(H.vertex2index[1]).long()
        ci_vec = torch.full((len(cx), 3), -1.0)
        ci_vec[cv2i[torch.from_numpy(cown)]] = torch.stack([gkey, gkey, gkey], dim=1)[torch.from_numpy(cown)]
        D.exchange(ci_vec, ci)
        assert torch.equal(ci_vec[cv2i, 0], gkey), "coarse ghosts (index order)"
        # ownership is a partition of the global vertex set
        cnt = torch.zeros(len(gpts))
        cnt[off:off + len(pts)][torch.from_numpy(owned)] = 1
        dist.all_reduce(cnt)
        assert torch.all(cnt == 1)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nz", [(2, NZ), (4, NZ), (8, 2 * NZ)])
def test_halo_plans_gloo(world, nz):
    """world 4 and 8 (2 cell layers per rank) add interior ranks, which talk to both neighbours; 8 is the size of the
    round-end scaling run."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cpu_worker, args=(r, world, port, out, nz)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    res = sorted(out.get(timeout=5) for _ in range(world))
    assert res == [(r, "ok") for r in range(world)], res


def test_slab_planes_validation():
    assert D.slab_planes(8, 2, 0) == (0, 4, 0, 4) and D.slab_planes(8, 2, 1) == (4, 8, 2, 8)
    with pytest.raises(ValueError):
        D.slab_planes(9, 2, 0)
    with pytest.raises(ValueError):
        D.slab_planes(12, 4, 0)            # 3 planes per rank: odd


def _gpu_worker(rank, world, port, steps, out):
    torch.cuda.set_device(rank)
    _init(rank, world, port, "nccl")
    try:
        so = D.SlabObject(12, 10, 16, DX, device=f"cuda:{rank}")      # 16 cell layers: 8 per rank at world 2, 4 at world 4
        so.FASv3(steps)
        ids, xs = so.owned_positions()
        out.put((rank, ids.cpu().numpy(), xs.cpu().numpy(), so.force_inf_norm()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_distributed_fasv3_matches_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from fem_simulation_b200 import Object
    steps = 3
    pts, hx = mesh.box_mesh(12, 10, 16, DX)
    ref = Object((pts, hx), DX, [], layers=2, use_graphs=False)
    ref.FASv3(steps)
    xref = ref.x_gpu_layer[0].cpu().numpy()
    ref._force(ref.x_gpu_layer[0], ref.grad_gpu)
    fref = float(ref.grad_gpu.abs().max())
    world = 4 if torch.cuda.device_count() >= 4 else 2               # 4 ranks exercise interior slabs (two neighbours)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gpu_worker, args=(r, world, port, steps, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(60)
    x = np.full_like(xref, np.nan)
    for rank, ids, xs, fn in res:
        x[ids] = xs
        assert abs(fn - fref) <= 1e-6 * fref
    assert not np.isnan(x).any()
    assert np.abs(x - xref).max() <= 1e-6 * np.abs(xref).max(), f"max diff {np.abs(x - xref).max():.3e}"
    disp = xref - pts
    assert np.abs((x - pts) - disp).max() <= 1e-4 * np.abs(disp).max()
