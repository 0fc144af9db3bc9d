"""Host side of the C-ABI communicator (include/femb200.h, csrc/femb_comm.cu): one NCCL communicator per rank created
by libfemb200 itself, halo plans as ``FembHalo`` structs, whole-step CUDA-graph capture with the exchanges inside.

torch.distributed is used once, to hand rank 0's NCCL unique id to the other ranks."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class HaloComm:
    def __init__(self, obj, rank, world, group=None):
        self.lib = obj.lib
        self.device = obj.device
        self.rank, self.world = rank, world
        idb = (C.c_ubyte * _lib.COMM_ID_BYTES)()
        if rank == 0:
            _lib.check(self.lib.femb_comm_unique_id(idb), "femb_comm_unique_id")
        t = torch.tensor(list(idb), dtype=torch.uint8, device=self.device if dist.get_backend(group) == "nccl" else "cpu")
        dist.broadcast(t, src=0, group=group)
        for i, b in enumerate(t.cpu().tolist()):
            idb[i] = b
        torch.cuda.set_device(self.device)
        self.handle = C.c_void_p()
        _lib.check(self.lib.femb_comm_init(C.byref(self.handle), idb, rank, world), "femb_comm_init")
        self._keep = []
        self.transport = "nccl"

    def enable_p2p(self, neighbours, slot_bytes):
        """Peer-memory transport (femb_comm_p2p_enable): collective; False (on every rank) if CUDA IPC / peer access is not
        available between all neighbours -- the exchanges then stay on NCCL send / recv."""
        nb = (C.c_int * max(len(neighbours), 1))(*[int(v) for v in neighbours])
        rc = self.lib.femb_comm_p2p_enable(self.handle, nb, len(neighbours), C.c_size_t(int(slot_bytes)),
                                           C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc == 0:
            self.transport = "peer memory (CUDA IPC over NVLink: one push-flag-wait-pull kernel per exchange)"
        return rc == 0

    def check(self):
        """Raise if a pull kernel ever gave up waiting for a neighbour (synchronises)."""
        err = C.c_int(0)
        _lib.check(self.lib.femb_comm_p2p_error(self.handle, C.byref(err)), "femb_comm_p2p_error")
        if err.value:
            raise _lib.FembError("halo exchange: a neighbour's data did not arrive (peer-memory transport timed out)")

    def plan(self, hp):
        """FembHalo for a distributed.HaloPlan (send / recv row lists per neighbour, already on the device)."""
        nbs = hp.neighbours()
        assert len(nbs) <= _lib.MAX_NEIGHBOURS
        h = _lib.FembHalo()
        h.n_nb = len(nbs)
        send, recv = [], []
        for k, nb in enumerate(nbs):
            s = hp.send.get(nb)
            r = hp.recv.get(nb)
            h.nb[k] = nb
            h.n_send[k] = 0 if s is None else int(s.numel())
            h.n_recv[k] = 0 if r is None else int(r.numel())
            if s is not None:
                send.append(s)
            if r is not None:
                recv.append(r)
        dev = self.device
        srows = (torch.cat(send) if send else torch.zeros(0, dtype=torch.long, device=dev)).to(torch.int32).contiguous()
        rrows = (torch.cat(recv) if recv else torch.zeros(0, dtype=torch.long, device=dev)).to(torch.int32).contiguous()
        sbuf = torch.empty(max(srows.numel(), 1) * 3, device=dev)
        rbuf = torch.empty(max(rrows.numel(), 1) * 3, device=dev)
        h.send_rows, h.recv_rows = srows.data_ptr(), rrows.data_ptr()
        h.send_buf, h.recv_buf = sbuf.data_ptr(), rbuf.data_ptr()
        self._keep.append((srows, rrows, sbuf, rbuf))
        return h

    def exchange(self, halo, vec):
        width = 1 if vec.dim() == 1 else int(vec.shape[1])
        assert width <= 3 and vec.is_contiguous() and vec.dtype == torch.float32
        _lib.check(self.lib.femb_halo_exchange(self.handle, C.byref(halo), vec.data_ptr(), width,
                                               C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "femb_halo_exchange")

    def allreduce(self, buf, op="sum"):
        _lib.check(self.lib.femb_comm_allreduce(self.handle, buf.data_ptr(), buf.numel(), 0 if op == "sum" else 1,
                                                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "femb_comm_allreduce")

    def capture(self, body):
        """One CUDA graph of a whole step, NCCL exchanges included (the communicator's connections exist: the caller has
        run `body` eagerly once)."""
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            body()
        return g

    def close(self):
        """Collective: every rank closes its communicator at the same point (a barrier first: with the peer-memory transport
        no rank may still be storing into another's arena)."""
        if self.handle and dist.is_available() and dist.is_initialized():
            torch.cuda.synchronize(self.device)
            dist.barrier()
        if self.handle:
            self.lib.femb_comm_destroy(self.handle)
            self.handle = C.c_void_p()
