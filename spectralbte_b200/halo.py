"""Ghost-cell exchange between the slabs of neighbouring ranks (one process per GPU).

Replaces the blocking, even/odd-ordered MPI_Send/MPI_Recv sequence of the reference
(/root/reference/src/transportroutines.c:107-172 order 1, :261-344 order 2) with one grouped
non-blocking exchange per upwind pass over torch.distributed (NCCL on GPUs, gloo in the CPU tests):
each rank sends its first/last `order` owned cells and receives its left/right ghost cells.
Init_field 6 at order 1 is periodic: rank 0 and rank n-1 also exchange (:156-166).

On one NVLink node the default mode has no messages at all (SlabHalo(..., mode="p2p"); SBTE_HALO=nccl selects messages):
the ranks swap CUDA IPC handles once, and from then on the upwind kernels read the neighbours' boundary cells
directly from peer memory, ordered by device-side counters (csrc/slab.cu peer_begin/peer_end).
"""
import os


import torch
import torch.distributed as dist


class DevicePtr:
    """Exposes a raw device pointer (n doubles) through __cuda_array_interface__ so torch can wrap it
    without a copy."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def wrap(ptr, n, device):
    return torch.as_tensor(DevicePtr(ptr, n), device=device)


def neighbour_ops(rank, nranks, regions, periodic):
    """Builds the P2POp list of one exchange. regions(side) -> (send_tensor, recv_tensor) with side 0 =
    left, 1 = right. Pure function of (rank, nranks, periodic): unit-tested on CPU with gloo."""
    ops = []
    left, right = rank - 1, rank + 1
    if rank > 0:
        s, r = regions(0)
        ops += [dist.P2POp(dist.isend, s, left), dist.P2POp(dist.irecv, r, left)]
    if rank < nranks - 1:
        s, r = regions(1)
        ops += [dist.P2POp(dist.isend, s, right), dist.P2POp(dist.irecv, r, right)]
    if periodic and nranks > 1:
        if rank == 0:
            s, r = regions(0)
            ops += [dist.P2POp(dist.isend, s, nranks - 1), dist.P2POp(dist.irecv, r, nranks - 1)]
        if rank == nranks - 1:
            s, r = regions(1)
            ops += [dist.P2POp(dist.isend, s, 0), dist.P2POp(dist.irecv, r, 0)]
    return ops


def peer_neighbours(rank, nranks, periodic):
    """[(side, neighbour_rank)] of the peer-memory halo; side 0 = left, 1 = right."""
    out = []
    if rank > 0:
        out.append((0, rank - 1))
    elif periodic and nranks > 1:
        out.append((0, nranks - 1))
    if rank < nranks - 1:
        out.append((1, rank + 1))
    elif periodic and nranks > 1:
        out.append((1, 0))
    return out


def exchange(rank, nranks, regions, periodic):
    ops = neighbour_ops(rank, nranks, regions, periodic)
    if not ops:
        return
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class SlabHalo:
    """Halo exchange for a spectralbte_b200.Slab on a GPU: wraps the regions the library reports
    (sbte_slab_halo_regions) as torch tensors on the library's stream."""

    def __init__(self, slab, device, mode=None):
        self.slab, self.device = slab, device
        self.stream = torch.cuda.ExternalStream(slab.coll.stream, device=device)
        self.periodic = False
        self._cache = {}
        self.mode = mode or os.environ.get("SBTE_HALO", "p2p")
        if self.mode not in ("nccl", "p2p"):
            raise ValueError("halo mode must be 'nccl' or 'p2p'")
        if slab.nranks == 1:
            self.mode = "nccl"   # nothing to exchange
        if self.mode == "p2p" and not self._connect_peers():
            self.mode = "nccl"   # no peer access between these GPUs: fall back to messages (every rank agrees)

    def _connect_peers(self):
        """Swap IPC handles and map the neighbours' slabs (the ring closes for the periodic order-1 case).
        Returns False -- on every rank -- if any rank could not map its neighbours."""
        slab = self.slab
        ok = 1
        try:
            mine = (slab.ipc_export(), slab.cells_local)
        except Exception:
            mine, ok = None, 0
        everyone = [None] * slab.nranks
        dist.all_gather_object(everyone, mine)
        if any(e is None for e in everyone):
            ok = 0
        if ok:
            try:
                for side, nb in peer_neighbours(slab.rank, slab.nranks, slab.init_field == 6 and slab.order == 1):
                    handles, cells = everyone[nb]
                    slab.ipc_import(side, handles, cells)
                slab.set_peer_halo(True)
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # also the barrier: nobody polls flags before every mapping exists
        if int(flag.item()) == 0:
            try:
                slab.set_peer_halo(False)
            except Exception:
                pass
            return False
        return True

    def close(self):
        """Collective: every rank unmaps its neighbours, then the ranks synchronise; after that the slabs may be
        destroyed in any order (memory exported over CUDA IPC must not be freed while a peer still maps it)."""
        if self.slab.nranks > 1:
            self.slab.coll.sync()
            if self.mode == "p2p":
                self.slab.peer_detach()
            dist.barrier()

    def _regions(self, which, stage):
        key = (which, stage)
        if key not in self._cache:
            out = {}
            for side in (0, 1):
                s, r, n = self.slab.halo_regions(which, stage, side)
                out[side] = (wrap(s, n, self.device), wrap(r, n, self.device))
            self._cache[key] = out
        return self._cache[key]

    def exchange(self, which, stage, periodic=False):
        reg = self._regions(which, stage)
        with torch.cuda.stream(self.stream):
            exchange(self.slab.rank, self.slab.nranks, lambda side: reg[side], periodic)


def advect(slab, halo, which, periodic):
    """advectOne / advectTwo across ranks: halo, upwind pass (x2 for order 2), average."""
    for stage in range(slab.order):
        if slab.nranks > 1 and halo.mode == "nccl":
            halo.exchange(which, stage, periodic)
        slab.upwind_stage(which, stage)
    slab.advect_finish(which)


def step(slab, halo, Kn, init_field, k2=0):
    """One time step of exec/boltz.c:264-353 on this rank's slab."""
    if slab.nranks == 1 or halo.mode == "p2p":
        slab.step(Kn, k2)   # nothing but launches on one stream: the library replays it from a CUDA graph
        return
    periodic = (init_field == 6 and slab.order == 1)
    advect(slab, halo, 0, periodic)
    slab.collide(Kn, k2)
    if slab.order == 2:
        advect(slab, halo, 1, periodic)
