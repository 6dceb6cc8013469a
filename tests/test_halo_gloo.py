"""N>1 host logic on CPU: block partition and the grouped neighbour halo exchange over
torch.distributed (gloo, world_size 2 and 3), as used by bench.py --workload shock1p2 on GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectralbte_b200 import initial


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, order, periodic, nX, n3, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spectralbte_b200 import halo as H
    lo, hi = initial.partition(nX, world)[rank]
    n = hi - lo
    slab = torch.full((n + 2 * order, n3), -1.0, dtype=torch.float64)
    for l in range(n):  # owned cell with global index g holds the value g + 0.25
        slab[order + l] = float(lo + l) + 0.25
    flat = slab.view(-1)

    def regions(side):
        if side == 0:
            return flat[order * n3:2 * order * n3], flat[0:order * n3]
        return flat[n * n3:(n + order) * n3], flat[(n + order) * n3:(n + 2 * order) * n3]

    H.exchange(rank, world, regions, periodic)
    out[rank] = slab[:, 0].clone().numpy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,order,periodic", [(2, 1, False), (2, 2, False), (2, 1, True), (3, 2, False), (3, 1, True)])
def test_halo_exchange_matches_global_array(world, order, periodic):
    nX, n3 = 11, 8  # uneven partition on purpose
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, order, periodic, nX, n3, out), nprocs=world, join=True)
    parts = initial.partition(nX, world)
    assert parts[0][0] == 0 and parts[-1][1] == nX and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    for r, (lo, hi) in enumerate(parts):
        col = out[r]
        n = hi - lo
        for g in range(order):
            left_global = lo - order + g
            right_global = hi + g
            if left_global >= 0:
                assert col[g] == left_global + 0.25
            elif periodic:
                assert col[g] == (nX - order + g) + 0.25          # wraps to the last cells
            else:
                assert col[g] == -1.0                               # physical boundary: library fills it
            if right_global < nX:
                assert col[n + order + g] == right_global + 0.25
            elif periodic:
                assert col[n + order + g] == g + 0.25
            else:
                assert col[n + order + g] == -1.0
        assert np.array_equal(col[order:n + order], np.arange(lo, hi) + 0.25)


def test_partition_allows_uneven_blocks():
    assert initial.partition(601, 8) == [(0, 76), (76, 151), (151, 226), (226, 301), (301, 376), (376, 451),
                                         (451, 526), (526, 601)]
    assert [b - a for a, b in initial.partition(250, 8)] == [32, 32, 31, 31, 31, 31, 31, 31]
    assert initial.partition(640, 8) == [(80 * r, 80 * (r + 1)) for r in range(8)]


def test_peer_neighbours_topology():
    """Who maps whom in the peer-memory halo: chain for walls / no-flux ends, ring for the periodic order-1 shock."""
    from spectralbte_b200.halo import peer_neighbours
    assert peer_neighbours(0, 1, True) == []
    assert peer_neighbours(0, 4, False) == [(1, 1)]
    assert peer_neighbours(2, 4, False) == [(0, 1), (1, 3)]
    assert peer_neighbours(3, 4, False) == [(0, 2)]
    assert peer_neighbours(0, 4, True) == [(0, 3), (1, 1)]
    assert peer_neighbours(3, 4, True) == [(0, 2), (1, 0)]
    assert peer_neighbours(0, 2, True) == [(0, 1), (1, 1)]     # two ranks: both sides are the same neighbour
    for n in (2, 3, 8):                                          # symmetry: if a maps b on the right, b maps a on the left
        for periodic in (False, True):
            for r in range(n):
                for side, nb in peer_neighbours(r, n, periodic):
                    assert (1 - side, r) in peer_neighbours(nb, n, periodic)
