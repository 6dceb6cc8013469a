"""Host logic of the batched convolution, checked without a GPU: the stream-K schedule the library uploads
(sbte_batch_schedule_host == the tables ensure_batch_schedule() builds, csrc/capi.cu) is walked exactly as
qhat_batch2_kernel / qhat_batch3_kernel walk it (csrc/qhat_batch.cu) and must visit every (zeta column, cell
group, xi_x plane, xi_y) step of the reference's N^6 loop (src/collisions.c:127-165) exactly once."""
import ctypes as C

import numpy as np
import pytest


def _lib():
    from spectralbte_b200 import _lib, build
    build.build()
    return _lib.load()


def schedule(N, cells, sym, ctas, split=False):
    L = _lib()
    dims = (C.c_int * 6)()
    assert L.sbte_batch_schedule_host(N, cells, int(sym), ctas, int(split), None, None, None, None, None, dims) == 0
    G, T, P, np_cols, kmax, np_len = list(dims)
    begin = np.zeros(P + 1, dtype=np.int64)
    tbegin = np.zeros(T + 1, dtype=np.int64)
    ctile = np.zeros(P, dtype=np.int32)
    first = np.zeros(T, dtype=np.int32)
    npt = np.zeros(np_len, dtype=np.uint8)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    assert L.sbte_batch_schedule_host(N, cells, int(sym), ctas, int(split), p(begin, C.c_longlong), p(tbegin, C.c_longlong),
                                      p(ctile, C.c_int), p(first, C.c_int), p(npt, C.c_ubyte), dims) == 0
    return dict(G=G, T=T, P=P, np_cols=np_cols, kmax=kmax, begin=begin, tbegin=tbegin, ctile=ctile, first=first, np=npt)


def sym_nrep(N, zx):
    a = (zx + N // 2) % N
    return a // 2 + 1 + (a + N) // 2 - a


def sym_rep(N, zx, c):
    a = (zx + N // 2) % N
    h = a // 2
    return c if c <= h else c + (a - h)


@pytest.mark.parametrize("N,cells,sym,ctas", [(8, 5, True, 148), (8, 37, False, 148), (16, 640, True, 148), (16, 80, True, 148),
                                              (16, 33, False, 7), (24, 250, True, 148), (24, 33, False, 148),
                                              (22, 250, True, 148), (22, 34, False, 148), (20, 33, True, 148),
                                              (22, 31, True, 148), (22, 70, True, 13), (16, 1, True, 148)])
def test_schedule_covers_every_step_once(N, cells, sym, ctas):
    s = schedule(N, cells, sym, ctas)
    cols = 8 if N >= 16 else 4                       # Batch2Cfg / Batch3Cfg::COLS
    bpx = -(-N // cols)                              # row-blocks per zeta_x plane
    ring = True                                      # every kernel works on whole xi_x chunks (canonical summation order)
    G, T, P = s["G"], s["T"], s["P"]
    assert G == -(-cells // 32) and T == G * N * bpx and 1 <= P <= ctas
    begin, tbegin = s["begin"], s["tbegin"]
    assert begin[0] == 0 and begin[-1] == tbegin[-1] and np.all(np.diff(begin) >= 0)
    if ring:
        assert np.all(begin % N == 0)
    # tile lengths: (visited xi_x planes) * N steps
    for t in range(T):
        zx = (t // G) // bpx
        assert tbegin[t + 1] - tbegin[t] == (sym_nrep(N, zx) if sym else N) * N
    # walk every CTA's range as the kernels do
    seen = {}                                        # (zx, zy, cg, ex, ey) -> part index
    writes = {}                                      # tile -> set of part indices written
    for p in range(P):
        g0, n = int(begin[p]), int(begin[p + 1] - begin[p])
        if n <= 0:
            continue
        t = int(s["ctile"][p])
        assert tbegin[t] <= g0 < tbegin[t + 1]
        te = int(tbegin[t + 1])
        sl = g0 - int(tbegin[t])
        for k in range(n):
            if g0 + k == te:
                t += 1
                te = int(tbegin[t + 1])
                sl = 0
            rb, cg = divmod(t, G)
            zx, zy0 = rb // bpx, (rb % bpx) * cols
            c, ey = divmod(sl, N)
            ex = sym_rep(N, zx, c) if sym else c
            assert 0 <= ex < N
            # chunk_end() of the kernels: the CTA owning the tile's first chunk folds into part 0, a later CTA writes
            # every chunk as part 1 + (c - e0), e0 = chunks held by the first CTA
            first = int(s["first"][t])
            if p == first:
                assert sl == k if t == int(s["ctile"][p]) and g0 == tbegin[t] else True
                part = 0
            else:
                assert p > first
                e0 = (int(begin[first + 1]) - int(tbegin[t])) // N
                part = 1 + c - e0
            assert 0 <= part < s["kmax"]
            writes.setdefault(t, set()).add(part)
            for w in range(cols):
                if zy0 + w < N:                      # surplus warps of a partly empty row-block do nothing
                    key = (zx, zy0 + w, cg, ex, ey)
                    assert key not in seen
                    seen[key] = part
            if ring and ey == 0:
                assert n - k >= N                    # a chunk never straddles the end of the CTA's range
            sl += 1
    # every step of the reference loop exactly once (restricted to the representative planes when symmetrised)
    want = 0
    for zx in range(N):
        want += (sym_nrep(N, zx) if sym else N) * N * N * G
    assert len(seen) == want
    # the part count the inverse transform reads per zeta column equals the number of parts written for its tile
    for zx in range(N):
        for zy in range(N):
            for cg in range(G):
                t = (zx * bpx + zy // cols) * G + cg
                q = zx * N + zy
                got = int(s["np"][(q // s["np_cols"]) * G + cg])
                assert writes[t] == set(range(got)), (zx, zy, cg)
    assert s["kmax"] == max(len(v) for v in writes.values())


def test_schedule_rejects_unscheduled_n():
    L = _lib()
    dims = (C.c_int * 6)()
    assert L.sbte_batch_schedule_host(12, 40, 1, 148, 0, None, None, None, None, None, dims) != 0
    assert b"any-N" in L.sbte_last_error()
    assert L.sbte_batch_schedule_host(16, 0, 1, 148, 0, None, None, None, None, None, dims) != 0


@pytest.mark.parametrize("N", [16, 20, 22, 24])
def test_line_ring_arrival_counts_complete_every_slot(N):
    """qhat_batch3_kernel: each of the L = N + COLS - 1 lines of a chunk is read by the warps w with
    0 <= jl - (COLS-1) + w < N; the reader that comes last in step order also arrives for the warps that never
    read the line, so every slot's empty barrier (COLS arrivals) completes exactly once per use."""
    COLS = 8
    L = N + COLS - 1
    arrivals = [0] * L
    readers = [0] * L
    for ey in range(N):
        for w in range(COLS):
            jl = COLS - 1 + ey - w
            assert 0 <= jl < L
            readers[jl] += 1
            cnt = 1
            if jl < COLS - 1 and w == COLS - 1:
                cnt = COLS - jl
            if jl > N - 1 and w == N + COLS - 2 - jl:
                cnt = jl - N + 2
            arrivals[jl] += cnt
            # the line warp w reads at step ey is zeta_y + N/2 - xi_y (mod N) of its column
            zy0 = 8
            Y_line = (zy0 + N // 2 + COLS - 1 - jl) % N
            assert Y_line == (zy0 + w + N // 2 - ey) % N
    assert arrivals == [COLS] * L
    assert all(1 <= r <= COLS for r in readers)
    # the producer needs line j' <= COLS-1 + ey before step ey: never more than COLS + 1 lines ahead of the oldest
    # line still being read, which must fit the ring
    RING = 8 + (4 if N <= 16 else 3 if N <= 20 else 2)   # Batch3Cfg::RING = COLS + STAGES
    for ey in range(N):
        newest = min(L - 1, COLS - 1 + ey)
        oldest = max(0, ey)          # warp COLS-1 reads line ey at step ey
        assert newest - oldest + 1 <= RING
