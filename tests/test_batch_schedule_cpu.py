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


def schedule(N, cells, sym, ctas, split=False, whole=False, anywhere=False):
    """split: the split-tile launch of a remainder group; whole: cuts at whole xi_x chunks only (SBTE_CHUNK_CUTS=1);
    anywhere: cuts at any step wherever the kernel takes them (SBTE_CHUNK_CUTS=0); neither: the library's choice."""
    L = _lib()
    dims = (C.c_int * 6)()
    split = int(split) | (2 if whole else 0) | (4 if anywhere else 0)
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


def segments(N, g0, g1):
    """The order in which qhat_batch3_kernel walks its range [g0, g1): the head of the chunk cut by g1, the whole
    chunks, the tail of the chunk cut by g0 (csrc/qhat_batch.cu).  Yields (chunk start, ey0, ey1)."""
    w0, w1 = -(-g0 // N) * N, (g1 // N) * N
    if g1 > w1:
        yield w1, 0, g1 - w1
    for cs in range(w0, w1, N):
        yield cs, 0, N
    if g0 < w0:
        yield w0 - N, g0 - (w0 - N), N


@pytest.mark.parametrize("cuts", ["auto", "whole", "anywhere"])
@pytest.mark.parametrize("N,cells,sym,ctas", [(8, 5, True, 148), (8, 37, False, 148), (16, 640, True, 148), (16, 80, True, 148),
                                              (16, 33, False, 7), (24, 250, True, 148), (24, 33, False, 148),
                                              (22, 250, True, 148), (22, 34, False, 148), (20, 33, True, 148),
                                              (22, 31, True, 148), (22, 70, True, 13), (16, 1, True, 148), (16, 32, True, 148)])
def test_schedule_covers_every_step_once(N, cells, sym, ctas, cuts):
    s = schedule(N, cells, sym, ctas, whole=cuts == "whole", anywhere=cuts == "anywhere")
    cols = 8 if N >= 16 else 4                       # Batch2Cfg / Batch3Cfg::COLS
    bpx = -(-N // cols)                              # row-blocks per zeta_x plane
    G, T, P = s["G"], s["T"], s["P"]
    assert G == -(-cells // 32) and T == G * N * bpx and 1 <= P <= ctas
    begin, tbegin = s["begin"], s["tbegin"]
    assert begin[0] == 0 and begin[-1] == tbegin[-1] and np.all(np.diff(begin) >= 0)
    aligned = bool(np.all(begin % N == 0))
    if cuts == "whole" or N == 8:                    # the resident-plane kernel works on whole xi_x chunks
        assert aligned
    if cuts == "auto" and not aligned:
        # chosen only where it is predicted to pay: equal shares beat the busiest CTA of the whole-chunk schedule
        w = schedule(N, cells, sym, ctas, whole=True)
        assert np.diff(begin).max() < np.diff(w["begin"]).max()
    if not aligned:
        # any step may be a cut: the shares differ by at most one step, and every CTA holds at least one chunk's worth, so
        # that the two chunks it shares with its neighbours are different chunks
        assert np.diff(begin).max() - np.diff(begin).min() <= 1 and np.diff(begin).min() >= N
    # tile lengths: (visited xi_x planes) * N steps
    for t in range(T):
        zx = (t // G) // bpx
        assert tbegin[t + 1] - tbegin[t] == (sym_nrep(N, zx) if sym else N) * N
    # walk every CTA's range as the kernels do
    seen = {}                                        # (zx, zy, cg, ex, ey) -> CTA
    writes = {}                                      # tile -> set of part indices written
    done_by = {}                                     # (tile, chunk) -> CTA that completes it (and writes it)
    for p in range(P):
        g0, g1 = int(begin[p]), int(begin[p + 1])
        if g1 <= g0:
            continue
        t0 = int(s["ctile"][p])
        assert tbegin[t0] <= g0 < tbegin[t0 + 1]
        for cs, ey0, ey1 in segments(N, g0, g1):
            t = int(np.searchsorted(tbegin, cs, side="right")) - 1
            assert t >= t0
            rb, cg = divmod(t, G)
            zx, zy0 = rb // bpx, (rb % bpx) * cols
            c = (cs - int(tbegin[t])) // N
            ex = sym_rep(N, zx, c) if sym else c
            assert 0 <= ex < N
            for ey in range(ey0, ey1):
                for w in range(cols):
                    if zy0 + w < N:                  # surplus warps of a partly empty row-block do nothing
                        key = (zx, zy0 + w, cg, ex, ey)
                        assert key not in seen
                        seen[key] = p
            if ey0 > 0:                              # continued from the previous CTA, which computed [0, ey0) of it FIRST
                assert p > 0 and int(begin[p]) == cs + ey0 and next(iter(segments(N, int(begin[p - 1]), g0))) == (cs, 0, ey0)
            if ey1 < N:
                continue                             # handed over: the next CTA completes it
            # chunk_end() of the kernels: the CTA owning the tile's first step folds the chunks it completes into part 0, a
            # later CTA writes every chunk it completes as part (e0 > 0) + (c - e0), e0 = chunks the first CTA completes
            first = int(s["first"][t])
            assert (t, c) not in done_by
            done_by[(t, c)] = p
            if p == first:
                part = 0
            else:
                assert p > first
                e0 = (int(begin[first + 1]) - int(tbegin[t])) // N
                part = (1 if e0 > 0 else 0) + c - e0
                assert part >= (1 if e0 > 0 else 0)
            assert 0 <= part < s["kmax"]
            if p != first or c == 0:
                assert part not in writes.get(t, set())   # a plain store: nobody else writes that part
            writes.setdefault(t, set()).add(part)
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
    """qhat_batch3_kernel: a segment [ey0, ey1) of a chunk streams the lines j' = ey0 .. COLS-2 + ey1; line jl is read by
    the warps w with ey0 <= jl - (COLS-1) + w < ey1; the reader that comes last in step order also arrives for the warps
    that never read the line, so every slot's empty barrier (COLS arrivals) completes exactly once per use.  Whole
    chunks are the segment [0, N)."""
    COLS = 8
    RING = 8 + (4 if N <= 16 else 3 if N <= 20 else 2)   # Batch3Cfg::RING = COLS + STAGES
    for ey0, ey1 in [(0, N), (0, 1), (0, 5), (N - 1, N), (3, N), (0, N - 1), (N // 2, N), (0, COLS), (N - COLS, N)]:
        lo, hi = ey0, COLS - 1 + ey1                  # lines [lo, hi)
        arrivals = {j: 0 for j in range(lo, hi)}
        readers = {j: 0 for j in range(lo, hi)}
        for ey in range(ey0, ey1):
            for w in range(COLS):
                jl = COLS - 1 + ey - w
                assert lo <= jl < hi
                readers[jl] += 1
                cmax = min(COLS - 1, ey1 - 1 - jl + COLS - 1)
                cmin = max(0, ey0 - jl + COLS - 1)
                arrivals[jl] += COLS - (cmax - cmin) if w == cmax else 1
                # the line warp w reads at step ey is zeta_y + N/2 - xi_y (mod N) of its column
                zy0 = 8
                Y_line = (zy0 + N // 2 + COLS - 1 - jl) % N
                assert Y_line == (zy0 + w + N // 2 - ey) % N
        assert all(a == COLS for a in arrivals.values()), (ey0, ey1)
        assert all(1 <= r <= COLS for r in readers.values())
        # the producer needs line j' <= COLS-1 + ey before step ey: never more than COLS + 1 lines ahead of the oldest
        # line still being read, which must fit the ring
        for ey in range(ey0, ey1):
            newest = min(hi - 1, COLS - 1 + ey)
            oldest = max(lo, ey)         # warp COLS-1 reads line ey at step ey
            assert newest - oldest + 1 <= RING
