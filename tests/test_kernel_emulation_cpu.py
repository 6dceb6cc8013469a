"""Control flow of the batched convolution kernels, checked without a GPU.  tests/emul/kernel_emul.cpp compiles the
kernel sources of spectralbte_b200/csrc verbatim with g++ through a host shim (tests/emul/cuda_emul.h: one OS thread
per CUDA thread, emulated mbarriers / TMA copies / shared memory -- test infrastructure, never part of the library)
and runs them CTA by CTA on the stream-K schedule the library builds (sbte_batch_schedule_host).  The partial sums
they write, combined the way the inverse transform combines them, must equal the oracle's convolution
(src/collisions.c:127-165) for every cell.  This covers what arithmetic checks cannot: barrier counts (a wrong one
hangs; the shim's bounded wait aborts), TMA coordinates, shared-memory offsets, tile switches, flushes -- so a
change to the kernels' protocol is caught here before any GPU time is spent on it."""
import ctypes as C
import multiprocessing as mp
import os
import subprocess

import numpy as np
import pytest

from conftest import relmax, seeded_f
from oracle import oracle as orc
from test_batch_schedule_cpu import schedule

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emul", "kernel_emul.cpp")
LIB = os.path.join(HERE, "emul", "libkernel_emul.so")
DEPS = [SRC, os.path.join(HERE, "emul", "cuda_emul.h")] + [
    os.path.join(ROOT, "spectralbte_b200", "csrc", f) for f in ("qhat_batch.cu", "common.cuh", "internal.h")]


def _lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in DEPS):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-I", cuda_inc, "-I", os.path.join(ROOT, "include"),
                        "-o", LIB, SRC], check=True, capture_output=True)
    L = C.CDLL(LIB)
    dp, llp, ip, ubp = C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_ubyte)
    L.emul_batched.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, llp, llp, ip, ip, ubp, C.c_int, C.c_int, C.c_int, C.c_int,
                               dp, dp, dp, C.c_double, C.c_double]
    return L


def symmetrise_standard(W, N):
    """symmetrize_weights_kernel (csrc/qhat.cu): W + W o sigma on the smaller plane of each pair, W on self-paired planes, 0 elsewhere."""
    n3 = N ** 3
    Wm = W.reshape(n3, n3)
    idx = np.arange(n3)
    x, y, z = idx // (N * N), (idx // N) % N, idx % N
    out = np.zeros_like(Wm)
    for zeta in range(n3):
        zx, zy, zz = zeta // (N * N), (zeta // N) % N, zeta % N
        X, Y, Z = (zx + N // 2 - x) % N, (zy + N // 2 - y) % N, (zz + N // 2 - z) % N
        sig = (X * N + Y) * N + Z
        row = Wm[zeta]
        out[zeta] = np.where(x < X, row + row[sig], np.where(x == X, row, 0.0))
    return out.reshape(-1)


CASES = [  # kind, N, cells, sym, ctas
    (0, 8, 37, True, 5), (0, 8, 5, False, 3),            # qhat_batch2_kernel<8>
    (0, 16, 3, True, 3),                                   # qhat_batch2_kernel<16> (SBTE_N16_PLANE=1)
    (1, 16, 3, True, 3),                                   # qhat_batch3_kernel<16> (the default at N = 16)
    (2, 16, 3, True, 3),                                   # its instance for schedules cut at whole chunks only
    (0, 20, 3, True, 4),                                   # qhat_batch3_kernel<20> (partly empty row-blocks)
    (0, 22, 2, False, 3),                                  # qhat_batch3_kernel<22>
    (0, 24, 1, True, 3),                                   # qhat_batch3_kernel<24>
]


def run_emulation(tmp_path, kind, N, cells, sym, ctas, tag="", split_group=None, whole=False, anywhere=False):
    """Runs the kernels CTA by CTA on the library's schedule; returns (combined Q^ per cell, oracle, W, spectra).
    split_group = g: only the split-tile launch that serves the last cell group g (at most 16 live cells) is run; the
    result list then holds None for the cells of the other groups.  whole / anywhere: stream-K cuts at whole xi_x chunks only / at any step (default: the library's choice)."""
    L = _lib()
    o = orc.Oracle(N, 9.0, 1)
    n3 = N ** 3
    W = np.random.default_rng(N).standard_normal(n3 * n3) if N <= 8 else orc.synthetic_weights(N)
    Wk = symmetrise_standard(W, N) if sym else W
    Gtot = -(-cells // 32)
    if kind in (0, 2) and N == 16 and split_group is None:
        whole = True   # (kind 2: the line-ring instance without hand-over code)
    if kind == 0 and N == 16 and split_group is None:
        whole = True   # the resident-plane kernel (what SBTE_N16_PLANE=1 selects, with whole-chunk cuts) cannot take a cut chunk
    if split_group is None:
        s = schedule(N, cells, sym, ctas, whole=whole, anywhere=anywhere and not whole)
        assert s["G"] == Gtot
    else:
        assert split_group == Gtot - 1 and 1 <= cells - 32 * split_group <= 16
        s = schedule(N, cells - 32 * split_group, sym, ctas, split=True, whole=whole, anywhere=anywhere)
        assert s["G"] == 1
        kind = 100 + split_group
    G, T, P, kmax = s["G"], s["T"], s["P"], s["kmax"]
    # spectra, cell-minor [Gtot][n3][32]; padding cells are zero
    spec = np.zeros((Gtot, n3, 32), dtype=complex)
    F = []
    for b in range(cells):
        f = seeded_f(o.v, 700 + b, noise=0.3) * (1.0 + 0.05 * b)
        Fb = o.fft3d(f.astype(complex))
        F.append(Fb)
        spec[b // 32, :, b % 32] = Fb
    stride = Gtot * 32 * n3
    parts = np.full(kmax * stride, np.nan + 1j * np.nan, dtype=complex)
    dv = o.v[1] - o.v[0]
    L_eta = 0.5 * N * (2.0 * np.pi / (N * dv))
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    out = str(tmp_path / ("parts%s.npy" % tag))

    def child():   # a deadlocked protocol ends the emulation with _Exit: keep that out of the test runner
        rc = L.emul_batched(kind, N, cells, int(sym), P, p(s["begin"], C.c_longlong), p(s["tbegin"], C.c_longlong), p(s["ctile"], C.c_int),
                            p(s["first"], C.c_int), p(s["np"], C.c_ubyte), G, T, s["np_cols"], kmax,
                            p(Wk, C.c_double), p(spec.view(np.float64), C.c_double), p(parts.view(np.float64), C.c_double), L_eta, 9.0)
        if rc == 0:
            np.save(out, parts)
        os._exit(rc)

    proc = mp.get_context("fork").Process(target=child)
    proc.start()
    proc.join(900)
    assert proc.exitcode == 0, "kernel emulation failed or deadlocked (exit code %s)" % proc.exitcode
    parts = np.load(out)
    # combine the partial sums the way the inverse transform does: a left fold over np[(column / np_cols) * G + cell group]
    # parts per column, starting from zero (csrc/fft.cu)
    parts = parts.reshape(kmax, Gtot * 32, n3)
    Q = []
    for b in range(cells):
        if split_group is not None and b // 32 != split_group:
            Q.append(None)
            continue
        g = 0 if split_group is not None else b // 32
        q = np.zeros(n3, dtype=complex)
        for col in range(N * N):
            npc = int(s["np"][(col // s["np_cols"]) * G + g])
            assert npc >= 1
            seg = parts[:npc, b, col * N:(col + 1) * N]
            assert not np.isnan(seg.view(np.float64)).any(), (b, col)       # every part the table promises was written
            acc = np.zeros(N, dtype=complex)
            for m in range(npc):
                acc = acc + seg[m]
            q[col * N:(col + 1) * N] = acc
        Q.append(q)
    return Q, o, W, F


@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")   # the oracle's OpenMP pool; the child only runs the emulation
@pytest.mark.parametrize("kind,N,cells,sym,ctas", CASES)
def test_kernel_control_flow_on_host(tmp_path, kind, N, cells, sym, ctas):
    Q, o, W, F = run_emulation(tmp_path, kind, N, cells, sym, ctas)
    for b in range(cells):
        assert relmax(Q[b], o.qhat(W, F[b], F[b])) < 1e-12, b


@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")
@pytest.mark.parametrize("N,sym,a,b", [(8, True, (37, 5), (37, 3)), (8, True, (37, 5), (5, 4)), (8, False, (40, 7), (9, 2)),
                                       (20, True, (3, 4), (2, 7))])
def test_summation_order_is_independent_of_the_schedule(tmp_path, N, sym, a, b):
    """Rank-count invariance at the source: the same cell gives the same BITS whatever the number of CTAs and however
    many other cells share the launch (the reference's per-cell loop, exec/boltz.c:285-345, has that property by
    construction).  The kernels get it from a canonical order -- left fold over xi_x chunk sums -- in chunk_end()."""
    Qa, *_ = run_emulation(tmp_path, 0, N, a[0], sym, a[1], "a")
    Qb, *_ = run_emulation(tmp_path, 0, N, b[0], sym, b[1], "b")
    for cell in range(min(a[0], b[0])):
        assert np.array_equal(Qa[cell].view(np.float64), Qb[cell].view(np.float64)), cell


@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")
@pytest.mark.parametrize("sym,cells,ctas_split", [(True, 35, 40), (False, 44, 7)])
def test_split_tiles_give_the_bits_of_ordinary_tiles(tmp_path, sym, cells, ctas_split):
    """N = 16: a last cell group with at most 16 live cells runs on split tiles (two zeta_y columns per warp, 16 cells each,
    Batch3Cfg<16, true>).  Same per-lane arithmetic in the same canonical order: its cells must come out bit-identical to
    the ordinary line-ring tiles (kind 1) serving the padded group -- and both must equal the oracle to 1e-12."""
    g = (cells - 1) // 32
    Qs, o, W, F = run_emulation(tmp_path, 0, 16, cells, sym, ctas_split, "s", split_group=g)
    Qr, *_ = run_emulation(tmp_path, 1, 16, cells, sym, 5, "r")
    for b in range(32 * g, cells):
        assert relmax(Qs[b], o.qhat(W, F[b], F[b])) < 1e-12, b
        assert np.array_equal(Qs[b].view(np.float64), Qr[b].view(np.float64)), b


@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")
@pytest.mark.parametrize("kind,N,cells,sym,ctas,split", [(1, 16, 3, True, 7, None), (0, 20, 2, False, 9, None), (0, 24, 1, True, 11, None),
                                                         (0, 16, 35, True, 13, 1)])
def test_cuts_inside_a_chunk_give_the_bits_of_whole_chunk_cuts(tmp_path, kind, N, cells, sym, ctas, split):
    _cuts_vs_whole(tmp_path, kind, N, cells, sym, ctas, split)


@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")
def test_whole_chunk_instance_gives_the_same_bits(tmp_path):
    """N = 16 has two instances of the line-ring kernel: the general one and, for schedules cut at whole chunks, one
    without hand-over code.  Same chunk sums, same fold: same bits."""
    Qg, *_ = run_emulation(tmp_path, 1, 16, 3, True, 7, "g", anywhere=True)
    Qw, *_ = run_emulation(tmp_path, 2, 16, 3, True, 7, "w")
    for b in range(3):
        assert np.array_equal(Qg[b].view(np.float64), Qw[b].view(np.float64)), b


def _cuts_vs_whole(tmp_path, kind, N, cells, sym, ctas, split):
    """Line-ring kernels: a stream-K range may end inside a xi_x chunk.  The CTA that begins the chunk hands its running
    sum to the next one (BatchSched::carry), which continues with the same operations in the same order -- so the chunk
    sum, and with it every cell, has the BITS of a schedule cut at whole chunks (and of the oracle to 1e-12).  The
    emulation also fails if a running sum is awaited that nobody published, or published and never taken."""
    Qc, o, W, F = run_emulation(tmp_path, kind, N, cells, sym, ctas, "c", split_group=split, anywhere=True)
    Qw, *_ = run_emulation(tmp_path, kind, N, cells, sym, ctas, "w", split_group=split, whole=True)
    s = schedule(N, cells - 32 * split if split is not None else cells, sym, ctas, split=split is not None, anywhere=True)
    assert np.any(s["begin"][1:-1] % N != 0)          # the case does cut chunks
    for b in range(cells):
        if Qc[b] is None:
            continue
        assert relmax(Qc[b], o.qhat(W, F[b], F[b])) < 1e-12, b
        assert np.array_equal(Qc[b].view(np.float64), Qw[b].view(np.float64)), b
