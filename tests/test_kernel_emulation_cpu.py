"""Control flow of the batched convolution kernels, checked without a GPU.  tests/emul/kernel_emul.cpp compiles the
kernel sources of spectralbte_b200/csrc verbatim with g++ through a host shim (tests/emul/cuda_emul.h: one OS thread
per CUDA thread, emulated mbarriers / TMA copies / shared memory -- test infrastructure, never part of the library)
and runs them CTA by CTA on the stream-K schedule the library builds (sbte_batch_schedule_host).  The partial sums
they write, combined the way the inverse transform combines them, must equal the oracle's convolution
(src/collisions.c:127-165) for every cell.  This covers what arithmetic checks cannot: barrier counts (a wrong one
hangs; the shim's bounded wait aborts), TMA coordinates, shared-memory offsets, tile switches, flushes -- for the
GPU-verified kernels (as a check of the emulation itself) and for the opt-in kernels that have not run on a GPU yet."""
import ctypes as C
import multiprocessing as mp
import os
import subprocess

import numpy as np
import pytest

from conftest import relmax, seeded_f
from oracle import oracle as orc
from test_batch_schedule_cpu import schedule
from test_mirror_emulation_cpu import _emul as _mirror_rule_lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emul", "kernel_emul.cpp")
LIB = os.path.join(HERE, "emul", "libkernel_emul.so")
DEPS = [SRC, os.path.join(HERE, "emul", "cuda_emul.h")] + [
    os.path.join(ROOT, "spectralbte_b200", "csrc", f) for f in ("qhat_batch.cu", "qhat_mirror.cu", "mirror.cuh", "common.cuh", "internal.h")]


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


# (0, 16, 3, True, 3) and (0, 22, 2, True, 3) -- the GPU-verified N=16 / N=22 kernels -- pass as well; left out to keep the
# CPU suite short (same templates as the N=8 / N=20 cases)
CASES = [  # kind, N, cells, sym, ctas
    (0, 8, 37, True, 5), (0, 8, 5, False, 3),            # qhat_batch2_kernel<8>      (GPU-verified: checks the emulation)
    (0, 20, 3, True, 4),                                   # qhat_batch3_kernel<20>     (GPU-verified, partly empty row-blocks)
    (1, 8, 37, True, 5), (1, 8, 5, False, 3),             # qhat_mirror_kernel<8>
    (1, 16, 3, True, 3),                                   # qhat_mirror_kernel<16>
    (1, 20, 3, True, 4), (1, 20, 2, False, 3),            # qhat_mirror_ring_kernel<20>
    (1, 22, 2, True, 3),                                   # qhat_mirror_ring_kernel<22> (odd N/2, padded box slots, partial tiles)
    (1, 24, 1, True, 3),                                   # qhat_mirror_ring_kernel<24>
    (2, 24, 1, True, 3),                                   # qhat_batch3_kernel<24, ROLL=3> (opt-in rolled xi_z loop)
    (2, 22, 1, False, 3),                                  # qhat_batch3_kernel<22, ROLL=11>
    (3, 8, 37, True, 5), (3, 8, 5, False, 3),             # qhat_mirror_kernel<8> on the folded tensor (combined body)
    (3, 16, 3, True, 3),                                   # qhat_mirror_kernel<16> on the folded tensor
]


@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")   # the oracle's OpenMP pool; the child only runs the emulation
@pytest.mark.parametrize("kind,N,cells,sym,ctas", CASES)
def test_kernel_control_flow_on_host(tmp_path, kind, N, cells, sym, ctas):
    L = _lib()
    o = orc.Oracle(N, 9.0, 1)
    n3 = N ** 3
    W = np.random.default_rng(N).standard_normal(n3 * n3) if N <= 8 else orc.synthetic_weights(N)
    if kind == 3:
        Wk = np.empty_like(W)
        R = _mirror_rule_lib()
        assert R.mirror_emul_fold(N, W.ctypes.data_as(C.POINTER(C.c_double)), int(sym), Wk.ctypes.data_as(C.POINTER(C.c_double))) == 0
    elif sym and kind == 1:
        Wk = np.empty_like(W)
        R = _mirror_rule_lib()
        assert R.mirror_emul_symmetrize(N, W.ctypes.data_as(C.POINTER(C.c_double)), Wk.ctypes.data_as(C.POINTER(C.c_double))) == 0
    elif sym:
        Wk = symmetrise_standard(W, N)
    else:
        Wk = W
    s = schedule(N, cells, sym, ctas, mirror=(kind in (1, 3)))
    G, T, P, kmax = s["G"], s["T"], s["P"], s["kmax"]
    # spectra, cell-minor [G][n3][32]; padding cells are zero
    spec = np.zeros((G, n3, 32), dtype=complex)
    F, fs = [], []
    for b in range(cells):
        f = seeded_f(o.v, 700 + b, noise=0.3) * (1.0 + 0.05 * b)
        Fb = o.fft3d(f.astype(complex))
        F.append(Fb)
        fs.append(f)
        spec[b // 32, :, b % 32] = Fb
    stride = G * 32 * n3
    parts = np.full(kmax * stride, np.nan + 1j * np.nan, dtype=complex)
    dv = o.v[1] - o.v[0]
    L_eta = 0.5 * N * (2.0 * np.pi / (N * dv))
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))  # noqa: E731
    out = str(tmp_path / "parts.npy")

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
    # combine the partial sums the way the inverse transform does: np[(column / np_cols) * G + cell group] parts per column
    parts = parts.reshape(kmax, G * 32, n3)
    for b in range(cells):
        q = np.zeros(n3, dtype=complex)
        for col in range(N * N):
            npc = int(s["np"][(col // s["np_cols"]) * G + b // 32])
            assert npc >= 1
            seg = parts[:npc, b, col * N:(col + 1) * N]
            assert not np.isnan(seg.view(np.float64)).any(), (b, col)       # every part the table promises was written
            q[col * N:(col + 1) * N] = seg.sum(axis=0)
        if kind == 3:   # the folded tensor does not give Q^ but a spectrum with the same Q = Re(fft3D^-1(.))
            assert relmax(np.real(o.fft3d(q, invert=True)), o.compute_q(W, fs[b], fs[b])) < 1e-12, b
        else:
            assert relmax(q, o.qhat(W, F[b], F[b])) < 1e-12, b


@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")
@pytest.mark.parametrize("N,nsplit,packed,npairs", [(16, 2, 1, 1), (16, 1, 1, 2)])   # (16, 1, 0, 1) -- leftovers gathered -- passes too
def test_half_spectrum_0d_kernels_on_host(tmp_path, N, nsplit, packed, npairs):
    """qhat_stream_half_kernel + qhat_half_leftover_kernel (csrc/qhat_half.cu, opt-in SBTE_HALF0D=1): the 0D stream kernel on the
    folded tensor, mirror columns skipping the folded steps, leftovers added by the second kernel -- for ComputeQ(f, f) (one
    operand pair) and ComputeQ_maxPreserve (two pairs sharing the weight pass, src/collisions.c:178-210).  The sum of the
    partial spectra is not the reference's Q^, but Re(fft3D^-1(.)) must be the oracle's Q (src/collisions.c:212-221)."""
    L = _lib()
    dp = C.POINTER(C.c_double)
    L.emul_half0d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
    o = orc.Oracle(N, 5.0, 0)
    n3 = N ** 3
    W = orc.synthetic_weights(N)
    Wh = np.empty_like(W)
    R = _mirror_rule_lib()
    assert R.mirror_emul_fold(N, W.ctypes.data_as(dp), 1, Wh.ctypes.data_as(dp)) == 0
    f = seeded_f(o.v, 41, noise=0.3)
    z = np.arange(N)

    def parity(x):   # spectrum of a real field in parity-split lines [x][y][z & 1][z >> 1] (LAY_PARITY, csrc/internal.h)
        F = o.fft3d(np.asarray(x).astype(complex)).reshape(N, N, N)
        out = np.empty((N, N, N), dtype=complex)
        out[:, :, (z & 1) * (N // 2) + (z >> 1)] = F
        return np.ascontiguousarray(out)

    if npairs == 1:
        xiA = dfA = parity(f)
        xiB = dfB = xiA
        want = o.compute_q(W, f, f)
    else:                       # one species: M_i = M_j = M, g_i = g_j = f - M  (csrc/capi.cu: compute_q_maxpreserve_dev)
        M, _ = o.find_maxwellian(f)
        g = f - M
        xiA, dfA = parity(g), parity(f)      # g_j^[xi] f^[zeta - xi]
        xiB, dfB = parity(M), parity(g)      # M_j^[xi] g_i^[zeta - xi]
        want = o.compute_q_maxpreserve(W, f, f)
    parts = np.full((nsplit + 1) * n3, np.nan + 1j * np.nan, dtype=complex)
    out = str(tmp_path / "parts.npy")
    pd = lambda a: a.view(np.float64).ctypes.data_as(dp)  # noqa: E731

    def child():
        rc = L.emul_half0d(N, nsplit, packed, npairs, Wh.ctypes.data_as(dp), pd(xiA), pd(dfA), pd(xiB), pd(dfB), pd(parts))
        if rc == 0:
            np.save(out, parts)
        os._exit(rc)

    proc = mp.get_context("fork").Process(target=child)
    proc.start()
    proc.join(900)
    assert proc.exitcode == 0, "kernel emulation failed or deadlocked (exit code %s)" % proc.exitcode
    parts = np.load(out).reshape(nsplit + 1, n3)
    assert not np.isnan(parts.view(np.float64)).any()
    S = parts.sum(axis=0)
    assert relmax(np.real(o.fft3d(S, invert=True)), want) < 1e-12
