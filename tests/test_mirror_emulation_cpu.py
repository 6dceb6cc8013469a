"""The mirror-paired batched convolution (spectralbte_b200/csrc/mirror.cuh, opt-in SBTE_MIRROR=1) shares its
per-lane arithmetic, its column pairing and its symmetrisation rule with a host emulation built from the same
header (tests/emul/mirror_emul.cu -- test infrastructure, not part of the library).  Here the emulation runs one
cell tile by tile and step by step as the kernel does and must reproduce the oracle's convolution
(src/collisions.c:127-165) of a REAL distribution function to 1e-12, for arbitrary unsymmetric weights, with the
plain tensor and with the mirror-symmetrised one."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import relmax, seeded_f
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "mirror_emul.cu")
LIB = os.path.join(HERE, "emul", "libmirror_emul.so")
HDR = os.path.join(os.path.dirname(HERE), "spectralbte_b200", "csrc", "mirror.cuh")


def _emul():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--expt-relaxed-constexpr",
                        "-Xcompiler", "-fPIC", "-shared", "-o", LIB, SRC], check=True, capture_output=True)
    L = C.CDLL(LIB)
    dp = C.POINTER(C.c_double)
    L.mirror_emul_symmetrize.argtypes = [C.c_int, dp, dp]
    L.mirror_emul_qhat.argtypes = [C.c_int, dp, C.c_int, C.c_int, dp, C.c_double, C.c_double, dp]
    L.mirror_emul_fold.argtypes = [C.c_int, dp, C.c_int, dp]
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("N,L_v,rule", [(8, 5.0, 0), (8, 9.0, 1), (16, 9.0, 1), (12, 7.0, 1), (22, 9.0, 1)])
def test_mirror_emulation_matches_oracle(N, L_v, rule):
    L = _emul()
    o = orc.Oracle(N, L_v, rule)
    n3 = N ** 3
    if N <= 12:
        W = np.random.default_rng(N).standard_normal(n3 * n3)      # arbitrary, unsymmetric
    else:
        W = orc.synthetic_weights(N)
    f = seeded_f(o.v, 5, noise=0.3)
    F = np.ascontiguousarray(o.fft3d(f.astype(complex)))
    want = o.qhat(W, F, F)
    dv = o.v[1] - o.v[0]
    L_eta = 0.5 * N * (2.0 * np.pi / (N * dv))
    got = np.empty(n3, dtype=complex)
    assert L.mirror_emul_qhat(N, _p(W), 0, 0, _p(F.view(np.float64)), L_eta, L_v, _p(got.view(np.float64))) == 0
    assert not np.isnan(got.view(np.float64)).any()                # the pairing covers every zeta exactly
    assert relmax(got, want) < 1e-12
    Ws2 = np.empty_like(W)
    assert L.mirror_emul_symmetrize(N, _p(W), _p(Ws2)) == 0
    got2 = np.empty(n3, dtype=complex)
    assert L.mirror_emul_qhat(N, _p(Ws2), 1, 0, _p(F.view(np.float64)), L_eta, L_v, _p(got2.view(np.float64))) == 0
    assert not np.isnan(got2.view(np.float64)).any()
    assert relmax(got2, want) < 1e-12
    # folded tensors (mirror rows folded into their partners where the phase exponent is 0): not Q^ any more, but the
    # same Q = Re(fft3D^-1(.)), src/collisions.c:212-221
    Qwant = o.compute_q(W, f, f)
    for sym in (0, 1):
        Wh = np.empty_like(W)
        assert L.mirror_emul_fold(N, _p(W), sym, _p(Wh)) == 0
        S = np.empty(n3, dtype=complex)
        assert L.mirror_emul_qhat(N, _p(Wh), sym, 1, _p(F.view(np.float64)), L_eta, L_v, _p(S.view(np.float64))) == 0
        assert not np.isnan(S.view(np.float64)).any()
        assert relmax(np.real(o.fft3d(S, invert=True)), Qwant) < 1e-12, sym
        assert relmax(S, want) > 1e-6          # it really is a different spectrum


def test_mirror_tiles_pair_every_column_once():
    """Restates build_mirror_tiles (mirror.cuh): every zeta (x,y) column is an A column, a B column or unpaired, once."""
    for N, pairs in ((8, 2), (16, 4)):
        nu = lambda i: (N - i) % N  # noqa: E731
        seen = set()
        for zx in range(N // 2 + 1):
            if zx in (0, N // 2):
                groups = [(list(range(1, N // 2)), True), ([0], False), ([N // 2], False)]
            else:
                groups = [(list(range(N)), True)]
            for cols, paired in groups:
                for zy in cols:
                    assert (zx, zy) not in seen
                    seen.add((zx, zy))
                    if paired:
                        b = (nu(zx), nu(zy))
                        assert b not in seen and b != (zx, zy)
                        seen.add(b)
        assert len(seen) == N * N
