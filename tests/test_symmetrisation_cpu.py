"""Host-side check of the identity the symmetrised weight stream relies on, with the oracle's
own convolution as the reference: for f == g,

    sum_xi W[zeta][xi] f^[xi] f^[sigma(xi)]  ==  sum_{xi_x in reps(zeta_x)} Ws[zeta][xi] f^[xi] f^[sigma(xi)],
    sigma(xi) = wrap(zeta + N/2 - xi),  Ws = W + W o sigma on planes xi_x < sigma_x, W on self-paired planes,

for ARBITRARY weights (no symmetry of W assumed).  The plane enumeration below restates sym_nrep / sym_rep of
spectralbte_b200/csrc/common.cuh and the rule of symmetrize_weights_kernel (csrc/qhat.cu)."""
import numpy as np
import pytest

from conftest import relmax
from oracle import oracle as orc


def wrap(i, N):
    return i + N if i < 0 else (i - N if i > N - 1 else i)


def reps(N, zx):
    a = (zx + N // 2) % N
    return list(range(0, a // 2 + 1)) + list(range(a + 1, (a + N) // 2 + 1))


@pytest.mark.parametrize("N", [6, 8])
def test_symmetrised_sum_equals_full_sum(N):
    o = orc.Oracle(N, 5.0, 0)
    n3 = N ** 3
    rng = np.random.default_rng(N)
    W = rng.standard_normal((n3, n3))                      # arbitrary, unsymmetric weights
    fh = rng.standard_normal(n3) + 1j * rng.standard_normal(n3)
    want = o.qhat(W.reshape(-1).copy(), fh, fh)           # src/collisions.c:127-165
    got = np.zeros(n3, dtype=complex)
    idx = lambda x, y, z: z + N * (y + N * x)  # noqa: E731
    visited_planes = 0
    for zx in range(N):
        rp = reps(N, zx)
        a = (zx + N // 2) % N
        assert len(rp) == a // 2 + 1 + (a + N) // 2 - a      # sym_nrep
        assert len(rp) in (N // 2, N // 2 + 1)
        visited_planes += len(rp)
        for zy in range(N):
            for zz in range(N):
                zeta = idx(zx, zy, zz)
                acc = 0.0
                for ex in rp:
                    X = wrap(zx + N // 2 - ex, N)
                    assert ex <= X                           # representatives: the smaller plane of each pair
                    for ey in range(N):
                        Y = wrap(zy + N // 2 - ey, N)
                        for ez in range(N):
                            Z = wrap(zz + N // 2 - ez, N)
                            xi, sig = idx(ex, ey, ez), idx(X, Y, Z)
                            ws = W[zeta, xi] + W[zeta, sig] if ex < X else W[zeta, xi]
                            acc += ws * fh[xi] * fh[sig]
                got[zeta] = acc
    assert relmax(got, want) < 1e-12
    # bytes actually streamed: 8 N^4 sum(nrep) instead of 8 N^6
    assert visited_planes == sum(((zx + N // 2) % N) // 2 + 1 + (((zx + N // 2) % N) + N) // 2 - ((zx + N // 2) % N)
                                 for zx in range(N))
    assert 0.5 < visited_planes / float(N * N) < 0.6 + 1.0 / N
