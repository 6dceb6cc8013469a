"""Host-side check of an identity the batched (FP64-bound) convolution can exploit next, with
the oracle's own transform and convolution as the reference.  For REAL f the reference's spectrum
(src/collisions.c:232-283) satisfies, with nu(i) = (N - i) mod N per dimension and z(idx) = number of zero components,

    f^[nu(idx)] = theta^z(idx) conj(f^[idx]),   theta = exp(-2i L_eta L_v)

(eta_0 = -L_eta has no mirror node on the grid; the transform is quasi-periodic there).  The convolution index
sigma_zeta(xi) = wrap(zeta + N/2 - xi) (src/collisions.c:141-160) commutes with nu, so the complex product
P = f^[xi] f^[sigma_zeta(xi)] of row zeta is, up to conjugation and that phase, the product row nu(zeta) needs at
nu(xi): two weights can share one complex product (8 instead of 12 FP64 instructions) wherever no index component is
zero -- for ARBITRARY real weights (no symmetry of W assumed)."""
import numpy as np
import pytest

from conftest import relmax, seeded_f
from oracle import oracle as orc


@pytest.mark.parametrize("N,L_v,rule", [(8, 5.0, 0), (6, 7.0, 1)])
def test_mirror_rows_share_the_complex_product(N, L_v, rule):
    o = orc.Oracle(N, L_v, rule)
    n3 = N ** 3
    rng = np.random.default_rng(N)
    W = rng.standard_normal((n3, n3))                      # arbitrary, unsymmetric weights
    f = seeded_f(o.v, 3, noise=0.3)                        # real distribution function
    F = o.fft3d(f.astype(complex))
    dv = o.v[1] - o.v[0]
    L_eta = 0.5 * N * (2.0 * np.pi / (N * dv))             # src/initializer.c:66-82 (even N)
    theta = np.exp(-2j * L_eta * L_v)
    I, J, K = np.meshgrid(range(N), range(N), range(N), indexing="ij")
    nu = lambda a: (N - a) % N  # noqa: E731
    flat = lambda x, y, z: z + N * (y + N * x)  # noqa: E731
    zeros = (I == 0).astype(int) + (J == 0) + (K == 0)
    F3 = F.reshape(N, N, N)
    assert relmax(F3[nu(I), nu(J), nu(K)], theta ** zeros * np.conj(F3)) < 1e-12

    want = o.qhat(W.reshape(-1).copy(), F, F)              # src/collisions.c:127-165
    got = np.zeros(n3, dtype=complex)
    xi = flat(I, J, K).reshape(-1)
    nuxi = flat(nu(I), nu(J), nu(K)).reshape(-1)
    done = np.zeros(n3, dtype=bool)
    shared_rows, clean = 0, 0
    for zx in range(N):
        for zy in range(N):
            for zz in range(N):
                ze, zen = flat(zx, zy, zz), flat(nu(zx), nu(zy), nu(zz))
                if done[ze]:
                    continue
                SX, SY, SZ = (zx + N // 2 - I) % N, (zy + N // 2 - J) % N, (zz + N // 2 - K) % N
                sig = flat(SX, SY, SZ).reshape(-1)
                P = F[xi] * F[sig]                          # one complex product per (zeta, xi) ...
                got[ze] = (W[ze, xi] * P).sum()
                done[ze] = True
                if zen != ze:                               # ... serves the mirror row as well
                    m = zeros.reshape(-1) + ((SX == 0).astype(int) + (SY == 0) + (SZ == 0)).reshape(-1)
                    got[zen] = (W[zen, nuxi] * theta ** m * np.conj(P)).sum()
                    done[zen] = True
                    shared_rows += 1
                    clean += int((m == 0).sum())
    assert done.all()
    assert relmax(got, want) < 1e-12
    # all rows but the 8 self-mirrored ones pair up; the phase-free share of their pairs is about ((N-2)/N)^3
    assert shared_rows == (n3 - 8) // 2
    assert abs(clean / float(shared_rows * n3) - ((N - 2.0) / N) ** 3) < 0.08
