"""Host-side check of the next algorithmic step, with the oracle as the reference.

The reference keeps only the REAL part of the inverse transform, Q = Re(fft3D^-1(Q^)) (src/collisions.c:212-221), so of
Q^ only its Hermitian part matters.  With nu(i) = (N - i) mod N, omega = the trapezoid weights the inverse transform
applies in eta space (src/collisions.c:246-263), theta = exp(-2i L_eta L_v) and z(idx) = number of zero components,

    Q = Re fft3D^-1(S),   S[zeta] = Q^[zeta] + (omega[nu zeta] / omega[zeta]) theta^z(zeta) conj(Q^[nu zeta]),  S[nu zeta] = 0

for one zeta of every mirror pair, and since row nu(zeta) needs the conjugate of the products row zeta forms
(tests/test_hermitian_sharing_cpu.py),

    S[zeta] = sum_xi ( W[zeta][xi] + rho theta^e W[nu zeta][nu xi] ) g^[xi] f^[zeta - xi],   e = z(zeta) - z(xi) - z(zeta - xi),

i.e. ONE weight -- real wherever e = 0 -- per product and HALF of the zeta rows: half the weight bytes (0D, HBM-bound)
and half the FP64 work (1D) for arbitrary real W and real f, g.  Only Q is reproduced, not the reference's Q^ itself."""
import numpy as np
import pytest

from conftest import relmax, seeded_f
from oracle import oracle as orc


@pytest.mark.parametrize("N,L_v,rule", [(8, 5.0, 0), (6, 7.0, 1)])
def test_half_of_the_rows_with_combined_weights_reproduce_q(N, L_v, rule):
    o = orc.Oracle(N, L_v, rule)
    n3 = N ** 3
    W = np.random.default_rng(N + 1).standard_normal((n3, n3))        # arbitrary, unsymmetric weights
    f, g = seeded_f(o.v, 3, noise=0.3), seeded_f(o.v, 9, noise=0.3)
    dv = o.v[1] - o.v[0]
    L_eta = 0.5 * N * (2.0 * np.pi / (N * dv))
    theta = np.exp(-2j * L_eta * L_v)
    I, J, K = np.meshgrid(range(N), range(N), range(N), indexing="ij")
    nu = lambda a: (N - a) % N  # noqa: E731
    flat = lambda x, y, z: z + N * (y + N * x)  # noqa: E731
    wt = np.ones(N)
    wt[0] = wt[-1] = 0.5
    om = (wt[I] * wt[J] * wt[K]).reshape(-1)
    zc = ((I == 0).astype(int) + (J == 0) + (K == 0)).reshape(-1)
    xi, nuxi = flat(I, J, K).reshape(-1), flat(nu(I), nu(J), nu(K)).reshape(-1)
    for ff, gg in ((f, f), (f, g)):                                   # ComputeQ(f, f) and the two-species form ComputeQ(f, g)
        Fh, Gh = o.fft3d(ff.astype(complex)), o.fft3d(gg.astype(complex))
        Qref = o.compute_q(W.reshape(-1).copy(), ff, gg)
        S = np.zeros(n3, dtype=complex)
        done = np.zeros(n3, dtype=bool)
        rows, real_w, all_w = 0, 0, 0
        for zx in range(N):
            for zy in range(N):
                for zz in range(N):
                    ze, zen = flat(zx, zy, zz), flat(nu(zx), nu(zy), nu(zz))
                    if done[ze]:
                        continue
                    done[ze] = done[zen] = True
                    rows += 1
                    SX, SY, SZ = (zx + N // 2 - I) % N, (zy + N // 2 - J) % N, (zz + N // 2 - K) % N
                    P = Gh[xi] * Fh[flat(SX, SY, SZ).reshape(-1)]     # g^[xi] f^[zeta - xi]   (src/collisions.c:162)
                    if zen == ze:
                        S[ze] = (W[ze, xi] * P).sum()
                        continue
                    e = zc[ze] - zc - ((SX == 0).astype(int) + (SY == 0) + (SZ == 0)).reshape(-1)
                    S[ze] = ((W[ze, xi] + (om[zen] / om[ze]) * theta ** e * W[zen, nuxi]) * P).sum()
                    real_w += int((e == 0).sum())
                    all_w += n3
        Q = np.real(o.fft3d(S, invert=True))
        assert relmax(Q, Qref) < 1e-12
        assert rows == (n3 + 8) // 2                                  # the 8 self-mirrored rows stay as they are
        assert 0.3 < real_w / all_w < ((N - 1.0) / N) ** 3            # the combined weight is real wherever e = 0


@pytest.mark.parametrize("N,L_v,rule", [(8, 5.0, 0)])
def test_partial_fold_with_real_weights_is_exact_with_the_unchanged_formula(N, L_v, rule):
    """The variant planned for the 0D stream kernel: fold an entry into its mirror row only where the phase exponent is 0
    and the z components are regular; every other entry stays in its own row.  The folded tensor Wh is REAL and
    Q = Re fft3D^-1( sum_xi Wh[zeta][xi] g^[xi] f^[zeta - xi] ) -- the kernels' formula, unchanged -- is exact; the mirror
    rows keep only a sparse set of entries (what they may skip is the speed-up)."""
    o = orc.Oracle(N, L_v, rule)
    n3 = N ** 3
    W = np.random.default_rng(5).standard_normal((n3, n3))
    f = seeded_f(o.v, 3, noise=0.3)
    I, J, K = np.meshgrid(range(N), range(N), range(N), indexing="ij")
    nu = lambda a: (N - a) % N  # noqa: E731
    flat = lambda x, y, z: z + N * (y + N * x)  # noqa: E731
    wt = np.ones(N)
    wt[0] = wt[-1] = 0.5
    om = (wt[I] * wt[J] * wt[K]).reshape(-1)
    xi, nuxi = flat(I, J, K).reshape(-1), flat(nu(I), nu(J), nu(K)).reshape(-1)
    Wh = W.copy()
    kept_in_mirror_rows, mirror_entries = 0, 0
    done = np.zeros(n3, dtype=bool)
    for zx in range(N):
        for zy in range(N):
            for zz in range(N):
                ze, zen = flat(zx, zy, zz), flat(nu(zx), nu(zy), nu(zz))
                if done[ze] or zen == ze:
                    done[ze] = True
                    continue
                done[ze] = done[zen] = True
                SX, SY, SZ = (zx + N // 2 - I) % N, (zy + N // 2 - J) % N, (zz + N // 2 - K) % N
                ex = int(zx == 0) - (I == 0).astype(int) - (SX == 0)
                ey = int(zy == 0) - (J == 0).astype(int) - (SY == 0)
                zreg = (zz != 0) & (K != 0) & (SZ != 0)
                fold = ((ex + ey == 0) & zreg).reshape(-1)
                # fold[xi] decides for the pair (zeta, xi) <-> (nu zeta, nu xi)
                Wh[ze, xi[fold]] = W[ze, xi[fold]] + (om[zen] / om[ze]) * W[zen, nuxi[fold]]
                Wh[zen, nuxi[fold]] = 0.0
                kept_in_mirror_rows += int((~fold).sum())
                mirror_entries += n3
    Fh = o.fft3d(f.astype(complex))
    S = o.qhat(Wh.reshape(-1).copy(), Fh, Fh)                       # the kernels' formula with the folded tensor
    Q = np.real(o.fft3d(S, invert=True))
    assert relmax(Q, o.compute_q(W.reshape(-1).copy(), f, f)) < 1e-12
    # what stays in the mirror rows: about 1 - ((N-2)/N)^3 (N-1)/N of their entries -- 63 % at this toy N = 8, 20 % at N = 32
    assert kept_in_mirror_rows / mirror_entries < 1.0 - ((N - 2.0) / N) ** 3 * (N - 1.0) / N + 0.05
