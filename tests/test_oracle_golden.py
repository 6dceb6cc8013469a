"""Pins the oracle (oracle/oracle.c) against the reference's own golden vectors and against
vectors produced by the reference's own code (tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest

from conftest import check_diff_two_sided, load_moments, relmax, seeded_f
from oracle import oracle as orc


def test_bkw8_golden_moments(W_bkw8):
    """tests/BKW8: 0D, N=8, L_v=5, lambda=0, dt=0.01, 100 RK2 steps, Init_field 2, output each step."""
    o = orc.Oracle(8, 5.0, 0)
    f = o.init_hom(2)
    rows = [np.concatenate([[0.0], o.row_0d(f)])]
    for t in range(100):
        o.step_0d(W_bkw8, f, 0.01, 1.0, 2)
        rows.append(np.concatenate([[0.01 * (t + 1)], o.row_0d(f)]))
    got = np.array(rows)
    want = load_moments("moments_BKW8.test.in")
    assert got.shape == want.shape == (101, 14)
    # the file holds 7 significant digits (%le): compare at the reference's own tolerance
    # after rounding our values the way %le does
    got_r = np.array([[float("%le" % x) for x in r] for r in got])
    bad = check_diff_two_sided(got_r, want)
    # column 2 (u_x) is 1e-16 round-off noise in the reference itself: abs tol only
    noise = np.abs(want[:, 2]).max()
    assert noise < 1e-14
    assert check_diff_two_sided(np.delete(got_r, 2, axis=1), np.delete(want, 2, axis=1)) == 0, bad
    assert np.abs(got[:, 2]).max() < 1e-13


def test_heat_transport_golden_moments(W_heat8):
    """tests/heat_transport: 1D, N=8, L_v=9, Kn=3.2, lambda=1, dt=1e-4, 10 Euler steps, order 1,
    Init_field 3 (diffuse walls T=1,2), 250 cells on [0,1]."""
    N, nX, order, ic, dt, Kn = 8, 250, 1, 3, 1e-4, 3.2
    o = orc.Oracle(N, 9.0, 1)
    nX_, x, dx = orc.make_mesh([250], [1.0], order)
    assert nX_ == nX
    f = o.init_inhom(ic, nX, order)
    fc, f1, ft = np.zeros_like(f), np.zeros_like(f), np.zeros_like(f)

    def dump(t):
        return [np.concatenate([[t, x[l]], o.row_1d(f[l])]) for l in range(order, nX + order)]

    rows = dump(0.0)
    for t in range(10):
        o.step_1d(W_heat8, nX, x, dx, dt, Kn, order, ic, f, fc, f1, ft)
        rows += dump(dt * (t + 1))
    got = np.array(rows)
    want = load_moments("moments_heat_transport.test.in")
    assert got.shape == want.shape == (2750, 6)
    got_r = np.array([[float("%le" % x) for x in r] for r in got])
    # u_x (column 3) starts as 1e-15 noise; everything else at the reference tolerance
    assert check_diff_two_sided(np.delete(got_r, 3, axis=1), np.delete(want, 3, axis=1)) == 0
    big = np.abs(want[:, 3]) > 1e-9
    assert check_diff_two_sided(got_r[big, 3], want[big, 3]) == 0
    assert np.abs(got[~big, 3] - want[~big, 3]).max() < 1e-12


CASES = [("n8_l0", 8, 5.0, 0, "bkw"), ("n8_l1", 8, 9.0, 1, "heat"),
         ("n12_syn", 12, 6.0, 0, "syn"), ("n16_syn", 16, 5.0, 0, "syn")]


@pytest.mark.parametrize("tag,N,L_v,rule,wkind", CASES)
def test_hot_path_vs_reference_vectors(ref_vectors, W_bkw8, W_heat8, tag, N, L_v, rule, wkind):
    o = orc.Oracle(N, L_v, rule)
    W = {"bkw": W_bkw8, "heat": W_heat8}.get(wkind)
    if W is None:
        W = orc.synthetic_weights(N)
    f, g = seeded_f(o.v, 11), seeded_f(o.v, 12)
    rng = np.random.default_rng(5)
    z = rng.standard_normal(o.n3) + 1j * rng.standard_normal(o.n3)
    V = ref_vectors
    assert relmax(o.fft3d(z, False), V[f"{tag}_fft_fwd"]) < 1e-14
    assert relmax(o.fft3d(z, True), V[f"{tag}_fft_inv"]) < 1e-14
    Qff = o.compute_q(W, f, f)
    assert relmax(Qff, V[f"{tag}_Q_ff"]) < 1e-13
    assert relmax(o.compute_q(W, f, g), V[f"{tag}_Q_fg"]) < 1e-13
    Qmp = o.compute_q_maxpreserve(W, f, f)
    assert relmax(Qmp, V[f"{tag}_Qmp_ff"]) < 1e-13
    assert relmax(o.compute_q_maxpreserve(W, f, g), V[f"{tag}_Qmp_fg"]) < 1e-13
    assert relmax(o.conserve(V[f"{tag}_Q_ff"]), V[f"{tag}_cons_Q_ff"]) < 1e-14
    assert relmax(o.conserve(V[f"{tag}_Qmp_ff"]), V[f"{tag}_cons_Qmp_ff"]) < 1e-14
    rho = o.density(f)
    u = o.bulk_velocity(f, rho)
    T = o.temperature(f, u, rho)
    e = o.energy(f)
    got = np.array([rho, u[0], u[1], u[2], T, e[0], e[1]])
    np.testing.assert_allclose(got, V[f"{tag}_moments_f"], rtol=1e-14, atol=1e-16)
    # conservation achieved by the projection: |C Q| after vs before
    before = np.abs(o.moment_functionals(V[f"{tag}_Q_ff"])).max()
    after = np.abs(o.moment_functionals(o.conserve(V[f"{tag}_Q_ff"]))).max()
    assert after < 1e-13 * max(1.0, before) and after < 1e-13


@pytest.mark.parametrize("tag,N,L_v,nX,ic,dt", [("tr_ic3", 8, 9.0, 12, 3, 1e-3), ("tr_ic6", 8, 9.0, 12, 6, 1e-3),
                                                 ("tr_ic0", 6, 7.0, 10, 0, 2e-3), ("tr_ic5", 8, 9.0, 12, 5, 1e-3),
                                                 ("tr_ic1", 8, 9.0, 12, 1, 1e-3)])
def test_transport_vs_reference_vectors(ref_vectors, tag, N, L_v, nX, ic, dt):
    o = orc.Oracle(N, L_v, 1)
    V = ref_vectors
    for order in (1, 2):
        _, x, dx = orc.make_mesh([nX // 2, nX - nX // 2], [0.4, 0.6], order)
        f = np.zeros((nX + 2 * order, o.n3))
        f[order:nX + order] = V[f"{tag}_o{order}_in"]
        fc = o.upwind_one(nX, x, dx, dt, ic, f) if order == 1 else o.advect_two(nX, x, dx, dt, ic, f)
        want = V[f"{tag}_o{order}_out"]
        assert np.array_equal(fc[order:nX + order], want) or relmax(fc[order:nX + order], want) < 1e-15
    fin = seeded_f(o.v, 3)
    for bdry, TW in ((0, 1.0), (1, 2.0)):
        res = o.diffuse_bc(fin, np.zeros(o.n3), TW, bdry)
        assert relmax(res, V[f"{tag}_bc{bdry}"]) < 1e-15


def test_dft_against_numpy_pocketfft():
    """Independent check of the dense DFT both the oracle and the FFTW shim rely on."""
    for N in (6, 8, 12, 16):
        o = orc.Oracle(N, 5.0, 0)
        rng = np.random.default_rng(N)
        z = rng.standard_normal((N, N, N)) + 1j * rng.standard_normal((N, N, N))
        assert relmax(o.dft3(z, -1).reshape(N, N, N), np.fft.fftn(z)) < 1e-14
        assert relmax(o.dft3(z, +1).reshape(N, N, N), np.fft.ifftn(z) * N ** 3) < 1e-14


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (no /root/reference here)")
def test_live_reference_matches_committed_vectors(ref_vectors, W_bkw8):
    """Where the reference can be compiled, the committed vectors must be reproducible."""
    R = orc.Reference(8, 5.0, 0)
    o = orc.Oracle(8, 5.0, 0)
    f = seeded_f(o.v, 11)
    Q = R.compute_q(R.rows(W_bkw8), f, f)
    assert np.array_equal(Q, ref_vectors["n8_l0_Q_ff"])
