"""GPU parity tests proper: the CUDA path, called through the C ABI (spectralbte_b200 -> ctypes ->
libsbte_b200.so), against the CPU oracle on the same seeded inputs, against the reference's golden
files, and through size-independent properties at full size.  Tolerances follow BASELINE.json:
Q^ relative 1e-12 (normwise), conservation 1e-13, golden moments abs 1e-14 or rel 1e-6 two-sided."""
import os

import numpy as np
import pytest

from conftest import check_diff_two_sided, load_moments, relmax, seeded_f
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL_QHAT = 1e-12


def _sbte():
    import spectralbte_b200 as sb
    return sb


@pytest.fixture(scope="module")
def sb():
    return _sbte()


# ---------------------------------------------------------------- transforms
@pytest.mark.parametrize("N,rule", [(6, 0), (8, 0), (8, 1), (12, 0), (16, 1), (22, 1), (24, 0), (32, 0)])
def test_fft3d_matches_oracle(sb, N, rule):
    o = orc.Oracle(N, 7.0, rule)
    c = sb.Collisions(N, 7.0, inhomogeneous=bool(rule))
    assert np.array_equal(c.v, o.v) and np.array_equal(c.eta, o.eta)
    rng = np.random.default_rng(N)
    for inv in (False, True):
        z = rng.standard_normal((3, o.n3)) + 1j * rng.standard_normal((3, o.n3))
        got = c.fft3D(z, inv).reshape(3, -1)
        for b in range(3):
            assert relmax(got[b], o.fft3d(z[b], inv)) < 1e-13


@pytest.mark.parametrize("N", [8, 12, 16])
def test_fft3d_whole_cell_kernel_matches_oracle(sb, N):
    """batches >= 8 take the one-launch whole-cell transform (radix-2 in registers for N = 8, 16)."""
    o = orc.Oracle(N, 7.0, 1)
    c = sb.Collisions(N, 7.0, inhomogeneous=True)
    rng = np.random.default_rng(N + 100)
    for inv in (False, True):
        z = rng.standard_normal((11, o.n3)) + 1j * rng.standard_normal((11, o.n3))
        got = c.fft3D(z, inv).reshape(11, -1)
        small = c.fft3D(z[:3], inv).reshape(3, -1)      # plane-parallel pair of kernels
        for b in range(11):
            assert relmax(got[b], o.fft3d(z[b], inv)) < 1e-13
        assert relmax(got[:3], small) < 1e-13


@pytest.mark.parametrize("N,batch,rule", [(16, 1, 0), (32, 1, 0), (16, 12, 1), (24, 3, 1), (22, 9, 1)])
def test_fft3d_matches_cufft(sb, N, batch, rule):
    """cuFFT (torch.fft on the device) as a second, independent validation of the hand-written transforms -- the cluster
    kernel (batch 1, N = 16 / 32), the whole-cell kernel (N = 16, batch >= 8) and the plane-parallel pair: fft3D is a
    trapezoid-weighted, shifted DFT (src/collisions.c:232-283),
        out[a] = post[a] * sum_b exp(-+ 2 pi i a.b / N) * pre[b] * w[b] * in[b] * (2 pi)^(-3/2) delta^3,
    pre[b] = exp(s i (b_x+b_y+b_z) L_start delta), post[a] = exp(s i L_end (g_ax+g_ay+g_az)), s = +1 forward, -1 inverse."""
    import torch
    o = orc.Oracle(N, 7.0, rule)
    c = sb.Collisions(N, 7.0, inhomogeneous=bool(rule))
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(7 * N + batch)
    dv, deta = c.v[1] - c.v[0], c.eta[1] - c.eta[0]
    L_v, L_eta = -c.v[0], -c.eta[0]
    wt = np.ones(N)
    wt[0] = wt[-1] = 0.5
    idx = np.arange(N)
    for inv in (False, True):
        z = rng.standard_normal((batch, N, N, N)) + 1j * rng.standard_normal((batch, N, N, N))
        got = c.fft3D(z.reshape(batch, -1), inv).reshape(batch, N, N, N)
        sign, delta, L_start, L_end, grid = (-1.0, deta, L_v, L_eta, c.v) if inv else (1.0, dv, L_eta, L_v, c.eta)
        s3 = idx[:, None, None] + idx[None, :, None] + idx[None, None, :]
        pre = np.exp(1j * sign * s3 * L_start * delta)
        post = np.exp(1j * sign * L_end * (grid[:, None, None] + grid[None, :, None] + grid[None, None, :]))
        w3 = wt[:, None, None] * wt[None, :, None] * wt[None, None, :]
        x = torch.from_numpy(z * (pre * w3 * (2.0 * np.pi) ** -1.5 * delta ** 3)).to(dev)
        y = torch.fft.ifftn(x, dim=(1, 2, 3), norm="forward") if inv else torch.fft.fftn(x, dim=(1, 2, 3))
        want = y.cpu().numpy() * post
        assert relmax(got, want) < 1e-13
        assert relmax(want[0].reshape(-1), o.fft3d(z[0].reshape(-1), inv)) < 1e-13   # and cuFFT agrees with the oracle


# ---------------------------------------------------------------- the convolution
def _weights(name, N, W_bkw8, W_heat8):
    if name == "bkw":
        return W_bkw8
    if name == "heat":
        return W_heat8
    return orc.synthetic_weights(N)


@pytest.mark.parametrize("N,L_v,rule,wname,k2", [
    (8, 5.0, 0, "bkw", 1), (8, 9.0, 1, "heat", 1), (6, 4.0, 0, "syn", 1), (12, 6.0, 0, "syn", 1),
    (16, 5.0, 0, "syn", 1), (16, 5.0, 0, "syn", 2), (16, 5.0, 0, "syn", 4), (24, 9.0, 1, "syn", 2)])
def test_qhat_matches_oracle(sb, W_bkw8, W_heat8, N, L_v, rule, wname, k2):
    o = orc.Oracle(N, L_v, rule)
    W = _weights(wname, N, W_bkw8, W_heat8)
    c = sb.Collisions(N, L_v, inhomogeneous=bool(rule))
    c.set_weights(W)
    f, g = seeded_f(o.v, 11), seeded_f(o.v, 12)
    for ff, gg in ((f, f), (f, g)):
        _, want = o.compute_q(W, ff, gg, want_qhat=True)
        got = c.Qhat(ff, None if gg is ff else gg, k2=k2)
        assert relmax(got, want) < TOL_QHAT


def test_qhat_rows_upload_equals_contiguous_upload(sb, W_bkw8):
    o = orc.Oracle(8, 5.0, 0)
    f = seeded_f(o.v, 11)
    a = sb.Collisions(8, 5.0)
    a.set_weights(W_bkw8)
    b = sb.Collisions(8, 5.0)
    b.set_weights_rows(W_bkw8)
    assert np.array_equal(a.Qhat(f), b.Qhat(f))
    assert np.array_equal(b.weights_to_host(), W_bkw8)


def test_qhat_n32_full_size_properties(sb):
    """BASELINE config 3 size (N=32, 1.07e9 weights, 8.6 GB): stream kernel vs the independent generic
    kernel on every row, sampled rows vs a numpy evaluation of the reference formula, linearity."""
    N, L_v = 32, 5.0
    o = orc.Oracle(N, L_v, 0)
    c = sb.Collisions(N, L_v)
    c.synthetic_weights(20261017)
    f, g = seeded_f(o.v, 11), seeded_f(o.v, 12)
    assert relmax(c.Qhat(f, None, k2=sb.K2_STREAM), c.Qhat(f, None, k2=sb.K2_GENERIC)) < TOL_QHAT
    q_stream = c.Qhat(f, g, k2=sb.K2_STREAM)
    q_deep = c.Qhat(f, g, k2=sb.K2_STREAM_DEEP)
    q_generic = c.Qhat(f, g, k2=sb.K2_GENERIC)
    assert relmax(q_stream, q_generic) < TOL_QHAT
    assert relmax(q_deep, q_generic) < TOL_QHAT
    # sampled rows against the formula of src/collisions.c:127-165 evaluated with numpy
    fh = o.fft3d(f.astype(np.complex128)).reshape(N, N, N)
    gh = o.fft3d(g.astype(np.complex128)).reshape(-1)
    n3 = N ** 3
    ar = np.arange(N)
    rng = np.random.default_rng(0)
    Wrow = np.empty(n3)
    scale = np.abs(q_generic).max()
    for zeta in rng.integers(0, n3, 24):
        zx, zy, zz = zeta // (N * N), (zeta // N) % N, zeta % N
        wrap = lambda z: np.where(z < 0, z + N, np.where(z > N - 1, z - N, z))  # noqa: E731
        X, Y, Z = wrap(zx + N // 2 - ar), wrap(zy + N // 2 - ar), wrap(zz + N // 2 - ar)
        fsel = fh[np.ix_(X, Y, Z)].reshape(-1)
        sb._lib.check(c.L.sbte_d2h(c.h, Wrow.ctypes.data, c.L.sbte_weights_device(c.h) + int(zeta) * n3 * 8, n3 * 8))
        want = np.sum(Wrow * gh * fsel)
        assert abs(q_stream[zeta] - want) < TOL_QHAT * scale
    # bilinearity: Q^(a f1 + b f2, g) = a Q^(f1, g) + b Q^(f2, g)
    f2 = seeded_f(o.v, 13)
    lhs = c.Qhat(2.0 * f - 0.5 * f2, g, k2=sb.K2_STREAM)
    rhs = 2.0 * q_stream - 0.5 * c.Qhat(f2, g, k2=sb.K2_STREAM)
    assert relmax(lhs, rhs) < TOL_QHAT


def test_synthetic_weights_device_equals_host_generator(sb):
    c = sb.Collisions(8, 5.0)
    c.synthetic_weights(20261017)
    assert np.array_equal(c.weights_to_host(), orc.synthetic_weights(8, 20261017))


# ---------------------------------------------------------------- ComputeQ / maxPreserve / conserve
@pytest.mark.parametrize("N,L_v,rule,wname", [(8, 5.0, 0, "bkw"), (8, 9.0, 1, "heat"), (12, 6.0, 0, "syn"),
                                              (16, 5.0, 0, "syn")])
def test_computeq_and_maxpreserve_match_oracle(sb, ref_vectors, W_bkw8, W_heat8, N, L_v, rule, wname):
    o = orc.Oracle(N, L_v, rule)
    W = _weights(wname, N, W_bkw8, W_heat8)
    c = sb.Collisions(N, L_v, inhomogeneous=bool(rule))
    c.set_weights(W)
    f, g = seeded_f(o.v, 11), seeded_f(o.v, 12)
    assert relmax(c.ComputeQ(f), o.compute_q(W, f, f)) < TOL_QHAT
    assert relmax(c.ComputeQ(f, g), o.compute_q(W, f, g)) < TOL_QHAT
    assert relmax(c.ComputeQ_maxPreserve(f), o.compute_q_maxpreserve(W, f, f)) < TOL_QHAT
    assert relmax(c.ComputeQ_maxPreserve(f, g), o.compute_q_maxpreserve(W, f, g)) < TOL_QHAT
    # and against the vectors produced by the reference's own code
    tag = {(8, "bkw"): "n8_l0", (8, "heat"): "n8_l1", (12, "syn"): "n12_syn", (16, "syn"): "n16_syn"}[(N, wname)]
    assert relmax(c.ComputeQ(f), ref_vectors[f"{tag}_Q_ff"]) < TOL_QHAT
    assert relmax(c.ComputeQ_maxPreserve(f, g), ref_vectors[f"{tag}_Qmp_fg"]) < TOL_QHAT


@pytest.mark.parametrize("N", [8, 12, 16, 32])
def test_conserve_and_moments(sb, N):
    o = orc.Oracle(N, 6.0, 0)
    c = sb.Collisions(N, 6.0)
    rng = np.random.default_rng(N)
    Q = np.stack([seeded_f(o.v, s) * rng.standard_normal(o.n3) for s in (1, 2, 3)])
    got = c.conserveAllMoments(Q).reshape(3, -1)
    for b in range(3):
        assert relmax(got[b], o.conserve(Q[b])) < 1e-13
    # residual |C Q| after projection: round-off of the corrected field, i.e. ~eps * sum |C| |Q|
    # (these random fields are O(1); a physical Q is checked at 1e-13 absolute below)
    vx, vy, vz = np.meshgrid(o.v, o.v, o.v, indexing="ij")
    w = np.ones(N); w[0] = w[-1] = 0.5
    w3 = (w[:, None, None] * w[None, :, None] * w[None, None, :] * (o.v[1] - o.v[0]) ** 3).reshape(-1)
    e2 = (0.5 * (vx * vx + vy * vy + vz * vz)).reshape(-1)
    res = c.moment_functionals(got)
    for b in range(3):
        scale = float(np.sum(w3 * e2 * np.abs(got[b])))
        # the solve amplifies round-off by the conditioning of C C^T; the reference's own arithmetic
        # (oracle) is the yardstick for these O(1) random fields
        ref_res = np.abs(o.moment_functionals(o.conserve(Q[b]))).max()
        assert np.abs(res[b]).max() <= 4 * max(ref_res, 64 * np.finfo(float).eps * scale)
    f = np.stack([seeded_f(o.v, s) for s in (4, 5)])
    m = c.moments(f)
    for b in range(2):
        rho = o.density(f[b]); u = o.bulk_velocity(f[b], rho); T = o.temperature(f[b], u, rho); e = o.energy(f[b])
        want = np.array([rho, u[0], u[1], u[2], T, e[0], e[1], rho * T])
        np.testing.assert_allclose(m[b], want, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("N,L_v,wname", [(8, 5.0, "bkw"), (16, 5.0, "syn")])
def test_conservation_to_1e13_on_collision_output(sb, W_bkw8, W_heat8, N, L_v, wname):
    """BASELINE: mass/momentum/energy conserved to 1e-13 after projection, on a real Q = ComputeQ(f,f)."""
    o = orc.Oracle(N, L_v, 0)
    W = _weights(wname, N, W_bkw8, W_heat8)
    if wname == "syn":
        W = W * 1e-3
    c = sb.Collisions(N, L_v)
    c.set_weights(W)
    f = o.init_hom(2)
    Q = c.ComputeQ(f)
    raw = np.abs(c.moment_functionals(Q)).max()
    Qc = c.conserveAllMoments(Q)
    assert np.abs(c.moment_functionals(Qc)).max() < 1e-13
    assert raw > 1e-9  # the projection had something to remove
    assert relmax(Qc, o.conserve(o.compute_q(W, f, f))) < 1e-12


def test_capacity_regrowth_keeps_results(sb, W_bkw8):
    """single-cell call first, then a larger batch on the same context (scratch is reallocated)."""
    o = orc.Oracle(8, 5.0, 0)
    c = sb.Collisions(8, 5.0)
    c.set_weights(W_bkw8)
    f = o.init_hom(2)
    c.ComputeQ_maxPreserve(f)
    cells = np.stack([f * (1.0 + 0.01 * b) for b in range(40)])
    for k2 in (sb.K2_BATCH, sb.K2_GENERIC):
        Q = c.ComputeQ(cells, k2=k2).reshape(40, -1)
        for b in (0, 17, 39):
            assert relmax(Q[b], o.compute_q(W_bkw8, cells[b], cells[b])) < TOL_QHAT


# ---------------------------------------------------------------- the reference's own goldens
def test_bkw8_golden_on_gpu(sb, W_bkw8):
    """tests/BKW8 of the reference, run through sbte_step_0d with f resident on the device."""
    o = orc.Oracle(8, 5.0, 0)
    c = sb.Collisions(8, 5.0)
    c.set_weights(W_bkw8)
    f = c.array(o.n3).put(o.init_hom(2))
    rows = [np.concatenate([[0.0], c.row_0d(f)])]
    for t in range(100):
        c.step_0d(f, 0.01, 1.0, 2)
        rows.append(np.concatenate([[0.01 * (t + 1)], c.row_0d(f)]))
    got = np.array(rows)
    want = load_moments("moments_BKW8.test.in")
    got_r = np.array([[float("%le" % x) for x in r] for r in got])
    assert check_diff_two_sided(np.delete(got_r, 2, axis=1), np.delete(want, 2, axis=1)) == 0
    assert np.abs(got[:, 2]).max() < 1e-13
    # full-precision agreement with the oracle's own trajectory at the end
    fo = o.init_hom(2)
    for t in range(100):
        o.step_0d(W_bkw8, fo, 0.01, 1.0, 2)
    assert relmax(f.get(), fo) < 1e-11


def test_heat_transport_golden_on_gpu(sb, W_heat8):
    """tests/heat_transport of the reference through the device-resident slab (batched K2)."""
    N, nX, order, ic, dt, Kn = 8, 250, 1, 3, 1e-4, 3.2
    o = orc.Oracle(N, 9.0, 1)
    _, x, dx = orc.make_mesh([250], [1.0], order)
    c = sb.Collisions(N, 9.0, inhomogeneous=True)
    c.set_weights(W_heat8)
    s = sb.Slab(c, nX, order, x, dx, ic, dt)
    s.upload(o.init_inhom(ic, nX, order))

    def dump(t):
        m = s.moments()
        return [[t, x[l + order], m[l, 0], m[l, 1], m[l, 4], m[l, 7]] for l in range(nX)]

    rows = dump(0.0)
    for t in range(10):
        s.step(Kn)
        rows += dump(dt * (t + 1))
    got = np.array(rows)
    want = load_moments("moments_heat_transport.test.in")
    got_r = np.array([[float("%le" % v) for v in r] for r in got])
    assert check_diff_two_sided(np.delete(got_r, 3, axis=1), np.delete(want, 3, axis=1)) == 0
    big = np.abs(want[:, 3]) > 1e-9
    assert check_diff_two_sided(got_r[big, 3], want[big, 3]) == 0
    assert np.abs(got[~big, 3] - want[~big, 3]).max() < 1e-12


# ---------------------------------------------------------------- batched convolution (1D)
@pytest.mark.parametrize("N,cells,k2", [(8, 5, 3), (8, 37, 3), (16, 33, 3), (24, 33, 3), (8, 5, 1), (12, 3, 1),
                                        (12, 9, 0), (22, 34, 3), (6, 40, 3)])
def test_batched_computeq_matches_oracle(sb, N, cells, k2):
    o = orc.Oracle(N, 9.0, 1)
    W = orc.synthetic_weights(N)
    c = sb.Collisions(N, 9.0, inhomogeneous=True)
    c.set_weights(W)
    f = np.stack([seeded_f(o.v, 100 + b, noise=0.2) * (1.0 + 0.1 * b) for b in range(cells)])
    qh = c.Qhat(f, k2=k2).reshape(cells, -1)
    Q = c.ComputeQ(f, k2=k2).reshape(cells, -1)
    for b in range(cells):
        Qo, qo = o.compute_q(W, f[b], f[b], want_qhat=True)
        assert relmax(qh[b], qo) < TOL_QHAT
        assert relmax(Q[b], Qo) < TOL_QHAT


@pytest.mark.parametrize("cells", [12, 44, 80])
def test_split_tiles_of_the_remainder_group_n16(sb, cells):
    """N = 16: a last cell group with at most 16 live cells runs on split tiles (two zeta_y columns per warp) in a second
    launch.  Every cell against the generic kernel and three against the oracle; and the cells of the split group must
    carry the same BITS as when they sit in a full group on ordinary tiles (rank-count invariance: which tile shape
    serves a cell depends on how many cells the rank holds)."""
    N = 16
    o = orc.Oracle(N, 9.0, 1)
    c = sb.Collisions(N, 9.0, inhomogeneous=True)
    c.synthetic_weights(11)
    W = c.weights_to_host()
    full = -(-cells // 32) * 32
    f = np.stack([seeded_f(o.v, 300 + b, noise=0.2) * (1.0 + 0.01 * b) for b in range(full)])
    for sym in (True, False):
        c.set_symmetrize(sym)
        qh = c.Qhat(f[:cells], k2=sb.K2_BATCH).reshape(cells, -1)
        qg = c.Qhat(f[:cells], k2=sb.K2_GENERIC).reshape(cells, -1)
        for b in range(cells):
            assert relmax(qh[b], qg[b]) < TOL_QHAT, (sym, b)
        Qs = c.ComputeQ(f[:cells], k2=sb.K2_BATCH).reshape(cells, -1)
        Qf = c.ComputeQ(f, k2=sb.K2_BATCH).reshape(full, -1)          # the same cells inside a full last group
        assert np.array_equal(Qs, Qf[:cells]), sym
    for b in (0, cells - 1):
        assert relmax(Qs[b], o.compute_q(W, f[b], f[b])) < TOL_QHAT   # Qs: the unsymmetrised pass


_CUT_SNIPPET = """
import hashlib, sys
import numpy as np
sys.path.insert(0, %(root)r)
import spectralbte_b200 as sb
N, cells = %(N)d, %(cells)d
c = sb.Collisions(N, 9.0, inhomogeneous=True)
c.synthetic_weights(11)
v = np.linspace(-9.0, 9.0, N)
X, Y, Z = np.meshgrid(v, v, v, indexing="ij")
f = np.stack([np.exp(-((X - 0.1 * b) ** 2 + Y ** 2 + (Z + 0.05 * b) ** 2) / (1.0 + 0.01 * b)) for b in range(cells)])
Q = c.ComputeQ(f, k2=sb.K2_BATCH)
print("DIGEST", hashlib.sha256(np.ascontiguousarray(Q).tobytes()).hexdigest(), float(np.abs(Q).max()))
"""


@pytest.mark.parametrize("N,cells", [(16, 70), (16, 160), (20, 40), (24, 33)])
def test_cuts_inside_chunks_give_whole_chunk_bits(N, cells):
    """Line-ring kernels: a stream-K range may end inside a xi_x chunk, the running sum then passes from one CTA to the
    next through global memory (flag word per warp, release / acquire).  The same batch computed by a process that cuts
    at any step (SBTE_CHUNK_CUTS=0) and by one that cuts at whole chunks only (=1) must agree bit for bit -- on the
    hardware, where the hand-over is a real inter-CTA exchange (the CPU emulation runs the CTAs one after the other)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    for mode in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", _CUT_SNIPPET % dict(root=root, N=N, cells=cells)], capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, SBTE_CHUNK_CUTS=mode))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][-1].split()
        assert float(line[2]) > 0.0
        digests[mode] = line[1]
    assert digests["0"] == digests["1"]


@pytest.mark.parametrize("N,cells", [(22, 70), (20, 33), (22, 250)])
def test_line_ring_with_partial_row_blocks(sb, N, cells):
    """N = 20, 22: 8 zeta_y columns per CTA do not tile a zeta_x plane, so every third row-block is partly
    empty. All cells against the generic kernel (one CTA per (zeta, cell), no tiling), three cells against the
    oracle; several cell groups so that stream-K cuts tiles into partial sums."""
    o = orc.Oracle(N, 9.0, 1)
    c = sb.Collisions(N, 9.0, inhomogeneous=True)
    c.synthetic_weights(11)
    W = c.weights_to_host()
    f = np.stack([seeded_f(o.v, 300 + b, noise=0.2) * (1.0 + 0.01 * b) for b in range(cells)])
    for sym in (True, False):
        c.set_symmetrize(sym)
        qh = c.Qhat(f, k2=sb.K2_BATCH).reshape(cells, -1)
        qg = np.concatenate([c.Qhat(f[b0:b0 + 50], k2=sb.K2_GENERIC).reshape(-1, qh.shape[1])
                             for b0 in range(0, cells, 50)])
        for b in range(cells):
            assert relmax(qh[b], qg[b]) < TOL_QHAT, (sym, b)
    c.set_symmetrize(True)
    Q = c.ComputeQ(f, k2=sb.K2_BATCH).reshape(cells, -1)
    for b in (0, 31, cells - 1):
        Qo, qo = o.compute_q(W, f[b], f[b], want_qhat=True)
        assert relmax(qh[b], qo) < TOL_QHAT   # qh: the unsymmetrised pass of the loop above
        assert relmax(Q[b], Qo) < TOL_QHAT


# ---------------------------------------------------------------- transport
@pytest.mark.parametrize("N,L_v,nX,ic,dt", [(8, 9.0, 12, 3, 1e-3), (8, 9.0, 12, 6, 1e-3), (6, 7.0, 10, 0, 2e-3),
                                             (8, 9.0, 12, 1, 1e-3), (8, 9.0, 12, 5, 1e-3), (8, 9.0, 12, 2, 1e-3)])
def test_transport_matches_oracle(sb, N, L_v, nX, ic, dt):
    o = orc.Oracle(N, L_v, 1)
    c = sb.Collisions(N, L_v, inhomogeneous=True)
    rng = np.random.default_rng(77)
    for order in (1, 2):
        _, x, dx = orc.make_mesh([nX // 2, nX - nX // 2], [0.4, 0.6], order)
        f = o.init_inhom(ic, nX, order)
        f[order:nX + order] *= 1.0 + 0.2 * rng.standard_normal((nX, o.n3))
        s = sb.Slab(c, nX, order, x, dx, ic, dt)
        s.upload(f)
        s.advect(0)
        got = s.download_fconv()[order:nX + order]
        want = (o.upwind_one(nX, x, dx, dt, ic, f.copy()) if order == 1
                else o.advect_two(nX, x, dx, dt, ic, f.copy()))[order:nX + order]
        assert relmax(got, want) < 1e-14


@pytest.mark.parametrize("order,ic", [(1, 3), (2, 3), (2, 6), (1, 6), (1, 5), (2, 5), (1, 1), (2, 1), (2, 0)])
def test_1d_step_matches_oracle(sb, W_heat8, order, ic):
    """exec/boltz.c:264-353 end to end (advect + batched collide + advect), 3 steps, N=8."""
    N, nX, dt, Kn = 8, 14, 2e-3, 1.52
    o = orc.Oracle(N, 9.0, 1)
    _, x, dx = orc.make_mesh([nX], [0.7], order)
    c = sb.Collisions(N, 9.0, inhomogeneous=True)
    c.set_weights(W_heat8)
    s = sb.Slab(c, nX, order, x, dx, ic, dt)
    f = o.init_inhom(ic, nX, order)
    s.upload(f)
    fc, f1, ft = np.zeros_like(f), np.zeros_like(f), np.zeros_like(f)
    for _ in range(3):
        s.step(Kn)
        o.step_1d(W_heat8, nX, x, dx, dt, Kn, order, ic, f, fc, f1, ft)
    assert relmax(s.download()[order:nX + order], f[order:nX + order]) < 1e-11


@pytest.mark.parametrize("order,ic", [(1, 3), (2, 3), (1, 6), (2, 6)])
def test_peer_memory_halo_equals_single_rank(sb, W_heat8, order, ic):
    """The message-free halo (upwind kernels read the neighbour's boundary cells from peer memory, ordered by
    device-side counters) on two slabs driven from two streams reproduces the single-slab run bit for bit."""
    N, nX, dt, Kn = 8, 16, 2e-3, 1.52
    o = orc.Oracle(N, 9.0, 1)
    _, x, dx = orc.make_mesh([nX], [0.8], order)
    f0 = o.init_inhom(ic, nX, order)
    rng = np.random.default_rng(5)
    f0[order:nX + order] *= 1.0 + 0.1 * rng.standard_normal((nX, o.n3))
    ctxs = [sb.Collisions(N, 9.0, inhomogeneous=True) for _ in range(3)]
    for c in ctxs:
        c.set_weights(W_heat8)
    one = sb.Slab(ctxs[2], nX, order, x, dx, ic, dt)
    one.upload(f0)
    h = nX // 2
    parts = []
    for r in range(2):
        lo = r * h
        p = sb.Slab(ctxs[r], h, order, x[lo:lo + h + 2 * order].copy(), dx[lo:lo + h + 2 * order].copy(), ic, dt,
                    rank=r, nranks=2)
        p.upload(f0[lo:lo + h + 2 * order].copy())
        parts.append(p)
    parts[0].peer_attach(1, parts[1])
    parts[1].peer_attach(0, parts[0])
    if ic == 6 and order == 1:   # the ring closes
        parts[0].peer_attach(0, parts[1])
        parts[1].peer_attach(1, parts[0])
    for p in parts:
        p.set_peer_halo(True)
    for c in ctxs:
        c.sync()

    def advect(which):   # enqueue pass by pass on both streams so neither waits for work not yet submitted
        for stage in range(order):
            for p in parts:
                p.upwind_stage(which, stage)
        for p in parts:
            p.advect_finish(which)

    for _ in range(4):
        one.step(Kn)
        advect(0)
        for p in parts:
            p.collide(Kn)
        if order == 2:
            advect(1)
    want = one.download()[order:nX + order]
    got = np.concatenate([p.download()[order:h + order] for p in parts])
    assert np.array_equal(got, want)


def test_peer_halo_wait_times_out_without_killing_the_context(sb, W_heat8):
    """A neighbour that never arrives (it may be writing files, or gone): the waiting stencil kernels must not trap.
    They raise the slab's error word after the configured bound and stop waiting; the host learns it from the next
    synchronising call; the CUDA context -- and every other slab in it -- keeps working."""
    N, nX, order, ic, dt, Kn = 8, 16, 2, 6, 2e-3, 1.52
    o = orc.Oracle(N, 9.0, 1)
    _, x, dx = orc.make_mesh([nX], [0.8], order)
    f0 = o.init_inhom(ic, nX, order)
    ctxs = [sb.Collisions(N, 9.0, inhomogeneous=True) for _ in range(2)]
    for c in ctxs:
        c.set_weights(W_heat8)
    h = nX // 2
    parts = []
    for r in range(2):
        lo = r * h
        p = sb.Slab(ctxs[r], h, order, x[lo:lo + h + 2 * order].copy(), dx[lo:lo + h + 2 * order].copy(), ic, dt, rank=r, nranks=2)
        p.upload(f0[lo:lo + h + 2 * order].copy())
        parts.append(p)
    parts[0].peer_attach(1, parts[1])
    parts[1].peer_attach(0, parts[0])
    for p in parts:
        p.set_peer_halo(True)
        p.set_halo_timeout(0.05)
    parts[0].upwind_stage(0, 0)          # rank 1 never issues its pass: rank 0 waits 50 ms, then gives up
    ctxs[0].sync()
    assert parts[0].halo_state()[3] == 1
    with pytest.raises(sb._lib.SbteError, match="did not arrive"):
        parts[0].moments()
    # the context is alive: an ordinary single-rank slab on the same context steps and reproduces the oracle
    one = sb.Slab(ctxs[0], nX, order, x, dx, ic, dt)
    one.upload(f0)
    f = f0.copy()
    fc, f1, ft = np.zeros_like(f), np.zeros_like(f), np.zeros_like(f)
    one.step(Kn)
    o.step_1d(W_heat8, nX, x, dx, dt, Kn, order, ic, f, fc, f1, ft)
    assert relmax(one.download()[order:nX + order], f[order:nX + order]) < 1e-11


@pytest.mark.parametrize("order,ic,nX,h", [(1, 3, 16, 8), (2, 3, 16, 8), (1, 6, 16, 8), (2, 6, 16, 8), (2, 6, 80, 40), (1, 3, 70, 35),
                                            (2, 6, 16, 5), (1, 3, 44, 4)])
def test_two_rank_split_equals_single_rank(sb, W_heat8, order, ic, nX, h):
    """Rank-count invariance: two slabs exchanging halos through the regions
    reported by the library reproduce the single-slab run bit for bit -- also when the halves form a different number
    of 32-cell groups than the whole (nX = 80, 70: another stream-K schedule, same canonical summation order) and when
    one rank holds only h = 4 or 5 cells (the same transform kernels at every slab size)."""
    N, dt, Kn = 8, 2e-3, 1.52
    o = orc.Oracle(N, 9.0, 1)
    _, x, dx = orc.make_mesh([nX], [0.05 * nX], order)
    c = sb.Collisions(N, 9.0, inhomogeneous=True)
    c.set_weights(W_heat8)
    f0 = o.init_inhom(ic, nX, order)
    rng = np.random.default_rng(3)
    f0[order:nX + order] *= 1.0 + 0.1 * rng.standard_normal((nX, o.n3))
    one = sb.Slab(c, nX, order, x, dx, ic, dt)
    one.upload(f0)
    parts = []
    for r, (lo, n) in enumerate(((0, h), (h, nX - h))):     # rank 0 holds the first h cells
        xs, dxs = x[lo:lo + n + 2 * order].copy(), dx[lo:lo + n + 2 * order].copy()
        p = sb.Slab(c, n, order, xs, dxs, ic, dt, rank=r, nranks=2)
        p.upload(f0[lo:lo + n + 2 * order].copy())
        parts.append(p)

    def exchange(which, stage):
        # rank 0's right neighbour is rank 1 and vice versa; periodic wrap for IC 6 at order 1
        s0R, r0R, n = parts[0].halo_regions(which, stage, 1)
        s1L, r1L, _ = parts[1].halo_regions(which, stage, 0)
        sb._lib.check(c.L.sbte_d2d(c.h, r1L, s0R, n * 8))
        sb._lib.check(c.L.sbte_d2d(c.h, r0R, s1L, n * 8))
        if ic == 6 and order == 1:
            s0L, r0L, _ = parts[0].halo_regions(which, stage, 0)
            s1R, r1R, _ = parts[1].halo_regions(which, stage, 1)
            sb._lib.check(c.L.sbte_d2d(c.h, r1R, s0L, n * 8))
            sb._lib.check(c.L.sbte_d2d(c.h, r0L, s1R, n * 8))

    def advect(which):
        for stage in range(order):
            exchange(which, stage)
            for p in parts:
                p.upwind_stage(which, stage)
        for p in parts:
            p.advect_finish(which)

    for _ in range(3):
        one.step(Kn)
        advect(0)
        for p in parts:
            p.collide(Kn)
        if order == 2:
            advect(1)
    whole = one.download()[order:nX + order]
    split = np.concatenate([p.download()[order:p.nX + order] for p in parts])
    assert np.array_equal(whole, split)


# ---------------------------------------------------------------- drop-in symbols
def test_dropin_symbols_reproduce_reference_vectors(sb, ref_vectors, W_bkw8):
    """initialize_coll / ComputeQ / ComputeQ_maxPreserve / conserveAllMoments / fft3D called exactly as
    exec/boltz.c and src/initializer.c call them (host pointers, N^3 weight row pointers)."""
    import ctypes as C
    L = sb._lib.load()
    dp = C.POINTER(C.c_double)
    o = orc.Oracle(8, 5.0, 0)
    v, eta = o.v.copy(), o.eta.copy()
    L.initialize_coll(8, 5.0, v.ctypes.data_as(dp), eta.ctypes.data_as(dp))
    L.initialize_conservation_fast(8, v[1] - v[0], v.ctypes.data_as(dp))
    n3 = 512
    Wm = W_bkw8.reshape(n3, n3).copy()
    rows = (dp * n3)(*[C.cast(Wm.ctypes.data + i * n3 * 8, dp) for i in range(n3)])
    f = seeded_f(o.v, 11)
    Q = np.empty(n3)
    L.ComputeQ(f.ctypes.data_as(dp), f.ctypes.data_as(dp), Q.ctypes.data_as(dp), rows)
    assert relmax(Q, ref_vectors["n8_l0_Q_ff"]) < TOL_QHAT
    Qp = (dp * 1)(Q.ctypes.data_as(dp))
    L.conserveAllMoments(Qp)
    assert relmax(Q, ref_vectors["n8_l0_cons_Q_ff"]) < 1e-12
    L.ComputeQ_maxPreserve(f.ctypes.data_as(dp), f.ctypes.data_as(dp), Q.ctypes.data_as(dp), rows)
    assert relmax(Q, ref_vectors["n8_l0_Qmp_ff"]) < TOL_QHAT
    rng = np.random.default_rng(5)
    z = rng.standard_normal(n3) + 1j * rng.standard_normal(n3)
    out = np.empty(n3, dtype=np.complex128)
    L.fft3D(z.ctypes.data, out.ctypes.data, 0)
    assert relmax(out, ref_vectors["n8_l0_fft_fwd"]) < 1e-13
    L.dealloc_conservation()
    L.dealloc_coll()


def test_dropin_advect_reproduces_reference_vectors(sb, ref_vectors):
    import ctypes as C
    L = sb._lib.load()
    dp = C.POINTER(C.c_double)
    N, L_v, nX, ic, dt = 8, 9.0, 12, 3, 1e-3
    o = orc.Oracle(N, L_v, 1)
    v, eta = o.v.copy(), o.eta.copy()
    L.initialize_coll(N, L_v, v.ctypes.data_as(dp), eta.ctypes.data_as(dp))
    for order, fn in ((1, L.advectOne), (2, L.advectTwo)):
        _, x, dx = orc.make_mesh([nX // 2, nX - nX // 2], [0.4, 0.6], order)
        L.initialize_transport(N, nX, L_v, x.ctypes.data_as(dp), dx.ctypes.data_as(dp), v.ctypes.data_as(dp), ic, dt,
                               1.0, None)
        f = np.zeros((nX + 2 * order, o.n3))
        f[order:nX + order] = ref_vectors[f"tr_ic3_o{order}_in"]
        fc = np.zeros_like(f)
        cells = lambda a: (dp * a.shape[0])(*[C.cast(a.ctypes.data + i * a.shape[1] * 8, dp) for i in range(a.shape[0])])  # noqa: E731
        fn(cells(f), cells(fc), 0)
        assert relmax(fc[order:nX + order], ref_vectors[f"tr_ic3_o{order}_out"]) < 1e-14
        L.dealloc_trans()
    L.dealloc_coll()


# ---------------------------------------------------------------- device weight generator (row f1)
@pytest.mark.parametrize("fix,N,L_v,lam,rule", [("W_bkw8", 8, 5.0, 0.0, 0), ("W_heat8", 8, 9.0, 1.0, 1)])
def test_device_weight_generator_matches_golden_wts(request, sb, tmp_path, fix, N, L_v, lam, rule):
    want = request.getfixturevalue(fix)
    c = sb.Collisions(N, L_v, inhomogeneous=bool(rule))
    c.generate_weights(lam)
    got = c.weights_to_host()
    # device sin/pow are not bit-identical to glibc's, so only round-off-level agreement is asked for
    # (the CPU restatement of the same algorithm is 99.8 % byte-identical, tests/test_oracle_weights.py)
    scale = np.abs(want).max()
    assert np.median(np.abs(got - want)) <= 1e-15 * scale
    assert np.quantile(np.abs(got - want), 0.999) <= 1e-13 * scale
    assert np.abs(got - want).max() <= 1e-7 * scale          # adaptive decisions may differ within the 1e-8 tolerance
    path = str(tmp_path / sb.weights_filename(N, L_v, lam).split("/")[-1])
    c.save_weights(path)
    assert np.array_equal(np.fromfile(path), got) and os.path.getsize(path) == 8 * N ** 6
    d = sb.Collisions(N, L_v, inhomogeneous=bool(rule))
    d.load_weights(path)
    o = orc.Oracle(N, L_v, rule)
    f = seeded_f(o.v, 11)
    assert np.array_equal(c.Qhat(f), d.Qhat(f))
    assert relmax(c.ComputeQ(f), o.compute_q(got, f, f)) < TOL_QHAT


def test_device_weight_generator_n16_matches_oracle_samples(sb):
    c = sb.Collisions(16, 9.0, inhomogeneous=True)
    c.generate_weights(1.0)
    o = orc.Oracle(16, 9.0, 1)
    rng = np.random.default_rng(1)
    n3 = 16 ** 3
    row = np.empty(n3)
    worst = 0.0
    for zeta in rng.integers(0, n3, 6):
        sb._lib.check(c.L.sbte_d2h(c.h, row.ctypes.data, c.L.sbte_weights_device(c.h) + int(zeta) * n3 * 8, n3 * 8))
        xs = rng.integers(0, n3, 200)
        want = np.array([o.weight_one(1.0, int(zeta), int(x)) for x in xs])
        worst = max(worst, np.abs(row[xs] - want).max() / max(np.abs(want).max(), 1e-300))
    assert worst < 1e-7


# ---------------------------------------------------------------- BASELINE configs with real (generated) weights
def test_config_bkw16_relaxation_matches_oracle(sb):
    """BASELINE config 2 (input_examples/BKW16.in): 0D BKW at N=16, L_v=5, lambda=0, dt=0.01, RK2.
    Weights generated on the device; the SAME tensor is handed to the oracle. 8 steps (48 evaluations)."""
    N, L_v, lam, dt = 16, 5.0, 0.0, 0.01
    c = sb.Collisions(N, L_v)
    c.generate_weights(lam)
    W = c.weights_to_host()
    o = orc.Oracle(N, L_v, 0)
    fo = o.init_hom(2)
    f = c.array(o.n3).put(fo)
    rows = []
    for t in range(8):
        c.step_0d(f, dt, 1.0, 2)
        o.step_0d(W, fo, dt, 1.0, 2)
        rows.append(c.row_0d(f))
    assert relmax(f.get(), fo) < 1e-11
    np.testing.assert_allclose(rows[-1], o.row_0d(fo), rtol=1e-10, atol=1e-14)
    m = c.moments(f.get())[0]
    # BKW is an exact solution with constant density and temperature: conservation over the run
    m0 = c.moments(o.init_hom(2))[0]
    assert abs(m[0] - m0[0]) < 1e-13 and abs(m[4] - m0[4]) < 1e-12 and np.abs(m[1:4]).max() < 1e-13


def test_config_n32_hard_spheres_real_weights_vs_oracle(sb):
    """BASELINE config 3 with REAL weights: N=32, lambda=1 (1.07e9 weights generated on the device),
    one ComputeQ evaluation on the shifted-isotropic initial data against the CPU oracle reading the
    same tensor, plus the fused maxPreserve pass against the oracle's three separate passes."""
    N, L_v = 32, 5.0
    c = sb.Collisions(N, L_v)
    c.generate_weights(1.0)
    W = c.weights_to_host()                     # 8.59 GB on the host
    o = orc.Oracle(N, L_v, 0)
    f = o.init_hom(0)
    Qo, qo = o.compute_q(W, f, f, want_qhat=True)
    assert relmax(c.Qhat(f, k2=sb.K2_STREAM), qo) < TOL_QHAT
    Q = c.ComputeQ(f)
    assert relmax(Q, Qo) < TOL_QHAT
    Qc = c.conserveAllMoments(Q)
    assert np.abs(c.moment_functionals(Qc)).max() < 1e-13
    assert relmax(c.ComputeQ_maxPreserve(f), o.compute_q_maxpreserve(W, f, f)) < TOL_QHAT
    # two different distributions (the f != g entry of the mixture-shaped interface, src/collisions.c:178-210): the BKW
    # data against the shifted-isotropic one, both orders, plain ComputeQ and maxPreserve with its g - M_i quirk (:104)
    g = o.init_hom(2)
    assert relmax(c.ComputeQ(f, g), o.compute_q(W, f, g)) < TOL_QHAT
    assert relmax(c.ComputeQ_maxPreserve(f, g), o.compute_q_maxpreserve(W, f, g)) < TOL_QHAT
    assert relmax(c.ComputeQ_maxPreserve(g, f), o.compute_q_maxpreserve(W, g, f)) < TOL_QHAT


def test_n32_weight_file_round_trip_is_bit_exact(sb, tmp_path):
    """The 8.59 GB Weights/N32_isotropic_L_v5_lambda1.wts in the reference's format (headerless native doubles, row by
    row: src/weights.c:78-88 load, :100-103 store): save from one context, load into another, and the convolution on
    the loaded tensor must give the bits of the convolution on the original (plain and symmetrised stream)."""
    import shutil
    N, L_v = 32, 5.0
    if shutil.disk_usage(tmp_path).free < 10 * 2 ** 30:
        pytest.skip("less than 10 GiB free for the 8.59 GB weight file")
    a = sb.Collisions(N, L_v)
    a.generate_weights(1.0)
    path = str(tmp_path / "N32_isotropic_L_v5_lambda1.wts")
    a.save_weights(path)
    assert os.path.getsize(path) == 8 * N ** 6
    b = sb.Collisions(N, L_v)
    b.load_weights(path)
    os.remove(path)
    f = orc.Oracle(N, L_v, 0).init_hom(0)
    for sym in (True, False):
        a.set_symmetrize(sym)
        b.set_symmetrize(sym)
        assert np.array_equal(a.Qhat(f, k2=sb.K2_STREAM).view(np.float64), b.Qhat(f, k2=sb.K2_STREAM).view(np.float64)), sym
    # a few rows of the tensors themselves, bit for bit, through the device pointers
    import ctypes as C
    n3 = N ** 3
    for row in (0, n3 // 2 + 17, n3 - 1):
        ra, rb = np.empty(n3), np.empty(n3)
        for ctx, out in ((a, ra), (b, rb)):
            sb._lib.check(ctx.L.sbte_d2h(ctx.h, out.ctypes.data, C.c_void_p(ctx.L.sbte_weights_device(ctx.h) + row * n3 * 8), out.nbytes))
        assert np.array_equal(ra, rb), row


@pytest.mark.parametrize("N", [22, 24])
def test_config_heattrans_full_step_matches_oracle(sb, N):
    """BASELINE config 5 (input_examples/heatTrans.long.in: N = 22 as shipped, 24 as BASELINE states): L_v=9, Kn=0.3,
    lambda=1, Init_field 3 (diffuse walls), Space_order 1, dt=1e-4, dx=1/250 -- a 40-cell slab between the walls, 3 full
    time steps (exec/boltz.c:264-353) with device-generated weights against the oracle reading the same tensor."""
    L_v, Kn, order, ic, dt, nX = 9.0, 0.3, 1, 3, 1e-4, 40
    c = sb.Collisions(N, L_v, inhomogeneous=True)
    c.generate_weights(1.0)
    W = c.weights_to_host()
    o = orc.Oracle(N, L_v, 1)
    _, x, dx = orc.make_mesh([nX], [nX / 250.0], order)
    f = o.init_inhom(ic, nX, order)
    s = sb.Slab(c, nX, order, x, dx, ic, dt)
    s.upload(f)
    fc, f1, ft = np.zeros_like(f), np.zeros_like(f), np.zeros_like(f)
    for _ in range(3):
        s.step(Kn)
        o.step_1d(W, nX, x, dx, dt, Kn, order, ic, f, fc, f1, ft)
    got = s.download()[order:nX + order]
    assert relmax(got, f[order:nX + order]) < 1e-11
    mom = s.moments()
    for l in (0, 1, nX // 2, nX - 1):
        np.testing.assert_allclose(mom[l, [0, 1, 4, 7]], o.row_1d(f[l + order]), rtol=1e-10, atol=1e-13)


def test_config_shock1p2_derived_matches_oracle(sb):
    """BASELINE config 4 (derived): N=16, L_v=9, Kn=1.52, lambda=1, Init_field 6, Space_order 2,
    dt=1e-3, dx=6/640 -- a 40-cell window around the shock, 3 steps, against the oracle."""
    N, L_v, Kn, order, ic, dt, nX = 16, 9.0, 1.52, 2, 6, 1e-3, 40
    c = sb.Collisions(N, L_v, inhomogeneous=True)
    c.generate_weights(1.0)
    W = c.weights_to_host()
    o = orc.Oracle(N, L_v, 1)
    _, x, dx = orc.make_mesh([nX], [6.0 * nX / 640.0], order)
    f = o.init_inhom(ic, nX, order)
    s = sb.Slab(c, nX, order, x, dx, ic, dt)
    s.upload(f)
    fc, f1, ft = np.zeros_like(f), np.zeros_like(f), np.zeros_like(f)
    for _ in range(3):
        s.step(Kn)
        o.step_1d(W, nX, x, dx, dt, Kn, order, ic, f, fc, f1, ft)
    got = s.download()[order:nX + order]
    assert relmax(got, f[order:nX + order]) < 1e-11
    mom = s.moments()
    for l in (0, nX // 2 - 1, nX // 2, nX - 1):
        np.testing.assert_allclose(mom[l, [0, 1, 4, 7]], o.row_1d(f[l + order]), rtol=1e-10, atol=1e-13)


# ---------------------------------------------------------------- error behaviour of the extension surface
def test_error_paths_report_like_the_reference(sb, tmp_path):
    from spectralbte_b200._lib import SbteError
    c = sb.Collisions(8, 5.0)
    with pytest.raises(SbteError, match="no weights bound"):
        c.ComputeQ(np.zeros(512))
    with pytest.raises(SbteError, match="cannot open weight file"):
        c.load_weights(str(tmp_path / "missing.wts"))
    bad = tmp_path / "short.wts"
    bad.write_bytes(b"\0" * 1000)                       # truncated file: src/weights.c:82-86
    with pytest.raises(SbteError, match="Error reading weight file"):
        c.load_weights(str(bad))
    with pytest.raises(SbteError, match="N must be even"):
        sb.Collisions(7, 5.0)
    c.synthetic_weights(1)
    with pytest.raises(SbteError, match="requires f == g"):
        c.ComputeQ(np.ones((40, 512)), np.ones((40, 512)), k2=sb.K2_BATCH)
    with pytest.raises(SbteError, match="Space_order must be 1 or 2"):
        sb.Slab(c, 8, 3, np.zeros(14), np.ones(14), 3, 1e-3)
    with pytest.raises(SbteError, match="too few cells"):
        sb.Slab(c, 3, 2, np.zeros(7), np.ones(7), 3, 1e-3)


# ---------------------------------------------------------------- symmetrised weight stream (f == g)
@pytest.mark.parametrize("N,k2,cells", [(16, 2, 1), (24, 2, 1), (16, 4, 1), (8, 3, 37), (16, 3, 40), (24, 3, 33), (22, 3, 33)])
def test_symmetrised_stream_equals_plain_stream(sb, N, k2, cells):
    """Ws = W + W o sigma over the representative xi_x planes must reproduce the full sum for f == g."""
    o = orc.Oracle(N, 7.0, 1)
    c = sb.Collisions(N, 7.0, inhomogeneous=True)
    c.synthetic_weights(7)
    f = np.stack([seeded_f(o.v, 50 + b, noise=0.3) for b in range(cells)])
    c.set_symmetrize(True)
    a = c.Qhat(f, k2=k2)
    qa = c.ComputeQ(f, k2=k2)
    c.set_symmetrize(False)
    b = c.Qhat(f, k2=k2)
    qb = c.ComputeQ(f, k2=k2)
    assert relmax(a, b) < 1e-13 and relmax(qa, qb) < 1e-13
    assert relmax(b.reshape(cells, -1)[0], c.Qhat(f[0], k2=sb.K2_GENERIC)) < TOL_QHAT
    if cells == 1:
        c.set_symmetrize(True)
        ma = c.ComputeQ_maxPreserve(f[0], k2=k2)
        c.set_symmetrize(False)
        mb = c.ComputeQ_maxPreserve(f[0], k2=k2)
        assert relmax(ma, mb) < 1e-13


@pytest.mark.parametrize("N,L_v,lam", [(16, 5.0, 0.0), (16, 9.0, 1.0), (32, 5.0, 1.0)])
def test_transposed_pairing_of_the_0d_stream(sb, N, L_v, lam):
    """Isotropic weights are invariant under x <-> y of both indices, so the 0D stream reads only the zeta columns
    zx >= zy and forms column (zy, zx) from the same weights against the transposed spectrum (csrc/qhat.cu, TP).  The
    library must (i) find the generated tensor invariant, (ii) reproduce the full stream's Q^ far inside the 1e-12
    tolerance, plain and symmetrised, (iii) refuse a tensor that is not invariant (synthetic weights) and give the very
    bits of the full stream there."""
    rule = 0
    c = sb.Collisions(N, L_v)
    c.generate_weights(lam)
    state, dev = c.xy_pairing_state()
    assert state == 1 and dev <= 1e-14, (state, dev)
    o = orc.Oracle(N, L_v, rule)
    f = seeded_f(o.v, 77, noise=0.2)
    for sym in (True, False):
        c.set_symmetrize(sym)
        c.set_xy_pairing(True)
        qa, Qa, Ma = c.Qhat(f, k2=sb.K2_STREAM), c.ComputeQ(f, k2=sb.K2_STREAM), c.ComputeQ_maxPreserve(f)
        c.set_xy_pairing(False)
        qb, Qb, Mb = c.Qhat(f, k2=sb.K2_STREAM), c.ComputeQ(f, k2=sb.K2_STREAM), c.ComputeQ_maxPreserve(f)
        assert not np.array_equal(qa, qb) and not np.array_equal(Ma, Mb)   # different kernels really ran
        assert relmax(qa, qb) < 1e-13 and relmax(Qa, Qb) < 1e-13, sym
        # ComputeQ_maxPreserve: its two summed products in both orientations (four operand pairs, half-chunk staging at N = 32)
        assert relmax(Ma, Mb) < 1e-13, sym
    c.set_xy_pairing(True)
    if N == 16:
        W = c.weights_to_host()
        assert relmax(qa, o.qhat(W, o.fft3d(f.astype(complex)), o.fft3d(f.astype(complex)))) < TOL_QHAT
        assert relmax(Ma, o.compute_q_maxpreserve(W, f, f)) < TOL_QHAT
    # a tensor without the invariance keeps the full stream
    c.synthetic_weights(5)
    state, dev = c.xy_pairing_state()
    assert state == 0 and dev > 1e-3
    qa = c.Qhat(f, k2=sb.K2_STREAM)
    c.set_xy_pairing(False)
    assert np.array_equal(qa, c.Qhat(f, k2=sb.K2_STREAM))


def test_symmetrised_representatives_cover_every_pair_once():
    """Host-side check of the xi_x representative rule used by the kernels (common.cuh sym_nrep/sym_rep)."""
    for N in (8, 16, 22, 24, 32):
        for zx in range(N):
            a = (zx + N // 2) % N
            h = a // 2
            nrep = h + 1 + (a + N) // 2 - a
            reps = [c if c <= h else c + (a - h) for c in range(nrep)]
            want = [ex for ex in range(N) if ex <= (a - ex) % N]
            assert reps == want
            seen = set()
            for ex in reps:
                seen.add(ex); seen.add((a - ex) % N)
            assert seen == set(range(N))
