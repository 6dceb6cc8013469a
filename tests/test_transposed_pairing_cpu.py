"""Host-side check of the identity the transposed pairing of the 0D stream relies on, with the
oracle's own convolution and the REFERENCE'S golden weight files as the witnesses:

  (i)  isotropic weights are invariant under swapping the x and y axes of both indices,
           W[(zy,zx,zz)][(ey,ex,ez)] == W[(zx,zy,zz)][(ex,ey,ez)]        (src/weights.c:265-281),
       exactly -- in the reference's golden .wts files and in the oracle's restatement of the generator;
  (ii) hence, for f == g, the rows of zeta column (zy, zx) are the rows of column (zx, zy) read against the transposed
       spectrum F(ex,ey,ez) = f^(ey,ex,ez):
           Q^(zy,zx,zz) = sum_xi W[(zx,zy,zz)][xi] F(xi) F(sigma(xi)),
       which is what csrc/qhat.cu (TP) computes from the columns zx >= zy alone."""
import numpy as np
import pytest

from conftest import relmax, seeded_f
from oracle import oracle as orc


def transpose_xy(a, N):
    return np.ascontiguousarray(a.reshape(N, N, N).transpose(1, 0, 2)).reshape(-1)


def transpose_rows_and_columns(W, N):
    n3 = N ** 3
    perm = transpose_xy(np.arange(n3), N)      # perm[(x,y,z)] = flat index of (y,x,z)
    return W.reshape(n3, n3)[perm][:, perm]


@pytest.mark.parametrize("wname", ["W_bkw8", "W_heat8"])
def test_reference_golden_weights_are_xy_invariant(request, wname):
    W = request.getfixturevalue(wname)
    N = 8
    assert np.array_equal(transpose_rows_and_columns(W, N), W.reshape(N ** 3, N ** 3))   # bit for bit


@pytest.mark.parametrize("N,lam", [(6, 1.0)])
def test_oracle_generator_weights_are_xy_invariant(N, lam):
    o = orc.Oracle(N, 5.0, 0)
    W = o.weights_iso(lam)
    assert np.array_equal(transpose_rows_and_columns(W, N), W.reshape(N ** 3, N ** 3))


def test_transposed_rows_give_the_transposed_columns(W_bkw8):
    N = 8
    o = orc.Oracle(N, 5.0, 0)
    fh = o.fft3d(seeded_f(o.v, 5, noise=0.3).astype(complex))
    full = o.qhat(W_bkw8, fh, fh).reshape(N, N, N)                       # src/collisions.c:127-165, every row
    Fh = transpose_xy(fh, N)
    paired = o.qhat(W_bkw8, Fh, Fh).reshape(N, N, N)                     # the same rows against the transposed spectrum
    # row (zx, zy, zz) of `paired` is row (zy, zx, zz) of the full result
    assert relmax(paired.transpose(1, 0, 2), full) < 1e-13
    # so the columns zx >= zy determine everything
    rebuilt = np.empty_like(full)
    for zx in range(N):
        for zy in range(zx + 1):
            rebuilt[zx, zy] = full[zx, zy]
            rebuilt[zy, zx] = paired[zx, zy]
    assert relmax(rebuilt, full) < 1e-13


def test_a_generic_tensor_is_not_invariant():
    N = 6
    W = orc.synthetic_weights(N)
    d = np.abs(transpose_rows_and_columns(W, N) - W.reshape(N ** 3, N ** 3)).max()
    assert d > 1e-3 * np.abs(W).max()      # the library's check (1e-14 of the largest entry) refuses such a tensor
