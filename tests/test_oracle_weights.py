"""Pins oracle/qag21.c + the weight formula against the reference's byte-exact golden .wts files
(tests/run_test.sh:41-52 diffs the generated file against target/)."""
import numpy as np
import pytest

from oracle import oracle as orc


@pytest.mark.parametrize("fix,N,L_v,lam,rule", [("W_bkw8", 8, 5.0, 0.0, 0), ("W_heat8", 8, 9.0, 1.0, 1)])
def test_generated_weights_match_golden(request, fix, N, L_v, lam, rule):
    want = request.getfixturevalue(fix)
    o = orc.Oracle(N, L_v, rule)
    got = o.weights_iso(lam)
    assert got.shape == want.shape == (N ** 6,)
    exact = (got == want).mean()
    # GSL itself is absent: bit-equality of every entry depends on libm's pow/sin; demand >= 99.5 %
    # identical bytes and the rest within 2e-14 of the largest weight
    assert exact > 0.995, exact
    assert np.abs(got - want).max() <= 2e-14 * np.abs(want).max()
    # the row whose zeta has eta = 0 in every dimension (index N/2) has an identically zero
    # integrand: sinc(0)*sinc(r|xi|) - sinc(r|xi|)  (src/weights.c:159,192-194)
    n3 = N ** 3
    z0 = (N // 2) * (1 + N + N * N)
    assert np.all(want.reshape(n3, n3)[z0] == 0.0) and np.all(got.reshape(n3, n3)[z0] == 0.0)


def test_wts_file_layout(W_bkw8):
    """Headerless raw doubles, exactly 8*N^6 bytes (src/weights.c:81-87,101-103)."""
    assert W_bkw8.nbytes == 8 * 8 ** 6 == 2097152
