"""One-off CPU check of the N = 32 instantiation of the 0D half-spectrum kernels (csrc/qhat_half.cu) on the host emulation
(tests/emul): about 18 GB of tensors and ten minutes, so it is not part of the test-suite.
    python tools/emul_half0d_n32.py [npairs]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import relmax, seeded_f  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import test_kernel_emulation_cpu as tk  # noqa: E402
import test_mirror_emulation_cpu as tm  # noqa: E402

N, nsplit, packed = 32, 2, 1
npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if npairs == 2:
    nsplit = 1
dp = C.POINTER(C.c_double)
L = tk._lib()
L.emul_half0d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
R = tm._emul()
o = orc.Oracle(N, 5.0, 0)
n3 = N ** 3
t0 = time.time()
W = orc.synthetic_weights(N)
print("weights %.0f s" % (time.time() - t0), flush=True)
Wh = np.empty_like(W)
assert R.mirror_emul_fold(N, W.ctypes.data_as(dp), 1, Wh.ctypes.data_as(dp)) == 0
print("folded %.0f s" % (time.time() - t0), flush=True)
f = seeded_f(o.v, 41, noise=0.3)
z = np.arange(N)


def parity(x):
    F = o.fft3d(np.asarray(x).astype(complex)).reshape(N, N, N)
    out = np.empty((N, N, N), dtype=complex)
    out[:, :, (z & 1) * (N // 2) + (z >> 1)] = F
    return np.ascontiguousarray(out)


if npairs == 1:
    xiA = dfA = xiB = dfB = parity(f)
    want = o.compute_q(W, f, f)
else:
    M, _ = o.find_maxwellian(f)
    g = f - M
    xiA, dfA, xiB, dfB = parity(g), parity(f), parity(M), parity(g)
    want = o.compute_q_maxpreserve(W, f, f)
print("oracle %.0f s" % (time.time() - t0), flush=True)
parts = np.full((nsplit + 1) * n3, np.nan + 1j * np.nan, dtype=complex)
pd = lambda a: a.view(np.float64).ctypes.data_as(dp)  # noqa: E731
rc = L.emul_half0d(N, nsplit, packed, npairs, Wh.ctypes.data_as(dp), pd(xiA), pd(dfA), pd(xiB), pd(dfB), pd(parts))
print("emulation rc %d, %.0f s" % (rc, time.time() - t0), flush=True)
parts = parts.reshape(nsplit + 1, n3)
assert rc == 0 and not np.isnan(parts.view(np.float64)).any()
err = relmax(np.real(o.fft3d(parts.sum(axis=0), invert=True)), want)
print("N=32 npairs=%d: Q rel err %.3e" % (npairs, err))
assert err < 1e-12
