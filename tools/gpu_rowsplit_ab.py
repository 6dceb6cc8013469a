"""Row-split batched convolution (SBTE_ROW_SPLIT=1: two warps per zeta column, four compute warps per SM
sub-partition) against the default (one warp per column) at N=16: convolution time and bitwise comparison
of Q.  numpy + ctypes only (no torch import)."""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())


def run(N, cells, out, reps=20):
    import spectralbte_b200 as sb
    from spectralbte_b200._lib import check
    c = sb.Collisions(N, 9.0, inhomogeneous=True)
    c.synthetic_weights(11)
    rng = np.random.default_rng(N)
    f = np.abs(rng.standard_normal(cells * c.n3)) * 1e-2
    d = c.array(f.size).put(f)
    q = c.array(f.size)
    call = lambda: check(c.L.sbte_compute_q(c.h, d.ptr, d.ptr, q.ptr, cells, sb.K2_BATCH))  # noqa: E731
    for _ in range(3):
        call()
    c.sync()
    c.k2_profile(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    c.sync()
    wall = (time.perf_counter() - t0) / reps * 1e3
    k2ms, n = c.k2_profile_read()
    c.k2_profile(False)
    Q = q.get()
    if out:
        np.save(out, Q)
    nrep = sum((lambda a: a // 2 + 1 + (a + N) // 2 - a)((zx + N // 2) % N) for zx in range(N))
    flops = 10.0 * N ** 4 * nrep * cells
    print("N=%d cells=%d row_split=%s: %.3f ms/ComputeQ, convolution %.4f ms = %.2f TFLOP/s executed (%.1f%% of 36.5)"
          % (N, cells, os.environ.get("SBTE_ROW_SPLIT", "default"), wall, k2ms / n, flops / (k2ms / n) * 1e-9,
             flops / (k2ms / n) * 1e-9 / 36.5 * 100), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        for N, cells in ((16, 640), (16, 80)):
            outs = []
            for split in ("0", "1"):
                e = dict(os.environ)
                e["SBTE_ROW_SPLIT"] = split
                o = "/tmp/rowsplit_%d_%d_%s.npy" % (N, cells, split)
                outs.append(o)
                try:
                    subprocess.run([sys.executable, __file__, str(N), str(cells), o], env=e, timeout=12)
                except subprocess.TimeoutExpired:
                    print("N=%d cells=%d split=%s: timeout" % (N, cells, split), flush=True)
            try:
                a, b = np.load(outs[0]), np.load(outs[1])
                print("N=%d cells=%d: bitwise equal %s, max |diff| %.3e, max |Q| %.3e"
                      % (N, cells, bool(np.array_equal(a, b)), float(np.abs(a - b).max()), float(np.abs(a).max())), flush=True)
            except Exception as ex:  # noqa: BLE001
                print("compare failed:", ex, flush=True)
