"""Batched ComputeQ timing: `gpu_batch_time.py N [cells]` times one configuration under the current environment
(SBTE_NO_BATCH3G, SBTE_BATCH_CTAS, ...); without arguments, heatTrans-sized batches (250 cells) at N = 20, 22, 24:
line-ring kernel with partial row-blocks (default) against the any-N kernel (SBTE_NO_BATCH3G=1).
numpy + ctypes only (no torch import), so it starts in a second on a fresh box."""
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())


def run(N, cells=250, reps=10):
    import spectralbte_b200 as sb
    from spectralbte_b200._lib import check
    c = sb.Collisions(N, 9.0, inhomogeneous=True)
    c.synthetic_weights(11)
    rng = np.random.default_rng(N)
    f = np.abs(rng.standard_normal(cells * c.n3)) * 1e-2
    d = c.array(f.size).put(f)
    q = c.array(f.size)
    call = lambda: check(c.L.sbte_compute_q(c.h, d.ptr, d.ptr, q.ptr, cells, sb.K2_BATCH))
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
    nrep = sum((lambda a: a // 2 + 1 + (a + N) // 2 - a)((zx + N // 2) % N) for zx in range(N))
    flops = 10.0 * N ** 4 * nrep * cells
    print("N=%d cells=%d %s: %.3f ms/ComputeQ, convolution %.3f ms (%d launches) = %.2f TFLOP/s executed, checksum %.17g"
          % (N, cells, "any-N kernel" if os.environ.get("SBTE_NO_BATCH3G") else "batched kernel", wall, k2ms / max(n, 1), n,
             flops / (k2ms / max(n, 1)) * 1e-9, float(np.abs(Q).sum())), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(int(sys.argv[1]), *(int(a) for a in sys.argv[2:3]))
    else:
        for N in (22, 20, 24):
            for env in ({}, {"SBTE_NO_BATCH3G": "1"}):
                if N == 24 and env:
                    continue
                e = dict(os.environ); e.update(env)
                try:
                    subprocess.run([sys.executable, __file__, str(N)], env=e, timeout=60)
                except subprocess.TimeoutExpired:
                    print("N=%d %s: timeout" % (N, env), flush=True)
