"""Hot-loop timing of the single-cell transforms (back-to-back launches, device time per launch)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
for N in (32, 24, 16):
    c = sb.Collisions(N, 5.0)
    a = c.array(2 * c.n3).put(np.random.default_rng(0).standard_normal(2 * c.n3))
    b = c.array(2 * c.n3)
    for inv in (0, 1):
        for _ in range(10): sb._lib.check(c.L.sbte_fft3d(c.h, a.ptr, b.ptr, inv, 1))
        c.sync(); t0 = time.perf_counter()
        n = 2000
        for _ in range(n): sb._lib.check(c.L.sbte_fft3d(c.h, a.ptr, b.ptr, inv, 1))
        c.sync(); print("N", N, "invert", inv, "us per transform", (time.perf_counter() - t0) / n * 1e6)
