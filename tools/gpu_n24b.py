import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
from spectralbte_b200 import initial
for N in (24, 16):
    c = sb.Collisions(N, 5.0)
    c.synthetic_weights(1)
    f = initial.init_hom(c.v, 5.0, 0)
    d = c.array(c.n3).put(f); q = c.array(c.n3)
    for _ in range(5): sb._lib.check(c.L.sbte_compute_q(c.h, d.ptr, d.ptr, q.ptr, 1, 0))
    c.sync(); t0 = time.perf_counter()
    for _ in range(200): sb._lib.check(c.L.sbte_compute_q(c.h, d.ptr, d.ptr, q.ptr, 1, 0))
    c.sync(); print("N", N, "ComputeQ device ms", (time.perf_counter() - t0) / 200 * 1e3)
