import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import spectralbte_b200 as sb
from spectralbte_b200 import initial
# heatTrans-like slab at N=24: 256 cells, order 1, time the collide step
N, L_v, Kn, order, ic, dt, nX = 24, 9.0, 0.3, 1, 3, 1e-4, 256
c = sb.Collisions(N, L_v, inhomogeneous=True)
t0 = time.perf_counter(); c.generate_weights(1.0); print("weights N=24 generated in %.1f s" % (time.perf_counter() - t0), flush=True)
_, x, dx = initial.make_mesh([nX], [1.0], order)
s = sb.Slab(c, nX, order, x, dx, ic, dt)
s.upload(initial.init_inhom(c.v, ic, nX, order, 0, nX))
for _ in range(2): s.step(Kn)
c.sync(); c.k2_profile(True)
t0 = time.perf_counter()
for _ in range(5): s.step(Kn)
c.sync(); dtm = (time.perf_counter() - t0) / 5
ms, n = c.k2_profile_read()
fl = 10.0 * N ** 6 * nX
print("N=24 %d cells: step %.2f ms, K2 %.2f ms/launch -> %.2f TFLOP/s (%.1f%% of 37), %.0f cells*steps/s" % (nX, dtm * 1e3, ms / n, fl / (ms / n * 1e-3) / 1e12, fl / (ms / n * 1e-3) / 1e12 / 37 * 100, nX / dtm), flush=True)
mom = s.moments(); print("rho range", mom[:, 0].min(), mom[:, 0].max(), "T range", mom[:, 4].min(), mom[:, 4].max())
