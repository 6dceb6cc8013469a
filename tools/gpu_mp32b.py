"""maxPreserve at N=32: device-pointer API, back-to-back vs synchronised, host clock."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
from spectralbte_b200 import initial
c = sb.Collisions(32, 5.0)
c.synthetic_weights(20261017)
f = initial.init_hom(c.v, 5.0, 0)
d = c.array(c.n3).put(f); q = c.array(c.n3)
call = lambda: sb._lib.check(c.L.sbte_compute_q_maxpreserve(c.h, d.ptr, d.ptr, q.ptr, 0))
for _ in range(5): call()
c.sync()
for mode in ("back-to-back", "sync each", "back-to-back"):
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        call()
        if mode == "sync each": c.sync()
    c.sync()
    print(mode, "ms/call", (time.perf_counter() - t0) / n * 1e3)
c.k2_profile(True)
for _ in range(20): call()
c.sync()
ms, cnt = c.k2_profile_read()
print("K2 event-timed ms", ms / cnt, cnt)
