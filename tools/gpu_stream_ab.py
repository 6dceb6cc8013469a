"""Same-box A/B of the 0D N=32 stream kernels (pairing / symmetrised / plain; ComputeQ and maxPreserve) between the current
library and another build given by SBTE_LIB_PATH (run twice, once with and once without the variable)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
from spectralbte_b200 import initial
from spectralbte_b200._lib import check
c = sb.Collisions(32, 5.0)
c.generate_weights(1.0)
f = initial.init_hom(c.v, 5.0, 0)
d, q = c.array(f.size).put(f), c.array(f.size)
tag = "other build" if os.environ.get("SBTE_LIB_PATH") else "current"
def run(name, fn, reps=300):
    for _ in range(5):
        fn()
    c.sync()
    c.k2_profile(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    c.sync()
    wall = (time.perf_counter() - t0) / reps * 1e3
    k2, n = c.k2_profile_read()
    c.k2_profile(False)
    print("%-12s %-22s %.4f ms/call, kernel %.4f ms, checksum %.17g" % (tag, name, wall, k2 / n, float(np.abs(q.get()).sum())), flush=True)
cq = lambda: check(c.L.sbte_compute_q(c.h, d.ptr, d.ptr, q.ptr, 1, sb.K2_AUTO))
mp = lambda: check(c.L.sbte_compute_q_maxpreserve(c.h, d.ptr, d.ptr, q.ptr, sb.K2_AUTO))
has_xy = hasattr(c.L, "sbte_set_xy_pairing")
run("ComputeQ default", cq)
if has_xy:
    c.set_xy_pairing(False)
run("ComputeQ sym", cq)
c.set_symmetrize(False)
run("ComputeQ plain", cq, 150)
c.set_symmetrize(True)
run("maxPreserve sym", mp)
if has_xy:
    c.set_xy_pairing(True)
    run("maxPreserve default", mp)
