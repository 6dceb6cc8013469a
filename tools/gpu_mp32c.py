"""Does the two-pair kernel slow down under sustained load? Batches of 50 back-to-back calls with clocks/power sampled between."""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
from spectralbte_b200 import initial
c = sb.Collisions(32, 5.0)
c.synthetic_weights(20261017)
f = initial.init_hom(c.v, 5.0, 0)
d = c.array(c.n3).put(f); q = c.array(c.n3)
mp = lambda: sb._lib.check(c.L.sbte_compute_q_maxpreserve(c.h, d.ptr, d.ptr, q.ptr, 0))
cq = lambda: sb._lib.check(c.L.sbte_compute_q(c.h, d.ptr, d.ptr, q.ptr, 1, 0))
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for name, call in (("maxPreserve", mp), ("ComputeQ", cq), ("maxPreserve", mp)):
    for b in range(8):
        t0 = time.perf_counter()
        for _ in range(50): call()
        c.sync()
        dt = (time.perf_counter() - t0) / 50 * 1e3
        print(name, "batch", b, "ms/call %.4f" % dt, "|", smi() if b % 2 == 1 else "")
