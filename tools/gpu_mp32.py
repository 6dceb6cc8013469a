"""ComputeQ_maxPreserve at N=32 (synthetic weights): a few calls, timed; used under ncu to profile the NP=2 stream kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import spectralbte_b200 as sb
from spectralbte_b200 import initial
c = sb.Collisions(32, 5.0)
c.synthetic_weights(20261017)
f = initial.init_hom(c.v, 5.0, 0)
for _ in range(3):
    Q = c.ComputeQ_maxPreserve(f)
c.sync()
n = int(os.environ.get("REPS", "20"))
t0 = time.perf_counter()
for _ in range(n):
    Q = c.ComputeQ_maxPreserve(f)
c.sync()
print("maxPreserve ms/call (host API)", (time.perf_counter() - t0) / n * 1e3, "checksum", float(np.abs(Q).sum()))
